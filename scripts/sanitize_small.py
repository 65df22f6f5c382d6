"""Tiny run for compute-sanitizer: TN + pedigree + AFD on a few hundred loci through the host entry."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import synth, engine
for name, (sc, b) in {"tn": synth.tumor_normal(192, seed=3), "ped": synth.pedigree(192, seed=4),
                      "skew": synth.tumor_normal(48, seed=5, depth_range=(10, 1200))}.items():
    eng = engine.PosteriorEngine(sc.flatten())
    out = eng.call_batch(b, afd_capacity=64)
    print(name, "ok", int((out.status & 0x83f).sum()))
    eng.close()
