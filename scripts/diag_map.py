"""Which loci of the config-5 parity sample report another MAP allele frequency than the oracle, and by how much?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import engine, synth
from oracle import oracle
sc, b = synth.tumor_normal(2000, seed=synth.SEED_BASE + 5, depth_range=(10, 2000))
flat = sc.flatten()
o = oracle.call_batch(flat, b, afd_capacity=0, n_threads=os.cpu_count() or 1)
g = engine.PosteriorEngine(flat).call_batch(b)
bad = np.nonzero(np.any(o.map_vaf != g.map_vaf, axis=1) & ~o.knife_edge())[0]
S = 2
for i in bad:
    d = b.read_offsets[i * S + 1] - b.read_offsets[i * S], b.read_offsets[i * S + 2] - b.read_offsets[i * S + 1]
    print("locus", i, "depths", d, "oracle MAP", [float.hex(float(x)) for x in o.map_vaf[i]], "engine MAP", [float.hex(float(x)) for x in g.map_vaf[i]],
          "best", o.best_event[i], g.best_event[i], "max |d ln post|", float(np.nanmax(np.abs(o.log_posteriors[i] - g.log_posteriors[i]))),
          "evals", o.n_base_events[i], g.n_base_events[i])
print(len(bad), "loci differ")
