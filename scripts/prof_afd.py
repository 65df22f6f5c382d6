"""Host-entry throughput with an AFD requested (the reference always writes the AFD field)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from varlociraptor_b200 import synth, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sc, b = synth.tumor_normal(n, seed=11)
eng = engine.PosteriorEngine(sc.flatten())
pb = engine.pin_batch(b)
out = engine.pinned_results(n, 2, sc.flatten().n_events, cap)
for rep in range(3):
    t = time.perf_counter(); eng.call_batch(pb, afd_capacity=cap, out=out); dt = time.perf_counter() - t
    print("afd %d: %.1f ms -> %.0f loci/s, %d launches" % (cap, dt * 1e3, n / dt, eng.launches))
