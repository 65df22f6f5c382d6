"""Experiment: throughput against the CTAs of the class-1 resident kernel per SM (VLR_RES_CTAS), for one sub-chunk on
one stream (65 536 loci) and for three sub-chunks on three streams (196 608 loci)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import engine, synth
for n in (65536, 196608):
    sc, b = synth.config(2, n, seed=synth.SEED_BASE + 2)
    flat = sc.flatten(); S = flat.n_samples
    db = engine.DeviceBatch(b); dr = engine.DeviceResults(n, S, flat.n_events); s = torch.cuda.Stream()
    for ctas in (2, 3, 4, 5, 6):
        os.environ["VLR_RES_CTAS"] = str(ctas)
        eng = engine.PosteriorEngine(flat)
        best = 1e30
        for i in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); eng.call_batch_device(db, dr, s.cuda_stream); e1.record(s); torch.cuda.synchronize()
            if i: best = min(best, e0.elapsed_time(e1))
        print("%6d loci, %d CTAs (%2d warps) per SM: %.3f ms = %.3f M loci/s" % (n, ctas, 2 * ctas, best, n / best / 1e3), flush=True)
        eng.close()
    del db, dr
