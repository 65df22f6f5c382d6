"""Small single-launch run for ncu: one vlr_call_kernel over N loci of config 2 (device-resident)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from varlociraptor_b200 import synth, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc, b = synth.tumor_normal(n, seed=synth.SEED_BASE + 2)
flat = sc.flatten()
eng = engine.PosteriorEngine(flat)
db = engine.DeviceBatch(b)
dr = engine.DeviceResults(b.n_loci, 2, flat.n_events)
s = torch.cuda.Stream()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    eng.call_batch_device(db, dr, s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize()
    print("kernel ms %.3f  -> %.0f loci/s" % (e0.elapsed_time(e1), n / e0.elapsed_time(e1) * 1e3))
