"""One sub-chunk of the wavefront pipeline for ncu / timing: N loci of config 2, device-resident, `reps` calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from varlociraptor_b200 import synth, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sc, b = synth.config(cfg, n, seed=synth.SEED_BASE + cfg)
flat = sc.flatten()
eng = engine.PosteriorEngine(flat)
import numpy as np
S = flat.n_samples
eng.reserve(int(np.max(b.read_offsets[S::S] - b.read_offsets[:-S:S])))
db = engine.DeviceBatch(b)
dr = engine.DeviceResults(b.n_loci, S, flat.n_events)
s = torch.cuda.Stream()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    eng.call_batch_device(db, dr, s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize()
    print("call ms %.3f  -> %.0f loci/s, %d launches" % (e0.elapsed_time(e1), n / e0.elapsed_time(e1) * 1e3, eng.launches))
