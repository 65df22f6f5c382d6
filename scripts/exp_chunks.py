"""Experiment: chunk size of the host entry (VLR_CHUNK_LOCI: loci per chunk; reads per chunk = max(16 M, 256 x that))
against the device entry on one batch.   python scripts/exp_chunks.py N CFG chunk_loci ...   (0 = the built-in rule)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 5
chunks = [int(x) for x in sys.argv[3:]] or [0]
sc, b = synth.config(cfg, n, seed=synth.SEED_BASE + cfg)
flat = sc.flatten(); S = flat.n_samples
eng = engine.PosteriorEngine(flat)
pb = engine.pin_batch(b); pres = engine.pinned_results(n, S, flat.n_events)
db = engine.DeviceBatch(b); dr = engine.DeviceResults(n, S, flat.n_events); s = torch.cuda.Stream()
eng.reserve(int(np.max(b.read_offsets[S::S] - b.read_offsets[:-S:S])))
def t_dev():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); eng.call_batch_device(db, dr, s.cuda_stream); e1.record(s); torch.cuda.synchronize(); return e0.elapsed_time(e1)
def t_host():
    torch.cuda.synchronize(); t0 = time.perf_counter(); eng.call_batch(pb, out=pres); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
t_dev(); ts = [t_dev() for _ in range(3)]
print("cfg %d, %d loci, device entry: %.2f ms = %.3f M loci/s" % (cfg, n, min(ts), n / min(ts) / 1e3), flush=True)
ref = dr.log_posteriors.cpu().numpy()
for c in chunks:
    if c: os.environ["VLR_CHUNK_LOCI"] = str(c)
    else: os.environ.pop("VLR_CHUNK_LOCI", None)
    t_host(); ts = [t_host() for _ in range(3)]
    print("host f32, chunk knob %7d: %.2f ms = %.3f M loci/s, %d launches, bitwise equal to the device entry: %s"
          % (c, min(ts), n / min(ts) / 1e3, eng.launches, np.array_equal(pres.log_posteriors, ref, equal_nan=True)), flush=True)
