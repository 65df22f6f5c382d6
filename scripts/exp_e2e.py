"""Experiment: the host entries (plain f32 columns, packed columns) against the device entry on the same config-2 batch."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
sc, b = synth.config(2, n, seed=synth.SEED_BASE + 2)
flat = sc.flatten(); S = flat.n_samples
eng = engine.PosteriorEngine(flat)
pb = engine.pin_batch(b); pk = engine.PackedBatch(pb); pres = engine.pinned_results(n, S, flat.n_events)
db = engine.DeviceBatch(b); dr = engine.DeviceResults(n, S, flat.n_events); s = torch.cuda.Stream()
def t_dev():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); eng.call_batch_device(db, dr, s.cuda_stream); e1.record(s); torch.cuda.synchronize(); return e0.elapsed_time(e1)
def t_host(fn, arg):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(arg, out=pres); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
for name, f in (("device", t_dev), ("host f32", lambda: t_host(eng.call_batch, pb)), ("host packed", lambda: t_host(eng.call_batch_packed, pk)),
                ("device", t_dev), ("host packed", lambda: t_host(eng.call_batch_packed, pk)), ("host f32", lambda: t_host(eng.call_batch, pb))):
    f(); ts = [f() for _ in range(3)]
    print("%-12s best %.2f ms = %.3f M loci/s (chunk loci %s)" % (name, min(ts), n / min(ts) / 1e3, os.environ.get("VLR_CHUNK_LOCI", "65536")), flush=True)
