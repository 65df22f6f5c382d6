import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
sc, b = synth.config(2, n, seed=synth.SEED_BASE + 2)
flat = sc.flatten(); S = flat.n_samples
eng = engine.PosteriorEngine(flat)
pb = engine.pin_batch(b); pk = engine.PackedBatch(pb); pres = engine.pinned_results(n, S, flat.n_events)
eng.call_batch_packed(pk, out=pres); eng.call_batch_packed(pk, out=pres)
os.environ["VLR_CHUNK_TIMING"] = "1"
t0 = time.perf_counter(); eng.call_batch_packed(pk, out=pres); print("packed ms", (time.perf_counter() - t0) * 1e3)
t0 = time.perf_counter(); eng.call_batch(pb, out=pres); print("f32 ms", (time.perf_counter() - t0) * 1e3)
