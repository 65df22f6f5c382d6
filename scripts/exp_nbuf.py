"""Experiment: chunks in flight (VLR_NBUF) x loci per chunk (VLR_CHUNK_LOCI) of the host entries."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
combos = [tuple(int(x) for x in a.split(":")) for a in sys.argv[3:]] or [(3, 65536), (6, 65536), (8, 65536)]
sc, b = synth.config(cfg, n, seed=synth.SEED_BASE + cfg)
flat = sc.flatten(); S = flat.n_samples
pb = engine.pin_batch(b); pk = engine.PackedBatch(pb); pres = engine.pinned_results(n, S, flat.n_events)
for nbuf, chunk in combos:
    os.environ["VLR_NBUF"] = str(nbuf)
    os.environ["VLR_CHUNK_LOCI"] = str(chunk)
    eng = engine.PosteriorEngine(flat)
    out = []
    for fn, arg in ((eng.call_batch_packed, pk), (eng.call_batch, pb)):
        fn(arg, out=pres); ts = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(arg, out=pres); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        out.append(n / min(ts) / 1e3)
    print("nbuf %d, chunk %d: packed %.3f M loci/s, f32 %.3f M loci/s" % (nbuf, chunk, out[0], out[1]), flush=True)
    eng.close()
