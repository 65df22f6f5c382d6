"""A/B of several builds of libvlr_engine.so in one process (same device-resident batch, alternating rounds):
python scripts/ab_libs.py N REPS CFG name1 name2 ...   (name -> varlociraptor_b200/csrc/libvlr_engine_<name>.so, "new" = the
product library). Prints the best device-entry time of each and whether the results are bitwise those of the first."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 2
names = sys.argv[4:] or ["new", "prev"]
here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "varlociraptor_b200", "csrc")
libs = {k: os.path.join(here, "libvlr_engine.so" if k == "new" else "libvlr_engine_%s.so" % k) for k in names}
sc, b = synth.config(cfg, n, seed=synth.SEED_BASE + cfg)
flat = sc.flatten()
S = flat.n_samples
best = {k: 1e30 for k in libs}
ref = None
for rnd in range(2):
    for name, path in libs.items():
        engine._lib, engine._LIB_PATH = None, path
        eng = engine.PosteriorEngine(flat)
        eng.reserve(int(np.max(b.read_offsets[S::S] - b.read_offsets[:-S:S])))
        db = engine.DeviceBatch(b)
        dr = engine.DeviceResults(b.n_loci, S, flat.n_events)
        s = torch.cuda.Stream()
        for i in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            eng.call_batch_device(db, dr, s.cuda_stream)
            e1.record(s)
            torch.cuda.synchronize()
            if i:
                best[name] = min(best[name], e0.elapsed_time(e1))
        post = dr.log_posteriors.cpu().numpy().copy()
        if ref is None:
            ref = post
        same = np.array_equal(post, ref, equal_nan=True)
        with np.errstate(invalid="ignore"):
            dmax = float(np.nanmax(np.abs(np.where(np.isfinite(post) & np.isfinite(ref), post - ref, 0.0))))
        print("%-10s round %d: best %.3f ms = %.3f M loci/s, bitwise equal to the first: %s, max |d| %.2e"
              % (name, rnd, best[name], n / best[name] / 1e3, same, dmax), flush=True)
        del eng, db, dr
