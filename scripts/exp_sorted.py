"""Experiment: how much of the resident kernel's time is the tail of uneven lcs inside a warp? The same config-2 loci
in generated order, sorted by their number of joint evaluations (known from a first call), and sorted by predictors a
kernel could compute before the rounds (alt-supporting read counts of the two pileups)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
sc, b = synth.config(2, n, seed=synth.SEED_BASE + 2)
flat = sc.flatten()
S = flat.n_samples
eng = engine.PosteriorEngine(flat)
def run(batch, label):
    db = engine.DeviceBatch(batch); dr = engine.DeviceResults(batch.n_loci, S, flat.n_events); s = torch.cuda.Stream()
    best = 1e30
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); eng.call_batch_device(db, dr, s.cuda_stream); e1.record(s); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    print("%-46s %.3f ms = %.3f M loci/s" % (label, best, batch.n_loci / best / 1e3), flush=True)
    return dr.to_host()
r = run(b, "generated order")
run(b.select(np.argsort(r.n_base_events, kind="stable")), "sorted by joint evaluations (a posteriori)")
alt = (b.columns["prob_alt"] > b.columns["prob_ref"]).astype(np.int64)
cs = np.concatenate([[0], np.cumsum(alt)])
per = (cs[b.read_offsets[1:]] - cs[b.read_offsets[:-1]]).reshape(-1, S)  # alt-supporting reads per locus and sample
an, at = per[:, 0], per[:, 1]
run(b.select(np.argsort(at, kind="stable")), "sorted by tumor alt reads")
run(b.select(np.lexsort((at, an))), "sorted by (normal alt, tumor alt)")
bins = np.array([0, 1, 3, 8, 20, 40, 70, 1000])
bt, bn = np.digitize(at, bins), np.digitize(an, bins)
run(b.select(np.lexsort((bt, bn))), "8 x 8 bins of (normal alt, tumor alt)")
run(b.select(np.argsort(bt * 8 + bn, kind="stable")), "8 x 8 bins of (tumor alt, normal alt)")
run(b.select(np.argsort(np.minimum(bt, 3) * 4 + np.minimum(bn, 3), kind="stable")), "4 x 4 bins")
rng = np.random.default_rng(1)
run(b.select(rng.permutation(n)), "random permutation")
