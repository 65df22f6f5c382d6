"""Experiment: host-side knobs of the wavefront pipeline on one device-resident batch: streams the parts of a batch are
spread over (VLR_WAVE_STREAMS), loci per sub-chunk (VLR_WAVE_SUB, 0 = the built-in rule) and the batch size from which a
batch is split at all (VLR_WAVE_SPLIT_MIN).   python scripts/exp_knobs.py N CFG streams:sub:splitmin ..."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
combos = [tuple(int(x) for x in a.split(":")) for a in sys.argv[3:]] or [(3, 0, 0)]
sc, b = synth.config(cfg, n, seed=synth.SEED_BASE + cfg)
flat = sc.flatten(); S = flat.n_samples
db = engine.DeviceBatch(b); dr = engine.DeviceResults(n, S, flat.n_events); s = torch.cuda.Stream()
ref = None
for streams, sub, split in combos:
    os.environ["VLR_WAVE_STREAMS"] = str(streams)
    for k, v in (("VLR_WAVE_SUB", sub), ("VLR_WAVE_SPLIT_MIN", split)):
        if v: os.environ[k] = str(v)
        else: os.environ.pop(k, None)
    eng = engine.PosteriorEngine(flat)
    eng.reserve(int(np.max(b.read_offsets[S::S] - b.read_offsets[:-S:S])))
    best = 1e30
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); eng.call_batch_device(db, dr, s.cuda_stream); e1.record(s); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    post = dr.log_posteriors.cpu().numpy()
    if ref is None: ref = post.copy()
    print("cfg %d, %d loci, streams %d, sub-chunk %6d, split from %6d: %.3f ms = %.3f M loci/s, bitwise equal: %s, %.1f GB in use"
          % (cfg, n, streams, sub, split, best, n / best / 1e3, np.array_equal(post, ref, equal_nan=True),
             (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9), flush=True)
    eng.close()
