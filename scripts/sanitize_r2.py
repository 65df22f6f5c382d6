"""Small batches through every round-2 kernel for compute-sanitizer: tumor-normal (octet resident kernel) with and
without AFD, depth skew up to 2000 reads (the four resident size classes, the per-round kernels, deferred loci),
pedigree (all-Set pipeline), and the three entries of the ABI (host f32 columns, packed columns, device pointers).
Results of the three entries must be bitwise equal."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import engine, synth  # noqa: E402

cases = {"tn": synth.tumor_normal(96, seed=3), "ped": synth.pedigree(96, seed=4),
         "skew": synth.tumor_normal(24, seed=5, depth_range=(3, 2000))}
failed = False
for name, (sc, b) in cases.items():
    flat = sc.flatten()
    S, E = flat.n_samples, flat.n_events
    eng = engine.PosteriorEngine(flat)
    host = eng.call_batch(b)
    n_host = eng.launches
    afd = eng.call_batch(b, afd_capacity=48)
    packed = engine.PackedBatch(b)
    pk = eng.call_batch_packed(packed)
    packed.close()
    db = engine.DeviceBatch(b)
    dr = engine.DeviceResults(b.n_loci, S, E)
    eng.call_batch_device(db, dr)
    torch.cuda.synchronize()  # the device entry is asynchronous on the engine's stream: the caller synchronises
    dev = dr.to_host()
    def delta(o):
        a, c = host.log_posteriors, o.log_posteriors
        bits = int(np.count_nonzero(a.view(np.uint64) != c.view(np.uint64)))
        fin = np.isfinite(a) & np.isfinite(c)
        return bits, float(np.max(np.abs(a[fin] - c[fin]))) if fin.any() else 0.0
    rep = {k: delta(o) for k, o in (("afd", afd), ("packed", pk), ("device", dev))}
    print(name, "launches", n_host, "error bits", int(np.count_nonzero(host.status & 0x83f)),
          "entries that differ from the host entry (count, max |d|):", rep, flush=True)
    failed = failed or any(v[0] for v in rep.values())
    del db, dr
    eng.close()
sys.exit(1 if failed else 0)
