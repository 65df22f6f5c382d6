"""Small batches through every round-2 kernel for compute-sanitizer: tumor-normal (octet resident kernel) with and
without AFD, depth skew up to 2000 reads (the four resident size classes, the per-round kernels, deferred loci),
pedigree (all-Set pipeline), and the three entries of the ABI (host f32 columns, packed columns, device pointers).
Results of the three entries must be bitwise equal."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varlociraptor_b200 import engine, synth  # noqa: E402

cases = {"tn": synth.tumor_normal(160, seed=3), "ped": synth.pedigree(160, seed=4),
         "skew": synth.tumor_normal(40, seed=5, depth_range=(3, 2000))}
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
    dev = dr.to_host()
    same = all(np.array_equal(host.log_posteriors.view(np.uint64), o.log_posteriors.view(np.uint64))
               for o in (afd, pk, dev))
    print(name, "launches", n_host, "error bits", int(np.count_nonzero(host.status & 0x83f)),
          "entries bitwise equal:", same, flush=True)
    assert same
    del db, dr
    eng.close()
