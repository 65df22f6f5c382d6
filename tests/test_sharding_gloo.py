"""N>1 path on CPU: two gloo processes shard a batch, compute their ranges (with the test-only host emulation of
the engine standing in for the per-GPU engine) and gather to rank 0; the result must equal the single-process one."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from varlociraptor_b200 import synth
from varlociraptor_b200.sharding import call_sharded, pack_records, shard_ranges, unpack_records

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from tests import emu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc, b = synth.tumor_normal(40, seed=77, depth_range=(10, 120))
    flat = sc.flatten()
    res = call_sharded(lambda sub: emu.call_batch(flat, sub, afd_capacity=48), b, flat.n_events, rank, world)
    if rank == 0:
        np.save(out_path, pack_records(res))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(tmp_path):
    from tests import emu
    emu.build()
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    sc, b = synth.tumor_normal(40, seed=77, depth_range=(10, 120))
    flat = sc.flatten()
    want = pack_records(emu.call_batch(flat, b, afd_capacity=48))  # the distributions travel with the records
    got = np.load(out)
    assert got.shape == want.shape
    assert np.array_equal(got, want, equal_nan=True)


def test_shard_ranges_balance_and_cover_everything():
    from varlociraptor_b200.sharding import batch_depths, locus_work, shard_cuts
    _, b = synth.tumor_normal(500, seed=5, depth_range=(10, 2000))
    work = locus_work(batch_depths(b))
    for world in (1, 2, 4, 8):
        for by in ("work", "reads"):
            r = shard_ranges(b, world, by=by)
            assert r[0][0] == 0 and r[-1][1] == b.n_loci
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            if by == "reads":
                load = [int(b.read_offsets[hi * 2] - b.read_offsets[lo * 2]) for lo, hi in r]
                assert max(load) - min(load) <= 2 * 4000  # within one locus of perfect balance
            else:
                load = [float(work[lo:hi].sum()) for lo, hi in r]
                assert max(load) - min(load) <= 2 * work.max()
    assert shard_ranges(b, 3, by="loci")[1] == (166, 333)
    # the cut only needs the per-locus depths, which the generator provides without the reads (bench.py --config 5)
    assert np.array_equal(synth.tumor_normal_depths(500, seed=5, depth_range=(10, 2000)), batch_depths(b))
    assert shard_cuts(np.ones(10), 3) == [0, 4, 7, 10] and shard_cuts(np.zeros(0), 2) == [0, 0, 0]


def test_record_pack_roundtrip():
    from varlociraptor_b200.batch import CallResults
    r = CallResults(5, 2, 4)
    rng = np.random.default_rng(0)
    r.log_posteriors[...] = -rng.random((5, 5))
    r.log_posteriors[2, 1] = -np.inf
    r.log_marginal[...] = -rng.random(5) * 100
    r.map_vaf[...] = rng.random((5, 2))
    r.best_event[...] = [0, 2, 4, 6, 8]
    r.status[...] = [0, 1 << 8, 1 << 9, 0, 1 << 10]
    back = unpack_records(pack_records(r), 2, 4)
    assert np.array_equal(back.log_posteriors, r.log_posteriors) and np.array_equal(back.status, r.status)
    assert np.array_equal(back.best_event, r.best_event) and np.array_equal(back.map_vaf, r.map_vaf)
    a = CallResults(3, 2, 4, afd_capacity=5)
    a.afd_count[...] = [[1, 2], [0, 5], [3, 3]]
    a.afd_vaf[...] = rng.random((3, 2, 5))
    a.afd_logp[...] = -rng.random((3, 2, 5))
    back = unpack_records(pack_records(a), 2, 4, afd_capacity=5)
    assert np.array_equal(back.afd_count, a.afd_count) and np.array_equal(back.afd_vaf, a.afd_vaf)
    assert np.array_equal(back.afd_logp, a.afd_logp)
