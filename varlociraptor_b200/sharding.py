"""Multi-GPU sharding of a locus batch.

Loci are independent (breakend/haplotype groups are resolved on the host before batching, calling.rs:569-580,726-741),
so the batch is cut into one contiguous locus range per rank, balanced by reads; every rank runs its own engine on
its own GPU with no data-path collective. The only exchange is the final gather of the fixed-stride result records
to rank 0 (`torch.distributed.gather`: NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np

from .batch import CallResults, LocusBatch


def shard_ranges(batch: LocusBatch, world: int, by: str = "reads") -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) locus ranges, one per rank, with (nearly) equal read counts (or locus counts)."""
    L, S = batch.n_loci, batch.n_samples
    if by == "loci" or batch.n_reads == 0:
        cuts = [L * r // world for r in range(world + 1)]
    else:
        ends = batch.read_offsets[S::S].astype(np.float64)  # reads up to and including locus i
        targets = batch.n_reads * (np.arange(1, world) / world)
        inner = np.searchsorted(ends, targets, side="left") + 1
        cuts = [0] + [int(min(max(c, 0), L)) for c in inner] + [L]
        for i in range(1, len(cuts)):  # keep monotone
            cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def pack_records(res: CallResults) -> np.ndarray:
    """Fixed-stride records [log_post (E+1) | log_marginal | map_vaf (S) | best_event | map_config | status | n_base]."""
    cols = [res.log_posteriors, res.log_marginal[:, None], res.map_vaf, res.best_event[:, None].astype(np.float64),
            res.map_config[:, None].astype(np.float64), res.status[:, None].astype(np.float64),
            res.n_base_events[:, None].astype(np.float64)]
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


def unpack_records(rec: np.ndarray, n_samples: int, n_events: int) -> CallResults:
    out = CallResults(rec.shape[0], n_samples, n_events)
    E, S = n_events, n_samples
    out.log_posteriors[...] = rec[:, :E + 1]
    out.log_marginal[...] = rec[:, E + 1]
    out.map_vaf[...] = rec[:, E + 2:E + 2 + S]
    out.best_event[...] = rec[:, E + 2 + S].astype(np.int32)
    out.map_config[...] = rec[:, E + 3 + S].astype(np.int32)
    out.status[...] = rec[:, E + 4 + S].astype(np.uint32)
    out.n_base_events[...] = rec[:, E + 5 + S].astype(np.uint32)
    return out


def call_sharded(compute: Callable[[LocusBatch], CallResults], batch: LocusBatch, n_events: int, rank: int, world: int,
                 group=None, device: Optional[str] = None) -> Optional[CallResults]:
    """Every rank calls this with the same batch description; rank r computes its range, rank 0 gets all results
    (in input order), the others get None. `compute` is the rank's engine entry (PosteriorEngine.call_batch)."""
    import torch
    import torch.distributed as dist
    ranges = shard_ranges(batch, world)
    lo, hi = ranges[rank]
    mine = compute(batch.slice(lo, hi))
    rec = torch.from_numpy(pack_records(mine))
    if world == 1:
        return mine
    width = rec.shape[1]
    sizes = [h - l for l, h in ranges]
    pad = max(sizes)
    buf = torch.zeros((pad, width), dtype=torch.float64)
    buf[:rec.shape[0]] = rec
    if device is not None:
        buf = buf.to(device)
    gathered = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0, group=group)
    if rank != 0:
        return None
    parts = [g.cpu().numpy()[:n] for g, n in zip(gathered, sizes)]
    return unpack_records(np.concatenate(parts, axis=0), batch.n_samples, n_events)
