"""Multi-GPU sharding of a locus batch.

Loci are independent (breakend/haplotype groups are resolved on the host before batching, calling.rs:569-580,726-741),
so the batch is cut into one contiguous locus range per rank, balanced by estimated work; every rank runs its own
engine on its own GPU with no data-path collective. The only exchange is the final gather of the fixed-stride result
records to rank 0 (`torch.distributed.gather`: NCCL over NVLink on GPUs, gloo in the CPU tests).

Work estimate of a locus (class-agnostic: nothing of the result is known before it is computed): the number of joint
evaluations hardly depends on the depth, and one evaluation is a pass over the pileup of the sample being integrated
(the last sample of the scenario: the tumor) plus a fixed bookkeeping cost; the other samples' pileups are evaluated
once per integration. Measured on config 5 (depth 10..2000 per sample): work ~ OVERHEAD + depth_leaf + 0.15 * others.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .batch import CallResults, LocusBatch

WORK_OVERHEAD_READS = 60.0   # per-locus cost that does not scale with depth, in read-equivalents
WORK_OTHER_SAMPLES = 0.15    # weight of a read of a sample that is not the innermost integration variable


def locus_work(depths: np.ndarray) -> np.ndarray:
    """Estimated work per locus from the reads per locus and sample, shape [L, S] -> [L]."""
    d = np.asarray(depths, dtype=np.float64)
    return WORK_OVERHEAD_READS + d[:, -1] + WORK_OTHER_SAMPLES * d[:, :-1].sum(axis=1)


def shard_cuts(weights: np.ndarray, world: int) -> List[int]:
    """world + 1 monotone cut positions over len(weights) loci: contiguous ranges of (nearly) equal total weight."""
    w = np.asarray(weights, dtype=np.float64)
    L = len(w)
    total = float(w.sum())
    if L == 0 or total <= 0.0:
        return [L * r // world for r in range(world + 1)]
    ends = np.cumsum(w)  # weight up to and including locus i
    targets = total * (np.arange(1, world) / world)
    inner = np.searchsorted(ends, targets, side="left") + 1
    cuts = [0] + [int(min(max(c, 0), L)) for c in inner] + [L]
    for i in range(1, len(cuts)):  # keep monotone
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts


def batch_depths(batch: LocusBatch) -> np.ndarray:
    S = batch.n_samples
    return (batch.read_offsets[1:] - batch.read_offsets[:-1]).reshape(-1, S)


def shard_ranges(batch: LocusBatch, world: int, by: str = "work") -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) locus ranges, one per rank: equal estimated work (default), equal reads or equal loci."""
    L = batch.n_loci
    if by == "loci" or batch.n_reads == 0:
        cuts = [L * r // world for r in range(world + 1)]
    elif by == "reads":
        cuts = shard_cuts(batch_depths(batch).sum(axis=1), world)
    elif by == "work":
        cuts = shard_cuts(locus_work(batch_depths(batch)), world)
    else:
        raise ValueError("by must be 'work', 'reads' or 'loci'")
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def record_width(n_samples: int, n_events: int, afd_capacity: int = 0) -> int:
    return (n_events + 1) + 1 + n_samples + 4 + n_samples * (1 + 2 * afd_capacity)


def pack_records(res: CallResults) -> np.ndarray:
    """Fixed-stride records [log_post (E+1) | log_marginal | map_vaf (S) | best_event | map_config | status | n_base |
    afd_count (S) | afd_vaf (S x cap) | afd_logp (S x cap)] — the allele frequency distributions travel with the record
    so that the gathered results do not depend on the number of ranks."""
    L = res.log_posteriors.shape[0]
    cols = [res.log_posteriors, res.log_marginal[:, None], res.map_vaf, res.best_event[:, None].astype(np.float64),
            res.map_config[:, None].astype(np.float64), res.status[:, None].astype(np.float64),
            res.n_base_events[:, None].astype(np.float64)]
    if res.afd_capacity:
        cols += [res.afd_count.astype(np.float64), res.afd_vaf.reshape(L, -1), res.afd_logp.reshape(L, -1)]
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


def unpack_records(rec: np.ndarray, n_samples: int, n_events: int, afd_capacity: int = 0) -> CallResults:
    out = CallResults(rec.shape[0], n_samples, n_events, afd_capacity)
    E, S = n_events, n_samples
    out.log_posteriors[...] = rec[:, :E + 1]
    out.log_marginal[...] = rec[:, E + 1]
    out.map_vaf[...] = rec[:, E + 2:E + 2 + S]
    out.best_event[...] = rec[:, E + 2 + S].astype(np.int32)
    out.map_config[...] = rec[:, E + 3 + S].astype(np.int32)
    out.status[...] = rec[:, E + 4 + S].astype(np.uint32)
    out.n_base_events[...] = rec[:, E + 5 + S].astype(np.uint32)
    if afd_capacity:
        o = E + 6 + S
        out.afd_count[...] = rec[:, o:o + S].astype(np.int32)
        o += S
        out.afd_vaf[...] = rec[:, o:o + S * afd_capacity].reshape(-1, S, afd_capacity)
        o += S * afd_capacity
        out.afd_logp[...] = rec[:, o:o + S * afd_capacity].reshape(-1, S, afd_capacity)
    return out


def gather_records(rec, sizes: Sequence[int], rank: int, world: int, group=None):
    """`rec`: this rank's records (torch tensor [n, width], on the device the process group communicates on); rank 0
    gets the list of all ranks' records (padded to the largest shard), the others None. One collective."""
    import torch
    import torch.distributed as dist
    pad = max(sizes)
    if rec.shape[0] == pad:
        buf = rec.contiguous()
    else:
        buf = torch.zeros((pad, rec.shape[1]), dtype=rec.dtype, device=rec.device)
        buf[:rec.shape[0]] = rec
    gathered = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0, group=group)
    return gathered


def call_sharded(compute: Callable[[LocusBatch], CallResults], batch: LocusBatch, n_events: int, rank: int, world: int,
                 group=None, device: Optional[str] = None, by: str = "work") -> Optional[CallResults]:
    """Every rank calls this with the same batch description; rank r computes its range, rank 0 gets all results
    (in input order), the others get None. `compute` is the rank's engine entry (PosteriorEngine.call_batch)."""
    import torch
    ranges = shard_ranges(batch, world, by)
    lo, hi = ranges[rank]
    mine = compute(batch.slice(lo, hi))
    if world == 1:
        return mine
    rec = torch.from_numpy(pack_records(mine))
    if device is not None:
        rec = rec.to(device)
    sizes = [h - l for l, h in ranges]
    gathered = gather_records(rec, sizes, rank, world, group)
    if rank != 0:
        return None
    parts = [g.cpu().numpy()[:n] for g, n in zip(gathered, sizes)]
    return unpack_records(np.concatenate(parts, axis=0), batch.n_samples, n_events, mine.afd_capacity)
