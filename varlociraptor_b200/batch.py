"""LocusBatch: the structure-of-arrays input of `vlr_call_batch` (include/vlr_engine.h, vlr_batch_t).

Replaces the AoS `Vec<Pileup>` of `ReadObservation` structs the reference hands to
`model.compute` (src/variants/evidence/observations/read_observation.rs:221-280,
src/calling/variants/calling.rs:586-626). Reads of locus i, sample s occupy rows
read_offsets[i*S+s] .. read_offsets[i*S+s+1] of every per-read column, in pileup order.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import abi


def mini_logprob(x: np.ndarray) -> np.ndarray:
    """Round ln-probabilities through `MiniLogProb` (src/utils/mod.rs:448-474): f16 when the value is
    < -10 and the f16 projection keeps the integer floor, else f32. Returns float32 (f16 ⊂ f32)."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        half = x.astype(np.float16)
        proj = half.astype(np.float64)
        use_half = (x < -10.0) & (np.floor(proj) == np.floor(x))
    out = x.astype(np.float32)
    out[use_half] = half[use_half].astype(np.float32)
    return out


class LocusBatch:
    def __init__(self, n_samples: int, read_offsets: np.ndarray, columns: dict, read_flags: np.ndarray,
                 locus_flags: np.ndarray, prob_homopolymer_artifact: Optional[np.ndarray] = None,
                 prob_homopolymer_variant: Optional[np.ndarray] = None,
                 locus_heterozygosity_phred: Optional[np.ndarray] = None,
                 locus_semr_phred: Optional[np.ndarray] = None):
        self.n_samples = int(n_samples)
        self.read_offsets = np.ascontiguousarray(read_offsets, dtype=np.int64)
        self.n_loci = (len(self.read_offsets) - 1) // self.n_samples
        assert len(self.read_offsets) == self.n_loci * self.n_samples + 1
        self.n_reads = int(self.read_offsets[-1])
        self.columns = {k: np.ascontiguousarray(columns[k], dtype=np.float32) for k in abi.BATCH_F32_COLUMNS}
        for k, v in self.columns.items():
            assert len(v) == self.n_reads, (k, len(v), self.n_reads)
        self.read_flags = np.ascontiguousarray(read_flags, dtype=np.uint32)
        self.locus_flags = np.ascontiguousarray(locus_flags, dtype=np.uint32)
        assert len(self.read_flags) == self.n_reads and len(self.locus_flags) == self.n_loci

        def opt(a, n):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float32)
            assert len(a) == n
            return a
        self.prob_homopolymer_artifact = opt(prob_homopolymer_artifact, self.n_reads)
        self.prob_homopolymer_variant = opt(prob_homopolymer_variant, self.n_reads)
        self.locus_heterozygosity_phred = opt(locus_heterozygosity_phred, self.n_loci)
        self.locus_semr_phred = opt(locus_semr_phred, self.n_loci)

    # ------------------------------------------------------------------ views
    def as_c(self) -> abi.Batch:
        b = abi.Batch()
        b.n_loci = self.n_loci
        b.n_reads = self.n_reads
        b.read_offsets = abi.ptr(self.read_offsets, abi.C.c_int64)
        for k in abi.BATCH_F32_COLUMNS:
            setattr(b, k, abi.ptr(self.columns[k], abi.C.c_float))
        b.read_flags = abi.ptr(self.read_flags, abi.C.c_uint32)
        b.prob_homopolymer_artifact = abi.ptr(self.prob_homopolymer_artifact, abi.C.c_float)
        b.prob_homopolymer_variant = abi.ptr(self.prob_homopolymer_variant, abi.C.c_float)
        b.locus_flags = abi.ptr(self.locus_flags, abi.C.c_uint32)
        b.locus_heterozygosity_phred = abi.ptr(self.locus_heterozygosity_phred, abi.C.c_float)
        b.locus_semr_phred = abi.ptr(self.locus_semr_phred, abi.C.c_float)
        return b

    def nbytes(self) -> int:
        n = self.read_offsets.nbytes + self.read_flags.nbytes + self.locus_flags.nbytes
        n += sum(v.nbytes for v in self.columns.values())
        for a in (self.prob_homopolymer_artifact, self.prob_homopolymer_variant, self.locus_heterozygosity_phred,
                  self.locus_semr_phred):
            if a is not None:
                n += a.nbytes
        return n

    def slice(self, lo: int, hi: int) -> "LocusBatch":
        """Contiguous locus range [lo, hi) as a new batch (offsets rebased)."""
        S = self.n_samples
        r0, r1 = int(self.read_offsets[lo * S]), int(self.read_offsets[hi * S])
        offs = self.read_offsets[lo * S: hi * S + 1] - r0

        def cut(a, a0, a1):
            return None if a is None else a[a0:a1]
        return LocusBatch(S, offs, {k: v[r0:r1] for k, v in self.columns.items()}, self.read_flags[r0:r1],
                          self.locus_flags[lo:hi], cut(self.prob_homopolymer_artifact, r0, r1),
                          cut(self.prob_homopolymer_variant, r0, r1), cut(self.locus_heterozygosity_phred, lo, hi),
                          cut(self.locus_semr_phred, lo, hi))

    def select(self, loci: Sequence[int]) -> "LocusBatch":
        """Arbitrary subset / permutation of loci as a new batch."""
        S = self.n_samples
        loci = np.asarray(loci, dtype=np.int64)
        starts = self.read_offsets[:-1].reshape(self.n_loci, S)[loci].reshape(-1)
        ends = self.read_offsets[1:].reshape(self.n_loci, S)[loci].reshape(-1)
        lens = ends - starts
        offs = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        idx = np.repeat(starts - offs[:-1], lens) + np.arange(offs[-1], dtype=np.int64)

        def take(a, i):
            return None if a is None else a[i]
        return LocusBatch(S, offs, {k: v[idx] for k, v in self.columns.items()}, self.read_flags[idx],
                          self.locus_flags[loci], take(self.prob_homopolymer_artifact, idx),
                          take(self.prob_homopolymer_variant, idx), take(self.locus_heterozygosity_phred, loci),
                          take(self.locus_semr_phred, loci))

    def save(self, path: str) -> None:
        d = {"n_samples": np.array(self.n_samples), "read_offsets": self.read_offsets, "read_flags": self.read_flags,
             "locus_flags": self.locus_flags}
        d.update(self.columns)
        for k in ("prob_homopolymer_artifact", "prob_homopolymer_variant", "locus_heterozygosity_phred",
                  "locus_semr_phred"):
            if getattr(self, k) is not None:
                d[k] = getattr(self, k)
        np.savez_compressed(path, **d)

    @staticmethod
    def load(path: str) -> "LocusBatch":
        z = np.load(path)
        opt = {k: (z[k] if k in z.files else None) for k in (
            "prob_homopolymer_artifact", "prob_homopolymer_variant", "locus_heterozygosity_phred",
            "locus_semr_phred")}
        return LocusBatch(int(z["n_samples"]), z["read_offsets"], {k: z[k] for k in abi.BATCH_F32_COLUMNS},
                          z["read_flags"], z["locus_flags"], **opt)

    @staticmethod
    def concat(batches: List["LocusBatch"]) -> "LocusBatch":
        S = batches[0].n_samples
        offs = [np.zeros(1, dtype=np.int64)]
        base = 0
        for b in batches:
            offs.append(b.read_offsets[1:] + base)
            base += b.n_reads

        def cat(name, per_read=True):
            arrs = [getattr(b, name) for b in batches]
            if all(a is None for a in arrs):
                return None
            out = []
            for b, a in zip(batches, arrs):
                n = b.n_reads if per_read else b.n_loci
                out.append(np.full(n, np.nan, dtype=np.float32) if a is None else a)
            return np.concatenate(out)
        return LocusBatch(S, np.concatenate(offs),
                          {k: np.concatenate([b.columns[k] for b in batches]) for k in abi.BATCH_F32_COLUMNS},
                          np.concatenate([b.read_flags for b in batches]),
                          np.concatenate([b.locus_flags for b in batches]),
                          cat("prob_homopolymer_artifact"), cat("prob_homopolymer_variant"),
                          cat("locus_heterozygosity_phred", False), cat("locus_semr_phred", False))


class CallResults:
    """Host-side result buffers of `vlr_call_batch` (vlr_results_t)."""

    def __init__(self, n_loci: int, n_samples: int, n_events: int, afd_capacity: int = 0):
        self.n_loci, self.n_samples, self.n_events, self.afd_capacity = n_loci, n_samples, n_events, afd_capacity
        self.log_posteriors = np.full((n_loci, n_events + 1), np.nan, dtype=np.float64)
        self.log_marginal = np.full(n_loci, np.nan, dtype=np.float64)
        self.map_vaf = np.full((n_loci, n_samples), np.nan, dtype=np.float64)
        self.map_config = np.zeros(n_loci, dtype=np.int32)
        self.best_event = np.zeros(n_loci, dtype=np.int32)
        self.status = np.zeros(n_loci, dtype=np.uint32)
        self.n_base_events = np.zeros(n_loci, dtype=np.uint32)
        if afd_capacity > 0:
            self.afd_count = np.zeros((n_loci, n_samples), dtype=np.int32)
            self.afd_vaf = np.full((n_loci, n_samples, afd_capacity), np.nan, dtype=np.float64)
            self.afd_logp = np.full((n_loci, n_samples, afd_capacity), np.nan, dtype=np.float64)
        else:
            self.afd_count = self.afd_vaf = self.afd_logp = None

    def as_c(self) -> abi.Results:
        r = abi.Results()
        C = abi.C
        r.log_posteriors = abi.ptr(self.log_posteriors, C.c_double)
        r.log_marginal = abi.ptr(self.log_marginal, C.c_double)
        r.map_vaf = abi.ptr(self.map_vaf, C.c_double)
        r.map_config = abi.ptr(self.map_config, C.c_int32)
        r.best_event = abi.ptr(self.best_event, C.c_int32)
        r.status = abi.ptr(self.status, C.c_uint32)
        r.n_base_events = abi.ptr(self.n_base_events, C.c_uint32)
        r.afd_capacity = self.afd_capacity
        r.afd_count = abi.ptr(self.afd_count, C.c_int32)
        r.afd_vaf = abi.ptr(self.afd_vaf, C.c_double)
        r.afd_logp = abi.ptr(self.afd_logp, C.c_double)
        return r

    def slice(self, lo: int, hi: int) -> "CallResults":
        """Results of the loci [lo, hi) (views, not copies)."""
        r = CallResults.__new__(CallResults)
        r.n_loci, r.n_samples, r.n_events, r.afd_capacity = hi - lo, self.n_samples, self.n_events, self.afd_capacity
        for name in ("log_posteriors", "log_marginal", "map_vaf", "map_config", "best_event", "status", "n_base_events",
                     "afd_count", "afd_vaf", "afd_logp"):
            a = getattr(self, name)
            setattr(r, name, None if a is None else a[lo:hi])
        return r

    def afd(self, locus: int, sample: int):
        n = int(self.afd_count[locus, sample])
        return self.afd_vaf[locus, sample, :n].copy(), self.afd_logp[locus, sample, :n].copy()
