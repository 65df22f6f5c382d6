"""Shared helpers for the test-suite (small hand-built batches, comparison metrics)."""
import math

import numpy as np

from varlociraptor_b200 import abi
from varlociraptor_b200.batch import LocusBatch

LN05 = math.log(0.5)
SNV_FLAGS = (abi.LF_CHECK_ROB | abi.LF_CHECK_SB | abi.LF_CHECK_RPB | abi.LF_CHECK_SCB | abi.LF_CHECK_ALB |
             abi.LF_FILTER_NONSTANDARD | abi.LF_HAS_SNV | (ord("A") << abi.LF_REFBASE_SHIFT) |
             (ord("G") << abi.LF_ALTBASE_SHIFT))


def read(prob_mapping=0.0, prob_alt=0.0, prob_ref=-np.inf, prob_missed_allele=None, prob_sample_alt=0.0,
         prob_double_overlap=0.0, prob_hit_base=math.log(0.01), strand=abi.STRAND_BOTH, orientation=abi.ORIENT_NONE,
         major=False, softclipped=False, paired=True, max_mapq=True, alt_locus=abi.ALTLOCUS_NONE,
         hlen=None, hart=None, hvar=None):
    """One read in the shape of `model::tests::observation` (src/variants/model/mod.rs:374-402)."""
    if prob_missed_allele is None:
        prob_missed_allele = float(np.logaddexp(prob_ref, prob_alt)) - math.log(2.0)
    f = (strand << abi.RF_STRAND_SHIFT) | (orientation << abi.RF_ORIENT_SHIFT) | (alt_locus << abi.RF_ALTLOCUS_SHIFT)
    if major:
        f |= abi.RF_READPOS_MAJOR
    if softclipped:
        f |= abi.RF_SOFTCLIPPED
    if paired:
        f |= abi.RF_PAIRED
    if max_mapq:
        f |= abi.RF_MAX_MAPQ
    if hlen is not None:
        f |= abi.RF_HAS_HOMOPOLYMER_LEN | ((hlen & 0xff) << abi.RF_HOMOPOLYMER_LEN_SHIFT)
    return dict(prob_mapping=prob_mapping, prob_alt=prob_alt, prob_ref=prob_ref,
                prob_missed_allele=prob_missed_allele, prob_sample_alt=prob_sample_alt,
                prob_double_overlap=prob_double_overlap, prob_hit_base=prob_hit_base, flags=f,
                hart=np.nan if hart is None else hart, hvar=np.nan if hvar is None else hvar)


def batch_from_reads(loci, locus_flags=None, n_samples=None):
    """loci: list (per locus) of lists (per sample) of lists of `read()` dicts."""
    S = n_samples or len(loci[0])
    offs = [0]
    rows = []
    for locus in loci:
        assert len(locus) == S
        for pile in locus:
            rows.extend(pile)
            offs.append(offs[-1] + len(pile))
    cols = {k: np.array([r[k] for r in rows], dtype=np.float32) for k in abi.BATCH_F32_COLUMNS}
    flags = np.array([r["flags"] for r in rows], dtype=np.uint32)
    hart = np.array([r["hart"] for r in rows], dtype=np.float32)
    hvar = np.array([r["hvar"] for r in rows], dtype=np.float32)
    any_h = bool(np.any(~np.isnan(hart)) or np.any(~np.isnan(hvar)))
    if locus_flags is None:
        locus_flags = [SNV_FLAGS] * len(loci)
    return LocusBatch(S, np.array(offs, dtype=np.int64), cols, flags, np.array(locus_flags, dtype=np.uint32),
                      hart if any_h else None, hvar if any_h else None)


def max_abs_delta(a, b):
    """max |a - b| with the SURVEY §8(d) convention: 0 where both are -inf (or both NaN), +inf where one is."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b)
    d[same] = 0.0
    d[np.isnan(d)] = np.inf
    return float(d.max()) if d.size else 0.0


def phred(lp):
    return -10.0 * np.asarray(lp, dtype=np.float64) / math.log(10.0)


def pair_as_tumor_normal(single: LocusBatch, pairs):
    """Two-sample (normal, tumor) batch from a one-sample batch: locus k = (normal = single[i], tumor = single[j])
    for (i, j) in pairs; the locus flags are the tumor record's."""
    cols = {k: [] for k in abi.BATCH_F32_COLUMNS}
    flags, lflags, harts, hvars = [], [], [], []
    offs = [0]
    has_h = single.prob_homopolymer_artifact is not None
    for i, j in pairs:
        for s in (i, j):
            lo, hi = int(single.read_offsets[s]), int(single.read_offsets[s + 1])
            for k in cols:
                cols[k].append(single.columns[k][lo:hi])
            flags.append(single.read_flags[lo:hi])
            if has_h:
                harts.append(single.prob_homopolymer_artifact[lo:hi])
                hvars.append(single.prob_homopolymer_variant[lo:hi])
            offs.append(offs[-1] + hi - lo)
        lflags.append(single.locus_flags[j])
    cat = np.concatenate
    return LocusBatch(2, np.array(offs, dtype=np.int64), {k: cat(v) for k, v in cols.items()}, cat(flags),
                      np.array(lflags, dtype=np.uint32), cat(harts) if has_h else None, cat(hvars) if has_h else None)


FOUR_SAMPLE_YAML = """
samples:
  a:
    universe: "0.0 | 0.5 | 1.0"
  b:
    universe: "0.0 | 0.5 | 1.0"
  c:
    universe: "[0.0,1.0]"
    resolution: 0.1
  d:
    universe: "[0.0,1.0]"
    resolution: 0.1
    contamination:
      by: c
      fraction: 0.3
events:
  e1: "a:0.5 & b:0.0 & c:]0.0,1.0] & d:]0.0,1.0]"
  e2: "a:0.0 & b:0.0 & c:]0.0,0.5[ & d:0.0"
  e3: "a:1.0 & c:0.0"
  e4: "a:0.0 & b:0.5 & c:0.0 & d:]0.0,1.0]"
"""


def four_sample_batch(n_loci, seed, depth=14):
    """Four-sample loci (exercise the full-capacity engine variant: > 3 samples)."""
    from varlociraptor_b200 import synth
    _, b1 = synth.tumor_normal(n_loci, seed=seed, depth=depth)
    _, b2 = synth.tumor_normal(n_loci, seed=seed + 1, depth=depth)
    cols = {k: [] for k in abi.BATCH_F32_COLUMNS}
    flags, offs = [], [0]
    for i in range(n_loci):
        for b in (b1, b2):
            for s in range(2):
                lo, hi = int(b.read_offsets[2 * i + s]), int(b.read_offsets[2 * i + s + 1])
                for k in cols:
                    cols[k].append(b.columns[k][lo:hi])
                flags.append(b.read_flags[lo:hi])
                offs.append(offs[-1] + hi - lo)
    return LocusBatch(4, np.array(offs, dtype=np.int64), {k: np.concatenate(v) for k, v in cols.items()},
                      np.concatenate(flags), b1.locus_flags)
