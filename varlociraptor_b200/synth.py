"""Seeded synthetic locus batches for BASELINE.json's configs.

Every per-read probability is rounded through `MiniLogProb` (src/utils/mod.rs:448-474) so the
inputs are bit-realistic: what `varlociraptor preprocess variants` would have stored on disk.
Read model (per read): allele ~ Bernoulli(effective VAF); base quality Q ~ U{20..40},
e = 10^(-Q/10), with probability e the read shows the other allele;
(prob_alt, prob_ref) = (ln(1-e), ln(e/3)) or swapped; prob_missed_allele =
ln_add_exp(prob_alt, prob_ref) - ln 2 (src/variants/types/mod.rs:100-102); MAPQ = 60 for 95 % of
reads else U{0..59}, prob_mapping = ln(1 - 10^(-MAPQ/10)); prob_sample_alt = 0; strand F/R and
orientation F1R2/F2R1 50:50; prob_double_overlap = -inf; prob_hit_base = -ln 150; read position Some.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

from . import abi
from .batch import LocusBatch, mini_logprob
from .scenario import Scenario

SEED_BASE = 20260101

SIMPLE_PEDIGREE_YAML = """
species:
  heterozygosity: 0.001
  germline-mutation-rate: 1e-3
  ploidy: 2
samples:
  mother:
    sex: female
  father:
    sex: male
  child:
    sex: female
    inheritance:
      mendelian:
        from:
          - mother
          - father
events:
  denovo_child: "(child:0.5 | child:1.0) & mother:0.0 & father:0.0"
  inherited: "!mother:0.0 | !father:0.0"
"""


def _tables():
    q = np.arange(20, 41, dtype=np.float64)
    e = 10.0 ** (-q / 10.0)
    hi = np.log1p(-e)
    lo = np.log(e / 3.0)
    missed = np.logaddexp(hi, lo) - math.log(2.0)
    mapq = np.arange(0, 61, dtype=np.float64)
    with np.errstate(divide="ignore"):
        pm = np.log1p(-(10.0 ** (-mapq / 10.0)))
    return mini_logprob(hi), mini_logprob(lo), mini_logprob(missed), mini_logprob(pm)


def _reads(rng: np.random.Generator, eff_vaf_per_read: np.ndarray):
    """Per-read columns + flags for reads whose alt-sampling probability is given."""
    n = len(eff_vaf_per_read)
    hi, lo, missed, pm_tab = _tables()
    qi = rng.integers(0, 21, size=n, dtype=np.int8)
    e = (10.0 ** (-(qi.astype(np.float32) + 20.0) / 10.0)).astype(np.float32)
    is_alt = rng.random(n, dtype=np.float32) < eff_vaf_per_read.astype(np.float32)
    flip = rng.random(n, dtype=np.float32) < e
    shows_alt = is_alt ^ flip
    mapq = np.where(rng.random(n, dtype=np.float32) < 0.95, 60, rng.integers(0, 60, size=n, dtype=np.int8))
    bits = rng.integers(0, 4, size=n, dtype=np.uint8)
    cols = {
        "prob_alt": np.where(shows_alt, hi[qi], lo[qi]).astype(np.float32),
        "prob_ref": np.where(shows_alt, lo[qi], hi[qi]).astype(np.float32),
        "prob_missed_allele": missed[qi],
        "prob_mapping": pm_tab[mapq],
        "prob_sample_alt": np.zeros(n, dtype=np.float32),
        "prob_double_overlap": np.full(n, -np.inf, dtype=np.float32),
        "prob_hit_base": np.full(n, mini_logprob(np.array([-math.log(150.0)]))[0], dtype=np.float32),
    }
    strand = (bits & 1).astype(np.uint32)  # Forward / Reverse
    orient = ((bits >> 1) & 1).astype(np.uint32)  # F1R2 / F2R1
    flags = (strand << abi.RF_STRAND_SHIFT) | (orient << abi.RF_ORIENT_SHIFT) | \
        (np.uint32(abi.ALTLOCUS_NONE) << abi.RF_ALTLOCUS_SHIFT) | np.uint32(abi.RF_PAIRED)
    flags = flags | np.where(mapq == 60, abi.RF_MAX_MAPQ, 0).astype(np.uint32)
    return cols, flags.astype(np.uint32)


_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _snv_locus_flags(rng, n, snv_mask=None):
    ref = rng.integers(0, 4, size=n)
    alt = (ref + rng.integers(1, 4, size=n)) % 4
    snv = np.uint32(abi.LF_CHECK_ROB | abi.LF_CHECK_SB | abi.LF_CHECK_RPB | abi.LF_CHECK_SCB | abi.LF_CHECK_ALB |
                    abi.LF_FILTER_NONSTANDARD | abi.LF_HAS_SNV | (abi.VARTYPE_SNV << abi.LF_VARTYPE_SHIFT))
    f = snv | (_BASES[ref].astype(np.uint32) << abi.LF_REFBASE_SHIFT) | \
        (_BASES[alt].astype(np.uint32) << abi.LF_ALTBASE_SHIFT)
    if snv_mask is not None:
        # indel records: is_snv_or_mnv = false -> only strand bias and alt locus bias are checked
        # (src/calling/variants/calling.rs:555-566); no homopolymer info in the synthetic reads
        indel = np.uint32(abi.LF_CHECK_SB | abi.LF_CHECK_ALB | (abi.VARTYPE_INDEL << abi.LF_VARTYPE_SHIFT))
        f = np.where(snv_mask, f, indel)
    return f.astype(np.uint32)


def _assemble(n_samples, depths, eff_vaf, rng, locus_flags) -> LocusBatch:
    """depths, eff_vaf: [L, S] arrays (reads per locus/sample; alt sampling probability)."""
    lens = depths.reshape(-1).astype(np.int64)
    offsets = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    per_read_vaf = np.repeat(eff_vaf.reshape(-1).astype(np.float32), lens)
    cols, flags = _reads(rng, per_read_vaf)
    return LocusBatch(n_samples, offsets, cols, flags, locus_flags)


def _tn_draws(rng, n_loci: int, depth: int, depth_range):
    """The per-locus draws of `tumor_normal` that precede the reads: class, true allele frequencies, depths [L, 2]."""
    cls = rng.choice(5, size=n_loci, p=[0.50, 0.25, 0.15, 0.05, 0.05])
    u_t = rng.uniform(0.05, 0.6, size=n_loci)
    u_n = rng.uniform(0.05, 0.3, size=n_loci)
    if depth_range is None:
        depths = np.full((n_loci, 2), depth, dtype=np.int64)
    else:
        lo, hi = depth_range
        depths = np.exp(rng.uniform(math.log(lo), math.log(hi + 1), size=(n_loci, 2))).astype(np.int64)
        depths = np.clip(depths, lo, hi)
    return cls, u_t, u_n, depths


def tumor_normal_depths(n_loci: int, seed: int = SEED_BASE + 2, depth: int = 100,
                        depth_range: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """Reads per locus and sample [L, 2] of `tumor_normal(n_loci, seed, ...)` without generating the reads: what a
    rank needs of the WHOLE batch to cut it into shards of equal work (sharding.shard_cuts)."""
    return _tn_draws(np.random.Generator(np.random.PCG64(seed)), n_loci, depth, depth_range)[3]


def tumor_normal(n_loci: int, seed: int = SEED_BASE + 2, depth: int = 100, purity: float = 0.75,
                 depth_range: Optional[Tuple[int, int]] = None) -> Tuple[Scenario, LocusBatch]:
    """cfg-2 (and cfg-4 by n_loci, cfg-5 by depth_range=(10, 2000)): SNV loci, scenario of
    `call variants tumor-normal` (src/cli.rs:1151-1173). Sample order: normal = 0, tumor = 1."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cls, u_t, u_n, depths = _tn_draws(rng, n_loci, depth, depth_range)
    theta_t = np.zeros(n_loci)
    theta_n = np.zeros(n_loci)
    theta_t[cls == 1] = u_t[cls == 1]
    theta_t[cls == 2] = theta_n[cls == 2] = 0.5
    theta_t[cls == 3] = theta_n[cls == 3] = 1.0
    theta_t[cls == 4] = u_t[cls == 4]
    theta_n[cls == 4] = u_n[cls == 4]
    eff = np.stack([theta_n, purity * theta_t + (1.0 - purity) * theta_n], axis=1)
    batch = _assemble(2, depths, eff, rng, _snv_locus_flags(rng, n_loci))
    return Scenario.tumor_normal(purity=purity), batch


def pedigree(n_loci: int, seed: int = SEED_BASE + 3, depth: int = 100) -> Tuple[Scenario, LocusBatch]:
    """cfg-3: 70 % SNV / 30 % indel loci, simple-pedigree scenario (mother, father, child; the scenario text
    restates tests/resources/prior/scenarios/simple-pedigree.scenario.yaml plus the two events of SURVEY
    Appendix B). Sample order: child = 0, father = 1, mother = 2."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g_f = rng.choice(3, size=n_loci, p=[0.7, 0.2, 0.1])
    g_m = rng.choice(3, size=n_loci, p=[0.7, 0.2, 0.1])
    from_f = rng.random(n_loci) < g_f / 2.0
    from_m = rng.random(n_loci) < g_m / 2.0
    g_c = from_f.astype(np.int64) + from_m.astype(np.int64)
    denovo = (g_f == 0) & (g_m == 0) & (rng.random(n_loci) < 0.1)
    g_c[denovo] = 1
    eff = np.stack([g_c / 2.0, g_f / 2.0, g_m / 2.0], axis=1)
    depths = np.full((n_loci, 3), depth, dtype=np.int64)
    snv_mask = rng.random(n_loci) < 0.7
    batch = _assemble(3, depths, eff, rng, _snv_locus_flags(rng, n_loci, snv_mask))
    return Scenario.from_yaml(SIMPLE_PEDIGREE_YAML), batch


def config_depths(idx: int, n_loci: int, seed: Optional[int] = None) -> np.ndarray:
    """Reads per locus and sample of `config(idx, n_loci, seed)` without generating the reads."""
    seed = SEED_BASE + idx if seed is None else seed
    if idx in (2, 4):
        return np.full((n_loci, 2), 100, dtype=np.int64)
    if idx == 3:
        return np.full((n_loci, 3), 100, dtype=np.int64)
    if idx == 5:
        return tumor_normal_depths(n_loci, seed, depth_range=(10, 2000))
    raise ValueError("config index must be 2..5")


def config(idx: int, n_loci: int, seed: Optional[int] = None):
    """BASELINE.json configs by 1-based index (2..5); seed defaults to SEED_BASE + idx."""
    seed = SEED_BASE + idx if seed is None else seed
    if idx in (2, 4):
        return tumor_normal(n_loci, seed)
    if idx == 3:
        return pedigree(n_loci, seed)
    if idx == 5:
        return tumor_normal(n_loci, seed, depth_range=(10, 2000))
    raise ValueError("config index must be 2..5")
