"""Independent check of the oracle's log-space pileup likelihood: the model of likelihood.rs written
out in LINEAR space from the reference's formulas and evaluated with 60 significant digits (mpmath), on realistic
synthetic pileups (config 2 and 3 reads: f16/f32-quantised probabilities, MAPQ 0..60, both strands) and on edge values.

  per read (likelihood.rs:43-53, :171-220, :108-135; Artifacts::none(): bias/mod.rs:259-284, strand_bias.rs:28-53,
  read_orientation_bias.rs:17-29, read_position_bias.rs:17-61, softclip_bias.rs:14-24, alt_locus_bias.rs:62-112):
    s      = vaf == 1 ? 1 : min(vaf * e^prob_sample_alt, 1)
    L(vaf) = s * bA * e^prob_alt + (1 - s) * e^prob_ref * bR
    single:        e^pm * L(vaf)                                        + (1 - e^pm) * e^prob_missed_allele * bAny
    contaminated:  e^pm * (purity * L(vaf1) + (1 - purity) * L(vaf2))   + (1 - e^pm) * e^prob_missed_allele * bAny
  pileup = sum of the logs. This pins the oracle's arithmetic (ln_add_exp / ln_sum_exp / ln_one_minus_exp chains,
  -inf handling) to the stated model at 1e-10 absolute for 100-read pileups; it says nothing about whether the model
  is the reference's - that is what the golden pair and the testcase expectations are for."""
import math

import mpmath as mp
import numpy as np

from oracle import oracle
from tests.util import batch_from_reads, read
from varlociraptor_b200 import abi, synth

mp.mp.dps = 60


def _e(x):
    """e^x of an f32 column value, exactly; e^-inf = 0."""
    x = float(x)
    return mp.mpf(0) if x == -math.inf else mp.e ** mp.mpf(x)


def _pileup_ln_likelihood(b, lo, hi, vaf, vaf2=0.0, purity=1.0, contaminated=False):
    c = b.columns
    total = mp.mpf(0)
    for r in range(lo, hi):
        f = int(b.read_flags[r])
        strand = (f >> abi.RF_STRAND_SHIFT) & 3
        major = bool(f & abi.RF_READPOS_MAJOR)
        pm = _e(c["prob_mapping"][r])
        pdo = _e(c["prob_double_overlap"][r])
        phb = _e(c["prob_hit_base"][r])
        sb_alt = {0: mp.mpf("0.5") * (1 - pdo), 1: mp.mpf("0.5") * (1 - pdo), 2: pdo, 3: mp.mpf(1)}[strand]
        rpb = phb if major else 1 - phb
        half = mp.mpf("0.5")
        b_alt = sb_alt * half * rpb * half          # strand, orientation, position, (softclip 1, homopolymer 1), alt locus
        b_ref = half * half * rpb * half
        b_any = half * half * rpb * half

        def mapped(v):
            v = mp.mpf(v)
            s = mp.mpf(1) if v == 1 else min(v * _e(c["prob_sample_alt"][r]), mp.mpf(1))
            return s * b_alt * _e(c["prob_alt"][r]) + (1 - s) * _e(c["prob_ref"][r]) * b_ref
        mis = (1 - pm) * _e(c["prob_missed_allele"][r]) * b_any
        if contaminated:
            p = pm * (mp.mpf(purity) * mapped(vaf) + (1 - mp.mpf(purity)) * mapped(vaf2)) + mis
        else:
            p = pm * mapped(vaf) + mis
        total += mp.log(p) if p > 0 else mp.mpf("-inf")
    return total


def _check(b, lo, hi, vaf, vaf2=0.0, purity=1.0, contaminated=False, tol=1e-10):
    got = oracle.pileup_likelihood(b, lo, hi, vaf, vaf2, purity, contaminated)
    want = _pileup_ln_likelihood(b, lo, hi, vaf, vaf2, purity, contaminated)
    if want == mp.mpf("-inf"):
        assert got == -math.inf
        return 0.0
    d = abs(float(mp.mpf(got) - want))
    assert d <= tol, (vaf, vaf2, purity, contaminated, got, float(want), d)
    return d


def test_single_sample_pileups_config2_and_3():
    worst = 0.0
    for cfg, seed in ((2, 11), (3, 12)):
        _, b = synth.config(cfg, 6, seed=seed)
        S = b.n_samples
        for locus in range(b.n_loci):
            for s in range(S):
                lo, hi = int(b.read_offsets[locus * S + s]), int(b.read_offsets[locus * S + s + 1])
                for vaf in (0.0, 1e-4, 0.01, 0.25, 0.5, 0.73, 1.0):
                    worst = max(worst, _check(b, lo, hi, vaf))
    assert worst > 0.0  # (the comparison is not vacuous: the two evaluations do differ in the last bits)


def test_contaminated_sample_pileups():
    _, b = synth.config(2, 5, seed=13)
    S = b.n_samples
    for locus in range(b.n_loci):
        lo, hi = int(b.read_offsets[locus * S + 1]), int(b.read_offsets[locus * S + 2])  # the tumor sample's reads
        for purity in (0.75, 0.2, 1.0):
            for vaf, vaf2 in ((0.3, 0.0), (0.0, 0.5), (1.0, 0.5), (0.12, 1.0), (0.0, 0.0), (0.5, 0.5)):
                _check(b, lo, hi, vaf, vaf2, purity, True)


def test_edge_values():
    ninf = -np.inf
    reads = [
        read(prob_mapping=0.0, prob_alt=ninf, prob_ref=0.0),                       # certain reference read, MAPQ inf
        read(prob_mapping=ninf, prob_alt=-1.0, prob_ref=-2.0),                     # unmappable read
        read(prob_mapping=np.log1p(-1e-6), prob_alt=-0.001, prob_ref=-30.0, strand=2, prob_double_overlap=-0.5),
        read(prob_mapping=np.log(0.5), prob_alt=-40.0, prob_ref=-0.0001, strand=3),
        read(prob_mapping=np.log1p(-1e-3), prob_alt=-0.01, prob_ref=-12.0, prob_sample_alt=-0.7, strand=1),  # indel-like
        read(prob_mapping=np.log1p(-1e-3), prob_alt=-9.0, prob_ref=-0.02, prob_sample_alt=-0.7, strand=0,
             prob_double_overlap=-3.0),
    ]
    b = batch_from_reads([[reads]])
    for vaf in (0.0, 0.05, 0.5, 1.0):
        _check(b, 0, len(reads), vaf)
        for k in range(len(reads)):
            _check(b, k, k + 1, vaf)
        _check(b, 0, len(reads), vaf, 0.3, 0.6, True)
