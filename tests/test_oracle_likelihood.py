"""The reference's own unit tests for the pileup likelihood (src/variants/model/likelihood.rs:273-394),
re-expressed against the oracle, plus model-level properties."""
import math

import numpy as np

from oracle import oracle
from tests.util import LN05, batch_from_reads, read
from varlociraptor_b200 import Scenario, synth

NEG_INF = -np.inf


def _ref_obs():  # observation(ln_one, ln_zero, ln_one): a certain reference read
    return read(prob_mapping=0.0, prob_alt=NEG_INF, prob_ref=0.0)


def _alt_obs():
    return read(prob_mapping=0.0, prob_alt=0.0, prob_ref=NEG_INF)


def _none_prob_ref(r):
    """Artifacts::none().prob_ref(obs) (bias/mod.rs:268-275): ln .5 (strand) + ln .5 (orientation) +
    ln(1 - e^prob_hit_base) (read position Some) + 0 (softclip) + 0 (homopolymer) + ln .5 (alt locus)."""
    return LN05 + LN05 + math.log1p(-math.exp(float(np.float32(r["prob_hit_base"])))) + LN05


def test_likelihood_observation_absent_single():
    b = batch_from_reads([[[_ref_obs()]]])
    lh = oracle.pileup_likelihood(b, 0, 1, 0.0)
    assert math.isclose(lh, _none_prob_ref(_ref_obs()), rel_tol=1e-12)


def test_likelihood_observation_absent_contaminated():
    b = batch_from_reads([[[_ref_obs()]]])
    lh = oracle.pileup_likelihood(b, 0, 1, 0.0, 0.0, purity=1.0, contaminated=True)
    assert math.isclose(lh, _none_prob_ref(_ref_obs()), rel_tol=1e-12)


def test_likelihood_pileup_absent_both_models():
    b = batch_from_reads([[[_ref_obs() for _ in range(10)]]])
    want = 10 * _none_prob_ref(_ref_obs())
    assert math.isclose(oracle.pileup_likelihood(b, 0, 10, 0.0), want, rel_tol=1e-12)
    assert math.isclose(oracle.pileup_likelihood(b, 0, 10, 0.0, 0.0, 1.0, True), want, rel_tol=1e-12)


def test_likelihood_pileup_maximal_at_true_vaf():
    b = batch_from_reads([[[_alt_obs() for _ in range(5)] + [_ref_obs() for _ in range(5)]]])
    lh = oracle.pileup_likelihood(b, 0, 10, 0.5, 0.0, 1.0, True)
    for af in np.linspace(0.0, 1.0, 10):
        if af != 0.5:
            assert lh > oracle.pileup_likelihood(b, 0, 10, float(af), 0.0, 1.0, True)


def test_posteriors_sum_to_one_and_map_inside_best_event():
    sc, b = synth.tumor_normal(60, seed=7)
    flat = sc.flatten()
    out = oracle.call_batch(flat, b, afd_capacity=96)
    assert np.all(out.status & 0x3f == 0)
    total = np.logaddexp.reduce(out.log_posteriors, axis=1)
    assert np.max(np.abs(total)) < 1e-12
    names = flat.event_names
    for i in range(b.n_loci):
        e = out.best_event[i] // 2
        n, t = out.map_vaf[i]
        if out.status[i] & (1 << 8):  # is_artifact: MAP reports AF 0
            continue
        if names[e] == "absent":
            assert (n, t) == (0.0, 0.0)
        elif names[e] == "somatic_tumor":
            assert n == 0.0 and t > 0.0
        elif names[e] == "germline_het":
            assert n == 0.5 and t > 0.0
        elif names[e] == "germline_hom":
            assert n == 1.0 and t > 0.0
        else:
            assert 0.0 < n < 0.5 and t > 0.0


def test_clear_ref_shortcut():
    """All reads strongly reference in both samples (n > 10): Set nodes with only VAFs > 0 are cut to -inf
    (generic.rs:295-300); ranges that *start* at 0 (even exclusively) are still integrated (:342-347)."""
    pile = [read(prob_mapping=math.log1p(-1e-6), prob_alt=math.log(1e-3 / 3), prob_ref=math.log1p(-1e-3),
                 strand=i % 2, orientation=i % 2) for i in range(20)]
    b = batch_from_reads([[pile, pile]])
    flat = Scenario.tumor_normal(0.75).flatten()
    out = oracle.call_batch(flat, b)
    lp = dict(zip(flat.event_names + ["artifact"], out.log_posteriors[0]))
    assert np.isneginf(lp["germline_het"]) and np.isneginf(lp["germline_hom"])
    assert np.isfinite(lp["somatic_tumor"]) and np.isfinite(lp["somatic_normal"])
    assert lp["absent"] > -0.1 and lp["absent"] > lp["somatic_tumor"]
    assert np.isneginf(lp["artifact"])  # all reads support ref -> no bias is likely (bias/mod.rs:85-93)


def test_threads_match_single_thread():
    sc, b = synth.tumor_normal(40, seed=11)
    flat = sc.flatten()
    a = oracle.call_batch(flat, b, n_threads=1)
    c = oracle.call_batch(flat, b, n_threads=4)
    assert np.array_equal(a.log_posteriors, c.log_posteriors, equal_nan=True)
    assert np.array_equal(a.map_vaf, c.map_vaf, equal_nan=True)
