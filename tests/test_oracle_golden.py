"""Pins the CPU oracle against the reference's own golden pair
tests/resources/flamegraph_profiling/{normal.vcf -> calls.vcf} (fixtures made by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from tests.util import phred
from varlociraptor_b200 import LocusBatch, Scenario


@pytest.fixture(scope="module")
def golden(golden_dir):
    exp = json.load(open(os.path.join(golden_dir, "flamegraph_expected.json")))
    flat = Scenario.from_yaml(exp["scenario_yaml"]).flatten()
    batch = LocusBatch.load(os.path.join(golden_dir, "flamegraph_obs.npz"))
    return exp, flat, batch


def _f32_text(x):
    """bcftools prints f32 with %g (6 significant digits); inf as 'inf'."""
    x = np.float32(x)
    if np.isinf(x):
        return "inf"
    return "%g" % (x + np.float32(0.0))  # normalise -0


def _close_printed(got, want_text):
    if want_text == "inf":
        return np.isinf(got) and got > 0
    want = float(want_text)
    # one unit in the 6th significant digit of the printed value
    tol = 10.0 ** (np.floor(np.log10(max(abs(want), 1e-30))) - 5) * 1.01 if want != 0 else 1e-6
    return abs(float(np.float32(got)) - want) <= tol


def test_golden_pair_all_records_legacy_rules(golden):
    """With the bias-selection rules of the release that wrote calls.vcf (see vlr_oracle.cpp,
    g_legacy_is_likely) every printed number of all 11 records is reproduced."""
    exp, flat, batch = golden
    oracle.set_legacy_is_likely(True)
    try:
        out = oracle.call_batch(flat, batch, afd_capacity=64)
    finally:
        oracle.set_legacy_is_likely(False)
    names = flat.event_names + ["artifact"]
    n_numbers = 0
    for i, rec in enumerate(exp["records"]):
        ph = phred(out.log_posteriors[i])
        for name, val in zip(names, ph):
            want = rec["info"]["PROB_" + name.upper()]
            assert _close_printed(val, want), (rec["pos"], name, val, want)
            n_numbers += 1
        assert abs(out.map_vaf[i, 0] - rec["AF"]) < 1e-6
        vaf, logp = out.afd(i, 0)
        assert len(vaf) == len(rec["AFD"]), rec["pos"]
        for (v, p), (wv, wp) in zip(zip(vaf, phred(logp)), rec["AFD"]):
            assert "%.3f" % v == "%.3f" % wv  # identical adaptive grid abscissae
            assert abs(p - wp) <= 0.0101, (rec["pos"], v, p, wp)  # "%.2f" text, last digit
            n_numbers += 2
    assert n_numbers > 300


def test_golden_pair_head_rules(golden):
    """HEAD semantics (is_uniquely_mapping in Bias::is_likely, MAPQ-only alt-locus bias): the 8 records where
    no artifact twin survives are unchanged; the other 3 are the documented version drift."""
    exp, flat, batch = golden
    out = oracle.call_batch(flat, batch, afd_capacity=64)
    drift = {10471, 10489, 10542}
    for i, rec in enumerate(exp["records"]):
        ph = phred(out.log_posteriors[i])
        if rec["pos"] in drift:
            assert np.isfinite(ph[-1])
            continue
        for name, val in zip(flat.event_names + ["artifact"], ph):
            assert _close_printed(val, rec["info"]["PROB_" + name.upper()]), (rec["pos"], name)


def test_dp_matches(golden):
    """DP = round(sum exp(prob_mapping)) (read_observation.rs:43-47) — checks the codec's prob_mapping column."""
    exp, flat, batch = golden
    for i, rec in enumerate(exp["records"]):
        lo, hi = batch.read_offsets[i], batch.read_offsets[i + 1]
        n = hi - lo
        # the reference reports the pileup size as DP for these records
        assert rec["DP"] in (n, int(round(np.exp(batch.columns["prob_mapping"][lo:hi].astype(np.float64)).sum())))
