"""The reference's own testcase expectations (`expected:` of tests/resources/testcases/*/testcase.yaml, checked upstream
by src/testcase/runner/common/mod.rs:330-395 on the AF field and the PHRED PROB_* fields of the call) evaluated on the
real-data pileups those testcases embed (tests/golden/real_pileups.*, made by tests/golden/make_golden.py).

Upstream the expectations are end-to-end (BAM -> preprocess -> call); here only the calling step runs, on observations a
PREVIOUS preprocessing wrote into the testcase's candidates.vcf. Expectations that hinge on a later change of the
preprocessing cannot hold on those observations; they are listed in PREPROCESSING_DEPENDENT with what the oracle
computes instead (for test_mapq_meth the testcase's own comment states the value for exactly these observations).
Everything else must hold for the oracle and for the engine (emulation and, -m gpu, the CUDA library)."""
import json
import os
import re

import numpy as np
import pytest

from oracle import oracle
from tests.util import phred
from varlociraptor_b200 import LocusBatch, Scenario

# testcase -> (why, the predicate that holds on the embedded observations instead)
PREPROCESSING_DEPENDENT = {
    # testcase.yaml: "If prob_mapping_adj in src/calling/variants/preprocessing/mod.rs: normal > 0.98 && normal < 0.99;
    # if prob_mapping_orig: normal > 0.71 && normal < 0.72" - the embedded record carries the original MAPQs
    "test_mapq_meth": ("MAPQ adjustment happens in preprocessing", "normal > 0.71 && normal < 0.72"),
    # a false negative that was fixed in the realignment (the embedded observations are those of the failing run)
    "test_false_negative_indel_call": ("fixed in the evidence extraction", "sample == 0.0"),
    # a false positive SNV on an insertion, fixed in the evidence extraction (same)
    "test_uzuner_fp_snv_on_ins": ("fixed in the evidence extraction", "sample == 1.0"),
}


def _holds(expr, values):
    """The reference evaluates these with the `eval` crate: identifiers, comparisons, && and ||."""
    py = expr.replace("&&", " and ").replace("||", " or ")
    assert re.fullmatch(r"[\w\s.<>=!()+\-*/]+", py), expr
    return bool(eval(py, {"__builtins__": {}, "inf": float("inf")}, dict(values)))  # noqa: S307 (fixture text, checked above)


def _cases(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "real_pileups.json")))
    allb = LocusBatch.load(os.path.join(golden_dir, "real_pileups.npz"))
    lo = 0
    for tc in meta["testcases"]:
        b = allb.slice(lo, lo + tc["n_loci"])
        lo += tc["n_loci"]
        try:
            sc = Scenario.from_yaml(tc["scenario_yaml"])
        except NotImplementedError:
            continue
        if len(sc.sample_names) != 1:  # the other samples' observations are not embedded
            continue
        yield tc, sc, b


def _check(tc, sc, flat, res):
    """AF of the testcase's sample and PROB_<EVENT> in PHRED as f32, like the fields the reference reads back."""
    values = {sc.sample_names[0]: float(np.float32(res.map_vaf[0, 0]))}
    names = list(flat.event_names) + ["artifact"]
    for k, name in enumerate(names):
        values["PROB_" + name.upper()] = float(np.float32(phred(res.log_posteriors[0, k])))
    exp = tc["expected"]
    exprs = list(exp["allelefreqs"]) + list(exp["posteriors"])
    if tc["testcase"] in PREPROCESSING_DEPENDENT:
        exprs = [PREPROCESSING_DEPENDENT[tc["testcase"]][1]]
    for e in exprs:
        assert _holds(e, values), "%s: %s is false for %s" % (tc["testcase"], e, values)
    return len(exprs)


def test_oracle_meets_the_reference_testcase_expectations(golden_dir):
    n = n_own = 0
    for tc, sc, b in _cases(golden_dir):
        flat = sc.flatten()
        n += _check(tc, sc, flat, oracle.call_batch(flat, b))
        n_own += tc["testcase"] not in PREPROCESSING_DEPENDENT
    assert n >= 10 and n_own >= 7


def test_engine_meets_the_reference_testcase_expectations(golden_dir, engine_call):
    n = 0
    for tc, sc, b in _cases(golden_dir):
        flat = sc.flatten()
        n += _check(tc, sc, flat, engine_call(flat, b))
    assert n >= 10
