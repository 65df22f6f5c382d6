"""The reference's six prior scenarios (tests/resources/prior/scenarios, committed verbatim in
tests/golden/prior_scenarios.json): population heterozygosity, Mendelian inheritance with sex- and contig-specific
ploidy, clonal / subclonal inheritance and somatic mutation rates (src/variants/model/prior.rs:298-678). Upstream
only plots them; here every one goes through scenario front-end -> flattened trees -> engine (host emulation of the
kernel source, and the CUDA library with -m gpu) and is compared with the oracle, on autosomes and on the sex chromosomes where ploidies differ."""
import json
import os

import pytest

from oracle import oracle
from tests.test_emu_parity import _compare
from tests.util import four_sample_batch
from varlociraptor_b200 import Scenario, synth

CASES = [("population", "all"), ("tumor-normal", "all"), ("tumor-relapse", "all"), ("tumor-normal-relapse", "all"),
         ("simple-pedigree", "all"), ("pedigree", "all"), ("pedigree", "X"), ("pedigree", "Y")]


def _batch(n_samples, n_loci, seed):
    if n_samples == 2:
        return synth.tumor_normal(n_loci, seed=seed, depth=30)[1]
    if n_samples == 3:
        return synth.pedigree(n_loci, seed=seed, depth=30)[1]
    return four_sample_batch(n_loci, seed=seed, depth=16)


@pytest.mark.parametrize("name,contig", CASES)
def test_prior_scenario_engine_matches_oracle(engine_call, golden_dir, name, contig):
    text = json.load(open(os.path.join(golden_dir, "prior_scenarios.json")))["scenarios"][name]
    if engine_call.kind == "cuda" and name == "tumor-normal-relapse":
        # three nested full-range integrations with a per-point prior: ~1e6 joint evaluations x ~1e4 instructions of
        # prior per locus, and a locus is one warp of the generic engine: minutes per call for a handful of loci (8 min
        # measured for this test). The control flow is covered by the host emulation, the same branches on the device
        # by the two-sample relapse scenario and the fuzzers.
        pytest.skip("minutes on the device: one warp per locus, ~1e10 instructions per locus")
    for full_prior in (False, True):
        sc = Scenario.from_yaml(text, full_prior=full_prior).for_contig(contig)
        flat = sc.flatten()
        b = _batch(flat.n_samples, 16, seed=70 + len(name))
        # the relapse scenarios integrate two full ranges inside each other: more base events than the engine's log
        # holds (4096 per locus), so those loci take the second, filtered pass of process_locus (engine_core.cuh)
        afd = 64
        want = oracle.call_batch(flat, b, afd_capacity=afd, n_threads=4)
        _compare(want, engine_call(flat, b, afd_capacity=afd))
        assert not (want.status & (1 << 1 | 1 << 2 | 1 << 3)).any()  # no NaN / overshoot / positive prior


def test_ploidies_of_the_pedigree_on_sex_chromosomes(golden_dir):
    text = json.load(open(os.path.join(golden_dir, "prior_scenarios.json")))["scenarios"]["pedigree"]
    sc = Scenario.from_yaml(text)
    assert [sc.for_contig("X").ploidy(n) for n in ("mother", "father", "child", "sibling")] == [2, 1, 1, 2]
    assert [sc.for_contig("Y").ploidy(n) for n in ("mother", "father", "child", "sibling")] == [0, 1, 1, 0]
    y = sc.for_contig("Y")
    assert y.universe("mother") == [frozenset([0.0])] and y.universe("father") == [frozenset([0.0, 1.0])]
