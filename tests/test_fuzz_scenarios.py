"""Random scenarios (2-3 samples, set and range universes, nested and/or/not formulas) through the front-end, the
engine's host emulation and the oracle: tree shapes no hand-written test covers (many roots after the DNF step, true
nodes, branches that re-add missing samples). Posteriors, best events, status bits and the number of joint evaluations
must agree, and so must the MAP allele frequencies although random events overlap: upstream keeps one global map of
base events and reports the best one the strongest event contains, whichever event evaluated it. Every test runs on
the host emulation and (-m gpu) on the CUDA library."""
import random

import numpy as np
import pytest

from oracle import oracle
from tests.util import max_abs_delta
from varlociraptor_b200 import Scenario, synth

POINTS = [0.0, 0.25, 0.5, 0.75, 1.0]


def _spectrum(rng, kind):
    if kind == "set" or rng.random() < 0.5:
        k = rng.randint(1, 2)
        return "{%s}" % ",".join(str(x) for x in sorted(rng.sample(POINTS, k))) if k > 1 else str(rng.choice(POINTS))
    lo, hi = sorted(rng.sample(POINTS, 2))
    return "%s%s,%s%s" % (rng.choice("[]"), lo, hi, rng.choice("[]"))


def _formula(rng, samples, kinds, depth=0):
    if depth >= 2 or rng.random() < 0.35:
        s = rng.choice(samples)
        text = "%s:%s" % (s, _spectrum(rng, kinds[s]))
        return "!" + text if rng.random() < 0.15 else text
    op = " & " if rng.random() < 0.6 else " | "
    return "(" + op.join(_formula(rng, samples, kinds, depth + 1) for _ in range(rng.randint(2, 3))) + ")"


def random_scenario(seed):
    rng = random.Random(seed)
    samples = ["s%d" % i for i in range(rng.choice([2, 3]))]
    kinds = {s: rng.choice(["set", "range"]) for s in samples}
    text = "samples:\n"
    for s in samples:
        text += '  %s:\n    resolution: %s\n    universe: "%s"\n' % (
            s, rng.choice([0.1, 0.05]), "0.0 | 0.25 | 0.5 | 0.75 | 1.0" if kinds[s] == "set" else "[0.0,1.0]")
    text += "events:\n"
    for e in range(rng.randint(1, 3)):
        text += '  e%d: "%s"\n' % (e, _formula(rng, samples, kinds))
    return text, len(samples)


@pytest.mark.parametrize("seed", list(range(120)))
def test_random_scenario(engine_call, seed):
    text, n_samples = random_scenario(seed)
    try:
        flat = Scenario.from_yaml(text).flatten()
    except ValueError as err:  # the generator may produce events the reference rejects as well
        assert "not disjunct" in str(err)
        return
    gen = synth.tumor_normal if n_samples == 2 else synth.pedigree
    b = gen(5, seed=seed, depth=16)[1]
    want = oracle.call_batch(flat, b, n_threads=4)
    got = engine_call(flat, b)
    ok = ~want.knife_edge()
    assert max_abs_delta(want.log_posteriors[ok], got.log_posteriors[ok]) <= 1e-9
    assert np.array_equal(want.best_event[ok], got.best_event[ok])
    assert np.array_equal(want.n_base_events[ok], got.n_base_events[ok])
    assert np.array_equal(want.status[ok], got.status[ok])
    # random events overlap: the MAP is the best base event ANY event recorded that the strongest event contains
    # (calling.rs:851-864), which the engine reproduces by offering every base event to every event (Ctx::map_global)
    assert np.array_equal(want.map_vaf[ok], got.map_vaf[ok], equal_nan=True)
    assert np.array_equal(want.map_config[ok], got.map_config[ok])


def random_prior_scenario(seed):
    """Species / sample prior settings at random: ploidy 1-3, heterozygosity, germline and somatic rates, Mendelian,
    clonal and subclonal inheritance, contamination (src/variants/model/prior.rs:298-678)."""
    rng = random.Random(seed)
    names = ["a", "b", "c"][:rng.choice([2, 3])]
    ploidy = rng.choice([1, 2, 2, 2, 3])
    text = "species:\n  heterozygosity: %g\n  ploidy: %d\n" % (rng.choice([1e-3, 1e-2, 5e-4]), ploidy)
    if rng.random() < 0.5:
        text += "  germline-mutation-rate: %g\n" % rng.choice([1e-3, 1e-4])
    if rng.random() < 0.3:
        text += "  somatic-effective-mutation-rate: %g\n" % rng.choice([1e-6, 1e-5])
    text += "samples:\n"
    somatic = {}
    for i, n in enumerate(names):
        text += "  %s:\n    resolution: %s\n" % (n, rng.choice([0.1, 0.05]))
        somatic[n] = rng.random() < 0.4
        if somatic[n]:
            text += "    somatic-effective-mutation-rate: %g\n" % rng.choice([1e-6, 1e-4, 1e-10])
        if i > 0 and rng.random() < 0.6:
            r = rng.random()
            if len(names) == 3 and i == 2 and r < 0.4:
                text += "    inheritance:\n      mendelian:\n        from:\n          - a\n          - b\n"
            elif r < 0.7:
                text += "    inheritance:\n      clonal:\n        from: %s\n        somatic: %s\n" % (
                    names[rng.randrange(i)], rng.choice(["true", "false"]))
            else:
                text += "    inheritance:\n      subclonal:\n        from: %s\n" % names[rng.randrange(i)]
        if i > 0 and rng.random() < 0.25:
            text += "    contamination:\n      by: a\n      fraction: %s\n" % rng.choice([0.1, 0.3])
    points = [k / ploidy for k in range(ploidy + 1)]
    events = ['  het: "a:%s"' % points[1]]
    if ploidy >= 2:
        events.append('  hom: "a:%s"' % points[-1])
    if somatic[names[-1]]:
        events.append('  som: "a:0.0 & %s:]0.0,%s["' % (names[-1], points[1]))
    return text + "events:\n" + "\n".join(events) + "\n", len(names)


@pytest.mark.parametrize("seed", range(40))
def test_random_prior_configuration(engine_call, seed):
    """Everything must agree here (events are disjoint), MAP allele frequencies included. Loci whose prior raises an
    invariant violation (status bits NaN / overshoot / prior > 0: the reference panics, e.g. Mendelian inheritance
    with mismatching ploidies, prior.rs:672-676) only need the same status. 1500 configurations were explored."""
    text, n_samples = random_prior_scenario(seed)
    gen = synth.tumor_normal if n_samples == 2 else synth.pedigree
    b = gen(5, seed=seed, depth=16)[1]
    for full_prior in (False, True):
        flat = Scenario.from_yaml(text, full_prior=full_prior).flatten()
        want = oracle.call_batch(flat, b, n_threads=4)
        got = engine_call(flat, b)
        assert np.array_equal(want.status, got.status)
        ok = ~want.knife_edge() & ((want.status & 0xe) == 0)
        assert max_abs_delta(want.log_posteriors[ok], got.log_posteriors[ok]) <= 1e-9
        assert np.array_equal(want.best_event[ok], got.best_event[ok])
        assert np.array_equal(want.n_base_events[ok], got.n_base_events[ok])
        assert np.array_equal(want.map_vaf[ok], got.map_vaf[ok], equal_nan=True)
