"""`estimate contamination` (src/estimation/contamination.rs).

The reference holds no golden vector for this model ("parity unpinned"): the oracle's restatement
(oracle/vlr_oracle.cpp, vlr_oracle_contamination_posterior) is checked here against an independent pure-Python
restatement written from the same source lines, the device functions of csrc/contamination.cuh are checked against the
oracle through the host emulation, and the CUDA path (`-m gpu`) against the oracle through the C-ABI.
Tolerance: 1e-9 absolute on ln posteriors / ln marginal (sums over <= 4096 observations; the chunked device sum and the
reference's sequential sum differ by rounding only), identical -inf / NaN positions, identical max_vaf."""
import bisect
import math

import numpy as np
import pytest

from oracle import oracle
from tests import emu
from varlociraptor_b200 import abi, calling, contamination as ct, synth
from varlociraptor_b200.batch import LocusBatch

TOL = 1e-9


def make_observations(n, seed, with_full_vaf=True, points=(8, 40)):
    """Unimodal AFDs over [0, 1] whose mode is the MAP VAF (what `sample_infos` produces for a de-novo call). With
    `with_full_vaf` one observation has MAP 1.0, so every expected VAF lies on a rising segment (all values finite)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    obs = []
    for i in range(n):
        m = 1.0 if (with_full_vaf and i == n // 2) else float(np.round(rng.uniform(0.05, 0.9), 3))
        k = int(rng.integers(points[0], points[1]))
        grid = np.unique(np.concatenate([[0.0, m, 1.0], np.round(rng.uniform(0.0, 1.0, k), 3)]))
        w = rng.uniform(0.03, 0.2)
        logp = -0.5 * ((grid - m) / w) ** 2 - math.log(w) - rng.uniform(0.0, 3.0)
        if rng.random() < 0.3:
            logp[0] = -np.inf  # density exactly zero at VAF 0
        obs.append(ct.VariantObservation(float(np.log(rng.uniform(0.95, 1.0))), list(zip(grid.tolist(), logp.tolist())),
                                         m, "1", 1000 + i))
    return obs


def oracle_posterior(obs, prior_estimate=None):
    pd, mpv, off, vaf, logp = ct.pack_observations(obs)
    return oracle.contamination_posterior(pd, mpv, off, vaf, logp, ct.Prior(prior_estimate).table())


def python_posterior(obs, prior_estimate=None):
    """contamination.rs:84-115, 163-186, 213-240, 249-271 in plain Python (small inputs only)."""
    def ln_add_exp(a, b):
        if b > a:
            a, b = b, a
        if a == -math.inf:
            return -math.inf
        if b == -math.inf:
            return a
        if math.isnan(a) or math.isnan(b):
            return math.nan
        return a + math.log1p(math.exp(b - a))

    def ln_sum_exp(ps):
        pmax, imax = ps[0], 0
        for i, p in enumerate(ps):  # first maximum; a NaN never wins a `>` comparison but poisons the sum
            if p > pmax:
                pmax, imax = p, i
        if pmax == -math.inf:
            return -math.inf
        if any(math.isnan(p) for p in ps):
            return math.nan
        if pmax == math.inf:
            return math.inf
        return pmax + math.log1p(sum(math.exp(p - pmax) for i, p in enumerate(ps) if i != imax and p != -math.inf))

    def pdf(o, x):
        keys = [v for v, _ in o.vaf_dist]
        j = bisect.bisect_left(keys, x)
        if j < len(keys) and keys[j] == x:
            return o.vaf_dist[j][1]
        if j == 0 or j == len(keys):
            return -math.inf
        (iv, ip), (sv, sp) = o.vaf_dist[j - 1], o.vaf_dist[j]
        with np.errstate(all="ignore"):  # ln of a negative slope is NaN, of zero -inf (f64::ln), not an exception
            return ln_add_exp(ip, float(np.log(np.float64(math.exp(sp) - math.exp(ip)) / (sv - iv)) + np.log(x - iv)))

    max_vaf = max([0.0] + [o.max_posterior_vaf for o in obs])
    prior = ct.Prior(prior_estimate)
    joint = np.empty((4, 101))
    rows = []
    for k, emsv in enumerate(ct.EXPECTED_MAX_SOMATIC_VAFS):
        for i in range(101):
            c = ct.grid_contamination(i)
            purity = 1.0 - c
            lik = 0.0
            for o in obs:
                if purity == 0.0:
                    p = o.prob_denovo
                    with np.errstate(all="ignore"):  # P(denovo) = 1: ln(0) = -inf
                        lik += math.log1p(-math.exp(p)) if p < -0.693 else float(np.log(np.float64(-math.expm1(p))))
                else:
                    lik += pdf(o, emsv * purity * (o.max_posterior_vaf / max_vaf))
            joint[k, i] = prior.prob(c) + lik
        probs = [joint[k, i] + math.log(2 + (i % 2) * 2) for i in range(1, 100)] + [joint[k, 0], joint[k, 100]]
        rows.append(ln_sum_exp(probs) + math.log(1.0) - math.log(100.0) - math.log(3.0))
    marginal = ln_sum_exp(rows)
    with np.errstate(invalid="ignore"):  # -inf - -inf = NaN, as upstream
        return joint - marginal, marginal, max_vaf


def assert_same(got_post, got_marg, want_post, want_marg, tol=TOL):
    assert np.array_equal(np.isnan(got_post), np.isnan(want_post))
    assert np.array_equal(np.isneginf(got_post), np.isneginf(want_post))
    fin = np.isfinite(want_post)
    assert fin.any()
    assert np.max(np.abs(got_post[fin] - want_post[fin])) <= tol
    assert abs(got_marg - want_marg) <= tol


# ------------------------------------------------------------------------------------------ oracle and device functions
def test_oracle_matches_python_restatement():
    obs = make_observations(25, seed=7)
    post, lik, marg, max_vaf = oracle_posterior(obs, ct.PriorEstimate(0.2, 50))
    p_post, p_marg, p_max = python_posterior(obs, ct.PriorEstimate(0.2, 50))
    assert max_vaf == p_max == 1.0
    assert_same(post, marg, p_post, p_marg, tol=1e-12)
    # posteriors are a density over the Simpson rule: integrating them back gives 1
    rows = [np.logaddexp.reduce(np.concatenate([post[k, 1:-1] + np.log(2 + (np.arange(1, 100) % 2) * 2),
                                                post[k, [0, 100]]])) - math.log(300.0) for k in range(4)]
    assert abs(np.logaddexp.reduce(rows)) < 1e-12


@pytest.mark.parametrize("n,chunk", [(1, 1), (40, 7), (700, 37), (3000, 64)])
def test_device_functions_match_oracle(n, chunk):
    obs = make_observations(n, seed=100 + n)
    want_post, want_lik, want_marg, want_max = oracle_posterior(obs)
    got = emu.contamination_posterior(obs, chunk=chunk)
    assert got.max_vaf == want_max == 1.0
    assert_same(got.ln_posterior, got.ln_marginal, want_post, want_marg)
    fin = np.isfinite(want_lik)
    assert np.max(np.abs(got.ln_likelihood[fin] - want_lik[fin])) <= TOL * max(1.0, n / 1000.0)


def test_no_observations_and_prior_only():
    post, lik, marg, max_vaf = oracle_posterior([], ct.PriorEstimate(0.3, 20))
    got = emu.contamination_posterior([], ct.PriorEstimate(0.3, 20))
    assert max_vaf == got.max_vaf == 0.0 and np.all(lik == 0.0) and np.all(got.ln_likelihood == 0.0)
    assert_same(got.ln_posterior, got.ln_marginal, post, marg)
    # binomial prior with k = round(0.3 * 20) = 6: the mode of every row is at contamination 0.3
    assert all(int(np.argmax(post[k])) == 30 for k in range(4))


def test_falling_segment_is_nan_like_the_reference():
    """contamination.rs:98-101 takes `ln` of the density slope: on a falling AFD segment that is ln(negative) = NaN,
    and the NaN reaches the marginal. Reproduced, not repaired."""
    obs = make_observations(12, seed=5, with_full_vaf=False)
    post, _, marg, max_vaf = oracle_posterior(obs)
    got = emu.contamination_posterior(obs)
    assert max_vaf < 1.0 and math.isnan(marg) and math.isnan(got.ln_marginal)
    assert np.isnan(post).all() and np.isnan(got.ln_posterior).all()
    assert np.array_equal(np.isnan(got.ln_likelihood), np.isnan(oracle_posterior(obs)[1]))


# ------------------------------------------------------------------------------------------ host logic
def test_prior_and_number_formatting():
    assert ct.binomial_pdf(0, 0.0, 10) == 1.0 and ct.binomial_pdf(1, 0.0, 10) == 0.0 and ct.binomial_pdf(10, 1.0, 10) == 1.0
    assert abs(ct.binomial_pdf(3, 0.25, 12) - math.comb(12, 3) * 0.25 ** 3 * 0.75 ** 9) < 1e-15
    assert abs(sum(ct.binomial_pdf(k, 0.37, 30) for k in range(31)) - 1.0) < 1e-12
    assert ct.Prior(None).prob(0.4) == 0.0 and ct.Prior(ct.PriorEstimate(0.5, 10)).prob(0.0) == -math.inf
    assert ct._rust_round(2.5) == 3.0 and ct._rust_round(0.3 * 20) == 6.0
    assert [ct._rust_f64(x) for x in (1.0, 0.25, 0.0, 1e-7, 123456789.5, float("nan"))] == \
        ["1", "0.25", "0", "0.0000001", "123456789.5", "NaN"]
    assert ct.grid_contamination(100) == 1.0 and ct.grid_contamination(30) == 30 * 0.01


def _call(prob_denovo, dist, af=0.3):
    c = calling.Call("1", 5, "A", "G")
    c.event_probs = {"absent": -9.0, "denovo": prob_denovo, "other": -9.0, "artifact": -math.inf}
    c.sample_info = [calling.SampleCall(0.0, "none", [(0.0, 0.0)], 20),
                     None if dist == "nomap" else calling.SampleCall(af, "none", dist, 20)]
    return c


def test_variant_observation_selection():
    names = ["contaminant", "sample"]
    keep = ct.VariantObservation.new(_call(math.log(0.97), [(0.5, -1.0), (0.1, -3.0), (0.5, -2.0)]), names)
    assert keep.vaf_dist == [(0.1, -3.0), (0.5, -2.0)] and keep.max_posterior_vaf == 0.3  # BTreeMap: sorted, last wins
    assert ct.VariantObservation.new(_call(math.log(0.94), [(0.5, -1.0)]), names) is None   # P(denovo) < 0.95
    assert ct.VariantObservation.new(_call(math.log(0.99), None), names) is None            # artifact MAP: no AFD
    assert ct.VariantObservation.new(_call(math.log(0.99), "nomap"), names) is None


def _two_sample_records(n_loci, seed, depth=24, homozygous_first=False):
    """Somatic-looking loci of the tumor-normal generator as (contaminant, sample) observation records. With
    `homozygous_first` every sample read of locus 0 supports the alt allele and every contaminant read the reference
    (MAP VAF 1.0, so that the second model stays on rising AFD segments and is finite)."""
    from tests.test_calling_host import _records_from_batch
    _, b = synth.tumor_normal(n_loci, seed=seed, depth=depth)
    if homozygous_first:
        pa, pr = b.columns["prob_alt"], b.columns["prob_ref"]
        lo, mid, hi = (int(x) for x in b.read_offsets[:3])
        hi_v, lo_v = np.maximum(pa[lo:hi], pr[lo:hi]), np.minimum(pa[lo:hi], pr[lo:hi])
        pa[lo:mid], pr[lo:mid] = lo_v[:mid - lo], hi_v[:mid - lo]
        pa[mid:hi], pr[mid:hi] = hi_v[mid - lo:], lo_v[mid - lo:]

    def one(s):
        starts, ends = b.read_offsets[s:-1:2], b.read_offsets[s + 1::2]
        idx = np.concatenate([np.arange(a, e) for a, e in zip(starts, ends)])
        offs = np.concatenate([[0], np.cumsum(ends - starts)])
        return LocusBatch(1, offs, {k: v[idx] for k, v in b.columns.items()}, b.read_flags[idx], b.locus_flags)
    return _records_from_batch(one(0)), _records_from_batch(one(1))  # normal -> contaminant, tumor -> sample


def test_candidate_filter():
    contaminant, sample = _two_sample_records(40, seed=11)
    from varlociraptor_b200 import obs_codec
    f = ct.ContaminationCandidateFilter()
    names = ["contaminant", "sample"]
    kept = 0
    for i, (c, s) in enumerate(zip(contaminant, sample)):
        one = obs_codec.batch_from_records([[c], [s]])
        item = calling.WorkItem(i, "1", c["pos"], "A", "G", one, int(one.locus_flags[0]))
        lo, mid, hi = (int(x) for x in one.read_offsets[:3])
        pa, pr = one.columns["prob_alt"].astype(np.float64), one.columns["prob_ref"].astype(np.float64)
        want = bool(mid - lo >= 10 and np.all(pr[lo:mid] > pa[lo:mid]) and hi - mid >= 10
                    and np.any(np.exp(pa[mid:hi] - pr[mid:hi]) > 20.0))
        assert f.filter(item, names) == want
        kept += want
        item.locus_flags &= ~abi.LF_HAS_SNV  # not an SNV: never kept (contamination.rs:406)
        assert not f.filter(item, names)
    assert 0 < kept < 40


def test_estimate_contamination_end_to_end_with_emulated_engine(tmp_path):
    """call_generic -> filter -> VariantObservations -> second model -> tables, every stage the product's host code;
    the two device stages are replaced by their host emulations."""
    from tests.test_calling_host import EmuEngine
    from varlociraptor_b200 import Scenario
    contaminant, sample = _two_sample_records(30, seed=21, homozygous_first=True)
    sc = Scenario.from_yaml(ct.CONTAMINATION_SCENARIO)
    out, variants, plot = (str(tmp_path / n) for n in ("post.tsv", "maxvaf.csv", "plot.json"))
    est = ct.ContaminationEstimator(out, plot, variants, ct.PriorEstimate(0.25, 40),
                                    posterior_fn=emu.contamination_posterior)
    calling.call_generic(sc, {"sample": sample, "contaminant": contaminant}, call_processor=est,
                         candidate_filter=ct.ContaminationCandidateFilter(), engine=EmuEngine(sc.flatten()),
                         afd_capacity=256)
    assert len(est.variant_observations) >= 3
    for o in est.variant_observations:
        vafs = [v for v, _ in o.vaf_dist]
        assert vafs == sorted(set(vafs)) and math.exp(o.prob_denovo) >= 0.95
    want_post, _, want_marg, want_max = oracle_posterior(est.variant_observations, ct.PriorEstimate(0.25, 40))
    assert est.posterior.max_vaf == want_max == 1.0
    assert_same(est.posterior.ln_posterior, est.posterior.ln_marginal, want_post, want_marg)
    lines = open(out).read().splitlines()
    assert lines[0] == "maximum somatic VAF\tcontamination\tposterior density" and len(lines) == 1 + 404
    dens = [float(l.split("\t")[2]) for l in lines[1:]]
    finite = [d for d in dens if not math.isnan(d)]
    assert finite == sorted(finite, reverse=True)
    assert {l.split("\t")[0] for l in lines[1:]} == {"0.25", "0.5", "0.75", "1"}
    assert open(variants).read().splitlines()[0] == "chrom,pos" and len(open(variants).read().splitlines()) >= 2
    import json
    spec = json.load(open(plot))
    assert sum(b["count"] for b in spec["datasets"]["empirical_vaf_dist"]) == len(est.variant_observations)
    assert len(spec["datasets"]["densities"]) == 101 + 404


# ------------------------------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 33, 3000])
def test_gpu_contamination_matches_oracle(n):
    obs = make_observations(n, seed=300 + n)
    if n == 33:  # one AFD larger than the shared-memory tile: read from global memory by the kernel
        big = np.unique(np.round(np.random.Generator(np.random.PCG64(1)).uniform(0, 1, 6000), 5))
        m = obs[3].max_posterior_vaf
        grid = np.unique(np.concatenate([[0.0, m, 1.0], big]))
        obs[3].vaf_dist = list(zip(grid.tolist(), (-0.5 * ((grid - m) / 0.1) ** 2).tolist()))
        assert len(grid) > 2048
    prior = ct.PriorEstimate(0.1, 30) if n else None
    want_post, want_lik, want_marg, want_max = oracle_posterior(obs, prior)
    got = ct.contamination_posterior(obs, prior, device=0)
    assert got.max_vaf == want_max
    assert_same(got.ln_posterior, got.ln_marginal, want_post, want_marg)
    fin = np.isfinite(want_lik)
    assert np.array_equal(np.isneginf(got.ln_likelihood), np.isneginf(want_lik))
    assert np.max(np.abs(got.ln_likelihood[fin] - want_lik[fin])) <= TOL * max(1.0, n / 1000.0)


@pytest.mark.gpu
def test_gpu_contamination_nan_positions_and_bad_arguments():
    from varlociraptor_b200 import engine
    obs = make_observations(12, seed=5, with_full_vaf=False)
    want_lik = oracle_posterior(obs)[1]
    got = ct.contamination_posterior(obs, device=0)
    assert np.array_equal(np.isnan(got.ln_likelihood), np.isnan(want_lik)) and np.isnan(got.ln_posterior).all()
    with pytest.raises(engine.EngineError):
        ct.contamination_posterior(obs, device=0, n_grid=100)  # Simpson needs an odd grid (rust-bio asserts)
    with pytest.raises(engine.EngineError):
        ct.contamination_posterior(obs, device=99)


def test_random_small_inputs_agree_everywhere():
    """Random tiny observation sets (non-monotone AFDs, -inf densities, P(denovo) = 1, repeated MAP VAFs, expected VAFs
    that hit AFD points exactly or fall outside the support): oracle, pure-Python restatement and the device functions
    agree on every value, including where -inf and NaN appear."""
    rng = np.random.Generator(np.random.PCG64(99))
    grid = np.round(np.linspace(0.0, 1.0, 21), 2)
    for case in range(60):
        obs = []
        for i in range(int(rng.integers(1, 6))):
            k = int(rng.integers(1, 7))
            vafs = np.sort(rng.choice(grid, size=k, replace=False))
            logp = np.round(rng.normal(-2.0, 3.0, k), 3)
            logp[rng.random(k) < 0.15] = -np.inf
            prob_denovo = 0.0 if rng.random() < 0.1 else float(np.log(rng.uniform(0.95, 1.0)))
            obs.append(ct.VariantObservation(prob_denovo, list(zip(vafs.tolist(), logp.tolist())),
                                             float(rng.choice(grid[1:])), "1", i))
        prior = ct.PriorEstimate(float(rng.uniform(0, 1)), int(rng.integers(1, 40))) if rng.random() < 0.5 else None
        post, lik, marg, max_vaf = oracle_posterior(obs, prior)
        p_post, p_marg, p_max = python_posterior(obs, prior)
        got = emu.contamination_posterior(obs, prior, chunk=int(rng.integers(1, 4)))
        assert max_vaf == p_max == got.max_vaf
        for a, b in ((post, p_post), (post, got.ln_posterior)):
            assert np.array_equal(np.isnan(a), np.isnan(b)), case
            assert np.array_equal(np.isneginf(a), np.isneginf(b)) and np.array_equal(np.isposinf(a), np.isposinf(b)), case
            fin = np.isfinite(a)
            assert not fin.any() or np.max(np.abs(a[fin] - b[fin])) <= 1e-9, case
        assert np.array_equal(np.isnan(lik), np.isnan(got.ln_likelihood)), case


@pytest.mark.gpu
def test_device_resident_observations_match_the_call_processor():
    """`estimate contamination` without the calls leaving the device (contamination.rs:44-82 + :163-240 on the GPU):
    vlr_call_batch_device -> vlr_contamination_gather_device -> vlr_contamination_posterior_device. The selection, the
    packed AFDs and the posterior are exactly what the CallProcessor route (host results -> VariantObservation::new per
    call -> vlr_contamination_posterior) produces."""
    import torch
    from varlociraptor_b200 import Scenario, engine
    sc = Scenario.from_yaml(ct.CONTAMINATION_SCENARIO)
    flat = sc.flatten()
    names = list(flat.sample_names)
    s_idx, e_denovo = names.index("sample"), flat.event_names.index("denovo")
    _, b = synth.tumor_normal(700, seed=31, depth=24)  # normal = 0, tumor = 1
    if names.index("contaminant") != 0:  # the batch's sample order follows the scenario's
        off = b.read_offsets
        idx = np.concatenate([np.arange(off[2 * i + 1 - s], off[2 * i + 2 - s]) for i in range(b.n_loci) for s in (0, 1)])
        lens = np.array([off[2 * i + 2 - s] - off[2 * i + 1 - s] for i in range(b.n_loci) for s in (0, 1)])
        b = LocusBatch(2, np.concatenate([[0], np.cumsum(lens)]), {k: v[idx] for k, v in b.columns.items()}, b.read_flags[idx],
                       b.locus_flags)
    cap = 256
    eng = engine.PosteriorEngine(flat)
    host = eng.call_batch(b, afd_capacity=cap)
    obs, kept = [], []
    for i in range(b.n_loci):
        lp = float(host.log_posteriors[i, e_denovo])
        if (host.status[i] & abi.ST_NO_MAP) or host.map_config[i] != 0 or math.exp(lp) < 0.95:
            continue
        v, p = host.afd(i, s_idx)
        obs.append(ct.VariantObservation(lp, list(zip(v.tolist(), p.tolist())), float(host.map_vaf[i, s_idx]), "1", i))
        kept.append(i)
    assert 10 < len(obs) < b.n_loci
    prior = ct.PriorEstimate(0.2, 25)
    want = ct.contamination_posterior(obs, prior, device=0)

    db = engine.DeviceBatch(b)
    dr = engine.DeviceResults(b.n_loci, 2, flat.n_events, cap)
    eng.call_batch_device(db, dr)
    torch.cuda.synchronize()
    dobs = ct.DeviceObservations(dr, s_idx, e_denovo)
    assert dobs.n_obs == len(obs)
    assert dobs.kept_loci[:dobs.n_obs].cpu().tolist() == kept
    offs = dobs.afd_offsets[:dobs.n_obs + 1].cpu().numpy()
    assert np.array_equal(np.diff(offs), [len(o.vafs) for o in obs])
    assert np.array_equal(dobs.afd_vaf[:offs[-1]].cpu().numpy(), np.concatenate([o.vafs for o in obs]))
    assert np.array_equal(dobs.afd_logp[:offs[-1]].cpu().numpy(), np.concatenate([o.densities for o in obs]))
    got = ct.contamination_posterior_device(dobs, prior)
    assert got.max_vaf == want.max_vaf and (got.ln_marginal == want.ln_marginal or (math.isnan(got.ln_marginal) and math.isnan(want.ln_marginal)))
    assert np.array_equal(got.ln_posterior, want.ln_posterior, equal_nan=True)
    assert np.array_equal(got.ln_likelihood, want.ln_likelihood, equal_nan=True)
    # nothing passes a threshold of 1.0 + epsilon: an empty observation set is valid
    none = ct.DeviceObservations(dr, s_idx, e_denovo, min_prob=1.5)
    assert none.n_obs == 0 and np.isfinite(ct.contamination_posterior_device(none, None).ln_marginal)
