"""Parity of the CUDA engine (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerance: 1e-9 absolute in log space (BASELINE.json north_star) for event posteriors, marginal and AFD;
MAP VAFs, best events, status bits and AFD abscissae must be identical. Loci the oracle marks as knife-edge
(a discrete decision of the reference algorithm within rounding noise of flipping: DESIGN.md §5) are reported
and excluded from the value comparison; they are bounded to a small fraction.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from tests.util import FOUR_SAMPLE_YAML, batch_from_reads, four_sample_batch, max_abs_delta, pair_as_tumor_normal, read
from varlociraptor_b200 import LocusBatch, Scenario, abi, synth

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def engine_mod():
    from varlociraptor_b200 import engine
    return engine


def _compare(o, g, afd=True, max_knife_fraction=0.02):
    ke = o.knife_edge()
    assert ke.mean() <= max_knife_fraction, "too many knife-edge loci: %g" % ke.mean()
    ok = ~ke
    assert max_abs_delta(o.log_posteriors[ok], g.log_posteriors[ok]) <= TOL
    assert max_abs_delta(o.log_marginal[ok], g.log_marginal[ok]) <= TOL * 10  # marginal ~ -1e3: relative 1e-14
    # MAP allele frequencies are visited abscissae: identical, except where two visited abscissae are one ulp apart (a
    # closing linspace point next to a dyadic bracket point: lo + k * step rounds differently from (l + r) / 2) and carry
    # the same likelihood to the last bits - then the order of the factors in the pileup product decides which of the
    # two is "the first maximum" (config 5, locus 297: 0x1.aa00000000000p-8 vs 0x1.aa00000000001p-8)
    with np.errstate(invalid="ignore"):
        d_map = np.abs(o.map_vaf[ok] - g.map_vaf[ok])
    same = (o.map_vaf[ok] == g.map_vaf[ok]) | (np.isnan(o.map_vaf[ok]) & np.isnan(g.map_vaf[ok]))
    assert np.all(same | (d_map <= 2 * np.spacing(np.abs(o.map_vaf[ok]))))
    assert np.count_nonzero(~same) <= max(1, int(2e-3 * same.size))
    assert np.array_equal(o.best_event[ok], g.best_event[ok])
    assert np.array_equal(o.map_config[ok], g.map_config[ok])
    assert np.array_equal(o.status[ok], g.status[ok])
    assert np.array_equal(o.n_base_events[ok], g.n_base_events[ok])  # identical adaptive grids
    if afd and o.afd_capacity:
        assert np.array_equal(o.afd_count[ok], g.afd_count[ok])
        valid = np.arange(o.afd_capacity)[None, None, :] < o.afd_count[:, :, None]  # entries past the count are unspecified
        valid &= ok[:, None, None]
        assert max_abs_delta(o.afd_vaf[valid], g.afd_vaf[valid]) == 0.0
        assert max_abs_delta(o.afd_logp[valid], g.afd_logp[valid]) <= TOL
    return ke


def test_tumor_normal_config2_sample(engine_mod):
    sc, b = synth.tumor_normal(1500, seed=synth.SEED_BASE + 2)
    flat = sc.flatten()
    o = oracle.call_batch(flat, b, afd_capacity=96, n_threads=os.cpu_count() or 1)
    eng = engine_mod.PosteriorEngine(flat)
    g = eng.call_batch(b, afd_capacity=96)
    _compare(o, g)
    assert eng.launches >= 1


def test_pedigree_config3_sample(engine_mod):
    sc, b = synth.pedigree(3000, seed=synth.SEED_BASE + 3)
    flat = sc.flatten()
    o = oracle.call_batch(flat, b, afd_capacity=8, n_threads=os.cpu_count() or 1)
    g = engine_mod.PosteriorEngine(flat).call_batch(b, afd_capacity=8)
    _compare(o, g)


def test_depth_skew_config5_sample(engine_mod):
    sc, b = synth.tumor_normal(2000, seed=synth.SEED_BASE + 5, depth_range=(10, 2000))
    flat = sc.flatten()
    o = oracle.call_batch(flat, b, afd_capacity=0, n_threads=os.cpu_count() or 1)
    g = engine_mod.PosteriorEngine(flat).call_batch(b)
    _compare(o, g, afd=False, max_knife_fraction=0.05)


def test_golden_flamegraph(engine_mod, golden_dir):
    exp = json.load(open(os.path.join(golden_dir, "flamegraph_expected.json")))
    flat = Scenario.from_yaml(exp["scenario_yaml"]).flatten()
    b = LocusBatch.load(os.path.join(golden_dir, "flamegraph_obs.npz"))
    o = oracle.call_batch(flat, b, afd_capacity=64)
    g = engine_mod.PosteriorEngine(flat).call_batch(b, afd_capacity=64)
    _compare(o, g)
    # and directly against what the reference printed, for the records unaffected by the version drift
    ph = -10.0 * g.log_posteriors / np.log(10.0)
    for i, rec in enumerate(exp["records"]):
        if rec["pos"] in (10471, 10489, 10542):
            continue
        assert abs(np.float32(ph[i, 0]) - float(rec["info"]["PROB_ABSENT"])) <= 2e-3 * max(1.0, ph[i, 0] / 300)
        assert g.map_vaf[i, 0] == rec["AF"]


def test_real_pileups_single_sample(engine_mod, golden_dir):
    """Real-data pileups (depth 2..2991, f16/f32 quantised) under each testcase's own scenario."""
    meta = json.load(open(os.path.join(golden_dir, "real_pileups.json")))
    allb = LocusBatch.load(os.path.join(golden_dir, "real_pileups.npz"))
    lo = 0
    n_checked = n_knife = 0
    for tc in meta["testcases"]:
        b = allb.slice(lo, lo + tc["n_loci"])
        lo += tc["n_loci"]
        try:
            sc = Scenario.from_yaml(tc["scenario_yaml"])
        except NotImplementedError:
            continue
        if len(sc.sample_names) != 1:
            continue
        flat = sc.flatten()
        o = oracle.call_batch(flat, b, afd_capacity=128)
        g = engine_mod.PosteriorEngine(flat).call_batch(b, afd_capacity=128)
        n_knife += int(_compare(o, g, max_knife_fraction=1.0).sum())  # one locus per testcase: bounded over all of them
        n_checked += 1
    assert n_checked >= 5
    assert n_knife <= 1, "knife-edge loci among the real pileups: %d of %d" % (n_knife, n_checked)  # oracle: exactly one


def test_edge_cases(engine_mod):
    """Empty pileups, a single read, all-reference, singleton alt read, prob_mapping = -inf, vaf 1 bypass."""
    flat = Scenario.tumor_normal(0.75).flatten()
    ref = dict(prob_alt=np.log(1e-3 / 3), prob_ref=np.log1p(-1e-3), prob_mapping=np.log1p(-1e-6))
    alt = dict(prob_ref=np.log(1e-3 / 3), prob_alt=np.log1p(-1e-3), prob_mapping=np.log1p(-1e-6))
    mk = lambda d, i: read(strand=i % 2, orientation=i % 2, prob_double_overlap=-np.inf, **d)  # noqa: E731
    loci = [
        [[], []],                                                     # no observations at all
        [[mk(ref, 0)], []],                                           # one read, empty tumor
        [[mk(ref, i) for i in range(30)], [mk(ref, i) for i in range(30)]],   # clear reference
        [[mk(ref, i) for i in range(30)], [mk(ref, i) for i in range(29)] + [mk(alt, 0)]],  # singleton evidence
        [[mk(ref, i) for i in range(12)], [mk(alt, i) for i in range(12)]],   # tumor all alt
        [[mk(alt, i) for i in range(12)], [mk(alt, i) for i in range(12)]],   # germline hom
        [[read(prob_mapping=-np.inf, prob_alt=-1.0, prob_ref=-2.0, strand=0, orientation=0) for _ in range(6)],
         [mk(alt, i) for i in range(3)]],                             # unmappable reads, n_obs < 5 (Simpson 11)
        [[mk(ref, i) for i in range(8)] + [read(orientation=abi.ORIENT_F1F2, **alt)],  # filtered non-standard read
         [mk(alt, i) for i in range(4)] + [mk(ref, i) for i in range(4)]],
    ]
    b = batch_from_reads(loci)
    o = oracle.call_batch(flat, b, afd_capacity=128)
    g = engine_mod.PosteriorEngine(flat).call_batch(b, afd_capacity=128)
    ke = _compare(o, g, max_knife_fraction=1.0 / 8)  # locus 7: an exact tie of the adaptive argmax (identical reads)
    assert not ke[:7].any()
    assert g.status[3] & abi.ST_SINGLETON_ADJUSTED
    assert g.status[7] & abi.ST_FILTERED_NONSTANDARD


def test_device_pointer_entry_matches_host_entry(engine_mod):
    import torch
    sc, b = synth.tumor_normal(400, seed=99)
    flat = sc.flatten()
    eng = engine_mod.PosteriorEngine(flat)
    host = eng.call_batch(b, afd_capacity=64)
    db = engine_mod.DeviceBatch(b)
    dr = engine_mod.DeviceResults(b.n_loci, 2, flat.n_events, 64)
    eng.call_batch_device(db, dr)
    torch.cuda.synchronize()
    dev = dr.to_host()
    assert np.array_equal(host.log_posteriors, dev.log_posteriors, equal_nan=True)
    assert np.array_equal(host.map_vaf, dev.map_vaf, equal_nan=True)
    assert np.array_equal(host.afd_logp[host.afd_count > 0], dev.afd_logp[dev.afd_count > 0], equal_nan=True)


def test_workspace_overflow_is_reported_not_silent(engine_mod, monkeypatch):
    import torch
    sc, b = synth.tumor_normal(4, seed=5, depth=3000)  # 6000 reads/locus > default reserve of 4096
    flat = sc.flatten()
    # the wavefront pipeline sizes its coefficient arena from the batch and never needs the reserve; the generic
    # engine (VLR_WAVE=0, and every scenario the pipeline does not serve) does
    wave = engine_mod.PosteriorEngine(flat)
    dbw = engine_mod.DeviceBatch(b)
    drw = engine_mod.DeviceResults(b.n_loci, 2, flat.n_events)
    wave.call_batch_device(dbw, drw)
    torch.cuda.synchronize()
    assert not np.any(drw.to_host().status & abi.ST_WORKSPACE_OVERFLOW)
    assert max_abs_delta(oracle.call_batch(flat, b).log_posteriors, drw.to_host().log_posteriors) <= TOL
    monkeypatch.setenv("VLR_WAVE", "0")
    eng = engine_mod.PosteriorEngine(flat)
    monkeypatch.delenv("VLR_WAVE")
    db = engine_mod.DeviceBatch(b)
    dr = engine_mod.DeviceResults(b.n_loci, 2, flat.n_events)
    eng.call_batch_device(db, dr)
    torch.cuda.synchronize()
    assert np.all(dr.to_host().status & abi.ST_WORKSPACE_OVERFLOW)
    eng.reserve(6000)
    eng.call_batch_device(db, dr)
    torch.cuda.synchronize()
    got = dr.to_host()
    assert not np.any(got.status & abi.ST_WORKSPACE_OVERFLOW)
    o = oracle.call_batch(flat, b)
    assert max_abs_delta(o.log_posteriors, got.log_posteriors) <= TOL


def test_chunked_host_path_is_order_preserving(engine_mod):
    """More loci than one transfer chunk (65536 loci / 16M reads): results come back in input order across chunks."""
    sc, b = synth.tumor_normal(70000, seed=123, depth=20)  # 2 chunks
    flat = sc.flatten()
    eng = engine_mod.PosteriorEngine(flat)
    g = eng.call_batch(b)
    assert eng.launches >= 2
    idx = np.r_[0:50, 65500:65560, 69950:70000]
    o = oracle.call_batch(flat, b.select(idx), n_threads=os.cpu_count() or 1)
    ke = o.knife_edge()
    assert max_abs_delta(o.log_posteriors[~ke], g.log_posteriors[idx][~ke]) <= TOL
    total = np.logaddexp.reduce(g.log_posteriors, axis=1)
    assert np.nanmax(np.abs(total)) < 1e-9  # size-independent property: posteriors sum to one


def test_call_generic_reproduces_reference_text_fields(engine_mod, golden_dir):
    """The operator API end to end on the GPU: scenario + observation records in, final-record text fields out,
    compared with what the reference wrote to calls.vcf (records outside the documented version drift)."""
    from tests.test_calling_host import _records_from_batch
    from varlociraptor_b200 import calling
    exp = json.load(open(os.path.join(golden_dir, "flamegraph_expected.json")))
    b = LocusBatch.load(os.path.join(golden_dir, "flamegraph_obs.npz"))
    recs = _records_from_batch(b)
    for r, e in zip(recs, exp["records"]):
        r["pos"], r["ref"], r["alt"] = e["pos"], "CG", "<METH>"
    w = calling.call_generic(Scenario.from_yaml(exp["scenario_yaml"]), {"normal": recs})
    n = 0
    for c, e in zip(w.calls, exp["records"]):
        if e["pos"] in (10471, 10489, 10542):
            continue
        info = c.info_fields()
        for tag, text in e["info"].items():
            if text == "inf":
                assert np.isinf(info[tag])
            else:
                assert abs(float("%g" % info[tag]) - float(text)) <= 10 ** (np.floor(np.log10(max(abs(float(text)), 1e-9))) - 4)
        f = c.format_fields(0)
        assert f["AF"] == "%g" % e["AF"] and int(f["DP"]) == e["DP"]
        got = [tuple(map(float, kv.split("="))) for kv in f["AFD"].split(",")]
        assert [g[0] for g in got] == [x[0] for x in e["AFD"]]
        assert max(abs(g[1] - x[1]) for g, x in zip(got, e["AFD"])) <= 0.0101
        n += 1
    assert n == 8


def test_config1_real_pileups_paired_as_tumor_normal(engine_mod, golden_dir):
    """BASELINE config 1 (plumbing): ~100 tumor-normal loci built by pairing the real-data pileups embedded in the
    reference's testcases (depth 2..2991, indels with prob_sample_alt < 0, homopolymer columns, f16/f32 quantised)."""
    single = LocusBatch.load(os.path.join(golden_dir, "real_pileups.npz"))
    n = single.n_loci
    pairs = [(i, j) for i in range(n) for j in range(n) if i != j]
    b = pair_as_tumor_normal(single, pairs)
    assert b.n_loci >= 100
    flat = Scenario.tumor_normal(0.8).flatten()
    o = oracle.call_batch(flat, b, afd_capacity=128, n_threads=os.cpu_count() or 1)
    g = engine_mod.PosteriorEngine(flat).call_batch(b, afd_capacity=128)
    _compare(o, g, max_knife_fraction=0.1)


def test_four_samples_full_capacity_variant(engine_mod):
    """> 3 samples: served by the full-capacity kernel variant (vlr_call_kernel_vlr_full)."""
    flat = Scenario.from_yaml(FOUR_SAMPLE_YAML).flatten()
    b = four_sample_batch(60, seed=51)
    o = oracle.call_batch(flat, b, afd_capacity=128, n_threads=os.cpu_count() or 1)
    g = engine_mod.PosteriorEngine(flat).call_batch(b, afd_capacity=128)
    _compare(o, g, max_knife_fraction=0.1)


def test_wavefront_pipeline_equals_generic_engine(engine_mod, monkeypatch):
    """Tumor-normal scenarios run the wavefront pipeline (engine_wave.cuh); VLR_WAVE=0 forces the generic warp-per-locus
    engine. Same loci, both against the oracle and against each other, with and without AFD, several sub-chunks."""
    sc, b = synth.tumor_normal(20000, seed=4242, depth=40)
    flat = sc.flatten()
    wave = engine_mod.PosteriorEngine(flat)
    monkeypatch.setenv("VLR_WAVE", "0")
    generic = engine_mod.PosteriorEngine(flat)
    monkeypatch.delenv("VLR_WAVE")
    for afd in (0, 48):
        gw = wave.call_batch(b, afd_capacity=afd)
        gg = generic.call_batch(b, afd_capacity=afd)
        assert wave.launches > generic.launches
        assert np.array_equal(gw.status, gg.status)
        assert np.array_equal(gw.best_event, gg.best_event)
        same_grid = gw.n_base_events == gg.n_base_events
        assert same_grid.mean() > 0.995  # different product order: knife-edge loci may take another grid
        assert max_abs_delta(gw.log_posteriors[same_grid], gg.log_posteriors[same_grid]) <= TOL
        assert max_abs_delta(gw.map_vaf[same_grid], gg.map_vaf[same_grid]) == 0.0
        if afd:
            assert np.array_equal(gw.afd_count[same_grid], gg.afd_count[same_grid])
    idx = np.r_[0:300, 19700:20000]
    o = oracle.call_batch(flat, b.select(idx), afd_capacity=48, n_threads=os.cpu_count() or 1)
    ke = o.knife_edge()
    gw = wave.call_batch(b, afd_capacity=48)
    assert max_abs_delta(o.log_posteriors[~ke], gw.log_posteriors[idx][~ke]) <= TOL
    assert np.array_equal(o.n_base_events[~ke], gw.n_base_events[idx][~ke])
    valid = np.arange(48)[None, None, :] < o.afd_count[:, :, None]
    valid &= (~ke)[:, None, None]
    assert np.array_equal(o.afd_count[~ke], gw.afd_count[idx][~ke])
    assert max_abs_delta(o.afd_logp[valid], gw.afd_logp[idx][valid]) <= TOL


def test_large_device_batch_runs_on_internal_streams(engine_mod):
    """>= 131072 loci through the device entry: the wavefront pipeline splits the batch over internal streams (forked
    from / joined to the caller's stream). Bitwise the same results as the chunked host entry, in input order."""
    import torch
    sc, b = synth.tumor_normal(140000, seed=777, depth=12)
    flat = sc.flatten()
    eng = engine_mod.PosteriorEngine(flat)
    host = eng.call_batch(b)
    db = engine_mod.DeviceBatch(b)
    dr = engine_mod.DeviceResults(b.n_loci, 2, flat.n_events)
    s = torch.cuda.Stream()
    eng.call_batch_device(db, dr, s.cuda_stream)
    s.synchronize()
    dev = dr.to_host()
    assert np.array_equal(host.log_posteriors, dev.log_posteriors, equal_nan=True)
    assert np.array_equal(host.map_vaf, dev.map_vaf, equal_nan=True)
    assert np.array_equal(host.status, dev.status)
    assert np.array_equal(host.n_base_events, dev.n_base_events)
    total = np.logaddexp.reduce(dev.log_posteriors, axis=1)
    assert np.nanmax(np.abs(total)) < 1e-9
