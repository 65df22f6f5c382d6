"""Engine against the oracle, every test twice (fixture `engine_call`, tests/conftest.py): the single-lane host build of
engine_core.cuh (tests/emu, test infrastructure only: the engine's control flow without a GPU) and the CUDA library
through the C-ABI (-m gpu). More CUDA-only parity tests are in tests/test_gpu_parity.py."""
import json
import os

import numpy as np

from oracle import oracle
from tests import emu
from tests.util import FOUR_SAMPLE_YAML, batch_from_reads, four_sample_batch, max_abs_delta, pair_as_tumor_normal, read
from varlociraptor_b200 import LocusBatch, Scenario, abi, synth

TOL = 1e-9


def _compare(o, g, ok=None):
    ok = ~o.knife_edge() if ok is None else ok
    assert max_abs_delta(o.log_posteriors[ok], g.log_posteriors[ok]) <= TOL
    assert max_abs_delta(o.map_vaf[ok], g.map_vaf[ok]) == 0.0
    assert np.array_equal(o.best_event[ok], g.best_event[ok])
    assert np.array_equal(o.map_config[ok], g.map_config[ok])
    assert np.array_equal(o.status[ok], g.status[ok])
    assert np.array_equal(o.n_base_events[ok], g.n_base_events[ok])
    if o.afd_capacity:
        assert np.array_equal(o.afd_count[ok], g.afd_count[ok])
        valid = np.arange(o.afd_capacity)[None, None, :] < o.afd_count[:, :, None]  # entries past the count are unspecified
        valid &= ok[:, None, None]
        assert max_abs_delta(o.afd_vaf[valid], g.afd_vaf[valid]) == 0.0
        assert max_abs_delta(o.afd_logp[valid], g.afd_logp[valid]) <= TOL


def test_tumor_normal(engine_call):
    sc, b = synth.tumor_normal(150, seed=5)
    flat = sc.flatten()
    _compare(oracle.call_batch(flat, b, afd_capacity=96, n_threads=4), engine_call(flat, b, afd_capacity=96))


def test_pedigree_mixed_snv_indel(engine_call):
    sc, b = synth.pedigree(300, seed=6)
    flat = sc.flatten()
    _compare(oracle.call_batch(flat, b, afd_capacity=8, n_threads=4), engine_call(flat, b, afd_capacity=8))


def test_depth_skew(engine_call):
    sc, b = synth.tumor_normal(24, seed=8, depth_range=(10, 2000))
    flat = sc.flatten()
    _compare(oracle.call_batch(flat, b, n_threads=4), engine_call(flat, b))


def _tn_with_depths(depths, seed):
    """Tumor-normal SNV loci of the synthetic generator with the given (normal, tumor) depths."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    depths = np.asarray(depths, dtype=np.int64)
    n = len(depths)
    cls = rng.choice(5, size=n, p=[0.3, 0.3, 0.2, 0.1, 0.1])
    t = np.where(cls == 0, 0.0, np.where(cls == 2, 0.5, np.where(cls == 3, 1.0, rng.uniform(0.05, 0.6, n))))
    nn = np.where(cls == 2, 0.5, np.where(cls == 3, 1.0, np.where(cls == 4, rng.uniform(0.05, 0.3, n), 0.0)))
    eff = np.stack([nn, 0.75 * t + 0.25 * nn], axis=1)
    return Scenario.tumor_normal(purity=0.75), synth._assemble(2, depths, eff, rng, synth._snv_locus_flags(rng, n))


def test_resident_size_classes_at_their_boundaries(engine_call):
    """The lc-resident round kernels serve an lc by the deeper of its two pileups: an octet per lc up to 110 reads,
    a warp per lc with slots of 104 / 208 / 416 polynomials (about 510 / 1030 / 2070 reads) above, the per-round kernels
    beyond (engine_resident.cuh: r_class). Depth pairs on both sides of every boundary, shallow against deep in both
    orders, with allele frequency distributions."""
    pairs = [(109, 109), (110, 110), (111, 60), (60, 111), (12, 509), (510, 510), (511, 40), (40, 515), (1029, 1030),
             (1031, 25), (25, 1035), (2069, 2070), (2071, 300), (300, 2075), (10, 10), (2600, 2600)]
    sc, b = _tn_with_depths(pairs, seed=17)
    flat = sc.flatten()
    _compare(oracle.call_batch(flat, b, afd_capacity=64, n_threads=4), engine_call(flat, b, afd_capacity=64))


def test_golden_and_real_pileups(engine_call, golden_dir):
    exp = json.load(open(os.path.join(golden_dir, "flamegraph_expected.json")))
    flat = Scenario.from_yaml(exp["scenario_yaml"]).flatten()
    b = LocusBatch.load(os.path.join(golden_dir, "flamegraph_obs.npz"))
    _compare(oracle.call_batch(flat, b, afd_capacity=64), engine_call(flat, b, afd_capacity=64))
    meta = json.load(open(os.path.join(golden_dir, "real_pileups.json")))
    allb = LocusBatch.load(os.path.join(golden_dir, "real_pileups.npz"))
    lo = 0
    n = 0
    for tc in meta["testcases"]:
        bb = allb.slice(lo, lo + tc["n_loci"])
        lo += tc["n_loci"]
        try:
            sc = Scenario.from_yaml(tc["scenario_yaml"])
        except NotImplementedError:
            continue
        if len(sc.sample_names) != 1:
            continue
        flat = sc.flatten()
        _compare(oracle.call_batch(flat, bb, afd_capacity=128), engine_call(flat, bb, afd_capacity=128))
        n += 1
    assert n >= 5


def test_edge_cases(engine_call):
    flat = Scenario.tumor_normal(0.75).flatten()
    ref = dict(prob_alt=np.log(1e-3 / 3), prob_ref=np.log1p(-1e-3), prob_mapping=np.log1p(-1e-6))
    alt = dict(prob_ref=np.log(1e-3 / 3), prob_alt=np.log1p(-1e-3), prob_mapping=np.log1p(-1e-6))
    mk = lambda d, i: read(strand=i % 2, orientation=i % 2, prob_double_overlap=-np.inf, **d)  # noqa: E731
    loci = [
        [[], []],
        [[mk(ref, 0)], []],
        [[mk(ref, i) for i in range(30)], [mk(ref, i) for i in range(30)]],
        [[mk(ref, i) for i in range(30)], [mk(ref, i) for i in range(29)] + [mk(alt, 0)]],
        [[mk(ref, i) for i in range(12)], [mk(alt, i) for i in range(12)]],
        [[mk(alt, i) for i in range(12)], [mk(alt, i) for i in range(12)]],
        [[read(prob_mapping=-np.inf, prob_alt=-1.0, prob_ref=-2.0, strand=0, orientation=0) for _ in range(6)],
         [mk(alt, i) for i in range(3)]],
        [[mk(ref, i) for i in range(8)] + [read(orientation=abi.ORIENT_F1F2, **alt)],
         [mk(alt, i) for i in range(4)] + [mk(ref, i) for i in range(4)]],
    ]
    b = batch_from_reads(loci)
    o = oracle.call_batch(flat, b, afd_capacity=128)
    g = engine_call(flat, b, afd_capacity=128)
    _compare(o, g)
    assert g.status[3] & abi.ST_SINGLETON_ADJUSTED
    assert g.status[7] & abi.ST_FILTERED_NONSTANDARD


LFC_YAML = """
samples:
  a:
    universe: "[0.0,1.0]"
    resolution: 0.05
  b:
    universe: "[0.0,1.0]"
    resolution: 0.05
events:
  up: "l2fc(a,b) >= 1.0 & a:]0.0,1.0] & b:]0.0,1.0]"
  same: "l2fc(a,b) < 1.0 & a:]0.0,1.0] & b:]0.0,1.0]"
  only_a: "a:]0.0,1.0] & b:0.0"
"""


def test_log2_fold_change_and_variant_nodes(engine_call):
    sc = Scenario.from_yaml(LFC_YAML)
    flat = sc.flatten()
    _, b = synth.tumor_normal(40, seed=21, depth=30)
    o = oracle.call_batch(flat, b, afd_capacity=128, n_threads=4)
    g = engine_call(flat, b, afd_capacity=128)
    if engine_call.kind == "emu":  # same libm as the oracle: every decision is the same one
        _compare(o, g)
    else:
        # The limits the reference infers from a predicate put abscissae exactly ON its threshold (b = a / 2), where
        # `log2(a) - log2(b) >= 1` hangs on the last bit of log2 - glibc's for the reference and the oracle, CUDA's on
        # the device. Loci whose result depends on such ties (almost all: OracleOutput.lfc_threshold_ties) can only be
        # required to agree mostly; the others must agree like everywhere else.
        ties = o.lfc_threshold_ties()
        _compare(o, g, ok=~o.knife_edge() & ~ties)
        same = np.array([max_abs_delta(o.log_posteriors[i], g.log_posteriors[i]) <= TOL and
                         np.array_equal(o.map_vaf[i], g.map_vaf[i], equal_nan=True) for i in np.nonzero(ties)[0]])
        assert same.mean() >= 0.7, "log2-fold-change loci that agree with the oracle: %g" % same.mean()
    sc2 = Scenario.from_yaml("""
samples:
  s:
    universe: "[0.0,1.0]"
events:
  ct: "C>T & s:]0.0,1.0]"
  other: "!C>T & s:]0.0,1.0]"
""")
    flat2 = sc2.flatten()
    _, b2 = synth.tumor_normal(30, seed=22, depth=25)
    one = LocusBatch(1, b2.read_offsets[::2].copy(), {k: v for k, v in b2.columns.items()}, b2.read_flags,
                     b2.locus_flags)  # merge both pileups of each locus into one sample
    _compare(oracle.call_batch(flat2, one, afd_capacity=64), engine_call(flat2, one, afd_capacity=64))


def test_population_and_somatic_priors(engine_call):
    """Prior branches beyond uniform/Mendelian: somatic rate (germline odometer), clonal/subclonal inheritance."""
    sc = Scenario.from_yaml("""
species:
  heterozygosity: 0.001
  somatic-effective-mutation-rate: 1e-6
  ploidy: 2
samples:
  normal:
    resolution: 0.1
  tumor:
    resolution: 0.05
    inheritance:
      clonal:
        from: normal
        somatic: false
    contamination:
      by: normal
      fraction: 0.2
events:
  germline: "(normal:0.5 | normal:1.0)"
  somatic_normal: "normal:]0.0,0.5[ | normal:]0.5,1.0["
  somatic_tumor: "normal:0.0 & tumor:]0.0,1.0]"
""")
    flat = sc.flatten()
    _, b = synth.tumor_normal(25, seed=31, depth=40)
    _compare(oracle.call_batch(flat, b, afd_capacity=128, n_threads=4), engine_call(flat, b, afd_capacity=128))
    full = Scenario.from_yaml(synth.SIMPLE_PEDIGREE_YAML, full_prior=True).flatten()
    _, b3 = synth.pedigree(60, seed=32, depth=30)
    _compare(oracle.call_batch(full, b3), engine_call(full, b3))


def test_config1_real_pileups_paired_as_tumor_normal(engine_call, golden_dir):
    """BASELINE config 1 (plumbing): ~100 tumor-normal loci built by pairing the real-data pileups embedded in the
    reference's testcases (depth 2..2991, indels with prob_sample_alt < 0, homopolymer columns, f16/f32 quantised)."""
    single = LocusBatch.load(os.path.join(golden_dir, "real_pileups.npz"))
    n = single.n_loci
    pairs = [(i, j) for i in range(n) for j in range(n) if i != j and (i + 2 * j) % 3 != 0][:40]
    b = pair_as_tumor_normal(single, pairs)
    flat = Scenario.tumor_normal(0.8).flatten()
    _compare(oracle.call_batch(flat, b, afd_capacity=128, n_threads=4), engine_call(flat, b, afd_capacity=128))


def test_four_samples_nested_ranges(engine_call):
    flat = Scenario.from_yaml(FOUR_SAMPLE_YAML).flatten()
    b = four_sample_batch(12, seed=51)
    _compare(oracle.call_batch(flat, b, afd_capacity=128, n_threads=4), engine_call(flat, b, afd_capacity=128))


def test_map_must_be_contained_in_the_best_event(engine_call):
    """calling.rs:861-864 skips base events that are not contained in the strongest event. With fewer than 10 reads the
    integration limits are the range bounds themselves even when they are exclusive (formula.rs:1172-1224), so for
    `tumor:]0.0,1.0]` the point 0.0 is evaluated, can have the highest joint (two reference reads) and is still not a
    valid MAP: the reference reports the next best contained point (0.1). Found by the read-level fuzzer."""
    import math
    hi, lo = math.log1p(-1e-3), math.log(1e-3 / 3)

    def r(alt, k):
        return read(prob_mapping=math.log1p(-1e-6), prob_alt=hi if alt else lo, prob_ref=lo if alt else hi,
                    strand=abi.STRAND_FORWARD if k % 2 else abi.STRAND_REVERSE,
                    orientation=abi.ORIENT_F1R2 if k % 2 else abi.ORIENT_F2R1, prob_double_overlap=-math.inf)
    b = batch_from_reads([[[r(k % 2 == 0, k // 2) for k in range(40)], [r(False, 0), r(False, 1)]]])
    flat = Scenario.tumor_normal(0.75).flatten()
    want = oracle.call_batch(flat, b, afd_capacity=64)
    assert want.map_vaf.tolist() == [[0.5, 0.1]] and want.best_event.tolist() == [2]  # germline_het
    # (the wavefront pipeline defers this locus: an integration limit on an excluded bound)
    for got in (engine_call(flat, b, afd_capacity=64), emu.wave_call_batch(flat, b, afd_capacity=64)[0]):
        assert max_abs_delta(want.log_posteriors, got.log_posteriors) <= TOL
        assert np.array_equal(want.map_vaf, got.map_vaf)
