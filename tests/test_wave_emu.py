"""GPU-less check of the wavefront pipeline (engine_wave.cuh): the same device functions the CUDA kernels call, run
sequentially on the host (tests/emu, test infrastructure only), against the oracle. The parity tests proper are
tests/test_gpu_parity.py (-m gpu), which reach the pipeline through the C-ABI."""
import os

import numpy as np
import pytest

from oracle import oracle
from tests import emu
from tests.test_emu_parity import LFC_YAML, _compare
from tests.util import batch_from_reads, pair_as_tumor_normal, read
from varlociraptor_b200 import LocusBatch, Scenario, abi, synth


def test_tumor_normal_with_afd():
    sc, b = synth.tumor_normal(300, seed=5)
    flat = sc.flatten()
    g, deferred = emu.wave_call_batch(flat, b, afd_capacity=96)
    assert deferred == 0
    _compare(oracle.call_batch(flat, b, afd_capacity=96, n_threads=4), g)


def test_depth_skew_and_other_purity():
    sc, b = synth.tumor_normal(24, seed=8, depth_range=(10, 2000))
    flat = sc.flatten()
    g, _ = emu.wave_call_batch(flat, b)
    _compare(oracle.call_batch(flat, b, n_threads=4), g)
    flat2 = Scenario.tumor_normal(1.0).flatten()  # no contamination: the tumor pileup does not see the normal's VAF
    _, b2 = synth.tumor_normal(60, seed=9, depth=30)
    g2, _ = emu.wave_call_batch(flat2, b2, afd_capacity=64)
    _compare(oracle.call_batch(flat2, b2, afd_capacity=64, n_threads=4), g2)


def test_edge_cases_are_deferred_or_exact():
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
    g, deferred = emu.wave_call_batch(flat, b, afd_capacity=128)
    assert deferred >= 2  # empty / tiny pileups take the Simpson branches of the generic engine
    _compare(o, g)
    assert g.status[3] & abi.ST_SINGLETON_ADJUSTED
    assert g.status[7] & abi.ST_FILTERED_NONSTANDARD


def test_real_pileups_paired_as_tumor_normal(golden_dir):
    """indels (prob_sample_alt < 0: per-read x/y corrections), homopolymer columns, depth 2..2991"""
    single = LocusBatch.load(os.path.join(golden_dir, "real_pileups.npz"))
    n = single.n_loci
    pairs = [(i, j) for i in range(n) for j in range(n) if i != j and (i + 2 * j) % 3 != 0][:40]
    b = pair_as_tumor_normal(single, pairs)
    flat = Scenario.tumor_normal(0.8).flatten()
    g, deferred = emu.wave_call_batch(flat, b, afd_capacity=128)
    assert deferred < b.n_loci
    _compare(oracle.call_batch(flat, b, afd_capacity=128, n_threads=4), g)


def test_other_scenarios_are_not_eligible():
    for flat in (Scenario.from_yaml(LFC_YAML).flatten(), Scenario.from_yaml(synth.SIMPLE_PEDIGREE_YAML).flatten()):
        _, b = synth.tumor_normal(2, seed=1, depth=12)
        with pytest.raises(LookupError):
            emu.wave_call_batch(flat, b)


def test_two_level_chain_other_names_and_point_event():
    """Not the CLI's tumor-normal text: other resolutions / contamination and a second point event (both VAFs fixed)."""
    sc = Scenario.from_yaml("""
samples:
  a:
    universe: "0.0 | 0.5 | 1.0 | ]0.0,0.5["
    resolution: 0.05
  b:
    universe: "[0.0,1.0]"
    resolution: 0.02
    contamination:
      by: a
      fraction: 0.4
events:
  only_b: "a:0.0 & b:]0.0,1.0]"
  both_low: "a:]0.0,0.5[ & b:]0.0,1.0]"
  het: "a:0.5 & b:]0.0,1.0]"
  hom_point: "a:1.0 & b:1.0"
""")
    flat = sc.flatten()
    _, b = synth.tumor_normal(80, seed=77, depth=40)
    o = oracle.call_batch(flat, b, afd_capacity=128, n_threads=4)
    g, _ = emu.wave_call_batch(flat, b, afd_capacity=128)
    _compare(o, g)
