"""GPU-less check of the all-Set pipeline (engine_sets.cuh: pedigrees and other scenarios whose trees are Set nodes
only): the same device functions the CUDA kernels call, run sequentially on the host (tests/emu, test infrastructure
only), against the oracle. On the GPU the pipeline is reached through the C-ABI by every pedigree test (-m gpu)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from tests import emu
from tests.test_emu_parity import _compare
from tests.util import batch_from_reads, read
from varlociraptor_b200 import LocusBatch, Scenario, abi, synth


def test_pedigree_with_afd_and_indels():
    sc, b = synth.pedigree(400, seed=6)
    flat = sc.flatten()
    g, deferred = emu.sets_call_batch(flat, b, afd_capacity=8)
    assert deferred == 0
    _compare(oracle.call_batch(flat, b, afd_capacity=8, n_threads=4), g)
    full = Scenario.from_yaml(synth.SIMPLE_PEDIGREE_YAML, full_prior=True).flatten()
    g2, _ = emu.sets_call_batch(full, b)
    _compare(oracle.call_batch(full, b, n_threads=4), g2)


@pytest.mark.parametrize("name,contig", [("simple-pedigree", "all"), ("pedigree", "X"), ("pedigree", "Y"), ("population", "all")])
def test_reference_prior_scenarios(golden_dir, name, contig):
    from tests.test_prior_scenarios import _batch
    text = json.load(open(os.path.join(golden_dir, "prior_scenarios.json")))["scenarios"][name]
    flat = Scenario.from_yaml(text).for_contig(contig).flatten()
    b = _batch(flat.n_samples, 24, seed=11)
    try:
        g, _ = emu.sets_call_batch(flat, b, afd_capacity=16)
    except LookupError:
        pytest.skip("not an all-Set scenario of at most three samples")
    _compare(oracle.call_batch(flat, b, afd_capacity=16, n_threads=4), g)


def test_deferred_and_edge_loci():
    """Per-record prior overrides and pileups deeper than the shared-memory arena go to the generic engine; empty and
    clearly-reference pileups, a singleton alt read and filtered alignments stay in the pipeline."""
    sc = Scenario.from_yaml(synth.SIMPLE_PEDIGREE_YAML)
    flat = sc.flatten()
    ref = dict(prob_alt=np.log(1e-3 / 3), prob_ref=np.log1p(-1e-3), prob_mapping=np.log1p(-1e-6))
    alt = dict(prob_ref=np.log(1e-3 / 3), prob_alt=np.log1p(-1e-3), prob_mapping=np.log1p(-1e-6))
    mk = lambda d, i: read(strand=i % 2, orientation=i % 2, prob_double_overlap=-np.inf, **d)  # noqa: E731
    loci = [
        [[], [], []],
        [[mk(ref, 0)], [], [mk(alt, 1)]],
        [[mk(ref, i) for i in range(30)], [mk(ref, i) for i in range(30)], [mk(ref, i) for i in range(30)]],
        [[mk(ref, i) for i in range(30)], [mk(ref, i) for i in range(29)] + [mk(alt, 0)], [mk(ref, i) for i in range(12)]],
        [[mk(alt, i) for i in range(12)] + [mk(ref, i) for i in range(12)], [mk(ref, i) for i in range(20)],
         [mk(ref, i) for i in range(20)]],
        [[mk(ref, i) for i in range(8)] + [read(orientation=abi.ORIENT_F1F2, **alt)], [mk(alt, i) for i in range(9)],
         [mk(alt, i) for i in range(4)] + [mk(ref, i) for i in range(4)]],
        # 450 reads (not all-alt in the father: child het on reference reads would tie exactly with father het on alt reads)
        [[mk(ref, i) for i in range(150)], [mk(alt, i) for i in range(90)] + [mk(ref, i) for i in range(60)],
         [mk(ref, i) for i in range(150)]],
    ]
    b = batch_from_reads(loci)
    het = np.full(b.n_loci, np.nan, dtype=np.float32)
    het[4] = 12.0
    b = LocusBatch(3, b.read_offsets, b.columns, b.read_flags, b.locus_flags, None, None, het, None)
    o = oracle.call_batch(flat, b, afd_capacity=16)
    g, deferred = emu.sets_call_batch(flat, b, afd_capacity=16)
    assert deferred == 2  # the override and the deep locus
    _compare(o, g)
    assert g.status[3] & abi.ST_SINGLETON_ADJUSTED and g.status[5] & abi.ST_FILTERED_NONSTANDARD


def test_tumor_normal_is_not_an_all_set_scenario():
    sc, b = synth.tumor_normal(2, seed=1)
    with pytest.raises(LookupError):
        emu.sets_call_batch(sc.flatten(), b)
