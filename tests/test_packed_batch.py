"""Packed batch (include/vlr_engine.h: vlr_packed_batch_t): the lossless per-column encodings of the host -> device link.

CPU part: vlr_pack_batch is host code — every encoding decodes to the source column's exact f32 bit patterns, the
smallest lossless encoding is chosen, the byte count is what the header says. GPU part: vlr_call_batch_packed returns
bit for bit what vlr_call_batch returns (the device-side widening, vlr_unpack_kernel, is exact), including chunked
batches and every encoding."""
import numpy as np
import pytest

from varlociraptor_b200 import abi, engine, synth
from varlociraptor_b200.batch import LocusBatch


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _custom_batch(n_loci=1200, depth=37, seed=5):
    """Tumor-normal SNV loci whose columns exercise every encoding: prob_mapping const, prob_ref f16-exact with > 256
    values, prob_alt continuous (f32), prob_missed_allele 16-bit dictionary, the rest 8-bit dictionaries."""
    sc, b = synth.tumor_normal(n_loci, seed=seed, depth=depth)
    rng = np.random.default_rng(seed)
    n = b.n_reads
    b.columns["prob_mapping"] = np.full(n, np.float32(-1e-6))
    b.columns["prob_ref"] = (-rng.integers(0, 3000, n) / 64.0).astype(np.float16).astype(np.float32)
    b.columns["prob_alt"] = (-rng.random(n) * 9.0).astype(np.float32)
    pool = (-rng.random(5000) * 4.0 - 0.3).astype(np.float32)
    b.columns["prob_missed_allele"] = pool[rng.integers(0, len(pool), n)]
    return sc, b


def test_pack_is_lossless_and_picks_the_smallest_encoding():
    sc, b = _custom_batch()
    pk = engine.PackedBatch(b, n_threads=3)
    enc = pk.encodings
    assert enc["prob_mapping"] == ("const", 1)
    assert enc["prob_ref"][0] == "f16"
    assert enc["prob_alt"][0] == "f32"
    assert enc["prob_missed_allele"][0] == "dict16" and 256 < enc["prob_missed_allele"][1] <= 5000
    assert enc["prob_double_overlap"] == ("const", 1) and enc["prob_sample_alt"] == ("const", 1)
    assert enc["read_flags"][0] == "dict8"
    for name in abi.PACKED_COLUMNS:
        src = b.read_flags if name == "read_flags" else b.columns[name]
        assert np.array_equal(pk.decode(name), _bits(src)), name
    per_read = sum(abi.ENC_BYTES[abi.ENC_NAMES.index(e)] for e, _ in enc.values())
    dicts = sum(4 * n for _, n in enc.values())
    assert pk.nbytes() == per_read * b.n_reads + dicts + 8 * (b.n_loci * 2 + 1) + 4 * b.n_loci
    assert pk.nbytes() < b.nbytes() / 2


def test_pack_synthetic_config_columns_are_dictionaries():
    """The synthetic generator's reads (21 base qualities, 61 MAPQs) pack to 9 bytes per read; -0.0 / NaN / inf
    patterns survive (dictionaries hold bit patterns, not values)."""
    sc, b = synth.config(2, 500)
    b.columns["prob_hit_base"] = b.columns["prob_hit_base"].copy()
    b.columns["prob_hit_base"][:4] = np.array([-0.0, 0.0, np.nan, -np.inf], dtype=np.float32)
    pk = engine.PackedBatch(b)
    for name in abi.PACKED_COLUMNS:
        src = b.read_flags if name == "read_flags" else b.columns[name]
        assert np.array_equal(pk.decode(name), _bits(src)), name
    assert all(e in ("dict8", "const") for e, _ in pk.encodings.values())
    assert pk.nbytes() <= 7 * b.n_reads + 8 * (b.n_loci * 2 + 1) + 4 * b.n_loci + 4096


def test_pack_empty_and_single_thread():
    sc, b = synth.config(2, 3)
    e = b.slice(0, 0)
    pk = engine.PackedBatch(e)
    assert pk.n_reads == 0 and all(x == ("f32", 0) for x in pk.encodings.values())
    pk1 = engine.PackedBatch(b, n_threads=1)
    pk8 = engine.PackedBatch(b, n_threads=8)
    assert pk1.encodings == pk8.encodings
    for name in abi.PACKED_COLUMNS:
        assert np.array_equal(pk1.decode(name), pk8.decode(name))


def _same(a, b):
    for f in ("log_posteriors", "log_marginal", "map_vaf"):
        x, y = getattr(a, f), getattr(b, f)
        assert np.array_equal(x.view(np.uint64), y.view(np.uint64)), f
    for f in ("map_config", "best_event", "status", "n_base_events"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [2, 3])
def test_call_batch_packed_is_bitwise_call_batch(cfg):
    sc, b = synth.config(cfg, 3000)
    flat = sc.flatten()
    eng = engine.PosteriorEngine(flat)
    want = eng.call_batch(b)
    pk = engine.PackedBatch(b)
    got = eng.call_batch_packed(pk)
    assert eng.launches >= 2  # the unpack kernel + the pipeline
    _same(want, got)


@pytest.mark.gpu
def test_call_batch_packed_every_encoding_chunked_with_afd():
    """Every encoding in one batch, more loci than one chunk (65 536) with an odd number of reads per chunk (row
    offsets of the encoded columns are not multiples of 4), AFDs on a slice."""
    sc, b = _custom_batch(n_loci=70001, depth=9, seed=11)
    flat = sc.flatten()
    eng = engine.PosteriorEngine(flat)
    pk = engine.PackedBatch(b)
    assert {e for e, _ in pk.encodings.values()} == {"const", "f16", "f32", "dict16", "dict8"}
    _same(eng.call_batch(b), eng.call_batch_packed(pk))
    small = b.slice(100, 400)
    w = eng.call_batch(small, afd_capacity=256)
    g = eng.call_batch_packed(engine.PackedBatch(small), afd_capacity=256)
    _same(w, g)
    assert np.array_equal(w.afd_count, g.afd_count)
    assert np.array_equal(w.afd_logp.view(np.uint64), g.afd_logp.view(np.uint64))
