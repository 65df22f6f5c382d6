"""Host-side logic: scenario normalisation/flattening and the observation codec."""
import json
import os

import numpy as np

from varlociraptor_b200 import LocusBatch, Scenario, abi, obs_codec, synth
from varlociraptor_b200.batch import mini_logprob
from varlociraptor_b200.scenario import VAFRange, parse_formula, parse_vafdef


def test_tumor_normal_trees_match_appendix_a():
    flat = Scenario.tumor_normal(0.75).flatten()
    assert flat.sample_names == ["normal", "tumor"]  # BTreeMap order (grammar/mod.rs:178-190)
    assert flat.event_names == ["absent", "germline_het", "germline_hom", "somatic_normal", "somatic_tumor"]
    assert flat.c.samples[1].contamination_by == 0
    assert abs(flat.c.samples[1].contamination_fraction - 0.25) < 1e-15
    d = flat.describe()
    assert "somatic_normal:\n  normal:]0,0.5[\n    tumor:]0,1]" in d
    assert "absent:\n  normal:{0}\n    tumor:{0}" in d


def test_range_overlap_and_intersection():  # formula.rs:1598-1735 (range tests)
    a, b = VAFRange(0.0, 0.7, False, False), VAFRange(0.3, 1.0, False, False)
    assert a.intersect(b) == VAFRange(0.3, 0.7, False, False)
    assert VAFRange(0.0, 0.5, False, True).no_overlap(VAFRange(0.5, 1.0, False, False))
    assert not VAFRange(0.0, 0.5, False, False).no_overlap(VAFRange(0.5, 1.0, False, False))
    assert parse_vafdef("]0.0,0.5[") == VAFRange(0.0, 0.5, True, True)
    assert parse_vafdef("{0.0,0.5}") == frozenset([0.0, 0.5])


def test_negation_against_universe_and_expressions():
    sc = Scenario.from_yaml(synth.SIMPLE_PEDIGREE_YAML)
    f = sc.normalize(parse_formula("!mother:0.0"))
    assert f.sample == "mother" and f.vafs == frozenset([0.5, 1.0])
    flat = sc.flatten()
    assert flat.sample_names == ["child", "father", "mother"]
    assert flat.event_names == ["absent", "denovo_child", "inherited"]
    assert flat.c.samples[0].inheritance == abi.INHERIT_MENDELIAN
    assert (flat.c.samples[0].parent_a, flat.c.samples[0].parent_b) == (2, 1)


def test_mini_logprob_rounding():
    x = np.array([-0.5, -9.99, -10.5, -800.123, -70000.0, -np.inf, 0.0])
    y = mini_logprob(x)
    assert y[0] == np.float32(-0.5) and y[1] == np.float32(-9.99)
    assert y[2] == np.float32(np.float16(-10.5))
    assert y[3] == np.float32(-800.123)  # f16 projection changes the integer floor -> f32
    assert y[4] == np.float32(-70000.0)
    assert np.isneginf(y[5]) and y[6] == 0.0


def test_codec_decodes_golden_records(golden_dir):
    b = LocusBatch.load(os.path.join(golden_dir, "flamegraph_obs.npz"))
    exp = json.load(open(os.path.join(golden_dir, "flamegraph_expected.json")))
    assert b.n_loci == 11 and b.n_samples == 1
    # <METH> records: not SNV/MNV -> only strand and alt-locus bias are checked (calling.rs:555-566)
    assert all(int(f) & 0x3f == (abi.LF_CHECK_SB | abi.LF_CHECK_ALB) for f in b.locus_flags)
    assert np.all(b.columns["prob_mapping"] <= 0) and np.all(np.isfinite(b.columns["prob_hit_base"]))
    strands = (b.read_flags >> abi.RF_STRAND_SHIFT) & 3
    assert set(np.unique(strands)) <= {0, 1, 2, 3}
    assert len(exp["records"]) == 11


def test_codec_roundtrip_of_synthetic_record():
    """encode (as the reference's write_observations would) -> decode."""
    vals = np.array([-0.25, -12.5, -1000.0, -np.inf], dtype=np.float64)
    body = bytearray(np.uint64(len(vals)).tobytes())
    for v in vals:
        h = np.float16(v)
        if v < -10 and np.floor(np.float64(h)) == np.floor(v):
            body += np.uint32(0).tobytes() + h.tobytes()
        else:
            body += np.uint32(1).tobytes() + np.float32(v).tobytes()
    if len(body) % 2:
        body.append(0)
    ints = np.frombuffer(bytes(body), dtype="<u2").astype(np.int32).tolist()
    out = obs_codec.decode_mini_logprobs(ints)
    assert np.array_equal(out, mini_logprob(vals))


def test_batch_slice_select_concat():
    _, b = synth.tumor_normal(10, seed=3, depth_range=(3, 9))
    parts = [b.slice(0, 4), b.slice(4, 10)]
    c = LocusBatch.concat(parts)
    assert np.array_equal(c.read_offsets, b.read_offsets)
    assert all(np.array_equal(c.columns[k], b.columns[k], equal_nan=True) for k in abi.BATCH_F32_COLUMNS)
    s = b.select([9, 0])
    assert np.array_equal(s.slice(1, 2).columns["prob_alt"], b.slice(0, 1).columns["prob_alt"])


def test_codec_encode_reproduces_the_reference_files_bit_for_bit(golden_dir):
    """decode -> encode of records taken verbatim from the reference's observation files gives back the identical
    INFO integer arrays (every tag the engine's batch layout carries)."""
    raw = json.load(open(os.path.join(golden_dir, "raw_info_records.json")))
    n_tags = 0
    for rec in raw["records"]:
        cols, flags, hart, hvar = obs_codec.decode_record(rec["info"])
        back = obs_codec.encode_record(cols, flags, hart, hvar)
        for tag, ints in back.items():
            assert ints == rec["info"][tag], (rec["pos"], tag)
            n_tags += 1
    assert n_tags >= 3 * 14


def test_codec_round_trip_property():
    """encode_record -> decode_record is the identity on MiniLogProb-representable columns and on every flag
    combination (hypothesis), including -inf, -0.0, half-precision boundaries and empty pileups."""
    from hypothesis import given, settings, strategies as st
    special = st.sampled_from([0.0, -0.0, -np.inf, -10.0, -10.000001, -9.999999, -65504.0, -65520.0, -1e-8, -745.0])
    values = st.lists(st.one_of(special, st.floats(min_value=-1e5, max_value=0.0, allow_nan=False)), max_size=12)

    @settings(max_examples=150, deadline=None)
    @given(values, st.integers(0, 2 ** 31 - 1))
    def check(vals, seed):
        n = len(vals)
        rng = np.random.Generator(np.random.PCG64(seed))
        cols = {k: mini_logprob(np.array(vals, dtype=np.float64) if i == 0 else -rng.exponential(20.0, n))
                for i, k in enumerate(abi.BATCH_F32_COLUMNS)}
        flags = ((rng.integers(0, 4, n) << abi.RF_STRAND_SHIFT) | (rng.integers(0, 9, n) << abi.RF_ORIENT_SHIFT)
                 | (rng.integers(0, 3, n) << abi.RF_ALTLOCUS_SHIFT) | (rng.integers(0, 2, n) * abi.RF_READPOS_MAJOR)
                 | (rng.integers(0, 2, n) * abi.RF_SOFTCLIPPED) | (rng.integers(0, 2, n) * abi.RF_PAIRED)
                 | (rng.integers(0, 2, n) * abi.RF_MAX_MAPQ)).astype(np.uint32)
        got_cols, got_flags, hart, hvar = obs_codec.decode_record(obs_codec.encode_record(cols, flags))
        assert hart is None and hvar is None and np.array_equal(got_flags, flags)
        for k in cols:
            assert np.array_equal(got_cols[k].view(np.uint32), cols[k].astype(np.float32).view(np.uint32)), k
    check()
