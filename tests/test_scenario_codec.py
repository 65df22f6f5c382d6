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
