"""Host logic of the operator API mirror (varlociraptor_b200/calling.py) without a GPU: the engine is replaced by
the test-only host emulation, everything else (lock-step reading, candidate filter, batching, ordering, formatting
of the final-record fields) is the product code."""
import json
import os

import numpy as np
import pytest

from tests import emu
from oracle import oracle
from varlociraptor_b200 import Scenario, calling, obs_codec, synth
from varlociraptor_b200.batch import LocusBatch

REF_FLAME = "/root/reference/tests/resources/flamegraph_profiling"


class EmuEngine:
    def __init__(self, flat):
        self.flat = flat
        self.calls = 0

    def call_batch(self, batch, afd_capacity=0, out=None):
        self.calls += 1
        return emu.call_batch(self.flat, batch, afd_capacity=afd_capacity)


def _records_from_batch(batch: LocusBatch):
    """Re-encode a one-sample LocusBatch into the record dicts `parse_observation_vcf` produces."""
    def enc_probs(vals):
        body = bytearray(np.uint64(len(vals)).tobytes())
        for v in vals:
            body += np.uint32(1).tobytes() + np.float32(v).tobytes()
        return body

    def enc_enum(vals):
        body = bytearray(np.uint64(len(vals)).tobytes())
        for v in vals:
            body += np.uint32(int(v)).tobytes()
        return body

    def enc_bits(bits):
        n = len(bits)
        blocks = np.packbits(np.asarray(bits, dtype=np.uint8), bitorder="little").tobytes()
        return bytearray(b"\x01" + np.uint64(len(blocks)).tobytes() + blocks + np.uint64(n).tobytes())

    def ints(body):
        if len(body) % 2:
            body.append(0)
        return np.frombuffer(bytes(body), dtype="<u2").astype(np.int64).tolist()
    recs = []
    for i in range(batch.n_loci):
        lo, hi = batch.read_offsets[i], batch.read_offsets[i + 1]
        f = batch.read_flags[lo:hi]
        info = {tag: ints(enc_probs(batch.columns[col][lo:hi])) for tag, col in obs_codec._PROB_TAGS.items()}
        info["STRAND"] = ints(enc_enum(f & 3))
        info["READ_ORIENTATION"] = ints(enc_enum((f >> 2) & 15))
        info["READ_POSITION"] = ints(enc_enum(np.where(f & (1 << 6), 0, 1)))
        info["ALT_LOCUS"] = ints(enc_enum((f >> 10) & 3))
        info["SOFTCLIPPED"] = ints(enc_bits((f >> 7) & 1))
        info["PAIRED"] = ints(enc_bits((f >> 8) & 1))
        info["IS_MAX_MAPQ"] = ints(enc_bits((f >> 9) & 1))
        recs.append({"chrom": "1", "pos": 100 + i, "ref": "A", "alt": "G", "info": info, "flags": set()})
    return recs


@pytest.fixture(scope="module")
def tn_records():
    from varlociraptor_b200 import synth
    _, b = synth.tumor_normal(12, seed=3, depth=20)
    normal = LocusBatch(1, b.read_offsets[0::2][:13] * 0, {}, np.zeros(0, np.uint32), np.zeros(0, np.uint32)) \
        if False else None
    # split the two-sample batch into two one-sample batches
    def one(s):
        starts = b.read_offsets[s:-1:2]
        ends = b.read_offsets[s + 1::2]
        idx = np.concatenate([np.arange(a, e) for a, e in zip(starts, ends)])
        offs = np.concatenate([[0], np.cumsum(ends - starts)])
        return LocusBatch(1, offs, {k: v[idx] for k, v in b.columns.items()}, b.read_flags[idx], b.locus_flags)
    return _records_from_batch(one(0)), _records_from_batch(one(1)), b


def test_call_generic_orders_batches_and_formats(tn_records):
    normal, tumor, b = tn_records
    sc = Scenario.tumor_normal(0.75)
    flat = sc.flatten()
    eng = EmuEngine(flat)
    writer = calling.call_generic(sc, {"tumor": tumor, "normal": normal}, engine=eng, batch_size=5)
    assert eng.calls == 3  # 12 records in batches of 5
    assert [c.pos for c in writer.calls] == [100 + i for i in range(12)]
    want = emu.call_batch(flat, b, afd_capacity=128)
    for i, c in enumerate(writer.calls):
        for e, name in enumerate(flat.event_names):
            assert c.event_probs[name] == want.log_posteriors[i, e] or \
                (np.isnan(c.event_probs[name]) and np.isnan(want.log_posteriors[i, e]))
        assert set(c.info_fields()) == {"PROB_ABSENT", "PROB_GERMLINE_HET", "PROB_GERMLINE_HOM",
                                        "PROB_SOMATIC_NORMAL", "PROB_SOMATIC_TUMOR", "PROB_ARTIFACT"}
        assert all(v >= 0 for v in c.info_fields().values())
        assert c.sample_info[1].allelefreq_estimate == want.map_vaf[i, 1]
        assert c.sample_info[0].depth == 20
    lines = writer.lines()
    assert len(lines) == 12 and lines[0].split("\t")[8] == "DP:AF:SAOBS:SROBS:OBS:OOBS:SB:ROB:RPB:SCB:HE:ALB:AFD"


def test_candidate_filter_and_missing_sample(tn_records):
    normal, tumor, _ = tn_records

    class EveryOther(calling.CandidateFilter):
        def filter(self, work_item, sample_names):  # noqa: A003
            assert work_item.pileups.n_samples == 2 and sample_names == ["normal", "tumor"]
            return work_item.index % 2 == 0
    sc = Scenario.tumor_normal(0.75)
    w = calling.call_generic(sc, {"tumor": tumor, "normal": normal}, engine=EmuEngine(sc.flatten()),
                             candidate_filter=EveryOther())
    assert [c.pos for c in w.calls] == [100, 102, 104, 106, 108, 110]
    # a sample without observations is legal: zero coverage (calling.rs:605-607)
    sc2 = Scenario.tumor_normal(0.75)
    w2 = calling.call_generic(sc2, {"tumor": tumor}, engine=EmuEngine(sc2.flatten()))
    assert len(w2.calls) == 12 and w2.calls[0].sample_info[0].depth == 0


def test_errors_mirror_the_reference(tn_records):
    normal, tumor, _ = tn_records
    sc = Scenario.tumor_normal(0.75)
    with pytest.raises(ValueError, match="invalid observation sample name"):
        calling.call_generic(sc, {"tumour": tumor}, engine=EmuEngine(sc.flatten()))
    with pytest.raises(ValueError, match="different numbers of records"):
        calling.call_generic(sc, {"tumor": tumor, "normal": normal[:-1]}, engine=EmuEngine(sc.flatten()))
    shifted = [dict(r, pos=r["pos"] + 1) for r in normal]
    with pytest.raises(ValueError, match="inconsistent observations"):
        calling.call_generic(sc, {"tumor": tumor, "normal": shifted}, engine=EmuEngine(sc.flatten()))


def test_engine_is_required_without_gpu():
    """The product path has no CPU fallback: constructing the default engine without a device must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from varlociraptor_b200.engine import EngineError
    with pytest.raises(EngineError):
        calling.call_generic(Scenario.tumor_normal(0.75), {"tumor": [], "normal": []})


def test_golden_text_fields_with_emulated_engine(golden_dir):
    """Final-record text fields (PROB_*, AF, DP, AFD) against what the reference printed (calls.vcf), for the
    records not affected by the version drift of the bias selection (see test_oracle_golden.py)."""
    exp = json.load(open(os.path.join(golden_dir, "flamegraph_expected.json")))
    b = LocusBatch.load(os.path.join(golden_dir, "flamegraph_obs.npz"))
    recs = _records_from_batch(b)
    for r, e in zip(recs, exp["records"]):
        r["pos"], r["ref"], r["alt"] = e["pos"], "CG", "<METH>"
        r["info"]["THIRD_ALLELE_EVIDENCE"] = e["THIRD_ALLELE_EVIDENCE"]
    sc = Scenario.from_yaml(exp["scenario_yaml"])
    w = calling.call_generic(sc, {"normal": recs}, engine=EmuEngine(sc.flatten()))
    n = 0
    for c, e in zip(w.calls, exp["records"]):
        if e["pos"] in (10471, 10489, 10542):
            continue
        info = c.info_fields()
        for tag, text in e["info"].items():
            got = info[tag]
            if text == "inf":
                assert np.isinf(got)
            else:
                assert abs(float("%g" % got) - float(text)) <= 10 ** (np.floor(np.log10(max(abs(float(text)), 1e-9))) - 4)
        f = c.format_fields(0)
        assert f["AF"] == "%g" % e["AF"] and int(f["DP"]) == e["DP"]
        got_afd = [tuple(map(float, kv.split("="))) for kv in f["AFD"].split(",")]
        assert len(got_afd) == len(e["AFD"])
        for (gv, gp), (wv, wp) in zip(got_afd, e["AFD"]):
            assert gv == wv and abs(gp - wp) <= 0.0101
        assert f["SB"] == "." and f["ALB"] == "."
        # SAOBS / SROBS: same (count, letter) entries; entries with equal counts come out of a hash map upstream
        import re
        PAT = {"SAOBS": r"\d+[A-Za-z]", "SROBS": r"\d+[A-Za-z]", "OBS": r"\d+[A-Za-z]{1,2}(?:\d+|\.)[ps][#*.][*+.-][><*!][\^*][$.][*.]"}
        assert f["OOBS"] == e["OOBS"]
        for key in ("SAOBS", "SROBS", "OBS"):
            assert sorted(re.findall(PAT[key], f[key])) == sorted(re.findall(PAT[key], e[key])), (key, f[key], e[key])
        n += 1
    assert n == 8


def test_haplotype_groups_reuse_the_first_members_result(tn_records):
    """calling.rs:569-580, 726-741: records sharing INFO/EVENT (or a breakend MATEID pair) are computed once."""
    tumor, normal, _ = tn_records
    import copy
    tumor, normal = copy.deepcopy(tumor), copy.deepcopy(normal)
    for recs in (tumor, normal):
        recs[2]["info"]["EVENT"] = "ev1"
        recs[7]["info"]["EVENT"] = "ev1"        # reuses record 2's result although its own reads differ
        recs[4]["id"], recs[4]["info"]["MATEID"] = "bnd_b", "bnd_a"
        recs[9]["id"], recs[9]["info"]["MATEID"] = "bnd_a", "bnd_b"
    assert calling.haplotype_identifier(tumor[4]) == calling.haplotype_identifier(tumor[9]) == "bnd_a-bnd_b"
    sc = Scenario.tumor_normal(0.75)
    eng = EmuEngine(sc.flatten())
    seen = []
    orig = eng.call_batch
    eng.call_batch = lambda batch, afd_capacity=0, out=None: (seen.append(batch.n_loci), orig(batch, afd_capacity))[1]
    w = calling.call_generic(sc, {"tumor": tumor, "normal": normal}, engine=eng, batch_size=5)
    assert [c.pos for c in w.calls] == [r["pos"] for r in tumor]          # input order kept
    assert sum(seen) == len(tumor) - 2                                     # two records were never packed
    assert w.calls[7].event_probs == w.calls[2].event_probs and w.calls[7].sample_info == w.calls[2].sample_info
    assert w.calls[9].event_probs == w.calls[4].event_probs
    ref = calling.call_generic(sc, {"tumor": tn_records[0], "normal": tn_records[1]}, engine=EmuEngine(sc.flatten()))
    assert ref.calls[7].event_probs != w.calls[7].event_probs              # it really is the group's result
    assert ref.calls[3].event_probs == w.calls[3].event_probs
    with pytest.raises(ValueError, match="without record ID"):
        calling.haplotype_identifier({"info": {"MATEID": "x"}, "id": "."})


def test_missing_data_record(tn_records):
    """mod.rs:424-466, 559-571: a record without any read in any sample is written with missing probabilities,
    DP 0 and the hint `missing-data`; its neighbours are unaffected."""
    import copy
    tumor, normal, b = tn_records
    tumor, normal = copy.deepcopy(tumor), copy.deepcopy(normal)
    empty = _records_from_batch(LocusBatch(1, np.zeros(2, np.int64), {k: v[:0] for k, v in b.columns.items()},
                                           b.read_flags[:0], b.locus_flags[:1]))[0]
    for recs in (tumor, normal):
        recs[5]["info"] = copy.deepcopy(empty["info"])
    sc = Scenario.tumor_normal(0.75)
    w = calling.call_generic(sc, {"tumor": tumor, "normal": normal}, engine=EmuEngine(sc.flatten()))
    ref = calling.call_generic(sc, {"tumor": tn_records[0], "normal": tn_records[1]}, engine=EmuEngine(sc.flatten()))
    c = w.calls[5]
    assert c.is_missing_data and c.all_hints()[-1] == "missing-data"
    assert all(np.isnan(v) for v in c.info_fields().values())
    assert c.format_fields(0)["DP"] == "0" and c.format_fields(1)["AF"] == "." and c.format_fields(0)["OBS"] == "."
    line = w.lines()[5].split("\t")
    assert "PROB_ABSENT=." in line[7] and line[7].endswith("HINTS=missing-data") and line[9].startswith("0:.:")
    assert not w.calls[4].is_missing_data and w.calls[4].event_probs == ref.calls[4].event_probs
    assert w.lines()[4] == ref.lines()[4] and w.lines()[6] == ref.lines()[6]


def test_model_is_reconfigured_per_contig(golden_dir):
    """calling.rs:632-718: ploidy (hence universes, trees and prior) follows the contig of the record."""
    from tests.test_prior_scenarios import _batch
    text = json.load(open(os.path.join(golden_dir, "prior_scenarios.json")))["scenarios"]["pedigree"]
    sc = Scenario.from_yaml(text)
    assert sc.is_contig_dependent() and not Scenario.tumor_normal(0.75).is_contig_dependent()
    b = _batch(4, 9, seed=3)
    names = sc.sample_names

    def one(s):
        starts, ends = b.read_offsets[s:-1:4], b.read_offsets[s + 1::4]
        idx = np.concatenate([np.arange(a, e) for a, e in zip(starts, ends)])
        offs = np.concatenate([[0], np.cumsum(ends - starts)])
        return LocusBatch(1, offs, {k: v[idx] for k, v in b.columns.items()}, b.read_flags[idx], b.locus_flags)
    records = {n: _records_from_batch(one(s)) for s, n in enumerate(names)}
    contigs = ["1", "1", "2", "X", "X", "Y", "1", "X", "2"]
    for recs in records.values():
        for r, c in zip(recs, contigs):
            r["chrom"] = c
    built = []

    def factory(flat):
        built.append(flat)
        return EmuEngine(flat)
    w = calling.call_generic(sc, records, engine_factory=factory)
    assert [c.chrom for c in w.calls] == contigs
    assert len(built) == 3  # autosomes ("all", "1", "2" share one model), X, Y
    for contig in ("1", "X", "Y"):
        flat = sc.for_contig(contig).flatten()
        want = emu.call_batch(flat, b, afd_capacity=128)
        for i, c in enumerate(w.calls):
            if c.chrom == contig:
                assert [c.event_probs[e] for e in flat.event_names] == want.log_posteriors[i, :-1].tolist()
    x, a = [c for c in w.calls if c.chrom == "X"][0], [c for c in w.calls if c.chrom == "1"][0]
    father = names.index("father")
    assert x.sample_info[father].allelefreq_estimate in (0.0, 1.0)          # haploid on X
    assert a.sample_info[father].allelefreq_estimate in (0.0, 0.5, 1.0)


def test_read_observation_summary_codes():
    """FORMAT/OBS (mod.rs:277-333): one nine-part code per read, counted; most common first, `E` then `N` codes last."""
    from varlociraptor_b200 import abi
    ln = np.log

    def flags(strand, orient, altlocus, major=False, soft=False, paired=False, maxq=False, hlen=None):
        f = (strand << abi.RF_STRAND_SHIFT) | (orient << abi.RF_ORIENT_SHIFT) | (altlocus << abi.RF_ALTLOCUS_SHIFT)
        f |= (abi.RF_READPOS_MAJOR if major else 0) | (abi.RF_SOFTCLIPPED if soft else 0)
        f |= (abi.RF_PAIRED if paired else 0) | (abi.RF_MAX_MAPQ if maxq else 0)
        if hlen is not None:
            f |= abi.RF_HAS_HOMOPOLYMER_LEN | ((hlen & 0xff) << abi.RF_HOMOPOLYMER_LEN_SHIFT)
        return f
    pa = np.array([ln(0.9), ln(0.9), ln(0.5), ln(0.02), ln(0.5), ln(0.4)])
    pr = np.array([ln(0.001), ln(0.001), ln(0.5), ln(0.9), ln(0.5), ln(0.6)])
    rf = np.array([flags(0, 0, 2, paired=True, maxq=True), flags(0, 0, 2, paired=True, maxq=True),
                   flags(3, 8, 2), flags(1, 1, 0, major=True, soft=True, maxq=True, hlen=-2),
                   flags(3, 8, 2), flags(2, 5, 1, hlen=0)], dtype=np.uint32)
    third = (np.array([False, False, False, True, False, False]), np.array([0, 0, 0, 3, 0, 0], dtype=np.uint32))
    got = calling.read_observation_summary(pa, pr, rf, third)
    # two very strong alt reads (max MAPQ -> upper case), one strong ref read with third-allele distance 3, major read
    # position, softclip and a homopolymer error, one barely-ref read on both strands in a non-standard orientation,
    # two reads with equal evidence
    # (the `E`-codes-last rule is case sensitive upstream: a lower-case `e` of a read below the maximum MAPQ sorts by count)
    assert got == "2AV.p.+>*..2e.s..**..1RS3s#-<^$*1rb.s**!*.."
    upper = calling.read_observation_summary(pa[2:5], pr[2:5], rf[2:5] | abi.RF_MAX_MAPQ)
    assert upper == "1RS.s#-<^$*2E.s..**.."  # ... an upper-case `E` goes last despite its higher count
    assert calling.read_observation_summary(pa[:0], pr[:0], rf[:0]) == "."
    assert calling.read_observation_summary(pa[:1], pr[:1], rf[:1]) == "1AV.p.+>*.."


def test_batching_groups_and_filters_do_not_change_what_is_delivered(tn_records):
    """Property: whatever the batch size, calls arrive in input order with the same content; haplotype groups and a
    candidate filter interact with batching only through which records are computed."""
    import copy
    import random
    tumor0, normal0, _ = tn_records
    sc = Scenario.tumor_normal(0.75)

    class EveryThird(calling.CandidateFilter):
        def filter(self, work_item, sample_names):  # noqa: A003
            return work_item.index % 3 != 1

    def run(batch_size, groups, use_filter):
        tumor, normal = copy.deepcopy(tumor0), copy.deepcopy(normal0)
        for name, members in groups.items():
            for i in members:
                tumor[i]["info"]["EVENT"] = normal[i]["info"]["EVENT"] = name
        w = calling.call_generic(sc, {"tumor": tumor, "normal": normal}, engine=EmuEngine(sc.flatten()),
                                 batch_size=batch_size, candidate_filter=EveryThird() if use_filter else None)
        return [(c.pos, tuple(sorted(c.event_probs.items())), repr(c.sample_info)) for c in w.calls]

    rng = random.Random(5)
    for trial in range(6):
        idx = list(range(len(tumor0)))
        rng.shuffle(idx)
        groups = {"g1": sorted(idx[:3]), "g2": sorted(idx[3:5])} if trial % 2 else {}
        use_filter = trial >= 3
        want = run(10 ** 6, groups, use_filter)
        assert [p for p, _, _ in want] == sorted(p for p, _, _ in want)
        for batch_size in (1, 2, 5):
            assert run(batch_size, groups, use_filter) == want


def test_scenario_without_the_pseudo_contig_all():
    """A per-contig ploidy map need not define `all`: upstream resolves only the contigs that occur (calling.rs:632-718)."""
    sc = Scenario.from_yaml("""
species:
  heterozygosity: 0.001
  ploidy:
    chr1: 2
    chrX: 1
samples:
  s:
    resolution: 0.1
events:
  het: "s:0.5"
  hom: "s:1.0"
""")
    assert sc.is_contig_dependent()
    _, b2 = synth.tumor_normal(6, seed=12, depth=20)
    one = LocusBatch(1, b2.read_offsets[::2].copy(), dict(b2.columns), b2.read_flags, b2.locus_flags)
    recs = _records_from_batch(one)
    contigs = ["chr1", "chr1", "chrX", "chrX", "chr1", "chrX"]
    for r, c in zip(recs, contigs):
        r["chrom"] = c
    built = []

    def factory(flat):
        built.append(flat)
        return EmuEngine(flat)
    w = calling.call_generic(sc, {"s": recs}, engine_factory=factory)
    assert [c.chrom for c in w.calls] == contigs and len(built) == 2
    x = [c for c in w.calls if c.chrom == "chrX"]
    assert all(set(c.event_probs) == {"absent", "hom", "artifact"} or "het" in c.event_probs for c in x)
    want = emu.call_batch(sc.for_contig("chrX").flatten(), one, afd_capacity=128)
    flat_x = sc.for_contig("chrX").flatten()
    for i, c in enumerate(w.calls):
        if c.chrom == "chrX":
            assert [c.event_probs[e] for e in flat_x.event_names] == want.log_posteriors[i, :-1].tolist()


def test_variant_specific_priors_reach_the_engine():
    """INFO HETEROZYGOSITY / SOMATIC_EFFECTIVE_MUTATION_RATE of a record (PHRED; calling.rs:472-494) override the
    scenario's rates for that record: the operator API must hand them to the engine like the batch columns do."""
    sc = Scenario.from_yaml(synth.SIMPLE_PEDIGREE_YAML)
    flat = sc.flatten()
    _, b = synth.pedigree(8, seed=41, depth=14)
    from tests.util import SNV_FLAGS
    b.locus_flags[:] = SNV_FLAGS  # the re-encoded records are all A>G SNVs
    names = sc.sample_names

    def one(s):
        starts, ends = b.read_offsets[s:-1:3], b.read_offsets[s + 1::3]
        idx = np.concatenate([np.arange(a, e) for a, e in zip(starts, ends)])
        offs = np.concatenate([[0], np.cumsum(ends - starts)])
        return LocusBatch(1, offs, {k: v[idx] for k, v in b.columns.items()}, b.read_flags[idx], b.locus_flags)
    records = {n: _records_from_batch(one(s)) for s, n in enumerate(names)}
    het = np.full(8, np.nan, dtype=np.float32)
    het[[1, 4, 5]] = [13.5, 40.25, 7.0]
    for recs in records.values():
        for i, r in enumerate(recs):
            if not np.isnan(het[i]):
                r["info"]["HETEROZYGOSITY"] = float(het[i])
    w = calling.call_generic(sc, records, engine_factory=EmuEngine)
    with_override = LocusBatch(3, b.read_offsets, b.columns, b.read_flags, b.locus_flags, None, None, het, None)
    want = oracle.call_batch(flat, with_override, afd_capacity=128)
    plain = oracle.call_batch(flat, b, afd_capacity=128)
    assert not np.allclose(want.log_posteriors[[1, 4, 5]], plain.log_posteriors[[1, 4, 5]])  # the override matters
    for i, c in enumerate(w.calls):
        got = [c.event_probs[e] for e in flat.event_names]
        assert np.allclose(got, want.log_posteriors[i, :-1], rtol=0, atol=1e-9, equal_nan=True)


def test_observation_file_version_and_float_tags(tmp_path):
    rec = _records_from_batch(LocusBatch(1, *(lambda b: (b.read_offsets[::2].copy(), dict(b.columns), b.read_flags,
                                                         b.locus_flags))(synth.tumor_normal(1, seed=3, depth=4)[1])))[0]
    info = ";".join("%s=%s" % (k, ",".join(map(str, v))) for k, v in rec["info"].items())
    body = "chr1\t100\t.\tA\tG\t.\t.\t%s;HETEROZYGOSITY=30.1;SOMATIC_EFFECTIVE_MUTATION_RATE=.\n" % info
    good = tmp_path / "obs.vcf"
    good.write_text("##fileformat=VCFv4.2\n##varlociraptor_observation_format_version=%s\n#CHROM\tPOS\n%s" % (
        obs_codec.OBSERVATION_FORMAT_VERSION, body))
    r = obs_codec.parse_observation_vcf(str(good))[0]
    assert r["info"]["HETEROZYGOSITY"] == 30.1 and r["info"]["SOMATIC_EFFECTIVE_MUTATION_RATE"] is None
    b = obs_codec.batch_from_records([[r]])
    assert b.locus_heterozygosity_phred is not None and abs(float(b.locus_heterozygosity_phred[0]) - 30.1) < 1e-5
    assert b.locus_semr_phred is None
    for header in ("##varlociraptor_observation_format_version=14\n", ""):
        bad = tmp_path / "old.vcf"
        bad.write_text("##fileformat=VCFv4.2\n%s#CHROM\tPOS\n%s" % (header, body))
        with pytest.raises(obs_codec.InvalidObservationFormat):
            obs_codec.parse_observation_vcf(str(bad))
    r2 = dict(r, info=dict(r["info"], STRAND=r["info"]["STRAND"][:-4]))  # one read fewer in one vector
    import struct
    with pytest.raises((obs_codec.InvalidObservationFormat, ValueError, IndexError, struct.error)):
        obs_codec.decode_record(r2["info"])
    shorter = dict(r["info"])
    n = len(obs_codec.decode_enum(shorter["STRAND"]))
    shorter["STRAND"] = obs_codec.encode_enum([0] * (n - 1))  # a well-formed vector with one read fewer
    with pytest.raises(obs_codec.InvalidObservationFormat):
        obs_codec.decode_record(shorter)
