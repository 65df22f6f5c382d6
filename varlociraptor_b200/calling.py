"""Host-side mirror of the reference's operator API for `call variants`, over the CUDA engine.

Same names and argument meaning as src/calling/variants/calling.rs:

  * `call_generic(scenario, observations, omit_*..., output, log_each_record, call_processor, candidate_filter,
    propagate_info_fields, full_prior)`  (calling.rs:1022-1116)
  * `CallProcessor.setup / process_call / finalize`  (calling.rs:964-976), `CallWriter` (calling.rs:978-1006)
  * `CandidateFilter.filter(work_item, sample_names) -> bool`  (calling.rs:1008-1011), `DefaultCandidateFilter`
  * `Call` with the fields `Call::write_final_record` emits (src/calling/variants/mod.rs:178-576): `PROB_<EVENT>` as
    PHRED f32, `AF` f32, `AFD` "vaf=phred" text, the bias labels SB/ROB/RPB/SCB/HE/ALB, `DP`, `HINTS`.

Differences by design: records are processed in batches (`batch_size` loci per engine call) instead of one by one;
calls are still delivered to the processor in input order, and the candidate filter still sees every work item with
its pileups before it is packed. Observation input here is the text form of the observation BCF (`bcftools view`);
binary BCF I/O stays with the host application (out of scope).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Iterable, Iterator, List, Optional, Sequence

import numpy as np

from . import abi, obs_codec
from .batch import CallResults, LocusBatch
from .scenario import Scenario

_PHRED = -10.0 / math.log(10.0)


def _kass_raftery_letter(m1: float, m2: float) -> str:
    """`bayes_factor_to_letter` (src/utils/mod.rs:158-167) of BayesFactor::new(m1, m2) = exp(m1 - m2)."""
    with np.errstate(over="ignore", invalid="ignore"):
        k = float(np.exp(np.float64(m1) - np.float64(m2)))
    if k <= 1.0:
        return "E" if k == 1.0 or abs(k - 1.0) <= 2.220446049250313e-16 else "N"
    if k <= 3.0:
        return "B"
    if k <= 20.0:
        return "P"
    if k <= 150.0:
        return "S"
    return "V"


def simple_observations(prob_alt, prob_ref, is_max_mapq, alt_allele: bool) -> str:
    """`fmt_simple_obs` (src/calling/variants/mod.rs:337-376): generalized CIGAR of Kass-Raftery letters of the reads
    favouring the alt (or the ref / a third) allele, upper case iff the read has the maximum MAPQ; most common
    first, 'E' entries last."""
    from collections import Counter
    items = []
    for pa, pr, mq in zip(prob_alt, prob_ref, is_max_mapq):
        keep = (pa > pr) if alt_allele else (pa <= pr)
        if not keep:
            continue
        letter = _kass_raftery_letter(pa, pr) if alt_allele else _kass_raftery_letter(pr, pa)
        items.append(letter.upper() if mq else letter.lower())
    if not items:
        return "."
    common = Counter(items).most_common()
    common.sort(key=lambda kv: 1 if kv[0].endswith("E") else 0)  # stable: keeps the count order
    return "".join("%d%s" % (n, it) for it, n in common)


def read_observation_summary(prob_alt, prob_ref, read_flags, third_allele=None) -> str:
    """FORMAT/OBS (src/calling/variants/mod.rs:277-333): generalized CIGAR over one code per read -
    max Bayes factor (`A`/`R` + Kass-Raftery letter or `E`, upper case iff maximum MAPQ), third-allele edit distance
    or '.', p(aired)/s(ingle), alt locus (#, *, .), strand (*, -, +, .), read orientation (>, <, *, !), read position
    (^, *), softclip ($, .), homopolymer error (*, .); most common first, then `E...` codes, then `N...` codes."""
    from collections import Counter
    items = []
    has3, val3 = third_allele if third_allele is not None else (None, None)
    for k, (pa, pr, f) in enumerate(zip(prob_alt, prob_ref, read_flags)):
        f = int(f)
        if pa > pr:      # bf_alt > bf_ref  <=>  prob_alt > prob_ref (exp is monotone; equal -> Equal)
            score = "A" + _kass_raftery_letter(pa, pr)
        elif pr > pa:
            score = "R" + _kass_raftery_letter(pr, pa)
        else:
            score = "E"
        score = score.upper() if f & abi.RF_MAX_MAPQ else score.lower()
        third = str(int(val3[k])) if has3 is not None and k < len(has3) and has3[k] else "."
        altlocus = "#*."[(f >> abi.RF_ALTLOCUS_SHIFT) & 3]
        strand = "+-*."[(f >> abi.RF_STRAND_SHIFT) & 3]          # Forward, Reverse, Both, None
        orient = {0: ">", 1: "<", 8: "*"}.get((f >> abi.RF_ORIENT_SHIFT) & 15, "!")
        hom = bool(f & abi.RF_HAS_HOMOPOLYMER_LEN) and ((f >> abi.RF_HOMOPOLYMER_LEN_SHIFT) & 0xff) != 0
        items.append("%s%s%s%s%s%s%s%s%s" % (score, third, "p" if f & abi.RF_PAIRED else "s", altlocus, strand, orient,
                                               "^" if f & abi.RF_READPOS_MAJOR else "*",
                                               "$" if f & abi.RF_SOFTCLIPPED else ".", "*" if hom else "."))
    if not items:
        return "."
    common = Counter(items).most_common()
    common.sort(key=lambda kv: 2 if kv[0].startswith("N") else (1 if kv[0].startswith("E") else 0))  # stable
    return "".join("%d%s" % (n, it) for it, n in common)


def event_tag_name(event: str) -> str:
    """src/utils/mod.rs `event_tag_name`: PROB_<EVENT upper-cased>."""
    return "PROB_" + event.upper()


@dataclass
class SampleCall:
    """`SampleInfo` of a called record (src/calling/variants/mod.rs:578-600)."""
    allelefreq_estimate: float
    artifact: str            # abi.ARTIFACT_CONFIG_NAMES entry, "none" if the MAP is not an artifact
    vaf_dist: Optional[List[tuple]]  # [(vaf, ln posterior density)] ascending, None for artifact MAPs
    depth: int               # DP = round(sum(exp(prob_mapping)))
    saobs: str = "."         # SAOBS / SROBS: simplified observation summaries (mod.rs:337-379)
    srobs: str = "."
    obs: str = "."           # OBS: full per-read codes (mod.rs:277-333)
    oobs: int = 0            # OOBS: reads dropped by remove_nonstandard_alignments (pileup.rs:26-43, mod.rs:382)
    n_reads: int = -1        # reads left in the pileup (Pileup::is_empty decides `missing-data`); -1 = not recorded


@dataclass
class Call:
    chrom: str
    pos: int                 # 1-based
    ref: str
    alt: str
    event_probs: Dict[str, float] = field(default_factory=dict)  # event name (+ "artifact") -> ln posterior
    sample_info: List[Optional[SampleCall]] = field(default_factory=list)
    hints: List[str] = field(default_factory=list)
    status: int = 0

    @property
    def is_missing_data(self) -> bool:
        """mod.rs:424-431: no sample has a read left (after the non-standard alignment filter): the record is written
        with missing PROB_* / AF, DP 0 and the hint `missing-data`."""
        return all(si is None or si.n_reads == 0 for si in self.sample_info)

    def all_hints(self) -> List[str]:
        return self.hints + (["missing-data"] if self.is_missing_data else [])

    def info_fields(self) -> Dict[str, np.float32]:
        """PROB_* INFO values exactly as written: PHRED, absolute value, f32 (mod.rs:447-466); NaN = missing value."""
        if self.is_missing_data:
            return {event_tag_name(e): np.float32(np.nan) for e in self.event_probs}
        return {event_tag_name(e): np.float32(abs(_PHRED * p)) for e, p in self.event_probs.items()}

    def format_fields(self, sample: int) -> Dict[str, str]:
        si = self.sample_info[sample]
        if self.is_missing_data:  # mod.rs:559-571
            return {"DP": "0", "AF": ".", "SAOBS": ".", "SROBS": ".", "OBS": ".", "OOBS": ".", "SB": ".", "ROB": ".",
                    "RPB": ".", "AFD": "."}
        if si is None:
            return {"DP": ".", "AF": ".", "AFD": ".", "OBS": ".", "OOBS": "."}
        labels = {"SB": ".", "ROB": ".", "RPB": ".", "SCB": ".", "HE": ".", "ALB": "."}
        a = si.artifact
        if a == "SB_FWD":
            labels["SB"] = "+"
        elif a == "SB_REV":
            labels["SB"] = "-"
        elif a == "ROB_F1R2":
            labels["ROB"] = ">"
        elif a == "ROB_F2R1":
            labels["ROB"] = "<"
        elif a == "RPB":
            labels["RPB"] = "^"
        elif a == "SCB":
            labels["SCB"] = "$"
        elif a == "HE":
            labels["HE"] = "*"
        elif a == "ALB":
            labels["ALB"] = "*"
        afd = "." if si.vaf_dist is None else ",".join(
            "%.3f=%.2f" % (v, _PHRED * p) for v, p in si.vaf_dist)  # mod.rs:546-556
        out = {"DP": str(si.depth), "AF": "%g" % np.float32(si.allelefreq_estimate), "AFD": afd,
               "SAOBS": si.saobs, "SROBS": si.srobs, "OBS": si.obs, "OOBS": str(si.oobs)}
        out.update(labels)
        return out


@dataclass
class WorkItem:
    """What a CandidateFilter may look at (calling.rs:943-962): the record and its per-sample pileups (as a
    one-locus LocusBatch)."""
    index: int
    chrom: str
    pos: int
    ref: str
    alt: str
    pileups: LocusBatch
    locus_flags: int
    haplotype: Optional[str] = None  # EVENT / breakend group: later members reuse the first member's result
    third_allele_evidence: Optional[list] = None  # per sample (is_some, value) arrays or None; output only (OBS)


def haplotype_identifier(record: dict) -> Optional[str]:
    """`HaplotypeIdentifier::from` (src/variants/model/mod.rs:87-133): INFO/EVENT, else the sorted pair of record id and
    INFO/MATEID joined by '-' (breakend mates); a MATEID without a record id is an error upstream too."""
    info = record.get("info", {})
    event = info.get("EVENT")
    if event is not None:
        return str(event).split(",")[0]
    mateid = info.get("MATEID")
    if mateid is not None:
        recid = record.get("id", ".")
        if recid in (".", "", None):
            raise ValueError("breakend with MATEID but without record ID")  # errors::Error::BreakendMateidWithoutRecid
        return "-".join(sorted([str(recid), str(mateid).split(",")[0]]))
    return None


class CandidateFilter:
    def filter(self, work_item: WorkItem, sample_names: Sequence[str]) -> bool:  # noqa: A003
        raise NotImplementedError


class DefaultCandidateFilter(CandidateFilter):
    def filter(self, work_item, sample_names):  # noqa: A003
        return True


class CallProcessor:
    def setup(self, caller: "Caller") -> None:
        pass

    def process_call(self, call: Call, sample_names: Sequence[str]) -> None:
        raise NotImplementedError

    def finalize(self) -> None:
        pass


class CallWriter(CallProcessor):
    """Collects calls; `lines()` renders the INFO/FORMAT columns the way the final BCF shows them in text."""

    def __init__(self):
        self.calls: List[Call] = []
        self.sample_names: List[str] = []

    def setup(self, caller):
        self.sample_names = list(caller.sample_names)

    def process_call(self, call, sample_names):
        self.calls.append(call)

    def lines(self) -> List[str]:
        out = []
        for c in self.calls:
            info = ";".join("%s=%s" % (k, "." if np.isnan(v) else ("inf" if np.isinf(v) else "%g" % v))
                            for k, v in c.info_fields().items())
            if c.all_hints():
                info += ";HINTS=" + ",".join(c.all_hints())
            keys = ["DP", "AF", "SAOBS", "SROBS", "OBS", "OOBS", "SB", "ROB", "RPB", "SCB", "HE", "ALB", "AFD"]
            fmt = []
            for s in range(len(c.sample_info)):
                f = c.format_fields(s)
                fmt.append(":".join(f.get(k, ".") for k in keys))
            out.append("\t".join([c.chrom, str(c.pos), ".", c.ref, c.alt, ".", ".", info, ":".join(keys)] + fmt))
        return out


class Caller:
    """`Caller` (calling.rs:56-130, 320-455) over the CUDA engine."""

    def __init__(self, scenario: Scenario, observations: Dict[str, Iterable[dict]], call_processor: CallProcessor,
                 candidate_filter: CandidateFilter, omit: Dict[str, bool], full_prior: bool = False,
                 batch_size: int = 65536, afd_capacity: int = 128, device: int = 0, engine=None,
                 engine_factory=None):
        """`engine`: a ready engine for `scenario.flatten()` (tests). `engine_factory(flat_scenario)`: how to build one
        per configured model; default = the CUDA `PosteriorEngine`. A scenario whose ploidies or universes depend on
        the contig gets one model per contig, like `configure_model` upstream (calling.rs:632-718)."""
        scenario.full_prior = full_prior
        self.scenario = scenario
        self.sample_names = list(scenario.sample_names)
        for name in observations:
            if name not in self.sample_names:  # errors::Error::InvalidObservationSampleName
                raise ValueError("invalid observation sample name: %s" % name)
        self.observations = observations
        self.call_processor = call_processor
        self.candidate_filter = candidate_filter
        self.omit = omit
        self.batch_size = batch_size
        self.afd_capacity = afd_capacity
        if engine_factory is None:
            def engine_factory(flat):
                from .engine import PosteriorEngine
                return PosteriorEngine(flat, device=device)
        self._engine_factory = engine_factory
        self._per_contig = engine is None and scenario.is_contig_dependent()
        self._models: Dict[str, tuple] = {}
        self._model_by_signature: Dict[tuple, tuple] = {}
        if self._per_contig:
            # ploidies / universes are given per contig: like upstream (calling.rs:632-718) a model exists only for
            # the contigs that occur in the records; the scenario need not define the pseudo-contig "all"
            self._contig = None
            self.flat = self.engine = None
        else:
            self._contig = scenario.contig
            self.flat = scenario.flatten()
            if engine is None:
                engine = engine_factory(self.flat)
            self.engine = engine
            self._models[scenario.contig] = (self.flat, engine)
            self._model_by_signature[scenario.contig_signature()] = (self.flat, engine)
        self._haplotype_results: Dict[str, Optional[Call]] = {}

    def _configure_model(self, contig: str) -> None:
        """calling.rs:632-718: the event universe follows the contig (ploidy- and contig-specific universes)."""
        if contig not in self._models:
            sc = self.scenario.for_contig(contig)
            key = sc.contig_signature()
            if key not in self._model_by_signature:  # contigs with the same ploidies and universes share a model
                flat = sc.flatten()
                self._model_by_signature[key] = (flat, self._engine_factory(flat))
            self._models[contig] = self._model_by_signature[key]
        self.flat, self.engine = self._models[contig]
        self._contig = contig

    def _records(self) -> Iterator[List[Optional[dict]]]:
        """One record per sample in lock-step (calling.rs:353-398): same chrom/pos/alleles required."""
        its = [iter(self.observations[n]) if n in self.observations else None for n in self.sample_names]
        while True:
            recs: List[Optional[dict]] = []
            eof = 0
            for it in its:
                if it is None:
                    recs.append(None)
                    continue
                r = next(it, None)
                if r is None:
                    eof += 1
                recs.append(r)
            active = sum(1 for it in its if it is not None)
            if eof == active:
                return
            if eof:
                raise ValueError("observation files have different numbers of records")  # calling.rs:369-376
            first = next(r for r in recs if r is not None)
            for r in recs:
                if r is not None and (r["chrom"], r["pos"], r["ref"], r["alt"]) != \
                        (first["chrom"], first["pos"], first["ref"], first["alt"]):
                    raise ValueError("inconsistent observations: records differ at %s:%d" % (first["chrom"], first["pos"]))
            yield recs

    def call(self) -> None:
        self.call_processor.setup(self)
        pending: List[tuple] = []
        index = 0
        for recs in self._records():
            one = obs_codec.batch_from_records([[r] for r in recs], **self.omit)
            first = next(r for r in recs if r is not None)
            if self._per_contig and first["chrom"] != self._contig:
                if pending:  # records of the previous contig are called with its model
                    self._flush(pending)
                    pending = []
                self._configure_model(first["chrom"])
            item = WorkItem(index, first["chrom"], first["pos"], first["ref"], first["alt"], one,
                            int(one.locus_flags[0]), haplotype_identifier(first),
                            [obs_codec.decode_optional_u32(r["info"]["THIRD_ALLELE_EVIDENCE"])
                             if r is not None and "THIRD_ALLELE_EVIDENCE" in r["info"] else None for r in recs])
            index += 1
            if not self.candidate_filter.filter(item, self.sample_names):
                continue
            pending.append((item, one))
            if len(pending) >= self.batch_size:
                self._flush(pending)
                pending = []
        if pending:
            self._flush(pending)
        self.call_processor.finalize()

    def _flush(self, pending):
        # Breakend / haplotype groups (calling.rs:569-580, 726-741, 820-839): only the first member of a group is
        # computed; the others take its event probabilities and sample infos. Upstream drops the stored result after
        # the member with the last record index; groups are keyed here for the whole run, which is the same as long
        # as an identifier is not reused by a later, unrelated group.
        computed = []
        for k, (item, _) in enumerate(pending):
            if item.haplotype is None or item.haplotype not in self._haplotype_results:
                computed.append(k)
                if item.haplotype is not None:
                    self._haplotype_results[item.haplotype] = None  # claimed by this record, filled below
        slot = {k: i for i, k in enumerate(computed)}
        res: Optional[CallResults] = None
        if computed:
            batch = LocusBatch.concat([pending[k][1] for k in computed])
            res = self.engine.call_batch(batch, afd_capacity=self.afd_capacity)
        S = len(self.sample_names)
        names = self.flat.event_names
        for k, (item, one) in enumerate(pending):
            if k not in slot:
                first = self._haplotype_results[item.haplotype]
                call = Call(item.chrom, item.pos, item.ref, item.alt, dict(first.event_probs), list(first.sample_info),
                            status=first.status)
                self.call_processor.process_call(call, self.sample_names)
                continue
            i = slot[k]
            call = Call(item.chrom, item.pos, item.ref, item.alt, status=int(res.status[i]))
            for e, name in enumerate(names):
                call.event_probs[name] = float(res.log_posteriors[i, e])
            call.event_probs["artifact"] = float(res.log_posteriors[i, len(names)])
            if res.status[i] & abi.ST_SINGLETON_ADJUSTED:
                call.hints.append("adjusted-singleton-evidence")
            if res.status[i] & abi.ST_FILTERED_NONSTANDARD:
                call.hints.append("filtered-non-standard-alignments")
            no_map = bool(res.status[i] & abi.ST_NO_MAP)
            for s in range(S):
                if no_map:
                    call.sample_info.append(None)
                    continue
                lo, hi = one.read_offsets[s], one.read_offsets[s + 1]
                keep = np.ones(hi - lo, dtype=bool)
                if item.locus_flags & abi.LF_FILTER_NONSTANDARD:
                    o = (one.read_flags[lo:hi] >> abi.RF_ORIENT_SHIFT) & 15
                    keep = (o == 0) | (o == 1) | (o == 8)
                depth = int(round(float(np.exp(one.columns["prob_mapping"][lo:hi][keep].astype(np.float64)).sum())))
                pa = one.columns["prob_alt"][lo:hi][keep].astype(np.float64)
                pr = one.columns["prob_ref"][lo:hi][keep].astype(np.float64)
                rf = one.read_flags[lo:hi][keep]
                mq = (rf & abi.RF_MAX_MAPQ) != 0
                third = item.third_allele_evidence[s] if item.third_allele_evidence else None
                if third is not None:
                    third = (third[0][keep], third[1][keep])
                cfg = int(res.map_config[i])
                dist = None
                if cfg == 0 and res.afd_count is not None:
                    v, p = res.afd(i, s)
                    dist = list(zip(v.tolist(), p.tolist()))
                call.sample_info.append(SampleCall(float(res.map_vaf[i, s]), abi.ARTIFACT_CONFIG_NAMES[cfg], dist, depth,
                                                   simple_observations(pa, pr, mq, True),
                                                   simple_observations(pa, pr, mq, False),
                                                   read_observation_summary(pa, pr, rf, third),
                                                   int((~keep).sum()), int(keep.sum())))
            if item.haplotype is not None:
                self._haplotype_results[item.haplotype] = call
            self.call_processor.process_call(call, self.sample_names)


def call_generic(scenario: Scenario, observations: Dict[str, Iterable[dict]], omit_strand_bias: bool = False,
                 omit_read_orientation_bias: bool = False, omit_read_position_bias: bool = False,
                 omit_softclip_bias: bool = False, omit_homopolymer_artifact_detection: bool = False,
                 omit_alt_locus_bias: bool = False, output=None, log_each_record: bool = False,
                 call_processor: Optional[CallProcessor] = None, candidate_filter: Optional[CandidateFilter] = None,
                 propagate_info_fields: Sequence[str] = (), full_prior: bool = False, **engine_kwargs) -> CallProcessor:
    """`call_generic` (calling.rs:1022-1116). `observations` maps sample name -> records of its observation file
    (`obs_codec.parse_observation_vcf(path)`); a sample without an entry has zero coverage (calling.rs:605-607)."""
    omit = dict(omit_strand_bias=omit_strand_bias, omit_read_orientation_bias=omit_read_orientation_bias,
                omit_read_position_bias=omit_read_position_bias, omit_softclip_bias=omit_softclip_bias,
                omit_homopolymer_artifact_detection=omit_homopolymer_artifact_detection,
                omit_alt_locus_bias=omit_alt_locus_bias)
    cp = call_processor or CallWriter()
    cf = candidate_filter or DefaultCandidateFilter()
    Caller(scenario, observations, cp, cf, omit, full_prior=full_prior, **engine_kwargs).call()
    return cp


def call_tumor_normal(tumor_records, normal_records, purity: float = 1.0, **kwargs) -> CallProcessor:
    """`call variants tumor-normal` (src/cli.rs:1114-1196): the fixed two-sample scenario with tumor purity."""
    return call_generic(Scenario.tumor_normal(purity=purity), {"tumor": tumor_records, "normal": normal_records},
                        **kwargs)
