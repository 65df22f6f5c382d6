"""Observation wire format -> SoA columns.

Decodes the per-record INFO arrays `varlociraptor preprocess variants` writes
(src/calling/variants/preprocessing/mod.rs:921-1038, read side :818-919) straight into
the per-read columns of `LocusBatch`, without materialising one struct per read.

Wire format (observation format version 15, preprocessing/mod.rs:810): every INFO tag is a
bincode-1.3 byte stream (little endian, fixed-width ints), zero-padded to even length, split
into u16s and stored as BCF integers. `Vec<T>` = u64 length + items; `MiniLogProb`
(src/utils/mod.rs:448-474) = u32 variant (0 = f16 bits, 1 = f32 bits) + payload; plain enums =
u32 variant index; `Option<T>` = u8 tag + payload; `bv::BitVec<u8>` = Option<Box<[u8]>>
(u8 tag, u64 #blocks, blocks) + u64 bit length.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import abi
from .batch import LocusBatch

OBSERVATION_FORMAT_VERSION = "15"

_PROB_TAGS = {
    "PROB_MAPPING": "prob_mapping", "PROB_REF": "prob_ref", "PROB_ALT": "prob_alt",
    "PROB_MISSED_ALLELE": "prob_missed_allele", "PROB_SAMPLE_ALT": "prob_sample_alt",
    "PROB_DOUBLE_OVERLAP": "prob_double_overlap", "PROB_HIT_BASE": "prob_hit_base",
}


class _Reader:
    def __init__(self, ints: Sequence[int]):
        self.b = np.asarray(ints, dtype=np.int64).astype(np.uint16).astype("<u2").tobytes()
        self.i = 0

    def u8(self) -> int:
        v = self.b[self.i]
        self.i += 1
        return v

    def u32(self) -> int:
        v = struct.unpack_from("<I", self.b, self.i)[0]
        self.i += 4
        return v

    def u64(self) -> int:
        v = struct.unpack_from("<Q", self.b, self.i)[0]
        self.i += 8
        return v

    def mini_logprob(self) -> float:
        variant = self.u32()
        if variant == 0:
            v = np.frombuffer(self.b, dtype="<f2", count=1, offset=self.i)[0]
            self.i += 2
        elif variant == 1:
            v = np.frombuffer(self.b, dtype="<f4", count=1, offset=self.i)[0]
            self.i += 4
        else:
            raise ValueError("invalid MiniLogProb variant %d" % variant)
        return float(v)


def decode_mini_logprobs(ints: Sequence[int]) -> np.ndarray:
    r = _Reader(ints)
    n = r.u64()
    return np.array([r.mini_logprob() for _ in range(n)], dtype=np.float32)


def decode_optional_mini_logprobs(ints: Sequence[int]) -> np.ndarray:
    """Vec<Option<MiniLogProb>>; None -> NaN."""
    r = _Reader(ints)
    n = r.u64()
    out = np.full(n, np.nan, dtype=np.float32)
    for k in range(n):
        if r.u8():
            out[k] = r.mini_logprob()
    return out


def decode_enum(ints: Sequence[int]) -> np.ndarray:
    r = _Reader(ints)
    n = r.u64()
    return np.array([r.u32() for _ in range(n)], dtype=np.uint32)


def decode_optional_i8(ints: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    r = _Reader(ints)
    n = r.u64()
    has = np.zeros(n, dtype=bool)
    val = np.zeros(n, dtype=np.int8)
    for k in range(n):
        if r.u8():
            has[k] = True
            val[k] = np.frombuffer(r.b, dtype=np.int8, count=1, offset=r.i)[0]
            r.i += 1
    return has, val


def decode_optional_u32(ints: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """Vec<Option<u32>> (THIRD_ALLELE_EVIDENCE, preprocessing/mod.rs:818-1038) -> (is_some, value)."""
    r = _Reader(ints)
    n = r.u64()
    has = np.zeros(n, dtype=bool)
    val = np.zeros(n, dtype=np.uint32)
    for k in range(n):
        if r.u8():
            has[k] = True
            val[k] = r.u32()
    return has, val


def decode_bitvec(ints: Sequence[int]) -> np.ndarray:
    r = _Reader(ints)
    blocks = b""
    if r.u8():
        nblocks = r.u64()
        blocks = r.b[r.i:r.i + nblocks]
        r.i += nblocks
    nbits = r.u64()
    arr = np.frombuffer(blocks, dtype=np.uint8)
    idx = np.arange(nbits)
    if nbits == 0:
        return np.zeros(0, dtype=bool)
    return ((arr[idx // 8] >> (idx % 8)) & 1).astype(bool)


def decode_record(info: Dict[str, Sequence[int]]):
    """One record's INFO arrays -> (columns dict of f32 arrays, read_flags u32, hom_artifact, hom_variant)."""
    cols = {dst: decode_mini_logprobs(info[tag]) for tag, dst in _PROB_TAGS.items()}
    n = len(cols["prob_mapping"])
    strand = decode_enum(info["STRAND"])
    orient = decode_enum(info["READ_ORIENTATION"])
    readpos = decode_enum(info["READ_POSITION"])  # Major = 0, Some = 1
    altlocus = decode_enum(info["ALT_LOCUS"])
    softclipped = decode_bitvec(info["SOFTCLIPPED"])
    paired = decode_bitvec(info["PAIRED"])
    max_mapq = decode_bitvec(info["IS_MAX_MAPQ"])
    lengths = {k: len(v) for k, v in cols.items()}
    lengths.update(STRAND=len(strand), READ_ORIENTATION=len(orient), READ_POSITION=len(readpos), ALT_LOCUS=len(altlocus),
                   SOFTCLIPPED=len(softclipped), PAIRED=len(paired), IS_MAX_MAPQ=len(max_mapq))
    if any(m != n for m in lengths.values()):  # one entry per read in every vector (preprocessing/mod.rs:869-910)
        raise InvalidObservationFormat("per-read vectors of a record differ in length: %r" % lengths)
    flags = (strand << abi.RF_STRAND_SHIFT) | (orient << abi.RF_ORIENT_SHIFT) | (altlocus << abi.RF_ALTLOCUS_SHIFT)
    flags = flags.astype(np.uint32)
    flags |= np.where(readpos == 0, abi.RF_READPOS_MAJOR, 0).astype(np.uint32)
    flags |= np.where(softclipped[:n], abi.RF_SOFTCLIPPED, 0).astype(np.uint32)
    flags |= np.where(paired[:n], abi.RF_PAIRED, 0).astype(np.uint32)
    flags |= np.where(max_mapq[:n], abi.RF_MAX_MAPQ, 0).astype(np.uint32)
    hart = hvar = None
    # is_homopolymer_indel = the artifact tag is present and non-empty (preprocessing/mod.rs:874)
    if "PROB_HOMOPOLYMER_ARTIFACT_OBSERVABLE" in info:
        hart = decode_optional_mini_logprobs(info["PROB_HOMOPOLYMER_ARTIFACT_OBSERVABLE"])
        hvar = decode_optional_mini_logprobs(info["PROB_HOMOPOLYMER_VARIANT_OBSERVABLE"])
        has, val = decode_optional_i8(info["HOMOPOLYMER_INDEL_LEN"])
        flags |= np.where(has, abi.RF_HAS_HOMOPOLYMER_LEN, 0).astype(np.uint32)
        flags |= (val.view(np.uint8).astype(np.uint32) << abi.RF_HOMOPOLYMER_LEN_SHIFT)
    return cols, flags, hart, hvar


class InvalidObservationFormat(ValueError):
    """errors::Error::InvalidObservationFormat: the file was written by another observation format version."""


_FLOAT_INFO_TAGS = ("HETEROZYGOSITY", "SOMATIC_EFFECTIVE_MUTATION_RATE")  # PHRED floats (calling.rs:472-494)


def parse_observation_vcf(path: str, check_version: bool = True) -> List[dict]:
    """Text dump (`bcftools view`) of an observation BCF -> list of records with integer INFO arrays. Like
    `Caller::call` (calling.rs:324-339) it refuses files whose `varlociraptor_observation_format_version` header is
    missing or differs from the version this codec decodes: another bincode layout would be mis-decoded silently."""
    out = []
    version = None
    with open(path) as f:
        for line in f:
            if line.startswith("##varlociraptor_observation_format_version="):
                version = line.rstrip("\n").split("=", 1)[1]
            if line.startswith("#"):
                continue
            t = line.rstrip("\n").split("\t")
            info = {}
            flags = set()
            for kv in t[7].split(";"):
                if "=" in kv:
                    k, v = kv.split("=", 1)
                    if k in _FLOAT_INFO_TAGS:
                        x = v.split(",")[0]  # all records output by preprocess are single allele
                        info[k] = None if x == "." else float(x)
                        continue
                    try:
                        info[k] = [int(x) for x in v.split(",")]
                    except ValueError:
                        info[k] = v
                else:
                    flags.add(kv)
            out.append({"chrom": t[0], "pos": int(t[1]), "id": t[2], "ref": t[3], "alt": t[4], "info": info,
                        "flags": flags})
    if check_version and version != OBSERVATION_FORMAT_VERSION:
        raise InvalidObservationFormat("%s: observation format version %s, this codec reads version %s" % (
            path, "missing" if version is None else version, OBSERVATION_FORMAT_VERSION))
    return out


def locus_flags_for(ref: str, alt: str, is_homopolymer_indel: bool, imprecise: bool = False,
                    omit_strand_bias=False, omit_read_orientation_bias=False, omit_read_position_bias=False,
                    omit_softclip_bias=False, omit_homopolymer_artifact_detection=False,
                    omit_alt_locus_bias=False, vartype: Optional[int] = None) -> int:
    """The per-record switches of `Caller::preprocess_record` (src/calling/variants/calling.rs:513-566)."""
    if len(ref) == 1 and len(alt) == 1:
        is_snv_or_mnv, snv = True, True
    elif len(ref) == len(alt):
        is_snv_or_mnv, snv = True, False
    else:
        is_snv_or_mnv, snv = False, False
    precise = not imprecise
    f = 0
    if is_snv_or_mnv and not omit_read_orientation_bias and precise:
        f |= abi.LF_CHECK_ROB
    if not omit_strand_bias and precise:
        f |= abi.LF_CHECK_SB
    if is_snv_or_mnv and not omit_read_position_bias and precise:
        f |= abi.LF_CHECK_RPB
    if is_snv_or_mnv and not omit_softclip_bias and precise:
        f |= abi.LF_CHECK_SCB
    if is_homopolymer_indel and not omit_homopolymer_artifact_detection:
        f |= abi.LF_CHECK_HE
    if not omit_alt_locus_bias:
        f |= abi.LF_CHECK_ALB
    if is_snv_or_mnv and not omit_read_orientation_bias:
        f |= abi.LF_FILTER_NONSTANDARD
    if vartype is None:
        # model::VariantType of the record (src/variants/model/mod.rs) -> variant-type fraction class
        if alt in ("<DEL>", "<INS>", "<REP>"):
            vartype = abi.VARTYPE_INDEL
        elif alt in ("<INV>", "<DUP>", "<BND>") or "[" in alt or "]" in alt:
            vartype = abi.VARTYPE_SV
        elif alt.startswith("<"):
            vartype = abi.VARTYPE_SNV  # METH, REF, ...: fraction 1 (grammar/mod.rs:403-412)
        elif len(ref) == 1 and len(alt) == 1:
            vartype = abi.VARTYPE_SNV
        elif len(ref) == len(alt):
            vartype = abi.VARTYPE_MNV
        else:
            vartype = abi.VARTYPE_INDEL
    f |= vartype << abi.LF_VARTYPE_SHIFT
    if snv:
        f |= abi.LF_HAS_SNV | (ord(ref) << abi.LF_REFBASE_SHIFT) | (ord(alt) << abi.LF_ALTBASE_SHIFT)
    return f


def batch_from_records(per_sample_records: List[List[dict]], **omit) -> LocusBatch:
    """Lock-step records of S samples (outer list = samples in sample-index order; an entry may be None
    for a sample without observations, calling.rs:605-607) -> LocusBatch."""
    S = len(per_sample_records)
    L = len(per_sample_records[0])
    cols = {k: [] for k in abi.BATCH_F32_COLUMNS}
    flags, harts, hvars, lflags = [], [], [], []
    het, semr = np.full(L, np.nan, dtype=np.float32), np.full(L, np.nan, dtype=np.float32)
    offsets = [0]
    any_h = False
    for i in range(L):
        hom = False
        first = None
        for s in range(S):
            rec = per_sample_records[s][i]
            if rec is None:
                offsets.append(offsets[-1])
                continue
            first = first or rec
            c, f, hart, hvar = decode_record(rec["info"])
            n = len(f)
            for k in abi.BATCH_F32_COLUMNS:
                cols[k].append(c[k])
            flags.append(f)
            if hart is not None and len(hart) > 0:
                hom = True
                any_h = True
                harts.append(hart)
                hvars.append(hvar)
            else:
                harts.append(np.full(n, np.nan, dtype=np.float32))
                hvars.append(np.full(n, np.nan, dtype=np.float32))
            offsets.append(offsets[-1] + n)
        lflags.append(locus_flags_for(first["ref"], first["alt"], hom, "IMPRECISE" in first["flags"], **omit))
        # variant-specific priors of the first record with observations (calling.rs:472-494): PHRED, None = not given
        h, r = first["info"].get("HETEROZYGOSITY"), first["info"].get("SOMATIC_EFFECTIVE_MUTATION_RATE")
        if isinstance(h, (int, float)):
            het[i] = h
        if isinstance(r, (int, float)):
            semr[i] = r

    def cat(lst, dt):
        return np.concatenate(lst) if lst else np.zeros(0, dtype=dt)
    return LocusBatch(S, np.array(offsets, dtype=np.int64), {k: cat(v, np.float32) for k, v in cols.items()},
                      cat(flags, np.uint32), np.array(lflags, dtype=np.uint32),
                      cat(harts, np.float32) if any_h else None, cat(hvars, np.float32) if any_h else None,
                      het if np.any(~np.isnan(het)) else None, semr if np.any(~np.isnan(semr)) else None)


# ----------------------------------------------------------------------------- encoding (write_observations mirror)
def _to_ints(body: bytearray) -> List[int]:
    """bincode bytes -> the u16-as-integer INFO array (preprocessing/mod.rs:978-1000: zero-pad to even length)."""
    if len(body) % 2:
        body = body + b"\x00"
    return np.frombuffer(bytes(body), dtype="<u2").astype(np.int64).tolist()


def encode_mini_logprobs(values: Sequence[float]) -> List[int]:
    """Vec<MiniLogProb> with `MiniLogProb::new` (src/utils/mod.rs:458-466): f16 if < -10 and the f16 projection keeps
    the integer floor, else f32."""
    body = bytearray(np.uint64(len(values)).tobytes())
    with np.errstate(over="ignore", invalid="ignore"):
        for v in np.asarray(values, dtype=np.float64):
            h = np.float16(v)
            if v < -10.0 and np.floor(np.float64(h)) == np.floor(v):
                body += np.uint32(0).tobytes() + h.tobytes()
            else:
                body += np.uint32(1).tobytes() + np.float32(v).tobytes()
    return _to_ints(body)


def encode_optional_mini_logprobs(values: Sequence[float]) -> List[int]:
    """Vec<Option<MiniLogProb>>; NaN -> None."""
    body = bytearray(np.uint64(len(values)).tobytes())
    with np.errstate(over="ignore", invalid="ignore"):
        for v in np.asarray(values, dtype=np.float64):
            if np.isnan(v):
                body += b"\x00"
                continue
            body += b"\x01"
            h = np.float16(v)
            if v < -10.0 and np.floor(np.float64(h)) == np.floor(v):
                body += np.uint32(0).tobytes() + h.tobytes()
            else:
                body += np.uint32(1).tobytes() + np.float32(v).tobytes()
    return _to_ints(body)


def encode_enum(values: Sequence[int]) -> List[int]:
    body = bytearray(np.uint64(len(values)).tobytes())
    for v in values:
        body += np.uint32(int(v)).tobytes()
    return _to_ints(body)


def encode_bitvec(bits: Sequence[bool]) -> List[int]:
    """bv::BitVec<u8>: Option<Box<[u8]>> blocks + u64 bit length (an empty vector has no block storage)."""
    bits = np.asarray(bits, dtype=np.uint8)
    n = len(bits)
    if n == 0:
        return _to_ints(bytearray(b"\x00" + np.uint64(0).tobytes()))
    blocks = np.packbits(bits, bitorder="little").tobytes()
    return _to_ints(bytearray(b"\x01" + np.uint64(len(blocks)).tobytes() + blocks + np.uint64(n).tobytes()))


def encode_optional_i8(has: Sequence[bool], values: Sequence[int]) -> List[int]:
    body = bytearray(np.uint64(len(has)).tobytes())
    for h, v in zip(has, values):
        body += (b"\x01" + np.int8(v).tobytes()) if h else b"\x00"
    return _to_ints(body)


def encode_record(cols: Dict[str, np.ndarray], flags: np.ndarray, hart: Optional[np.ndarray] = None,
                  hvar: Optional[np.ndarray] = None) -> Dict[str, List[int]]:
    """SoA columns of one pileup -> the INFO arrays `write_observations` emits (preprocessing/mod.rs:921-1038).
    FRAGMENT_ID and THIRD_ALLELE_EVIDENCE are not carried by the engine's batch layout and are not produced."""
    flags = np.asarray(flags, dtype=np.uint32)
    info = {tag: encode_mini_logprobs(cols[dst]) for tag, dst in _PROB_TAGS.items()}
    info["STRAND"] = encode_enum((flags >> abi.RF_STRAND_SHIFT) & 3)
    info["READ_ORIENTATION"] = encode_enum((flags >> abi.RF_ORIENT_SHIFT) & 15)
    info["READ_POSITION"] = encode_enum(np.where(flags & abi.RF_READPOS_MAJOR, 0, 1))
    info["ALT_LOCUS"] = encode_enum((flags >> abi.RF_ALTLOCUS_SHIFT) & 3)
    info["SOFTCLIPPED"] = encode_bitvec((flags & abi.RF_SOFTCLIPPED) != 0)
    info["PAIRED"] = encode_bitvec((flags & abi.RF_PAIRED) != 0)
    info["IS_MAX_MAPQ"] = encode_bitvec((flags & abi.RF_MAX_MAPQ) != 0)
    if hart is not None and np.any(~np.isnan(hart)):  # only written if any read has homopolymer info (:1018-1035)
        info["PROB_HOMOPOLYMER_ARTIFACT_OBSERVABLE"] = encode_optional_mini_logprobs(hart)
        info["PROB_HOMOPOLYMER_VARIANT_OBSERVABLE"] = encode_optional_mini_logprobs(hvar)
        has = (flags & abi.RF_HAS_HOMOPOLYMER_LEN) != 0
        val = ((flags >> abi.RF_HOMOPOLYMER_LEN_SHIFT) & 0xff).astype(np.uint8).view(np.int8)
        info["HOMOPOLYMER_INDEL_LEN"] = encode_optional_i8(has, val)
    return info
