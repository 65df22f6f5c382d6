"""ctypes mirror of include/vlr_engine.h (the C-ABI boundary).

Only declarations live here: struct layouts, constants and helpers that turn numpy
arrays into the plain pointers the ABI takes. Nothing in this module computes.
"""
import ctypes as C

import numpy as np

VLR_ABI_VERSION = 1
VLR_MAX_SAMPLES = 8
VLR_MAX_EVENTS = 24
VLR_N_ARTIFACT_CONFIGS = 9

# node kinds
NODE_SET, NODE_RANGE, NODE_LFC, NODE_VARIANT, NODE_TRUE, NODE_FALSE = range(6)
CMP_EQ, CMP_GT, CMP_GE, CMP_LT, CMP_LE, CMP_NE = range(6)
SPECTRUM_SET, SPECTRUM_RANGE = 0, 1
INHERIT_NONE, INHERIT_MENDELIAN, INHERIT_CLONAL, INHERIT_SUBCLONAL = range(4)

# read flags
RF_STRAND_SHIFT = 0
RF_ORIENT_SHIFT = 2
RF_READPOS_MAJOR = 1 << 6
RF_SOFTCLIPPED = 1 << 7
RF_PAIRED = 1 << 8
RF_MAX_MAPQ = 1 << 9
RF_ALTLOCUS_SHIFT = 10
RF_HAS_HOMOPOLYMER_LEN = 1 << 12
RF_HOMOPOLYMER_LEN_SHIFT = 16
STRAND_FORWARD, STRAND_REVERSE, STRAND_BOTH, STRAND_NONE = range(4)
ORIENT_F1R2, ORIENT_F2R1, ORIENT_R1F2, ORIENT_R2F1, ORIENT_F1F2, ORIENT_R1R2, ORIENT_F2F1, ORIENT_R2R1, ORIENT_NONE = range(9)
ALTLOCUS_MAJOR, ALTLOCUS_SOME, ALTLOCUS_NONE = range(3)

# locus flags
LF_CHECK_ROB = 1 << 0
LF_CHECK_SB = 1 << 1
LF_CHECK_RPB = 1 << 2
LF_CHECK_SCB = 1 << 3
LF_CHECK_HE = 1 << 4
LF_CHECK_ALB = 1 << 5
LF_FILTER_NONSTANDARD = 1 << 6
LF_VARTYPE_SHIFT = 8
LF_HAS_SNV = 1 << 10
LF_REFBASE_SHIFT = 16
LF_ALTBASE_SHIFT = 24
VARTYPE_SNV, VARTYPE_INDEL, VARTYPE_MNV, VARTYPE_SV = range(4)

# status bits
ST_MARGINAL_ZERO = 1 << 0
ST_NAN = 1 << 1
ST_OVERSHOOT = 1 << 2
ST_PRIOR_POSITIVE = 1 << 3
ST_GRID_OVERFLOW = 1 << 4
ST_BASE_EVENTS_OVERFLOW = 1 << 5
ST_AFD_TRUNCATED = 1 << 6
ST_NO_MAP = 1 << 7
ST_IS_ARTIFACT = 1 << 8
ST_SINGLETON_ADJUSTED = 1 << 9
ST_FILTERED_NONSTANDARD = 1 << 10
ST_WORKSPACE_OVERFLOW = 1 << 11

ARTIFACT_CONFIG_NAMES = ["none", "ALB", "HE", "SCB", "RPB", "ROB_F1R2", "ROB_F2R1", "SB_FWD", "SB_REV"]


class Node(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("sample", C.c_int32), ("sample_b", C.c_int32), ("cmp", C.c_int32),
        ("first_child", C.c_int32), ("n_children", C.c_int32), ("vaf_offset", C.c_int32), ("n_vafs", C.c_int32),
        ("left_exclusive", C.c_int32), ("right_exclusive", C.c_int32), ("variant_positive", C.c_int32),
        ("refmask", C.c_int32), ("altmask", C.c_int32), ("_pad", C.c_int32),
        ("start", C.c_double), ("end", C.c_double), ("lfc_value", C.c_double),
    ]


class Event(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("first_root", C.c_int32), ("n_roots", C.c_int32),
                ("has_artifact_twin", C.c_int32), ("_pad", C.c_int32)]


class Spectrum(C.Structure):
    _fields_ = [("kind", C.c_int32), ("vaf_offset", C.c_int32), ("n_vafs", C.c_int32),
                ("left_exclusive", C.c_int32), ("right_exclusive", C.c_int32), ("_pad", C.c_int32),
                ("start", C.c_double), ("end", C.c_double)]


class Sample(C.Structure):
    _fields_ = [
        ("resolution", C.c_double), ("contamination_fraction", C.c_double),
        ("germline_mutation_rate", C.c_double), ("somatic_effective_mutation_rate", C.c_double),
        ("contamination_by", C.c_int32), ("uniform_prior", C.c_int32), ("ploidy", C.c_int32),
        ("inheritance", C.c_int32), ("parent_a", C.c_int32), ("parent_b", C.c_int32),
        ("clonal_somatic", C.c_int32), ("universe_offset", C.c_int32), ("n_universe", C.c_int32),
        ("_pad", C.c_int32),
    ]


class Scenario(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_samples", C.c_int32), ("n_events", C.c_int32), ("n_nodes", C.c_int32),
        ("n_set_vafs", C.c_int32), ("n_spectra", C.c_int32),
        ("samples", C.POINTER(Sample)), ("events", C.POINTER(Event)), ("nodes", C.POINTER(Node)),
        ("set_vafs", C.POINTER(C.c_double)), ("spectra", C.POINTER(Spectrum)),
        ("heterozygosity", C.c_double), ("vtf_indel", C.c_double), ("vtf_mnv", C.c_double), ("vtf_sv", C.c_double),
        ("full_prior", C.c_int32), ("_pad", C.c_int32),
    ]


_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


class Batch(C.Structure):
    _fields_ = [
        ("n_loci", C.c_int64), ("n_reads", C.c_int64), ("read_offsets", _i64p),
        ("prob_mapping", _f32p), ("prob_ref", _f32p), ("prob_alt", _f32p), ("prob_missed_allele", _f32p),
        ("prob_sample_alt", _f32p), ("prob_double_overlap", _f32p), ("prob_hit_base", _f32p),
        ("read_flags", _u32p), ("prob_homopolymer_artifact", _f32p), ("prob_homopolymer_variant", _f32p),
        ("locus_flags", _u32p), ("locus_heterozygosity_phred", _f32p), ("locus_semr_phred", _f32p),
    ]


# packed batch (vlr_packed_batch_t): lossless per-column encodings for the host -> device link
ENC_F32, ENC_F16, ENC_DICT16, ENC_DICT8, ENC_CONST = 0, 1, 2, 3, 4
ENC_NAMES = ["f32", "f16", "dict16", "dict8", "const"]
ENC_BYTES = [4, 2, 2, 1, 0]
PACKED_COLUMNS = ["prob_mapping", "prob_ref", "prob_alt", "prob_missed_allele", "prob_sample_alt",
                  "prob_double_overlap", "prob_hit_base", "read_flags"]


class Column(C.Structure):  # vlr_column_t
    _fields_ = [("encoding", C.c_int32), ("n_dict", C.c_int32), ("data", C.c_void_p), ("dict", _u32p)]


class PackedBatch(C.Structure):  # vlr_packed_batch_t
    _fields_ = [
        ("n_loci", C.c_int64), ("n_reads", C.c_int64), ("read_offsets", _i64p),
        ("columns", Column * 8),
        ("prob_homopolymer_artifact", _f32p), ("prob_homopolymer_variant", _f32p),
        ("locus_flags", _u32p), ("locus_heterozygosity_phred", _f32p), ("locus_semr_phred", _f32p),
    ]


class Results(C.Structure):
    _fields_ = [
        ("log_posteriors", _f64p), ("log_marginal", _f64p), ("map_vaf", _f64p), ("map_config", _i32p),
        ("best_event", _i32p), ("status", _u32p), ("n_base_events", _u32p),
        ("afd_capacity", C.c_int32), ("_pad", C.c_int32),
        ("afd_count", _i32p), ("afd_vaf", _f64p), ("afd_logp", _f64p),
    ]


BATCH_F32_COLUMNS = ["prob_mapping", "prob_ref", "prob_alt", "prob_missed_allele", "prob_sample_alt",
                     "prob_double_overlap", "prob_hit_base"]
BATCH_OPTIONAL_F32_COLUMNS = ["prob_homopolymer_artifact", "prob_homopolymer_variant"]


class ContaminationInput(C.Structure):  # vlr_contamination_input_t
    _fields_ = [
        ("n_obs", C.c_int64), ("prob_denovo", _f64p), ("max_posterior_vaf", _f64p), ("afd_offsets", _i64p),
        ("afd_vaf", _f64p), ("afd_logp", _f64p), ("n_grid", C.c_int32), ("n_max_vafs", C.c_int32),
        ("expected_max_somatic_vaf", _f64p), ("ln_prior", _f64p),
    ]


class ContaminationOutput(C.Structure):  # vlr_contamination_output_t
    _fields_ = [("ln_posterior", _f64p), ("ln_likelihood", _f64p), ("ln_marginal", _f64p), ("max_vaf", _f64p)]


def ptr(arr, ctype):
    """Pointer to a C-contiguous numpy array (or NULL for None)."""
    if arr is None:
        return C.cast(None, C.POINTER(ctype))
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(C.POINTER(ctype))


def devptr(addr, ctype):
    """Typed pointer from a raw (device) address."""
    return C.cast(C.c_void_p(int(addr) if addr else None), C.POINTER(ctype))
