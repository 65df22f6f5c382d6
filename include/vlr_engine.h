/*
 * vlr_engine.h — C-ABI of the B200-native per-locus posterior engine.
 *
 * This is the drop-in boundary for varlociraptor's `call variants` inner loop.
 * The reference has no FFI; the seam this ABI replaces is the Rust call
 *
 *     let m = model.compute(event_universe.iter().cloned(), &data);
 *                                      (src/calling/variants/calling.rs:760)
 *
 * together with its readers in `Caller::call_record` / `Caller::sample_infos`
 * (calling.rs:762-813, 844-937) and the per-record preparation that feeds it
 * (`preprocess_record`, calling.rs:457-630; `configure_model`, calling.rs:632-718).
 * A Rust host keeps the CLI, scenario grammar and BCF I/O, flattens its
 * `grammar::Scenario` / `VAFTree`s into `vlr_scenario_t` once per contig, packs
 * batches of records into `vlr_batch_t` (structure of arrays) and gets back
 * event posteriors, MAP allele frequencies and allele frequency distributions.
 *
 * Conventions
 *  - plain C, no C++/torch types; every pointer is caller-owned memory;
 *  - every entry point returns a vlr_status_t, never throws or aborts;
 *    model invariants that `panic!` in the reference (NaN, prior > 0, ...)
 *    become per-locus bits in `vlr_results_t::status`;
 *  - a context is bound to one CUDA device and one stream and must not be
 *    used from two threads at once; several contexts (one per GPU) may run
 *    concurrently.
 */
#ifndef VLR_ENGINE_H
#define VLR_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLR_ABI_VERSION 1
#define VLR_MAX_SAMPLES 8
#define VLR_MAX_EVENTS 24
#define VLR_MAX_TREE_DEPTH 24
#define VLR_N_ARTIFACT_CONFIGS 9 /* 0 = none, 1..8 = single-artifact configs */

typedef int32_t vlr_status_t;
enum {
    VLR_OK = 0,
    VLR_ERR_INVALID_ARGUMENT = 1,
    VLR_ERR_UNSUPPORTED = 2,
    VLR_ERR_CUDA = 3,
    VLR_ERR_OUT_OF_MEMORY = 4,
    VLR_ERR_NO_DEVICE = 5
};

/* ---- VAF tree (flattened grammar::vaftree::Node, src/grammar/vaftree.rs:58-88) ---- */
enum {
    VLR_NODE_SET = 0,     /* NodeKind::Sample with VAFSpectrum::Set   */
    VLR_NODE_RANGE = 1,   /* NodeKind::Sample with VAFSpectrum::Range */
    VLR_NODE_LFC = 2,     /* NodeKind::Log2FoldChange                 */
    VLR_NODE_VARIANT = 3, /* NodeKind::Variant                        */
    VLR_NODE_TRUE = 4,
    VLR_NODE_FALSE = 5
};

/* utils/comparison.rs ComparisonOperator */
enum { VLR_CMP_EQ = 0, VLR_CMP_GT = 1, VLR_CMP_GE = 2, VLR_CMP_LT = 3, VLR_CMP_LE = 4, VLR_CMP_NE = 5 };

typedef struct {
    int32_t kind;        /* VLR_NODE_* */
    int32_t sample;      /* SET/RANGE: sample index; LFC: sample_a */
    int32_t sample_b;    /* LFC */
    int32_t cmp;         /* LFC: VLR_CMP_* */
    int32_t first_child; /* children are nodes[first_child .. first_child+n_children) */
    int32_t n_children;
    int32_t vaf_offset;  /* SET: vafs are set_vafs[vaf_offset .. vaf_offset+n_vafs), ascending */
    int32_t n_vafs;
    int32_t left_exclusive;  /* RANGE */
    int32_t right_exclusive; /* RANGE */
    int32_t variant_positive; /* VARIANT */
    int32_t refmask;          /* VARIANT: IUPAC set as bit mask A=1,C=2,G=4,T=8 */
    int32_t altmask;
    int32_t _pad;
    double start; /* RANGE */
    double end;   /* RANGE */
    double lfc_value; /* LFC predicate value */
} vlr_node_t;

/* One scenario event (model::Event, src/variants/model/mod.rs:34-39) before the
 * plain/artifact split; the engine derives the artifact twin itself
 * (calling.rs:655-687). */
typedef struct {
    char name[64];
    int32_t first_root; /* roots are nodes[first_root .. first_root+n_roots) */
    int32_t n_roots;
    int32_t has_artifact_twin; /* 0 for the absent event, 1 otherwise */
    int32_t _pad;
} vlr_event_t;

enum { VLR_SPECTRUM_SET = 0, VLR_SPECTRUM_RANGE = 1 };
typedef struct {
    int32_t kind;
    int32_t vaf_offset, n_vafs; /* into set_vafs */
    int32_t left_exclusive, right_exclusive;
    int32_t _pad;
    double start, end;
} vlr_spectrum_t;

enum { VLR_INHERIT_NONE = 0, VLR_INHERIT_MENDELIAN = 1, VLR_INHERIT_CLONAL = 2, VLR_INHERIT_SUBCLONAL = 3 };

/* grammar::Sample + calling.rs SampleInfos (calling.rs:1130-1213) */
typedef struct {
    double resolution;             /* grammar/mod.rs:427 */
    double contamination_fraction; /* valid iff contamination_by >= 0 */
    double germline_mutation_rate; /* NaN = none */
    double somatic_effective_mutation_rate; /* NaN = none */
    int32_t contamination_by;      /* -1 = not contaminated */
    int32_t uniform_prior;         /* sample declares `universe` (grammar/mod.rs:499) */
    int32_t ploidy;                /* -1 = none */
    int32_t inheritance;           /* VLR_INHERIT_* */
    int32_t parent_a, parent_b;    /* mendelian: (a,b); clonal/subclonal: a */
    int32_t clonal_somatic;
    int32_t universe_offset, n_universe; /* contig universe spectra (for the prior) */
    int32_t _pad;
} vlr_sample_t;

typedef struct {
    int32_t abi_version;
    int32_t n_samples;
    int32_t n_events;
    int32_t n_nodes;
    int32_t n_set_vafs;
    int32_t n_spectra;
    const vlr_sample_t* samples;
    const vlr_event_t* events;
    const vlr_node_t* nodes;
    const double* set_vafs;
    const vlr_spectrum_t* spectra;
    double heterozygosity;  /* species heterozygosity (linear), NaN = none */
    double vtf_indel, vtf_mnv, vtf_sv; /* VariantTypeFraction, grammar/mod.rs:375-415 */
    int32_t full_prior;     /* !is_absent_only (calling.rs:1086) */
    int32_t _pad;
} vlr_scenario_t;

/* ---- per-read flag word ---- */
#define VLR_RF_STRAND_SHIFT 0      /* 2 bits: 0 Forward, 1 Reverse, 2 Both, 3 None */
#define VLR_RF_ORIENT_SHIFT 2      /* 4 bits: bio-types SequenceReadPairOrientation:
                                      0 F1R2, 1 F2R1, 2 R1F2, 3 R2F1, 4 F1F2, 5 R1R2, 6 F2F1, 7 R2R1, 8 None */
#define VLR_RF_READPOS_MAJOR (1u << 6)
#define VLR_RF_SOFTCLIPPED (1u << 7)
#define VLR_RF_PAIRED (1u << 8)
#define VLR_RF_MAX_MAPQ (1u << 9)
#define VLR_RF_ALTLOCUS_SHIFT 10   /* 2 bits: 0 Major, 1 Some, 2 None */
#define VLR_RF_HAS_HOMOPOLYMER_LEN (1u << 12)
#define VLR_RF_HOMOPOLYMER_LEN_SHIFT 16 /* 8 bits, two's complement i8 */

/* ---- per-locus flag word (WorkItem, calling.rs:943-962) ---- */
#define VLR_LF_CHECK_ROB (1u << 0)
#define VLR_LF_CHECK_SB (1u << 1)
#define VLR_LF_CHECK_RPB (1u << 2)
#define VLR_LF_CHECK_SCB (1u << 3)
#define VLR_LF_CHECK_HE (1u << 4)
#define VLR_LF_CHECK_ALB (1u << 5)
#define VLR_LF_FILTER_NONSTANDARD (1u << 6) /* is_snv_or_mnv && !omit_read_orientation_bias (calling.rs:595-602) */
#define VLR_LF_VARTYPE_SHIFT 8  /* 2 bits: 0 fraction 1 (SNV, METH, ...), 1 indel, 2 MNV, 3 SV */
#define VLR_LF_HAS_SNV (1u << 10)
#define VLR_LF_REFBASE_SHIFT 16 /* 8 bits ASCII */
#define VLR_LF_ALTBASE_SHIFT 24

/* Batch of loci, structure of arrays. Reads of locus i, sample s are rows
 * read_offsets[i*S+s] .. read_offsets[i*S+s+1] of every per-read column, in
 * pileup order. Probabilities are natural-log, f32 (lossless: on disk every
 * value is f16 or f32, utils/mod.rs:448-474). */
typedef struct {
    int64_t n_loci;
    int64_t n_reads;
    const int64_t* read_offsets; /* [n_loci*S + 1] */
    const float* prob_mapping;
    const float* prob_ref;
    const float* prob_alt;
    const float* prob_missed_allele;
    const float* prob_sample_alt;
    const float* prob_double_overlap;
    const float* prob_hit_base;
    const uint32_t* read_flags;
    const float* prob_homopolymer_artifact; /* optional (NULL); NaN = None */
    const float* prob_homopolymer_variant;  /* optional (NULL); NaN = None */
    const uint32_t* locus_flags;            /* [n_loci] */
    const float* locus_heterozygosity_phred; /* optional; NaN = none (INFO HETEROZYGOSITY) */
    const float* locus_semr_phred;           /* optional; NaN = none */
} vlr_batch_t;

/* per-locus status bits */
#define VLR_ST_MARGINAL_ZERO (1u << 0)     /* all events have probability zero */
#define VLR_ST_NAN (1u << 1)               /* a NaN appeared (reference: assert!/panic!) */
#define VLR_ST_OVERSHOOT (1u << 2)         /* cap_numerical_overshoot would panic */
#define VLR_ST_PRIOR_POSITIVE (1u << 3)    /* prior > 0 (prior.rs:378) */
#define VLR_ST_GRID_OVERFLOW (1u << 4)     /* adaptive grid exceeded engine capacity */
#define VLR_ST_BASE_EVENTS_OVERFLOW (1u << 5) /* base-event log exceeded capacity: MAP/AFD incomplete */
#define VLR_ST_AFD_TRUNCATED (1u << 6)     /* AFD exceeded afd_capacity */
#define VLR_ST_NO_MAP (1u << 7)            /* no MAP estimate (sample_infos returned None) */
#define VLR_ST_IS_ARTIFACT (1u << 8)       /* artifact beats every other event (calling.rs:801-803) */
#define VLR_ST_SINGLETON_ADJUSTED (1u << 9)      /* Hint::AdjustedSingletonEvidence */
#define VLR_ST_FILTERED_NONSTANDARD (1u << 10)   /* Hint::FilteredNonStandardAlignments */
#define VLR_ST_WORKSPACE_OVERFLOW (1u << 11)     /* locus has more reads than vlr_ctx_reserve() provided for
                                                    (device-pointer entry only); results of the locus are invalid */

typedef struct {
    /* [n_loci][n_events+1]: ln posterior of every plain event in scenario order,
     * then ln P(artifact) = ln_sum_exp(artifact twins) (calling.rs:772-799) */
    double* log_posteriors;
    double* log_marginal;  /* [n_loci], optional */
    double* map_vaf;       /* [n_loci][S]; NaN if VLR_ST_NO_MAP */
    int32_t* map_config;   /* [n_loci]; artifact config of the MAP base event (0 none) */
    int32_t* best_event;   /* [n_loci]; index into the event universe: 2*e (plain) or 2*e+1 (twin) */
    uint32_t* status;      /* [n_loci] */
    uint32_t* n_base_events; /* [n_loci], optional: number of joint (prior x likelihood) evaluations */
    /* allele frequency distribution (calling.rs:891-928), optional (afd_capacity = 0) */
    int32_t afd_capacity;  /* entries per locus and sample */
    int32_t _pad;
    int32_t* afd_count;    /* [n_loci][S] */
    double* afd_vaf;       /* [n_loci][S][afd_capacity] */
    double* afd_logp;      /* [n_loci][S][afd_capacity] ln posterior */
} vlr_results_t;

typedef struct vlr_ctx vlr_ctx_t;

/* Replaces Caller::configure_model (calling.rs:632-718): build the event
 * universe (absent + plain + artifact twins) for one contig on `device`. */
vlr_status_t vlr_ctx_create(const vlr_scenario_t* scenario, int32_t device, vlr_ctx_t** out);
void vlr_ctx_destroy(vlr_ctx_t* ctx);

/* Replaces model.compute + call_record + sample_infos (calling.rs:720-937) for a
 * batch of records. Host buffers in, host buffers out; H2D, kernels and D2H
 * run on the context's stream; returns after the results are in host memory. */
vlr_status_t vlr_call_batch(vlr_ctx_t* ctx, const vlr_batch_t* batch, vlr_results_t* results);

/* ---- packed batch: the same columns, losslessly encoded for the host -> device link ------------------------------
 * vlr_call_batch() moves 32 bytes per read over PCIe, which bounds the end-to-end rate of the all-Set pipeline
 * (pedigrees) and, from ~8 M loci/s on, of the tumor-normal pipeline. Most observation columns hold few distinct
 * values: prob_mapping is a function of MAPQ, prob_hit_base of the read length, prob_double_overlap is mostly ln 0,
 * prob_sample_alt is ln 1 for SNVs, SNV base-quality emissions come from <= 94 quality values; on disk the reference
 * shrinks them value by value (MiniLogProb, utils/mod.rs:448-474: f16 where that keeps the integer part). A packed
 * column keeps the exact f32 BIT PATTERNS (so results are bitwise those of vlr_call_batch) in the smallest of:
 *   F32     4 bytes per read, as in vlr_batch_t
 *   F16     2 bytes per read: every value is exactly representable as IEEE half
 *   DICT16  2 bytes per read: code into a dictionary of <= 65536 bit patterns
 *   DICT8   1 byte per read:  code into a dictionary of <= 256 bit patterns
 *   CONST   0 bytes per read: every row holds dict[0]
 * The engine widens the chunk on the device (vlr_unpack_kernel: one pass, HBM-bound) in front of the pre-pass. */
enum { VLR_ENC_F32 = 0, VLR_ENC_F16 = 1, VLR_ENC_DICT16 = 2, VLR_ENC_DICT8 = 3, VLR_ENC_CONST = 4 };
enum {
    VLR_COL_PROB_MAPPING = 0, VLR_COL_PROB_REF = 1, VLR_COL_PROB_ALT = 2, VLR_COL_PROB_MISSED_ALLELE = 3,
    VLR_COL_PROB_SAMPLE_ALT = 4, VLR_COL_PROB_DOUBLE_OVERLAP = 5, VLR_COL_PROB_HIT_BASE = 6, VLR_COL_READ_FLAGS = 7,
    VLR_N_PACKED_COLUMNS = 8
};
typedef struct {
    int32_t encoding;     /* VLR_ENC_* */
    int32_t n_dict;       /* DICT16 / DICT8: entries of `dict`; CONST: 1; else 0 */
    const void* data;     /* [n_reads] float / uint32 (F32), uint16 half bits (F16), uint16 (DICT16), uint8 (DICT8); NULL (CONST) */
    const uint32_t* dict; /* [n_dict] bit patterns of the f32 values (of the flag words for VLR_COL_READ_FLAGS) */
} vlr_column_t;
typedef struct {
    int64_t n_loci;
    int64_t n_reads;
    const int64_t* read_offsets;                 /* as in vlr_batch_t */
    vlr_column_t columns[VLR_N_PACKED_COLUMNS];  /* VLR_COL_*; read_flags: F32 means plain uint32 words, F16 is invalid */
    const float* prob_homopolymer_artifact;      /* optional, plain (indel records only) */
    const float* prob_homopolymer_variant;
    const uint32_t* locus_flags;
    const float* locus_heterozygosity_phred;
    const float* locus_semr_phred;
} vlr_packed_batch_t;

/* Encodes `batch` (host buffers) column by column into the smallest lossless encoding; the packed arrays are
 * page-locked (vlr_host_alloc) and owned by `*out` until vlr_packed_batch_free(). Pointers of `batch` that need no
 * encoding (read_offsets, locus columns, homopolymer columns) are borrowed, not copied: they must outlive `*out`.
 * `n_threads` <= 0: all host threads. This is host work a producer does while it decodes observation records
 * (obs_codec): one hash lookup per value. */
vlr_status_t vlr_pack_batch(const vlr_batch_t* batch, int32_t n_threads, vlr_packed_batch_t** out);
void vlr_packed_batch_free(vlr_packed_batch_t* packed);
/* Bytes vlr_call_batch_packed() moves host -> device for `packed`. */
int64_t vlr_packed_batch_bytes(const vlr_packed_batch_t* packed, int32_t n_samples);
/* vlr_call_batch() on a packed batch: same chunking, same results bit for bit. */
vlr_status_t vlr_call_batch_packed(vlr_ctx_t* ctx, const vlr_packed_batch_t* packed, vlr_results_t* results);

/* Same, with every pointer in `batch` and `results` a DEVICE pointer on the
 * context's device. Asynchronous on `cuda_stream` (a cudaStream_t, 0 = the
 * context's own stream); the caller synchronises. Calls on one context share
 * its workspace: a call issued on another stream than the previous one waits
 * (on the device, through an event) until that one has finished. */
vlr_status_t vlr_call_batch_device(vlr_ctx_t* ctx, const vlr_batch_t* batch, vlr_results_t* results,
                                   void* cuda_stream);

/* Capacity of the per-warp read-coefficient workspace, in reads per locus summed over samples (default 4096).
 * vlr_call_batch() sizes it from the batch itself; callers of vlr_call_batch_device() whose loci can be deeper
 * reserve once up front (the call allocates, so it is not asynchronous). */
vlr_status_t vlr_ctx_reserve(vlr_ctx_t* ctx, int64_t max_reads_per_locus);

/* Page-locked host memory for batch columns and result arrays: vlr_call_batch() overlaps its chunked host<->device
 * copies with compute only when the caller's buffers are pinned (pageable memory still works, staged by the driver).
 * The pages are placed on the NUMA node of the CURRENT CUDA device when the host exposes it (call cudaSetDevice /
 * create the context first; VLR_NUMA=0 keeps the default placement). */
void* vlr_host_alloc(size_t bytes);
void vlr_host_free(void* p);

/* Measured fp64 FMA throughput of `device` in TFLOP/s (2 flops per FMA): a register-resident DFMA microbenchmark,
 * the denominator of the fp64 roofline this path is bound by (HBM is not the bound here). */
vlr_status_t vlr_measure_fp64_peak(int32_t device, double* tflops);

/* ---- contamination estimator (src/estimation/contamination.rs), SURVEY.md §8(f)-4 --------------------------------
 * The second Bayesian model `estimate contamination` runs over the calls of the `denovo`/`other` scenario
 * (contamination.rs:436-449): events = expected maximum somatic VAF x contamination grid, likelihood = sum over the
 * kept VariantObservations of the observation's allele frequency distribution, interpolated at the VAF the event
 * predicts (Likelihood::compute :163-186, VariantObservation::pdf :84-115), marginal = ln_sum_exp over the rows of
 * a Simpson rule over the contamination grid (Marginal::compute :213-240). The host keeps what the reference does per
 * call (VariantObservation::new :44-82: P(denovo) >= 0.95 and an AFD present) and the output tables. */
typedef struct {
    int64_t n_obs;
    const double* prob_denovo;       /* [n_obs] ln P(denovo) */
    const double* max_posterior_vaf; /* [n_obs] MAP allele frequency of the sample */
    const int64_t* afd_offsets;      /* [n_obs+1] rows of afd_vaf/afd_logp per observation */
    const double* afd_vaf;           /* ascending within an observation (BTreeMap order) */
    const double* afd_logp;          /* ln posterior density */
    int32_t n_grid;                  /* contamination grid points per row, odd, >= 3 (reference: 101) */
    int32_t n_max_vafs;              /* rows, <= 8 (reference: 4); n_grid * n_max_vafs <= 1024 */
    const double* expected_max_somatic_vaf; /* [n_max_vafs] (reference: 0.25, 0.5, 0.75, 1.0) */
    const double* ln_prior;          /* [n_grid] Prior::prob (contamination.rs:137-147) at linspace(0, 1, n_grid) */
} vlr_contamination_input_t;

typedef struct {
    double* ln_posterior;  /* [n_max_vafs][n_grid] joint - marginal (ModelInstance::event_posteriors) */
    double* ln_likelihood; /* [n_max_vafs][n_grid], optional (NULL) */
    double* ln_marginal;   /* [1] */
    double* max_vaf;       /* [1] VAFDist::max_vaf (contamination.rs:249-258), optional (NULL) */
} vlr_contamination_output_t;

/* Host buffers in, host buffers out, on `device`. n_obs = 0 is valid (likelihood ln_one everywhere). */
vlr_status_t vlr_contamination_posterior(int32_t device, const vlr_contamination_input_t* in,
                                         vlr_contamination_output_t* out);
/* Same with every pointer a DEVICE pointer on `device` (e.g. AFDs packed from device-resident results);
 * asynchronous on `cuda_stream` except for the scratch allocation; the caller synchronises. */
vlr_status_t vlr_contamination_posterior_device(int32_t device, const vlr_contamination_input_t* in,
                                                vlr_contamination_output_t* out, void* cuda_stream);

/* VariantObservation::new (contamination.rs:44-82) for a whole batch of calls without leaving the device: `results`
 * holds DEVICE pointers as written by vlr_call_batch_device() with afd_capacity > 0 for `n_loci` calls of a context with
 * `n_samples` samples and `n_events` events. Keeps the calls that have a MAP estimate outside the artifact events and
 * exp(ln P(`denovo_event`)) >= `min_prob` (reference: 0.95), and packs {ln P(denovo), MAP allele frequency of `sample`,
 * its allele frequency distribution} into the CSR columns vlr_contamination_posterior_device() reads. All output
 * pointers are DEVICE buffers of the caller: prob_denovo, max_posterior_vaf, kept_loci (optional, the call index of
 * every observation) [n_loci], afd_offsets [n_loci + 1], afd_vaf / afd_logp [n_loci * afd_capacity]. `n_obs` is a HOST
 * pointer: the call waits on `cuda_stream` for this one number (the posterior's launch geometry depends on it). */
vlr_status_t vlr_contamination_gather_device(int32_t device, const vlr_results_t* results, int64_t n_loci,
                                             int32_t n_samples, int32_t n_events, int32_t sample, int32_t denovo_event,
                                             double min_prob, double* prob_denovo, double* max_posterior_vaf,
                                             int64_t* afd_offsets, double* afd_vaf, double* afd_logp, int64_t* kept_loci,
                                             int64_t* n_obs, void* cuda_stream);

/* Number of kernels the last vlr_call_batch* launched (for bench accounting). */
int64_t vlr_last_launch_count(const vlr_ctx_t* ctx);
/* The context's stream (cudaStream_t) so callers can time with CUDA events. */
void* vlr_ctx_stream(const vlr_ctx_t* ctx);
const char* vlr_last_error(const vlr_ctx_t* ctx);
const char* vlr_status_string(vlr_status_t status);
int32_t vlr_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VLR_ENGINE_H */
