// engine_types.cuh — warp primitives, device views, LogProb / VAFRange helpers shared by every engine variant.
// See engine_core.cuh for the engine itself.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/vlr_engine.h"

#ifdef VLR_HOST_EMU
#define VLR_DEV inline
#define VLR_DEV_NOINLINE
struct alignas(16) double2 {
    double x, y;
};
#else
#define VLR_DEV __device__ __forceinline__
#define VLR_DEV_NOINLINE __device__ __noinline__
#endif

namespace vlrcore {

constexpr int WS_LEVELS = VLR_MAX_SAMPLES; // nested integration levels with scratch grids
#ifndef VLR_WARPS_PER_CTA
#define VLR_WARPS_PER_CTA 8
#endif
#ifndef VLR_MIN_CTAS
#define VLR_MIN_CTAS 2
#endif
constexpr int WARPS_PER_CTA = VLR_WARPS_PER_CTA;
constexpr int SM_READS = 208; // shared-memory coefficient arena per warp, in reads (6.5 KB): 2 x 100 reads fit
#ifndef VLR_HOST_EMU
extern __shared__ __align__(16) unsigned char vlr_smem[];
#endif
constexpr int NCFG = VLR_N_ARTIFACT_CONFIGS;
constexpr int GRID_CAP = 128;   // points per adaptive integration (res >= ~1e-5)
constexpr int LC_WAYS = 4;      // per-sample pileup-likelihood cache entries
constexpr int MAX_LFC_NODES = 32;
constexpr int AFD_TMP = 512;

constexpr double NUMERICAL_EPSILON = 1e-3;        // utils/mod.rs:41
constexpr double LN_05 = -0.6931471805599453;     // ln 0.5 (utils/mod.rs:45-47)
constexpr double LN_2 = 0.6931471805599453;
constexpr double LN_095 = -0.05129329438755058;   // ln 0.95 (utils/mod.rs:49-51)
constexpr double LN_3 = 1.0986122886681098;       // Kass-Raftery thresholds 3, 20, 150 in log space
constexpr double LN_20 = 2.995732273553991;
constexpr double LN_150 = 5.0106352940962555;

// ------------------------------------------------------------------------------------------------ warp layer
#ifdef VLR_HOST_EMU
constexpr int LANES = 1;
VLR_DEV int lane_id() { return 0; }
VLR_DEV void warp_sync() {}
VLR_DEV int w_sum_i(int v) { return v; }
VLR_DEV unsigned w_or_u(unsigned v) { return v; }
VLR_DEV int w_max_i(int v) { return v; }
VLR_DEV double w_sum_d(double v) { return v; }
VLR_DEV double w_mul_d(double v) { return v; }
VLR_DEV double w_max_d(double v) { return v; }
VLR_DEV bool w_any(bool p) { return p; }
VLR_DEV double w_bcast_d(double v, int) { return v; }
VLR_DEV int d_hi(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return (int)(u >> 32);
}
VLR_DEV int d_lo(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return (int)(u & 0xffffffffu);
}
VLR_DEV double d_make(int hi, int lo) {
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x;
    memcpy(&x, &u, 8);
    return x;
}
#else
// A "logical warp" of LANES lanes (32, or 16 = two loci share one physical warp's instruction stream: the kernel is
// instruction-fetch bound, so halving the streams per locus is worth more than the lanes). All collectives below are
// restricted to the logical warp through its member mask.
#ifndef VLR_LANES
#define VLR_LANES 32
#endif
constexpr int LANES = VLR_LANES;
VLR_DEV int lane_id() { return (int)(threadIdx.x & (LANES - 1)); }
VLR_DEV int group_in_cta() { return (int)(threadIdx.x / LANES); }
VLR_DEV unsigned group_shift() { return (threadIdx.x & 31u) & ~(unsigned)(LANES - 1); }
VLR_DEV unsigned gmask() { return LANES == 32 ? 0xffffffffu : (((1u << (LANES & 31)) - 1u) << group_shift()); }
#define FULL (vlrcore::gmask())
VLR_DEV void warp_sync() { __syncwarp(FULL); }
VLR_DEV int w_sum_i(int v) { return __reduce_add_sync(FULL, v); }
VLR_DEV unsigned w_or_u(unsigned v) { return __reduce_or_sync(FULL, v); }
VLR_DEV int w_max_i(int v) { return __reduce_max_sync(FULL, v); }
// xor butterflies: a+b == b+a bitwise, so every lane ends with the identical value (needed for uniform control flow)
VLR_DEV double w_sum_d(double v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
VLR_DEV double w_mul_d(double v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v *= __shfl_xor_sync(FULL, v, o);
    return v;
}
VLR_DEV double w_max_d(double v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
VLR_DEV bool w_any(bool p) { return __any_sync(FULL, p) != 0; }
VLR_DEV double w_bcast_d(double v, int src) { return __shfl_sync(FULL, v, src, LANES); }
// ballot of the logical warp with bit i = lane i of the logical warp
VLR_DEV unsigned w_ballot(bool p) { return __ballot_sync(FULL, p) >> group_shift(); }
VLR_DEV int d_hi(double x) { return __double2hiint(x); }
VLR_DEV int d_lo(double x) { return __double2loint(x); }
VLR_DEV double d_make(int hi, int lo) { return __hiloint2double(hi, lo); }
#endif

VLR_DEV double neg_inf() { return -INFINITY; }

// Streaming loads for the batch columns: every input byte is consumed within microseconds of its first touch, so it
// is marked evict-first and does not push the warps' stack / scratch lines out of L2 (measured: 203 GB of DRAM traffic
// per 1M-locus launch with plain loads, almost all of it re-fetched / written-back spill lines).
#ifdef VLR_HOST_EMU
template <class T> VLR_DEV T ldin(const T* p) { return *p; }
#else
template <class T> VLR_DEV T ldin(const T* p) { return __ldcs(p); }
#endif

// ------------------------------------------------------------------------------------------------ device views
// Context-level memo of Prior::compute for all-discrete VAF vectors (prior.rs:718-736 keeps an LRU(1000) per contig):
// the value depends only on the scenario, the variant type and the VAF vector, so it is shared by all loci of a context.
// Open addressing; an entry goes 0 (empty) -> 1 (being written by the warp that won the CAS) -> 2 (readable).
constexpr int PRIOR_TAB_N = 1024; // power of two
struct PriorTabEntry {
    unsigned state;
    unsigned side_status; // status bits the computation raised (they must reach every locus that uses the entry)
    int vartype, pad;
    double vaf[VLR_MAX_SAMPLES];
    double value;
};

struct DevScenario {
    PriorTabEntry* prior_tab; // NULL = no table
    int S, E, n_nodes, n_set_vafs, n_spectra, full_prior, all_uniform, n_lfc_nodes;
    int events_overlap; // two events may share a VAF combination (scenario_prep.h): MAP candidates are offered to every event
    int pad_;
    const vlr_sample_t* samples;
    const vlr_event_t* events;
    const vlr_node_t* nodes;
    const double* set_vafs;
    const vlr_spectrum_t* spectra;
    const int* lfc_nodes; // ordinal -> node index (LFC nodes only)
    const int* lfc_ordinal; // node index -> ordinal (or -1)
    double heterozygosity; // linear, NaN = none
    double vtf[4];         // variant type fraction by VLR_LF_VARTYPE class
};

struct DevBatch {
    int64_t n_loci;
    int64_t read_base; // first row held by the column pointers (chunked transfers)
    const int64_t* read_offsets;
    const float *pm, *pr, *pa, *pmiss, *psa, *pdo, *phb;
    const uint32_t* rflags;
    const float *hart, *hvar;
    const uint32_t* lflags;
    const float *het_phred, *semr_phred;
};

struct DevResults {
    double* log_post;
    double* log_marginal;
    double* map_vaf;
    int32_t* map_config;
    int32_t* best_event;
    uint32_t* status;
    uint32_t* n_base_events;
    int32_t afd_capacity;
    int32_t* afd_count;
    double* afd_vaf;
    double* afd_logp;
};

// Per-warp scratch in global memory (private to the warp, so it lives in L1/L2).
constexpr int BE_CAP = 4096; // recorded base events per locus (only when an AFD is requested)

constexpr int MT = 8;      // leaf integrations a warp advances concurrently (an outer batch has at most 7 abscissae)
constexpr int MSLOTS = 56; // (task, abscissa) slots of one step: at most 7 tasks x 7 points (+ padding)

struct WarpWs {
    double grid_x[WS_LEVELS][GRID_CAP];
    double grid_f[WS_LEVELS][GRID_CAP];
    double mgrid_x[MT][GRID_CAP]; // grids of the concurrent leaf integrations
    double mgrid_f[MT][GRID_CAP];
    double sort_x[GRID_CAP];
    double sort_f[GRID_CAP];
    double afd_x[AFD_TMP];
    double afd_p[AFD_TMP];
};

// ------------------------------------------------------------------------------------------------ wavefront plan
// Scenario-level description of the "two-level chain" shape the wavefront pipeline (engine_wave.cuh) serves: every event
// is  root(sample P: one VAF | Range) -> leaf(sample T: one VAF | Range)  (tumor-normal: P = normal, T = tumor), flat
// priors, no log2-fold-change / variant nodes. Built on the host (scenario_prep.h), passed to the kernels by value.
constexpr int WAVE_MAXE = 8;
struct WavePlan {
    int eligible;
    int P, T;        // root (parent) sample and leaf sample
    int outer_event; // the event whose root is a Range (nested integration), or -1
    int max_rounds;  // task rounds that complete every locus: 1 + outer iterations + 1
    int root_node[WAVE_MAXE], child_node[WAVE_MAXE];
};

// ------------------------------------------------------------------------------------------------ all-Set plan
// Scenario-level description of an "all-Set" scenario (engine_sets.cuh): every node of every event tree is a Set node,
// e.g. pedigrees (samples without an explicit universe have the allele frequencies their ploidy allows,
// grammar/mod.rs:539-543). GenericPosterior::density (generic.rs:294-330) then only sums joint probabilities of
// discrete allele frequency combinations: the trees are flattened on the host into the list of their root-to-leaf
// combinations ("leaves", in the order density() visits them) and the distinct pileup evaluations those need ("folds").
constexpr int SETS_MAXL = 192; // leaves of all events together
constexpr int SETS_MAXF = 48;  // distinct (sample, allele frequency, contaminant's allele frequency) pileup evaluations
constexpr int SETS_MAXS = 4;   // samples (the small engine variant serves <= 3)
struct SetsFold {
    double vaf, vaf_by;
    int sample, pad;
};
struct SetsLeaf {
    uint8_t fold[SETS_MAXS]; // pileup evaluation of each sample
    uint8_t event;           // scenario event the leaf belongs to
    uint8_t posmask;         // samples whose Set node on the path has only positive allele frequencies: the path is cut
                             // when such a sample is clearly reference (generic.rs:294-299)
    uint8_t discmask;        // samples pushed as discrete events on the path (all, unless a tree leaves a sample out)
    uint8_t pad;
};
struct SetsPlan {
    int eligible, n_folds, n_leaves, pad;
    int ev_first[VLR_MAX_EVENTS], ev_count[VLR_MAX_EVENTS]; // leaves of event e: [ev_first, ev_first + ev_count)
    const SetsFold* folds;
    const SetsLeaf* leaves;
    const double* leaf_vaf; // [n_leaves][S]
    // prior of every leaf per variant type class (Prior::compute depends on scenario, variant type and the allele
    // frequencies only; prior.rs:718-736 caches it per contig): filled by the first warps that need it
    double* prior_val;      // [4][n_leaves]
    uint32_t* prior_side;   // [4][n_leaves] status bits the computation raised
    int* prior_state;       // [4]: 0 = not computed, 2 = readable
};

// ------------------------------------------------------------------------------------------------ math
// Out-of-line fp64 transcendentals: CUDA inlines ~50-100 instructions per call site, and with dozens of call sites
// the kernel outgrows the instruction caches (ncu: stall_no_instruction dominated the first versions). One copy each.
VLR_DEV_NOINLINE double m_log(double x) { return log(x); }
VLR_DEV_NOINLINE double m_exp(double x) { return exp(x); }
VLR_DEV_NOINLINE double m_log1p(double x) { return log1p(x); }
VLR_DEV_NOINLINE double m_expm1(double x) { return expm1(x); }
VLR_DEV_NOINLINE double m_log2(double x) { return log2(x); }
VLR_DEV_NOINLINE double m_exp2(double x) { return exp2(x); }


// Correctly rounded log2 (double-double: ln m = 2 atanh((m - 1) / (m + 1)), 24 series terms). The log2-fold-change
// predicate `log2(a) - log2(b) >= v` (utils/log2_fold_change.rs:17-52) is evaluated exactly ON its threshold whenever an
// integration limit was inferred from the predicate (b = a / 2^v, generic.rs:148-174): the outcome is then decided by
// the last bit of log2. The reference (Rust f64::log2 -> glibc) is correctly rounded for 99.8 % of arguments; CUDA's
// log2 (1 ulp) agreed with it on only 74 % of such ties (measured, scripts/diag_lfc.py), this one on 99.9 %.
struct DD {
    double hi, lo;
};
VLR_DEV DD dd_two_sum(double a, double b) {
    const double s = a + b, bb = s - a;
    return DD{s, (a - (s - bb)) + (b - bb)};
}
VLR_DEV DD dd_quick(double a, double b) {
    const double s = a + b;
    return DD{s, b - (s - a)};
}
VLR_DEV DD dd_add(DD a, DD b) {
    DD s = dd_two_sum(a.hi, b.hi);
    const DD t = dd_two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = dd_quick(s.hi, s.lo);
    s.lo += t.lo;
    return dd_quick(s.hi, s.lo);
}
VLR_DEV DD dd_mul(DD a, DD b) {
    const double p = a.hi * b.hi;
    double e = fma(a.hi, b.hi, -p);
    e += a.hi * b.lo + a.lo * b.hi;
    return dd_quick(p, e);
}
VLR_DEV DD dd_div(DD a, DD b) {
    const double q1 = a.hi / b.hi;
    DD r = dd_add(a, dd_mul(DD{-q1, 0.0}, b));
    const double q2 = r.hi / b.hi;
    r = dd_add(r, dd_mul(DD{-q2, 0.0}, b));
    const double q3 = r.hi / b.hi;
    return dd_add(dd_quick(q1, q2), DD{q3, 0.0});
}
VLR_DEV_NOINLINE double log2_cr(double x) {
    if (x != x || x < 0.0) return NAN;
    if (x == 0.0) return -INFINITY;
    if (x == INFINITY) return x;
    int k;
    double m = frexp(x, &k); // [0.5, 1)
    if (m < 0.70710678118654752) {
        m *= 2.0;
        k -= 1;
    }
    // ln m = 2 atanh(s), s = (m - 1) / (m + 1), |s| <= 0.1716
    const DD s = dd_div(DD{m - 1.0, 0.0}, dd_two_sum(m, 1.0));
    const DD s2 = dd_mul(s, s);
    const double C[24][2] = {
    {1.0, 0.0},
    {0.3333333333333333, 1.850371707708594e-17},
    {0.2, -1.1102230246251566e-17},
    {0.14285714285714285, 7.93016446160826e-18},
    {0.1111111111111111, 6.1679056923619804e-18},
    {0.09090909090909091, -2.523234146875356e-18},
    {0.07692307692307693, -4.270088556250602e-18},
    {0.06666666666666667, 9.251858538542971e-19},
    {0.058823529411764705, 8.163404592832033e-19},
    {0.05263157894736842, 2.921639538487254e-18},
    {0.047619047619047616, 2.64338815386942e-18},
    {0.043478260869565216, 1.206764157201257e-18},
    {0.04, -8.326672684688674e-19},
    {0.037037037037037035, 2.05596856412066e-18},
    {0.034482758620689655, 4.785444071660157e-19},
    {0.03225806451612903, 8.953411488912552e-19},
    {0.030303030303030304, -8.410780489584519e-19},
    {0.02857142857142857, 8.921435019309293e-19},
    {0.02702702702702703, -1.50030138462859e-18},
    {0.02564102564102564, 8.896017825522087e-19},
    {0.024390243902439025, -8.46206573647223e-19},
    {0.023255813953488372, 3.2273925134452225e-19},
    {0.022222222222222223, -8.480870326997723e-19},
    {0.02127659574468085, 5.167261417803255e-19},
    };
    DD p = DD{C[23][0], C[23][1]};
    for (int i = 22; i >= 0; --i) p = dd_add(dd_mul(p, s2), DD{C[i][0], C[i][1]});
    DD ln = dd_mul(p, s);
    ln = DD{2.0 * ln.hi, 2.0 * ln.lo};
    const DD l2 = dd_mul(ln, DD{1.4426950408889634, 2.0355273740931033e-17});
    const DD r = dd_add(DD{(double)k, 0.0}, l2);
    return r.hi + r.lo;
}

#ifdef VLR_HOST_EMU
VLR_DEV double m_log2_lfc(double x) { return log2(x); } // the host build shares glibc with the oracle
#else
VLR_DEV double m_log2_lfc(double x) { return log2_cr(x); }
#endif

// ------------------------------------------------------------------------------------------------ LogProb helpers
// (rust-bio LogProb semantics)
VLR_DEV_NOINLINE double ln_add_exp(double a, double b) {
    double p0, p1;
    if (b > a) {
        p0 = b;
        p1 = a;
    } else {
        p0 = a;
        p1 = b;
    }
    if (p0 == neg_inf()) return neg_inf();
    if (p1 == neg_inf()) return p0;
    return p0 + m_log1p(m_exp(p1 - p0));
}
VLR_DEV_NOINLINE double ln_one_minus_exp(double p) {
    if (p < -0.693) return m_log1p(-m_exp(p));
    return m_log(-m_expm1(p));
}
// streaming ln_sum_exp accumulator (differs from the reference's max-first two-pass form by rounding only)
struct Lse {
    double m, s; // max so far, sum of m_exp(x - m)
    int n;
    VLR_DEV void init() {
        m = neg_inf();
        s = 0.0;
        n = 0;
    }
    VLR_DEV_NOINLINE void add(double x) {
        n++;
        if (x == neg_inf()) return;
        if (x != x) { // NaN poisons like the reference's arithmetic would
            m = x;
            return;
        }
        if (x > m) {
            s = (m == neg_inf()) ? 1.0 : s * m_exp(m - x) + 1.0;
            m = x;
        } else {
            s += m_exp(x - m);
        }
    }
    VLR_DEV double value() const {
        if (m == neg_inf() || m != m || m == INFINITY) return m;
        return m + m_log1p(s - 1.0);
    }
};

VLR_DEV int kass_raftery(double m1, double m2) {
    // BayesFactor::new(m1, m2) = m_exp(m1 - m2) compared with 1, 3, 20, 150; evaluated in log space.
    double d = m1 - m2;
    if (d <= 0.0) return 0;
    if (d <= LN_3) return 1;
    if (d <= LN_20) return 2;
    if (d <= LN_150) return 3;
    return 4; // incl. NaN, like the chain of failed comparisons in the reference
}

VLR_DEV bool relative_eq(double a, double b) { // approx 0.5 defaults (epsilon = max_relative = f64::EPSILON)
    if (a == b) return true;
    if (isinf(a) || isinf(b)) return false;
    double diff = fabs(a - b);
    const double eps = 2.220446049250313e-16;
    if (diff <= eps) return true;
    double largest = fmax(fabs(a), fabs(b));
    return diff <= largest * eps;
}

// ------------------------------------------------------------------------------------------------ VAFRange
struct Range {
    double start, end;
    bool lex, rex;
};
VLR_DEV Range range_empty() { return Range{0.0, 0.0, true, true}; }
VLR_DEV bool range_is_empty(const Range& r) { return r.start == r.end && (r.lex || r.rex); }
VLR_DEV bool range_is_singleton(const Range& r) { return r.start == r.end && !(r.lex || r.rex); }
VLR_DEV bool range_contains(const Range& r, double v) {
    bool l = r.lex ? (r.start < v) : (r.start <= v);
    bool rr = r.rex ? (r.end > v) : (r.end >= v);
    return l && rr;
}
VLR_DEV bool range_no_overlap(const Range& a, const Range& o) { // formula.rs:1137-1170
    if (a.start == o.start && a.end == o.end && a.lex == o.lex && a.rex == o.rex) return false;
    return (a.end < o.start || a.start > o.end) || (a.end <= o.start && (a.rex || o.lex)) ||
           (a.start >= o.end && (a.lex || o.rex));
}
VLR_DEV_NOINLINE Range range_intersect(const Range& a, const Range& o) {
    if (range_no_overlap(a, o)) return range_empty();
    Range r;
    r.start = fmax(a.start, o.start);
    r.end = fmin(a.end, o.end);
    r.lex = a.start > o.start ? a.lex : (a.start < o.start ? o.lex : (a.lex || o.lex));
    r.rex = a.end < o.end ? a.rex : (a.end > o.end ? o.rex : (a.rex || o.rex));
    return r;
}
VLR_DEV_NOINLINE double range_observable_max(const Range& r, int n) { // formula.rs:1202-1224
    if (n < 10 || !((double)n * (r.end - r.start) > 1.0)) return r.end;
    double c = (double)n * r.end;
    if (r.rex && c == floor(c)) c -= 1.0; // (c % 1.0 == 0.0 of the reference: c is a whole number; fmod is a loop on the device)
    c = floor(c);
    if (c == 0.0) return r.end;
    return c / (double)n;
}
VLR_DEV_NOINLINE double range_observable_min(const Range& r, int n) { // formula.rs:1172-1200
    double min_vaf;
    if (n < 10 || !((double)n * (r.end - r.start) > 1.0)) {
        min_vaf = r.start;
    } else {
        double c = (double)n * r.start;
        if (r.lex && c == floor(c)) {
            double adjusted_end = range_observable_max(r, n);
            double s1 = ceil(c + 1.0) / (double)n;
            if (s1 <= 1.0 && s1 <= adjusted_end) return s1;
            double s0 = ceil(c) / (double)n;
            if (s0 <= 1.0 && s0 <= adjusted_end) return s0;
        }
        min_vaf = ceil(c) / (double)n;
    }
    if (min_vaf >= range_observable_max(r, n)) return r.start;
    return min_vaf;
}

// log2 fold change predicates (utils/log2_fold_change.rs)
VLR_DEV int lfc_invert_cmp(int cmp) {
    switch (cmp) {
    case VLR_CMP_GT: return VLR_CMP_LE;
    case VLR_CMP_GE: return VLR_CMP_LT;
    case VLR_CMP_LT: return VLR_CMP_GE;
    case VLR_CMP_LE: return VLR_CMP_GT;
    default: return cmp;
    }
}
VLR_DEV_NOINLINE Range lfc_infer_bounds(int cmp, double value, double vaf) {
    double proj = vaf / m_exp2(value);
    if (proj < 0.0 || proj > 1.0) return range_empty();
    switch (cmp) {
    case VLR_CMP_EQ: return Range{proj, proj, false, false};
    case VLR_CMP_GT: return Range{0.0, proj, false, true};
    case VLR_CMP_GE: return Range{0.0, proj, false, false};
    case VLR_CMP_LT: return Range{proj, 1.0, true, false};
    case VLR_CMP_LE: return Range{proj, 1.0, false, false};
    default: return Range{0.0, 1.0, false, false};
    }
}


} // namespace vlrcore
