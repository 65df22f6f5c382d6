// engine_core.cuh — the per-locus posterior engine, written warp-cooperatively.
//
// One warp evaluates one locus. Control flow (VAF-tree walk, adaptive integration decisions, prior, event
// bookkeeping) is warp-uniform: every lane holds bit-identical copies of the deciding values. Lanes split the two
// things that are wide: the reads of a pileup (pre-pass statistics, per-read coefficients, likelihood product)
// and the points of an integration grid (sort + trapezoid).
//
// What it replaces in the reference (file:line under /root/reference/src):
//   calling/variants/calling.rs:586-626     pileup filtering, singleton adjustment      -> locus_prepass
//   variants/model/bias/*.rs                artifact models, learn_parameters, filters  -> locus_prepass, read_coefficients
//   variants/model/likelihood.rs:43-249     per-read emission + pileup fold             -> read_coefficients, sample_likelihood
//   variants/model/modes/generic.rs:191-554 VAF-tree density, joint likelihood          -> density, joint
//   utils/adaptive_integration.rs:25-141    unimodal adaptive integration               -> integrate_adaptive
//   variants/model/prior.rs:298-761         scenario prior                              -> prior_*
//   calling/variants/calling.rs:720-937     event posteriors, artifact, MAP, AFD        -> process_locus
//
// Numerical restructuring: in linear space the emission of one read is affine in the effective
// alt-sampling probability x,  L_r = e^{K_r} (alpha_r x + beta_r (1-x) + gamma_r),  so all VAF-independent work
// (bias terms, exp/log of the read's probabilities) is hoisted into three fp64 coefficients per (read, artifact
// config) and one pileup evaluation is 2 FMAs + 1 MUL per read and ONE log per pileup.
//
// The file compiles for sm_100a (nvcc) and, with -DVLR_HOST_EMU, as a single-lane host build that exists only so
// the control flow can be unit-tested without a GPU (tests/emu); the product library never contains that build.
//
// This header is a TEMPLATE BY INCLUSION: define VLR_VARIANT (namespace), VLR_VAR_MAXS, VLR_VAR_MAXE, VLR_VAR_MAXD
// (capacity in samples, events, tree depth) and include it; vlr_engine.cu does so twice (a small variant whose
// per-warp state fits a few hundred bytes of shared memory, and the full-capacity one).
#include "engine_types.cuh"

namespace VLR_VARIANT {
using namespace vlrcore;
constexpr int MAXS = VLR_VAR_MAXS;
constexpr int MAXE = VLR_VAR_MAXE;
constexpr int MAXD = VLR_VAR_MAXD;
constexpr int PRIOR_CACHE = 32;

// ------------------------------------------------------------------------------------------------ per-locus state
struct Ops { // generic.rs LikelihoodOperands
    double vaf[MAXS];
    uint32_t set_mask;  // samples whose event is present
    uint32_t disc_mask; // is_discrete per sample
    uint32_t lfc_mask;  // pushed LFC constraints by ordinal
};

struct Art { // bias::Artifacts restricted to "none" or exactly one artifact (bias/mod.rs:131-218)
    int id;  // 0 none; 1 ALB 2 HE 3 SCB 4 RPB 5 ROB_F1R2 6 ROB_F2R1 7 SB_FWD 8 SB_REV
    double forward_rate;
    double ln_fwd, ln_rev; // ln(forward_rate), ln(1 - forward_rate): computed once per locus (BiasPlan)
    bool has_alt_loci;
};

// utils/adaptive_integration.rs:25-141 as a small state machine that hands out the abscissae in batches:
// [min, max]; per iteration [middle, m1, m2]; finally the abandoned-arm midpoint and the 3 + 3 points around the optimum.
// The visit order equals the reference's. Argmax ties: ascending x, first maximum wins (the reference iterates a HashMap).
struct Adaptive {
    double minp, maxp, res;
    double left, right, f_left, f_right, middle, first_middle;
    int phase;
    bool have_middle, stop;
    VLR_DEV void init(double a, double b, double r) {
        minp = a;
        maxp = b;
        res = r;
        left = a;
        right = b;
        f_left = f_right = middle = first_middle = 0.0;
        phase = 0;
        have_middle = false;
        stop = false;
    }
    // fills xs[0..k) and returns k (2, 3 or 7)
    VLR_DEV int points(double* xs) const {
        if (phase == 0) {
            xs[0] = minp;
            xs[1] = maxp;
            return 2;
        }
        if (phase == 1) {
            xs[0] = middle;
            xs[1] = (middle + left) / 2.0;
            xs[2] = (right + middle) / 2.0;
            return 3;
        }
        xs[0] = (middle < first_middle) ? (maxp + first_middle) / 2.0 : (first_middle + minp) / 2.0;
        const double lo = fmax(middle - (res * 3.0), minp);
        const double slo = (middle - lo) / 3.0; // itertools-num linspace(lo, middle, 4).take(3)
        xs[1] = lo + slo * 0.0;
        xs[2] = lo + slo * 1.0;
        xs[3] = lo + slo * 2.0;
        const double hi = fmin(middle + (res * 3.0), maxp);
        const double shi = (hi - middle) / 3.0; // linspace(middle, hi, 4).skip(1)
        xs[4] = middle + shi * 1.0;
        xs[5] = middle + shi * 2.0;
        xs[6] = middle + shi * 3.0;
        return 7;
    }
    // takes the values of the batch; returns false when the integration grid is complete
    VLR_DEV bool consume(const double* xs, const double* fs, bool overflow) {
        if (phase == 2) return false;
        if (phase == 0) {
            f_left = fs[0];
            f_right = fs[1];
        } else {
            if (!have_middle) first_middle = middle;
            have_middle = true;
            const double m1 = xs[1], m2 = xs[2], f_m1 = fs[1], f_m2 = fs[2];
            int idx = 0;
            double fb = f_left;
            if (f_m1 > fb) {
                idx = 1;
                fb = f_m1;
            }
            if (f_m2 > fb) {
                idx = 2;
                fb = f_m2;
            }
            if (f_right > fb) idx = 3;
            // neighbours of the argmax in [left, m1, m2, right] become the new bounds (the middle is not a candidate)
            const double nl = idx <= 1 ? left : (idx == 2 ? m1 : m2), nfl = idx <= 1 ? f_left : (idx == 2 ? f_m1 : f_m2);
            const double nr = idx == 0 ? m1 : (idx == 1 ? m2 : right), nfr = idx == 0 ? f_m1 : (idx == 1 ? f_m2 : f_right);
            left = nl;
            f_left = nfl;
            right = nr;
            f_right = nfr;
        }
        // while (((right - left) >= res && left < right) || middle.is_none())
        if (!overflow && ((((right - left) >= res) && left < right) || !have_middle)) {
            phase = 1;
            middle = (right + left) / 2.0;
        } else {
            phase = 2;
        }
        return true;
    }
};


// ---- leaf fast path state -------------------------------------------------------------------------------------
// Along a leaf-level integration only sample t's VAF changes. Everything else is hoisted: the prior factors and the
// pileup likelihoods of the samples that do not see t (GenericLikelihood::compute, generic.rs:511-551, hits its
// per-sample cache for them at every point), and for the (at most two) samples that do, their pileup constants.
struct LeafDep {
    const double2* co; // coefficient arena of the sample (shared or global memory)
    double ksum, rho, iota, fixed_vaf, fixed_by;
    int n, s;
    bool vaf_is_x, by_is_x, has_by, vaf_is_parent, by_is_parent;
};
struct LeafFast {
    LeafDep dep[3]; // [0..1]: the pileups that see the leaf sample; [2]: the parent sample's pileup (setup only)
    int dep_lo, dep_hi; // the descriptors multi_eval walks
    int n_dep, n_tasks, parent, t;
    bool prior_per_point; // the prior has to be evaluated per point (non-uniform priors, or a universe with holes)
    bool uniform, t_ploidy0, coef_in_sm, record;
    bool parent_disc; // the parent sample's event is discrete (Set node) rather than an integration abscissa
};
// One of the (up to MT) leaf integrations a warp advances concurrently: they differ in the VAF of the parent sample
// (the abscissae of the enclosing integration's current batch).
struct MultiTask {
    Adaptive st;
    double parent_x, lh_const, prior_const, best_f, best_x;
    double xs[8], fs[8];
    int n, k, slot_base;
    int event_slot;   // MAP slot the task's base events belong to (2 * event + artifact config ? 1 : 0)
    bool have_best, active, overflow;
};

// One instance per warp, in SHARED memory: every lane sees the same (uniform) state, so a single copy replaces
// 32 per-lane local-memory copies. Uniform state is updated by ALL lanes storing the identical value (each lane's own
// store precedes its own load, so no lane can observe a stale value), lane-owned state (Ctx::mt[k] by lane k, slot
// arrays) is published with warp_sync() before other lanes read it, and the warp collectives at the end of every
// reads loop keep the lanes converged between those points. compute-sanitizer racecheck reports these same-value
// stores as "warp level programming" hazards (profiles/README.md); memcheck and synccheck are clean.
struct Ctx {
    const DevScenario* sc;
    const DevBatch* b;
    const DevResults* res;
    WarpWs* ws;
    double* coef; // per-warp coefficient arena: 4 doubles per kept read (alpha, beta, gamma, s)
    double* be;   // per-warp base-event log (NULL unless an AFD is requested): BE_CAP x (2 + S) doubles
    uint32_t n_rec;
    int coef_in_sm; // the coefficients of this locus live in the warp's shared-memory arena
    int coef_cap;   // capacity of `coef` in reads
    int coef_total; // kept reads of this locus over all samples
    int64_t locus;
    uint32_t lf;
    uint32_t status;
    int vartype;
    double het_override, semr_override; // ln, NaN = none
    bool has_snv;
    int refbase, altbase;
    int64_t singleton_row; // absolute row of the read whose evidence is adjusted, or -1
    Art art;
    // per sample
    int n_obs[MAXS];
    int n_notref[MAXS]; // kept reads that do not favour the reference (prob_ref <= prob_alt)
    int clear_ref[MAXS];
    int s_one[MAXS];   // every kept read has prob_sample_alt == 0
    int s_gt1[MAXS];   // some kept read has prob_sample_alt > 0
    int coef_off[MAXS];
    double ksum[MAXS];
    // events
    int cur_slot; // 2*event + (artifact config ? 1 : 0)
    double map_joint[2 * MAXE];
    double map_vaf[2 * MAXE][MAXS];
    uint32_t map_disc[2 * MAXE];
    int map_cfg[2 * MAXE];
    uint8_t map_set[2 * MAXE];
    uint32_t map_seq[2 * MAXE]; // evaluation number of the slot's base event (ties: the earlier one wins, like the oracle's stable sort)
    // MAP candidates are offered to every event's slots (instead of the walked event's own): set for scenarios whose
    // events overlap, and per locus as soon as an integration's limits lie on an excluded range bound (calling.rs:861-864)
    int map_global;
    // second pass of a locus whose base-event log overflowed: only base events that differ from the MAP in at most
    // one sample are logged (all an allele frequency distribution can use, calling.rs:891-928)
    int be_filter;
    double flt_vaf[MAXS];
    uint32_t flt_disc;
    // per-event accumulators and end-of-locus scratch
    Lse ev_plain[MAXE], ev_twin[MAXE];
    double joint_u[2 * MAXE];
    double my_lp[MAXE + 1];
    int scen_u[2 * MAXE];
    uint8_t art_u[2 * MAXE];
    double prior_absent; // Prior of the all-zero event (absent-only mode), NaN = not computed yet
    int pc_n;
    uint32_t n_base;
    uint32_t n_pileup_evals;
    // ---- everything above is what the pre-pass, the coefficients, a pileup evaluation, the prior and locus_tail use:
    // the pipelines' kernels (engine_wave.cuh, engine_sets.cuh) may give a warp only CTX_LEAN bytes of shared memory.
    // Below: state of the generic engine's tree walk only.
    // pileup likelihood cache (generic.rs:43-53; here LC_WAYS most recent entries per sample and config)
    double lc_k1[MAXS][LC_WAYS], lc_k2[MAXS][LC_WAYS], lc_v[MAXS][LC_WAYS];
    int lc_n[MAXS];
    // operand stack of the tree walk (generic.rs clones LikelihoodOperands per branch / grid point)
    Ops ops[MAXD + 2];
    // adaptive integration state per nesting level (at most one Range level per sample) and the leaf fast path
    Adaptive ad[MAXS];
    double xs[MAXS][8], fs[MAXS][8];
    LeafFast leaf;
    MultiTask mt[MT];
    double slot_x[MSLOTS], slot_f[MSLOTS];
    int slot_task[MSLOTS];
    int slot_slow[MSLOTS];
    // prior cache for all-discrete VAF vectors (prior.rs:718-736 keeps an LRU(1000) keyed by the VAF vector; the
    // prior does not depend on the artifact config, so the 27 combinations of a trio are computed once per locus)
    double pc_key[PRIOR_CACHE][MAXS], pc_val[PRIOR_CACHE];
};

// Shared memory of a CTA: [WARPS_PER_CTA x Ctx][WARPS_PER_CTA x SM_READS x 4 doubles]. Deriving the per-warp
// references from the __shared__ symbol (instead of carrying generic pointers through calls) lets the compiler
// emit LDS/STS for all of the uniform state.
constexpr int CTX_STRIDE = (int)((sizeof(Ctx) + 15) & ~(size_t)15);
constexpr int CTX_LEAN = (int)((offsetof(Ctx, lc_k1) + 15) & ~(size_t)15); // see the comment inside Ctx
#ifdef VLR_HOST_EMU
VLR_DEV Ctx& warp_ctx(Ctx& c) { return c; }
#else
// The same object, addressed through the __shared__ symbol (LDS/STS instead of generic accesses) wherever the kernel
// placed it inside its dynamic shared memory.
VLR_DEV Ctx& warp_ctx(Ctx& c) {
    return *reinterpret_cast<Ctx*>(vlr_smem + (__cvta_generic_to_shared(&c) - __cvta_generic_to_shared(vlr_smem)));
}
VLR_DEV double* warp_coef_sm() {
    return reinterpret_cast<double*>(vlr_smem + WARPS_PER_CTA * CTX_STRIDE) + group_in_cta() * (SM_READS * 4);
}
#endif

// ------------------------------------------------------------------------------------------------ reads
struct Read {
    double pm, pa, pr, pmiss, psa, pdo, phb, hart, hvar;
    uint32_t f;
    bool has_hart, has_hvar;
};
VLR_DEV Read load_read(const DevBatch* b, int64_t row) {
    int64_t i = row - b->read_base;
    Read r;
    r.pm = (double)ldin(b->pm + i);
    r.pa = (double)ldin(b->pa + i);
    r.pr = (double)ldin(b->pr + i);
    r.pmiss = (double)ldin(b->pmiss + i);
    r.psa = (double)ldin(b->psa + i);
    r.pdo = (double)ldin(b->pdo + i);
    r.phb = (double)ldin(b->phb + i);
    r.f = ldin(b->rflags + i);
    float ha = b->hart ? ldin(b->hart + i) : NAN, hv = b->hvar ? ldin(b->hvar + i) : NAN;
    r.has_hart = !(ha != ha);
    r.has_hvar = !(hv != hv);
    r.hart = r.has_hart ? (double)ha : 0.0;
    r.hvar = r.has_hvar ? (double)hv : 0.0;
    return r;
}
VLR_DEV int rd_strand(uint32_t f) { return (f >> VLR_RF_STRAND_SHIFT) & 3; }
VLR_DEV int rd_orient(uint32_t f) { return (f >> VLR_RF_ORIENT_SHIFT) & 15; }
VLR_DEV int rd_altlocus(uint32_t f) { return (f >> VLR_RF_ALTLOCUS_SHIFT) & 3; }
VLR_DEV int rd_hlen(uint32_t f) {
    return (f & VLR_RF_HAS_HOMOPOLYMER_LEN) ? (int)(int8_t)((f >> VLR_RF_HOMOPOLYMER_LEN_SHIFT) & 0xff) : 0;
}
// pileup.rs:26-43 (only applied when the locus asks for it)
VLR_DEV bool rd_kept(uint32_t lf, uint32_t f) {
    if (!(lf & VLR_LF_FILTER_NONSTANDARD)) return true;
    int o = rd_orient(f);
    return o == 0 || o == 1 || o == 8;
}

// Outcome of the bias pre-pass
struct BiasPlan {
    int n_twins;           // unfiltered number of single-artifact configs (generic.rs:437-441)
    int surviving[NCFG];   // config ids that pass is_possible && is_informative && is_likely
    int n_surviving;
    double forward_rate;
    double ln_fwd, ln_rev; // ln(forward_rate), ln(1 - forward_rate)
    bool has_alt_loci;
};

// One pass over all reads of the locus: everything the reference derives from the pileups before the model runs.
//   calling.rs:595-616 (filter + singleton), generic.rs:270-291 (n_obs, is_clear_ref),
//   strand_bias.rs:79-123, read_orientation_bias.rs:38-97, read_position_bias.rs:63-121, softclip_bias.rs:32-39,
//   homopolymer_error.rs:46-72, alt_locus_bias.rs:47-59,124-144, bias/mod.rs:37-104.
VLR_DEV_NOINLINE void locus_prepass(Ctx& c_, BiasPlan& plan) {
    Ctx& c = warp_ctx(c_);
    const DevBatch* b = c.b;
    const int S = c.sc->S;
    const uint32_t lf = c.lf;
    // global accumulators
    double g_sb_all = 0.0, g_sb_fwd = 0.0;
    int g_n = 0, g_uncertain = 0, g_sr_std = 0, g_sr_f1r2 = 0, g_n_alt = 0, g_nm_alt = 0, g_n_ref = 0, g_nm_ref = 0;
    int g_alt_support = 0;
    long long g_alt_row = -1;
    unsigned g_flags = 0; // bit0 has_alt_loci, 1 softclip, 2 pos_sb_fwd, 3 pos_sb_rev, 4 pos_rob1, 5 pos_rob2, 6 pos_rpb,
                          // 7 pos_alb_loci, 8 pos_alb_mapq
    bool he_informative = true;
    bool rpb_valid = false;
    // per sample "likely" inputs
    int uq_salt[MAXS];
    int ev_cnt[MAXS][10]; // index by config id 1..8; 9 = ALB evidence when no alt loci exist
    int all_ref[MAXS];
    int coef_total = 0;
    for (int s = 0; s < S; ++s) {
        int64_t lo = b->read_offsets[c.locus * S + s], hi = b->read_offsets[c.locus * S + s + 1];
        int n = 0, n_notposref = 0, n_strong_alt = 0, n_ins = 0, n_del = 0, n_notref = 0, n_uq = 0;
        int e1 = 0, e2 = 0, e3 = 0, e4 = 0, e5 = 0, e6 = 0, e7 = 0, e8 = 0, e9 = 0;
        int n_snot0 = 0, n_sgt0 = 0;
        double r_all = 0.0, r_major = 0.0, r_rate = 0.0;
        double memo_pm = NAN, memo_w = 0.0, memo_pmh = NAN, memo_wh = 0.0;
        for (int64_t row = lo + lane_id(); row < hi; row += LANES) {
            int64_t i = row - b->read_base;
            uint32_t f = ldin(b->rflags + i);
            if (!rd_kept(lf, f)) continue;
            double pm = (double)ldin(b->pm + i), pa = (double)ldin(b->pa + i), pr = (double)ldin(b->pr + i);
            double phb = (double)ldin(b->phb + i);
            float psa = ldin(b->psa + i);
            int kr_ref = kass_raftery(pr, pa), kr_alt = kass_raftery(pa, pr);
            bool strong_ref = kr_ref >= 3, strong_alt = kr_alt >= 3;
            int strand = rd_strand(f), orient = rd_orient(f), al = rd_altlocus(f);
            bool major = (f & VLR_RF_READPOS_MAJOR) != 0, softclip = (f & VLR_RF_SOFTCLIPPED) != 0;
            bool maxq = (f & VLR_RF_MAX_MAPQ) != 0;
            int hl = rd_hlen(f);
            n++;
            n_snot0 += (psa != 0.0f);
            n_sgt0 += (psa > 0.0f);
            n_notposref += !(kr_ref >= 2);
            n_strong_alt += strong_alt;
            n_ins += hl > 0;
            n_del += hl < 0;
            n_notref += !(pr > pa);
            if (pa > pr) {
                g_alt_support++;
                g_alt_row = row;
            }
            if (strong_ref) {
                if (pm != memo_pm) { // exp of a repeated argument (MAPQ takes a handful of values): same bits
                    memo_pm = pm;
                    memo_w = m_exp(pm);
                }
                const double w = memo_w;
                if (strand != 2) g_sb_all += w;
                if (strand == 0) g_sb_fwd += w;
                bool std_or = orient == 0 || orient == 1;
                g_sr_std += std_or;
                g_sr_f1r2 += orient == 0;
                g_n_ref++;
                g_nm_ref += !maxq;
                r_all += w;
                if (major) r_major += w;
                const double pmh = pm + phb;
                if (pmh != memo_pmh) {
                    memo_pmh = pmh;
                    memo_wh = m_exp(pmh);
                }
                r_rate += memo_wh;
            }
            g_uncertain += !(orient == 0 || orient == 1);
            if (strong_alt) {
                g_n_alt++;
                g_nm_alt += !maxq;
            }
            unsigned fl = 0;
            fl |= (al != 2) ? 1u : 0u;
            fl |= softclip ? 2u : 0u;
            bool ev_sbf = strand == 0 || strand == 3, ev_sbr = strand == 1 || strand == 3;
            bool ev_rob1 = orient != 1, ev_rob2 = orient != 0;
            fl |= ev_sbf ? 4u : 0u;
            fl |= ev_sbr ? 8u : 0u;
            fl |= ev_rob1 ? 16u : 0u;
            fl |= ev_rob2 ? 32u : 0u;
            fl |= major ? 64u : 0u;
            fl |= (al == 0) ? 128u : 0u;
            fl |= (!maxq) ? 256u : 0u;
            g_flags |= fl;
            if (pm >= LN_095 && strong_alt) { // is_uniquely_mapping && is_strong_alt_support (bias/mod.rs:66-72)
                n_uq++;
                e1 += al == 0;
                e9 += !maxq;
                e2 += hl != 0;
                e3 += softclip;
                e4 += major;
                e5 += ev_rob1;
                e6 += ev_rob2;
                e7 += ev_sbf;
                e8 += ev_sbr;
            }
        }
        n = w_sum_i(n);
        c.n_obs[s] = n;
        c.clear_ref[s] = (n > 10) && (w_sum_i(n_notposref) == 0);
        c.s_one[s] = w_sum_i(n_snot0) == 0;
        c.s_gt1[s] = w_sum_i(n_sgt0) != 0;
        c.coef_off[s] = coef_total;
        coef_total += n;
        g_n += n;
        bool any_strong_alt = w_sum_i(n_strong_alt) > 0;
        bool has_ins = w_sum_i(n_ins) > 0, has_del = w_sum_i(n_del) > 0;
        if (!(!any_strong_alt || (has_ins && has_del))) he_informative = false;
        c.n_notref[s] = w_sum_i(n_notref);
        all_ref[s] = c.n_notref[s] == 0;
        uq_salt[s] = w_sum_i(n_uq);
        ev_cnt[s][0] = 0;
        ev_cnt[s][1] = w_sum_i(e1);
        ev_cnt[s][2] = w_sum_i(e2);
        ev_cnt[s][3] = w_sum_i(e3);
        ev_cnt[s][4] = w_sum_i(e4);
        ev_cnt[s][5] = w_sum_i(e5);
        ev_cnt[s][6] = w_sum_i(e6);
        ev_cnt[s][7] = w_sum_i(e7);
        ev_cnt[s][8] = w_sum_i(e8);
        ev_cnt[s][9] = w_sum_i(e9);
        // read position bias: any pileup with a valid major rate (read_position_bias.rs:63-121)
        double ra = w_sum_d(r_all);
        if (ra > 10.0) {
            double rm = w_sum_d(r_major), rr = w_sum_d(r_rate);
            double major_rate = rm / ra;
            if (rm > 0.0 && fabs(major_rate - rr) < 0.05) rpb_valid = true;
        }
    }
    c.coef_total = coef_total;
    g_sb_all = w_sum_d(g_sb_all);
    g_sb_fwd = w_sum_d(g_sb_fwd);
    g_uncertain = w_sum_i(g_uncertain);
    g_sr_std = w_sum_i(g_sr_std);
    g_sr_f1r2 = w_sum_i(g_sr_f1r2);
    g_n_alt = w_sum_i(g_n_alt);
    g_nm_alt = w_sum_i(g_nm_alt);
    g_n_ref = w_sum_i(g_n_ref);
    g_nm_ref = w_sum_i(g_nm_ref);
    g_flags = w_or_u(g_flags);
    g_alt_support = w_sum_i(g_alt_support);
    // adjust_singleton_evidence (read_observation.rs:548-562)
    c.singleton_row = -1;
    if (g_alt_support == 1) {
        int hi32 = w_max_i((int)(g_alt_row >> 31)); // rows are < 2^62; split to use the integer reductions
        int lo31 = w_max_i((g_alt_row >> 31) == hi32 ? (int)(g_alt_row & 0x7fffffff) : -1);
        c.singleton_row = ((int64_t)hi32 << 31) | (int64_t)lo31;
        c.status |= VLR_ST_SINGLETON_ADJUSTED;
    }
    {
        // Hint::FilteredNonStandardAlignments (calling.rs:600-602, 620-624)
        int total = 0;
        for (int s = 0; s < S; ++s)
            total += (int)(b->read_offsets[c.locus * S + s + 1] - b->read_offsets[c.locus * S + s]);
        if (total != g_n) c.status |= VLR_ST_FILTERED_NONSTANDARD;
    }

    // strand bias forward rate (strand_bias.rs:79-123)
    bool sb_informative = false;
    plan.forward_rate = 0.5;
    if (g_sb_all > 2.0) {
        double ff = g_sb_fwd / g_sb_all;
        if (g_sb_all > 100.0 && ff > 0.0 && ff < 1.0) {
            plan.forward_rate = ff;
            sb_informative = true;
        } else if (ff >= 0.4 && ff <= 0.6) {
            plan.forward_rate = 0.5;
            sb_informative = true;
        }
    }
    plan.ln_fwd = m_log(plan.forward_rate);
    plan.ln_rev = m_log(1.0 - plan.forward_rate);
    plan.has_alt_loci = (g_flags & 1u) != 0;
    // read orientation bias (read_orientation_bias.rs:38-97)
    bool rob_informative = false;
    {
        bool enough = (double)g_uncertain < ((double)g_n / 2.0);
        bool uniform = false;
        if (g_sr_std > 2) {
            double fraction = (double)g_sr_f1r2 / (double)g_sr_std;
            uniform = fraction >= 0.3 && fraction <= 0.7;
        }
        rob_informative = enough && uniform;
    }
    // alt locus bias (alt_locus_bias.rs:124-144)
    bool alb_informative;
    {
        bool enough_alt = g_n_alt > 0 && (double)g_nm_alt > ((double)g_n_alt * 0.1) && (g_n_alt - g_nm_alt) < 10;
        bool enough_ref = g_n_ref > 0 && ((double)g_nm_ref < ((double)g_n_ref * 0.9));
        alb_informative = enough_alt && (plan.has_alt_loci || enough_ref);
    }

    plan.n_twins = 0;
    plan.n_surviving = 0;
    for (int id = 1; id <= 8; ++id) {
        uint32_t need = id == 1 ? VLR_LF_CHECK_ALB
                        : id == 2 ? VLR_LF_CHECK_HE
                        : id == 3 ? VLR_LF_CHECK_SCB
                        : id == 4 ? VLR_LF_CHECK_RPB
                        : id <= 6 ? VLR_LF_CHECK_ROB
                                  : VLR_LF_CHECK_SB;
        if (!(lf & need)) continue;
        plan.n_twins++;
        bool possible, informative;
        switch (id) {
        case 1:
            possible = plan.has_alt_loci ? (g_flags & 128u) != 0 : (g_flags & 256u) != 0;
            informative = alb_informative;
            break;
        case 2: possible = informative = he_informative; break;
        case 3:
            possible = (g_flags & 2u) != 0;
            informative = (g_flags & 2u) != 0;
            break;
        case 4:
            possible = (g_flags & 64u) != 0;
            informative = rpb_valid;
            break;
        case 5:
            possible = (g_flags & 16u) != 0;
            informative = rob_informative;
            break;
        case 6:
            possible = (g_flags & 32u) != 0;
            informative = rob_informative;
            break;
        case 7:
            possible = (g_flags & 4u) != 0;
            informative = sb_informative;
            break;
        default:
            possible = (g_flags & 8u) != 0;
            informative = sb_informative;
            break;
        }
        if (!(possible && informative)) continue;
        bool likely = false;
        if (id == 2) {
            likely = he_informative; // homopolymer_error.rs:78-80
        } else {
            for (int s = 0; s < S; ++s) { // bias/mod.rs:62-104
                bool l;
                if (uq_salt[s] >= 10) {
                    int ev = (id == 1 && !plan.has_alt_loci) ? ev_cnt[s][9] : ev_cnt[s][id];
                    double ratio = (double)ev / (double)uq_salt[s];
                    l = ratio >= 0.66666;
                } else if (all_ref[s]) {
                    l = false; // also the empty pileup
                } else {
                    l = true;
                }
                if (l) likely = true;
            }
        }
        if (likely) plan.surviving[plan.n_surviving++] = id;
    }
}

// Hoist everything VAF-independent of one (sample, artifact config) into per-read coefficients.
//   L_r(x) = e^{K_r} * (alpha_r * x + beta_r * (1 - x) + gamma_r); the arena keeps [alpha, beta, gamma, 1 - s_r]
//   ln alpha' = prob_mapping + bias.prob_alt + prob_alt     (likelihood.rs:213, :185/:107)
//   ln beta'  = prob_mapping + prob_ref + bias.prob_ref      (likelihood.rs:215)
//   ln gamma' = prob_mismapping + prob_missed_allele + bias.prob_any   (likelihood.rs:186-188)
//   K_r = max of the three; x = effective alt-sampling probability (likelihood.rs:43-53, :98-103).
// Per-warp table of ln(1 - e^{prob_mapping}) keyed by the f32 bits of prob_mapping (read_observation.rs:283-286): MAPQs
// come from a dictionary of a few dozen values, but 32 lanes with one-entry memos miss somewhere in almost every
// iteration (5 % of the reads differ from their lane's previous one), and a warp pays for ln_one_minus_exp (an expm1 or
// an exp, and a log) whenever ONE lane needs it. Direct mapped, filled by the warp itself at converged points only
// (one elected lane per slot writes value and key: entries are never torn), read in between: same function of the same
// argument, so the coefficients are bitwise what they were. The kernel that owns the table clears it (memo_clear).
struct MemoTab {
    static constexpr int N = 128;
    uint32_t k[N];
    double v[N];
};
constexpr uint32_t MEMO_EMPTY = 0xffffffffu; // a NaN pattern: never a key (NaNs bypass the table)
VLR_DEV void memo_clear(MemoTab* t) { // warp cooperative
    for (int i = lane_id(); i < MemoTab::N; i += LANES) t->k[i] = MEMO_EMPTY;
    warp_sync();
}
VLR_DEV uint32_t memo_bits(float f) {
#ifdef VLR_HOST_EMU
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#else
    return __float_as_uint(f);
#endif
}

VLR_DEV_NOINLINE void read_coefficients(Ctx& c_, int s, MemoTab* tab = nullptr) {
    Ctx& c = warp_ctx(c_);
    const DevBatch* b = c.b;
    const int S = c.sc->S;
    const Art a = c.art;
    int64_t lo = b->read_offsets[c.locus * S + s], hi = b->read_offsets[c.locus * S + s + 1];
    double* out = c.coef + (int64_t)c.coef_off[s] * 4;
    double ksum = 0.0;
    bool dead = false, bad = false;
    int base = 0; // kept reads before this chunk of LANES rows
    // Transcendentals dominate this function (profiles/: exp/log/log1p were 2/3 of the prep kernel's instructions), and
    // most of their arguments repeat: the strand rate is a per-locus constant, MAPQ and prob_hit_base take a handful of
    // values. Same function of the same argument = same bits, so one-entry per-lane memos and the exact special values
    // (exp(0) = 1, exp(-inf) = 0, ln(1 - e^-inf) = -0) change nothing numerically.
    const double ln_rate_fwd = a.ln_fwd, ln_rate_rev = a.ln_rev;
    double memo_pm = NAN, memo_pmis = 0.0, memo_phb = NAN, memo_l1m_phb = 0.0, memo_pdo = NAN, memo_l1m_pdo = 0.0;
    for (int64_t row0 = lo; row0 < hi; row0 += LANES) {
        int64_t row = row0 + lane_id();
        bool valid = row < hi;
        Read r;
        bool kept = false;
        if (valid) {
            r = load_read(b, row);
            kept = rd_kept(c.lf, r.f);
        }
        bool memo_miss = false; // this lane computed a value the table does not hold yet
        uint32_t memo_h = 0, memo_key = 0;
#ifdef VLR_HOST_EMU
        int pos = base;
        int n_kept = kept ? 1 : 0;
#else
        unsigned m = w_ballot(kept);
        int pos = base + __popc(m & ((1u << lane_id()) - 1u));
        int n_kept = __popc(m);
#endif
        if (kept) {
            double pa = r.pa, pr = r.pr;
            if (row == c.singleton_row) pa = pr = LN_05;
            int strand = rd_strand(r.f), orient = rd_orient(r.f), al = rd_altlocus(r.f);
            bool major = (r.f & VLR_RF_READPOS_MAJOR) != 0, softclip = (r.f & VLR_RF_SOFTCLIPPED) != 0;
            bool maxq = (r.f & VLR_RF_MAX_MAPQ) != 0;
            // strand bias (strand_bias.rs:28-53)
            double sb_alt;
            if (strand == 3) sb_alt = 0.0;
            else if (a.id == 7) sb_alt = strand == 0 ? 0.0 : neg_inf();
            else if (a.id == 8) sb_alt = strand == 1 ? 0.0 : neg_inf();
            else if (strand == 2) sb_alt = r.pdo;
            else {
                if (r.pdo != memo_pdo) {
                    memo_pdo = r.pdo;
                    memo_l1m_pdo = r.pdo == neg_inf() ? -0.0 : ln_one_minus_exp(r.pdo);
                }
                sb_alt = (strand == 0 ? ln_rate_fwd : ln_rate_rev) + memo_l1m_pdo;
            }
            // read orientation bias (read_orientation_bias.rs:17-29)
            double rob_alt = LN_05;
            if (a.id == 5) rob_alt = orient == 0 ? 0.0 : (orient == 1 ? neg_inf() : LN_05);
            else if (a.id == 6) rob_alt = orient == 1 ? 0.0 : (orient == 0 ? neg_inf() : LN_05);
            // read position bias (read_position_bias.rs:17-61)
            double rpb_any = r.phb;
            if (!major) {
                if (r.phb != memo_phb) {
                    memo_phb = r.phb;
                    memo_l1m_phb = r.phb != 0.0 ? ln_one_minus_exp(r.phb) : 0.0;
                }
                rpb_any = memo_l1m_phb;
            }
            double rpb_alt = a.id == 4 ? (major ? 0.0 : neg_inf()) : rpb_any;
            // softclip bias (softclip_bias.rs:14-24)
            double scb_alt = a.id == 3 ? (softclip ? 0.0 : neg_inf()) : 0.0;
            // homopolymer error (homopolymer_error.rs:22-40): same term for alt and ref, none for any
            double he = a.id == 2 ? (r.has_hart ? r.hart : 0.0) : (r.has_hvar ? r.hvar : 0.0);
            // alt locus bias (alt_locus_bias.rs:62-112)
            double alb_alt = LN_05, alb_ref = LN_05;
            if (a.id == 1) {
                if (a.has_alt_loci) {
                    alb_alt = al == 0 ? 0.0 : neg_inf();
                    alb_ref = al == 0 ? neg_inf() : 0.0;
                } else {
                    alb_alt = maxq ? neg_inf() : 0.0;
                }
            }
            // bias/mod.rs:259-284 (summation order of the reference)
            double bA = sb_alt + rob_alt + rpb_alt + scb_alt + he + alb_alt;
            double bR = LN_05 + LN_05 + rpb_any + 0.0 + he + alb_ref;
            double bAny = LN_05 + LN_05 + rpb_any + 0.0 + 0.0 + LN_05;
            if (r.pm != memo_pm) {
                memo_pm = r.pm;
                const float pf = (float)r.pm;
                const bool tabbed = tab != nullptr && (double)pf == r.pm; // (the column is f32: always, unless NaN)
                memo_key = memo_bits(pf);
                memo_h = (memo_key * 0x9E3779B1u) >> 25;
                if (tabbed && tab->k[memo_h] == memo_key) {
                    memo_pmis = tab->v[memo_h];
                } else {
                    memo_pmis = ln_one_minus_exp(r.pm);
                    memo_miss = tabbed;
                }
            }
            const double pmis = memo_pmis; // read_observation.rs:283-286
            double lnA = r.pm + (bA + pa);
            double lnR = r.pm + (pr + bR);
            double lnC = pmis + r.pmiss + bAny;
            double K = fmax(lnA, fmax(lnR, lnC));
            double al_ = 0.0, be_ = 0.0, ga_ = 0.0;
            if (K != K || lnA != lnA || lnR != lnR || lnC != lnC || K == INFINITY) {
                bad = true;
            } else if (K == neg_inf()) {
                dead = true;
            } else {
                // two exps for the three terms: the largest is exp(0) = 1, e1 and e2 are the other two (exp(0) = 1 and
                // exp(-inf) = 0 exactly: the same bits as a guarded call per term, one call site less per read)
                const bool aK = lnA == K, cK = lnC == K;
                const double e1 = m_exp((aK ? lnR : lnA) - K), e2 = m_exp((cK ? lnR : lnC) - K);
                al_ = aK ? 1.0 : e1;
                be_ = aK ? e1 : (cK ? e2 : 1.0);
                ga_ = cK ? 1.0 : e2;
                ksum += K;
            }
            double* o = out + (int64_t)pos * 4;
            o[0] = al_;
            o[1] = be_;
            o[2] = ga_;
            o[3] = r.psa == 0.0 ? -0.0 : -m_expm1(r.psa); // u_r = 1 - s_r, s_r = e^{prob_sample_alt}: 0 when prob_sample_alt = 0
        }
        if (tab != nullptr) { // the warp is converged here: new values go into the table, one writer per slot
#ifdef VLR_HOST_EMU
            if (memo_miss) {
                tab->v[memo_h] = memo_pmis;
                tab->k[memo_h] = memo_key;
            }
#else
            if (w_any(memo_miss)) {
                const unsigned grp = __match_any_sync(FULL, memo_miss ? memo_h : MEMO_EMPTY);
                if (memo_miss && (int)(threadIdx.x & 31u) == __ffs(grp) - 1) {
                    tab->v[memo_h] = memo_pmis;
                    tab->k[memo_h] = memo_key;
                }
                warp_sync();
            }
#endif
        }
        base += n_kept;
    }
    ksum = w_sum_d(ksum);
    if (w_any(bad)) {
        c.status |= VLR_ST_NAN;
        ksum = NAN;
    } else if (w_any(dead)) {
        ksum = neg_inf();
    }
    c.ksum[s] = ksum;
    warp_sync();
}

// Product over the reads of one pileup of (alpha x + beta y + gamma), as mantissa in [1,2) per lane + binary exponent.
// MODE 0: every read has prob_sample_alt = 0 (x, y are per-evaluation scalars); 1: per-read s_r <= 1; 2: some s_r > 1
// (cap_numerical_overshoot per component, likelihood.rs:43-53). The arena holds [alpha, beta, gamma, s] per read.
struct PileupArgs {
    double xu, Yp, X1, X0, Y0, vaf, vaf_by, rho, iota;
    bool p1, s1;
};
VLR_DEV void pileup_product(const double2* __restrict__ co, int n, const int MODE, const PileupArgs& a, double& acc_out,
                            int& ex_out, bool& zero_out, bool& overshoot_out) {
    double acc = 1.0;
    int ex = 0, k = 0;
    bool zero = false, overshoot = false;
    for (int r = lane_id(); r < n; r += LANES) {
        const double2 ab = co[2 * r];
        const double2 gs = co[2 * r + 1];
        double t;
        if (MODE == 0) {
            t = fma(ab.x, a.xu, fma(ab.y, a.Yp, gs.x));
        } else if (MODE == 1) {
            const double sr = 1.0 - gs.y; // the arena keeps u_r = 1 - s_r
            t = fma(ab.x, fma(sr, a.X1, a.X0), fma(ab.y, fma(-sr, a.X1, a.Y0), gs.x));
        } else {
            const double sr = 1.0 - gs.y;
            double xp = a.p1 ? 1.0 : a.vaf * sr, xs = a.s1 ? 1.0 : a.vaf_by * sr;
            if (xp > 1.0) {
                if (m_log(xp) > NUMERICAL_EPSILON) overshoot = true;
                xp = 1.0;
            }
            if (xs > 1.0) {
                if (a.iota != 0.0 && m_log(xs) > NUMERICAL_EPSILON) overshoot = true;
                xs = 1.0;
            }
            double x = a.rho * xp + a.iota * xs, y = a.rho * (1.0 - xp) + a.iota * (1.0 - xs);
            t = fma(ab.x, x, fma(ab.y, y, gs.x));
        }
        if (t < 1e-30) { // rare: keep the running product a normal number
            if (t <= 0.0) {
                zero = true;
                t = 1.0;
            } else {
                int e2;
                t = frexp(t, &e2) * 2.0;
                ex += e2 - 1;
            }
        }
        acc *= t;
        if (++k == 8) { // <= 8 factors >= 1e-30 each: still a normal number; pull the exponent out
            k = 0;
            int hi = d_hi(acc);
            ex += ((hi >> 20) & 0x7ff) - 1023;
            acc = d_make((hi & 0x800fffff) | (1023 << 20), d_lo(acc));
        }
    }
    int hi = d_hi(acc);
    ex += ((hi >> 20) & 0x7ff) - 1023;
    acc_out = d_make((hi & 0x800fffff) | (1023 << 20), d_lo(acc));
    ex_out = ex;
    zero_out = zero;
    overshoot_out = overshoot;
}

// Pileup log-likelihood of sample s at (vaf, contaminant vaf): likelihood.rs:122-158 / :227-249.
VLR_DEV double sample_likelihood(Ctx& c, int s, double vaf, double vaf_by) {
    const vlr_sample_t& sm = c.sc->samples[s];
    const int n = c.n_obs[s];
    if (n == 0) return 0.0; // empty fold = ln 1
    const double ksum = c.ksum[s];
    if (ksum != ksum) return NAN;
    c.n_pileup_evals++;
    PileupArgs a;
    a.rho = 1.0;
    a.iota = 0.0;
    if (sm.contamination_by >= 0) {
        a.rho = 1.0 - sm.contamination_fraction; // e^{purity}
        a.iota = 1.0 - a.rho;                    // e^{impurity} (likelihood.rs:77-84)
    }
    // x = rho * xp + iota * xs with xp = (vaf == 1 ? 1 : vaf * s_r), y = 1 - x accordingly
    a.vaf = vaf;
    a.vaf_by = vaf_by;
    a.p1 = vaf == 1.0;
    a.s1 = vaf_by == 1.0;
    const bool sec = a.iota != 0.0;
    a.X1 = (a.p1 ? 0.0 : a.rho * vaf) + ((sec && !a.s1) ? a.iota * vaf_by : 0.0);
    a.X0 = (a.p1 ? a.rho : 0.0) + ((sec && a.s1) ? a.iota : 0.0);
    a.Yp = (a.p1 ? 0.0 : a.rho * (1.0 - vaf)) + ((sec && !a.s1) ? a.iota * (1.0 - vaf_by) : 0.0);
    a.Y0 = (a.p1 ? 0.0 : a.rho) + ((sec && !a.s1) ? a.iota : 0.0);
    a.xu = a.X1 + a.X0;
    double acc;
    int ex;
    bool zero, overshoot;
    const int mode = c.s_one[s] ? 0 : (c.s_gt1[s] ? 2 : 1);
    // careful, rarely-hot path: one generic-pointer loop (shared or global arena) keeps the code small
    pileup_product(reinterpret_cast<const double2*>(c.coef) + (size_t)c.coef_off[s] * 2, n, mode, a, acc, ex, zero, overshoot);
    acc = w_mul_d(acc); // 32 mantissas in [1,2): < 2^32
    ex = w_sum_i(ex);
    if (mode == 2 && w_any(overshoot)) c.status |= VLR_ST_OVERSHOOT;
    if (w_any(zero) || ksum == neg_inf()) return neg_inf();
    if (acc != acc) {
        c.status |= VLR_ST_NAN;
        return NAN;
    }
    return (m_log(acc) + (double)ex * LN_2) + ksum;
}

VLR_DEV_NOINLINE double sample_likelihood_call(Ctx& c_, int s, double vaf, double vaf_by) {
    Ctx& c = warp_ctx(c_);
    return sample_likelihood(c, s, vaf, vaf_by);
}

// generic.rs:43-53 keeps an LRU of 10000 pileup likelihoods per sample; a VAF-tree walk revisits only the few most
// recent (vaf, contaminant vaf) pairs of the outer levels, so LC_WAYS entries per sample catch the same hits.
VLR_DEV_NOINLINE double cached_sample_likelihood(Ctx& c_, int s, double vaf, double vaf_by) {
    Ctx& c = warp_ctx(c_);
    int n = c.lc_n[s];
    int ways = n < LC_WAYS ? n : LC_WAYS;
    for (int w = 0; w < ways; ++w)
        if (c.lc_k1[s][w] == vaf && c.lc_k2[s][w] == vaf_by) return c.lc_v[s][w];
    double v = sample_likelihood_call(c, s, vaf, vaf_by);
    int slot = n % LC_WAYS;
    c.lc_k1[s][slot] = vaf;
    c.lc_k2[s][slot] = vaf_by;
    c.lc_v[s][slot] = v;
    c.lc_n[s] = n + 1;
    return v;
}

// ------------------------------------------------------------------------------------------------ prior
VLR_DEV double vtf(const Ctx& c) { return c.sc->vtf[c.vartype]; }
VLR_DEV bool prior_semr(const Ctx& c, int s, double& out) { // prior.rs:250-257
    if (!(c.semr_override != c.semr_override)) {
        out = c.semr_override;
        return true;
    }
    double r = c.sc->samples[s].somatic_effective_mutation_rate;
    if (r != r) return false;
    out = m_log(r * vtf(c));
    return true;
}
VLR_DEV bool prior_het(const Ctx& c, double& out) { // prior.rs:263-270
    if (!(c.het_override != c.het_override)) {
        out = c.het_override;
        return true;
    }
    double h = c.sc->heterozygosity;
    if (h != h) return false;
    out = m_log(m_exp(m_log(h)) * vtf(c));
    return true;
}
VLR_DEV_NOINLINE bool universe_contains_sc(const DevScenario* sc, int s, double v) {
    const vlr_sample_t& sm = sc->samples[s];
    for (int i = 0; i < sm.n_universe; ++i) {
        const vlr_spectrum_t& sp = sc->spectra[sm.universe_offset + i];
        if (sp.kind == VLR_SPECTRUM_SET) {
            for (int j = 0; j < sp.n_vafs; ++j)
                if (sc->set_vafs[sp.vaf_offset + j] == v) return true;
        } else {
            Range r{sp.start, sp.end, sp.left_exclusive != 0, sp.right_exclusive != 0};
            if (range_contains(r, v)) return true;
        }
    }
    return false;
}
VLR_DEV bool universe_contains(const Ctx& c, int s, double v) { return universe_contains_sc(c.sc, s, v); }
VLR_DEV double prob_somatic_mutation(double rate, double somatic_vaf) { // prior.rs:440-456
    if (relative_eq(somatic_vaf, 0.0)) return ln_one_minus_exp(rate);
    return rate;
}
VLR_DEV double binomial_coeff(unsigned n, unsigned k) { // statrs factorial::binomial
    if (k > n) return 0.0;
    if (n <= 30) { // floor(0.5 + exp(ln n! - ln k! - ln (n-k)!)) is the exact integer for small n: compute it exactly
        // (ploidies are 1..4 in practice; the log/exp form below was 57 % of the pedigree kernel's instructions)
        if (k > n - k) k = n - k;
        unsigned long long r = 1;
        for (unsigned i = 1; i <= k; ++i) r = r * (unsigned long long)(n - k + i) / (unsigned long long)i;
        return (double)r;
    }
    double fn = 1.0, fk = 1.0, fnk = 1.0;
    for (unsigned i = 2; i <= n; ++i) fn *= (double)i;
    for (unsigned i = 2; i <= k; ++i) fk *= (double)i;
    for (unsigned i = 2; i <= n - k; ++i) fnk *= (double)i;
    return floor(0.5 + m_exp(m_log(fn) - m_log(fk) - m_log(fnk)));
}
VLR_DEV_NOINLINE double prob_select(unsigned ploidy, unsigned source_alt, unsigned target_alt, unsigned target_ref) {
    unsigned draws = target_alt + target_ref, x = target_alt; // Hypergeometric(N = ploidy, K = source_alt, n = draws).pmf(x)
    double pmf = x > draws ? 0.0
                           : binomial_coeff(source_alt, x) * binomial_coeff(ploidy - source_alt, draws - x) /
                                 binomial_coeff(ploidy, draws);
    return m_log(pmf);
}
VLR_DEV_NOINLINE double prob_mendelian_alt_counts(Ctx& c, unsigned sp0, unsigned sp1, unsigned tp, unsigned sa0, unsigned sa1,
                                         unsigned ta, double rate) { // prior.rs:600-678
    unsigned c0[2], c1[2];
    int n0, n1;
    if (sp0 % 2 == 0) {
        c0[0] = sp0 / 2;
        n0 = 1;
    } else {
        c0[0] = (unsigned)floor((double)sp0 / 2.0);
        c0[1] = (unsigned)ceil((double)sp0 / 2.0);
        n0 = 2;
    }
    if (sp1 % 2 == 0) {
        c1[0] = sp1 / 2;
        n1 = 1;
    } else {
        c1[0] = (unsigned)floor((double)sp1 / 2.0);
        c1[1] = (unsigned)ceil((double)sp1 / 2.0);
        n1 = 2;
    }
    Lse acc;
    acc.init();
    bool valid = false;
    const double ln_rate = m_log(rate);
    for (int i = 0; i < n0; ++i)
        for (int j = 0; j < n1; ++j) {
            unsigned p1 = c0[i], p2 = c1[j];
            if (p1 + p2 != tp) continue;
            valid = true;
            unsigned m1 = sa0 < p1 ? sa0 : p1, m2 = sa1 < p2 ? sa1 : p2;
            for (unsigned a1 = 0; a1 <= m1; ++a1)
                for (unsigned a2 = 0; a2 <= m2; ++a2)
                    if (a1 + a2 <= ta) {
                        double p = prob_select(sp0, sa0, a1, p1 - a1) + prob_select(sp1, sa1, a2, p2 - a2);
                        int missing = (int)ta - (int)(a1 + a2);
                        acc.add(p + ln_rate * (double)missing);
                    }
        }
    if (!valid) {
        c.status |= VLR_ST_NAN; // the reference panics (prior.rs:672-676)
        return neg_inf();
    }
    return acc.value();
}

// prior.rs:298-384, recursion end: all germline VAFs chosen
VLR_DEV_NOINLINE double prior_leaf(Ctx& c_, const Ops& ev, const double* g) {
    Ctx& c = warp_ctx(c_);
    const DevScenario* sc = c.sc;
    const int S = sc->S;
    double prob = 0.0, het;
    if (prior_het(c, het)) { // prior.rs:554-582
        unsigned m = 0, n = 0;
        for (int s = 0; s < S; ++s) {
            const vlr_sample_t& sm = sc->samples[s];
            if (sm.inheritance == VLR_INHERIT_NONE && sm.ploidy >= 0 && !sm.uniform_prior) {
                m += (unsigned)round((double)sm.ploidy * g[s]);
                n += (unsigned)sm.ploidy;
            }
        }
        if (m > 0) {
            prob = het - m_log((double)m);
        } else {
            Lse acc;
            acc.init();
            for (unsigned i = 1; i <= n; ++i) acc.add(het - m_log((double)i));
            prob = ln_one_minus_exp(acc.value());
        }
    }
    double sum = 0.0;
    for (int s = 0; s < S; ++s) {
        const vlr_sample_t& sm = sc->samples[s];
        if (sm.uniform_prior) continue;
        double rate;
        switch (sm.inheritance) {
        case VLR_INHERIT_MENDELIAN: {
            const vlr_sample_t& pa = sc->samples[sm.parent_a];
            const vlr_sample_t& pb = sc->samples[sm.parent_b];
            unsigned na = (unsigned)round(g[sm.parent_a] * (double)pa.ploidy);
            unsigned nb = (unsigned)round(g[sm.parent_b] * (double)pb.ploidy);
            unsigned nc = (unsigned)round(g[s] * (double)sm.ploidy);
            double gr = sm.germline_mutation_rate * vtf(c);
            double p = prob_mendelian_alt_counts(c, (unsigned)pa.ploidy, (unsigned)pb.ploidy, (unsigned)sm.ploidy, na, nb,
                                                 nc, gr);
            if (prior_semr(c, s, rate)) p += prob_somatic_mutation(rate, ev.vaf[s] - g[s]);
            sum += p;
            break;
        }
        case VLR_INHERIT_CLONAL: { // prior.rs:458-514
            int parent = sm.parent_a;
            double p;
            if (!relative_eq(g[s], g[parent])) p = neg_inf();
            else {
                bool has_rate = prior_semr(c, s, rate);
                double psv = ev.vaf[parent] - g[parent], sv = ev.vaf[s] - g[s];
                if (sm.clonal_somatic && has_rate) p = (psv != 0.0) ? 0.0 : prob_somatic_mutation(rate, sv);
                else if (sm.clonal_somatic) p = relative_eq(sv, psv) ? 0.0 : neg_inf();
                else if (has_rate) p = prob_somatic_mutation(rate, sv);
                else p = 0.0;
            }
            sum += p;
            break;
        }
        case VLR_INHERIT_SUBCLONAL: { // prior.rs:516-552
            int parent = sm.parent_a;
            double p;
            if (!relative_eq(g[s], g[parent])) p = neg_inf();
            else if (prior_semr(c, s, rate)) {
                p = (ev.vaf[parent] == 0.0 && g[s] == 0.0) ? prob_somatic_mutation(rate, ev.vaf[s]) : 0.0;
            } else {
                p = relative_eq(ev.vaf[s] - g[s], ev.vaf[parent] - g[parent]) ? 0.0 : neg_inf();
            }
            sum += p;
            break;
        }
        default:
            if (prior_semr(c, s, rate)) sum += prob_somatic_mutation(rate, ev.vaf[s] - g[s]);
        }
    }
    prob += sum;
    if (!(prob <= 0.0)) c.status |= VLR_ST_PRIOR_POSITIVE; // assert!(*prob <= 0.0), prior.rs:378
    return prob;
}

// prior.rs:385-437: the recursion over germline VAF assignments, unrolled into an odometer. Nested ln_sum_exp of the
// reference == one ln_sum_exp over the flattened product (rounding-level difference only).
VLR_DEV_NOINLINE double prior_full(Ctx& c_, const Ops& ev) {
    Ctx& c = warp_ctx(c_);
    const DevScenario* sc = c.sc;
    const int S = sc->S;
    int nchoice[MAXS], idx[MAXS];
    double g[MAXS];
    bool any_multi = false;
    for (int s = 0; s < S; ++s) {
        const vlr_sample_t& sm = sc->samples[s];
        double v = ev.vaf[s];
        idx[s] = 0;
        if (sm.ploidy == 0 && v != 0.0) return neg_inf();
        if (sm.uniform_prior) {
            if (!universe_contains(c, s, v)) return neg_inf();
            nchoice[s] = 1;
            g[s] = 0.0;
        } else if (!(sm.somatic_effective_mutation_rate != sm.somatic_effective_mutation_rate)) {
            if (sm.ploidy < 0) { // unreachable!() in the reference
                c.status |= VLR_ST_NAN;
                return neg_inf();
            }
            nchoice[s] = sm.ploidy + 1;
            g[s] = 0.0;
            any_multi = any_multi || nchoice[s] > 1;
        } else if (sm.ploidy >= 0 && !(sc->heterozygosity != sc->heterozygosity)) {
            double n_alt = (double)sm.ploidy * v;
            if (!relative_eq(n_alt, round(n_alt))) return neg_inf();
            nchoice[s] = 1;
            g[s] = v;
        } else {
            c.status |= VLR_ST_NAN; // unreachable!() in the reference
            return neg_inf();
        }
    }
    if (!any_multi) return prior_leaf(c, ev, g);
    Lse acc;
    acc.init();
    for (;;) {
        for (int s = 0; s < S; ++s) {
            const vlr_sample_t& sm = sc->samples[s];
            if (nchoice[s] > 1 || (!sm.uniform_prior && !(sm.somatic_effective_mutation_rate !=
                                                         sm.somatic_effective_mutation_rate)))
                g[s] = sm.ploidy > 0 ? (double)idx[s] / (double)sm.ploidy : 0.0;
        }
        acc.add(prior_leaf(c, ev, g));
        int s = S - 1;
        for (; s >= 0; --s) {
            if (++idx[s] < nchoice[s]) break;
            idx[s] = 0;
        }
        if (s < 0) break;
    }
    return acc.value();
}

// Prior::compute (prior.rs:718-761)
VLR_DEV_NOINLINE double prior_compute_uncached(Ctx& c_, const Ops& ev);

// All-discrete VAF vector, no per-record prior overrides: look the prior up in the context's table, computing and
// publishing it on a miss. Lane 0 talks to the table and broadcasts, so the warp's control flow stays uniform.
VLR_DEV_NOINLINE double prior_tab_compute(Ctx& c_, const Ops& ev) {
    Ctx& c = warp_ctx(c_);
    const DevScenario* sc = c.sc;
    PriorTabEntry* tab = sc->prior_tab;
    if (tab == nullptr || !(c.het_override != c.het_override) || !(c.semr_override != c.semr_override))
        return prior_compute_uncached(c, ev);
    const int S = sc->S;
    unsigned h = 2166136261u ^ (unsigned)c.vartype;
    for (int s = 0; s < S; ++s) {
        h = (h ^ (unsigned)d_lo(ev.vaf[s])) * 16777619u;
        h = (h ^ (unsigned)d_hi(ev.vaf[s])) * 16777619u;
    }
    h ^= h >> 16; // murmur3 finaliser: the VAFs' low mantissa bits are all zero, FNV alone leaves the low hash bits equal
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    // ---- probe (lane 0): found -> value; else the first empty slot (or none)
    int found = 0, slot = -1;
    double value = 0.0;
    unsigned side = 0;
    if (lane_id() == 0) {
        for (int k = 0; k < 8; ++k) {
            PriorTabEntry* e = tab + ((h + (unsigned)k) & (PRIOR_TAB_N - 1));
#ifdef VLR_HOST_EMU
            const unsigned st = e->state;
#else
            unsigned st; // acquire: the entry's fields are read after (and only if) the state says "readable"
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(st) : "l"(&e->state) : "memory");
#endif
            if (st == 2u) {
                bool same = e->vartype == c.vartype;
                for (int s = 0; s < S; ++s) same = same && (e->vaf[s] == ev.vaf[s]);
                if (same) {
                    found = 1;
                    value = e->value;
                    side = e->side_status;
                    break;
                }
            } else if (st == 0u) {
                slot = (int)((h + (unsigned)k) & (PRIOR_TAB_N - 1));
                break;
            } // st == 1: someone is writing it; keep probing (worst case we compute without publishing)
        }
    }
#ifndef VLR_HOST_EMU
    found = __shfl_sync(FULL, found, 0, LANES);
    slot = __shfl_sync(FULL, slot, 0, LANES);
    value = __shfl_sync(FULL, value, 0, LANES);
    side = __shfl_sync(FULL, side, 0, LANES);
#endif
    if (found) {
        c.status |= side;
        return value;
    }
    const uint32_t s0 = c.status;
    c.status = 0;
    const double p = prior_compute_uncached(c, ev);
    side = c.status;
    c.status = s0 | side;
    if (slot >= 0 && lane_id() == 0) {
        PriorTabEntry* e = tab + slot;
#ifdef VLR_HOST_EMU
        const bool won = e->state == 0u;
        if (won) e->state = 1u;
#else
        const bool won = atomicCAS(&e->state, 0u, 1u) == 0u;
#endif
        if (won) {
            e->vartype = c.vartype;
            e->side_status = side;
            for (int s = 0; s < S; ++s) e->vaf[s] = ev.vaf[s];
            e->value = p;
#ifdef VLR_HOST_EMU
            e->state = 2u;
#else
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&e->state), "r"(2u) : "memory");
#endif
        }
    }
    warp_sync();
    return p;
}

VLR_DEV_NOINLINE double prior_compute(Ctx& c_, const Ops& ev) {
    Ctx& c = warp_ctx(c_);
    const int S = c.sc->S;
    const bool discrete = (ev.disc_mask & ((1u << S) - 1u)) == ((1u << S) - 1u);
    if (!discrete) return prior_compute_uncached(c, ev);
    for (int i = 0; i < c.pc_n; ++i) {
        bool same = true;
        for (int s = 0; s < S; ++s) same = same && (c.pc_key[i][s] == ev.vaf[s]);
        if (same) return c.pc_val[i];
    }
    const double p = prior_tab_compute(c, ev);
    if (c.pc_n < PRIOR_CACHE) {
        const int i = c.pc_n;
        for (int s = 0; s < S; ++s) c.pc_key[i][s] = ev.vaf[s];
        c.pc_val[i] = p;
        c.pc_n = i + 1;
    }
    return p;
}
VLR_DEV_NOINLINE double prior_compute_uncached(Ctx& c_, const Ops& ev) {
    Ctx& c = warp_ctx(c_);
    const DevScenario* sc = c.sc;
    if (!sc->full_prior && !sc->all_uniform) {
        bool absent = true;
        for (int s = 0; s < sc->S; ++s)
            if (ev.vaf[s] != 0.0) absent = false;
        if (!absent) {
            double full = prior_full(c, ev);
            if (full == neg_inf()) return full;
            if (c.prior_absent != c.prior_absent) {
                Ops z;
                for (int s = 0; s < MAXS; ++s) z.vaf[s] = 0.0;
                z.set_mask = (1u << sc->S) - 1u;
                z.disc_mask = z.set_mask;
                z.lfc_mask = 0;
                c.prior_absent = prior_full(c, z);
            }
            return ln_one_minus_exp(c.prior_absent);
        }
    }
    return prior_full(c, ev);
}

// ------------------------------------------------------------------------------------------------ joint
// GenericLikelihood::compute (generic.rs:500-554) + Prior + rust-bio Model::joint_prob bookkeeping.
// `od` indexes the operand stack c.ops[]. FORCE-INLINED: its one hot call site is the evaluation site of the
// leaf-level adaptive integrator; everything else reaches it through joint_call().
VLR_DEV_NOINLINE bool node_contains(const DevScenario* sc, int ni, const double* vaf, uint32_t& lfcs, int exclude);

// `best_event.contains(map_estimates, None)` (calling.rs:861-864, vaftree.rs:42-51) for event e
VLR_DEV_NOINLINE bool event_contains(const DevScenario* sc, int e, const double* vaf, uint32_t lfc_mask) {
    const vlr_event_t& ev = sc->events[e];
    for (int r = 0; r < ev.n_roots; ++r) {
        uint32_t l = lfc_mask;
        if (node_contains(sc, ev.first_root + r, vaf, l, -1)) return true;
    }
    return false;
}

// The reference keeps ONE map of base events and reports the best one the strongest event contains (calling.rs:851-864).
// Offer a base event to the MAP slot of every event (same plain / artifact class): needed when events overlap or when a
// point on an excluded range bound was evaluated (fewer than 10 reads, narrow ranges: formula.rs:1172-1224), which
// belongs to another event than the one whose tree produced it.
VLR_DEV_NOINLINE void map_offer_all(Ctx& c_, const Ops& ops, double j) {
    Ctx& c = warp_ctx(c_);
    const DevScenario* sc = c.sc;
    const int S = sc->S, E = sc->E;
    const int cls = c.art.id != 0 ? 1 : 0;
    // Recording order of the oracle (= universe order: event, plain before twin, then evaluation order) decides between
    // base events of equal probability (the reference: HashMap order). Two events can evaluate the SAME allele
    // frequencies, once as a discrete and once as a continuous event (e.g. normal = 0.0 as the set value of
    // somatic_tumor and as the excluded lower bound of somatic_normal): an exact tie upstream, so it is decided by the
    // recording order here too, whatever the last bit of the two evaluations.
    const uint32_t seq = ((uint32_t)c.cur_slot << 24) | (c.n_base & 0xffffffu);
    for (int e = 0; e < E; ++e) {
        const int slot = 2 * e + cls;
        if (c.map_set[slot]) {
            bool same = c.map_cfg[slot] == c.art.id;
            for (int s = 0; s < S; ++s) same = same && c.map_vaf[slot][s] == ops.vaf[s];
            const double cur = c.map_joint[slot];
            const bool better = same ? seq < c.map_seq[slot] : (j > cur || (j == cur && seq < c.map_seq[slot]));
            if (!better) continue;
        }
        if (!event_contains(sc, e, ops.vaf, ops.lfc_mask)) continue;
        c.map_set[slot] = 1;
        c.map_joint[slot] = j;
        c.map_cfg[slot] = c.art.id;
        c.map_disc[slot] = ops.disc_mask;
        c.map_seq[slot] = seq;
        for (int s = 0; s < S; ++s) c.map_vaf[slot][s] = ops.vaf[s];
    }
}

// base-event log filter of the second pass (Ctx::be_filter)
VLR_DEV bool be_keep(const Ctx& c, const double* vaf, uint32_t disc) {
    if (!c.be_filter) return true;
    int mismatches = 0;
    for (int s = 0; s < c.sc->S; ++s)
        mismatches += !(vaf[s] == c.flt_vaf[s] && ((disc >> s) & 1u) == ((c.flt_disc >> s) & 1u));
    return mismatches <= 1;
}

VLR_DEV double joint(Ctx& c, int od) {
    const DevScenario* sc = c.sc;
    const int S = sc->S;
    const Ops& ops = c.ops[od];
    double prior;
    if (sc->all_uniform) { // every sample declares a universe: flat prior inside it (prior.rs:385-406, :737)
        prior = 0.0;
        for (int s = 0; s < S; ++s) {
            const double v = ops.vaf[s];
            if ((sc->samples[s].ploidy == 0 && v != 0.0) || !universe_contains(c, s, v)) prior = neg_inf();
        }
    } else {
        prior = prior_compute(c, ops);
    }
    double lh = 0.0;
    bool lfc_ok = true;
    for (int k = 0; k < sc->n_lfc_nodes; ++k) {
        if (!(ops.lfc_mask & (1u << k))) continue;
        const vlr_node_t& nd = sc->nodes[sc->lfc_nodes[k]];
        double a = ops.vaf[nd.sample], b2 = ops.vaf[nd.sample_b];
        double lfc;
        if (a == 0.0 && b2 == 0.0) lfc = 0.0;
        else {
            lfc = m_log2_lfc(a) - m_log2_lfc(b2);
            if (lfc != lfc) c.status |= VLR_ST_NAN;
        }
        bool t;
        switch (nd.cmp) {
        case VLR_CMP_EQ: t = relative_eq(lfc, nd.lfc_value); break;
        case VLR_CMP_GT: t = lfc > nd.lfc_value; break;
        case VLR_CMP_GE: t = lfc >= nd.lfc_value; break;
        case VLR_CMP_LT: t = lfc < nd.lfc_value; break;
        case VLR_CMP_LE: t = lfc <= nd.lfc_value; break;
        default: t = !relative_eq(lfc, nd.lfc_value);
        }
        if (!t) {
            lfc_ok = false;
            break;
        }
    }
    if (!lfc_ok) lh = neg_inf();
    else {
        for (int s = 0; s < S; ++s) {
            const vlr_sample_t& sm = sc->samples[s];
            double by = sm.contamination_by >= 0 ? ops.vaf[sm.contamination_by] : 0.0;
            lh += cached_sample_likelihood(c, s, ops.vaf[s], by);
        }
    }
    double j = prior + lh;
    if (j != j) c.status |= VLR_ST_NAN;
    // rust-bio Model::joint_prob records every base event; only artifact-free ones can enter an AFD (calling.rs:912)
    if (c.be != nullptr && c.art.id == 0 && be_keep(c, ops.vaf, ops.disc_mask)) {
        if (c.n_rec < (uint32_t)BE_CAP) {
            // every lane stores the same record (no lane-divergent block in front of the uniform counter update)
            double* e = c.be + (int64_t)c.n_rec * (2 + S);
            e[0] = j;
            e[1] = d_make((int)ops.lfc_mask, (int)ops.disc_mask);
            for (int s = 0; s < S; ++s) e[2 + s] = ops.vaf[s];
            warp_sync();
            c.n_rec++;
        } else {
            c.status |= VLR_ST_BASE_EVENTS_OVERFLOW;
        }
    }
    // bookkeeping for MAP (calling.rs:851-870): best base event per (event, plain | artifact)
    c.n_base++;
    int slot = c.cur_slot;
    if (c.map_global) {
        map_offer_all(c, ops, j);
    } else if (!c.map_set[slot] || j > c.map_joint[slot]) {
        c.map_set[slot] = 1;
        c.map_joint[slot] = j;
        c.map_cfg[slot] = c.art.id;
        c.map_disc[slot] = ops.disc_mask;
        c.map_seq[slot] = ((uint32_t)slot << 24) | (c.n_base & 0xffffffu);
        for (int s = 0; s < S; ++s) c.map_vaf[slot][s] = ops.vaf[s];
    }
    return j;
}
VLR_DEV_NOINLINE double joint_call(Ctx& c_, int od) {
    Ctx& c = warp_ctx(c_);
    return joint(c, od);
}

// ------------------------------------------------------------------------------------------------ density
VLR_DEV_NOINLINE double density(Ctx& c, int ni, int od, int level);

// Value of the subtree below `node` for the operands c.ops[od] (the `subdensity` closure of generic.rs:199-231).
VLR_DEV_NOINLINE double subdensity(Ctx& c_, const vlr_node_t& node, int od, int level) {
    Ctx& c = warp_ctx(c_);
    double p;
    if (node.n_children == 0) p = joint_call(c, od);
    else if (node.n_children > 1) {
        Lse acc;
        acc.init();
        for (int k = 0; k < node.n_children; ++k) {
            c.ops[od + 1] = c.ops[od];
            acc.add(density(c, node.first_child + k, od + 1, level));
        }
        p = acc.value();
    } else {
        p = density(c, node.first_child, od, level);
    }
    if (p != p) c.status |= VLR_ST_NAN;
    return p;
}

VLR_DEV bool ops_lfc_bounds(const Ctx& c, const Ops& ops, int sample, Range& out) { // generic.rs:148-174
    const DevScenario* sc = c.sc;
    bool have = false;
    for (int k = 0; k < sc->n_lfc_nodes; ++k) {
        if (!(ops.lfc_mask & (1u << k))) continue;
        const vlr_node_t& nd = sc->nodes[sc->lfc_nodes[k]];
        bool got = false;
        Range b;
        if (nd.sample == sample) {
            if (ops.set_mask & (1u << nd.sample_b)) {
                int cmp = lfc_invert_cmp(nd.cmp);
                double val = (nd.cmp == VLR_CMP_EQ || nd.cmp == VLR_CMP_NE) ? nd.lfc_value : -nd.lfc_value;
                b = lfc_infer_bounds(cmp, val, ops.vaf[nd.sample_b]);
                got = true;
            }
        } else if (nd.sample_b == sample) {
            if (ops.set_mask & (1u << nd.sample)) {
                b = lfc_infer_bounds(nd.cmp, nd.lfc_value, ops.vaf[nd.sample]);
                got = true;
            }
        }
        if (got) {
            out = have ? range_intersect(out, b) : b;
            have = true;
        }
    }
    return have;
}

// Integration limits on an excluded bound of the node's range (formula.rs:1172-1224 returns the bound itself for fewer
// than 10 reads or narrow ranges): that point is evaluated but belongs to another event -> global MAP bookkeeping.
VLR_DEV void mark_bound_points(Ctx& c, const Range& vafs, double min_vaf, double max_vaf) {
    if ((vafs.lex && min_vaf <= vafs.start) || (vafs.rex && max_vaf >= vafs.end)) c.map_global = 1;
}

VLR_DEV void ops_push(Ops& o, int sample, double vaf, bool discrete) {
    o.vaf[sample] = vaf;
    o.set_mask |= 1u << sample;
    if (discrete) o.disc_mask |= 1u << sample;
    else o.disc_mask &= ~(1u << sample);
}

// ln_simpsons_integrate_exp (rust-bio): interior points first, then the two ends
VLR_DEV_NOINLINE double integrate_simpson(Ctx& c_, const vlr_node_t& node, int od, double a, double b, int n, int level) {
    Ctx& c = warp_ctx(c_);
    Lse acc;
    acc.init();
    const double step = (b - a) / (double)(n - 1);
    for (int j = 0; j < n; ++j) {
        double x, lw = 0.0;
        if (j < n - 2) {
            const int i = j + 1;
            x = a + step * (double)i;
            lw = m_log((double)(2 + (i % 2) * 2));
        } else {
            x = (j == n - 2) ? a : b;
        }
        c.ops[od + 1] = c.ops[od];
        ops_push(c.ops[od + 1], node.sample, x, false);
        acc.add(subdensity(c, node, od + 1, level + 1) + lw);
    }
    return acc.value() + m_log(b - a) - m_log((double)(n - 1)) - m_log(3.0);
}

// ln_trapezoidal_integrate_grid_exp over the n visited points of `level` (rust-bio): rank sort with
// lanes over points, then lanes over intervals and one warp-wide log-sum-exp.
VLR_DEV_NOINLINE double grid_trapezoid(Ctx& c_, const double* gx, const double* gf, int n) {
    Ctx& c = warp_ctx(c_);
    const int lane = lane_id();
    double* sx = c.ws->sort_x;
    double* sf = c.ws->sort_f;
    warp_sync();
#pragma unroll 1
    for (int a = lane; a < n; a += LANES) {
        const double xi = gx[a];
        int rank = 0;
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const double xj = gx[j];
            rank += (xj < xi) || (xj == xi && j < a);
        }
        sx[rank] = xi;
        sf[rank] = gf[a];
    }
    warp_sync();
    // interval terms ln((f_i + f_{i+1}) / 2 * dx), kept in registers for the sum pass (n <= GRID_CAP = 4 x 32 lanes)
    constexpr int PER = (GRID_CAP + LANES - 1) / LANES;
    double tl[PER];
    double tmax = neg_inf();
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int a = lane + q * LANES;
        double t = neg_inf();
        if (a + 1 < n) {
            const double dx = sx[a + 1] - sx[a];
            if (dx > 0.0) t = ln_add_exp(sf[a], sf[a + 1]) + m_log(dx) - LN_2;
            if (t != t) t = INFINITY; // poison through the max
        }
        tl[q] = t;
        tmax = fmax(tmax, t);
    }
    tmax = w_max_d(tmax);
    if (tmax == neg_inf()) return neg_inf();
    if (tmax == INFINITY) {
        c.status |= VLR_ST_NAN;
        return NAN;
    }
    double ssum = 0.0;
#pragma unroll
    for (int q = 0; q < PER; ++q)
        if (tl[q] != neg_inf()) ssum += m_exp(tl[q] - tmax);
    ssum = w_sum_d(ssum);
    warp_sync();
    return tmax + m_log(ssum);
}

VLR_DEV_NOINLINE bool try_child_batch(Ctx& c_, const vlr_node_t& node, int od, int k, const double* xs, double* fs);

// Generic adaptive integration: every point goes through subdensity()/joint() like in the reference, unless the
// whole batch can be handed to the concurrent leaf integrator (try_child_batch).
VLR_DEV_NOINLINE double integrate_adaptive_generic(Ctx& c_, const vlr_node_t& node, int od, double a, double b, double res,
                                                   int level) {
    Ctx& c = warp_ctx(c_);
    double* gx = c.ws->grid_x[level];
    double* gf = c.ws->grid_f[level];
    Adaptive& st = c.ad[level];
    st.init(a, b, res);
    int n = 0;
    bool overflow = false;
    double* xs = c.xs[level];
    double* fs = c.fs[level];
    for (;;) {
        const int k = st.points(xs);
        const bool batched = try_child_batch(c, node, od, k, xs, fs);
        for (int i = 0; i < k; ++i) {
            if (!batched) {
                c.ops[od + 1] = c.ops[od];
                ops_push(c.ops[od + 1], node.sample, xs[i], false);
                fs[i] = subdensity(c, node, od + 1, level + 1);
            }
            const double f = fs[i];
            if (n < GRID_CAP) {
                gx[n] = xs[i]; // every lane stores the same value: no divergence
                gf[n] = f;
                n++;
            } else {
                overflow = true;
            }
        }
        if (!st.consume(xs, fs, overflow)) break;
    }
    if (overflow) c.status |= VLR_ST_GRID_OVERFLOW;
    return grid_trapezoid(c, gx, gf, n);
}

// ---- leaf fast path -----------------------------------------------------------------------------------------
// Lanes = evaluations. A warp advances up to MT leaf-level integrations of one locus at the same time (the abscissae
// of the enclosing integration's batch, i.e. different VAFs of the parent sample) and, per step, all abscissae the
// adaptive searches ask for: every (task, abscissa) pair is a SLOT, the 32 lanes are split evenly over the slots
// (G = 32 / next_pow2(#slots) lanes per slot striding over the reads), so one instruction stream serves up to 32
// evaluations. The reads loop is a dozen instructions that stay in the L0 instruction cache; the coefficient loads
// are shared-memory broadcasts. (The first versions evaluated one point per warp and were instruction-fetch bound:
// profiles/README.md.)

// ln-likelihood of the dependent pileups for the slots [base, base + m) -> c.slot_f / c.slot_slow
template <bool SM>
VLR_DEV void multi_eval_impl(Ctx& c, int base, int m) {
    const int lane = lane_id();
    int p2 = 1;
    while (p2 < m) p2 <<= 1;
    const int G = LANES / p2 > 0 ? LANES / p2 : 1; // lanes per slot
    const int slot_in = lane / G, sub = lane - slot_in * G;
    const bool valid = slot_in < m;
    const int slot = base + (valid ? slot_in : 0);
    const double x = c.slot_x[slot];
    const MultiTask& task = c.mt[c.slot_task[slot]];
    double lnl = 0.0;
    bool slow = false;
    const int dep_hi = c.leaf.dep_hi;
    for (int di = c.leaf.dep_lo; di < dep_hi; ++di) {
        const LeafDep& d = c.leaf.dep[di];
        const int n = d.n;
        if (n == 0) continue; // empty fold = ln 1
        const double rho = d.rho, iota = d.iota;
        const double vaf = d.vaf_is_x ? x : (d.vaf_is_parent ? task.parent_x : d.fixed_vaf);
        const double vby = d.has_by ? (d.by_is_x ? x : (d.by_is_parent ? task.parent_x : d.fixed_by)) : 0.0;
        const bool p1 = vaf == 1.0, s1 = vby == 1.0, sec = iota != 0.0;
        // x = rho xp + iota xs with xp = (vaf == 1 ? 1 : vaf s_r), y = 1 - x without cancellation; see sample_likelihood
        const double X1 = (p1 ? 0.0 : rho * vaf) + ((sec && !s1) ? iota * vby : 0.0);
        const double X0 = (p1 ? rho : 0.0) + ((sec && s1) ? iota : 0.0);
        const double Yp = (p1 ? 0.0 : rho * (1.0 - vaf)) + ((sec && !s1) ? iota * (1.0 - vby) : 0.0);
        const double xu = X1 + X0;
        const double2* co = d.co;
#ifndef VLR_HOST_EMU
        if (SM) co = reinterpret_cast<const double2*>(vlr_smem + (__cvta_generic_to_shared(d.co) - __cvta_generic_to_shared(vlr_smem)));
#endif
        double acc = 1.0;
        int ex = 0, k = 0;
        const int nn = valid ? n : 0;
#pragma unroll 1
        for (int r = sub; r < nn; r += G) {
            const double2 ab = co[2 * r];
            const double2 gu = co[2 * r + 1];
            // per read x_r = xu - u_r X1, y_r = Yp + u_r X1 with u_r = 1 - s_r (u_r = 0 reproduces xu / Yp exactly)
            const double xr = fma(-gu.y, X1, xu);
            const double yr = fma(gu.y, X1, Yp);
            acc *= fma(ab.x, xr, fma(ab.y, yr, gu.x));
            if (++k == 4) { // <= 4 factors between exponent pulls; a tiny (< 1e-60) or zero factor -> careful path
                k = 0;
                slow = slow || !(acc >= 1e-240);
                const int hi = d_hi(acc);
                ex += ((hi >> 20) & 0x7ff) - 1023;
                acc = d_make((hi & 0x800fffff) | (1023 << 20), d_lo(acc));
            }
        }
        {
            slow = slow || !(acc >= 1e-240);
            const int hi = d_hi(acc);
            ex += ((hi >> 20) & 0x7ff) - 1023;
            acc = d_make((hi & 0x800fffff) | (1023 << 20), d_lo(acc));
        }
#ifndef VLR_HOST_EMU
        for (int o = G >> 1; o > 0; o >>= 1) { // butterfly inside the slot's lane group
            acc *= __shfl_xor_sync(FULL, acc, o);
            ex += __shfl_xor_sync(FULL, ex, o);
            slow = slow || (__shfl_xor_sync(FULL, slow ? 1 : 0, o) != 0);
        }
#endif
        lnl += (log(acc) + (double)ex * LN_2) + d.ksum; // a NaN ksum (invalid inputs) propagates
    }
    if (valid && sub == 0) {
        c.slot_f[slot] = lnl;
        c.slot_slow[slot] = slow ? 1 : 0;
    }
    warp_sync();
}
// the shared-memory variant is inlined into the step loop (measured: +18 % over an out-of-line call); the
// global-memory variant (loci deeper than the shared arena) stays out of line to keep the hot function small
VLR_DEV void multi_eval_sm(Ctx& c_, int base, int m) { multi_eval_impl<true>(warp_ctx(c_), base, m); }
VLR_DEV_NOINLINE void multi_eval_gl(Ctx& c_, int base, int m) { multi_eval_impl<false>(warp_ctx(c_), base, m); }

// careful re-evaluation of the slots whose fast product met a zero / denormal-range factor (rare)
VLR_DEV_NOINLINE void multi_eval_slow(Ctx& c_, int total) {
    Ctx& c = warp_ctx(c_);
    for (int slot = 0; slot < total; ++slot) {
        if (!c.slot_slow[slot]) continue;
        const double x = c.slot_x[slot];
        const MultiTask& task = c.mt[c.slot_task[slot]];
        double lnl = 0.0;
        for (int di = c.leaf.dep_lo; di < c.leaf.dep_hi; ++di) {
            const LeafDep& d = c.leaf.dep[di];
            const double vaf = d.vaf_is_x ? x : (d.vaf_is_parent ? task.parent_x : d.fixed_vaf);
            const double vby = d.has_by ? (d.by_is_x ? x : (d.by_is_parent ? task.parent_x : d.fixed_by)) : 0.0;
            lnl += sample_likelihood_call(c, d.s, vaf, vby);
        }
        c.slot_f[slot] = lnl;
        c.slot_slow[slot] = 0;
    }
    warp_sync();
}

// Materialises the operands of (task, x) at c.ops[od + 1] (parent sample at the task's abscissa, leaf sample at x).
VLR_DEV void multi_ops(Ctx& c, int od, int task, double x) {
    c.ops[od + 1] = c.ops[od];
    if (c.leaf.parent >= 0) ops_push(c.ops[od + 1], c.leaf.parent, c.mt[task].parent_x, c.leaf.parent_disc);
    ops_push(c.ops[od + 1], c.leaf.t, x, false);
}

// Decides whether the fast path can serve the leaf integration(s) of `node` over [a, b] and hoists the constants.
// parent >= 0: n_tasks integrations that differ in the VAF of sample `parent` (parent_xs); parent < 0: one.
VLR_DEV_NOINLINE bool leaf_setup(Ctx& c_, const vlr_node_t& node, int od, double a, double b, double res, int parent,
                                 int n_tasks, const double* parent_xs, const int* event_slots = nullptr,
                                 bool parent_disc = false) {
    Ctx& c = warp_ctx(c_);
    const DevScenario* sc = c.sc;
    const int S = sc->S;
    const int t = node.sample;
    LeafFast& L = c.leaf;
    if (c.ops[od].lfc_mask != 0 || n_tasks > MT) return false;
    if (c.map_global || c.be_filter) return false; // every base event goes through joint() (offers to all events / filtered log)
    int n_dep = 0;
    for (int s = 0; s < S; ++s) {
        const int by = sc->samples[s].contamination_by;
        if (s != t && by != t) continue;
        if (n_dep == 2 || c.s_gt1[s]) return false;
        LeafDep& d = L.dep[n_dep];
        n_dep++;
        d.s = s;
        d.co = reinterpret_cast<const double2*>(c.coef) + (size_t)c.coef_off[s] * 2;
        d.ksum = c.ksum[s];
        d.n = c.n_obs[s];
        d.rho = 1.0;
        d.iota = 0.0;
        d.has_by = by >= 0;
        if (by >= 0) {
            d.rho = 1.0 - sc->samples[s].contamination_fraction;
            d.iota = 1.0 - d.rho;
        }
        d.vaf_is_x = s == t;
        d.by_is_x = by == t;
        d.vaf_is_parent = s == parent;
        d.by_is_parent = by >= 0 && by == parent;
        d.fixed_vaf = c.ops[od].vaf[s];
        d.fixed_by = by >= 0 ? c.ops[od].vaf[by] : 0.0;
    }
    L.n_dep = n_dep;
    L.n_tasks = n_tasks;
    L.parent_disc = parent_disc;
    L.parent = parent;
    L.t = t;
    L.coef_in_sm = c.coef_in_sm != 0;
    L.uniform = sc->all_uniform != 0;
    L.t_ploidy0 = sc->samples[t].ploidy == 0;
    L.record = c.be != nullptr && c.art.id == 0;
    bool per_point = true;
    if (L.uniform) {
        // one Range spectrum of t's universe covering [a, b] covers every abscissa of this integration
        const vlr_sample_t& sm = sc->samples[t];
        for (int i = 0; i < sm.n_universe; ++i) {
            const vlr_spectrum_t& sp = sc->spectra[sm.universe_offset + i];
            if (sp.kind != VLR_SPECTRUM_RANGE) continue;
            Range r{sp.start, sp.end, sp.left_exclusive != 0, sp.right_exclusive != 0};
            if (range_contains(r, a) && range_contains(r, b) && !(L.t_ploidy0 && b != 0.0)) per_point = false;
        }
    }
    L.prior_per_point = per_point;
    // ---- loop constants per task: likelihoods of the samples that do not see t (generic.rs:511-551 cache hits).
    // Only the parent sample's differs between tasks: its n_tasks pileup likelihoods are evaluated together, one
    // slot per task, with the same slot machinery as the leaf evaluations.
    double lh_shared = 0.0;
    bool parent_by_slots = false;
    for (int s = 0; s < S; ++s) {
        const int by = sc->samples[s].contamination_by;
        if (s == t || by == t) continue;
        if (s == parent && n_tasks > 1 && !c.s_gt1[s] && by != parent) {
            LeafDep& d = L.dep[2];
            d.s = s;
            d.co = reinterpret_cast<const double2*>(c.coef) + (size_t)c.coef_off[s] * 2;
            d.ksum = c.ksum[s];
            d.n = c.n_obs[s];
            d.rho = 1.0;
            d.iota = 0.0;
            d.has_by = by >= 0;
            if (by >= 0) {
                d.rho = 1.0 - sc->samples[s].contamination_fraction;
                d.iota = 1.0 - d.rho;
            }
            d.vaf_is_x = true; // the slot abscissa is the parent's VAF here
            d.by_is_x = false;
            d.vaf_is_parent = false;
            d.by_is_parent = false;
            d.fixed_vaf = 0.0;
            d.fixed_by = by >= 0 ? c.ops[od].vaf[by] : 0.0;
            parent_by_slots = true;
        } else if (s != parent) {
            if (by >= 0 && by == parent) return false; // a non-dependent sample contaminated by the parent: generic path
            lh_shared += cached_sample_likelihood(c, s, c.ops[od].vaf[s], by >= 0 ? c.ops[od].vaf[by] : 0.0);
        }
    }
    for (int k = 0; k < n_tasks; ++k) {
        c.mt[k].parent_x = parent >= 0 ? parent_xs[k] : 0.0;
        c.slot_x[k] = c.mt[k].parent_x;
        c.slot_task[k] = k;
    }
    if (parent_by_slots) {
        L.dep_lo = 2;
        L.dep_hi = 3;
        warp_sync();
        for (int base = 0; base < n_tasks; base += LANES)
            multi_eval_gl(c, base, n_tasks - base < LANES ? n_tasks - base : LANES);
        bool any_slow = false;
        for (int k = lane_id(); k < n_tasks; k += LANES) any_slow = any_slow || c.slot_slow[k] != 0;
        if (w_any(any_slow)) multi_eval_slow(c, n_tasks);
    }
    for (int k = 0; k < n_tasks; ++k) {
        MultiTask& m = c.mt[k];
        double lh_const = lh_shared, prior_const = 0.0;
        if (parent_by_slots) lh_const += c.slot_f[k];
        else if (parent >= 0 && parent != t && sc->samples[parent].contamination_by != t) {
            const int by = sc->samples[parent].contamination_by;
            lh_const += cached_sample_likelihood(c, parent, m.parent_x, by >= 0 ? (by == parent ? m.parent_x : c.ops[od].vaf[by]) : 0.0);
        }
        if (L.uniform) { // flat prior inside every sample's universe (prior.rs:385-406)
            for (int s = 0; s < S; ++s) {
                if (s == t) continue;
                const double v = s == parent ? m.parent_x : c.ops[od].vaf[s];
                if ((sc->samples[s].ploidy == 0 && v != 0.0) || !universe_contains(c, s, v)) prior_const = neg_inf();
            }
        }
        m.lh_const = lh_const;
        m.prior_const = prior_const;
        m.event_slot = event_slots ? event_slots[k] : c.cur_slot;
        m.st.init(a, b, res);
        m.n = 0;
        m.k = 0;
        m.slot_base = 0;
        m.have_best = false;
        m.best_f = m.best_x = 0.0;
        m.active = true;
        m.overflow = false;
    }
    L.dep_lo = 0;
    L.dep_hi = n_dep;
    warp_sync();
    return true;
}

// prior of (task, x) when it is not a loop constant (non-uniform priors, universes with holes)
VLR_DEV_NOINLINE double multi_prior_point(Ctx& c_, int od, int task, double x) {
    Ctx& c = warp_ctx(c_);
    if (c.leaf.uniform) {
        if ((c.leaf.t_ploidy0 && x != 0.0) || !universe_contains(c, c.leaf.t, x)) return neg_inf();
        return c.mt[task].prior_const;
    }
    multi_ops(c, od, task, x);
    return prior_compute(c, c.ops[od + 1]);
}

// Per-task bookkeeping runs with LANES = TASKS: lane k owns task k (its state is c.mt[k] in shared memory), so the
// adaptive state machines of all tasks advance in one pass instead of one after the other.
#ifdef VLR_HOST_EMU
#define VLR_FOR_EACH_TASK(k, T) for (int k = 0; k < (T); ++k)
#else
#define VLR_FOR_EACH_TASK(k, T) for (int k = lane_id(); k < (T); k += LANES)
#endif

// prior per slot when it is not a loop constant: added into c.slot_f (uniform code: it uses the operand stack)
VLR_DEV_NOINLINE void multi_prior_slots(Ctx& c_, int od, int total) {
    Ctx& c = warp_ctx(c_);
    for (int slot = 0; slot < total; ++slot) c.slot_f[slot] += multi_prior_point(c, od, c.slot_task[slot], c.slot_x[slot]);
    warp_sync();
}

// base-event log for the AFD (the fast-path twin of the recording in joint()); called by the lane that owns the task
VLR_DEV_NOINLINE unsigned multi_record(Ctx& c_, int od, double parent_x, double x, double f, uint32_t at, unsigned disc) {
    Ctx& c = warp_ctx(c_);
    const int S = c.sc->S;
    const int t = c.leaf.t;
    if (at >= (uint32_t)BE_CAP) return VLR_ST_BASE_EVENTS_OVERFLOW;
    double* e = c.be + (int64_t)at * (2 + S);
    e[0] = f;
    e[1] = d_make(0, (int)disc);
    for (int s = 0; s < S; ++s) e[2 + s] = s == t ? x : (s == c.leaf.parent ? parent_x : c.ops[od].vaf[s]);
    return 0u;
}

// per task: MAP bookkeeping of joint() (first maximum over tasks in order), then the trapezoid over its grid
VLR_DEV_NOINLINE void leaf_multi_finish(Ctx& c_, int od, unsigned disc, double* out) {
    Ctx& c = warp_ctx(c_);
    const int T = c.leaf.n_tasks;
    const int S = c.sc->S;
    const int t = c.leaf.t;
    for (int k = 0; k < T; ++k) {
        MultiTask& m = c.mt[k];
        const int slot = m.event_slot;
        if (m.have_best && (!c.map_set[slot] || m.best_f > c.map_joint[slot])) {
            c.map_set[slot] = 1;
            c.map_joint[slot] = m.best_f;
            c.map_cfg[slot] = c.art.id;
            c.map_disc[slot] = disc;
            c.map_seq[slot] = ((uint32_t)slot << 24) | (c.n_base & 0xffffffu);
            for (int s = 0; s < S; ++s)
                c.map_vaf[slot][s] = s == t ? m.best_x : (s == c.leaf.parent ? m.parent_x : c.ops[od].vaf[s]);
        }
        if (m.overflow) c.status |= VLR_ST_GRID_OVERFLOW;
        out[k] = grid_trapezoid(c, c.ws->mgrid_x[k], c.ws->mgrid_f[k], m.n);
    }
}

// Runs the c.leaf.n_tasks prepared leaf integrations to completion; integrals -> out[task].
VLR_DEV_NOINLINE void leaf_multi_run(Ctx& c_, int od, double* out) {
    Ctx& c = warp_ctx(c_);
    const int T = c.leaf.n_tasks;
    const int t = c.leaf.t;
    const bool in_sm = c.leaf.coef_in_sm, per_point = c.leaf.prior_per_point, record = c.leaf.record;
    unsigned disc = c.ops[od].disc_mask & ~(1u << t);
    if (c.leaf.parent >= 0) {
        if (c.leaf.parent_disc) disc |= 1u << c.leaf.parent;
        else disc &= ~(1u << c.leaf.parent);
    }
    for (;;) {
        // ---- every active task asks its adaptive search for the next batch; slots = exclusive scan over tasks
        int my_k = 0;
        VLR_FOR_EACH_TASK(k, T) {
            MultiTask& m = c.mt[k];
            m.k = m.active ? m.st.points(m.xs) : 0;
#ifndef VLR_HOST_EMU
            my_k = m.k;
#endif
        }
        int total;
#ifdef VLR_HOST_EMU
        total = 0;
        for (int k = 0; k < T; ++k) {
            c.mt[k].slot_base = total;
            total += c.mt[k].k;
        }
        (void)my_k;
#else
        {
            int incl = my_k; // inclusive scan over the first 8 lanes (T <= MT = 8)
#pragma unroll
            for (int o = 1; o < MT; o <<= 1) {
                int v = __shfl_up_sync(FULL, incl, o, LANES);
                if (lane_id() >= o) incl += v;
            }
            total = __shfl_sync(FULL, incl, MT - 1, LANES);
            if (lane_id() < T) c.mt[lane_id()].slot_base = incl - my_k;
        }
#endif
        if (total == 0) break;
        VLR_FOR_EACH_TASK(k, T) {
            const MultiTask& m = c.mt[k];
            for (int i = 0; i < m.k; ++i) {
                c.slot_x[m.slot_base + i] = m.xs[i];
                c.slot_task[m.slot_base + i] = k;
            }
        }
        warp_sync();
        // ---- all slots in parallel, 32 at a time
        for (int base = 0; base < total; base += LANES) {
            const int m = total - base < LANES ? total - base : LANES;
            if (in_sm) multi_eval_sm(c, base, m);
            else multi_eval_gl(c, base, m);
        }
        {
            bool any_slow = false;
            for (int sidx = lane_id(); sidx < total; sidx += LANES) any_slow = any_slow || c.slot_slow[sidx] != 0;
            if (w_any(any_slow)) multi_eval_slow(c, total);
        }
        if (per_point) multi_prior_slots(c, od, total);
        // ---- consume: lane k takes the values of task k in visit order
        const uint32_t rec0 = c.n_rec;
        unsigned flags = 0;
        VLR_FOR_EACH_TASK(k, T) {
            MultiTask& m = c.mt[k];
            if (m.active) {
                double* gx = c.ws->mgrid_x[k];
                double* gf = c.ws->mgrid_f[k];
                const double base_prior = per_point ? 0.0 : m.prior_const;
                for (int i = 0; i < m.k; ++i) {
                    const double x = m.xs[i];
                    const double f = base_prior + (m.lh_const + c.slot_f[m.slot_base + i]);
                    m.fs[i] = f;
                    if (f != f) flags |= VLR_ST_NAN;
                    if (record) flags |= multi_record(c, od, m.parent_x, x, f, rec0 + (uint32_t)(m.slot_base + i), disc);
                    if (!m.have_best || f > m.best_f) { // first maximum in visit order
                        m.have_best = true;
                        m.best_f = f;
                        m.best_x = x;
                    }
                    if (m.n < GRID_CAP) {
                        gx[m.n] = x;
                        gf[m.n] = f;
                        m.n++;
                    } else {
                        m.overflow = true;
                    }
                }
                m.active = m.st.consume(m.xs, m.fs, m.overflow);
            }
        }
        flags = w_or_u(flags);
        if (flags) c.status |= flags;
        c.n_base += (uint32_t)total;
        if (record) {
            const uint32_t nr = rec0 + (uint32_t)total;
            c.n_rec = nr < (uint32_t)BE_CAP ? nr : (uint32_t)BE_CAP;
        }
        warp_sync();
    }
    leaf_multi_finish(c, od, disc, out);
}

VLR_DEV_NOINLINE double integrate_adaptive_leaf(Ctx& c_, const vlr_node_t& node, int od, double a, double b, double res,
                                                int level) {
    Ctx& c = warp_ctx(c_);
    if (!leaf_setup(c, node, od, a, b, res, -1, 1, nullptr)) return integrate_adaptive_generic(c, node, od, a, b, res, level);
    double* out = c.fs[level]; // scratch of this level
    leaf_multi_run(c, od, out);
    return out[0];
}

// The batch of an enclosing integration, when every point leads straight into a fast-path leaf integration of the
// single child: decides as density() would for the child (generic.rs:331-395) and runs the k integrations together.
// Returns false if the batch has to go through subdensity() point by point.
VLR_DEV_NOINLINE bool try_child_batch(Ctx& c_, const vlr_node_t& node, int od, int k, const double* xs, double* fs) {
    Ctx& c = warp_ctx(c_);
    const DevScenario* sc = c.sc;
    if (node.n_children != 1) return false;
    const vlr_node_t& child = sc->nodes[node.first_child];
    if (child.kind != VLR_NODE_RANGE || child.n_children != 0) return false;
    if (c.ops[od].lfc_mask != 0 || sc->n_lfc_nodes != 0) return false;
    const int cs = child.sample;
    if (cs == node.sample) return false;
    const int n_obs = c.n_obs[cs];
    Range vafs{child.start, child.end, child.left_exclusive != 0, child.right_exclusive != 0};
    if (range_is_empty(vafs) || range_is_singleton(vafs)) return false;
    if (c.clear_ref[cs] && vafs.start > 0.0) return false;
    const double res = sc->samples[cs].resolution;
    const double min_vaf = range_observable_min(vafs, n_obs), max_vaf = range_observable_max(vafs, n_obs);
    if (!(min_vaf <= max_vaf) || (max_vaf - min_vaf) < res || n_obs < 5) return false;
    mark_bound_points(c, vafs, min_vaf, max_vaf);
    // operands of the children: the parent's event is pushed per task (continuous => not discrete)
    c.ops[od + 1] = c.ops[od];
    ops_push(c.ops[od + 1], node.sample, xs[0], false);
    if (!leaf_setup(c, child, od + 1, min_vaf, max_vaf, res, node.sample, k, xs)) return false;
    leaf_multi_run(c, od + 1, fs);
    for (int i = 0; i < k; ++i)
        if (fs[i] != fs[i]) c.status |= VLR_ST_NAN;
    return true;
}

VLR_DEV bool iupac_contains(int mask, int base) {
    int bit = 0;
    switch (base) {
    case 'A': case 'a': bit = 1; break;
    case 'C': case 'c': bit = 2; break;
    case 'G': case 'g': bit = 4; break;
    case 'T': case 't': bit = 8; break;
    }
    return (mask & bit) != 0;
}

// GenericPosterior::density (generic.rs:191-422). c.ops[od] are the operands so far (modified in place where the
// reference passes its &mut on, copied to c.ops[od + 1] where it clones).
VLR_DEV_NOINLINE double density(Ctx& c_, int ni, int od, int level) {
    Ctx& c = warp_ctx(c_);
    const DevScenario* sc = c.sc;
    const vlr_node_t& node = sc->nodes[ni];
    switch (node.kind) {
    case VLR_NODE_LFC: c.ops[od].lfc_mask |= 1u << sc->lfc_ordinal[ni]; return subdensity(c, node, od, level);
    case VLR_NODE_FALSE: return neg_inf();
    case VLR_NODE_TRUE: return 0.0;
    case VLR_NODE_VARIANT: {
        if (c.has_snv) {
            bool contains = iupac_contains(node.refmask, c.refbase) && iupac_contains(node.altmask, c.altbase);
            if ((node.variant_positive && !contains) || (!node.variant_positive && contains)) return neg_inf();
            return subdensity(c, node, od, level);
        } else if (node.variant_positive) {
            return neg_inf();
        }
        return subdensity(c, node, od, level);
    }
    default: break;
    }
    const int sample = node.sample;
    Range bounds = range_empty();
    bool have_bounds = ops_lfc_bounds(c, c.ops[od], sample, bounds);
    if (have_bounds && range_is_empty(bounds)) return neg_inf();
    const int n_obs = c.n_obs[sample];
    const bool is_clear_ref = c.clear_ref[sample] != 0;
    if (node.kind == VLR_NODE_SET) {
        bool all_pos = true;
        int n_in = 0;
        double only = 0.0;
        for (int i = 0; i < node.n_vafs; ++i) {
            double v = sc->set_vafs[node.vaf_offset + i];
            if (!(v > 0.0)) all_pos = false;
            if (!have_bounds || range_contains(bounds, v)) {
                n_in++;
                only = v;
            }
        }
        if (is_clear_ref && all_pos) return neg_inf();
        if (n_in == 1) {
            ops_push(c.ops[od], sample, only, true);
            return subdensity(c, node, od, level);
        }
        Lse acc;
        acc.init();
        for (int i = 0; i < node.n_vafs; ++i) {
            double v = sc->set_vafs[node.vaf_offset + i];
            if (have_bounds && !range_contains(bounds, v)) continue;
            c.ops[od + 1] = c.ops[od];
            ops_push(c.ops[od + 1], sample, v, true);
            acc.add(subdensity(c, node, od + 1, level));
        }
        return acc.value();
    }
    Range vafs{node.start, node.end, node.left_exclusive != 0, node.right_exclusive != 0};
    if (have_bounds) vafs = range_intersect(vafs, bounds);
    if (range_is_empty(vafs)) return neg_inf();
    if (is_clear_ref && vafs.start > 0.0) return neg_inf();
    if (range_is_singleton(vafs)) {
        ops_push(c.ops[od], sample, vafs.start, true);
        return subdensity(c, node, od, level);
    }
    const double res = sc->samples[sample].resolution;
    const double min_vaf = range_observable_min(vafs, n_obs);
    const double max_vaf = range_observable_max(vafs, n_obs);
    if (!(min_vaf <= max_vaf)) c.status |= VLR_ST_NAN; // assert in the reference
    mark_bound_points(c, vafs, min_vaf, max_vaf);
    if ((max_vaf - min_vaf) < res) return integrate_simpson(c, node, od, min_vaf, max_vaf, 3, level);
    if (n_obs < 5) return integrate_simpson(c, node, od, min_vaf, max_vaf, 11, level);
    if (level >= MAXS) {
        c.status |= VLR_ST_GRID_OVERFLOW;
        return neg_inf();
    }
    if (node.n_children == 0) return integrate_adaptive_leaf(c, node, od, min_vaf, max_vaf, res, level);
    return integrate_adaptive_generic(c, node, od, min_vaf, max_vaf, res, level);
}

// ------------------------------------------------------------------------------------------------ locus driver
// Caller::call_record + sample_infos (calling.rs:720-937) around rust-bio Model::compute.
// vaftree.rs:116-164 (Node::contains); LFC constraints of the base event as a bit mask over LFC-node ordinals
VLR_DEV_NOINLINE bool node_contains(const DevScenario* sc, int ni, const double* vaf, uint32_t& lfcs, int exclude) {
    const vlr_node_t& node = sc->nodes[ni];
    bool contained = true;
    switch (node.kind) {
    case VLR_NODE_SET:
    case VLR_NODE_RANGE: {
        if (exclude == node.sample) return true;
        double v = vaf[node.sample];
        if (node.kind == VLR_NODE_SET) {
            contained = false;
            for (int i = 0; i < node.n_vafs; ++i)
                if (sc->set_vafs[node.vaf_offset + i] == v) contained = true;
        } else {
            Range r{node.start, node.end, node.left_exclusive != 0, node.right_exclusive != 0};
            contained = range_contains(r, v);
        }
        break;
    }
    case VLR_NODE_LFC: {
        bool found = false;
        for (int k = 0; k < sc->n_lfc_nodes; ++k) {
            if (!(lfcs & (1u << k))) continue;
            const vlr_node_t& o = sc->nodes[sc->lfc_nodes[k]];
            if (o.sample == node.sample && o.sample_b == node.sample_b && o.cmp == node.cmp &&
                o.lfc_value == node.lfc_value) {
                found = true;
                lfcs &= ~(1u << k);
            }
        }
        contained = found;
        break;
    }
    case VLR_NODE_FALSE: contained = false; break;
    default: contained = true;
    }
    if (node.n_children == 0) return contained && lfcs == 0;
    if (!contained) return false;
    for (int k = 0; k < node.n_children; ++k) {
        if (node.n_children == 1) {
            if (node_contains(sc, node.first_child + k, vaf, lfcs, exclude)) return true;
        } else {
            uint32_t cl = lfcs;
            if (node_contains(sc, node.first_child + k, vaf, cl, exclude)) return true;
        }
    }
    return false;
}

// Allele frequency distribution per sample (calling.rs:891-928) from the recorded artifact-free base events:
// those compatible with the best event (ignoring the sample's own node) whose other samples equal the MAP.
VLR_DEV_NOINLINE void afd_pass(Ctx& c_, int best_scen, int map_slot, double marginal) {
    Ctx& c = warp_ctx(c_);
    const DevScenario* sc = c.sc;
    const DevResults* res = c.res;
    const int S = sc->S;
    if (map_slot < 0 || c.map_cfg[map_slot] != 0) return; // artifact MAP: vaf_dist = None
    const vlr_event_t& ev = sc->events[best_scen];
    const int n_rec = (int)c.n_rec;
    const int stride = 2 + S;
    double* tx = c.ws->afd_x;
    double* tp = c.ws->afd_p;
    warp_sync();
    for (int s = 0; s < S; ++s) {
        int cnt = 0;
        bool trunc = false;
        for (int k0 = 0; k0 < n_rec; k0 += LANES) {
            int k = k0 + lane_id();
            bool ok = false;
            double x = 0.0, p = 0.0;
            if (k < n_rec) {
                const double* e = c.be + (int64_t)k * stride;
                uint32_t disc = (uint32_t)d_lo(e[1]), lfcs = (uint32_t)d_hi(e[1]);
                ok = true;
                for (int t = 0; t < S; ++t) {
                    if (t == s) continue;
                    bool d1 = (disc >> t) & 1u, d2 = (c.map_disc[map_slot] >> t) & 1u;
                    if (!(e[2 + t] == c.map_vaf[map_slot][t] && d1 == d2)) ok = false;
                }
                if (ok) {
                    bool contained = false;
                    for (int r = 0; r < ev.n_roots && !contained; ++r) {
                        uint32_t l = lfcs;
                        contained = node_contains(sc, ev.first_root + r, e + 2, l, s);
                    }
                    ok = contained;
                }
                x = e[2 + s];
                p = e[0] - marginal;
            }
#ifdef VLR_HOST_EMU
            int pos = cnt, tot = ok ? 1 : 0;
#else
            unsigned m = w_ballot(ok);
            int pos = cnt + __popc(m & ((1u << lane_id()) - 1u)), tot = __popc(m);
#endif
            if (ok) {
                if (pos < AFD_TMP) {
                    tx[pos] = x;
                    tp[pos] = p;
                }
            }
            cnt += tot;
        }
        if (cnt > AFD_TMP) {
            cnt = AFD_TMP;
            trunc = true;
        }
        warp_sync();
        // unique by vaf (a later, i.e. lower-posterior, duplicate overwrites in the reference), ascending
        const int cap = res->afd_capacity;
        int n_unique = 0;
        for (int i0 = 0; i0 < cnt; i0 += LANES) {
            int i = i0 + lane_id();
            bool first = false;
            int rank = 0;
            double xi = 0.0, pi = 0.0;
            if (i < cnt) {
                xi = tx[i];
                pi = tp[i];
                first = true;
                for (int j = 0; j < cnt; ++j) {
                    double xj = tx[j];
                    if (xj == xi) {
                        if (j < i) first = false;
                        double pj = tp[j];
                        if (pj < pi) pi = pj;
                    }
                }
                if (first) { // rank among first occurrences
                    for (int j = 0; j < cnt; ++j) {
                        double xj = tx[j];
                        if (xj < xi) {
                            bool jfirst = true;
                            for (int q = 0; q < j; ++q)
                                if (tx[q] == xj) {
                                    jfirst = false;
                                    break;
                                }
                            rank += jfirst;
                        }
                    }
                }
            }
            if (first) {
                if (rank < cap) {
                    int64_t o = ((int64_t)c.locus * S + s) * cap + rank;
                    res->afd_vaf[o] = xi;
                    res->afd_logp[o] = pi;
                } else {
                    trunc = true;
                }
            }
            n_unique += w_sum_i(first ? 1 : 0);
        }
        if (w_any(trunc)) c.status |= VLR_ST_AFD_TRUNCATED;
        if (lane_id() == 0) res->afd_count[(int64_t)c.locus * S + s] = n_unique < cap ? n_unique : cap;
        warp_sync();
    }
}

// Events of the shape  parent:{one VAF} -> leaf:Range  (tumor-normal: somatic_tumor, germline_het, germline_hom) are
// independent leaf integrations that differ only in the parent's VAF: they run as ONE concurrent group instead of
// one after the other. Mirrors, per event, exactly what density() does for the Set root (generic.rs:294-317: clear-ref
// shortcut, single-VAF push as a discrete event) and for the Range child (:331-395). `done` marks the events handled
// here; the caller walks the remaining ones through density().
VLR_DEV_NOINLINE void run_grouped_chain_events(Ctx& c_, bool twin, double ln_event_prior, unsigned& done) {
    Ctx& c = warp_ctx(c_);
    const DevScenario* sc = c.sc;
    const int E = sc->E;
    done = 0;
    if (sc->n_lfc_nodes != 0) return;
    int ev_idx[MT];
    int slots[MT];
    double* xs = c.xs[0];
    int n = 0, parent = -1, child_ni = -1;
    for (int e = 0; e < E && n < MT - 1; ++e) {
        const vlr_event_t& ev = sc->events[e];
        if (twin && !ev.has_artifact_twin) continue;
        if (ev.n_roots != 1) continue;
        const vlr_node_t& root = sc->nodes[ev.first_root];
        if (root.kind != VLR_NODE_SET || root.n_vafs != 1 || root.n_children != 1) continue;
        const vlr_node_t& child = sc->nodes[root.first_child];
        if (child.kind != VLR_NODE_RANGE || child.n_children != 0 || child.sample == root.sample) continue;
        if (n > 0) { // same parent sample and the same leaf integration as the group's first event
            const vlr_node_t& c0 = sc->nodes[child_ni];
            if (root.sample != parent || child.sample != c0.sample || child.start != c0.start || child.end != c0.end ||
                child.left_exclusive != c0.left_exclusive || child.right_exclusive != c0.right_exclusive)
                continue;
        } else {
            parent = root.sample;
            child_ni = root.first_child;
        }
        ev_idx[n] = e;
        n++;
    }
    if (n < 2) return;
    // the leaf integration as density() would set it up for the Range child (identical for every member)
    const vlr_node_t& child = sc->nodes[child_ni];
    const int cs = child.sample;
    const int n_obs = c.n_obs[cs];
    Range vafs{child.start, child.end, child.left_exclusive != 0, child.right_exclusive != 0};
    if (range_is_empty(vafs) || range_is_singleton(vafs)) return;
    if (c.clear_ref[cs] && vafs.start > 0.0) return;
    const double res = sc->samples[cs].resolution;
    const double min_vaf = range_observable_min(vafs, n_obs), max_vaf = range_observable_max(vafs, n_obs);
    if (!(min_vaf <= max_vaf) || (max_vaf - min_vaf) < res || n_obs < 5) return;
    mark_bound_points(c, vafs, min_vaf, max_vaf);
    if (c.map_global || c.be_filter) return; // no fast path (leaf_setup): the events go through density()
    // members cut by the Set node's clear-ref shortcut contribute ln 0 without any evaluation
    int m = 0;
    int live[MT];
    for (int i = 0; i < n; ++i) {
        const int e = ev_idx[i];
        const double v = sc->set_vafs[sc->nodes[sc->events[e].first_root].vaf_offset];
        done |= 1u << e;
        if (c.clear_ref[parent] && v > 0.0) {
            if (twin) c.ev_twin[e].add(ln_event_prior + neg_inf());
            else c.ev_plain[e].add(ln_event_prior + neg_inf());
            continue;
        }
        live[m] = e;
        xs[m] = v;
        slots[m] = 2 * e + (twin ? 1 : 0);
        m++;
    }
    if (m == 0) return;
    Ops& base = c.ops[0];
    for (int s = 0; s < MAXS; ++s) base.vaf[s] = 0.0;
    base.set_mask = base.disc_mask = base.lfc_mask = 0;
    c.ops[1] = base;
    ops_push(c.ops[1], parent, xs[0], true);
    c.cur_slot = slots[0];
    if (!leaf_setup(c, child, 1, min_vaf, max_vaf, res, parent, m, xs, slots, true)) {
        for (int i = 0; i < m; ++i) done &= ~(1u << live[i]); // generic path for the live members
        return;
    }
    double* out = c.fs[0];
    leaf_multi_run(c, 1, out);
    for (int i = 0; i < m; ++i) {
        const double d = out[i];
        if (d != d) c.status |= VLR_ST_NAN;
        if (twin) c.ev_twin[live[i]].add(ln_event_prior + d);
        else c.ev_plain[live[i]].add(ln_event_prior + d);
    }
}

// End of a locus (calling.rs:760-937 after the joint probabilities are known): marginal, posteriors, artifact
// probability, best event, MAP and AFD from the per-event accumulators and MAP slots in the Ctx. Shared by the
// generic warp-per-locus engine and the wavefront pipeline (engine_wave.cuh).
// Returns the MAP slot (or -1). with_afd = false: everything but the allele frequency distributions (first pass of a locus
// whose base-event log overflowed).
VLR_DEV_NOINLINE int locus_tail(Ctx& c_, int n_twins, bool with_afd = true) {
    Ctx& c = warp_ctx(c_);
    const DevScenario* sc = c.sc;
    const DevResults* res = c.res;
    const int S = sc->S, E = sc->E;
    const int64_t locus = c.locus;
    Lse* ev_plain = c.ev_plain;
    Lse* ev_twin = c.ev_twin;
    // marginal over the event universe, in universe order [e plain, e twin]...
    double* joint_u = c.joint_u;
    int* scen_u = c.scen_u;
    uint8_t* art_u = c.art_u;
    int nu = 0;
    Lse marg;
    marg.init();
    for (int e = 0; e < E; ++e) {
        joint_u[nu] = ev_plain[e].value();
        scen_u[nu] = e;
        art_u[nu] = 0;
        marg.add(joint_u[nu]);
        nu++;
        if (sc->events[e].has_artifact_twin && n_twins > 0) {
            joint_u[nu] = ev_twin[e].value();
            scen_u[nu] = e;
            art_u[nu] = 1;
            marg.add(joint_u[nu]);
            nu++;
        }
    }
    const double marginal = marg.value();
    if (marginal == neg_inf()) c.status |= VLR_ST_MARGINAL_ZERO;
    if (marginal != marginal) c.status |= VLR_ST_NAN;

    int best = 0;
    {
        double bestv = joint_u[0] - marginal;
        for (int i = 1; i < nu; ++i) {
            double v = joint_u[i] - marginal;
            if (v >= bestv) { // itertools minmax_by_key: the last maximum wins (calling.rs:762-769)
                bestv = v;
                best = i;
            }
        }
    }
    Lse art;
    art.init();
    double* lp = res->log_post + locus * (int64_t)(E + 1);
    double* my_lp = c.my_lp;
    for (int i = 0; i < nu; ++i) {
        double post = joint_u[i] - marginal;
        if (art_u[i]) art.add(post);
        else my_lp[scen_u[i]] = post;
    }
    const double prob_artifact = art.value();
    my_lp[E] = prob_artifact;
    bool is_artifact = true;
    for (int e = 0; e < E; ++e)
        if (!(my_lp[e] < prob_artifact)) is_artifact = false;
    if (is_artifact) c.status |= VLR_ST_IS_ARTIFACT;

    // MAP (calling.rs:844-890): events of a valid scenario are disjoint (grammar/mod.rs:238-272 rejects overlaps),
    // so the base events contained in the best event are the ones its own tree produced.
    const int best_scen = scen_u[best];
    int map_slot = -1;
    {
        int sp = 2 * best_scen, st = 2 * best_scen + 1;
        if (c.map_set[sp]) map_slot = sp;
        if (is_artifact && c.map_set[st] &&
            (map_slot < 0 || c.map_joint[st] > c.map_joint[sp] ||
             (c.map_joint[st] == c.map_joint[sp] && c.map_seq[st] < c.map_seq[sp])))
            map_slot = st;
    }
    if (map_slot < 0) c.status |= VLR_ST_NO_MAP;

    if (lane_id() == 0) {
        for (int e = 0; e <= E; ++e) lp[e] = my_lp[e];
        if (res->log_marginal) res->log_marginal[locus] = marginal;
        if (res->best_event) res->best_event[locus] = 2 * best_scen + (art_u[best] ? 1 : 0);
        if (res->n_base_events) res->n_base_events[locus] = c.n_base;
        for (int s = 0; s < S; ++s) {
            double v = NAN;
            if (map_slot >= 0) v = c.map_cfg[map_slot] != 0 ? 0.0 : c.map_vaf[map_slot][s];
            res->map_vaf[locus * S + s] = v;
        }
        if (res->map_config) res->map_config[locus] = map_slot >= 0 ? c.map_cfg[map_slot] : 0;
        if (res->afd_capacity > 0)
            for (int s = 0; s < S; ++s) res->afd_count[locus * S + s] = 0;
    }
    if (res->afd_capacity > 0 && with_afd) afd_pass(c, best_scen, map_slot, marginal);
    if (lane_id() == 0) res->status[locus] = c.status;
    return map_slot;
}

// `coef_sm` (capacity sm_reads) is the warp's shared-memory coefficient arena, `coef` (capacity coef_cap) the global
// one used when the locus has more kept reads than fit in shared memory.
VLR_DEV void process_locus(const DevScenario* sc, const DevBatch* b, const DevResults* res, WarpWs* ws, double* coef,
                           double* coef_sm, int sm_reads, double* be, int coef_cap, int64_t locus, Ctx& c) {
    const int S = sc->S, E = sc->E;
    c.sc = sc;
    c.b = b;
    c.res = res;
    c.ws = ws;
    c.coef = coef;
    c.be = res->afd_capacity > 0 ? be : nullptr;
    c.n_rec = 0;
    c.coef_cap = coef_cap;
    c.locus = locus;
    c.status = 0;
    c.lf = b->lflags[locus];
    c.vartype = (c.lf >> VLR_LF_VARTYPE_SHIFT) & 3;
    c.has_snv = (c.lf & VLR_LF_HAS_SNV) != 0;
    c.refbase = (c.lf >> VLR_LF_REFBASE_SHIFT) & 0xff;
    c.altbase = (c.lf >> VLR_LF_ALTBASE_SHIFT) & 0xff;
    c.het_override = NAN;
    c.semr_override = NAN;
    const double phred_to_ln = -0.23025850929940456; // -ln(10)/10 (PHREDProb -> LogProb)
    if (b->het_phred) {
        float h = b->het_phred[locus];
        if (!(h != h)) c.het_override = (double)h * phred_to_ln;
    }
    if (b->semr_phred) {
        float h = b->semr_phred[locus];
        if (!(h != h)) c.semr_override = (double)h * phred_to_ln;
    }
    c.prior_absent = NAN;
    c.pc_n = 0;
    c.n_base = 0;
    c.n_pileup_evals = 0;
    for (int i = 0; i < 2 * E; ++i) c.map_set[i] = 0;

    BiasPlan plan;
    locus_prepass(c, plan);
    if (c.coef_total > c.coef_cap) { // only reachable through vlr_call_batch_device without a sufficient reserve
        if (lane_id() == 0) {
            for (int e = 0; e <= E; ++e) res->log_post[locus * (int64_t)(E + 1) + e] = NAN;
            for (int s = 0; s < S; ++s) res->map_vaf[locus * S + s] = NAN;
            if (res->log_marginal) res->log_marginal[locus] = NAN;
            if (res->best_event) res->best_event[locus] = 0;
            if (res->map_config) res->map_config[locus] = 0;
            if (res->n_base_events) res->n_base_events[locus] = 0;
            if (res->afd_capacity > 0)
                for (int s = 0; s < S; ++s) res->afd_count[locus * S + s] = 0;
            res->status[locus] = c.status | VLR_ST_WORKSPACE_OVERFLOW | VLR_ST_NO_MAP;
        }
        return;
    }
    c.coef_in_sm = c.coef_total <= sm_reads;
    if (c.coef_in_sm) c.coef = coef_sm;

    // joint probability per universe event: plain events get ln 0.5, twins ln 0.5 + ln(1/#configs) (generic.rs:437-441)
    Lse* ev_plain = c.ev_plain;
    Lse* ev_twin = c.ev_twin;
    const double twin_prior = plan.n_twins > 0 ? LN_05 + m_log(1.0 / (double)plan.n_twins) : neg_inf();
    const uint32_t prepass_status = c.status;
    c.map_global = sc->events_overlap;
    c.be_filter = 0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int e = 0; e < E; ++e) {
            ev_plain[e].init();
            ev_twin[e].init();
        }
        for (int ci = 0; ci <= plan.n_surviving; ++ci) {
            c.art.id = ci == 0 ? 0 : plan.surviving[ci - 1];
            c.art.forward_rate = plan.forward_rate;
            c.art.ln_fwd = plan.ln_fwd;
            c.art.ln_rev = plan.ln_rev;
            c.art.has_alt_loci = plan.has_alt_loci;
            for (int s = 0; s < S; ++s) {
                c.lc_n[s] = 0;
                read_coefficients(c, s);
            }
            unsigned grouped = 0;
            run_grouped_chain_events(c, ci > 0, ci == 0 ? LN_05 : twin_prior, grouped);
            for (int e = 0; e < E; ++e) {
                const vlr_event_t& ev = sc->events[e];
                if (ci > 0 && !ev.has_artifact_twin) continue;
                if (grouped & (1u << e)) continue;
                c.cur_slot = 2 * e + (ci > 0 ? 1 : 0);
                for (int r = 0; r < ev.n_roots; ++r) {
                    Ops& ops = c.ops[0];
                    for (int s = 0; s < MAXS; ++s) ops.vaf[s] = 0.0;
                    ops.set_mask = ops.disc_mask = ops.lfc_mask = 0;
                    double d = density(c, ev.first_root + r, 0, 0);
                    if (ci == 0) ev_plain[e].add(LN_05 + d);
                    else ev_twin[e].add(twin_prior + d);
                }
            }
        }
        // The base-event log holds BE_CAP joint evaluations; scenarios that nest full ranges (the reference's
        // tumor-relapse priors) record more. The reference keeps them all (rust-bio Model::compute), but an allele
        // frequency distribution only uses those that equal the MAP in all other samples (calling.rs:891-928): take
        // the MAP from this pass and walk the trees once more, logging only base events within one sample of the MAP.
        const bool overflow = c.be != nullptr && (c.status & VLR_ST_BASE_EVENTS_OVERFLOW) != 0;
        if (pass == 1 || !overflow) {
            locus_tail(c, plan.n_twins);
            break;
        }
        const int map_slot = locus_tail(c, plan.n_twins, false);
        if (map_slot < 0 || c.map_cfg[map_slot] != 0) { // no MAP or an artifact MAP: no distribution to report
            if (lane_id() == 0) res->status[locus] = c.status & ~(uint32_t)VLR_ST_BASE_EVENTS_OVERFLOW;
            break;
        }
        for (int s = 0; s < S; ++s) c.flt_vaf[s] = c.map_vaf[map_slot][s];
        c.flt_disc = c.map_disc[map_slot];
        c.be_filter = 1;
        c.n_rec = 0;
        c.status = prepass_status;
        c.n_base = 0;
        c.n_pileup_evals = 0;
        c.pc_n = 0;
        c.prior_absent = NAN;
        for (int i = 0; i < 2 * E; ++i) c.map_set[i] = 0;
        warp_sync();
    }
}


#ifdef VLR_VAR_WAVE
#include "engine_wave.cuh"
#include "engine_sets.cuh"
#endif

} // namespace VLR_VARIANT
#undef VLR_VAR_WAVE
#undef VLR_VARIANT
#undef VLR_VAR_MAXS
#undef VLR_VAR_MAXE
#undef VLR_VAR_MAXD
