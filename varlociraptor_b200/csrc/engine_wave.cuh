// engine_wave.cuh — the wavefront pipeline: the same per-locus model as engine_core.cuh, restructured so that the
// hot work (adaptive leaf integrations over one sample's allele frequency) runs ONE THREAD PER INTEGRATION instead of
// one warp per locus.
//
// Why: the warp-per-locus engine executes ~300 warp-uniform bookkeeping instructions per joint evaluation and is
// instruction-fetch bound (profiles/README.md: stall_no_instruction ~13 cycles per issued instruction). Here the
// bookkeeping of an integration is private to a thread, so one instruction stream advances 32 integrations, and the
// inner loop (per read: 2 FMAs + 1 MUL per abscissa, three abscissae per thread) is a few dozen instructions that live in
// the L0 instruction cache and keep the fp64 pipe busy.
//
// Shape served (WavePlan, engine_types.cuh): two samples, every event a chain root(P) -> leaf(T) — tumor-normal
// (cli.rs:1151-1173), which is what configs 1, 2, 4 and 5 of BASELINE.json run. Everything else, and every locus that
// needs a branch of generic.rs the pipeline does not implement (Simpson fallbacks for tiny pileups, prob_sample_alt > 0,
// workspace overflow), is DEFERRED to the generic engine (vlr_call_kernel over a locus list): same results, slower.
//
// Pipeline per sub-chunk of loci (kernels in vlr_engine.cu, all on one stream):
//   prep    warp per locus : pre-pass (bias plan), per artifact config ("lc" = locus-config) the per-read coefficients
//                           into a global arena, point events evaluated in place, round-0 tasks emitted
//   round r CTA per group of lcs : stage the leaf sample's coefficients in shared memory, thread per task runs its
//                           adaptive integration to completion, then one thread per lc advances the lc's enclosing
//                           (outer) integration and emits the tasks of round r + 1
//   finish  warp per locus : combine the lcs of a locus (event posteriors, artifact, MAP, AFD: locus_tail)
//
// What the reference does where (file:line under /root/reference/src) is cited at the corresponding code below.
// This header is included at the end of engine_core.cuh, inside the variant namespace, when VLR_VAR_WAVE is defined.

constexpr int W_OGRID = 96;      // points of an outer (root Range) integration grid
constexpr int W_GCAP = GRID_CAP; // points of a leaf integration grid (same capacity as the generic engine)
constexpr int W_MAXT = 8;        // tasks an lc can have in one round
constexpr int W_MAXROUNDS = 64;
#ifndef VLR_WAVE_GROUP
#define VLR_WAVE_GROUP 16 // measured: 4 CTAs x 16 lcs per SM overlap their stage / task / advance phases best (vs 2 x 32)
#endif
constexpr int W_GROUP = VLR_WAVE_GROUP; // lcs per CTA group (= shared-memory coefficient slots), <= 32
constexpr int W_SLOT_READS = 104; // reads per shared-memory coefficient slot (3.3 KB); deeper pileups are read from L2
constexpr int W_SLOT_STRIDE = W_SLOT_READS * 4 + 2; // doubles between slots: 16 bytes of skew against bank conflicts

#ifdef VLR_HOST_EMU
VLR_DEV unsigned wa_add_u32(unsigned* p, unsigned v) {
    unsigned o = *p;
    *p += v;
    return o;
}
VLR_DEV unsigned long long wa_add_u64(unsigned long long* p, unsigned long long v) {
    unsigned long long o = *p;
    *p += v;
    return o;
}
#else
VLR_DEV unsigned wa_add_u32(unsigned* p, unsigned v) { return atomicAdd(p, v); }
VLR_DEV unsigned long long wa_add_u64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
#endif

constexpr int W_RCLASSES = 4; // size classes of the lc-resident round kernels (engine_resident.cuh: r_class)

struct WaveCounters {
    unsigned long long ticket[8]; // pre, finish, deferred (generic kernel), coefficients, resident rounds (4 + class - 1)
    unsigned long long coef_used; // doubles allocated in the coefficient arena
    unsigned int n_lc, n_deferred;
    // lcs served by the lc-resident round kernels (engine_resident.cuh), ordered so that the groups of a warp work on
    // lcs of the same shape and cost and stay in step (a warp's round lasts as long as its longest task): key = those
    // with an outer integration (five rounds of 5, 3, 3, 3, 7 tasks) first, the single-round ones behind them, and within
    // each by the shares of reads in the two pileups that do not favour the reference (8 x 8 bins: where the likelihood
    // peaks decides how many steps the adaptive searches take). Measured on config 2: +17 % loci/s against lc order,
    // as good as sorting by the number of joint evaluations known afterwards (scripts/exp_sorted.py).
    // One list per size class (engine_resident.cuh: r_class), laid out by KEY: the pre-pass counts the lcs of every key
    // (rkey_n), the lc-init kernel places them behind the exclusive prefix sums through the cursors (rkey_cur).
    unsigned int rlist_total[W_RCLASSES];
    unsigned int rkey_n[W_RCLASSES][128], rkey_cur[W_RCLASSES][128];
    unsigned int list_n[W_MAXROUNDS + 2];  // lcs of the round whose pileups fit a coefficient slot
    unsigned int dlist_n[W_MAXROUNDS + 2]; // lcs of the round with a deeper pileup
    unsigned int task_n[W_MAXROUNDS + 2];
};

struct WaveLocus {
    int lc_base, n_cfg; // n_cfg = 0: deferred to the generic engine
    int n_twins;
    uint32_t status;
    // what the per-lc kernels need from the pre-pass
    uint32_t lf;
    int resident;   // size class (r_class, 1..4) when the locus' lcs are served by the lc-resident round kernel
                    // (polynomial arena format), else 0
    int cost_bin;   // 8 * bin(leaf pileup) + bin(parent pileup) of the shares of reads not favouring the reference
    int lc_doubles; // arena doubles per lc
    int coef_total, has_alt_loci;
    int n_obs[2], s_one[2], coef_off[2];
    int surviving[NCFG];
    int64_t coef_base, singleton_row;
    double forward_rate, ln_fwd, ln_rev;
    double pa, pb;                    // limits of the outer (root Range) integration
    double ev_a[MAXE], ev_b[MAXE];    // limits of the leaf integration per event
    int8_t ev_kind[MAXE];             // 0 pruned, 1 leaf task under a discrete parent, 2 point, 3 outer
};

struct WaveLC { // one (locus, artifact config)
    int li;     // locus index inside the sub-chunk
    int ci, art_id;
    int nP, nT;
    int m0P, m0T; // every kept read of the sample has prob_sample_alt == 0
    int resident;               // size class (1..4): the arena holds pileup polynomials (engine_resident.cuh), else 0: per-read coefficients
    int nqPx, nqPy, nqTx, nqTy; // ... this many per group, parent pileup at coefP, leaf pileup right behind it
    int task_base, task_count;
    uint32_t status, n_base;
    int outer_pending, outer_n, outer_overflow;
    int64_t coefP, coefT; // first double of the sample's coefficients in the arena
    double ksumP, ksumT;
    double ta, tb; // leaf integration limits under the outer event
    double dens[MAXE];
    double map_joint[MAXE], map_vp[MAXE], map_vt[MAXE];
    uint8_t map_set[MAXE], map_disc[MAXE]; // disc: bit 0 parent event discrete, bit 1 leaf event discrete
    Adaptive outer;
    double outer_xs[8];
    // lc-resident rounds: m1 and m2 of the outer integration's first iteration (abscissa, integral, joint evaluations).
    // The closing batch revisits one of them (the abandoned arm's midpoint, adaptive_integration.rs:96-106, is bitwise
    // that m1 or m2): the reference evaluates it again and gets the same number, here it is remembered.
    double om_x[2], om_val[2];
    uint32_t om_nev[2];
    int outer_skip; // the next round's first outer abscissa is om_x[outer_skip - 1] and has no task
};

struct WaveTask {
    int lc;
    short event;
    uint8_t parent_disc, pad;
    double parent_x, a, b; // in
    double value, best_f, best_x; // out
    uint32_t n_evals, status;
    int n_grid, pad2;
};

struct WaveBufs {
    WaveCounters* cnt;
    WaveLocus* loci;
    WaveLC* lcs;
    double* og_x; // [lc_cap][W_OGRID]
    double* og_f;
    double* coef; // arena: 4 doubles per read [alpha, beta, gamma, u], or 6 per polynomial for resident lcs
    WaveTask* tasks[2];
    int* list[2];
    int* dlist[2];
    int* deferred; // absolute locus indices
    double* gx;    // per-thread leaf grid rows: [threads][W_GCAP]
    double* gf;
    double* be;      // base-event log per locus of the sub-chunk (AFD only): [n_sub][BE_CAP][2 + S]
    unsigned* be_n;  // [n_sub]
    int64_t coef_cap; // doubles
    int lc_cap;
    // lc-resident round kernel (engine_resident.cuh)
    int* rlist;        // its lcs
    double* rgx;       // per octet and task: visited abscissae [W_GCAP] ...
    double* rgm;       // ... mantissas ...
    int* rge;          // ... and binary exponents of the values
    double* rscratch;  // per octet: 3 x W_GCAP doubles (sorting grids longer than the shared-memory scratch)
    double* cscratch;  // per warp of the coefficient kernel: 6 doubles per read x cscratch_reads
    int cscratch_reads; // deepest lc (reads of both samples) the scratch holds: deeper ones are not resident
    int allow_resident;
};

// ---------------------------------------------------------------------------------------------- pileup evaluation
struct WArgs {
    double X1, xu, Yp;
};
// x = rho xp + iota xs with xp = (vaf == 1 ? 1 : vaf s_r), y = 1 - x without cancellation (likelihood.rs:43-53,
// :98-103); per read x_r = xu - u_r X1, y_r = Yp + u_r X1 with u_r = 1 - s_r. Same expressions as multi_eval_impl.
VLR_DEV WArgs wave_args(double rho, double iota, double vaf, double vby) {
    const bool p1 = vaf == 1.0, s1 = vby == 1.0, sec = iota != 0.0;
    WArgs a;
    a.X1 = (p1 ? 0.0 : rho * vaf) + ((sec && !s1) ? iota * vby : 0.0);
    const double X0 = (p1 ? rho : 0.0) + ((sec && s1) ? iota : 0.0);
    a.Yp = (p1 ? 0.0 : rho * (1.0 - vaf)) + ((sec && !s1) ? iota * (1.0 - vby) : 0.0);
    a.xu = a.X1 + X0;
    return a;
}

VLR_DEV void wave_pull(double& acc, int& ex, unsigned& slow, unsigned bit) {
    slow |= (acc >= 1e-240) ? 0u : bit; // a zero, tiny or NaN factor: the point is re-evaluated carefully
    const int hi = d_hi(acc);
    ex += ((hi >> 20) & 0x7ff) - 1023;
    acc = d_make((hi & 0x800fffff) | (1023 << 20), d_lo(acc));
}

// How the lanes of a task split the reads of a pileup: H = 1, 2, 4 or 8 neighbouring lanes (h = position inside the
// group, mask = the group's lanes) each take every H-th block of 4 reads and combine their partial products with an
// xor butterfly (bitwise identical on all H lanes, which keeps their control flow identical). (The text above says
// "block of 4 reads" loosely: lane h takes reads h, h + H, ...; see wave_eval.)
struct WSplit {
    int h, H;
    unsigned mask;
};

// ln-likelihood of one pileup at NP abscissae at once: product over the reads of alpha x + beta y + gamma as mantissa
// + binary exponent, one log per abscissa.
// M0: every read has prob_sample_alt = 0 (u_r = 0), the per-read x/y corrections vanish.
// SM: `co` points into the CTA's dynamic shared memory (LDS instead of generic loads).
template <bool M0, int NP, bool SM>
VLR_DEV void wave_eval(const double2* __restrict__ co, int n, double ksum, const WArgs* a, double* lnl, unsigned& slowmask,
                       const WSplit sp) {
#ifndef VLR_HOST_EMU
    if (SM) co = reinterpret_cast<const double2*>(vlr_smem + (__cvta_generic_to_shared(co) - __cvta_generic_to_shared(vlr_smem)));
#endif
    double acc[NP];
    int ex[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        acc[i] = 1.0;
        ex[i] = 0;
    }
    unsigned slow = 0;
    auto term = [&](int r, int i) -> double {
        const double2 ab = co[2 * r];
        const double2 gu = co[2 * r + 1];
        if (M0) return fma(ab.x, a[i].xu, fma(ab.y, a[i].Yp, gu.x));
        const double xr = fma(-gu.y, a[i].X1, a[i].xu);
        const double yr = fma(gu.y, a[i].X1, a[i].Yp);
        return fma(ab.x, xr, fma(ab.y, yr, gu.x));
    };
    auto step = [&](int r) {
#pragma unroll
        for (int i = 0; i < NP; ++i) acc[i] *= term(r, i);
    };
    // lane h takes the reads h, h + H, h + 2H, ...: neighbouring lanes touch neighbouring 32-byte records, and the
    // slots of neighbouring lcs are skewed by 16 bytes (W_SLOT_STRIDE), so the 128-bit loads of a quarter warp fall
    // into different bank groups (the first version, block-split and unskewed, ran ~7-way bank conflicts).
    // Four reads per block, their 4 x NP terms independent and multiplied as a tree: enough instruction-level
    // parallelism for one warp to keep the fp64 pipe busy (a serial acc *= t chain stalls on the pipe latency).
    const int H = sp.H;
    int r = sp.h;
#ifdef VLR_PULL8
    // Tuning experiment (never set in the product build; A/B with scripts/ab_lib.py): 8 factors between exponent pulls.
    // A pull is an exact power-of-two rescaling, so skipping every other one leaves every bit of the result unchanged
    // (mantissa < 2 times 8 factors <= 3 stays far from overflow; a product below 1e-240 still takes the careful path).
#pragma unroll 1
    for (; r + 7 * H < n; r += 8 * H) {
        double t0[NP], t1[NP], t2[NP], t3[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            t0[i] = term(r, i);
            t1[i] = term(r + H, i);
            t2[i] = term(r + 2 * H, i);
            t3[i] = term(r + 3 * H, i);
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) acc[i] *= (t0[i] * t1[i]) * (t2[i] * t3[i]);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            t0[i] = term(r + 4 * H, i);
            t1[i] = term(r + 5 * H, i);
            t2[i] = term(r + 6 * H, i);
            t3[i] = term(r + 7 * H, i);
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            acc[i] *= (t0[i] * t1[i]) * (t2[i] * t3[i]);
            wave_pull(acc[i], ex[i], slow, 1u << i);
        }
    }
#endif
#pragma unroll 1
    for (; r + 3 * H < n; r += 4 * H) { // 4 factors (each <= 3) between exponent pulls
        double t0[NP], t1[NP], t2[NP], t3[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            t0[i] = term(r, i);
            t1[i] = term(r + H, i);
            t2[i] = term(r + 2 * H, i);
            t3[i] = term(r + 3 * H, i);
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            acc[i] *= (t0[i] * t1[i]) * (t2[i] * t3[i]);
            wave_pull(acc[i], ex[i], slow, 1u << i);
        }
    }
#pragma unroll 1
    for (; r < n; r += H) step(r);
#pragma unroll
    for (int i = 0; i < NP; ++i) wave_pull(acc[i], ex[i], slow, 1u << i);
#ifndef VLR_HOST_EMU
#pragma unroll 1
    for (int o = sp.H >> 1; o > 0; o >>= 1) {
        slow |= __shfl_xor_sync(sp.mask, slow, o);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            acc[i] *= __shfl_xor_sync(sp.mask, acc[i], o);
            ex[i] += __shfl_xor_sync(sp.mask, ex[i], o);
        }
    }
#endif
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        wave_pull(acc[i], ex[i], slow, 1u << i); // <= 8 mantissas in [1,2): < 2^8
        lnl[i] = (log(acc[i]) + (double)ex[i] * LN_2) + ksum; // a NaN ksum (invalid inputs) propagates
    }
    slowmask = slow;
}

// careful evaluation of one abscissa (zero / denormal-range factors): sample_likelihood() for a single thread
VLR_DEV_NOINLINE double wave_eval_careful(const double2* co, int n, double ksum, WArgs a) {
    if (n == 0) return 0.0;
    if (ksum != ksum) return NAN;
    double acc = 1.0;
    int ex = 0, k = 0;
    bool zero = false;
    for (int r = 0; r < n; ++r) {
        const double2 ab = co[2 * r];
        const double2 gu = co[2 * r + 1];
        const double xr = fma(-gu.y, a.X1, a.xu);
        const double yr = fma(gu.y, a.X1, a.Yp);
        double t = fma(ab.x, xr, fma(ab.y, yr, gu.x));
        if (t < 1e-30) {
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
        if (++k == 8) {
            k = 0;
            const int hi = d_hi(acc);
            ex += ((hi >> 20) & 0x7ff) - 1023;
            acc = d_make((hi & 0x800fffff) | (1023 << 20), d_lo(acc));
        }
    }
    const int hi = d_hi(acc);
    ex += ((hi >> 20) & 0x7ff) - 1023;
    acc = d_make((hi & 0x800fffff) | (1023 << 20), d_lo(acc));
    if (zero || ksum == neg_inf()) return neg_inf();
    if (acc != acc) return NAN;
    return (log(acc) + (double)ex * LN_2) + ksum;
}

template <int NP, bool SM>
VLR_DEV void wave_pileup(const double2* co, int n, double ksum, bool m0, const WArgs* a, int nvalid, double* lnl,
                         const WSplit sp) {
    if (n == 0) { // empty fold = ln 1
#pragma unroll
        for (int i = 0; i < NP; ++i) lnl[i] = 0.0;
        return;
    }
    unsigned slow;
    if (m0) wave_eval<true, NP, SM>(co, n, ksum, a, lnl, slow, sp);
    else wave_eval<false, NP, SM>(co, n, ksum, a, lnl, slow, sp);
    if (slow) { // every lane of the task re-evaluates the whole pileup (identical values, no exchange needed)
#pragma unroll
        for (int i = 0; i < NP; ++i)
            if (i < nvalid && ((slow >> i) & 1u)) lnl[i] = wave_eval_careful(co, n, ksum, a[i]);
    }
}

// flat prior of an all-uniform scenario for one sample's VAF (prior.rs:385-406, generic.rs joint())
VLR_DEV bool wave_prior_ok(const DevScenario* sc, int s, double v) {
    return !(sc->samples[s].ploidy == 0 && v != 0.0) && universe_contains_sc(sc, s, v);
}

// ---------------------------------------------------------------------------------------------- leaf task (thread)
// One adaptive integration of the leaf sample's allele frequency over [a, b] with the parent sample fixed at
// parent_x: utils/adaptive_integration.rs:25-141 driven exactly like integrate_adaptive_leaf / leaf_multi_run. The
// whole search state lives in registers; the visited points are appended, in visit order, to the thread's grid row
// (gx, gf) and integrated afterwards by wave_fin_coop.
// The H lanes of a task (WSplit) run these functions in lockstep on identical values; lane h = 0 writes.
//
// Pileup ln-likelihood of the parent sample at the task's parent_x: a constant of the leaf integration
// (GenericLikelihood::compute, generic.rs:511-551, hits its per-sample cache for it at every point).
VLR_DEV double wave_task_parent(const WaveLC& lc, const WaveTask& t, const double2* coP, const bool p_in_sm, const WSplit sp) {
    const WArgs ap = wave_args(1.0, 0.0, t.parent_x, 0.0);
    double lh;
    if (p_in_sm) wave_pileup<1, true>(coP, lc.nP, lc.ksumP, lc.m0P != 0, &ap, 1, &lh, sp);
    else wave_pileup<1, false>(coP, lc.nP, lc.ksumP, lc.m0P != 0, &ap, 1, &lh, sp);
    return lh;
}

VLR_DEV void wave_task_run(const DevScenario* sc, const WavePlan& wp, const WaveLC& lc, WaveTask& t, const double2* coT,
                           const bool t_in_sm, const double lh_const, double* gx, double* gf, const WSplit sp) {
    const int P = wp.P, T = wp.T;
    const vlr_sample_t& smT = sc->samples[T];
    double rhoT = 1.0, iotaT = 0.0;
    if (smT.contamination_by >= 0) {
        rhoT = 1.0 - smT.contamination_fraction; // e^{purity}
        iotaT = 1.0 - rhoT;                      // e^{impurity} (likelihood.rs:77-84)
    }
    const double px = t.parent_x;
    const double vby = smT.contamination_by >= 0 ? px : 0.0;
    const int nT = lc.nT; // (locals: the lc lives in global memory and the grid stores below may alias it)
    const double ksumT = lc.ksumT;
    const bool m0T = lc.m0T != 0;
    uint32_t status = 0;
    const double prior_const = wave_prior_ok(sc, P, px) ? 0.0 : neg_inf();
    const double a = t.a, b = t.b, res = smT.resolution;
    int n = 0;
    uint32_t n_evals = 0;
    bool overflow = false, have_best = false, any_nan = false;
    double best_f = 0.0, best_x = 0.0;

    auto visit = [&](double x, double f) {
        n_evals++;
        if (f != f) any_nan = true;
        if (!have_best || f > best_f) { // first maximum in visit order (calling.rs:851-870 via joint())
            have_best = true;
            best_f = f;
            best_x = x;
        }
        if (n >= W_GCAP) {
            overflow = true;
            return;
        }
        if (sp.h == 0) {
            gx[n] = x;
            gf[n] = f;
        }
        n++;
    };

    // One loop, one evaluation site (small code: every lane of the warp, whatever its phase, meets in the same pileup
    // loop). Steps: 0 [min, max]; 1 [middle, m1, m2] while the bracket is wider than the resolution; 2, 3 the 3 + 3
    // points around the optimum (adaptive_integration.rs:108-131). The midpoint of the arm abandoned by the first
    // iteration (:96-106) is bitwise that iteration's m1 or m2 — the reference's HashMap deduplicates it — so its
    // value is remembered instead of evaluated again (it still counts as a visit, like in the generic engine).
    double left = a, right = b, f_left = 0.0, f_right = 0.0, middle = 0.0, first_middle = 0.0;
    double f_first_m1 = 0.0, f_first_m2 = 0.0;
    double x4 = 0.0, x5 = 0.0, x6 = 0.0;
    bool have_middle = false;
    int step = 0;
    while (step < 4) {
        double x0, x1, x2;
        int nvalid = 3;
        if (step == 0) {
            x0 = a;
            x1 = x2 = b;
            nvalid = 2;
        } else if (step == 1) {
            middle = (right + left) / 2.0;
            x0 = middle;
            x1 = (middle + left) / 2.0;
            x2 = (right + middle) / 2.0;
        } else if (step == 2) {
            const bool upper = middle < first_middle;
            visit(upper ? (b + first_middle) / 2.0 : (first_middle + a) / 2.0, upper ? f_first_m2 : f_first_m1);
            const double lo = fmax(middle - (res * 3.0), a);
            const double slo = (middle - lo) / 3.0; // itertools-num linspace(lo, middle, 4).take(3)
            x0 = lo + slo * 0.0;
            x1 = lo + slo * 1.0;
            x2 = lo + slo * 2.0;
            const double hi = fmin(middle + (res * 3.0), b);
            const double shi = (hi - middle) / 3.0; // linspace(middle, hi, 4).skip(1)
            x4 = middle + shi * 1.0;
            x5 = middle + shi * 2.0;
            x6 = middle + shi * 3.0;
        } else {
            x0 = x4;
            x1 = x5;
            x2 = x6;
        }
        WArgs w[3];
        w[0] = wave_args(rhoT, iotaT, x0, vby);
        w[1] = wave_args(rhoT, iotaT, x1, vby);
        w[2] = wave_args(rhoT, iotaT, x2, vby);
        double lnl[3];
        if (t_in_sm) wave_pileup<3, true>(coT, nT, ksumT, m0T, w, nvalid, lnl, sp);
        else wave_pileup<3, false>(coT, nT, ksumT, m0T, w, nvalid, lnl, sp);
        const double f0 = prior_const + (lh_const + lnl[0]), f1 = prior_const + (lh_const + lnl[1]),
                     f2 = prior_const + (lh_const + lnl[2]);
        visit(x0, f0);
        visit(x1, f1);
        if (nvalid > 2) visit(x2, f2);
        if (step == 0) {
            f_left = f0;
            f_right = f1;
        } else if (step == 1) {
            if (!have_middle) {
                first_middle = middle;
                f_first_m1 = f1;
                f_first_m2 = f2;
            }
            have_middle = true;
            const double m1 = x1, m2 = x2, f_m1 = f1, f_m2 = f2;
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
        if (step <= 1) {
            // while (((right - left) >= res && left < right) || middle.is_none())   (adaptive_integration.rs:52)
            step = (!overflow && ((((right - left) >= res) && left < right) || !have_middle)) ? 1 : 2;
        } else {
            step++;
        }
    }
    if (any_nan) status |= VLR_ST_NAN;
    if (overflow) status |= VLR_ST_GRID_OVERFLOW;
    if (sp.h != 0) return;
    t.value = neg_inf(); // set by wave_fin_coop
    t.best_f = best_f;
    t.best_x = best_x;
    t.n_evals = n_evals;
    t.n_grid = n;
    t.status = status;
}

// A cooperating lane group (the whole warp, or an 8-lane quarter of it so that one warp closes four lcs at a time
// and their global-memory latencies overlap; a single lane in the host emulation).
struct WGroup {
    int lane, n;
    unsigned mask;
};
#ifdef VLR_HOST_EMU
VLR_DEV void grp_sync(const WGroup&) {}
VLR_DEV double grp_sum_d(double v, const WGroup&) { return v; }
VLR_DEV unsigned grp_bcast_u(unsigned v, const WGroup&) { return v; }
#else
VLR_DEV void grp_sync(const WGroup& g) { __syncwarp(g.mask); }
VLR_DEV double grp_sum_d(double v, const WGroup& g) { // xor butterfly: identical bits in every lane of the group
    for (int o = g.n >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(g.mask, v, o);
    return v;
}
VLR_DEV unsigned grp_bcast_u(unsigned v, const WGroup& g) { return __shfl_sync(g.mask, v, 0, g.n); }
#endif

// ---------------------------------------------------------------------------------------------- closing trapezoid (lane group)
// ln_trapezoidal_integrate_grid_exp over n visited points in any order (rust-bio): rank sort by (x, visit
// order) into the scratch arrays (shared memory on the device), lanes over points, then lanes over intervals, summed in
// linear space relative to the maximum: ln( sum_i (e^{f_i} + e^{f_i+1}) / 2 * (x_i+1 - x_i) ). scratch: 3 x W_GCAP doubles.
// Equal abscissae keep visit order (zero-width intervals contribute nothing), like grid_trapezoid of the generic engine.
VLR_DEV_NOINLINE double wave_fin_coop(const double* x, const double* f, int n, double fmx, bool any_nan, double* scratch, const WGroup grp) {
    if (any_nan) return NAN;
    if (n < 2 || fmx == neg_inf()) return neg_inf();
    double* ux = scratch;          // abscissae in visit order
    double* ue = scratch + W_GCAP; // e^{f - max} in visit order
    short* inv = reinterpret_cast<short*>(scratch + 2 * W_GCAP); // inv[rank] = visit index
    grp_sync(grp);
    for (int a = grp.lane; a < n; a += grp.n) { // both rows in flight together; the exp is off the sort's critical path
        const double xa = x[a], fa = f[a];
        ux[a] = xa;
        ue[a] = m_exp(fa - fmx);
    }
    grp_sync(grp);
    for (int a = grp.lane; a < n; a += grp.n) {
        const double xi = ux[a];
        int rank = 0;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const double xj = ux[j];
            rank += (xj < xi) || (xj == xi && j < a);
        }
        inv[rank] = (short)a;
    }
    grp_sync(grp);
    double sum = 0.0;
    for (int p = grp.lane; p + 1 < n; p += grp.n) {
        const int i0 = inv[p], i1 = inv[p + 1];
        sum += (ue[i0] + ue[i1]) * (ux[i1] - ux[i0]);
    }
    sum = grp_sum_d(sum, grp);
    grp_sync(grp);
    return fmx + m_log(sum * 0.5);
}

// The round's lcs come in two lists: pileups that fit a shared-memory coefficient slot go to the warp-per-4-lcs round
// kernel, deeper ones to the CTA-per-group kernel (passes, big slot, up to 32 lanes per task).
VLR_DEV void wave_list_append(const WaveBufs& wb, int round, int lci, int nP, int nT) {
    if (nP > W_SLOT_READS || nT > W_SLOT_READS) {
        const unsigned at = wa_add_u32(&wb.cnt->dlist_n[round], 1u);
        wb.dlist[round & 1][at] = lci;
    } else {
        const unsigned at = wa_add_u32(&wb.cnt->list_n[round], 1u);
        wb.list[round & 1][at] = lci;
    }
}

VLR_DEV void wave_emit_task(WaveTask& t, int lc, int event, double px, bool disc, double a, double b) {
    t.lc = lc;
    t.event = (short)event;
    t.parent_disc = disc ? 1 : 0;
    t.pad = 0;
    t.parent_x = px;
    t.a = a;
    t.b = b;
    t.value = neg_inf();
    t.best_f = neg_inf();
    t.best_x = 0.0;
    t.n_evals = 0;
    t.n_grid = 0;
    t.status = 0;
}

// ---------------------------------------------------------------------------------------------- lc advance (warp)
// After the tasks of round `round` of an lc ran (task i of the lc used grid row rows + i * row_stride): integrate each
// task's grid, MAP bookkeeping in visit order (calling.rs:851-870), event densities, base-event log for the AFD, and the
// next step of the enclosing integration (integrate_adaptive_generic over the root Range). Control flow is warp-uniform
// (every lane reads the same records); lane 0 alone writes.
VLR_DEV_NOINLINE void wave_lc_advance(const WavePlan& wp, const WaveBufs& wb, int lci, int round, const double* rows_x,
                                      const double* rows_f, int row_stride, double* scratch, bool want_be, const WGroup grp) {
    WaveLC& lc = wb.lcs[lci];
    WaveTask* tasks = wb.tasks[round & 1] + lc.task_base;
    const int cnt = lc.task_count;
    const int P = wp.P, T = wp.T;
    const bool l0 = grp.lane == 0;
    uint32_t status = lc.status, n_base = lc.n_base;
    double ofs[8];
    int no = 0;
    WaveTask tnext = tasks[0];
    for (int i = 0; i < cnt; ++i) {
        const WaveTask t = tnext;
        if (i + 1 < cnt) tnext = tasks[i + 1]; // in flight while task i's grid is integrated
        const int e = t.event;
        const double* gx = rows_x + (size_t)i * row_stride;
        const double* gf = rows_f + (size_t)i * row_stride;
        const double value = wave_fin_coop(gx, gf, t.n_grid, t.best_f, (t.status & VLR_ST_NAN) != 0, scratch, grp);
        status |= t.status;
        n_base += t.n_evals;
        if (want_be && lc.ci == 0) { // base events of the artifact-free config feed the AFD (calling.rs:891-928)
            unsigned base = 0;
            if (l0) base = wa_add_u32(&wb.be_n[lc.li], (unsigned)t.n_grid);
            base = grp_bcast_u(base, grp);
            double* be = wb.be + (size_t)lc.li * BE_CAP * 4;
            const double disc = d_make(0, (int)((t.parent_disc ? 1u : 0u) << P));
            if (base + (unsigned)t.n_grid > (unsigned)BE_CAP) status |= VLR_ST_BASE_EVENTS_OVERFLOW;
            for (int k = grp.lane; k < t.n_grid; k += grp.n) {
                const unsigned at = base + (unsigned)k;
                if (at >= (unsigned)BE_CAP) break;
                double* r = be + (size_t)at * 4;
                r[0] = gf[k];
                r[1] = disc;
                r[2 + P] = t.parent_x;
                r[2 + T] = gx[k];
            }
        }
        if (l0 && t.n_evals > 0 && (!lc.map_set[e] || t.best_f > lc.map_joint[e])) {
            lc.map_set[e] = 1;
            lc.map_joint[e] = t.best_f;
            lc.map_vp[e] = t.parent_x;
            lc.map_vt[e] = t.best_x;
            lc.map_disc[e] = t.parent_disc ? 1 : 0;
        }
        grp_sync(grp);
        if (e == wp.outer_event) {
            if (no < 8) ofs[no] = value;
            no++;
        } else if (l0) {
            lc.dens[e] = value;
        }
    }
    if (!lc.outer_pending) {
        if (l0) {
            lc.status = status;
            lc.n_base = n_base;
            lc.task_count = 0;
        }
        return;
    }
    double* ox = wb.og_x + (size_t)lci * W_OGRID;
    double* of = wb.og_f + (size_t)lci * W_OGRID;
    int outer_n = lc.outer_n, outer_overflow = lc.outer_overflow;
    for (int i = 0; i < no && i < 8; ++i) {
        if (ofs[i] != ofs[i]) status |= VLR_ST_NAN;
        if (outer_n < W_OGRID) {
            if (l0) {
                ox[outer_n] = lc.outer_xs[i];
                of[outer_n] = ofs[i];
            }
            outer_n++;
        } else {
            outer_overflow = 1;
        }
    }
    grp_sync(grp);
    Adaptive st = lc.outer;
    double oxs[8];
    for (int i = 0; i < 8; ++i) oxs[i] = lc.outer_xs[i];
    const double lta = lc.ta, ltb = lc.tb;
    grp_sync(grp); // every lane has read the lc before lane 0 updates it
    const bool more = st.consume(oxs, ofs, outer_overflow != 0);
    if (more && round + 1 < W_MAXROUNDS) {
        double xs[8];
        const int k = st.points(xs);
        if (l0) {
            lc.outer = st;
            const unsigned tb = wa_add_u32(&wb.cnt->task_n[round + 1], (unsigned)k);
            WaveTask* nt = wb.tasks[(round + 1) & 1] + tb;
            for (int i = 0; i < k; ++i) {
                lc.outer_xs[i] = xs[i];
                wave_emit_task(nt[i], lci, wp.outer_event, xs[i], false, lta, ltb);
            }
            lc.task_base = (int)tb;
            lc.task_count = k;
            wave_list_append(wb, round + 1, lci, lc.nP, lc.nT);
            lc.outer_n = outer_n;
            lc.outer_overflow = outer_overflow;
            lc.status = status;
            lc.n_base = n_base;
        }
    } else {
        if (outer_overflow || more) status |= VLR_ST_GRID_OVERFLOW;
        bool any_nan = false;
        double fmx = neg_inf();
        for (int i = 0; i < outer_n; ++i) {
            const double v = of[i];
            if (v != v) any_nan = true;
            if (v > fmx) fmx = v;
        }
        if (any_nan) status |= VLR_ST_NAN;
        const double d = wave_fin_coop(ox, of, outer_n, fmx, any_nan, scratch, grp);
        if (l0) {
            lc.outer = st;
            lc.outer_pending = 0;
            lc.outer_n = outer_n;
            lc.outer_overflow = outer_overflow;
            lc.dens[wp.outer_event] = d;
            lc.status = status;
            lc.n_base = n_base;
            lc.task_count = 0;
        }
    }
    grp_sync(grp);
}

#include "engine_resident.cuh"

// ---- order of the resident lists -------------------------------------------------------------------------------------
constexpr int W_BINS = 64;          // 8 x 8 cost bins
constexpr int W_KEYS = 2 * W_BINS;  // lcs with an outer integration, then the single-round ones (WaveCounters::rkey_n)
VLR_DEV int wave_share_bin(int n_alt, int n) { // share of the pileup's reads that do not favour the reference
    const long long a = 100LL * n_alt, m = n; // floor(100 n_alt / n) < k  <=>  100 n_alt < k n: no division
    if (n <= 0 || a < m) return 0;
    return a < 3 * m ? 1 : (a < 8 * m ? 2 : (a < 20 * m ? 3 : (a < 40 * m ? 4 : (a < 70 * m ? 5 : 6))));
}
// Tasks of round 0 of config `ci` of a locus, and whether an outer integration is among them (artifact configs only
// evaluate the events that have a twin).
VLR_DEV int wave_lc_shape(const DevScenario* sc, const WaveLocus& wl, int ci, bool& outer) {
    int n_tasks = 0;
    outer = false;
    for (int e = 0; e < sc->E; ++e) {
        if (ci > 0 && !sc->events[e].has_artifact_twin) continue;
        if (wl.ev_kind[e] == 1) n_tasks += 1;
        if (wl.ev_kind[e] == 3) {
            n_tasks += 2;
            outer = true;
        }
    }
    return n_tasks;
}
// Key of the lc in its class's resident list, -1: not on a list (not resident, or nothing to integrate).
VLR_DEV int wave_lc_key(const DevScenario* sc, const WaveLocus& wl, int ci) {
    if (!wl.resident) return -1;
    bool outer;
    if (wave_lc_shape(sc, wl, ci, outer) == 0) return -1;
    return (outer ? 0 : W_BINS) + wl.cost_bin;
}
// Exclusive prefix sums of a class's key counts: where each key's block starts in the list.
VLR_DEV void wave_key_offsets(const unsigned* key_n, unsigned* off) {
    unsigned acc = 0;
    for (int k = 0; k < W_KEYS; ++k) {
        off[k] = acc;
        acc += key_n[k];
    }
}


// ---------------------------------------------------------------------------------------------- prep
// Three kernels, so that each one's code fits the instruction caches (a single warp-per-locus prep kernel spent most of
// its cycles waiting for instruction fetch: 43 KB of text, 16 warps per SM in different phases):
//   wave_pre_locus  warp per locus : calling.rs:586-626 + bias/*.rs (locus_prepass); per event what
//                                    GenericPosterior::density (generic.rs:191-422) decides before any integration
//                                    starts; lc and coefficient-arena allocation
//   wave_lc_init    thread per lc  : the lc record and its round-0 tasks (cold scalar code, 32 lcs per instruction)
//   wave_lc_coef    warp per lc    : likelihood.rs hoisting (read_coefficients) for both samples, point events
VLR_DEV void wave_pre_locus(const DevScenario* sc, const DevBatch* b, const WavePlan& wp, const WaveBufs& wb, int64_t locus,
                            int li, bool want_be, Ctx& c) {
    const int E = sc->E, P = wp.P, T = wp.T;
    c.sc = sc;
    c.b = b;
    c.locus = locus;
    c.status = 0;
    c.lf = b->lflags[locus];
    BiasPlan plan;
    locus_prepass(c, plan);
    WaveLocus& wl = wb.loci[li];

    // ---- per event: pruned (ln 0 without evaluation), leaf task, point evaluation, or the outer integration
    int ev_kind[MAXE]; // 0 pruned, 1 leaf task under a discrete parent, 2 point, 3 outer
    double ev_a[MAXE], ev_b[MAXE];
    double pa = 0.0, pb = 0.0;
    bool defer = c.s_gt1[P] || c.s_gt1[T];
    const vlr_sample_t& smP = sc->samples[P];
    const vlr_sample_t& smT = sc->samples[T];
    for (int e = 0; e < E; ++e) {
        ev_kind[e] = 0;
        ev_a[e] = ev_b[e] = 0.0;
        const vlr_node_t& root = sc->nodes[wp.root_node[e]];
        const vlr_node_t& child = sc->nodes[wp.child_node[e]];
        if (root.kind == VLR_NODE_SET) {
            if (c.clear_ref[P] && sc->set_vafs[root.vaf_offset] > 0.0) continue; // generic.rs:294-299
        } else {
            if (c.clear_ref[P] && root.start > 0.0) continue; // generic.rs:342-347
            Range r{root.start, root.end, root.left_exclusive != 0, root.right_exclusive != 0};
            const double mn = range_observable_min(r, c.n_obs[P]), mx = range_observable_max(r, c.n_obs[P]);
            if (!(mn <= mx) || (mx - mn) < smP.resolution || c.n_obs[P] < 5) { // Simpson fallbacks: generic engine
                defer = true;
                continue;
            }
            // a limit on an excluded bound (< 10 reads, narrow ranges: formula.rs:1172-1224) is evaluated but belongs to
            // another event; the generic engine offers such points to every event's MAP slot (calling.rs:861-864)
            if ((r.lex && mn <= r.start) || (r.rex && mx >= r.end)) {
                defer = true;
                continue;
            }
            pa = mn;
            pb = mx;
        }
        if (child.kind == VLR_NODE_SET) {
            if (c.clear_ref[T] && sc->set_vafs[child.vaf_offset] > 0.0) continue;
            ev_kind[e] = 2;
        } else {
            if (c.clear_ref[T] && child.start > 0.0) continue;
            Range r{child.start, child.end, child.left_exclusive != 0, child.right_exclusive != 0};
            const double mn = range_observable_min(r, c.n_obs[T]), mx = range_observable_max(r, c.n_obs[T]);
            if (!(mn <= mx) || (mx - mn) < smT.resolution || c.n_obs[T] < 5) {
                defer = true;
                continue;
            }
            if ((r.lex && mn <= r.start) || (r.rex && mx >= r.end)) {
                defer = true;
                continue;
            }
            bool covered = false; // one Range spectrum of T's universe covers every abscissa: the prior is a constant
            for (int i = 0; i < smT.n_universe; ++i) {
                const vlr_spectrum_t& sp = sc->spectra[smT.universe_offset + i];
                if (sp.kind != VLR_SPECTRUM_RANGE) continue;
                Range u{sp.start, sp.end, sp.left_exclusive != 0, sp.right_exclusive != 0};
                if (range_contains(u, mn) && range_contains(u, mx) && !(smT.ploidy == 0 && mx != 0.0)) covered = true;
            }
            if (!covered) {
                defer = true;
                continue;
            }
            ev_a[e] = mn;
            ev_b[e] = mx;
            ev_kind[e] = root.kind == VLR_NODE_SET ? 1 : 3;
        }
    }
    const int n_cfg = 1 + plan.n_surviving;
    const int coef_total = c.coef_total;
    const int resident = (wb.allow_resident && c.s_one[P] && c.s_one[T]) ? r_class(c.n_obs[P], c.n_obs[T], wb.cscratch_reads) : 0;
    const int lc_doubles = resident ? R_QW * (r_qcap(c.n_obs[P]) + r_qcap(c.n_obs[T])) : 4 * coef_total;
    int lc_base = 0;
    int64_t coef_base = 0;
    bool allocated = false;
    if (!defer) {
        allocated = true;
        unsigned lb = 0;
        unsigned long long cb = 0;
        if (lane_id() == 0) {
            lb = wa_add_u32(&wb.cnt->n_lc, (unsigned)n_cfg);
            cb = wa_add_u64(&wb.cnt->coef_used, (unsigned long long)n_cfg * (unsigned long long)lc_doubles);
        }
#ifndef VLR_HOST_EMU
        lb = __shfl_sync(FULL, lb, 0, LANES);
        cb = __shfl_sync(FULL, cb, 0, LANES);
#endif
        lc_base = (int)lb;
        coef_base = (int64_t)cb;
        if ((int64_t)lb + n_cfg > (int64_t)wb.lc_cap || coef_base + (int64_t)n_cfg * lc_doubles > wb.coef_cap) defer = true;
    }
    if (lane_id() != 0) return;
    if (defer) {
        if (allocated) // lcs were allocated before a table or the arena ran out: mark the ones inside the table as dead
            for (int ci = 0; ci < n_cfg; ++ci)
                if ((int64_t)lc_base + ci < (int64_t)wb.lc_cap) wb.lcs[lc_base + ci].li = -1;
        const unsigned d = wa_add_u32(&wb.cnt->n_deferred, 1u);
        wb.deferred[d] = (int)locus;
        wl.lc_base = 0;
        wl.n_cfg = 0;
        wl.n_twins = 0;
        wl.status = 0;
        return;
    }
    wl.lc_base = lc_base;
    wl.n_cfg = n_cfg;
    wl.n_twins = plan.n_twins;
    wl.status = c.status; // hints of the pre-pass (singleton adjustment, filtered alignments)
    wl.lf = c.lf;
    wl.coef_base = coef_base;
    wl.coef_total = coef_total;
    wl.resident = resident;
    wl.lc_doubles = lc_doubles;
    wl.singleton_row = c.singleton_row;
    wl.forward_rate = plan.forward_rate;
    wl.ln_fwd = plan.ln_fwd;
    wl.ln_rev = plan.ln_rev;
    wl.has_alt_loci = plan.has_alt_loci ? 1 : 0;
    for (int k = 0; k < NCFG; ++k) wl.surviving[k] = k < plan.n_surviving ? plan.surviving[k] : 0;
    for (int s = 0; s < 2; ++s) {
        wl.n_obs[s] = c.n_obs[s];
        wl.s_one[s] = c.s_one[s];
        wl.coef_off[s] = c.coef_off[s];
    }
    wl.pa = pa;
    wl.pb = pb;
    for (int e = 0; e < MAXE; ++e) {
        wl.ev_kind[e] = e < E ? (int8_t)ev_kind[e] : 0;
        wl.ev_a[e] = e < E ? ev_a[e] : 0.0;
        wl.ev_b[e] = e < E ? ev_b[e] : 0.0;
    }
    wl.cost_bin = 8 * wave_share_bin(c.n_notref[T], c.n_obs[T]) + wave_share_bin(c.n_notref[P], c.n_obs[P]);
    if (want_be) wb.be_n[li] = 0;
    for (int ci = 0; ci < n_cfg; ++ci) { // so that the per-lc kernels find their locus
        WaveLC& lc = wb.lcs[lc_base + ci];
        lc.li = li;
        lc.ci = ci;
    }
    if (resident) { // how many lcs every key of the class's list will hold (config 0, then the artifact configs)
        int n_res = 0;
        for (int ci = 0; ci < 2 && ci < n_cfg; ++ci) {
            const int key = wave_lc_key(sc, wl, ci);
            const unsigned n = ci == 0 ? 1u : (unsigned)(n_cfg - 1);
            if (key < 0) continue;
            wa_add_u32(&wb.cnt->rkey_n[resident - 1][key], n);
            n_res += (int)n;
        }
        if (n_res) wa_add_u32(&wb.cnt->rlist_total[resident - 1], (unsigned)n_res);
    }
}

// One thread per lc: everything of the lc record that does not need the reads, and the round-0 tasks. Returns
// W_KEYS * (class - 1) + key when the lc belongs on a resident list (the caller places it: a warp's lcs of one key as
// one block in lc order, the lcs of a locus next to each other), else -1.
VLR_DEV int wave_lc_init(const DevScenario* sc, const WavePlan& wp, const WaveBufs& wb, int lci) {
    WaveLC& lc = wb.lcs[lci];
    if (lc.li < 0) return -1; // dead (its locus was deferred after allocation)
    const WaveLocus& wl = wb.loci[lc.li];
    const int E = sc->E, P = wp.P, T = wp.T;
    const int ci = lc.ci;
    lc.art_id = ci == 0 ? 0 : wl.surviving[ci - 1];
    lc.nP = wl.n_obs[P];
    lc.nT = wl.n_obs[T];
    lc.m0P = wl.s_one[P];
    lc.m0T = wl.s_one[T];
    lc.status = 0;
    lc.n_base = 0;
    lc.resident = wl.resident;
    lc.nqPx = lc.nqPy = lc.nqTx = lc.nqTy = 0; // wave_lc_coef
    lc.coefP = wl.coef_base + (int64_t)ci * wl.lc_doubles + (wl.resident ? 0 : 4 * wl.coef_off[P]);
    lc.coefT = wl.coef_base + (int64_t)ci * wl.lc_doubles + (wl.resident ? 0 : 4 * wl.coef_off[T]);
    lc.ksumP = lc.ksumT = 0.0; // wave_lc_coef
    lc.outer_pending = 0;
    lc.outer_n = 0;
    lc.outer_overflow = 0;
    lc.outer_skip = 0;
    lc.ta = lc.tb = 0.0;
    for (int e = 0; e < MAXE; ++e) {
        lc.dens[e] = neg_inf();
        lc.map_set[e] = 0;
        lc.map_joint[e] = lc.map_vp[e] = lc.map_vt[e] = 0.0;
        lc.map_disc[e] = 3;
    }
    bool outer;
    const int n_tasks = wave_lc_shape(sc, wl, ci, outer);
    lc.task_base = 0;
    lc.task_count = 0;
    if (n_tasks == 0) return -1;
    if (wl.resident) // its tasks live in the shared memory of the group that takes the lc (r_first_tasks)
        return W_KEYS * (wl.resident - 1) + wave_lc_key(sc, wl, ci);
    const unsigned tb = wa_add_u32(&wb.cnt->task_n[0], (unsigned)n_tasks);
    WaveTask* nt = wb.tasks[0] + tb;
    int k = 0;
    for (int e = 0; e < E; ++e) {
        if (ci > 0 && !sc->events[e].has_artifact_twin) continue;
        if (wl.ev_kind[e] == 1) {
            const double v = sc->set_vafs[sc->nodes[wp.root_node[e]].vaf_offset];
            wave_emit_task(nt[k++], lci, e, v, true, wl.ev_a[e], wl.ev_b[e]);
        } else if (wl.ev_kind[e] == 3) {
            Adaptive st;
            st.init(wl.pa, wl.pb, sc->samples[P].resolution);
            double xs[8];
            const int np = st.points(xs); // [min, max]
            for (int i = 0; i < np; ++i) {
                lc.outer_xs[i] = xs[i];
                wave_emit_task(nt[k++], lci, e, xs[i], false, wl.ev_a[e], wl.ev_b[e]);
            }
            lc.outer = st;
            lc.outer_pending = 1;
            lc.ta = wl.ev_a[e];
            lc.tb = wl.ev_b[e];
        }
    }
    lc.task_base = (int)tb;
    lc.task_count = k;
    wave_list_append(wb, 0, lci, lc.nP, lc.nT);
    return -1;
}

// One warp per lc: the per-read coefficients of both samples under the lc's artifact config, and the point events
// (both nodes a single VAF, e.g. the absent event): one joint evaluation each, like joint() of the generic engine.
VLR_DEV void wave_lc_coef(const DevScenario* sc, const DevBatch* b, const WavePlan& wp, const WaveBufs& wb, int lci,
                          int64_t sub_lo, bool want_be, Ctx& c, int warp_global, MemoTab* memo = nullptr) {
    WaveLC& lc = wb.lcs[lci];
    const int li = lc.li, ci = lc.ci;
    if (li < 0) return; // dead
    const WaveLocus& wl = wb.loci[li];
    const int E = sc->E, P = wp.P, T = wp.T;
    c.sc = sc;
    c.b = b;
    c.locus = sub_lo + li;
    c.status = 0;
    c.lf = wl.lf;
    c.singleton_row = wl.singleton_row;
    c.n_pileup_evals = 0;
    c.art.id = ci == 0 ? 0 : wl.surviving[ci - 1];
    c.art.forward_rate = wl.forward_rate;
    c.art.ln_fwd = wl.ln_fwd;
    c.art.ln_rev = wl.ln_rev;
    c.art.has_alt_loci = wl.has_alt_loci != 0;
    // resident lcs: per-read coefficients into the warp's scratch (the point events below evaluate them there), then
    // multiplied out into the arena's pileup polynomials; other lcs: per-read coefficients straight into the arena
    double* const scratch = wb.cscratch + (size_t)warp_global * 6 * (size_t)wb.cscratch_reads;
    c.coef = wl.resident ? scratch : wb.coef + (wl.coef_base + (int64_t)ci * wl.lc_doubles);
    c.coef_in_sm = 0;
    c.coef_cap = wl.coef_total;
    c.coef_total = wl.coef_total;
    for (int s = 0; s < 2; ++s) {
        c.n_obs[s] = wl.n_obs[s];
        c.s_one[s] = wl.s_one[s];
        c.s_gt1[s] = 0;
        c.coef_off[s] = wl.coef_off[s];
    }
    warp_sync();
    for (int s = 0; s < 2; ++s) read_coefficients(c, s, memo);
    double* be = want_be ? wb.be + (size_t)li * BE_CAP * 4 : nullptr;
    uint32_t n_base = 0;
    for (int e = 0; e < E; ++e) {
        if (wl.ev_kind[e] != 2 || (ci > 0 && !sc->events[e].has_artifact_twin)) continue;
        const double v = sc->set_vafs[sc->nodes[wp.root_node[e]].vaf_offset];
        const double w = sc->set_vafs[sc->nodes[wp.child_node[e]].vaf_offset];
        const double prior = (wave_prior_ok(sc, P, v) && wave_prior_ok(sc, T, w)) ? 0.0 : neg_inf();
        double lh = 0.0;
        for (int s = 0; s < 2; ++s) { // sample-index order (generic.rs:511-551)
            const double vs = s == P ? v : w;
            const double by = sc->samples[s].contamination_by >= 0 ? v : 0.0;
            lh += sample_likelihood_call(c, s, vs, by);
        }
        const double j = prior + lh;
        if (j != j) c.status |= VLR_ST_NAN;
        n_base++;
        if (lane_id() == 0) {
            lc.dens[e] = j;
            lc.map_set[e] = 1;
            lc.map_joint[e] = j;
            lc.map_vp[e] = v;
            lc.map_vt[e] = w;
            lc.map_disc[e] = 3;
            if (be != nullptr && ci == 0) {
                const unsigned at = wa_add_u32(&wb.be_n[li], 1u);
                if (at < (unsigned)BE_CAP) {
                    double* r = be + (size_t)at * 4;
                    r[0] = j;
                    r[1] = d_make(0, 3);
                    r[2 + P] = v;
                    r[2 + T] = w;
                } else {
                    c.status |= VLR_ST_BASE_EVENTS_OVERFLOW;
                }
            }
        }
        warp_sync();
    }
    int nq[4] = {0, 0, 0, 0};
    if (wl.resident) {
        double* cd = scratch + 4 * (size_t)wl.coef_total;
        double* out = wb.coef + lc.coefP;
        r_build_pileup(scratch + 4 * (size_t)wl.coef_off[P], wl.n_obs[P], cd, out, nq[0], nq[1]);
        out += (size_t)(nq[0] + nq[1]) * R_QW;
        r_build_pileup(scratch + 4 * (size_t)wl.coef_off[T], wl.n_obs[T], cd, out, nq[2], nq[3]);
    }
    if (lane_id() == 0) {
        lc.ksumP = c.ksum[P];
        lc.ksumT = c.ksum[T];
        lc.status |= c.status;
        lc.n_base += n_base;
        lc.nqPx = nq[0];
        lc.nqPy = nq[1];
        lc.nqTx = nq[2];
        lc.nqTy = nq[3];
    }
    warp_sync();
}

// ---------------------------------------------------------------------------------------------- finish (warp per locus)
// rust-bio Model::compute's event loop + GenericPosterior::compute (generic.rs:430-460) over the lcs of the locus, then
// call_record / sample_infos (calling.rs:720-937): locus_tail.
VLR_DEV void wave_finish_locus(const DevScenario* sc, const DevBatch* b, const DevResults* res, const WavePlan& wp,
                               const WaveBufs& wb, WarpWs* ws, int64_t locus, int li, Ctx& c) {
    const WaveLocus& wl = wb.loci[li];
    if (wl.n_cfg == 0) return; // deferred: the generic engine writes this locus
    const int E = sc->E, P = wp.P, T = wp.T;
    if (res->afd_capacity > 0 && wb.be_n[li] > (unsigned)BE_CAP) {
        // more base events than the log holds: the generic engine redoes the locus (its second pass logs only what the
        // allele frequency distributions can use)
        if (lane_id() == 0) {
            const unsigned d = wa_add_u32(&wb.cnt->n_deferred, 1u);
            wb.deferred[d] = (int)locus;
        }
        warp_sync();
        return;
    }
    if (lane_id() == 0) { // one lane fills the warp's context (shared memory), the warp reads it after the sync
        c.sc = sc;
        c.b = b;
        c.res = res;
        c.ws = ws;
        c.locus = locus;
        c.status = wl.status;
        c.be = nullptr;
        c.n_rec = 0;
        if (res->afd_capacity > 0) {
            c.be = wb.be + (size_t)li * BE_CAP * 4;
            const unsigned n = wb.be_n[li];
            c.n_rec = n < (unsigned)BE_CAP ? n : (unsigned)BE_CAP;
        }
        for (int i = 0; i < 2 * E; ++i) c.map_set[i] = 0;
        for (int e = 0; e < E; ++e) {
            c.ev_plain[e].init();
            c.ev_twin[e].init();
        }
        const double twin_prior = wl.n_twins > 0 ? LN_05 + m_log(1.0 / (double)wl.n_twins) : neg_inf();
        uint32_t n_base = 0;
        for (int ci = 0; ci < wl.n_cfg; ++ci) {
            const WaveLC& lc = wb.lcs[wl.lc_base + ci];
            c.status |= lc.status;
            if (lc.outer_pending || lc.task_count != 0) c.status |= VLR_ST_GRID_OVERFLOW; // round budget exceeded
            n_base += lc.n_base;
            for (int e = 0; e < E; ++e) {
                if (ci > 0 && !sc->events[e].has_artifact_twin) continue;
                const double d = lc.dens[e];
                if (d != d) c.status |= VLR_ST_NAN;
                if (ci == 0) c.ev_plain[e].add(LN_05 + d);
                else c.ev_twin[e].add(twin_prior + d);
                const int slot = 2 * e + (ci > 0 ? 1 : 0);
                if (lc.map_set[e] && (!c.map_set[slot] || lc.map_joint[e] > c.map_joint[slot])) {
                    c.map_set[slot] = 1;
                    c.map_joint[slot] = lc.map_joint[e];
                    c.map_cfg[slot] = lc.art_id;
                    c.map_seq[slot] = ((uint32_t)slot << 24) | (uint32_t)ci;
                    c.map_disc[slot] = ((uint32_t)(lc.map_disc[e] & 1u) << P) | ((uint32_t)((lc.map_disc[e] >> 1) & 1u) << T);
                    c.map_vaf[slot][P] = lc.map_vp[e];
                    c.map_vaf[slot][T] = lc.map_vt[e];
                }
            }
        }
        c.n_base = n_base;
    }
    warp_sync();
    locus_tail(c, wl.n_twins);
}
