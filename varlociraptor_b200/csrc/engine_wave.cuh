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
constexpr int W_GROUP = 32;      // lcs per CTA group (= shared-memory coefficient slots)
constexpr int W_SLOT_READS = 104; // reads per shared-memory coefficient slot (3.3 KB); deeper pileups are read from L2

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

struct WaveCounters {
    unsigned long long ticket[4]; // prep, finish, deferred (generic kernel), spare
    unsigned long long coef_used; // reads allocated in the coefficient arena
    unsigned int n_lc, n_deferred;
    unsigned int list_n[W_MAXROUNDS + 2];
    unsigned int task_n[W_MAXROUNDS + 2];
};

struct WaveLocus {
    int lc_base, n_cfg; // n_cfg = 0: deferred to the generic engine
    int n_twins;
    uint32_t status;
};

struct WaveLC { // one (locus, artifact config)
    int li;     // locus index inside the sub-chunk
    int ci, art_id;
    int nP, nT;
    int m0P, m0T; // every kept read of the sample has prob_sample_alt == 0
    int task_base, task_count;
    uint32_t status, n_base;
    int outer_pending, outer_n, outer_overflow;
    int64_t coefP, coefT; // first read of the sample's coefficients in the arena
    double ksumP, ksumT;
    double ta, tb; // leaf integration limits under the outer event
    double dens[MAXE];
    double map_joint[MAXE], map_vp[MAXE], map_vt[MAXE];
    uint8_t map_set[MAXE], map_disc[MAXE]; // disc: bit 0 parent event discrete, bit 1 leaf event discrete
    Adaptive outer;
    double outer_xs[8];
};

struct WaveTask {
    int lc;
    short event;
    uint8_t parent_disc, pad;
    double parent_x, a, b; // in
    double value, best_f, best_x; // out
    uint32_t n_evals, status;
};

struct WaveBufs {
    WaveCounters* cnt;
    WaveLocus* loci;
    WaveLC* lcs;
    double* og_x; // [lc_cap][W_OGRID]
    double* og_f;
    double* coef; // arena, 4 doubles per read
    WaveTask* tasks[2];
    int* list[2];
    int* deferred; // absolute locus indices
    double* gx;    // per-thread leaf grids: [W_GCAP][g_stride]
    double* gf;
    short* gn;
    double* be;      // base-event log per locus of the sub-chunk (AFD only): [n_sub][BE_CAP][2 + S]
    unsigned* be_n;  // [n_sub]
    int64_t coef_cap; // reads
    int lc_cap;
    int g_stride;
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

// ln-likelihood of one pileup at NP abscissae at once (one thread): product over the reads of
// alpha x + beta y + gamma as mantissa + binary exponent, one log per abscissa.
// M0: every read has prob_sample_alt = 0 (u_r = 0), the per-read x/y corrections vanish.
template <bool M0, int NP>
VLR_DEV void wave_eval(const double2* __restrict__ co, int n, double ksum, const WArgs* a, double* lnl, unsigned& slowmask) {
    double acc[NP];
    int ex[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        acc[i] = 1.0;
        ex[i] = 0;
    }
    unsigned slow = 0;
    auto step = [&](int r) {
        const double2 ab = co[2 * r];
        const double2 gu = co[2 * r + 1];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            double t;
            if (M0) {
                t = fma(ab.x, a[i].xu, fma(ab.y, a[i].Yp, gu.x));
            } else {
                const double xr = fma(-gu.y, a[i].X1, a[i].xu);
                const double yr = fma(gu.y, a[i].X1, a[i].Yp);
                t = fma(ab.x, xr, fma(ab.y, yr, gu.x));
            }
            acc[i] *= t;
        }
    };
    int r = 0;
#pragma unroll 1
    for (; r + 4 <= n; r += 4) { // 4 factors (each <= 3) between exponent pulls
        step(r);
        step(r + 1);
        step(r + 2);
        step(r + 3);
#pragma unroll
        for (int i = 0; i < NP; ++i) wave_pull(acc[i], ex[i], slow, 1u << i);
    }
#pragma unroll 1
    for (; r < n; ++r) step(r);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        wave_pull(acc[i], ex[i], slow, 1u << i);
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

template <int NP>
VLR_DEV void wave_pileup(const double2* co, int n, double ksum, bool m0, const WArgs* a, int nvalid, double* lnl) {
    if (n == 0) { // empty fold = ln 1
#pragma unroll
        for (int i = 0; i < NP; ++i) lnl[i] = 0.0;
        return;
    }
    unsigned slow;
    if (m0) wave_eval<true, NP>(co, n, ksum, a, lnl, slow);
    else wave_eval<false, NP>(co, n, ksum, a, lnl, slow);
    if (slow) {
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
// parent_x: utils/adaptive_integration.rs:25-141 driven exactly like integrate_adaptive_leaf / leaf_multi_run, with the
// visited points kept in a per-thread linked list in x order (so the closing trapezoid needs no sort).
VLR_DEV void wave_task_run(const DevScenario* sc, const WavePlan& wp, const WaveLC& lc, WaveTask& t, const double2* coT,
                           const double2* coP, double* gx, double* gf, short* gn, const int gs, double* be, unsigned* be_n) {
    const int P = wp.P, T = wp.T;
    const vlr_sample_t& smT = sc->samples[T];
    double rhoT = 1.0, iotaT = 0.0;
    if (smT.contamination_by >= 0) {
        rhoT = 1.0 - smT.contamination_fraction; // e^{purity}
        iotaT = 1.0 - rhoT;                      // e^{impurity} (likelihood.rs:77-84)
    }
    const double px = t.parent_x;
    const double vby = smT.contamination_by >= 0 ? px : 0.0;
    const int nT = lc.nT;
    const double ksumT = lc.ksumT;
    const bool m0T = lc.m0T != 0;
    uint32_t status = 0;
    const double prior_const = wave_prior_ok(sc, P, px) ? 0.0 : neg_inf();
    double lh_const;
    {
        const WArgs ap = wave_args(1.0, 0.0, px, 0.0);
        wave_pileup<1>(coP, lc.nP, lc.ksumP, lc.m0P != 0, &ap, 1, &lh_const);
    }
    const double a = t.a, b = t.b, res = smT.resolution;
    int n = 0;
    uint32_t n_evals = 0;
    bool overflow = false, have_best = false, any_nan = false;
    double best_f = 0.0, best_x = 0.0;

    auto insert = [&](double x, double f, int hint) -> int {
        n_evals++;
        if (f != f) any_nan = true;
        if (!have_best || f > best_f) { // first maximum in visit order (calling.rs:851-870 via joint())
            have_best = true;
            best_f = f;
            best_x = x;
        }
        if (n >= W_GCAP) {
            overflow = true;
            return hint;
        }
        const int idx = n++;
        gx[idx * gs] = x;
        gf[idx * gs] = f;
        if (idx == 0) {
            gn[0] = -1;
            return 0;
        }
        int cur = hint;
        for (;;) { // equal abscissae keep visit order, like the rank sort of grid_trapezoid
            const int nx = gn[cur * gs];
            if (nx < 0 || !(gx[nx * gs] <= x)) break;
            cur = nx;
        }
        gn[idx * gs] = gn[cur * gs];
        gn[cur * gs] = (short)idx;
        return idx;
    };
    auto eval3 = [&](double x0, double x1, double x2, int nvalid, double* f) {
        WArgs w[3];
        w[0] = wave_args(rhoT, iotaT, x0, vby);
        w[1] = wave_args(rhoT, iotaT, x1, vby);
        w[2] = wave_args(rhoT, iotaT, x2, vby);
        double lnl[3];
        wave_pileup<3>(coT, nT, ksumT, m0T, w, nvalid, lnl);
#pragma unroll
        for (int i = 0; i < 3; ++i) f[i] = prior_const + (lh_const + lnl[i]);
    };

    double left = a, right = b, f_left, f_right, middle = 0.0, first_middle = 0.0;
    bool have_middle = false;
    int il, ir;
    double f[3];
    eval3(a, b, b, 2, f);
    il = insert(a, f[0], 0);
    ir = insert(b, f[1], il);
    f_left = f[0];
    f_right = f[1];
    // while (((right - left) >= res && left < right) || middle.is_none())   (adaptive_integration.rs:52)
    while (!overflow && ((((right - left) >= res) && left < right) || !have_middle)) {
        middle = (right + left) / 2.0;
        const double m1 = (middle + left) / 2.0, m2 = (right + middle) / 2.0;
        eval3(middle, m1, m2, 3, f);
        const int im = insert(middle, f[0], il);
        const int i1 = insert(m1, f[1], il);
        const int i2 = insert(m2, f[2], im);
        if (!have_middle) first_middle = middle;
        have_middle = true;
        const double f_m1 = f[1], f_m2 = f[2];
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
        const int nil = idx <= 1 ? il : (idx == 2 ? i1 : i2), nir = idx == 0 ? i1 : (idx == 1 ? i2 : ir);
        left = nl;
        f_left = nfl;
        right = nr;
        f_right = nfr;
        il = nil;
        ir = nir;
    }
    (void)ir;
    { // the abandoned arm's midpoint and 3 + 3 points around the optimum (adaptive_integration.rs:96-131)
        const double x0 = (middle < first_middle) ? (b + first_middle) / 2.0 : (first_middle + a) / 2.0;
        const double lo = fmax(middle - (res * 3.0), a);
        const double slo = (middle - lo) / 3.0;
        const double x1 = lo + slo * 0.0, x2 = lo + slo * 1.0, x3 = lo + slo * 2.0;
        const double hi = fmin(middle + (res * 3.0), b);
        const double shi = (hi - middle) / 3.0;
        const double x4 = middle + shi * 1.0, x5 = middle + shi * 2.0, x6 = middle + shi * 3.0;
        eval3(x0, x1, x2, 3, f);
        insert(x0, f[0], 0);
        int h = insert(x1, f[1], 0);
        h = insert(x2, f[2], h);
        eval3(x3, x4, x5, 3, f);
        h = insert(x3, f[0], h);
        h = insert(x4, f[1], h);
        h = insert(x5, f[2], h);
        eval3(x6, x6, x6, 1, f);
        insert(x6, f[0], h);
    }
    // ln_trapezoidal_integrate_grid_exp over the visited points in x order (rust-bio), summed in linear
    // space relative to the maximum: ln( sum_i (e^{f_i} + e^{f_i+1}) / 2 * (x_i+1 - x_i) )
    double value;
    if (any_nan) {
        status |= VLR_ST_NAN;
        value = NAN;
    } else if (best_f == neg_inf()) {
        value = neg_inf();
    } else {
        double xp = gx[0], ep = exp(gf[0] - best_f), sum = 0.0;
        for (int nx = gn[0]; nx >= 0; nx = gn[nx * gs]) {
            const double xc = gx[nx * gs], ec = exp(gf[nx * gs] - best_f);
            sum += (ep + ec) * (xc - xp);
            xp = xc;
            ep = ec;
        }
        value = best_f + log(sum * 0.5);
    }
    if (overflow) status |= VLR_ST_GRID_OVERFLOW;
    if (be != nullptr && lc.ci == 0) { // base events of the artifact-free config feed the AFD (calling.rs:891-928)
        const unsigned base = wa_add_u32(be_n, (unsigned)n);
        const double disc = d_make(0, (int)((t.parent_disc ? 1u : 0u) << P));
        for (int i = 0; i < n; ++i) {
            const unsigned at = base + (unsigned)i;
            if (at >= (unsigned)BE_CAP) {
                status |= VLR_ST_BASE_EVENTS_OVERFLOW;
                break;
            }
            double* e = be + (size_t)at * 4;
            e[0] = gf[i * gs];
            e[1] = disc;
            e[2 + P] = px;
            e[2 + T] = gx[i * gs];
        }
    }
    t.value = value;
    t.best_f = best_f;
    t.best_x = best_x;
    t.n_evals = n_evals;
    t.status = status;
}

// ---------------------------------------------------------------------------------------------- lc advance (thread)
// trapezoid over an outer grid (<= W_OGRID points, in place): stable insertion sort by x, then the same sum as above
VLR_DEV_NOINLINE double wave_outer_trapezoid(double* x, double* f, int n, uint32_t& status) {
    bool any_nan = false;
    double fmx = neg_inf();
    for (int i = 0; i < n; ++i) {
        if (f[i] != f[i]) any_nan = true;
        if (f[i] > fmx) fmx = f[i];
    }
    if (any_nan) {
        status |= VLR_ST_NAN;
        return NAN;
    }
    if (n < 2 || fmx == neg_inf()) return neg_inf();
    for (int i = 1; i < n; ++i) {
        const double xi = x[i], fi = f[i];
        int j = i - 1;
        while (j >= 0 && x[j] > xi) {
            x[j + 1] = x[j];
            f[j + 1] = f[j];
            --j;
        }
        x[j + 1] = xi;
        f[j + 1] = fi;
    }
    double sum = 0.0, ep = m_exp(f[0] - fmx);
    for (int i = 1; i < n; ++i) {
        const double ec = m_exp(f[i] - fmx);
        sum += (ep + ec) * (x[i] - x[i - 1]);
        ep = ec;
    }
    return fmx + m_log(sum * 0.5);
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
    t.status = 0;
}

// After the tasks of round `round` of an lc are complete: MAP bookkeeping in visit order (calling.rs:851-870), event
// densities, and the next step of the enclosing integration (integrate_adaptive_generic over the root Range).
VLR_DEV_NOINLINE void wave_lc_advance(const WavePlan& wp, const WaveBufs& wb, int lci, int round) {
    WaveLC& lc = wb.lcs[lci];
    const WaveTask* tasks = wb.tasks[round & 1] + lc.task_base;
    const int cnt = lc.task_count;
    double ofs[8];
    int no = 0;
    for (int i = 0; i < cnt; ++i) {
        const WaveTask& t = tasks[i];
        const int e = t.event;
        lc.status |= t.status;
        lc.n_base += t.n_evals;
        if (t.n_evals > 0 && (!lc.map_set[e] || t.best_f > lc.map_joint[e])) {
            lc.map_set[e] = 1;
            lc.map_joint[e] = t.best_f;
            lc.map_vp[e] = t.parent_x;
            lc.map_vt[e] = t.best_x;
            lc.map_disc[e] = t.parent_disc ? 1 : 0;
        }
        if (e == wp.outer_event) {
            if (no < 8) ofs[no] = t.value;
            no++;
        } else {
            lc.dens[e] = t.value;
        }
    }
    lc.task_count = 0;
    if (!lc.outer_pending) return;
    double* ox = wb.og_x + (size_t)lci * W_OGRID;
    double* of = wb.og_f + (size_t)lci * W_OGRID;
    for (int i = 0; i < no && i < 8; ++i) {
        if (ofs[i] != ofs[i]) lc.status |= VLR_ST_NAN;
        if (lc.outer_n < W_OGRID) {
            ox[lc.outer_n] = lc.outer_xs[i];
            of[lc.outer_n] = ofs[i];
            lc.outer_n++;
        } else {
            lc.outer_overflow = 1;
        }
    }
    Adaptive st = lc.outer;
    const bool more = st.consume(lc.outer_xs, ofs, lc.outer_overflow != 0);
    if (more && round + 1 < W_MAXROUNDS) {
        double xs[8];
        const int k = st.points(xs);
        lc.outer = st;
        const unsigned tb = wa_add_u32(&wb.cnt->task_n[round + 1], (unsigned)k);
        WaveTask* nt = wb.tasks[(round + 1) & 1] + tb;
        for (int i = 0; i < k; ++i) {
            lc.outer_xs[i] = xs[i];
            wave_emit_task(nt[i], lci, wp.outer_event, xs[i], false, lc.ta, lc.tb);
        }
        lc.task_base = (int)tb;
        lc.task_count = k;
        const unsigned li = wa_add_u32(&wb.cnt->list_n[round + 1], 1u);
        wb.list[(round + 1) & 1][li] = lci;
    } else {
        lc.outer = st;
        lc.outer_pending = 0;
        if (lc.outer_overflow || more) lc.status |= VLR_ST_GRID_OVERFLOW;
        lc.dens[wp.outer_event] = wave_outer_trapezoid(ox, of, lc.outer_n, lc.status);
    }
}

// ---------------------------------------------------------------------------------------------- prep (warp per locus)
// calling.rs:586-626 + bias/*.rs (locus_prepass), per config likelihood.rs hoisting (read_coefficients), then per event
// what GenericPosterior::density (generic.rs:191-422) decides before any integration starts.
VLR_DEV void wave_prep_locus(const DevScenario* sc, const DevBatch* b, const WavePlan& wp, const WaveBufs& wb, int64_t locus,
                             int li, bool want_be, Ctx& c) {
    const int E = sc->E, P = wp.P, T = wp.T;
    c.sc = sc;
    c.b = b;
    c.res = nullptr;
    c.ws = nullptr;
    c.be = nullptr;
    c.n_rec = 0;
    c.locus = locus;
    c.status = 0;
    c.lf = b->lflags[locus];
    c.vartype = (c.lf >> VLR_LF_VARTYPE_SHIFT) & 3;
    c.n_base = 0;
    c.n_pileup_evals = 0;
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
    int lc_base = 0;
    int64_t coef_base = 0;
    if (!defer) {
        unsigned lb = 0;
        unsigned long long cb = 0;
        if (lane_id() == 0) {
            lb = wa_add_u32(&wb.cnt->n_lc, (unsigned)n_cfg);
            cb = wa_add_u64(&wb.cnt->coef_used, (unsigned long long)n_cfg * (unsigned long long)coef_total);
        }
#ifndef VLR_HOST_EMU
        lb = __shfl_sync(FULL, lb, 0, LANES);
        cb = __shfl_sync(FULL, cb, 0, LANES);
#endif
        lc_base = (int)lb;
        coef_base = (int64_t)cb;
        if ((int64_t)lb + n_cfg > (int64_t)wb.lc_cap || coef_base + (int64_t)n_cfg * coef_total > wb.coef_cap) defer = true;
    }
    if (defer) {
        if (lane_id() == 0) {
            const unsigned d = wa_add_u32(&wb.cnt->n_deferred, 1u);
            wb.deferred[d] = (int)locus;
            wl.lc_base = 0;
            wl.n_cfg = 0;
            wl.n_twins = 0;
            wl.status = 0;
        }
        return;
    }
    if (lane_id() == 0) {
        wl.lc_base = lc_base;
        wl.n_cfg = n_cfg;
        wl.n_twins = plan.n_twins;
        wl.status = c.status; // hints of the pre-pass (singleton adjustment, filtered alignments)
        if (want_be) wb.be_n[li] = 0;
    }
    warp_sync();
    double* be = want_be ? wb.be + (size_t)li * BE_CAP * 4 : nullptr;
    for (int ci = 0; ci < n_cfg; ++ci) {
        c.status = 0;
        c.art.id = ci == 0 ? 0 : plan.surviving[ci - 1];
        c.art.forward_rate = plan.forward_rate;
        c.art.has_alt_loci = plan.has_alt_loci;
        c.coef = wb.coef + (coef_base + (int64_t)ci * coef_total) * 4;
        c.coef_in_sm = 0;
        c.coef_cap = coef_total;
        for (int s = 0; s < 2; ++s) read_coefficients(c, s);
        const int lci = lc_base + ci;
        WaveLC& lc = wb.lcs[lci];
        uint32_t n_base = 0;
        // point events (both nodes a single VAF, e.g. the absent event): one joint evaluation here (generic.rs joint())
        double dens[MAXE], mj[MAXE], mvp[MAXE], mvt[MAXE];
        int mset[MAXE];
        int n_tasks = 0;
        for (int e = 0; e < E; ++e) {
            dens[e] = neg_inf();
            mset[e] = 0;
            mj[e] = mvp[e] = mvt[e] = 0.0;
            if (ci > 0 && !sc->events[e].has_artifact_twin) continue;
            if (ev_kind[e] == 1) n_tasks += 1;
            if (ev_kind[e] == 3) n_tasks += 2;
            if (ev_kind[e] != 2) continue;
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
            dens[e] = j;
            mset[e] = 1;
            mj[e] = j;
            mvp[e] = v;
            mvt[e] = w;
            if (be != nullptr && ci == 0 && lane_id() == 0) {
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
        if (lane_id() == 0) {
            lc.li = li;
            lc.ci = ci;
            lc.art_id = c.art.id;
            lc.nP = c.n_obs[P];
            lc.nT = c.n_obs[T];
            lc.m0P = c.s_one[P];
            lc.m0T = c.s_one[T];
            lc.status = c.status;
            lc.n_base = n_base;
            lc.coefP = coef_base + (int64_t)ci * coef_total + c.coef_off[P];
            lc.coefT = coef_base + (int64_t)ci * coef_total + c.coef_off[T];
            lc.ksumP = c.ksum[P];
            lc.ksumT = c.ksum[T];
            lc.outer_pending = 0;
            lc.outer_n = 0;
            lc.outer_overflow = 0;
            lc.ta = lc.tb = 0.0;
            for (int e = 0; e < MAXE; ++e) {
                lc.dens[e] = e < E ? dens[e] : neg_inf();
                lc.map_set[e] = e < E ? (uint8_t)mset[e] : 0;
                lc.map_joint[e] = e < E ? mj[e] : 0.0;
                lc.map_vp[e] = e < E ? mvp[e] : 0.0;
                lc.map_vt[e] = e < E ? mvt[e] : 0.0;
                lc.map_disc[e] = 3;
            }
            lc.task_base = 0;
            lc.task_count = 0;
            if (n_tasks > 0) {
                const unsigned tb = wa_add_u32(&wb.cnt->task_n[0], (unsigned)n_tasks);
                WaveTask* nt = wb.tasks[0] + tb;
                int k = 0;
                for (int e = 0; e < E; ++e) {
                    if (ci > 0 && !sc->events[e].has_artifact_twin) continue;
                    if (ev_kind[e] == 1) {
                        const double v = sc->set_vafs[sc->nodes[wp.root_node[e]].vaf_offset];
                        wave_emit_task(nt[k++], lci, e, v, true, ev_a[e], ev_b[e]);
                    } else if (ev_kind[e] == 3) {
                        lc.outer.init(pa, pb, smP.resolution);
                        double xs[8];
                        const int np = lc.outer.points(xs); // [min, max]
                        for (int i = 0; i < np; ++i) {
                            lc.outer_xs[i] = xs[i];
                            wave_emit_task(nt[k++], lci, e, xs[i], false, ev_a[e], ev_b[e]);
                        }
                        lc.outer_pending = 1;
                        lc.ta = ev_a[e];
                        lc.tb = ev_b[e];
                    }
                }
                lc.task_base = (int)tb;
                lc.task_count = k;
                const unsigned at = wa_add_u32(&wb.cnt->list_n[0], 1u);
                wb.list[0][at] = lci;
            }
        }
        warp_sync();
    }
}

// ---------------------------------------------------------------------------------------------- finish (warp per locus)
// rust-bio Model::compute's event loop + GenericPosterior::compute (generic.rs:430-460) over the lcs of the locus, then
// call_record / sample_infos (calling.rs:720-937): locus_tail.
VLR_DEV void wave_finish_locus(const DevScenario* sc, const DevBatch* b, const DevResults* res, const WavePlan& wp,
                               const WaveBufs& wb, WarpWs* ws, int64_t locus, int li, Ctx& c) {
    const WaveLocus& wl = wb.loci[li];
    if (wl.n_cfg == 0) return; // deferred: the generic engine writes this locus
    const int E = sc->E, P = wp.P, T = wp.T;
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
                c.map_disc[slot] = ((uint32_t)(lc.map_disc[e] & 1u) << P) | ((uint32_t)((lc.map_disc[e] >> 1) & 1u) << T);
                c.map_vaf[slot][P] = lc.map_vp[e];
                c.map_vaf[slot][T] = lc.map_vt[e];
            }
        }
    }
    c.n_base = n_base;
    warp_sync();
    locus_tail(c, wl.n_twins);
}
