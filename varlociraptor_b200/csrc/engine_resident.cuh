// engine_resident.cuh — the lc-resident round engine of the wavefront pipeline (included by engine_wave.cuh).
//
// An lc (locus x artifact config) of a two-level chain scenario needs ~20 adaptive leaf integrations ("tasks") in ~5
// dependent rounds (engine_wave.cuh). The first version launched two kernels per round and re-staged every lc's
// per-read coefficients from a global arena in every round. Here an OCTET (8 lanes) takes an lc by ticket, pulls its
// pileup polynomials into shared memory once (cp.async.bulk + mbarrier) and keeps it through ALL rounds: tasks, closing
// trapezoids, the outer integration's bookkeeping and the next round's tasks never leave the SM.
//
// Two algebraic steps make the evaluation itself cheap (same model):
//  * Pileup polynomials. In linear space a read's emission is affine in the effective alt-sampling probability x
//    (likelihood.rs:86-115): alpha x + beta (1 - x) + gamma. With y = 1 - x that is  d + c x  for reads with
//    alpha > beta (d = beta + gamma, c = alpha - beta) and  d + c y  for the others (d = alpha + gamma, c = beta -
//    alpha): non-negative coefficients, no cancellation anywhere on [0, 1]. FIVE reads of a group are multiplied out
//    once per lc into a degree-5 polynomial ("quint", 6 coefficients, all >= 0); evaluating it by Horner costs 5 FMA +
//    1 MUL for five reads instead of 15 operations, and 9.6 instead of 32 bytes per read of shared memory.
//  * No transcendental inside the search. A pileup likelihood is carried as mantissa in [1, 2) and binary exponent;
//    the adaptive search (adaptive_integration.rs:25-141) only COMPARES values, which (mantissa, exponent) pairs do
//    exactly, and the closing trapezoid needs ratios to the maximum: mantissa quotient and exponent difference. One
//    log for the maximum, one for the integral, per task (the first version: a log and an exp per visited point).
//
// Served here: lcs whose kept reads all have prob_sample_alt = 0 (every SNV; x is then the same for all reads of a
// pileup) and whose polynomials fit a slot; all others take the per-round kernels of engine_wave.cuh.

constexpr int R_DEG = 5;          // reads per polynomial
constexpr int R_QW = R_DEG + 1;   // coefficients (doubles) per polynomial: 48 bytes = 3 x 16
// A slot holds the polynomials of the LEAF pileup (evaluated ~30 times per task); the parent pileup is evaluated once per
// task and round and is read from the arena (L2): half the shared memory per lc, a third more lcs per SM.
constexpr int R_SLOT_Q = 24;      // polynomials per octet slot (a pileup of <= 110 reads): 1152 bytes
constexpr int R_SLOT_QM = 104;    // ... of a warp's slot, class 2 (~500 reads): 4992 bytes
constexpr int R_SLOT_QD = 208;    // ... class 3 (~1000 reads): 9984 bytes
constexpr int R_SLOT_QL = 416;    // ... class 4 (~2070 reads): 19968 bytes
constexpr int R_CLASSES = W_RCLASSES; // 1: an octet per lc; 2..4: a warp per lc (deep pileups, depth skew)
constexpr int R_POOL_N = 400;     // sorted-list nodes per lc, shared by the tasks of a round: one 32-bit word each,
                                  // (24-bit key << 8) | next node (a task visits ~60 points at resolution 0.01)
constexpr int R_POOL_D = R_POOL_N / 2; // ... as doubles: 1600 bytes
constexpr int R_SORT = 64;        // points of the OUTER grid sorted in the (then idle) pool (larger: global scratch)
constexpr int R_NONE = 255;       // end of a sorted list
constexpr int R_ZERO_E = -(1 << 29);
constexpr int R_MAXREADS = 2 * R_DEG * R_SLOT_QL; // no resident lc has more kept reads in its two pileups (r_class)
// coefficient kernel, per warp: 4 doubles per read + (c, d) pairs of a pileup = 6 doubles per read of the deepest
// resident lc the workspace allows (WaveBufs::cscratch_reads <= R_MAXREADS)
// upper bound of the polynomials of a pileup of n reads (two groups, each rounded up)
VLR_DEV int r_qcap(int n) { return n / R_DEG + 2; }
VLR_DEV int r_slot_q(int cls) { return cls == 1 ? R_SLOT_Q : (cls == 2 ? R_SLOT_QM : (cls == 3 ? R_SLOT_QD : R_SLOT_QL)); }
// size class of an lc whose pileups hold nP and nT kept reads (0: not resident, the per-round kernels serve it): by the
// deeper of the two (the slot holds the leaf pileup; a parent pileup of another order of magnitude, read from L2 by
// the few lanes of an octet's task, would stall the octet's neighbours)
VLR_DEV int r_class(int nP, int nT, int scratch_reads) {
    if (nP + nT > scratch_reads) return 0;
    const int q = r_qcap(nP > nT ? nP : nT);
    return q <= R_SLOT_Q ? 1 : (q <= R_SLOT_QM ? 2 : (q <= R_SLOT_QD ? 3 : (q <= R_SLOT_QL ? 4 : 0)));
}

struct MV { // value = m * 2^e with m in [1, 2), or exactly zero (m = 0, e = R_ZERO_E); a NaN m is a NaN value
    double m;
    int e;
};
VLR_DEV bool mv_gt(const MV& a, const MV& b) { return a.e > b.e || (a.e == b.e && a.m > b.m); }
VLR_DEV MV mv_zero() { return MV{0.0, R_ZERO_E}; }
// ln of the value (the pileup ln-likelihood without the sum of the per-read scales)
VLR_DEV double mv_ln(const MV& v) {
    if (v.m == 0.0) return neg_inf();
    return m_log(v.m) + (double)v.e * LN_2;
}

// ---------------------------------------------------------------------------------------------- building polynomials
// One lane multiplies out up to R_DEG linear factors d_i + c_i z (identity factors pad the last polynomial of a group).
VLR_DEV void r_poly_build(const double* c, const double* d, int n, double* q) {
    q[0] = n > 0 ? d[0] : 1.0;
    q[1] = n > 0 ? c[0] : 0.0;
#pragma unroll
    for (int k = 2; k < R_QW; ++k) q[k] = 0.0;
#pragma unroll
    for (int i = 1; i < R_DEG; ++i) {
        const double ci = i < n ? c[i] : 0.0, di = i < n ? d[i] : 1.0;
#pragma unroll
        for (int k = R_DEG; k >= 1; --k)
            if (k <= i + 1) q[k] = fma(q[k - 1], ci, q[k] * di);
        q[0] = q[0] * di;
    }
}

// The polynomials of one pileup from its per-read coefficients [alpha, beta, gamma, u] (read_coefficients): group X
// (alpha > beta, variable x) first, then group Y (variable y = 1 - x). `cd` is scratch for 2 x n (c, d) pairs. Warp
// cooperative; returns the number of polynomials of each group.
VLR_DEV void r_build_pileup(const double* co, int n, double* cd, double* out, int& nqx, int& nqy) {
    // pass 1: (c, d) of every read, X reads compacted from the front of `cd`, Y reads from the back
    int nx = 0, ny = 0;
    for (int r0 = 0; r0 < n; r0 += LANES) {
        const int r = r0 + lane_id();
        const bool valid = r < n;
        double al = 0.0, be = 0.0, ga = 0.0;
        if (valid) {
            al = co[4 * r];
            be = co[4 * r + 1];
            ga = co[4 * r + 2];
        }
        const bool isx = valid && al > be;
        const bool isy = valid && !isx;
#ifdef VLR_HOST_EMU
        const int px = nx, py = ny, tx = isx ? 1 : 0, ty = isy ? 1 : 0;
#else
        const unsigned mx = w_ballot(isx), my = w_ballot(isy);
        const unsigned below = (1u << lane_id()) - 1u;
        const int px = nx + __popc(mx & below), py = ny + __popc(my & below), tx = __popc(mx), ty = __popc(my);
#endif
        if (isx) {
            cd[2 * px] = al - be;
            cd[2 * px + 1] = be + ga;
        } else if (isy) {
            const int at = n - 1 - py;
            cd[2 * at] = be - al;
            cd[2 * at + 1] = al + ga;
        }
        nx += tx;
        ny += ty;
    }
    warp_sync();
    nqx = (nx + R_DEG - 1) / R_DEG;
    nqy = (ny + R_DEG - 1) / R_DEG;
    // pass 2: one polynomial per lane
    for (int k = lane_id(); k < nqx + nqy; k += LANES) {
        double c[R_DEG], d[R_DEG], q[R_QW];
        int cnt;
        if (k < nqx) {
            const int first = k * R_DEG;
            cnt = nx - first < R_DEG ? nx - first : R_DEG;
            for (int i = 0; i < R_DEG; ++i) {
                c[i] = i < cnt ? cd[2 * (first + i)] : 0.0;
                d[i] = i < cnt ? cd[2 * (first + i) + 1] : 1.0;
            }
        } else {
            const int first = (k - nqx) * R_DEG;
            cnt = ny - first < R_DEG ? ny - first : R_DEG;
            for (int i = 0; i < R_DEG; ++i) { // Y reads were stored from the back: read them in pileup order
                c[i] = i < cnt ? cd[2 * (n - 1 - (first + i))] : 0.0;
                d[i] = i < cnt ? cd[2 * (n - 1 - (first + i)) + 1] : 1.0;
            }
        }
        r_poly_build(c, d, cnt, q);
        double* o = out + (size_t)k * R_QW;
#pragma unroll
        for (int i = 0; i < R_QW; ++i) o[i] = q[i];
    }
    warp_sync();
}

// ---------------------------------------------------------------------------------------------- evaluation
// Exponent out of the running product, mantissa back to [1, 2). The smallest and largest exponent FIELD seen tell
// afterwards whether a zero, tiny (< 2^-797), infinite or NaN product occurred (then the points are re-evaluated
// carefully): two integer min/max instead of a compare-and-select chain per pull.
VLR_DEV void r_pull(double& acc, int& ex, int& ef_min, int& ef_max) {
    const int hi = d_hi(acc);
    const int ef = (hi >> 20) & 0x7ff;
    ef_min = ef < ef_min ? ef : ef_min;
    ef_max = ef > ef_max ? ef : ef_max;
    ex += ef - 1023;
    acc = d_make((hi & 0x000fffff) | (1023 << 20), d_lo(acc));
}

VLR_DEV double r_horner(const double2 q01, const double2 q23, const double2 q45, double z) {
    double p = fma(q45.y, z, q45.x);
    p = fma(p, z, q23.y);
    p = fma(p, z, q23.x);
    p = fma(p, z, q01.y);
    return fma(p, z, q01.x);
}

// Product of the polynomials [0, nqx) at x and [nqx, nqx + nqy) at y for NP abscissae at once; the H lanes of a task
// (WSplit) take every H-th polynomial and combine their partial products with an xor butterfly (bitwise identical in
// all H lanes: their control flow stays identical). Returns true when some product left the safe range.
// SM: `q` points into the CTA's dynamic shared memory (LDS instead of generic loads).
template <int NP, bool SM>
VLR_DEV bool r_eval(const double* __restrict__ q, int nqx, int nqy, const double* x, const double* y, MV* out, const WSplit sp) {
    const double2* __restrict__ q2 = reinterpret_cast<const double2*>(q);
#ifndef VLR_HOST_EMU
    if (SM) q2 = reinterpret_cast<const double2*>(vlr_smem + (__cvta_generic_to_shared(q) - __cvta_generic_to_shared(vlr_smem)));
#endif
    double acc[NP];
    int ex[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        acc[k] = 1.0;
        ex[k] = 0;
    }
    int ef_min = 1023, ef_max = 1023;
    const int H = sp.H;
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
        const double* z = g == 0 ? x : y;
        const int lo = g == 0 ? 0 : nqx, hi = g == 0 ? nqx : nqx + nqy;
        double zz[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) zz[k] = z[k];
        int i = lo + sp.h;
        // four polynomials (each in [0, 2^5]) between exponent pulls: the product stays below 2^21, and a product that
        // underflows on the way shows up as a small exponent field at the next pull
#pragma unroll 1
        for (; i + 3 * H < hi; i += 4 * H) {
            const double2* qa = q2 + 3 * i;
            const double2* qb = q2 + 3 * (i + H);
            const double2* qc = q2 + 3 * (i + 2 * H);
            const double2* qd = q2 + 3 * (i + 3 * H);
            const double2 a0 = qa[0], a1 = qa[1], a2 = qa[2];
            const double2 b0 = qb[0], b1 = qb[1], b2 = qb[2];
            const double2 c0 = qc[0], c1 = qc[1], c2 = qc[2];
            const double2 d0 = qd[0], d1 = qd[1], d2 = qd[2];
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const double pa = r_horner(a0, a1, a2, zz[k]);
                const double pb = r_horner(b0, b1, b2, zz[k]);
                const double pc = r_horner(c0, c1, c2, zz[k]);
                const double pd = r_horner(d0, d1, d2, zz[k]);
                acc[k] *= (pa * pb) * (pc * pd);
                r_pull(acc[k], ex[k], ef_min, ef_max);
            }
        }
        if (i < hi) { // up to three left
#pragma unroll 1
            for (; i < hi; i += H) {
                const double2* qa = q2 + 3 * i;
                const double2 a0 = qa[0], a1 = qa[1], a2 = qa[2];
#pragma unroll
                for (int k = 0; k < NP; ++k) acc[k] *= r_horner(a0, a1, a2, zz[k]);
            }
#pragma unroll
            for (int k = 0; k < NP; ++k) r_pull(acc[k], ex[k], ef_min, ef_max);
        }
    }
#ifndef VLR_HOST_EMU
#pragma unroll 1
    for (int o = H >> 1; o > 0; o >>= 1) {
        ef_min = min(ef_min, __shfl_xor_sync(sp.mask, ef_min, o));
        ef_max = max(ef_max, __shfl_xor_sync(sp.mask, ef_max, o));
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            acc[k] *= __shfl_xor_sync(sp.mask, acc[k], o);
            ex[k] += __shfl_xor_sync(sp.mask, ex[k], o);
            r_pull(acc[k], ex[k], ef_min, ef_max);
        }
    }
#endif
#pragma unroll
    for (int k = 0; k < NP; ++k) out[k] = MV{acc[k], ex[k]};
    return ef_min < 226 || ef_max == 0x7ff;
}

// careful evaluation of one abscissa (a zero or denormal-range factor): every polynomial normalised on its own
VLR_DEV_NOINLINE MV r_eval_careful(const double* q, int nqx, int nqy, double x, double y) {
    double acc = 1.0;
    int ex = 0;
    for (int i = 0; i < nqx + nqy; ++i) {
        const double z = i < nqx ? x : y;
        const double* c = q + (size_t)i * R_QW;
        double p = c[5];
        for (int k = 4; k >= 0; --k) p = fma(p, z, c[k]);
        if (p != p) return MV{NAN, 0};
        if (p <= 0.0) return mv_zero();
        int e2;
        p = frexp(p, &e2) * 2.0;
        ex += e2 - 1;
        acc *= p;
        const int hi = d_hi(acc);
        ex += ((hi >> 20) & 0x7ff) - 1023;
        acc = d_make((hi & 0x800fffff) | (1023 << 20), d_lo(acc));
    }
    return MV{acc, ex};
}

template <int NP, bool SM>
VLR_DEV void r_pileup(const double* q, int nqx, int nqy, const double* x, const double* y, int nvalid, MV* out, const WSplit sp) {
    if (r_eval<NP, SM>(q, nqx, nqy, x, y, out, sp)) { // rare: every lane of the task re-evaluates (identical values)
#pragma unroll
        for (int k = 0; k < NP; ++k)
            if (k < nvalid) out[k] = r_eval_careful(q, nqx, nqy, x[k], y[k]);
    }
}

// ---------------------------------------------------------------------------------------------- per-octet state
struct RTask {
    double parent_x;             // in
    double lh_const, prior_const; // ln-likelihood of the parent pileup at parent_x; flat prior of the parent's VAF
    double best_x;
    MV best;                     // maximum over the visited points (first one in visit order)
    int event, n_evals, n_grid;
    uint32_t status;
    uint8_t parent_disc, dead, sorted, pad1; // dead: the constant part of every point is ln 0 or NaN;
                                             // sorted: the task's list holds all its points in abscissa order
};

struct RLc { // what the rounds need from the lc record, read once
    int lci, li, ci;
    int nqPx, nqPy, nqTx, nqTy; // polynomials of the parent pileup (arena, qP) and of the leaf pileup (the slot)
    int pad;
    double ksumP, ksumT;
    const double* qP;
};

struct alignas(16) ROct { // per lc group (octet or warp) in shared memory; the slot of polynomials follows it
    double pool[R_POOL_D]; // during the tasks: list nodes (unsigned[R_POOL_N]), split evenly over the round's tasks;
                           // closing the outer integration: sort scratch (R_SORT abscissae, weights, ranks)
    RTask task[W_MAXT];
    RLc lc;
    unsigned long long bar; // mbarrier of the slot's bulk copy
    double* q;              // the slot: r_slot_q(class) polynomials (128-bit loads, 16-byte bulk copies)
    unsigned long long pad;
};
static_assert(sizeof(ROct) % 16 == 0, "the slot behind the record must stay 16-byte aligned");

// ---- a task's points in abscissa order: singly linked list of 32-bit nodes, node index = visit index
// node = (key << 8) | next; key = the upper 24 bits of the abscissa rounded to float (abscissae are allele frequencies
// in [0, 1]: rounding and truncation are monotone, so a smaller key means a smaller abscissa; equal keys are decided by
// the exact abscissae in the task's grid row); next = R_NONE at the end of the list.
#ifdef VLR_HOST_EMU
VLR_DEV unsigned r_key(double x) {
    const float f = (float)x;
    unsigned u;
    memcpy(&u, &f, 4);
    return u >> 8;
}
#else
VLR_DEV unsigned r_key(double x) { return __float_as_uint((float)x) >> 8; }
#endif

// Links up to three nodes i0, i1, i2 with ascending abscissae x0 <= x1 <= x2 into the list in ONE walk that starts at
// node `from` (abscissa <= x0). cnt = 1 links i0 only. Equal abscissae: behind the nodes already there (visit order).
// One copy (the kernel's text has to stay small).
VLR_DEV_NOINLINE void r_link3(unsigned* node, const double* gx, int from, int cnt, int i0, int i1, int i2, double x0, double x1,
                              double x2) {
    int s = from;
    unsigned ws = node[s];
#pragma unroll 1
    for (int k = 0; k < cnt; ++k) {
        const int idx = k == 0 ? i0 : (k == 1 ? i1 : i2);
        const double x = k == 0 ? x0 : (k == 1 ? x1 : x2);
        const unsigned kx = r_key(x);
        for (;;) {
            const unsigned nx = ws & 0xffu;
            if (nx == (unsigned)R_NONE) break;
            const unsigned wn = node[nx];
            const unsigned kn = wn >> 8;
            if (kx < kn || (kx == kn && x < gx[nx])) break;
            s = (int)nx;
            ws = wn;
        }
        // idx goes between s and its successor
        const unsigned wi = (kx << 8) | (ws & 0xffu);
        node[idx] = wi;
        ws = (ws & ~0xffu) | (unsigned)idx;
        node[s] = ws;
        s = idx; // the next abscissa is not smaller: go on from here
        ws = wi;
    }
}

// ---------------------------------------------------------------------------------------------- leaf task
// One adaptive integration of the leaf sample's allele frequency over [a, b] with the parent fixed at task.parent_x
// (utils/adaptive_integration.rs:25-141; same visit order and decisions as wave_task_run, on (mantissa, exponent) values).
// The H lanes of the task run it in lockstep on identical values; lane h = 0 writes the grid rows and the task record.
//
// The visited points are kept in abscissa order on the fly: a singly linked list in the task's share of the octet's
// pool (r_link3). Every batch of points is linked in one walk that starts at a node known to lie below it: an
// iteration's [m1, middle, m2] at the bracket's left end, the closing points at a node left behind by the bracket
// and at the last middle. The closing trapezoid then needs no sort: each point knows its right neighbour (r_fin_list).
// All H lanes perform the same stores (a lane reads back what it wrote).
// `wmask`: the lanes of the WARP that run a task of this round (all of them are in here together). They meet again at
// the top of every step, so the evaluation of a step runs once for all tasks of the warp however differently their
// list walks went; a lane whose search is over waits there for the others.
VLR_DEV void r_task_run(const DevScenario* sc, const WavePlan& wp, const ROct& oc, RTask& t, const bool dead, double a, double b,
                        double* gx, double* gm, int* ge, unsigned* node, int cap, const WSplit sp, const unsigned wmask) {
    const int T = wp.T;
    const vlr_sample_t& smT = sc->samples[T];
    double rhoT = 1.0, iotaT = 0.0;
    if (smT.contamination_by >= 0) {
        rhoT = 1.0 - smT.contamination_fraction; // e^{purity}
        iotaT = 1.0 - rhoT;                      // e^{impurity} (likelihood.rs:77-84)
    }
    const double px = t.parent_x;
    const double vby = smT.contamination_by >= 0 ? px : 0.0;
    const double* qT = oc.q;
    const int nqx = oc.lc.nqTx, nqy = oc.lc.nqTy;
    const double res = smT.resolution;
    int n = 0, n_evals = 0;
    bool overflow = false, have_best = false, any_nan = false;
    MV best = mv_zero();
    double best_x = 0.0;

    bool sorted = true;
    // records the point as node n of the grid rows; returns the node index (or -1 when the grid is full)
    auto visit = [&](double x, const MV& v) -> int {
        n_evals++;
        if (v.m != v.m) any_nan = true;
        if (!have_best || mv_gt(v, best)) { // first maximum in visit order (calling.rs:851-870)
            have_best = true;
            best = v;
            best_x = x;
        }
        if (n >= W_GCAP) {
            overflow = true;
            return -1;
        }
        const int idx = n++;
        gx[idx] = x;
        if (sp.h == 0) {
            gm[idx] = v.m;
            ge[idx] = v.e;
        }
        if (idx >= cap) sorted = false; // the task's share of the pool is full: the closing trapezoid sorts (r_fin)
        return idx;
    };

    double left = a, right = b, middle = 0.0, first_middle = 0.0;
    MV f_left = mv_zero(), f_right = mv_zero(), f_first_m1 = mv_zero(), f_first_m2 = mv_zero();
    double x4 = 0.0, x5 = 0.0, x6 = 0.0;
    bool have_middle = false;
    int i_left = 0, i_right = 1, i_middle = 0, i_first_m1 = 0, i_first_m2 = 0; // nodes of the bracket ends, ...
    // a node at least 3 resolutions below the bracket's left end: where the walk of the closing points below the middle
    // starts (the left end only moves up and the final middle lies above it, so the node stays below middle - 3 res)
    int i_lo = 0;
    int step = 0;
    for (;;) {
#ifdef VLR_HOST_EMU
        if (step >= 4) break;
#else
        __syncwarp(wmask);
        if (!__any_sync(wmask, step < 4)) break;
        if (step >= 4) continue;
#endif
        double xs[3];
        int nvalid = 3;
        if (step == 0) {
            xs[0] = a;
            xs[1] = xs[2] = b;
            nvalid = 2;
        } else if (step == 1) {
            middle = (right + left) / 2.0;
            xs[0] = middle;
            xs[1] = (middle + left) / 2.0;
            xs[2] = (right + middle) / 2.0;
        } else if (step == 2) {
            // the midpoint of the arm the first iteration abandoned (:96-106) is bitwise that iteration's m1 or m2: its
            // value is remembered (the reference's HashMap deduplicates it); it still counts as a visit
            const bool upper = middle < first_middle;
            const double xa = upper ? (b + first_middle) / 2.0 : (first_middle + a) / 2.0;
            const int ia = visit(xa, upper ? f_first_m2 : f_first_m1);
            if (sorted && ia >= 0) r_link3(node, gx, upper ? i_first_m2 : i_first_m1, 1, ia, 0, 0, xa, 0.0, 0.0);
            const double lo = fmax(middle - (res * 3.0), a);
            const double slo = (middle - lo) / 3.0; // itertools-num linspace(lo, middle, 4).take(3)
            xs[0] = lo + slo * 0.0;
            xs[1] = lo + slo * 1.0;
            xs[2] = lo + slo * 2.0;
            const double hi = fmin(middle + (res * 3.0), b);
            const double shi = (hi - middle) / 3.0; // linspace(middle, hi, 4).skip(1)
            x4 = middle + shi * 1.0;
            x5 = middle + shi * 2.0;
            x6 = middle + shi * 3.0;
        } else {
            xs[0] = x4;
            xs[1] = x5;
            xs[2] = x6;
        }
        MV v[3];
        if (dead) {
            v[0] = v[1] = v[2] = mv_zero();
        } else {
            double X[3], Y[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const WArgs w = wave_args(rhoT, iotaT, xs[k], vby);
                X[k] = w.xu;
                Y[k] = w.Yp;
            }
            r_pileup<3, true>(qT, nqx, nqy, X, Y, nvalid, v, sp);
        }
        const int i0 = visit(xs[0], v[0]);
        const int i1 = visit(xs[1], v[1]);
        const int i2 = nvalid > 2 ? visit(xs[2], v[2]) : -1;
        if (step == 0) { // nodes 0 and 1: the list is [a, b] (a < b: the integration is wider than the resolution)
            if (cap >= 2) {
                node[0] = (r_key(xs[0]) << 8) | 1u;
                node[1] = (r_key(xs[1]) << 8) | (unsigned)R_NONE;
            } else {
                sorted = false;
            }
        } else if (sorted && i2 >= 0) {
            // step 1: m1 < middle < m2, inside the bracket: from its left end; step 2: the three closing points up to the
            // middle, ascending, from a node below them; step 3: those above the middle, from the middle
            if (step == 1) r_link3(node, gx, i_left, 3, i1, i0, i2, xs[1], xs[0], xs[2]);
            else r_link3(node, gx, step == 2 ? i_lo : i_middle, 3, i0, i1, i2, xs[0], xs[1], xs[2]);
        }
        if (step == 0) {
            f_left = v[0];
            f_right = v[1];
        } else if (step == 1) {
            if (!have_middle) {
                first_middle = middle;
                f_first_m1 = v[1];
                f_first_m2 = v[2];
                i_first_m1 = i1;
                i_first_m2 = i2;
            }
            have_middle = true;
            i_middle = i0;
            const double m1 = xs[1], m2 = xs[2];
            int idx = 0;
            MV fb = f_left;
            if (mv_gt(v[1], fb)) {
                idx = 1;
                fb = v[1];
            }
            if (mv_gt(v[2], fb)) {
                idx = 2;
                fb = v[2];
            }
            if (mv_gt(f_right, fb)) idx = 3;
            // neighbours of the argmax in [left, m1, m2, right] become the new bounds (the middle is not a candidate)
            const double nl = idx <= 1 ? left : (idx == 2 ? m1 : m2);
            const MV nfl = idx <= 1 ? f_left : (idx == 2 ? v[1] : v[2]);
            const double nr = idx == 0 ? m1 : (idx == 1 ? m2 : right);
            const MV nfr = idx == 0 ? v[1] : (idx == 1 ? v[2] : f_right);
            const int il = idx <= 1 ? i_left : (idx == 2 ? i1 : i2), ir = idx == 0 ? i1 : (idx == 1 ? i2 : i_right);
            if (idx >= 2 && left <= nl - (res * 3.0)) i_lo = i_left; // (left: still the old left end here)
            left = nl;
            f_left = nfl;
            right = nr;
            f_right = nfr;
            i_left = il;
            i_right = ir;
            if (il < 0 || ir < 0) sorted = false; // (grid overflow)
        }
        if (step <= 1) {
            // while (((right - left) >= res && left < right) || middle.is_none())   (adaptive_integration.rs:52)
            step = (!overflow && ((((right - left) >= res) && left < right) || !have_middle)) ? 1 : 2;
        } else {
            step++;
        }
    }
    if (sp.h != 0) return;
    uint32_t status = 0;
    if (any_nan) status |= VLR_ST_NAN;
    if (overflow) status |= VLR_ST_GRID_OVERFLOW;
    t.best = best;
    t.best_x = best_x;
    t.n_evals = n_evals;
    t.n_grid = n;
    t.status = status;
    t.sorted = (sorted && !overflow) ? 1 : 0;
}

// ln of the constant part + value of a point: prior + (ln L_parent + ((ln m + e ln 2) + sum of the read scales)),
// associated like wave_task_run did
VLR_DEV double r_point_ln(const RTask& t, double ksumT, const MV& v) {
    if (t.dead) {
        const double lnl = (ksumT != ksumT || ksumT == neg_inf()) ? ksumT : 0.0;
        return t.prior_const + (t.lh_const + lnl);
    }
    return t.prior_const + (t.lh_const + (mv_ln(v) + ksumT));
}

// ---------------------------------------------------------------------------------------------- closing trapezoid
// ln_trapezoidal_integrate_grid_exp over the n visited points of a task (rust-bio), in linear space
// relative to the maximum: weights m_i / m_max * 2^(e_i - e_max) without exp or log, rank sort by (x, visit order),
// ln( sum_i (w_i + w_i+1) / 2 * (x_i+1 - x_i) ) + ln(max). Equal abscissae keep visit order (zero-width intervals).
VLR_DEV_NOINLINE double r_fin(const double* gx, const double* gm, const int* ge, int n, const MV best, double fmax, double* sx,
                              double* se, short* inv, const WGroup grp) {
    if (fmax != fmax) return NAN;
    if (n < 2 || fmax == neg_inf()) return neg_inf();
    const double inv_m = 1.0 / best.m;
    grp_sync(grp);
    for (int a = grp.lane; a < n; a += grp.n) {
        const double m = gm[a];
        const int de = ge[a] - best.e; // <= 0
        double w = 0.0;
        if (m != 0.0 && de > -1000) w = (m * inv_m) * d_make((1023 + de) << 20, 0);
        sx[a] = gx[a];
        se[a] = w;
    }
    grp_sync(grp);
    for (int a = grp.lane; a < n; a += grp.n) {
        const double xi = sx[a];
        int rank = 0;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const double xj = sx[j];
            rank += (xj < xi) || (xj == xi && j < a);
        }
        inv[rank] = (short)a;
    }
    grp_sync(grp);
    double sum = 0.0;
    for (int p = grp.lane; p + 1 < n; p += grp.n) {
        const int i0 = inv[p], i1 = inv[p + 1];
        sum += (se[i0] + se[i1]) * (sx[i1] - sx[i0]);
    }
    sum = grp_sum_d(sum, grp);
    grp_sync(grp);
    return fmax + m_log(sum * 0.5);
}

// The same for a task whose points are linked in abscissa order (r_task_run): every point knows its right neighbour,
// so the interval terms are independent of each other: lanes over points, no sort.
VLR_DEV_NOINLINE double r_fin_list(const double* gx, const double* gm, const int* ge, int n, const MV best, double fmax,
                                   const unsigned* node, const WGroup grp) {
    if (fmax != fmax) return NAN;
    if (n < 2 || fmax == neg_inf()) return neg_inf();
    const double inv_m = 1.0 / best.m;
    grp_sync(grp);
    double sum = 0.0;
#pragma unroll 2
    for (int i = grp.lane; i < n; i += grp.n) { // (independent loads: several points in flight)
        int j = (int)(node[i] & 0xffu);
        const bool last = j == R_NONE;
        if (last) j = i;
        const double mi = gm[i], mj = gm[j], xi = gx[i], xj = gx[j];
        const int di = ge[i] - best.e, dj = ge[j] - best.e; // <= 0
        double wi = 0.0, wj = 0.0;
        if (mi != 0.0 && di > -1000) wi = (mi * inv_m) * d_make((1023 + di) << 20, 0);
        if (mj != 0.0 && dj > -1000) wj = (mj * inv_m) * d_make((1023 + dj) << 20, 0);
        if (!last) sum += (wi + wj) * (xj - xi);
    }
    sum = grp_sum_d(sum, grp);
    grp_sync(grp);
    return fmax + m_log(sum * 0.5);
}

// the same over ln values (the outer grid holds the task integrals): wave_fin_coop with explicit scratch rows
VLR_DEV_NOINLINE double r_fin_ln(const double* x, const double* f, int n, double fmx, bool any_nan, double* sx, double* se,
                                 short* inv, const WGroup grp) {
    if (any_nan) return NAN;
    if (n < 2 || fmx == neg_inf()) return neg_inf();
    grp_sync(grp);
    for (int a = grp.lane; a < n; a += grp.n) {
        const double xa = x[a], fa = f[a];
        sx[a] = xa;
        se[a] = m_exp(fa - fmx);
    }
    grp_sync(grp);
    for (int a = grp.lane; a < n; a += grp.n) {
        const double xi = sx[a];
        int rank = 0;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const double xj = sx[j];
            rank += (xj < xi) || (xj == xi && j < a);
        }
        inv[rank] = (short)a;
    }
    grp_sync(grp);
    double sum = 0.0;
    for (int p = grp.lane; p + 1 < n; p += grp.n) {
        const int i0 = inv[p], i1 = inv[p + 1];
        sum += (se[i0] + se[i1]) * (sx[i1] - sx[i0]);
    }
    sum = grp_sum_d(sum, grp);
    grp_sync(grp);
    return fmx + m_log(sum * 0.5);
}

// ---------------------------------------------------------------------------------------------- round of an lc
// Round-0 tasks of an lc into the octet's task array (what wave_lc_init writes to the global task list for the other
// lcs). Uniform over the group; lane 0 writes.
VLR_DEV int r_first_tasks(const DevScenario* sc, const WavePlan& wp, const WaveBufs& wb, int lci, RTask* tasks, const WGroup grp) {
    WaveLC& lc = wb.lcs[lci];
    const WaveLocus& wl = wb.loci[lc.li];
    const int E = sc->E, P = wp.P;
    int k = 0;
    for (int e = 0; e < E; ++e) {
        if (lc.ci > 0 && !sc->events[e].has_artifact_twin) continue;
        if (wl.ev_kind[e] == 1) {
            if (grp.lane == 0) {
                tasks[k].parent_x = sc->set_vafs[sc->nodes[wp.root_node[e]].vaf_offset];
                tasks[k].event = e;
                tasks[k].parent_disc = 1;
            }
            k++;
        } else if (wl.ev_kind[e] == 3) {
            Adaptive st;
            st.init(wl.pa, wl.pb, sc->samples[P].resolution);
            double xs[8];
            const int np = st.points(xs); // [min, max]
            for (int i = 0; i < np; ++i) {
                if (grp.lane == 0) {
                    lc.outer_xs[i] = xs[i];
                    tasks[k].parent_x = xs[i];
                    tasks[k].event = e;
                    tasks[k].parent_disc = 0;
                }
                k++;
            }
            if (grp.lane == 0) {
                lc.outer = st;
                lc.outer_pending = 1;
                lc.ta = wl.ev_a[e];
                lc.tb = wl.ev_b[e];
            }
        }
    }
    grp_sync(grp);
    return k;
}

// Phase 1 of a round, per task (H lanes): the parent pileup at the task's parent_x (a constant of the leaf integration:
// GenericLikelihood::compute, generic.rs:511-551, hits its per-sample cache for it at every point), then the search.
VLR_DEV void r_task(const DevScenario* sc, const WavePlan& wp, const WaveBufs& wb, ROct& oc, int q, int cnt, double* gx, double* gm,
                    int* ge, const WSplit sp, const unsigned wmask) {
    RTask& t = oc.task[q];
    const WaveLocus& wl = wb.loci[oc.lc.li];
    const double px = t.parent_x;
    double lh = 0.0;
    if (oc.lc.nqPx + oc.lc.nqPy > 0) {
        const WArgs ap = wave_args(1.0, 0.0, px, 0.0);
        MV v;
        r_pileup<1, false>(oc.lc.qP, oc.lc.nqPx, oc.lc.nqPy, &ap.xu, &ap.Yp, 1, &v, sp); // (from the arena)
        lh = v.m != v.m ? NAN : (mv_ln(v) + oc.lc.ksumP);
    }
    const double prior = wave_prior_ok(sc, wp.P, px) ? 0.0 : neg_inf();
    const double ksumT = oc.lc.ksumT;
    const double cpart = prior + lh;
    const bool dead = !(cpart > neg_inf()) || ksumT != ksumT || ksumT == neg_inf(); // ln 0 or NaN whatever the point
    if (sp.h == 0) {
        t.lh_const = lh;
        t.prior_const = prior;
        t.dead = dead ? 1 : 0;
    }
    const double a = t.parent_disc ? wl.ev_a[t.event] : wb.lcs[oc.lc.lci].ta;
    const double b = t.parent_disc ? wl.ev_b[t.event] : wb.lcs[oc.lc.lci].tb;
    const int cap = R_POOL_N / cnt; // the task's share of the octet's list pool
    unsigned* node = reinterpret_cast<unsigned*>(oc.pool) + (size_t)q * cap;
    r_task_run(sc, wp, oc, t, dead, a, b, gx, gm, ge, node, cap < W_GCAP ? cap : W_GCAP, sp, wmask);
}

// Phase 2 of a round (the whole group): integrate each task's grid, MAP bookkeeping in visit order (calling.rs:851-870),
// event densities, base-event log for the AFD, the next step of the enclosing integration over the parent's allele
// frequency and its tasks for the next round. Returns the number of tasks of the next round (0: the lc is complete).
// Row i of (rows_x, rows_m, rows_e) is the grid of task i. Control flow is uniform over the group; lane 0 writes.
VLR_DEV_NOINLINE int r_advance(const DevScenario* sc, const WavePlan& wp, const WaveBufs& wb, ROct& oc, int round, int cnt,
                               const double* rows_x, const double* rows_m, const int* rows_e, double* big_scratch, bool want_be,
                               const WGroup grp) {
    const int lci = oc.lc.lci;
    WaveLC& lc = wb.lcs[lci];
    const int P = wp.P, T = wp.T;
    const bool l0 = grp.lane == 0;
    const double ksumT = oc.lc.ksumT;
    uint32_t status = lc.status, n_base = lc.n_base;
    double ofs[8];
    int no = 0;
    if (lc.outer_skip) { // the batch's first abscissa had no task: its integral is known from the first iteration
        ofs[0] = lc.om_val[lc.outer_skip - 1];
        n_base += lc.om_nev[lc.outer_skip - 1];
        no = 1;
    }
    const bool first_iteration = lc.outer_pending && lc.outer.phase == 1 && !lc.outer.have_middle;
    for (int i = 0; i < cnt; ++i) {
        const RTask t = oc.task[i];
        const int e = t.event;
        const double* gx = rows_x + (size_t)i * W_GCAP;
        const double* gm = rows_m + (size_t)i * W_GCAP;
        const int* ge = rows_e + (size_t)i * W_GCAP;
        const bool nan = (t.status & VLR_ST_NAN) != 0;
        const double fmax = nan ? NAN : r_point_ln(t, ksumT, t.best);
        double value;
        if (t.sorted) {
            const int cap = R_POOL_N / cnt;
            value = r_fin_list(gx, gm, ge, t.n_grid, t.best, fmax, reinterpret_cast<const unsigned*>(oc.pool) + (size_t)i * cap, grp);
        } else { // rare: more points than the task's share of the pool, or a full grid: sort in the global scratch
            value = r_fin(gx, gm, ge, t.n_grid, t.best, fmax, big_scratch, big_scratch + W_GCAP,
                          reinterpret_cast<short*>(big_scratch + 2 * W_GCAP), grp);
        }
        status |= t.status;
        if (fmax != fmax) status |= VLR_ST_NAN;
        n_base += (uint32_t)t.n_evals;
        if (want_be && oc.lc.ci == 0) { // base events of the artifact-free config feed the AFD (calling.rs:891-928)
            unsigned base = 0;
            if (l0) base = wa_add_u32(&wb.be_n[oc.lc.li], (unsigned)t.n_grid);
            base = grp_bcast_u(base, grp);
            double* be = wb.be + (size_t)oc.lc.li * BE_CAP * 4;
            const double disc = d_make(0, (int)((t.parent_disc ? 1u : 0u) << P));
            if (base + (unsigned)t.n_grid > (unsigned)BE_CAP) status |= VLR_ST_BASE_EVENTS_OVERFLOW;
            for (int k = grp.lane; k < t.n_grid; k += grp.n) {
                const unsigned at = base + (unsigned)k;
                if (at >= (unsigned)BE_CAP) break;
                double* r = be + (size_t)at * 4;
                r[0] = r_point_ln(t, ksumT, MV{gm[k], ge[k]});
                r[1] = disc;
                r[2 + P] = t.parent_x;
                r[2 + T] = gx[k];
            }
        }
        if (l0 && t.n_evals > 0 && (!lc.map_set[e] || fmax > lc.map_joint[e])) {
            lc.map_set[e] = 1;
            lc.map_joint[e] = fmax;
            lc.map_vp[e] = t.parent_x;
            lc.map_vt[e] = t.best_x;
            lc.map_disc[e] = t.parent_disc ? 1 : 0;
        }
        grp_sync(grp);
        if (e == wp.outer_event) {
            if (no < 8) ofs[no] = value;
            if (first_iteration && (no == 1 || no == 2) && l0) { // [middle, m1, m2]
                lc.om_x[no - 1] = t.parent_x;
                lc.om_val[no - 1] = value;
                lc.om_nev[no - 1] = (uint32_t)t.n_evals;
            }
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
        grp_sync(grp);
        return 0;
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
    grp_sync(grp); // every lane has read the lc before lane 0 updates it
    const bool more = st.consume(oxs, ofs, outer_overflow != 0);
    int next = 0;
    if (more && round + 1 < W_MAXROUNDS) {
        double xs[8];
        const int k = st.points(xs);
        int skip = 0; // the closing batch starts with the abandoned arm's midpoint = the first iteration's m1 or m2
        if (k == 7) skip = xs[0] == lc.om_x[0] ? 1 : (xs[0] == lc.om_x[1] ? 2 : 0);
        const int first = skip ? 1 : 0;
        grp_sync(grp);
        if (l0) {
            lc.outer = st;
            lc.outer_skip = skip;
            for (int i = 0; i < k; ++i) lc.outer_xs[i] = xs[i];
            for (int i = first; i < k; ++i) {
                oc.task[i - first].parent_x = xs[i];
                oc.task[i - first].event = wp.outer_event;
                oc.task[i - first].parent_disc = 0;
            }
            lc.outer_n = outer_n;
            lc.outer_overflow = outer_overflow;
            lc.status = status;
            lc.n_base = n_base;
        }
        next = k - first;
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
        // (the task lists in the pool are dead by now)
        double *sx = oc.pool, *se = oc.pool + R_SORT;
        short* inv = reinterpret_cast<short*>(oc.pool + 2 * R_SORT);
        if (outer_n > R_SORT) {
            sx = big_scratch;
            se = big_scratch + W_GCAP;
            inv = reinterpret_cast<short*>(big_scratch + 2 * W_GCAP);
        }
        const double d = r_fin_ln(ox, of, outer_n, fmx, any_nan, sx, se, inv, grp);
        if (l0) {
            lc.outer = st;
            lc.outer_pending = 0;
            lc.outer_skip = 0;
            lc.outer_n = outer_n;
            lc.outer_overflow = outer_overflow;
            lc.dens[wp.outer_event] = d;
            lc.status = status;
            lc.n_base = n_base;
            lc.task_count = 0;
        }
    }
    grp_sync(grp);
    return next;
}
