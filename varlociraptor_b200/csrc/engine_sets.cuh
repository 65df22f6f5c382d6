// engine_sets.cuh — the all-Set pipeline: scenarios whose event trees consist of Set nodes only (pedigrees: config 3 of
// BASELINE.json). Included at the end of engine_core.cuh, inside the variant namespace, when VLR_VAR_SETS is defined.
//
// GenericPosterior::density (generic.rs:294-330) over Set nodes is a sum of joint probabilities of discrete allele
// frequency combinations; GenericLikelihood::compute (generic.rs:500-554) caches the pileup likelihood per (sample,
// allele frequency), Prior::compute per combination (prior.rs:718-736). So per (locus, artifact config) the model is:
//   1. per-read coefficients of every sample (read_coefficients: the transcendental part),
//   2. one pileup evaluation per FOLD (distinct sample x allele frequency: 9 for a trio),
//   3. per LEAF (root-to-leaf combination of an event tree, 39 for the trio's three events)
//      joint = prior[leaf] + sum over samples of fold values — table look-ups,
//   4. per event ln_sum_exp over its leaves, MAP = first maximum in visit order.
// The generic engine does the same through its tree interpreter (warp-uniform recursion, operand stack, caches) and
// spends most of its time fetching instructions; here trees and priors are flattened once per context (SetsPlan,
// scenario_prep.h) and the per-locus work is three small kernels like the wavefront pipeline's:
//   sets_pre_locus  warp per locus : bias pre-pass (locus_prepass), lc allocation
//   sets_lc         warp per lc    : steps 1-4 for one artifact config, coefficients in shared memory
//   sets_finish     warp per locus : event posteriors, artifact probability, MAP, AFD (locus_tail)
// Loci the pipeline does not serve (per-record prior overrides, more reads than the shared-memory arena holds) are
// deferred to the generic engine.

constexpr int SETS_SM_READS = 320; // per-warp coefficient arena in shared memory (10 KB): a trio at depth 100

struct SetsCounters {
    unsigned long long ticket[4]; // pre, lc, finish, deferred (generic kernel)
    unsigned int n_lc, n_deferred;
};

struct SetsLocus {
    int lc_base, n_cfg; // n_cfg = 0: deferred to the generic engine
    int n_twins;
    uint32_t status, lf;
    int has_alt_loci, coef_total, clear_mask;
    int n_obs[MAXS], s_one[MAXS], s_gt1[MAXS], coef_off[MAXS];
    int surviving[NCFG];
    int64_t singleton_row;
    double forward_rate, ln_fwd, ln_rev;
};

struct SetsLC { // one (locus, artifact config)
    int li, ci, art_id;
    uint32_t status, n_base;
    int map_leaf[MAXE]; // MAP leaf of the event under this config, or -1
    double dens[MAXE], map_joint[MAXE];
};

struct SetsBufs {
    SetsCounters* cnt;
    SetsLocus* loci;
    SetsLC* lcs;
    int* deferred;   // absolute locus indices
    double* be;      // base-event log per locus of the sub-chunk (AFD only): [n_sub][SETS_MAXL][2 + S]
    unsigned* be_n;  // [n_sub]
    int lc_cap;
};

// Prior of every leaf for variant type class `vt`, published for all warps of the context. Several warps may compute it
// at the same time (same values); readers only trust it once the state says so.
VLR_DEV_NOINLINE void sets_fill_prior(Ctx& c_, const SetsPlan& sp, int vt) {
    Ctx& c = warp_ctx(c_);
    const int S = c.sc->S;
    const uint32_t s0 = c.status;
    for (int l = 0; l < sp.n_leaves; ++l) {
        Ops ev; // (a local: the kernel's shared-memory Ctx is lean, without the operand stack)
        for (int s = 0; s < MAXS; ++s) ev.vaf[s] = s < S ? sp.leaf_vaf[(size_t)l * S + s] : 0.0;
        ev.set_mask = (1u << S) - 1u;
        ev.disc_mask = sp.leaves[l].discmask;
        ev.lfc_mask = 0;
        c.status = 0;
        double p;
        if (c.sc->all_uniform) { // flat prior inside every sample's universe (prior.rs:385-406, joint() of the generic engine)
            p = 0.0;
            for (int s = 0; s < S; ++s)
                if ((c.sc->samples[s].ploidy == 0 && ev.vaf[s] != 0.0) || !universe_contains(c, s, ev.vaf[s])) p = neg_inf();
        } else {
            p = prior_compute_uncached(c, ev);
        }
        if (lane_id() == 0) {
            sp.prior_val[(size_t)vt * sp.n_leaves + l] = p;
            sp.prior_side[(size_t)vt * sp.n_leaves + l] = c.status;
        }
    }
    c.status = s0;
    warp_sync();
    if (lane_id() == 0) {
#ifdef VLR_HOST_EMU
        sp.prior_state[vt] = 2;
#else
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(sp.prior_state + vt), "r"(2) : "memory");
#endif
    }
    warp_sync();
}

VLR_DEV bool sets_prior_ready(const SetsPlan& sp, int vt) {
#ifdef VLR_HOST_EMU
    return sp.prior_state[vt] == 2;
#else
    int st = 0;
    if (lane_id() == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(st) : "l"(sp.prior_state + vt) : "memory");
    st = __shfl_sync(FULL, st, 0, LANES);
    return st == 2;
#endif
}

// ---------------------------------------------------------------------------------------------- pre (warp per locus)
// calling.rs:586-626 + bias/*.rs (locus_prepass) and the lc allocation.
VLR_DEV void sets_pre_locus(const DevScenario* sc, const DevBatch* b, const SetsBufs& sb, int64_t locus, int li, bool want_be,
                            Ctx& c) {
    const int S = sc->S;
    c.sc = sc;
    c.b = b;
    c.locus = locus;
    c.status = 0;
    c.lf = b->lflags[locus];
    BiasPlan plan;
    locus_prepass(c, plan);
    SetsLocus& wl = sb.loci[li];
    bool defer = c.coef_total > SETS_SM_READS;
    if (b->het_phred) { // variant-specific priors (calling.rs:472-494): not in the context's prior table
        const float h = b->het_phred[locus];
        if (!(h != h)) defer = true;
    }
    if (b->semr_phred) {
        const float h = b->semr_phred[locus];
        if (!(h != h)) defer = true;
    }
    const int n_cfg = 1 + plan.n_surviving;
    int lc_base = 0;
    if (!defer) {
        unsigned lb = 0;
        if (lane_id() == 0) lb = wa_add_u32(&sb.cnt->n_lc, (unsigned)n_cfg);
#ifndef VLR_HOST_EMU
        lb = __shfl_sync(FULL, lb, 0, LANES);
#endif
        lc_base = (int)lb;
        if ((int64_t)lb + n_cfg > (int64_t)sb.lc_cap) {
            defer = true;
            if (lane_id() == 0)
                for (int ci = 0; ci < n_cfg; ++ci)
                    if ((int64_t)lc_base + ci < (int64_t)sb.lc_cap) sb.lcs[lc_base + ci].li = -1; // dead
        }
    }
    if (lane_id() != 0) return;
    if (defer) {
        const unsigned d = wa_add_u32(&sb.cnt->n_deferred, 1u);
        sb.deferred[d] = (int)locus;
        wl.lc_base = 0;
        wl.n_cfg = 0;
        wl.n_twins = 0;
        wl.status = 0;
        return;
    }
    wl.lc_base = lc_base;
    wl.n_cfg = n_cfg;
    wl.n_twins = plan.n_twins;
    wl.status = c.status;
    wl.lf = c.lf;
    wl.has_alt_loci = plan.has_alt_loci ? 1 : 0;
    wl.coef_total = c.coef_total;
    wl.singleton_row = c.singleton_row;
    wl.forward_rate = plan.forward_rate;
    wl.ln_fwd = plan.ln_fwd;
    wl.ln_rev = plan.ln_rev;
    int clear = 0;
    for (int s = 0; s < S; ++s) {
        wl.n_obs[s] = c.n_obs[s];
        wl.s_one[s] = c.s_one[s];
        wl.s_gt1[s] = c.s_gt1[s];
        wl.coef_off[s] = c.coef_off[s];
        if (c.clear_ref[s]) clear |= 1 << s;
    }
    wl.clear_mask = clear;
    for (int k = 0; k < NCFG; ++k) wl.surviving[k] = k < plan.n_surviving ? plan.surviving[k] : 0;
    if (want_be) sb.be_n[li] = 0;
    for (int ci = 0; ci < n_cfg; ++ci) {
        SetsLC& lc = sb.lcs[lc_base + ci];
        lc.li = li;
        lc.ci = ci;
    }
}

// ---------------------------------------------------------------------------------------------- lc (warp per lc)
// `coef`: the warp's coefficient arena (SETS_SM_READS reads), `ll`: SETS_MAXF doubles of warp-private scratch.
// The warps of a CTA run the three phases of their lcs together (`phased`: a CTA barrier between coefficients, folds
// and leaves; `lci` < 0: this warp has no lc in this iteration but keeps the barriers): the kernel's text is 170 KB and
// the phases share almost none of it, so warps in different phases evicted each other's instructions
// (stall_no_instruction 3.4 cycles per issue before, profiles/ncu_r2b_sets_lc_summary.csv).
#ifdef VLR_HOST_EMU
VLR_DEV void sets_phase_sync() {}
#else
VLR_DEV void sets_phase_sync() { __syncthreads(); }
#endif
VLR_DEV void sets_lc(const DevScenario* sc, const DevBatch* b, const SetsPlan& sp, const SetsBufs& sb, int lci, int64_t sub_lo,
                     bool want_be, Ctx& c, double* coef, double* ll, bool phased = false, MemoTab* memo = nullptr) {
    const bool live = lci >= 0 && sb.lcs[lci].li >= 0; // (li < 0: dead)
    SetsLC& lc = sb.lcs[live ? lci : 0];
    const int li = live ? lc.li : 0, ci = lc.ci;
    const SetsLocus& wl = sb.loci[li];
    const int S = sc->S, E = sc->E;
    if (live) {
    c.sc = sc;
    c.b = b;
    c.locus = sub_lo + li;
    c.status = 0;
    c.lf = wl.lf;
    c.vartype = (wl.lf >> VLR_LF_VARTYPE_SHIFT) & 3;
    c.het_override = NAN;
    c.semr_override = NAN;
    c.prior_absent = NAN;
    c.pc_n = 0;
    c.singleton_row = wl.singleton_row;
    c.n_pileup_evals = 0;
    c.art.id = ci == 0 ? 0 : wl.surviving[ci - 1];
    c.art.forward_rate = wl.forward_rate;
    c.art.ln_fwd = wl.ln_fwd;
    c.art.ln_rev = wl.ln_rev;
    c.art.has_alt_loci = wl.has_alt_loci != 0;
    c.coef = coef;
    c.coef_in_sm = 1;
    c.coef_cap = SETS_SM_READS;
    c.coef_total = wl.coef_total;
    for (int s = 0; s < S; ++s) {
        c.n_obs[s] = wl.n_obs[s];
        c.s_one[s] = wl.s_one[s];
        c.s_gt1[s] = wl.s_gt1[s];
        c.coef_off[s] = wl.coef_off[s];
    }
    warp_sync();
    if (!sets_prior_ready(sp, c.vartype)) sets_fill_prior(c, sp, c.vartype);
    for (int s = 0; s < S; ++s) read_coefficients(c, s, memo);
    }
    if (phased) sets_phase_sync();
    // ---- folds: the pileup ln-likelihoods the leaves look up
    if (live) {
        for (int f = 0; f < sp.n_folds; ++f) {
            const SetsFold fd = sp.folds[f];
            const double v = sample_likelihood_call(c, fd.sample, fd.vaf, fd.vaf_by);
            if (lane_id() == 0) ll[f] = v;
        }
    }
    warp_sync();
    if (phased) sets_phase_sync();
    if (!live) return;
    // ---- leaves and events
    const double* prior = sp.prior_val + (size_t)c.vartype * sp.n_leaves;
    const uint32_t* pside = sp.prior_side + (size_t)c.vartype * sp.n_leaves;
    const unsigned clear = (unsigned)wl.clear_mask;
    const bool log_be = want_be && ci == 0;
    double* be = log_be ? sb.be + (size_t)li * SETS_MAXL * (2 + S) : nullptr;
    uint32_t status = c.status, n_base = 0;
    unsigned n_rec = 0;
    for (int e = 0; e < E; ++e) {
        double dens = neg_inf(), map_joint = 0.0;
        int map_leaf = -1;
        if (!(ci > 0 && !sc->events[e].has_artifact_twin)) {
            const int first = sp.ev_first[e], count = sp.ev_count[e];
            double best_j = neg_inf(), mx = neg_inf();
            int best_l = -1, n_eval = 0;
            bool any_nan = false;
            uint32_t side = 0;
            // pass 1: joints of this lane's leaves (kept for pass 2: at most SETS_MAXL / LANES per lane)
            constexpr int PER = (SETS_MAXL + LANES - 1) / LANES;
            double js[PER];
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int k = lane_id() + q * LANES;
                double j = neg_inf();
                bool evaluated = false;
                if (k < count) {
                    const int l = first + k;
                    const SetsLeaf lf = sp.leaves[l];
                    if (!(lf.posmask & clear)) { // (a clearly reference sample cuts the paths through all-positive Set nodes)
                        double lh = 0.0;
                        for (int s = 0; s < S; ++s) lh += ll[lf.fold[s]]; // sample-index order (generic.rs:511-551)
                        j = prior[l] + lh;
                        side |= pside[l];
                        evaluated = true;
                        n_eval++;
                        if (j != j) any_nan = true;
                        if (best_l < 0 || j > best_j) { // first maximum in visit order (calling.rs:851-870)
                            best_j = j;
                            best_l = l;
                        }
                        if (j > mx) mx = j;
                    }
                }
                js[q] = j;
                if (log_be) { // base events of the artifact-free config feed the AFD (calling.rs:891-928)
#ifdef VLR_HOST_EMU
                    const int pos = (int)n_rec, tot = evaluated ? 1 : 0;
#else
                    const unsigned m = w_ballot(evaluated);
                    const int pos = (int)n_rec + __popc(m & ((1u << lane_id()) - 1u)), tot = __popc(m);
#endif
                    if (evaluated) {
                        const int l = first + k;
                        double* r = be + (size_t)pos * (2 + S);
                        r[0] = j;
                        r[1] = d_make(0, (int)sp.leaves[l].discmask);
                        for (int s = 0; s < S; ++s) r[2 + s] = sp.leaf_vaf[(size_t)l * S + s];
                    }
                    n_rec += (unsigned)tot;
                }
            }
            n_eval = w_sum_i(n_eval);
            n_base += (uint32_t)n_eval;
            side = w_or_u(side);
            status |= side;
            any_nan = w_any(any_nan);
            if (any_nan) status |= VLR_ST_NAN;
            if (n_eval > 0) {
                if (any_nan) {
                    // rare: mirror the sequential bookkeeping (a NaN joint poisons the sum and blocks later MAP candidates)
                    dens = NAN;
                    double bj = 0.0;
                    int bl = -1;
                    for (int k = 0; k < count; ++k) {
                        const int l = first + k;
                        const SetsLeaf lf = sp.leaves[l];
                        if (lf.posmask & clear) continue;
                        double lh = 0.0;
                        for (int s = 0; s < S; ++s) lh += ll[lf.fold[s]];
                        const double j = prior[l] + lh;
                        if (bl < 0 || j > bj) {
                            bj = j;
                            bl = l;
                        }
                    }
                    map_joint = bj;
                    map_leaf = bl;
                } else {
                    mx = w_max_d(mx);
                    // first maximum in visit order: larger joint, then smaller leaf index
#ifndef VLR_HOST_EMU
                    for (int o = LANES / 2; o > 0; o >>= 1) {
                        const double oj = __shfl_xor_sync(FULL, best_j, o);
                        const int ol = __shfl_xor_sync(FULL, best_l, o);
                        if (ol >= 0 && (best_l < 0 || oj > best_j || (oj == best_j && ol < best_l))) {
                            best_j = oj;
                            best_l = ol;
                        }
                    }
#endif
                    map_joint = best_j;
                    map_leaf = best_l;
                    if (mx == neg_inf() || mx == INFINITY) {
                        dens = mx;
                    } else {
                        double sum = 0.0;
#pragma unroll
                        for (int q = 0; q < PER; ++q)
                            if (js[q] != neg_inf()) sum += m_exp(js[q] - mx);
                        sum = w_sum_d(sum);
                        dens = mx + m_log1p(sum - 1.0);
                    }
                }
            }
        }
        if (lane_id() == 0) {
            lc.dens[e] = dens;
            lc.map_joint[e] = map_joint;
            lc.map_leaf[e] = map_leaf;
        }
    }
    if (lane_id() == 0) {
        lc.art_id = c.art.id;
        lc.status = status;
        lc.n_base = n_base;
        if (log_be) sb.be_n[li] = n_rec;
    }
    warp_sync();
}

// ---------------------------------------------------------------------------------------------- finish (warp per locus)
// rust-bio Model::compute's event loop + GenericPosterior::compute (generic.rs:430-460) over the lcs of the locus, then
// call_record / sample_infos (calling.rs:720-937): locus_tail.
VLR_DEV void sets_finish_locus(const DevScenario* sc, const DevBatch* b, const DevResults* res, const SetsPlan& sp,
                               const SetsBufs& sb, WarpWs* ws, int64_t locus, int li, Ctx& c) {
    const SetsLocus& wl = sb.loci[li];
    if (wl.n_cfg == 0) return; // deferred: the generic engine writes this locus
    const int S = sc->S, E = sc->E;
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
            c.be = sb.be + (size_t)li * SETS_MAXL * (2 + S);
            c.n_rec = sb.be_n[li];
        }
        for (int i = 0; i < 2 * E; ++i) c.map_set[i] = 0;
        for (int e = 0; e < E; ++e) {
            c.ev_plain[e].init();
            c.ev_twin[e].init();
        }
        const double twin_prior = wl.n_twins > 0 ? LN_05 + m_log(1.0 / (double)wl.n_twins) : neg_inf();
        uint32_t n_base = 0;
        for (int ci = 0; ci < wl.n_cfg; ++ci) {
            const SetsLC& lc = sb.lcs[wl.lc_base + ci];
            c.status |= lc.status;
            n_base += lc.n_base;
            for (int e = 0; e < E; ++e) {
                if (ci > 0 && !sc->events[e].has_artifact_twin) continue;
                const double d = lc.dens[e];
                if (d != d) c.status |= VLR_ST_NAN;
                if (ci == 0) c.ev_plain[e].add(LN_05 + d);
                else c.ev_twin[e].add(twin_prior + d);
                const int slot = 2 * e + (ci > 0 ? 1 : 0);
                const int l = lc.map_leaf[e];
                if (l >= 0 && (!c.map_set[slot] || lc.map_joint[e] > c.map_joint[slot])) {
                    c.map_set[slot] = 1;
                    c.map_joint[slot] = lc.map_joint[e];
                    c.map_cfg[slot] = lc.art_id;
                    c.map_seq[slot] = ((uint32_t)slot << 24) | (uint32_t)ci;
                    c.map_disc[slot] = sp.leaves[l].discmask;
                    for (int s = 0; s < S; ++s) c.map_vaf[slot][s] = sp.leaf_vaf[(size_t)l * S + s];
                }
            }
        }
        c.n_base = n_base;
    }
    warp_sync();
    locus_tail(c, wl.n_twins);
}
