// Single-lane HOST build of varlociraptor_b200/csrc/engine_core.cuh (-DVLR_HOST_EMU).
//
// TEST INFRASTRUCTURE ONLY. It exists so the warp-uniform control flow of the CUDA engine (tree walk, adaptive
// integration, prior, bias selection, MAP/AFD bookkeeping) can be compared with the oracle in the GPU-less build
// container. It is never linked into or loaded by the product library (libvlr_engine.so), which has no CPU path.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#define VLR_VARIANT vlr_small
#define VLR_VAR_MAXS 3
#define VLR_VAR_MAXE 8
#define VLR_VAR_MAXD 6
#define VLR_VAR_WAVE 1
#include "../../varlociraptor_b200/csrc/engine_core.cuh"
#define VLR_VARIANT vlr_full
#define VLR_VAR_MAXS VLR_MAX_SAMPLES
#define VLR_VAR_MAXE VLR_MAX_EVENTS
#define VLR_VAR_MAXD VLR_MAX_TREE_DEPTH
#include "../../varlociraptor_b200/csrc/engine_core.cuh"
#include "../../varlociraptor_b200/csrc/scenario_prep.h"
#include "../../varlociraptor_b200/csrc/contamination.cuh"

static void emu_views(const vlr_batch_t* batch, vlr_results_t* results, vlrcore::DevBatch& db, vlrcore::DevResults& dr) {
    db.n_loci = batch->n_loci;
    db.read_base = 0;
    db.read_offsets = batch->read_offsets;
    db.pm = batch->prob_mapping;
    db.pr = batch->prob_ref;
    db.pa = batch->prob_alt;
    db.pmiss = batch->prob_missed_allele;
    db.psa = batch->prob_sample_alt;
    db.pdo = batch->prob_double_overlap;
    db.phb = batch->prob_hit_base;
    db.rflags = batch->read_flags;
    db.hart = batch->prob_homopolymer_artifact;
    db.hvar = batch->prob_homopolymer_variant;
    db.lflags = batch->locus_flags;
    db.het_phred = batch->locus_heterozygosity_phred;
    db.semr_phred = batch->locus_semr_phred;
    dr.log_post = results->log_posteriors;
    dr.log_marginal = results->log_marginal;
    dr.map_vaf = results->map_vaf;
    dr.map_config = results->map_config;
    dr.best_event = results->best_event;
    dr.status = results->status;
    dr.n_base_events = results->n_base_events;
    dr.afd_capacity = results->afd_capacity;
    dr.afd_count = results->afd_count;
    dr.afd_vaf = results->afd_vaf;
    dr.afd_logp = results->afd_logp;
}

extern "C" int32_t vlr_emu_call_batch(const vlr_scenario_t* sc, const vlr_batch_t* batch, vlr_results_t* results) {
    using namespace vlrcore;
    using namespace vlr_full;
    ScenarioPrep prep;
    if (!prep.build(sc)) return VLR_ERR_INVALID_ARGUMENT;
    DevScenario ds = prep.view(sc->samples, sc->events, sc->nodes, sc->set_vafs, sc->spectra, prep.lfc_nodes.data(),
                               prep.lfc_ordinal.data());
    std::vector<PriorTabEntry> ptab(PRIOR_TAB_N);
    std::memset(ptab.data(), 0, sizeof(PriorTabEntry) * PRIOR_TAB_N);
    ds.prior_tab = ptab.data();
    DevBatch db;
    DevResults dr;
    emu_views(batch, results, db, dr);
    int64_t max_reads = 1;
    const int S = sc->n_samples;
    for (int64_t i = 0; i < batch->n_loci; ++i) {
        int64_t n = batch->read_offsets[(i + 1) * S] - batch->read_offsets[i * S];
        if (n > max_reads) max_reads = n;
    }
    std::vector<double> coef((size_t)max_reads * 4);
    std::vector<double> be((size_t)BE_CAP * (2 + S));
    WarpWs* ws = new WarpWs;
    Ctx* c = new Ctx;
    for (int64_t i = 0; i < batch->n_loci; ++i) process_locus(&ds, &db, &dr, ws, coef.data(), nullptr, 0, be.data(), (int)max_reads, i, *c);
    delete c;
    delete ws;
    return VLR_OK;
}

// The wavefront pipeline (engine_wave.cuh), run sequentially: prep per locus, rounds of tasks, lc advance, finish, and
// the generic engine for deferred loci — the same device functions the CUDA kernels call, one "thread" at a time.
// Returns the number of deferred loci (>= 0) or a negative status; -100 = scenario not eligible.
extern "C" int32_t vlr_emu_wave_call_batch(const vlr_scenario_t* sc, const vlr_batch_t* batch, vlr_results_t* results) {
    using namespace vlrcore;
    using namespace vlr_small;
    ScenarioPrep prep;
    if (!prep.build(sc)) return -VLR_ERR_INVALID_ARGUMENT;
    const WavePlan wp = prep.wave_plan();
    if (!wp.eligible) return -100;
    DevScenario ds = prep.view(sc->samples, sc->events, sc->nodes, sc->set_vafs, sc->spectra, prep.lfc_nodes.data(),
                               prep.lfc_ordinal.data());
    DevBatch db;
    DevResults dr;
    emu_views(batch, results, db, dr);
    const int S = sc->n_samples;
    const int64_t L = batch->n_loci;
    int64_t max_reads = 1;
    for (int64_t i = 0; i < L; ++i) {
        int64_t n = batch->read_offsets[(i + 1) * S] - batch->read_offsets[i * S];
        if (n > max_reads) max_reads = n;
    }
    const bool want_be = results->afd_capacity > 0;
    WaveCounters cnt;
    std::memset(&cnt, 0, sizeof cnt);
    const int lc_cap = (int)L * 4 + 16; // some loci overflow on purpose (-> deferred)
    std::vector<WaveLocus> loci((size_t)L);
    std::vector<WaveLC> lcs((size_t)lc_cap);
    std::vector<double> og_x((size_t)lc_cap * W_OGRID), og_f((size_t)lc_cap * W_OGRID);
    const int64_t coef_cap = batch->n_reads * 4 + 64;
    std::vector<double> coef((size_t)coef_cap * 4);
    std::vector<WaveTask> tasks0((size_t)lc_cap * W_MAXT), tasks1((size_t)lc_cap * W_MAXT);
    std::vector<int> list0((size_t)lc_cap), list1((size_t)lc_cap), dlist0((size_t)lc_cap), dlist1((size_t)lc_cap), deferred((size_t)L + 1);
    std::vector<double> gx((size_t)W_GCAP * W_MAXT), gf((size_t)W_GCAP * W_MAXT), scratch(3 * W_GCAP);
    std::vector<double> be(want_be ? (size_t)L * BE_CAP * 4 : 4);
    std::vector<unsigned> be_n((size_t)L + 1);
    WaveBufs wb;
    wb.cnt = &cnt;
    wb.loci = loci.data();
    wb.lcs = lcs.data();
    wb.og_x = og_x.data();
    wb.og_f = og_f.data();
    wb.coef = coef.data();
    wb.tasks[0] = tasks0.data();
    wb.tasks[1] = tasks1.data();
    wb.list[0] = list0.data();
    wb.list[1] = list1.data();
    wb.dlist[0] = dlist0.data();
    wb.dlist[1] = dlist1.data();
    wb.deferred = deferred.data();
    wb.gx = gx.data();
    wb.gf = gf.data();
    wb.be = want_be ? be.data() : nullptr;
    wb.be_n = be_n.data();
    wb.coef_cap = coef_cap * 4;
    wb.lc_cap = lc_cap;
    std::vector<int> rlist((size_t)lc_cap * R_CLASSES);
    std::vector<double> rgx((size_t)W_MAXT * W_GCAP), rgm((size_t)W_MAXT * W_GCAP), rscratch(3 * W_GCAP);
    const int cs_reads = (int)std::min<int64_t>(R_MAXREADS, std::max<int64_t>(max_reads, 256));
    std::vector<double> cscratch((size_t)6 * cs_reads);
    std::vector<int> rge((size_t)W_MAXT * W_GCAP);
    wb.rlist = rlist.data();
    wb.rgx = rgx.data();
    wb.rgm = rgm.data();
    wb.rge = rge.data();
    wb.rscratch = rscratch.data();
    wb.cscratch = cscratch.data();
    wb.cscratch_reads = cs_reads;
    {
        const char* e = getenv("VLR_RESIDENT");
        wb.allow_resident = !(e && e[0] == '0');
    }
    WarpWs* ws = new WarpWs;
    Ctx* c = new Ctx;
    for (int64_t i = 0; i < L; ++i) wave_pre_locus(&ds, &db, wp, wb, i, (int)i, want_be, *c);
    const int n_lc = (int)std::min<unsigned>(cnt.n_lc, (unsigned)lc_cap);
    std::fill(rlist.begin(), rlist.end(), -1);
    unsigned key_off[R_CLASSES][W_KEYS];
    for (int cl = 0; cl < R_CLASSES; ++cl) wave_key_offsets(cnt.rkey_n[cl], key_off[cl]);
    for (int k = 0; k < n_lc; ++k) {
        const int id = wave_lc_init(&ds, wp, wb, k);
        if (id < 0) continue;
        const unsigned at = key_off[id / W_KEYS][id % W_KEYS] + wa_add_u32(&cnt.rkey_cur[id / W_KEYS][id % W_KEYS], 1u);
        rlist[(size_t)(id / W_KEYS) * lc_cap + at] = k;
    }
    for (int k = 0; k < n_lc; ++k) wave_lc_coef(&ds, &db, wp, wb, k, 0, want_be, *c, 0);
    // lc-resident rounds (engine_resident.cuh): one group of a single lane per size class, the slot filled by a plain copy
    {
        ROct* oc = new ROct;
        std::vector<double> slot((size_t)R_SLOT_QL * R_QW);
        oc->q = slot.data();
        const WGroup one{0, 1, 1u};
        const WSplit lane{0, 1, 1u};
        for (int cls = 1; cls <= R_CLASSES; ++cls) {
            const int* rl = rlist.data() + (size_t)(cls - 1) * lc_cap;
            const unsigned n_all = cnt.rlist_total[cls - 1];
            for (unsigned k = 0; k < n_all; ++k) {
                const int lci = rl[k];
                if (lci < 0) return -VLR_ERR_INVALID_ARGUMENT; // the key counts of the pre-pass and the placement disagree
                const WaveLC& L = lcs[lci];
                oc->lc.lci = lci;
                oc->lc.li = L.li;
                oc->lc.ci = L.ci;
                oc->lc.nqPx = L.nqPx;
                oc->lc.nqPy = L.nqPy;
                oc->lc.nqTx = L.nqTx;
                oc->lc.nqTy = L.nqTy;
                oc->lc.ksumP = L.ksumP;
                oc->lc.ksumT = L.ksumT;
                const int nqP = L.nqPx + L.nqPy, nq = L.nqTx + L.nqTy;
                if (nq > r_slot_q(cls)) return -VLR_ERR_INVALID_ARGUMENT;
                oc->lc.qP = wb.coef + L.coefP;
                std::memcpy(oc->q, wb.coef + L.coefP + (size_t)nqP * R_QW, sizeof(double) * R_QW * (size_t)nq);
                int n_tasks = r_first_tasks(&ds, wp, wb, lci, oc->task, one);
                for (int round = 0; n_tasks > 0; ++round) {
                    for (int t = 0; t < n_tasks; ++t)
                        r_task(&ds, wp, wb, *oc, t, n_tasks, wb.rgx + (size_t)t * W_GCAP, wb.rgm + (size_t)t * W_GCAP, wb.rge + (size_t)t * W_GCAP, lane, 1u);
                    n_tasks = r_advance(&ds, wp, wb, *oc, round, n_tasks, wb.rgx, wb.rgm, wb.rge, wb.rscratch, want_be, one);
                }
            }
        }
        delete oc;
    }
    for (int round = 0; round < wp.max_rounds; ++round) {
        for (int which = 0; which < 2; ++which) { // shallow list, then deep list
            const int n_list = (int)(which == 0 ? cnt.list_n[round] : cnt.dlist_n[round]);
            const int* list = which == 0 ? wb.list[round & 1] : wb.dlist[round & 1];
            for (int k = 0; k < n_list; ++k) {
                WaveLC& lc = lcs[list[k]];
                for (int t = 0; t < lc.task_count; ++t) {
                    WaveTask& task = wb.tasks[round & 1][lc.task_base + t];
                    const WSplit one{0, 1, 1u};
                    const double lh = wave_task_parent(lc, task, reinterpret_cast<const double2*>(wb.coef + lc.coefP), false, one);
                    wave_task_run(&ds, wp, lc, task, reinterpret_cast<const double2*>(wb.coef + lc.coefT), false, lh,
                                  wb.gx + (size_t)t * W_GCAP, wb.gf + (size_t)t * W_GCAP, one);
                }
                wave_lc_advance(wp, wb, list[k], round, wb.gx, wb.gf, W_GCAP, scratch.data(), want_be, WGroup{0, 1, 1u});
            }
        }
    }
    for (int64_t i = 0; i < L; ++i) wave_finish_locus(&ds, &db, &dr, wp, wb, ws, i, (int)i, *c);
    std::vector<double> coef2((size_t)max_reads * 4);
    std::vector<double> be2((size_t)BE_CAP * (2 + S));
    for (unsigned k = 0; k < cnt.n_deferred; ++k)
        process_locus(&ds, &db, &dr, ws, coef2.data(), nullptr, 0, be2.data(), (int)max_reads, deferred[k], *c);
    delete c;
    delete ws;
    return (int32_t)cnt.n_deferred;
}

// The all-Set pipeline (engine_sets.cuh), run sequentially: pre per locus, lc per (locus, config), finish, and the
// generic engine for deferred loci. Returns the number of deferred loci (>= 0) or a negative status; -100 = scenario
// not eligible.
extern "C" int32_t vlr_emu_sets_call_batch(const vlr_scenario_t* sc, const vlr_batch_t* batch, vlr_results_t* results) {
    using namespace vlrcore;
    using namespace vlr_small;
    ScenarioPrep prep;
    if (!prep.build(sc)) return -VLR_ERR_INVALID_ARGUMENT;
    if (sc->n_samples > 3 || sc->n_events > 8 || prep.max_depth > 6) return -100;
    SetsPlan sp = prep.sets_plan();
    if (!sp.eligible) return -100;
    DevScenario ds = prep.view(sc->samples, sc->events, sc->nodes, sc->set_vafs, sc->spectra, prep.lfc_nodes.data(),
                               prep.lfc_ordinal.data());
    std::vector<PriorTabEntry> ptab(PRIOR_TAB_N);
    std::memset(ptab.data(), 0, sizeof(PriorTabEntry) * PRIOR_TAB_N);
    ds.prior_tab = ptab.data();
    std::vector<double> prior_val((size_t)4 * sp.n_leaves);
    std::vector<uint32_t> prior_side((size_t)4 * sp.n_leaves);
    int prior_state[4] = {0, 0, 0, 0};
    sp.folds = prep.sets_folds.data();
    sp.leaves = prep.sets_leaves.data();
    sp.leaf_vaf = prep.sets_leaf_vaf.data();
    sp.prior_val = prior_val.data();
    sp.prior_side = prior_side.data();
    sp.prior_state = prior_state;
    DevBatch db;
    DevResults dr;
    emu_views(batch, results, db, dr);
    const int S = sc->n_samples;
    const int64_t L = batch->n_loci;
    int64_t max_reads = 1;
    for (int64_t i = 0; i < L; ++i) {
        int64_t n = batch->read_offsets[(i + 1) * S] - batch->read_offsets[i * S];
        if (n > max_reads) max_reads = n;
    }
    const bool want_be = results->afd_capacity > 0;
    SetsCounters cnt;
    std::memset(&cnt, 0, sizeof cnt);
    const int lc_cap = (int)L * 9 + 16;
    std::vector<SetsLocus> loci((size_t)L + 1);
    std::vector<SetsLC> lcs((size_t)lc_cap);
    std::vector<int> deferred((size_t)L + 1);
    std::vector<double> be(want_be ? (size_t)L * SETS_MAXL * (2 + S) : 4);
    std::vector<unsigned> be_n((size_t)L + 1);
    SetsBufs sb;
    sb.cnt = &cnt;
    sb.loci = loci.data();
    sb.lcs = lcs.data();
    sb.deferred = deferred.data();
    sb.be = want_be ? be.data() : nullptr;
    sb.be_n = be_n.data();
    sb.lc_cap = lc_cap;
    WarpWs* ws = new WarpWs;
    Ctx* c = new Ctx;
    std::vector<double> arena((size_t)4 * SETS_SM_READS + SETS_MAXF);
    for (int64_t i = 0; i < L; ++i) sets_pre_locus(&ds, &db, sb, i, (int)i, want_be, *c);
    const int n_lc = (int)std::min<unsigned>(cnt.n_lc, (unsigned)lc_cap);
    static MemoTab memo; // the table path of read_coefficients (cleared per call like the kernel does per launch)
    memo_clear(&memo);
    for (int k = 0; k < n_lc; ++k) sets_lc(&ds, &db, sp, sb, k, 0, want_be, *c, arena.data(), arena.data() + 4 * SETS_SM_READS, false, &memo);
    for (int64_t i = 0; i < L; ++i) sets_finish_locus(&ds, &db, &dr, sp, sb, ws, i, (int)i, *c);
    std::vector<double> coef2((size_t)max_reads * 4);
    std::vector<double> be2((size_t)BE_CAP * (2 + S));
    for (unsigned k = 0; k < cnt.n_deferred; ++k)
        process_locus(&ds, &db, &dr, ws, coef2.data(), nullptr, 0, be2.data(), (int)max_reads, deferred[k], *c);
    delete c;
    delete ws;
    return (int32_t)cnt.n_deferred;
}

// The contamination model's device functions (csrc/contamination.cuh) driven the way the two kernels drive them:
// per event a partial sum per chunk of `chunk` observations, combined in chunk order, then the Simpson rows.
extern "C" int32_t vlr_emu_contamination_posterior(const vlr_contamination_input_t* in, vlr_contamination_output_t* out,
                                                    int64_t chunk) {
    using namespace vlrcontam;
    const int n_events = in->n_grid * in->n_max_vafs;
    if (n_events > CONTAM_MAX_EVENTS || in->n_max_vafs > CONTAM_MAX_ROWS || chunk < 1) return VLR_ERR_INVALID_ARGUMENT;
    double max_vaf = 0.0;
    for (int64_t o = 0; o < in->n_obs; ++o) max_vaf = fmax(max_vaf, in->max_posterior_vaf[o]);
    std::vector<double> joint((size_t)n_events), scratch((size_t)n_events), rows((size_t)in->n_max_vafs);
    for (int e = 0; e < n_events; ++e) {
        const int k = e / in->n_grid, gi = e % in->n_grid;
        const double purity = 1.0 - grid_contamination(gi, in->n_grid);
        double lik = 0.0;
        for (int64_t o0 = 0; o0 < in->n_obs; o0 += chunk) {
            double sum = 0.0;
            for (int64_t o = o0; o < std::min(in->n_obs, o0 + chunk); ++o) {
                const int64_t base = in->afd_offsets[o];
                sum += obs_term(in->prob_denovo[o], in->max_posterior_vaf[o], max_vaf, in->expected_max_somatic_vaf[k],
                                purity, in->afd_vaf + base, in->afd_logp + base, (int)(in->afd_offsets[o + 1] - base));
            }
            lik += sum;
        }
        if (out->ln_likelihood) out->ln_likelihood[e] = lik;
        joint[e] = in->ln_prior[gi] + lik;
    }
    for (int k = 0; k < in->n_max_vafs; ++k)
        rows[k] = simpson_row(joint.data() + (size_t)k * in->n_grid, in->n_grid, scratch.data() + (size_t)k * in->n_grid);
    const double marginal = ln_sum_exp_slice(rows.data(), in->n_max_vafs);
    for (int e = 0; e < n_events; ++e) out->ln_posterior[e] = joint[e] - marginal;
    *out->ln_marginal = marginal;
    if (out->max_vaf) *out->max_vaf = max_vaf;
    return VLR_OK;
}
