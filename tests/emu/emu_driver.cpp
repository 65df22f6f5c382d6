// Single-lane HOST build of varlociraptor_b200/csrc/engine_core.cuh (-DVLR_HOST_EMU).
//
// TEST INFRASTRUCTURE ONLY. It exists so the warp-uniform control flow of the CUDA engine (tree walk, adaptive
// integration, prior, bias selection, MAP/AFD bookkeeping) can be compared with the oracle in the GPU-less build
// container. It is never linked into or loaded by the product library (libvlr_engine.so), which has no CPU path.
#include <cstdlib>
#include <cstring>
#include <vector>

#define VLR_VARIANT vlr_full
#define VLR_VAR_MAXS VLR_MAX_SAMPLES
#define VLR_VAR_MAXE VLR_MAX_EVENTS
#define VLR_VAR_MAXD VLR_MAX_TREE_DEPTH
#include "../../varlociraptor_b200/csrc/engine_core.cuh"
#include "../../varlociraptor_b200/csrc/scenario_prep.h"

extern "C" int32_t vlr_emu_call_batch(const vlr_scenario_t* sc, const vlr_batch_t* batch, vlr_results_t* results) {
    using namespace vlrcore;
    using namespace vlr_full;
    ScenarioPrep prep;
    if (!prep.build(sc)) return VLR_ERR_INVALID_ARGUMENT;
    DevScenario ds = prep.view(sc->samples, sc->events, sc->nodes, sc->set_vafs, sc->spectra, prep.lfc_nodes.data(),
                               prep.lfc_ordinal.data());
    DevBatch db;
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
    DevResults dr;
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
