// scenario_prep.h — host-side validation and derived tables of a vlr_scenario_t (shared by the CUDA host code and
// the test-only host emulation): LFC node ordinals, tree depth, the all-uniform flag (prior.rs:107-109).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "engine_types.cuh"

namespace vlrcore {

struct ScenarioPrep {
    std::vector<int> lfc_nodes, lfc_ordinal;
    int max_depth = 0;
    int all_uniform = 1;
    const vlr_scenario_t* src = nullptr;
    const char* error = "";

    int depth_of(const vlr_scenario_t* sc, int ni, int guard) {
        if (guard > sc->n_nodes + 1) return 1 << 20;
        const vlr_node_t& n = sc->nodes[ni];
        int d = 0;
        for (int k = 0; k < n.n_children; ++k) {
            int c = n.first_child + k;
            if (c < 0 || c >= sc->n_nodes) return 1 << 20;
            int dc = depth_of(sc, c, guard + 1);
            if (dc > d) d = dc;
        }
        return d + 1;
    }

    bool build(const vlr_scenario_t* sc) {
        src = sc;
        if (!sc || sc->abi_version != VLR_ABI_VERSION) return fail("ABI version mismatch");
        if (sc->n_samples < 1 || sc->n_samples > VLR_MAX_SAMPLES) return fail("n_samples out of range");
        if (sc->n_events < 1 || sc->n_events > VLR_MAX_EVENTS) return fail("n_events out of range");
        if (sc->n_nodes < 1 || !sc->nodes || !sc->events || !sc->samples) return fail("missing arrays");
        lfc_ordinal.assign(sc->n_nodes, -1);
        for (int i = 0; i < sc->n_nodes; ++i) {
            const vlr_node_t& n = sc->nodes[i];
            if (n.kind < VLR_NODE_SET || n.kind > VLR_NODE_FALSE) return fail("invalid node kind");
            if (n.kind == VLR_NODE_SET || n.kind == VLR_NODE_RANGE || n.kind == VLR_NODE_LFC)
                if (n.sample < 0 || n.sample >= sc->n_samples) return fail("node sample out of range");
            if (n.kind == VLR_NODE_LFC) {
                if (n.sample_b < 0 || n.sample_b >= sc->n_samples) return fail("node sample_b out of range");
                if ((int)lfc_nodes.size() >= MAX_LFC_NODES) return fail("too many log2-fold-change nodes (max 32)");
                lfc_ordinal[i] = (int)lfc_nodes.size();
                lfc_nodes.push_back(i);
            }
            if (n.kind == VLR_NODE_SET &&
                (n.vaf_offset < 0 || n.n_vafs < 0 || n.vaf_offset + n.n_vafs > sc->n_set_vafs))
                return fail("set node VAFs out of range");
            if (n.n_children < 0 || (n.n_children > 0 && (n.first_child <= i || n.first_child + n.n_children > sc->n_nodes)))
                return fail("children must follow their parent");
        }
        max_depth = 0;
        for (int e = 0; e < sc->n_events; ++e) {
            const vlr_event_t& ev = sc->events[e];
            if (ev.first_root < 0 || ev.n_roots < 0 || ev.first_root + ev.n_roots > sc->n_nodes)
                return fail("event roots out of range");
            for (int r = 0; r < ev.n_roots; ++r) {
                int d = depth_of(sc, ev.first_root + r, 0);
                if (d > max_depth) max_depth = d;
            }
        }
        if (max_depth > VLR_MAX_TREE_DEPTH) return fail("VAF tree deeper than VLR_MAX_TREE_DEPTH");
        all_uniform = 1;
        for (int s = 0; s < sc->n_samples; ++s) {
            const vlr_sample_t& sm = sc->samples[s];
            if (!sm.uniform_prior) all_uniform = 0;
            if (sm.contamination_by >= sc->n_samples) return fail("contamination_by out of range");
            if (sm.contamination_by >= 0 && !(sm.contamination_fraction >= 0.0 && sm.contamination_fraction < 1.0))
                return fail("contamination fraction must be in [0, 1)"); // likelihood.rs:78 assert
            if (sm.universe_offset < 0 || sm.n_universe < 0 || sm.universe_offset + sm.n_universe > sc->n_spectra)
                return fail("sample universe out of range");
            if (sm.inheritance != VLR_INHERIT_NONE && (sm.parent_a < 0 || sm.parent_a >= sc->n_samples))
                return fail("inheritance parent out of range");
            if (sm.inheritance == VLR_INHERIT_MENDELIAN && (sm.parent_b < 0 || sm.parent_b >= sc->n_samples))
                return fail("mendelian parent out of range");
            if (!(sm.resolution > 0.0)) return fail("resolution must be positive");
        }
        return true;
    }

    // Does the scenario have the two-level chain shape of the wavefront pipeline (engine_types.cuh WavePlan)?
    WavePlan wave_plan() const {
        WavePlan wp;
        std::memset(&wp, 0, sizeof wp);
        wp.outer_event = -1;
        wp.max_rounds = 1;
        const vlr_scenario_t* sc = src;
        if (!sc || sc->n_samples != 2 || sc->n_events > WAVE_MAXE || !all_uniform || !lfc_nodes.empty()) return wp;
        int P = -1, n_leaf_tasks = 0;
        for (int e = 0; e < sc->n_events; ++e) {
            const vlr_event_t& ev = sc->events[e];
            if (ev.n_roots != 1) return wp;
            const vlr_node_t& root = sc->nodes[ev.first_root];
            if (root.n_children != 1) return wp;
            const vlr_node_t& child = sc->nodes[root.first_child];
            if (child.n_children != 0) return wp;
            const vlr_node_t* both[2] = {&root, &child};
            for (const vlr_node_t* n : both) {
                if (n->kind == VLR_NODE_SET) {
                    if (n->n_vafs != 1) return wp;
                } else if (n->kind == VLR_NODE_RANGE) {
                    if (!(n->start < n->end)) return wp; // empty or singleton ranges take the generic path
                } else {
                    return wp;
                }
            }
            if (P < 0) P = root.sample;
            if (root.sample != P || child.sample != 1 - P) return wp;
            if (root.kind == VLR_NODE_RANGE) {
                if (wp.outer_event >= 0 || child.kind != VLR_NODE_RANGE) return wp;
                wp.outer_event = e;
                const double res = sc->samples[P].resolution, w0 = root.end - root.start;
                int iters = 1;
                if (w0 > res) iters = (int)std::ceil(std::log(w0 / res) / std::log(4.0 / 3.0)) + 2;
                if (iters > 60) return wp;
                wp.max_rounds = 1 + iters + 1;
                n_leaf_tasks += 2;
            } else if (child.kind == VLR_NODE_RANGE) {
                n_leaf_tasks += 1;
            }
            wp.root_node[e] = ev.first_root;
            wp.child_node[e] = root.first_child;
        }
        if (P < 0 || n_leaf_tasks > 8) return wp;
        const int T = 1 - P;
        if (sc->samples[P].contamination_by >= 0) return wp;
        if (sc->samples[T].contamination_by >= 0 && sc->samples[T].contamination_by != P) return wp;
        wp.P = P;
        wp.T = T;
        wp.eligible = 1;
        return wp;
    }

    bool fail(const char* msg) {
        error = msg;
        return false;
    }

    DevScenario view(const vlr_sample_t* samples, const vlr_event_t* events, const vlr_node_t* nodes,
                     const double* set_vafs, const vlr_spectrum_t* spectra, const int* lfc_nodes_p,
                     const int* lfc_ordinal_p) const {
        DevScenario d;
        d.prior_tab = nullptr;
        d.S = src->n_samples;
        d.E = src->n_events;
        d.n_nodes = src->n_nodes;
        d.n_set_vafs = src->n_set_vafs;
        d.n_spectra = src->n_spectra;
        d.full_prior = src->full_prior;
        d.all_uniform = all_uniform;
        d.n_lfc_nodes = (int)lfc_nodes.size();
        d.samples = samples;
        d.events = events;
        d.nodes = nodes;
        d.set_vafs = set_vafs;
        d.spectra = spectra;
        d.lfc_nodes = lfc_nodes_p;
        d.lfc_ordinal = lfc_ordinal_p;
        d.heterozygosity = src->heterozygosity;
        d.vtf[0] = 1.0;
        d.vtf[1] = src->vtf_indel;
        d.vtf[2] = src->vtf_mnv;
        d.vtf[3] = src->vtf_sv;
        return d;
    }
};

} // namespace vlrcore
