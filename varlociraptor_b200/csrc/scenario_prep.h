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
    int events_overlap = 0;
    const vlr_scenario_t* src = nullptr;
    const char* error = "";

    // ---- may two events contain the same VAF combination? ----------------------------------------------------------
    // The reference picks the MAP among ALL recorded base events that the best event's tree contains (calling.rs:851-864,
    // vaftree.rs:42-51). For pairwise disjoint events those are the base events the best event recorded itself (plus
    // points on excluded range bounds, which the engines detect per locus); for overlapping events (the reference's
    // validation only rejects containment, grammar/mod.rs:223-278) the engines offer every base event to every event.
    // Conservative test over root-to-leaf paths: per sample the Set / Range constraints of both paths must intersect;
    // log2-fold-change, variant and true nodes are treated as no constraint.
    struct PathSet {
        std::vector<std::vector<int>> paths; // node indices (Set / Range nodes) along each root-to-leaf path
        bool too_many = false;
    };
    void collect_paths(const vlr_scenario_t* sc, int ni, std::vector<int>& cur, PathSet& out) {
        if (out.too_many) return;
        const vlr_node_t& n = sc->nodes[ni];
        if (n.kind == VLR_NODE_FALSE) return;
        const bool constrains = n.kind == VLR_NODE_SET || n.kind == VLR_NODE_RANGE;
        if (constrains) cur.push_back(ni);
        if (n.n_children == 0) {
            if (out.paths.size() >= 4096) out.too_many = true;
            else out.paths.push_back(cur);
        } else {
            for (int k = 0; k < n.n_children; ++k) collect_paths(sc, n.first_child + k, cur, out);
        }
        if (constrains) cur.pop_back();
    }
    static bool h_range_contains(const vlr_node_t& r, double v) {
        const bool l = r.left_exclusive ? (r.start < v) : (r.start <= v);
        const bool rr = r.right_exclusive ? (r.end > v) : (r.end >= v);
        return l && rr;
    }
    bool paths_overlap(const vlr_scenario_t* sc, const std::vector<int>& a, const std::vector<int>& b) const {
        for (int s = 0; s < sc->n_samples; ++s) {
            std::vector<const vlr_node_t*> cons;
            for (int ni : a)
                if (sc->nodes[ni].sample == s) cons.push_back(&sc->nodes[ni]);
            for (int ni : b)
                if (sc->nodes[ni].sample == s) cons.push_back(&sc->nodes[ni]);
            if (cons.size() < 2) continue;
            const vlr_node_t* set = nullptr;
            for (const vlr_node_t* n : cons)
                if (n->kind == VLR_NODE_SET) set = n;
            bool any = false;
            if (set) { // some value of the set satisfies every constraint
                for (int i = 0; i < set->n_vafs && !any; ++i) {
                    const double v = sc->set_vafs[set->vaf_offset + i];
                    bool ok = true;
                    for (const vlr_node_t* n : cons) {
                        if (n->kind == VLR_NODE_SET) {
                            bool in = false;
                            for (int j = 0; j < n->n_vafs; ++j) in = in || sc->set_vafs[n->vaf_offset + j] == v;
                            ok = ok && in;
                        } else {
                            ok = ok && h_range_contains(*n, v);
                        }
                    }
                    any = ok;
                }
            } else { // intersection of intervals
                double lo = -INFINITY, hi = INFINITY;
                bool lex = false, rex = false;
                for (const vlr_node_t* n : cons) {
                    if (n->start > lo || (n->start == lo && n->left_exclusive)) {
                        lex = n->start > lo ? n->left_exclusive != 0 : (lex || n->left_exclusive != 0);
                        lo = n->start;
                    }
                    if (n->end < hi || (n->end == hi && n->right_exclusive)) {
                        rex = n->end < hi ? n->right_exclusive != 0 : (rex || n->right_exclusive != 0);
                        hi = n->end;
                    }
                }
                any = lo < hi || (lo == hi && !lex && !rex);
            }
            if (!any) return false;
        }
        return true;
    }
    int compute_events_overlap(const vlr_scenario_t* sc) {
        std::vector<PathSet> ps((size_t)sc->n_events);
        for (int e = 0; e < sc->n_events; ++e) {
            std::vector<int> cur;
            for (int r = 0; r < sc->events[e].n_roots; ++r) collect_paths(sc, sc->events[e].first_root + r, cur, ps[e]);
            if (ps[e].too_many) return 1;
        }
        for (int a = 0; a < sc->n_events; ++a)
            for (int b = a + 1; b < sc->n_events; ++b)
                for (const auto& pa : ps[a].paths)
                    for (const auto& pb : ps[b].paths)
                        if (paths_overlap(sc, pa, pb)) return 1;
        return 0;
    }

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
        events_overlap = compute_events_overlap(sc);
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
        if (events_overlap) return wp; // the pipeline's MAP bookkeeping is per event (see compute_events_overlap)
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

    // ---- all-Set scenarios (engine_sets.cuh): leaves in the order GenericPosterior::density visits them ------------
    std::vector<SetsFold> sets_folds;
    std::vector<SetsLeaf> sets_leaves;
    std::vector<double> sets_leaf_vaf;
    bool sets_ok = true;

    void sets_walk(const vlr_scenario_t* sc, int ni, int event, double* vaf, unsigned posmask, unsigned discmask) {
        if (!sets_ok) return;
        const vlr_node_t& n = sc->nodes[ni];
        if (n.kind != VLR_NODE_SET || n.n_vafs < 1) {
            sets_ok = false;
            return;
        }
        bool all_pos = true;
        for (int i = 0; i < n.n_vafs; ++i)
            if (!(sc->set_vafs[n.vaf_offset + i] > 0.0)) all_pos = false;
        const unsigned pm = posmask | (all_pos ? 1u << n.sample : 0u);
        const double saved = vaf[n.sample];
        for (int i = 0; i < n.n_vafs; ++i) {
            vaf[n.sample] = sc->set_vafs[n.vaf_offset + i];
            const unsigned dm = discmask | (1u << n.sample);
            if (n.n_children == 0) {
                if ((int)sets_leaves.size() >= SETS_MAXL) {
                    sets_ok = false;
                    return;
                }
                SetsLeaf lf;
                std::memset(&lf, 0, sizeof lf);
                lf.event = (uint8_t)event;
                lf.posmask = (uint8_t)pm;
                lf.discmask = (uint8_t)dm;
                for (int s = 0; s < sc->n_samples; ++s) {
                    const int by = sc->samples[s].contamination_by;
                    const double v = vaf[s], vb = by >= 0 ? vaf[by] : 0.0;
                    int f = -1;
                    for (size_t k = 0; k < sets_folds.size(); ++k)
                        if (sets_folds[k].sample == s && sets_folds[k].vaf == v && sets_folds[k].vaf_by == vb) f = (int)k;
                    if (f < 0) {
                        if ((int)sets_folds.size() >= SETS_MAXF) {
                            sets_ok = false;
                            return;
                        }
                        f = (int)sets_folds.size();
                        sets_folds.push_back(SetsFold{v, vb, s, 0});
                    }
                    lf.fold[s] = (uint8_t)f;
                }
                sets_leaves.push_back(lf);
                for (int s = 0; s < sc->n_samples; ++s) sets_leaf_vaf.push_back(vaf[s]);
            } else {
                for (int k = 0; k < n.n_children; ++k) sets_walk(sc, n.first_child + k, event, vaf, pm, dm);
            }
        }
        vaf[n.sample] = saved;
    }

    // Host part of the plan (the device pointers are filled in by the caller). Not eligible: anything but Set nodes,
    // overlapping events (MAP candidates would have to be offered across events, engine_core.cuh map_offer_all), more
    // than SETS_MAXS samples.
    SetsPlan sets_plan() {
        SetsPlan sp;
        std::memset(&sp, 0, sizeof sp);
        const vlr_scenario_t* sc = src;
        sets_folds.clear();
        sets_leaves.clear();
        sets_leaf_vaf.clear();
        sets_ok = true;
        if (!sc || sc->n_samples > SETS_MAXS || sc->n_samples > 3 || events_overlap || !lfc_nodes.empty()) return sp;
        for (int e = 0; e < sc->n_events && sets_ok; ++e) {
            sp.ev_first[e] = (int)sets_leaves.size();
            for (int r = 0; r < sc->events[e].n_roots && sets_ok; ++r) {
                double vaf[VLR_MAX_SAMPLES];
                for (int s = 0; s < VLR_MAX_SAMPLES; ++s) vaf[s] = 0.0; // an unset sample reads as 0.0 (Ops of the generic engine)
                sets_walk(sc, sc->events[e].first_root + r, e, vaf, 0u, 0u);
            }
            sp.ev_count[e] = (int)sets_leaves.size() - sp.ev_first[e];
        }
        if (!sets_ok || sets_leaves.empty()) return sp;
        sp.eligible = 1;
        sp.n_folds = (int)sets_folds.size();
        sp.n_leaves = (int)sets_leaves.size();
        return sp;
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
        d.events_overlap = events_overlap;
        d.pad_ = 0;
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
