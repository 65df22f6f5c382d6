// CPU test of host/vlr_caller.hpp with a recording mock engine (no GPU, no engine library).
#define VLR_CALLER_NO_DEFAULT_ENGINE
#include <cassert>
#include <cstdio>
#include <cstring>

#include "../../host/vlr_caller.hpp"

using namespace vlr;

static std::vector<int64_t> g_batch_sizes;
static vlr_status_t mock_create(const vlr_scenario_t*, int32_t, vlr_ctx_t** out) {
    *out = reinterpret_cast<vlr_ctx_t*>(0x1);
    return VLR_OK;
}
static void mock_destroy(vlr_ctx_t*) {}
static vlr_status_t mock_call(vlr_ctx_t*, const vlr_batch_t* b, vlr_results_t* r) {
    g_batch_sizes.push_back(b->n_loci);
    const int S = 2, E = 2;
    for (int64_t i = 0; i < b->n_loci; ++i) {
        int64_t n = b->read_offsets[(i + 1) * S] - b->read_offsets[i * S];
        r->log_posteriors[i * (E + 1) + 0] = -(double)n;           // encodes the locus' read count
        r->log_posteriors[i * (E + 1) + 1] = (double)b->prob_mapping[b->read_offsets[i * S]]; // and its first value
        r->log_posteriors[i * (E + 1) + 2] = -INFINITY;
        r->map_vaf[i * S] = 0.0;
        r->map_vaf[i * S + 1] = 0.25;
        r->map_config[i] = 0;
        r->status[i] = (b->locus_flags[i] & VLR_LF_HAS_SNV) ? 0u : VLR_ST_NO_MAP;
        for (int s = 0; s < S; ++s) {
            r->afd_count[i * S + s] = 1;
            r->afd_vaf[(i * S + s) * r->afd_capacity] = 0.5;
            r->afd_logp[(i * S + s) * r->afd_capacity] = -1.0;
        }
    }
    return VLR_OK;
}

static ObservationSource make_source(int n_records, int reads, float tag, bool shift_pos = false, const char* hap = "") {
    auto i = std::make_shared<int>(0);
    return [=](ObservationRecord& r) {
        if (*i >= n_records) return false;
        r.chrom = "1";
        r.pos = 100 + *i + (shift_pos ? 1 : 0);
        r.ref = "A";
        r.alt = (*i == 3) ? "AT" : "G"; // record 3 is an insertion
        int n = reads + *i;
        r.prob_mapping.assign(n, tag + (float)*i);
        r.prob_ref.assign(n, -0.1f);
        r.prob_alt.assign(n, -3.0f);
        r.prob_missed_allele.assign(n, -1.0f);
        r.prob_sample_alt.assign(n, 0.0f);
        r.prob_double_overlap.assign(n, -INFINITY);
        r.prob_hit_base.assign(n, -5.0f);
        r.read_flags.assign(n, 0u);
        if (hap[0] && (*i == 1 || *i == 4)) r.haplotype = hap;
        ++*i;
        return true;
    };
}

struct Collect : CallProcessor {
    std::vector<Call> calls;
    bool set_up = false, finalized = false;
    void setup(const Caller& c) override { set_up = c.sample_names().size() == 2; }
    void process_call(Call call, const std::vector<std::string>& names) override {
        assert(names.size() == 2);
        calls.push_back(std::move(call));
    }
    void finalize() override { finalized = true; }
};
struct SkipOdd : CandidateFilter {
    bool filter(const WorkItem& w, const std::vector<std::string>&) const override { return w.index % 2 == 0; }
};

int main() {
    vlr_event_t events[2];
    std::memset(events, 0, sizeof events);
    std::strcpy(events[0].name, "absent");
    std::strcpy(events[1].name, "present");
    vlr_scenario_t sc{};
    sc.abi_version = VLR_ABI_VERSION;
    sc.n_samples = 2;
    sc.n_events = 2;
    sc.events = events;
    EngineApi api;
    api.ctx_create = mock_create;
    api.ctx_destroy = mock_destroy;
    api.call_batch = mock_call;
    std::vector<std::string> names{"normal", "tumor"};
    DefaultCandidateFilter all;

    // 1. ordering across batches, flags, fields
    {
        Collect cp;
        g_batch_sizes.clear();
        call_generic(sc, names, {{"normal", make_source(7, 5, -0.5f)}, {"tumor", make_source(7, 9, -0.25f)}}, false, false,
                     false, false, false, false, cp, all, api, 0, 3);
        assert(cp.set_up && cp.finalized && cp.calls.size() == 7);
        assert((g_batch_sizes == std::vector<int64_t>{3, 3, 1}));
        for (size_t i = 0; i < 7; ++i) {
            const Call& c = cp.calls[i];
            assert(c.index == i && c.pos == (int64_t)(100 + i));
            assert(c.event_probs.at("absent") == -(double)(5 + i + 9 + i)); // normal + tumor reads of record i
            assert(c.event_probs.at("present") == (double)(-0.5f + (float)i)); // sample 0 (normal) comes first
            assert(std::isinf(c.event_probs.at("artifact")));
            if (i == 3) assert(!c.sample_info[0].has_value()); // the mock flags non-SNV records as NO_MAP
            else assert(c.sample_info[1]->allelefreq_estimate == 0.25 && c.sample_info[1]->vaf_dist.size() == 1);
        }
        assert(Call::phred(-1.0) > 4.34f && Call::phred(-1.0) < 4.35f);
    }
    // 2. candidate filter, missing sample
    {
        Collect cp;
        SkipOdd f;
        call_generic(sc, names, {{"tumor", make_source(6, 4, -0.25f)}}, false, false, false, false, false, false, cp, f, api);
        assert(cp.calls.size() == 3 && cp.calls[1].index == 2);
        assert(cp.calls[0].event_probs.at("absent") == -4.0); // normal has zero coverage
    }
    // 3. haplotype (breakend) groups: the second member reuses the first member's result, order is kept
    {
        Collect cp;
        g_batch_sizes.clear();
        call_generic(sc, names, {{"normal", make_source(6, 5, -0.5f, false, "bnd1")}, {"tumor", make_source(6, 9, -0.25f, false, "bnd1")}},
                     false, false, false, false, false, false, cp, all, api, 0, 100);
        assert(cp.calls.size() == 6 && g_batch_sizes[0] == 5); // record 4 not computed
        assert(cp.calls[4].event_probs.at("absent") == cp.calls[1].event_probs.at("absent"));
        assert(cp.calls[4].pos == 104 && cp.calls[5].event_probs.at("absent") == -(double)(10 + 14));
    }
    // 4. errors mirror the reference
    auto throws = [&](std::function<void()> fn, const char* what) {
        try {
            fn();
        } catch (const std::runtime_error& e) {
            return std::string(e.what()).find(what) != std::string::npos;
        }
        return false;
    };
    Collect cp;
    assert(throws([&] { call_generic(sc, names, {{"tumour", make_source(2, 3, 0.f)}}, 0, 0, 0, 0, 0, 0, cp, all, api); },
                  "invalid observation sample name"));
    assert(throws([&] { call_generic(sc, names, {{"normal", make_source(2, 3, 0.f)}, {"tumor", make_source(3, 3, 0.f)}}, 0, 0, 0, 0, 0, 0, cp, all, api); },
                  "different numbers of records"));
    assert(throws([&] { call_generic(sc, names, {{"normal", make_source(2, 3, 0.f)}, {"tumor", make_source(2, 3, 0.f, true)}}, 0, 0, 0, 0, 0, 0, cp, all, api); },
                  "inconsistent observations"));
    // 5. per-record switches (calling.rs:513-566)
    {
        ObservationRecord r;
        r.ref = "A";
        r.alt = "G";
        Omit none;
        uint32_t f = locus_flags_for(r, false, none);
        assert((f & 0x7f) == (VLR_LF_CHECK_ROB | VLR_LF_CHECK_SB | VLR_LF_CHECK_RPB | VLR_LF_CHECK_SCB | VLR_LF_CHECK_ALB |
                              VLR_LF_FILTER_NONSTANDARD));
        assert((f & VLR_LF_HAS_SNV) && ((f >> VLR_LF_REFBASE_SHIFT) & 0xff) == 'A' && ((f >> VLR_LF_ALTBASE_SHIFT) & 0xff) == 'G');
        r.alt = "AT";
        f = locus_flags_for(r, true, none);
        assert((f & 0x7f) == (VLR_LF_CHECK_SB | VLR_LF_CHECK_HE | VLR_LF_CHECK_ALB) && ((f >> VLR_LF_VARTYPE_SHIFT) & 3) == 1);
        r.imprecise = true;
        assert((locus_flags_for(r, false, none) & 0x7f) == VLR_LF_CHECK_ALB);
        Omit o;
        o.alt_locus_bias = true;
        r.imprecise = false;
        r.alt = "G";
        assert(!(locus_flags_for(r, false, o) & VLR_LF_CHECK_ALB));
    }
    std::puts("host caller mock test: ok");
    return 0;
}
