// GPU test of host/vlr_caller.hpp against the real engine library: reads a flattened scenario dumped by the Python
// test (raw C structs), feeds two samples' records through call_generic and prints the calls as lines of numbers.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>

#include "../../host/vlr_caller.hpp"

using namespace vlr;

template <class T> static std::vector<T> slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<T> out(raw.size() / sizeof(T));
    std::memcpy(out.data(), raw.data(), out.size() * sizeof(T));
    return out;
}

struct Print : CallProcessor {
    std::vector<std::string> events;
    void setup(const Caller& c) override { events = c.event_names(); }
    void process_call(Call call, const std::vector<std::string>&) override {
        std::printf("CALL %zu %u", call.index, call.status);
        for (auto& e : events) std::printf(" %.17g", call.event_probs.at(e));
        std::printf(" %.17g", call.event_probs.at("artifact"));
        for (auto& s : call.sample_info) std::printf(" %.17g %zu", s ? s->allelefreq_estimate : NAN, s ? s->vaf_dist.size() : 0);
        std::printf("\n");
    }
};

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    auto samples = slurp<vlr_sample_t>(dir + "/samples.bin");
    auto events = slurp<vlr_event_t>(dir + "/events.bin");
    auto nodes = slurp<vlr_node_t>(dir + "/nodes.bin");
    auto set_vafs = slurp<double>(dir + "/set_vafs.bin");
    auto spectra = slurp<vlr_spectrum_t>(dir + "/spectra.bin");
    vlr_scenario_t sc{};
    sc.abi_version = VLR_ABI_VERSION;
    sc.n_samples = (int)samples.size();
    sc.n_events = (int)events.size();
    sc.n_nodes = (int)nodes.size();
    sc.n_set_vafs = (int)set_vafs.size();
    sc.n_spectra = (int)spectra.size();
    sc.samples = samples.data();
    sc.events = events.data();
    sc.nodes = nodes.data();
    sc.set_vafs = set_vafs.data();
    sc.spectra = spectra.data();
    sc.heterozygosity = NAN;
    sc.vtf_indel = 0.0125;
    sc.vtf_mnv = 0.001;
    sc.vtf_sv = 0.01;
    sc.full_prior = 0;
    // records: per sample a file of [n_records] x { n_reads, then 7 float columns and the flag column }
    auto source = [&](const std::string& name) {
        auto raw = std::make_shared<std::vector<float>>(slurp<float>(dir + "/" + name + ".bin"));
        auto at = std::make_shared<size_t>(0);
        auto idx = std::make_shared<int>(0);
        return ObservationSource([=](ObservationRecord& r) {
            if (*at >= raw->size()) return false;
            size_t n = (size_t)(*raw)[(*at)++];
            r.chrom = "1";
            r.pos = 1000 + *idx;
            r.ref = "A";
            r.alt = "G";
            std::vector<float>* cols[7] = {&r.prob_mapping, &r.prob_ref, &r.prob_alt, &r.prob_missed_allele,
                                           &r.prob_sample_alt, &r.prob_double_overlap, &r.prob_hit_base};
            for (auto* c : cols) {
                c->assign(raw->begin() + *at, raw->begin() + *at + n);
                *at += n;
            }
            r.read_flags.resize(n);
            for (size_t i = 0; i < n; ++i) r.read_flags[i] = (uint32_t)(*raw)[*at + i];
            *at += n;
            ++*idx;
            return true;
        });
    };
    Print cp;
    DefaultCandidateFilter all;
    try {
        call_generic(sc, {"normal", "tumor"}, {{"normal", source("normal")}, {"tumor", source("tumor")}}, false, false, false,
                     false, false, false, cp, all, EngineApi::linked(), 0, 4);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
