// Decodes observation INFO arrays given as text ("TAG v1 v2 ..." per line, records separated by "--") with
// host/vlr_obs_codec.hpp and prints every decoded column as raw bits; tests/test_host_cpp.py compares the output with
// varlociraptor_b200.obs_codec.decode_record on the reference's own records.
#define VLR_CALLER_NO_DEFAULT_ENGINE
#include <cstdio>
#include <fstream>
#include <sstream>

#include "../../host/vlr_obs_codec.hpp"

static void print_f32(const char* name, const std::vector<float>& v) {
    std::printf("%s", name);
    for (float x : v) {
        uint32_t b;
        std::memcpy(&b, &x, 4);
        std::printf(" %08x", b);
    }
    std::printf("\n");
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream in(argv[1]);
    std::string line;
    vlr::InfoArrays info;
    auto flush = [&]() {
        if (info.empty()) return;
        vlr::ObservationRecord r;
        vlr::ThirdAlleleEvidence third;
        try {
            vlr::decode_observation_record(info, r, &third);
        } catch (const std::runtime_error& e) {
            std::printf("ERROR %s\n--\n", e.what());
            info.clear();
            return;
        }
        print_f32("prob_mapping", r.prob_mapping);
        print_f32("prob_ref", r.prob_ref);
        print_f32("prob_alt", r.prob_alt);
        print_f32("prob_missed_allele", r.prob_missed_allele);
        print_f32("prob_sample_alt", r.prob_sample_alt);
        print_f32("prob_double_overlap", r.prob_double_overlap);
        print_f32("prob_hit_base", r.prob_hit_base);
        std::printf("read_flags");
        for (uint32_t f : r.read_flags) std::printf(" %08x", f);
        std::printf("\n");
        print_f32("hart", r.prob_homopolymer_artifact);
        print_f32("hvar", r.prob_homopolymer_variant);
        std::printf("third");
        for (size_t i = 0; i < third.has.size(); ++i) std::printf(" %s", third.has[i] ? std::to_string(third.value[i]).c_str() : ".");
        std::printf("\n--\n");
        info.clear();
    };
    while (std::getline(in, line)) {
        if (line == "--") {
            flush();
            continue;
        }
        std::istringstream ss(line);
        std::string tag;
        ss >> tag;
        std::vector<int32_t> v;
        long x;
        while (ss >> x) v.push_back((int32_t)x);
        info[tag] = v;
    }
    flush();
    return 0;
}
