// Links host/vlr_contamination.hpp against the real libvlr_engine.so. Without a CUDA device the C-ABI entry must
// refuse (VLR_ERR_NO_DEVICE: there is no CPU path) and the estimator must surface that as an error; with a device it
// must produce a normalised posterior (the Simpson rule over every row integrates to one).
#include <cstdio>
#include <sstream>

#include "../../host/vlr_contamination.hpp"

int main() {
    using namespace vlr;
    std::ostringstream table;
    ContaminationEstimator est(table, std::nullopt, &vlr_contamination_posterior, 0);
    Call c;
    c.chrom = "1";
    c.pos = 5;
    c.event_probs = {{"denovo", std::log(0.99)}};
    c.sample_info.resize(2);
    c.sample_info[1] = SampleCall{1.0, 0, {{0.0, -30.0}, {0.5, -2.0}, {1.0, 1.0}}};
    est.process_call(c, {"contaminant", "sample"});
    try {
        est.finalize();
    } catch (const std::runtime_error& e) {
        std::printf("refused: %s\n", e.what());
        return 0;
    }
    double total = 0.0;
    for (const auto& r : est.rows()) {
        const int i = (int)std::lround(r.contamination * 100.0);
        const double w = (i == 0 || i == 100) ? 1.0 : (i % 2 ? 4.0 : 2.0);
        total += w * std::exp(r.ln_posterior) / 300.0;
    }
    std::printf("posterior integrates to %.12f over %zu events\n", total, est.rows().size());
    return std::fabs(total - 1.0) < 1e-9 ? 0 : 1;
}
