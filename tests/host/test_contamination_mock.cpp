// CPU test of host/vlr_contamination.hpp with a recording mock of vlr_contamination_posterior (no GPU, no library).
#define VLR_CALLER_NO_DEFAULT_ENGINE
#include <cassert>
#include <cstdio>
#include <sstream>

#include "../../host/vlr_contamination.hpp"

using namespace vlr;

static vlr_contamination_input_t g_seen;
static std::vector<double> g_prior;
static vlr_status_t mock_posterior(int32_t device, const vlr_contamination_input_t* in, vlr_contamination_output_t* out) {
    assert(device == 3);
    g_seen = *in;
    g_prior.assign(in->ln_prior, in->ln_prior + in->n_grid);
    for (int k = 0; k < in->n_max_vafs; ++k)
        for (int i = 0; i < in->n_grid; ++i) // a peak at contamination 0.2 of the second row; one NaN; ties elsewhere
            out->ln_posterior[k * in->n_grid + i] = (k == 1) ? -std::fabs(i - 20.0) : -200.0;
    out->ln_posterior[7] = NAN;
    *out->ln_marginal = -12.5;
    double m = 0.0;
    for (int64_t o = 0; o < in->n_obs; ++o) m = std::max(m, in->max_posterior_vaf[o]);
    *out->max_vaf = m;
    return VLR_OK;
}

static Call make_call(int64_t pos, double p_denovo, double af, std::vector<std::pair<double, double>> dist, bool has_info = true) {
    Call c;
    c.chrom = "2";
    c.pos = pos;
    c.event_probs = {{"absent", -9.0}, {"denovo", std::log(p_denovo)}, {"other", -9.0}, {"artifact", -INFINITY}};
    c.sample_info.resize(2);
    c.sample_info[0] = SampleCall{0.0, 0, {{0.0, 0.0}}};
    if (has_info) c.sample_info[1] = SampleCall{af, 0, std::move(dist)};
    return c;
}

int main() {
    const std::vector<std::string> names = {"contaminant", "sample"};
    std::ostringstream table;
    ContaminationEstimator est(table, PriorEstimate{0.3, 20}, &mock_posterior, 3);
    est.process_call(make_call(10, 0.97, 0.4, {{0.5, -1.0}, {0.1, -3.0}, {0.5, -2.0}}), names); // kept; sorted, last wins
    est.process_call(make_call(11, 0.94, 0.4, {{0.5, -1.0}}), names);                            // P(denovo) < 0.95
    est.process_call(make_call(12, 0.99, 0.4, {}), names);                                       // artifact MAP: no AFD
    est.process_call(make_call(13, 0.99, 0.4, {{0.5, -1.0}}, false), names);                     // no MAP at all
    est.process_call(make_call(14, 0.999, 0.7, {{0.0, -8.0}, {0.7, 0.5}, {1.0, -4.0}}), names);  // kept
    assert(est.n_observations() == 2);
    assert((est.afd_offsets() == std::vector<int64_t>{0, 2, 5}));
    assert((est.afd_vaf() == std::vector<double>{0.1, 0.5, 0.0, 0.7, 1.0}));
    est.finalize();
    assert(g_seen.n_obs == 2 && g_seen.n_grid == 101 && g_seen.n_max_vafs == 4);
    assert(g_seen.afd_logp[1] == -2.0 && g_seen.prob_denovo[1] == std::log(0.999) && g_seen.max_posterior_vaf[0] == 0.4);
    assert(g_seen.expected_max_somatic_vaf[0] == 0.25 && g_seen.expected_max_somatic_vaf[3] == 1.0);
    // binomial prior with k = round(0.3 * 20) = 6: -inf at the ends, mode at contamination 0.3
    assert(std::isinf(g_prior[0]) && std::isinf(g_prior[100]));
    assert(std::max_element(g_prior.begin(), g_prior.end()) - g_prior.begin() == 30);
    assert(std::fabs(std::exp(g_prior[30]) - binomial_pdf(6, ContaminationEstimator::grid_contamination(30), 20)) < 1e-15);
    assert(std::fabs(binomial_pdf(3, 0.25, 12) - 220.0 * std::pow(0.25, 3) * std::pow(0.75, 9)) < 1e-15);
    assert(est.ln_marginal() == -12.5 && est.max_vaf() == 0.7);
    assert((est.max_vaf_variants() == std::vector<std::pair<std::string, int64_t>>{{"2", 14}}));
    // table: header, 404 rows, best first, NaN last, Rust number formatting
    std::istringstream lines(table.str());
    std::string line;
    std::getline(lines, line);
    assert(line == "maximum somatic VAF\tcontamination\tposterior density");
    std::getline(lines, line);
    assert(line == "0.5\t0.2\t1");
    std::getline(lines, line);
    assert(line == "0.5\t0.19\t0.36787944117144233");
    int n = 2;
    std::string last;
    while (std::getline(lines, line)) {
        ++n;
        last = line;
    }
    assert(n == 404 && last == "0.25\t0.07\tNaN");
    assert(est.rows()[0].contamination == 0.2 && est.rows().size() == 404);
    assert(est.rows()[101].ln_posterior == -200.0 && est.rows()[101].expected_max_somatic_vaf == 0.25 &&
           est.rows()[101].contamination == 0.0 && est.rows()[201].expected_max_somatic_vaf == 0.75); // ties: grid order

    // candidate filter
    ObservationRecord cont, samp;
    for (int i = 0; i < 12; ++i) {
        cont.prob_ref.push_back(-0.01f);
        cont.prob_alt.push_back(-6.0f);
        samp.prob_ref.push_back(i == 5 ? -8.0f : -0.01f);
        samp.prob_alt.push_back(i == 5 ? -0.01f : -6.0f);
    }
    cont.prob_mapping.assign(12, 0.f);
    samp.prob_mapping.assign(12, 0.f);
    WorkItem item;
    item.locus_flags = VLR_LF_HAS_SNV;
    item.pileups = {&cont, &samp};
    ContaminationCandidateFilter f;
    assert(f.filter(item, names));
    item.locus_flags = 0;
    assert(!f.filter(item, names)); // not an SNV
    item.locus_flags = VLR_LF_HAS_SNV;
    cont.prob_alt[3] = 0.0f;
    assert(!f.filter(item, names)); // a contaminant read supports alt
    cont.prob_alt[3] = -6.0f;
    samp.prob_alt[5] = -6.0f;
    samp.prob_ref[5] = -0.01f;
    assert(!f.filter(item, names)); // no strong alt read in the sample
    cont.prob_mapping.resize(9);
    assert(!f.filter(item, names)); // fewer than 10 contaminant reads
    std::printf("host contamination mock test: ok\n");
    return 0;
}
