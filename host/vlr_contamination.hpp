// vlr_contamination.hpp — C++17 host side of `varlociraptor estimate contamination` above the C-ABI entry
// vlr_contamination_posterior (include/vlr_engine.h), mirroring src/estimation/contamination.rs by name:
//
//   PriorEstimate                      contamination.rs:276-280
//   VariantObservation::make           contamination.rs:44-82    (P(denovo) >= 0.95 and an AFD present)
//   Prior::prob                        contamination.rs:137-147  (binomial pdf around the prior estimate)
//   ContaminationEstimator             contamination.rs:282-395  (CallProcessor: collect, then calc_posterior)
//   ContaminationCandidateFilter       contamination.rs:397-419
//
// The second model itself (Likelihood/Marginal, contamination.rs:163-240) runs on the GPU; this header packs the
// observations into the CSR columns of vlr_contamination_input_t as they arrive and writes the posterior table.
#pragma once

#include <algorithm>
#include <cstdio>
#include <ostream>
#include <sstream>

#include "vlr_caller.hpp"

namespace vlr {

struct PriorEstimate {
    double contamination;
    uint32_t n_observed_cells;
};

// gsl_ran_binomial_pdf(k, p, n) (rgsl; library not vendored), restated from its published algorithm
inline double binomial_pdf(uint32_t k, double p, uint32_t n) {
    if (k > n) return 0.0;
    if (p == 0.0) return k == 0 ? 1.0 : 0.0;
    if (p == 1.0) return k == n ? 1.0 : 0.0;
    const double ln_cnk = std::lgamma(n + 1.0) - std::lgamma(k + 1.0) - std::lgamma((double)(n - k) + 1.0);
    return std::exp(ln_cnk + k * std::log(p) + (n - k) * std::log1p(-p));
}

struct ContaminationPrior { // contamination.rs:118-157
    std::optional<PriorEstimate> prior_estimate;
    double prob(double contamination) const {
        if (!prior_estimate) return 0.0; // LogProb::ln_one()
        const uint32_t n = prior_estimate->n_observed_cells;
        const uint32_t k = (uint32_t)std::round(prior_estimate->contamination * n); // f64::round: half away from zero
        return std::log(binomial_pdf(k, contamination, n));
    }
};

using ContaminationFn = vlr_status_t (*)(int32_t, const vlr_contamination_input_t*, vlr_contamination_output_t*);

struct ContaminationTableRow {
    double expected_max_somatic_vaf, contamination, ln_posterior;
};

class ContaminationEstimator : public CallProcessor {
  public:
    static constexpr int N_GRID = 101; // ln_simpsons_integrate_exp(density, 0.0, 1.0, 101)
    ContaminationEstimator(std::ostream& output, std::optional<PriorEstimate> prior_estimate, ContaminationFn fn,
                           int32_t device = 0)
        : out_(output), prior_{prior_estimate}, fn_(fn), device_(device) {
        offsets_.push_back(0);
    }

    // VariantObservation::new + push (contamination.rs:44-82, 379-385): the BTreeMap becomes a sorted CSR row
    void process_call(Call call, const std::vector<std::string>& sample_names) override {
        const size_t s = (size_t)(std::find(sample_names.begin(), sample_names.end(), "sample") - sample_names.begin());
        if (s >= call.sample_info.size()) throw std::runtime_error("invalid observation sample name: sample");
        const auto& info = call.sample_info[s];
        const double prob_denovo = call.event_probs.at("denovo");
        if (!info || info->vaf_dist.empty() || std::exp(prob_denovo) < 0.95) return; // no denovo variant, skip
        std::map<double, double> dist; // repeated keys keep the last value, like collect() into a BTreeMap
        for (const auto& kv : info->vaf_dist) dist[kv.first] = kv.second;
        for (const auto& kv : dist) {
            afd_vaf_.push_back(kv.first);
            afd_logp_.push_back(kv.second);
        }
        offsets_.push_back((int64_t)afd_vaf_.size());
        prob_denovo_.push_back(prob_denovo);
        max_posterior_vaf_.push_back(info->allelefreq_estimate);
        loci_.emplace_back(call.chrom, call.pos);
    }

    // calc_posterior (contamination.rs:308-377): the table "maximum somatic VAF, contamination, posterior density",
    // highest posterior first (ties / NaNs: grid order, last)
    void finalize() override {
        static const double emsv[4] = {0.25, 0.5, 0.75, 1.0};
        std::vector<double> ln_prior(N_GRID), post(4 * N_GRID), lik(4 * N_GRID);
        for (int i = 0; i < N_GRID; ++i) ln_prior[i] = prior_.prob(grid_contamination(i));
        vlr_contamination_input_t in{};
        in.n_obs = (int64_t)prob_denovo_.size();
        in.prob_denovo = prob_denovo_.data();
        in.max_posterior_vaf = max_posterior_vaf_.data();
        in.afd_offsets = offsets_.data();
        in.afd_vaf = afd_vaf_.data();
        in.afd_logp = afd_logp_.data();
        in.n_grid = N_GRID;
        in.n_max_vafs = 4;
        in.expected_max_somatic_vaf = emsv;
        in.ln_prior = ln_prior.data();
        vlr_contamination_output_t o{};
        o.ln_posterior = post.data();
        o.ln_likelihood = lik.data();
        o.ln_marginal = &ln_marginal_;
        o.max_vaf = &max_vaf_;
        const vlr_status_t st = fn_(device_, &in, &o);
        if (st != VLR_OK) throw std::runtime_error("vlr_contamination_posterior failed with status " + std::to_string(st));
        rows_.clear();
        for (int k = 0; k < 4; ++k)
            for (int i = 0; i < N_GRID; ++i) rows_.push_back({emsv[k], grid_contamination(i), post[k * N_GRID + i]});
        std::stable_sort(rows_.begin(), rows_.end(), [](const ContaminationTableRow& a, const ContaminationTableRow& b) {
            const bool na = std::isnan(a.ln_posterior), nb = std::isnan(b.ln_posterior);
            if (na != nb) return nb;
            return !na && a.ln_posterior > b.ln_posterior;
        });
        out_ << "maximum somatic VAF\tcontamination\tposterior density\n";
        for (const auto& r : rows_)
            out_ << fmt(r.expected_max_somatic_vaf) << '\t' << fmt(r.contamination) << '\t' << fmt(std::exp(r.ln_posterior))
                 << '\n';
    }

    // the variants whose MAP VAF is the maximum (output_max_vaf_variants, contamination.rs:350-361)
    std::vector<std::pair<std::string, int64_t>> max_vaf_variants() const {
        std::vector<std::pair<std::string, int64_t>> v;
        for (size_t i = 0; i < loci_.size(); ++i)
            if (max_posterior_vaf_[i] == max_vaf_) v.push_back(loci_[i]);
        return v;
    }

    static double grid_contamination(int i) { return 0.0 + (double)i * ((1.0 - 0.0) / (double)(N_GRID - 1)); }
    size_t n_observations() const { return prob_denovo_.size(); }
    const std::vector<int64_t>& afd_offsets() const { return offsets_; }
    const std::vector<double>& afd_vaf() const { return afd_vaf_; }
    const std::vector<ContaminationTableRow>& rows() const { return rows_; }
    double ln_marginal() const { return ln_marginal_; }
    double max_vaf() const { return max_vaf_; }

  private:
    static std::string fmt(double x) { // Rust's `{}` for f64: shortest round-trip digits, never an exponent
        if (std::isnan(x)) return "NaN";
        if (std::isinf(x)) return x > 0 ? "inf" : "-inf";
        char buf[64];
        for (int prec = 1; prec <= 17; ++prec) { // shortest %.{prec}g that reads back exactly
            std::snprintf(buf, sizeof buf, "%.*g", prec, x);
            if (std::strtod(buf, nullptr) == x) break;
        }
        std::string s(buf);
        const size_t e = s.find('e');
        if (e == std::string::npos) return s;
        // expand the exponent form
        int exp10 = std::atoi(s.c_str() + e + 1);
        std::string mant = s.substr(0, e), sign;
        if (!mant.empty() && mant[0] == '-') {
            sign = "-";
            mant.erase(0, 1);
        }
        const size_t dot = mant.find('.');
        std::string digits = mant;
        int point = (int)(dot == std::string::npos ? mant.size() : dot);
        if (dot != std::string::npos) digits.erase(dot, 1);
        point += exp10;
        if (point <= 0) return sign + "0." + std::string((size_t)-point, '0') + digits;
        if ((size_t)point >= digits.size()) return sign + digits + std::string((size_t)point - digits.size(), '0');
        return sign + digits.substr(0, (size_t)point) + "." + digits.substr((size_t)point);
    }

    std::ostream& out_;
    ContaminationPrior prior_;
    ContaminationFn fn_;
    int32_t device_;
    std::vector<double> prob_denovo_, max_posterior_vaf_, afd_vaf_, afd_logp_;
    std::vector<int64_t> offsets_;
    std::vector<std::pair<std::string, int64_t>> loci_;
    std::vector<ContaminationTableRow> rows_;
    double ln_marginal_ = 0.0, max_vaf_ = 0.0;
};

// contamination.rs:397-419: SNVs whose contaminant reads (>= 10) all support the reference and whose sample pileup
// (>= 10 reads) has at least one strong alt read (Kass-Raftery: exp(prob_alt - prob_ref) > 20)
struct ContaminationCandidateFilter : CandidateFilter {
    bool filter(const WorkItem& item, const std::vector<std::string>& sample_names) const override {
        auto idx = [&](const char* name) {
            return (size_t)(std::find(sample_names.begin(), sample_names.end(), name) - sample_names.begin());
        };
        const size_t c = idx("contaminant"), s = idx("sample");
        if (c >= item.pileups.size() || s >= item.pileups.size()) throw std::runtime_error("invalid observation sample name");
        if (!(item.locus_flags & VLR_LF_HAS_SNV)) return false;
        const ObservationRecord* cp = item.pileups[c];
        const ObservationRecord* sp = item.pileups[s];
        if (!cp || cp->n_reads() < 10 || !sp || sp->n_reads() < 10) return false;
        for (size_t i = 0; i < cp->n_reads(); ++i)
            if (!((double)cp->prob_ref[i] > (double)cp->prob_alt[i])) return false; // is_ref_support
        for (size_t i = 0; i < sp->n_reads(); ++i)
            if (std::exp((double)sp->prob_alt[i] - (double)sp->prob_ref[i]) > 20.0) return true; // is_strong_alt_support
        return false;
    }
};

} // namespace vlr
