// vlr_caller.hpp — C++17 host side above the C-ABI (include/vlr_engine.h), mirroring the reference's operator
// interface for `call variants` with the same names, argument meaning and error behaviour:
//
//   call_generic(...)                     src/calling/variants/calling.rs:1022-1116
//   Caller::call()                        calling.rs:320-455   (lock-step reading, filter, batching, ordered delivery)
//   trait CallProcessor                   calling.rs:964-976   -> struct CallProcessor { setup, process_call, finalize }
//   trait CandidateFilter                 calling.rs:1008-1011 -> struct CandidateFilter { filter(work_item, names) }
//   WorkItem                              calling.rs:943-962
//   preprocess_record's per-record flags  calling.rs:513-566   -> locus_flags_for()
//
// The reference is Rust; its toolchain is not available in the build image, so this header is the compiled-language
// host a maintainer can diff against calling.rs. BCF access (htslib) stays with the application: a record arrives here
// as per-read columns (`ObservationRecord`), exactly what `read_observations` (preprocessing/mod.rs:818-919) yields
// before it builds `ReadObservation` structs; vlr_obs_codec.hpp decodes the INFO integer arrays into them.
//
// Errors: like anyhow::Result in the reference, configuration / input problems throw std::runtime_error with the
// reference's message; model-invariant violations arrive as status bits per call (the reference panics there).
#pragma once

#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/vlr_engine.h"

namespace vlr {

// ---- the engine entry points the caller needs (defaults: the C-ABI symbols; tests may inject a recorder)
struct EngineApi {
    vlr_status_t (*ctx_create)(const vlr_scenario_t*, int32_t, vlr_ctx_t**) = nullptr;
    void (*ctx_destroy)(vlr_ctx_t*) = nullptr;
    vlr_status_t (*call_batch)(vlr_ctx_t*, const vlr_batch_t*, vlr_results_t*) = nullptr;
    const char* (*last_error)(const vlr_ctx_t*) = nullptr;
#ifndef VLR_CALLER_NO_DEFAULT_ENGINE
    static EngineApi linked() {
        EngineApi a;
        a.ctx_create = &vlr_ctx_create;
        a.ctx_destroy = &vlr_ctx_destroy;
        a.call_batch = &vlr_call_batch;
        a.last_error = &vlr_last_error;
        return a;
    }
#endif
};

// ---- one sample's observations of one record, decoded (preprocessing/mod.rs:856-880)
struct ObservationRecord {
    std::string chrom;
    int64_t pos = 0; // 0-based like bcf::Record::pos()
    std::string ref, alt;
    bool imprecise = false; // VariantPrecision::Imprecise
    std::vector<float> prob_mapping, prob_ref, prob_alt, prob_missed_allele, prob_sample_alt, prob_double_overlap,
        prob_hit_base;
    std::vector<uint32_t> read_flags; // VLR_RF_* packing of strand / orientation / position / softclip / ...
    std::vector<float> prob_homopolymer_artifact, prob_homopolymer_variant; // empty unless is_homopolymer_indel
    std::optional<float> heterozygosity_phred, somatic_effective_mutation_rate_phred; // INFO overrides
    std::string haplotype; // EVENT / breakend group id, empty if none (HaplotypeIdentifier)
    size_t n_reads() const { return prob_mapping.size(); }
};

// One sample's observation stream (the reference reads an indexed BCF; here: any pull source)
using ObservationSource = std::function<bool(ObservationRecord&)>; // false = end of file

struct Omit { // the six --omit-* switches of `call variants` (cli.rs)
    bool strand_bias = false, read_orientation_bias = false, read_position_bias = false, softclip_bias = false,
         homopolymer_artifact_detection = false, alt_locus_bias = false;
};

// calling.rs:513-566 + src/variants/model/mod.rs VariantType -> variant-type-fraction class
inline uint32_t locus_flags_for(const ObservationRecord& r, bool is_homopolymer_indel, const Omit& omit) {
    bool is_snv_or_mnv, snv;
    if (r.ref.size() == 1 && r.alt.size() == 1) is_snv_or_mnv = snv = true;
    else if (r.ref.size() == r.alt.size()) is_snv_or_mnv = true, snv = false;
    else is_snv_or_mnv = snv = false;
    const bool precise = !r.imprecise;
    uint32_t f = 0;
    if (is_snv_or_mnv && !omit.read_orientation_bias && precise) f |= VLR_LF_CHECK_ROB;
    if (!omit.strand_bias && precise) f |= VLR_LF_CHECK_SB;
    if (is_snv_or_mnv && !omit.read_position_bias && precise) f |= VLR_LF_CHECK_RPB;
    if (is_snv_or_mnv && !omit.softclip_bias && precise) f |= VLR_LF_CHECK_SCB;
    if (is_homopolymer_indel && !omit.homopolymer_artifact_detection) f |= VLR_LF_CHECK_HE;
    if (!omit.alt_locus_bias) f |= VLR_LF_CHECK_ALB;
    if (is_snv_or_mnv && !omit.read_orientation_bias) f |= VLR_LF_FILTER_NONSTANDARD;
    uint32_t vartype = 0; // fraction 1
    const std::string& a = r.alt;
    if (a == "<DEL>" || a == "<INS>" || a == "<REP>") vartype = 1;
    else if (a == "<INV>" || a == "<DUP>" || a == "<BND>" || a.find('[') != std::string::npos ||
             a.find(']') != std::string::npos) vartype = 3;
    else if (!a.empty() && a[0] == '<') vartype = 0;
    else if (snv) vartype = 0;
    else if (is_snv_or_mnv) vartype = 2;
    else vartype = 1;
    f |= vartype << VLR_LF_VARTYPE_SHIFT;
    if (snv) f |= VLR_LF_HAS_SNV | ((uint32_t)(uint8_t)r.ref[0] << VLR_LF_REFBASE_SHIFT) |
                  ((uint32_t)(uint8_t)r.alt[0] << VLR_LF_ALTBASE_SHIFT);
    return f;
}

// ---- what the plugins see
struct WorkItem { // calling.rs:943-962
    size_t index = 0;
    std::string chrom;
    int64_t pos = 0;
    std::string ref, alt, haplotype;
    uint32_t locus_flags = 0;
    std::vector<const ObservationRecord*> pileups; // per sample, nullptr = no observations (calling.rs:605-607)
};

struct SampleCall { // SampleInfo of the final record (calling/variants/mod.rs)
    double allelefreq_estimate = NAN;
    int artifact_config = 0;                         // 0 none, 1..8 = VLR artifact config ids
    std::vector<std::pair<double, double>> vaf_dist; // (vaf, ln posterior density), ascending; empty for artifact MAPs
};

struct Call {
    size_t index = 0;
    std::string chrom;
    int64_t pos = 0;
    std::string ref, alt;
    std::map<std::string, double> event_probs; // event name and "artifact" -> ln posterior (calling.rs:772-799)
    std::vector<std::optional<SampleCall>> sample_info;
    uint32_t status = 0;
    bool adjusted_singleton_evidence() const { return status & VLR_ST_SINGLETON_ADJUSTED; }
    bool filtered_non_standard_alignments() const { return status & VLR_ST_FILTERED_NONSTANDARD; }
    // PROB_* as written to the BCF: PHRED, absolute value, f32 (calling/variants/mod.rs:459-466)
    static float phred(double ln_prob) { return (float)std::fabs(-10.0 * ln_prob / std::log(10.0)); }
};

struct CandidateFilter {
    virtual ~CandidateFilter() = default;
    // Return true if work_item shall be processed, otherwise false.
    virtual bool filter(const WorkItem& work_item, const std::vector<std::string>& sample_names) const = 0;
};
struct DefaultCandidateFilter : CandidateFilter {
    bool filter(const WorkItem&, const std::vector<std::string>&) const override { return true; }
};

class Caller;
struct CallProcessor {
    virtual ~CallProcessor() = default;
    virtual void setup(const Caller&) {}
    virtual void process_call(Call call, const std::vector<std::string>& sample_names) = 0;
    virtual void finalize() {}
};

// ---- the caller
class Caller {
  public:
    // `scenario` is the flattened grammar::Scenario for the current contig (vlr_scenario_t); sample and event
    // names travel separately (events[i].name in the scenario).
    Caller(const vlr_scenario_t& scenario, std::vector<std::string> sample_names,
           std::vector<std::optional<ObservationSource>> observations, Omit omit, CallProcessor& call_processor,
           const CandidateFilter& candidate_filter, EngineApi api, int device = 0, size_t batch_size = 65536,
           int afd_capacity = 128)
        : scenario_(scenario), names_(std::move(sample_names)), obs_(std::move(observations)), omit_(omit),
          processor_(call_processor), filter_(candidate_filter), api_(api), batch_size_(batch_size), afd_(afd_capacity) {
        if ((int)names_.size() != scenario.n_samples || obs_.size() != names_.size())
            throw std::runtime_error("sample names / observations do not match the scenario");
        vlr_status_t st = api_.ctx_create(&scenario_, device, &ctx_);
        if (st != VLR_OK) throw std::runtime_error("vlr_ctx_create failed with status " + std::to_string(st));
        for (int e = 0; e < scenario.n_events; ++e) event_names_.push_back(scenario.events[e].name);
    }
    ~Caller() {
        if (ctx_) api_.ctx_destroy(ctx_);
    }
    Caller(const Caller&) = delete;
    Caller& operator=(const Caller&) = delete;

    const std::vector<std::string>& sample_names() const { return names_; }
    const std::vector<std::string>& event_names() const { return event_names_; }
    size_t n_samples() const { return names_.size(); }

    // Caller::call (calling.rs:320-455)
    void call() {
        processor_.setup(*this);
        size_t index = 0;
        for (;;) {
            // one record per sample in lock-step (calling.rs:353-398)
            std::vector<std::unique_ptr<ObservationRecord>> recs(obs_.size());
            size_t active = 0, eof = 0;
            for (size_t s = 0; s < obs_.size(); ++s) {
                if (!obs_[s]) continue;
                ++active;
                auto r = std::make_unique<ObservationRecord>();
                if ((*obs_[s])(*r)) recs[s] = std::move(r);
                else ++eof;
            }
            if (eof == active) break;
            if (eof) throw std::runtime_error("observation files have different numbers of records");
            const ObservationRecord* first = nullptr;
            for (auto& r : recs)
                if (r) {
                    if (!first) first = r.get();
                    else if (r->chrom != first->chrom || r->pos != first->pos || r->ref != first->ref || r->alt != first->alt)
                        throw std::runtime_error("inconsistent observations: records differ at " + first->chrom + ":" +
                                                 std::to_string(first->pos + 1));
                }
            Pending p;
            p.item.index = index++;
            p.item.chrom = first->chrom;
            p.item.pos = first->pos;
            p.item.ref = first->ref;
            p.item.alt = first->alt;
            p.item.haplotype = first->haplotype;
            bool hom = false;
            for (auto& r : recs) {
                p.item.pileups.push_back(r.get());
                if (r && !r->prob_homopolymer_artifact.empty()) hom = true;
            }
            p.item.locus_flags = locus_flags_for(*first, hom, omit_);
            p.records = std::move(recs);
            if (!filter_.filter(p.item, names_)) continue;
            // breakend / haplotype groups: later members reuse the first member's result (calling.rs:726-741)
            if (!p.item.haplotype.empty()) {
                auto it = group_first_.find(p.item.haplotype);
                if (it != group_first_.end()) p.reuse_of = (int64_t)it->second;
                else group_first_[p.item.haplotype] = p.item.index;
            }
            pending_.push_back(std::move(p));
            if (pending_.size() >= batch_size_) flush();
        }
        flush();
        processor_.finalize();
    }

  private:
    struct Pending {
        WorkItem item;
        std::vector<std::unique_ptr<ObservationRecord>> records;
        int64_t reuse_of = -1;
    };

    void flush() {
        if (pending_.empty()) return;
        const size_t S = names_.size();
        const int E = scenario_.n_events;
        // pack the SoA batch (only records that are computed)
        std::vector<size_t> computed;
        for (size_t i = 0; i < pending_.size(); ++i)
            if (pending_[i].reuse_of < 0) computed.push_back(i);
        const size_t L = computed.size();
        std::vector<int64_t> offsets(L * S + 1, 0);
        std::vector<float> cols[7], hart, hvar, het, semr;
        std::vector<uint32_t> rflags, lflags(L);
        bool any_h = false, any_het = false, any_semr = false;
        for (size_t li = 0; li < L; ++li) {
            const Pending& p = pending_[computed[li]];
            lflags[li] = p.item.locus_flags;
            const ObservationRecord* first = nullptr;
            for (size_t s = 0; s < S; ++s) {
                const ObservationRecord* r = p.item.pileups[s];
                size_t n = r ? r->n_reads() : 0;
                offsets[li * S + s + 1] = offsets[li * S + s] + (int64_t)n;
                if (!r) continue;
                if (!first) first = r;
                const std::vector<float>* src[7] = {&r->prob_mapping, &r->prob_ref, &r->prob_alt, &r->prob_missed_allele,
                                                    &r->prob_sample_alt, &r->prob_double_overlap, &r->prob_hit_base};
                for (int c = 0; c < 7; ++c) {
                    if (src[c]->size() != n) throw std::runtime_error("observation columns of unequal length");
                    cols[c].insert(cols[c].end(), src[c]->begin(), src[c]->end());
                }
                rflags.insert(rflags.end(), r->read_flags.begin(), r->read_flags.end());
                if (!r->prob_homopolymer_artifact.empty()) {
                    any_h = true;
                    hart.resize(rflags.size() - n, NAN);
                    hvar.resize(rflags.size() - n, NAN);
                    hart.insert(hart.end(), r->prob_homopolymer_artifact.begin(), r->prob_homopolymer_artifact.end());
                    hvar.insert(hvar.end(), r->prob_homopolymer_variant.begin(), r->prob_homopolymer_variant.end());
                }
            }
            het.push_back(first && first->heterozygosity_phred ? *first->heterozygosity_phred : NAN);
            semr.push_back(first && first->somatic_effective_mutation_rate_phred
                               ? *first->somatic_effective_mutation_rate_phred : NAN);
            any_het = any_het || (first && first->heterozygosity_phred);
            any_semr = any_semr || (first && first->somatic_effective_mutation_rate_phred);
        }
        if (any_h) {
            hart.resize(rflags.size(), NAN);
            hvar.resize(rflags.size(), NAN);
        }
        vlr_batch_t b{};
        b.n_loci = (int64_t)L;
        b.n_reads = (int64_t)rflags.size();
        b.read_offsets = offsets.data();
        b.prob_mapping = cols[0].data();
        b.prob_ref = cols[1].data();
        b.prob_alt = cols[2].data();
        b.prob_missed_allele = cols[3].data();
        b.prob_sample_alt = cols[4].data();
        b.prob_double_overlap = cols[5].data();
        b.prob_hit_base = cols[6].data();
        b.read_flags = rflags.data();
        b.prob_homopolymer_artifact = any_h ? hart.data() : nullptr;
        b.prob_homopolymer_variant = any_h ? hvar.data() : nullptr;
        b.locus_flags = lflags.data();
        b.locus_heterozygosity_phred = any_het ? het.data() : nullptr;
        b.locus_semr_phred = any_semr ? semr.data() : nullptr;
        std::vector<double> lp(L * (E + 1)), marg(L), mapv(L * S), afdv(L * S * afd_), afdp(L * S * afd_);
        std::vector<int32_t> mapc(L), best(L), afdn(L * S);
        std::vector<uint32_t> status(L), nbase(L);
        vlr_results_t r{};
        r.log_posteriors = lp.data();
        r.log_marginal = marg.data();
        r.map_vaf = mapv.data();
        r.map_config = mapc.data();
        r.best_event = best.data();
        r.status = status.data();
        r.n_base_events = nbase.data();
        r.afd_capacity = afd_;
        r.afd_count = afdn.data();
        r.afd_vaf = afdv.data();
        r.afd_logp = afdp.data();
        if (L > 0) {
            vlr_status_t st = api_.call_batch(ctx_, &b, &r);
            if (st != VLR_OK)
                throw std::runtime_error(std::string("vlr_call_batch failed: ") + (api_.last_error ? api_.last_error(ctx_) : ""));
        }
        // deliver in input order (calling.rs:447); group members copy the first member's result
        std::map<size_t, Call> group_results;
        size_t li = 0;
        for (Pending& p : pending_) {
            Call c;
            if (p.reuse_of >= 0) {
                auto it = results_of_group_.find((size_t)p.reuse_of);
                if (it == results_of_group_.end()) throw std::runtime_error("bug: haplotype group result missing");
                c = it->second;
            } else {
                for (int e = 0; e < E; ++e) c.event_probs[event_names_[e]] = lp[li * (E + 1) + e];
                c.event_probs["artifact"] = lp[li * (E + 1) + E];
                c.status = status[li];
                for (size_t s = 0; s < S; ++s) {
                    if (status[li] & VLR_ST_NO_MAP) {
                        c.sample_info.emplace_back(std::nullopt);
                        continue;
                    }
                    SampleCall sc;
                    sc.allelefreq_estimate = mapv[li * S + s];
                    sc.artifact_config = mapc[li];
                    if (mapc[li] == 0)
                        for (int k = 0; k < afdn[li * S + s]; ++k)
                            sc.vaf_dist.emplace_back(afdv[(li * S + s) * afd_ + k], afdp[(li * S + s) * afd_ + k]);
                    c.sample_info.emplace_back(std::move(sc));
                }
                if (!p.item.haplotype.empty()) results_of_group_[p.item.index] = c;
                ++li;
            }
            c.index = p.item.index;
            c.chrom = p.item.chrom;
            c.pos = p.item.pos;
            c.ref = p.item.ref;
            c.alt = p.item.alt;
            processor_.process_call(std::move(c), names_);
        }
        pending_.clear();
    }

    vlr_scenario_t scenario_;
    std::vector<std::string> names_, event_names_;
    std::vector<std::optional<ObservationSource>> obs_;
    Omit omit_;
    CallProcessor& processor_;
    const CandidateFilter& filter_;
    EngineApi api_;
    vlr_ctx_t* ctx_ = nullptr;
    size_t batch_size_;
    int afd_;
    std::vector<Pending> pending_;
    std::map<std::string, size_t> group_first_;
    std::map<size_t, Call> results_of_group_;
};

// call_generic (calling.rs:1022-1116): observations maps sample name -> source; a sample without an entry has zero
// coverage; a name that is not a scenario sample is errors::Error::InvalidObservationSampleName.
inline void call_generic(const vlr_scenario_t& scenario, const std::vector<std::string>& sample_names,
                         const std::map<std::string, ObservationSource>& observations, bool omit_strand_bias,
                         bool omit_read_orientation_bias, bool omit_read_position_bias, bool omit_softclip_bias,
                         bool omit_homopolymer_artifact_detection, bool omit_alt_locus_bias,
                         CallProcessor& call_processor, const CandidateFilter& candidate_filter, EngineApi api,
                         int device = 0, size_t batch_size = 65536) {
    std::vector<std::optional<ObservationSource>> per_sample(sample_names.size());
    for (const auto& kv : observations) {
        size_t s = 0;
        for (; s < sample_names.size(); ++s)
            if (sample_names[s] == kv.first) break;
        if (s == sample_names.size()) throw std::runtime_error("invalid observation sample name: " + kv.first);
        per_sample[s] = kv.second;
    }
    Omit omit{omit_strand_bias, omit_read_orientation_bias, omit_read_position_bias, omit_softclip_bias,
              omit_homopolymer_artifact_detection, omit_alt_locus_bias};
    Caller caller(scenario, sample_names, std::move(per_sample), omit, call_processor, candidate_filter, api, device,
                  batch_size);
    caller.call();
}

} // namespace vlr
