// vlr_obs_codec.hpp — C++17 decoder of varlociraptor's observation wire format into the columns the caller packs
// (ObservationRecord, host/vlr_caller.hpp), without materialising per-read structs.
//
// The reference writes every per-read field of a record as one INFO tag: the bincode-1.3 serialisation of a Vec,
// split into little-endian u16 words stored as BCF integers (`write_observations`, src/variants/evidence/observations/
// pileup.rs / preprocessing/mod.rs:978-1000) and reads them back in `read_observations` (preprocessing/mod.rs:818-919).
// Layouts: Vec<T> = u64 length + items; MiniLogProb = u32 variant (0: IEEE half, 1: f32) + payload
// (utils/mod.rs:448-474); plain enums = u32 variant index; Option<T> = u8 tag + payload; BitVec<u8> = u8 some-tag,
// u64 #blocks, blocks, u64 #bits (bit i = block[i / 8] >> (i % 8) & 1).
// Mirrors varlociraptor_b200/obs_codec.py (decode_record); tests/test_host_cpp.py compares the two bit for bit on the
// reference's own records.
#pragma once

#include <cstring>

#include "vlr_caller.hpp"

namespace vlr {

using InfoArrays = std::map<std::string, std::vector<int32_t>>; // INFO tag -> integer values as htslib hands them out

class WireReader {
  public:
    explicit WireReader(const std::vector<int32_t>& words) {
        bytes_.reserve(words.size() * 2);
        for (int32_t w : words) {
            bytes_.push_back((uint8_t)(w & 0xff));
            bytes_.push_back((uint8_t)((w >> 8) & 0xff));
        }
    }
    uint8_t u8() { return take<uint8_t>(); }
    uint32_t u32() { return take<uint32_t>(); }
    uint64_t u64() { return take<uint64_t>(); }
    float mini_logprob() {
        const uint32_t variant = u32();
        if (variant == 0) return half_to_float(take<uint16_t>());
        if (variant == 1) return take<float>();
        throw std::runtime_error("invalid MiniLogProb variant " + std::to_string(variant));
    }
    const uint8_t* raw(size_t n) {
        need(n);
        const uint8_t* p = bytes_.data() + at_;
        at_ += n;
        return p;
    }
    static float half_to_float(uint16_t h) {
        const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
        uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, bits;
        if (exp == 0) {
            if (man == 0) {
                bits = sign;
            } else { // subnormal half: normalise
                int e = -1;
                do {
                    man <<= 1;
                    ++e;
                } while (!(man & 0x400u));
                bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
            }
        } else if (exp == 31) {
            bits = sign | 0x7f800000u | (man << 13);
        } else {
            bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
        }
        float f;
        std::memcpy(&f, &bits, 4);
        return f;
    }

  private:
    template <class T> T take() {
        need(sizeof(T));
        T v;
        std::memcpy(&v, bytes_.data() + at_, sizeof(T)); // the wire is little-endian, like every target of this library
        at_ += sizeof(T);
        return v;
    }
    void need(size_t n) const {
        if (at_ + n > bytes_.size()) throw std::runtime_error("truncated observation INFO array");
    }
    std::vector<uint8_t> bytes_;
    size_t at_ = 0;
};

inline std::vector<float> decode_mini_logprobs(const std::vector<int32_t>& words) {
    WireReader r(words);
    std::vector<float> out((size_t)r.u64());
    for (float& v : out) v = r.mini_logprob();
    return out;
}
inline std::vector<float> decode_optional_mini_logprobs(const std::vector<int32_t>& words) { // None -> NaN
    WireReader r(words);
    std::vector<float> out((size_t)r.u64(), NAN);
    for (float& v : out)
        if (r.u8()) v = r.mini_logprob();
    return out;
}
inline std::vector<uint32_t> decode_enum(const std::vector<int32_t>& words) {
    WireReader r(words);
    std::vector<uint32_t> out((size_t)r.u64());
    for (uint32_t& v : out) v = r.u32();
    return out;
}
inline std::vector<bool> decode_bitvec(const std::vector<int32_t>& words) {
    WireReader r(words);
    const uint8_t* blocks = nullptr;
    uint64_t nblocks = 0;
    if (r.u8()) {
        nblocks = r.u64();
        blocks = r.raw((size_t)nblocks);
    }
    const uint64_t nbits = r.u64();
    if (nbits > nblocks * 8) throw std::runtime_error("truncated observation INFO array");
    std::vector<bool> out((size_t)nbits);
    for (uint64_t i = 0; i < nbits; ++i) out[(size_t)i] = (blocks[i / 8] >> (i % 8)) & 1;
    return out;
}
template <class T> inline void decode_optional_ints(const std::vector<int32_t>& words, std::vector<bool>& has, std::vector<T>& val) {
    WireReader r(words);
    const size_t n = (size_t)r.u64();
    has.assign(n, false);
    val.assign(n, T(0));
    for (size_t i = 0; i < n; ++i)
        if (r.u8()) {
            has[i] = true;
            if (sizeof(T) == 1) val[i] = (T)r.u8();
            else val[i] = (T)r.u32();
        }
}

struct ThirdAlleleEvidence { // output-only column (FORMAT/OBS), not part of the engine's batch
    std::vector<bool> has;
    std::vector<uint32_t> value;
};

// read_observations (preprocessing/mod.rs:818-919) for one record: INFO arrays -> per-read columns and packed flags
inline void decode_observation_record(const InfoArrays& info, ObservationRecord& out, ThirdAlleleEvidence* third = nullptr) {
    auto get = [&](const char* tag) -> const std::vector<int32_t>& {
        auto it = info.find(tag);
        if (it == info.end()) throw std::runtime_error(std::string("observation record lacks INFO/") + tag);
        return it->second;
    };
    out.prob_mapping = decode_mini_logprobs(get("PROB_MAPPING"));
    out.prob_ref = decode_mini_logprobs(get("PROB_REF"));
    out.prob_alt = decode_mini_logprobs(get("PROB_ALT"));
    out.prob_missed_allele = decode_mini_logprobs(get("PROB_MISSED_ALLELE"));
    out.prob_sample_alt = decode_mini_logprobs(get("PROB_SAMPLE_ALT"));
    out.prob_double_overlap = decode_mini_logprobs(get("PROB_DOUBLE_OVERLAP"));
    out.prob_hit_base = decode_mini_logprobs(get("PROB_HIT_BASE"));
    const size_t n = out.prob_mapping.size();
    const auto strand = decode_enum(get("STRAND")), orient = decode_enum(get("READ_ORIENTATION")),
               readpos = decode_enum(get("READ_POSITION")), altlocus = decode_enum(get("ALT_LOCUS"));
    const auto softclipped = decode_bitvec(get("SOFTCLIPPED")), paired = decode_bitvec(get("PAIRED")),
               max_mapq = decode_bitvec(get("IS_MAX_MAPQ"));
    for (const auto* v : {&out.prob_ref, &out.prob_alt, &out.prob_missed_allele, &out.prob_sample_alt,
                          &out.prob_double_overlap, &out.prob_hit_base})
        if (v->size() != n) throw std::runtime_error("observation INFO arrays differ in length");
    if (strand.size() != n || orient.size() != n || readpos.size() != n || altlocus.size() != n || softclipped.size() < n ||
        paired.size() < n || max_mapq.size() < n)
        throw std::runtime_error("observation INFO arrays differ in length");
    out.read_flags.assign(n, 0u);
    for (size_t i = 0; i < n; ++i) {
        uint32_t f = (strand[i] << VLR_RF_STRAND_SHIFT) | (orient[i] << VLR_RF_ORIENT_SHIFT) | (altlocus[i] << VLR_RF_ALTLOCUS_SHIFT);
        if (readpos[i] == 0) f |= VLR_RF_READPOS_MAJOR; // ReadPosition::Major = 0, Some = 1
        if (softclipped[i]) f |= VLR_RF_SOFTCLIPPED;
        if (paired[i]) f |= VLR_RF_PAIRED;
        if (max_mapq[i]) f |= VLR_RF_MAX_MAPQ;
        out.read_flags[i] = f;
    }
    out.prob_homopolymer_artifact.clear();
    out.prob_homopolymer_variant.clear();
    // is_homopolymer_indel = the artifact tag is present (preprocessing/mod.rs:874)
    if (info.count("PROB_HOMOPOLYMER_ARTIFACT_OBSERVABLE")) {
        out.prob_homopolymer_artifact = decode_optional_mini_logprobs(get("PROB_HOMOPOLYMER_ARTIFACT_OBSERVABLE"));
        out.prob_homopolymer_variant = decode_optional_mini_logprobs(get("PROB_HOMOPOLYMER_VARIANT_OBSERVABLE"));
        std::vector<bool> has;
        std::vector<uint8_t> len;
        decode_optional_ints<uint8_t>(get("HOMOPOLYMER_INDEL_LEN"), has, len);
        for (size_t i = 0; i < n && i < has.size(); ++i)
            if (has[i]) out.read_flags[i] |= VLR_RF_HAS_HOMOPOLYMER_LEN | ((uint32_t)len[i] << VLR_RF_HOMOPOLYMER_LEN_SHIFT);
    }
    if (third) {
        third->has.clear();
        third->value.clear();
        if (info.count("THIRD_ALLELE_EVIDENCE")) decode_optional_ints<uint32_t>(get("THIRD_ALLELE_EVIDENCE"), third->has, third->value);
    }
}

} // namespace vlr
