// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A C-ABI veneer over the *unmodified* reference translation units
//   /root/reference/team_alignment/team_alignment.cpp   (team::Align,          :49-350)
//   /root/reference/team_minimizers/team_minimizers.cpp (team::KMER::Minimize, :122-225)
// so that Python (ctypes) and bench.py can drive the reference's own CPU code.
// The reference sources are compiled where they lie (see oracle/Makefile); only
// the resulting shared object lands in oracle/_ref/ (git-ignored).
#include <cstdint>
#include <cstring>
#include <string>
#include <tuple>
#include <vector>
#include <exception>

#include "team_alignment.hpp"
#include "team_minimizers.hpp"

extern "C" {

// Returns 0 on success, -1 if the reference threw. *cigar_len receives the
// full byte length (CIGARs may contain a NUL: the "1\0" empty-path case).
int ref_align(const char* q, uint32_t ql, const char* t, uint32_t tl, int type,
              int match, int mismatch, int gap, int want_cigar,
              int32_t* score, uint32_t* target_begin,
              char* cigar_buf, uint64_t cigar_cap, uint64_t* cigar_len) {
    try {
        std::string cg;
        unsigned int tb = 0xdeadbeefu;
        int s = team::Align(q, ql, t, tl, static_cast<team::AlignmentType>(type), match, mismatch,
                            gap, want_cigar ? &cg : nullptr, &tb);
        *score = s;
        *target_begin = tb;
        if (cigar_len) *cigar_len = cg.size();
        if (want_cigar && cigar_buf) {
            if (cg.size() > cigar_cap) return -2;
            std::memcpy(cigar_buf, cg.data(), cg.size());
        }
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// Packed batch (single thread; bench.py runs one call per host thread on disjoint slices).
// Returns the number of pairs aligned, or -1 if the reference threw.
int64_t ref_align_batch(uint64_t n, const char* qbuf, const uint64_t* qoff, const char* tbuf,
                        const uint64_t* toff, int type, int match, int mismatch, int gap, int want_cigar,
                        int32_t* score, uint32_t* target_begin, uint64_t* cigar_bytes) {
    uint64_t bytes = 0;
    try {
        std::string cg;
        for (uint64_t i = 0; i < n; ++i) {
            unsigned int tb = 0;
            score[i] = team::Align(qbuf + qoff[i], (unsigned)(qoff[i + 1] - qoff[i]), tbuf + toff[i],
                                   (unsigned)(toff[i + 1] - toff[i]), static_cast<team::AlignmentType>(type),
                                   match, mismatch, gap, want_cigar ? &cg : nullptr, &tb);
            target_begin[i] = tb;
            bytes += cg.size();
        }
    } catch (const std::exception&) {
        return -1;
    }
    if (cigar_bytes) *cigar_bytes = bytes;
    return (int64_t)n;
}

// The same batch with every output kept (scores, target_begin, CIGAR bytes back to back with n+1 offsets).
// team::Align is re-entrant (locals only), so tests run one call per host thread on disjoint slices.
// Returns n, -1 if the reference threw, -2 when cigar_cap is too small.
int64_t ref_align_batch_cigar(uint64_t n, const char* qbuf, const uint64_t* qoff, const char* tbuf,
                              const uint64_t* toff, int type, int match, int mismatch, int gap,
                              int32_t* score, uint32_t* target_begin, char* cigar_buf, uint64_t cigar_cap,
                              uint64_t* cigar_off) {
    uint64_t at = 0;
    cigar_off[0] = 0;
    try {
        std::string cg;
        for (uint64_t i = 0; i < n; ++i) {
            unsigned int tb = 0;
            score[i] = team::Align(qbuf + qoff[i], (unsigned)(qoff[i + 1] - qoff[i]), tbuf + toff[i],
                                   (unsigned)(toff[i + 1] - toff[i]), static_cast<team::AlignmentType>(type),
                                   match, mismatch, gap, &cg, &tb);
            target_begin[i] = tb;
            if (at + cg.size() > cigar_cap) return -2;
            std::memcpy(cigar_buf + at, cg.data(), cg.size());
            at += cg.size();
            cigar_off[i + 1] = at;
        }
    } catch (const std::exception&) {
        return -1;
    }
    return (int64_t)n;
}

// Packed batch of Minimize calls on ONE thread (the reference keeps process-global state, team_minimizers.cpp:19-22,
// so calls must not overlap): tuples of sequence i at out_off[i] .. out_off[i+1]. Returns the tuple total, or -2
// when cap is too small (out_off is still filled).
int64_t ref_minimize_batch(uint64_t n, const char* buf, const uint64_t* off, uint32_t k, uint32_t w, int is_fwd,
                           uint32_t* hash, uint32_t* pos, uint8_t* flag, uint64_t cap, uint64_t* out_off) {
    uint64_t at = 0;
    bool over = false;
    out_off[0] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        team::KMER km(is_fwd != 0);
        auto v = km.Minimize(buf + off[i], (unsigned)(off[i + 1] - off[i]), k, w);
        if (at + v.size() > cap) over = true;
        if (!over)
            for (size_t x = 0; x < v.size(); ++x) {
                hash[at + x] = std::get<0>(v[x]);
                pos[at + x] = std::get<1>(v[x]);
                flag[at + x] = std::get<2>(v[x]) ? 1 : 0;
            }
        at += v.size();
        out_off[i + 1] = at;
    }
    return over ? -2 : (int64_t)at;
}

// Two-call protocol: call with cap = 0 to learn the tuple count, then again with
// buffers. `seq` must stay readable for the few bytes past `len` that the
// reference touches when len < k + w - 2: callers pass a
// zero-padded buffer.
int64_t ref_minimize(const char* seq, uint32_t len, uint32_t k, uint32_t w, int is_fwd,
                     uint32_t* hash, uint32_t* pos, uint8_t* flag, uint64_t cap) {
    team::KMER km(is_fwd != 0);
    auto v = km.Minimize(seq, len, k, w);
    if (cap >= v.size()) {
        for (size_t i = 0; i < v.size(); ++i) {
            hash[i] = std::get<0>(v[i]);
            pos[i] = std::get<1>(v[i]);
            flag[i] = std::get<2>(v[i]) ? 1 : 0;
        }
    }
    return static_cast<int64_t>(v.size());
}

// Side state after the last ref_minimize (process-global in the reference,
// team_minimizers.cpp:19-22): number of distinct tuples and of distinct hashes.
void ref_minimize_side_state(uint64_t* n_unique_tuples, uint64_t* n_distinct_hashes) {
    team::KMER km(true);
    *n_unique_tuples = km.GetUniqueMinimizers().size();
    *n_distinct_hashes = km.GetMinimizerFrequencies().size();
}

}  // extern "C"
