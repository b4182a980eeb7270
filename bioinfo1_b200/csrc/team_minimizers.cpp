// team_minimizers.cpp -- C++ drop-in wrappers for team::KMER over the C ABI (include/b200map.h).
// Mirrors reference team_minimizers/team_minimizers.hpp:13-33. Minimize() itself is one call to
// the MinimizeBatch kernel; the small string helpers are plain host code.
#include "team_minimizers.hpp"

#include <cstdint>
#include <mutex>
#include <stdexcept>

#include "b200map.h"

namespace team {

using Tuple = std::tuple<unsigned int, unsigned int, bool>;

// The reference keeps this state in namespace-scope globals shared by every KMER object
// (team_minimizers.cpp:19-22); callers (team_mapper.cpp:420, :429, :433-434) rely on "the last
// Minimize wins". Same contract here, minus the data race. The histogram / distinct set are
// derived lazily from the last result.
namespace {
std::mutex g_mu;
bool g_count_frequencies = true;   // SetFrequenciesCount
bool g_last_counted = true;        // value of the flag when the last Minimize ran
std::vector<Tuple> g_last;
}  // namespace

KMER::KMER(bool is_fwd_) : is_fwd(is_fwd_) {}

std::vector<std::vector<Tuple>> MinimizeBatch(const std::vector<MinimizeJob>& jobs, unsigned int k, unsigned int w,
                                              int device) {
    const size_t n = jobs.size();
    std::vector<const char*> seq(n);
    std::vector<uint32_t> len(n);
    std::vector<uint8_t> fwd(n);
    uint64_t total = 0;
    for (size_t i = 0; i < n; ++i) {
        seq[i] = jobs[i].sequence; len[i] = jobs[i].sequence_len; fwd[i] = jobs[i].is_fwd ? 1 : 0;
        total += b200_minimize_count(len[i], k, w);
    }
    std::vector<uint32_t> hash(total ? total : 1), pos(total ? total : 1);
    std::vector<uint8_t> flag(total ? total : 1);
    std::vector<uint64_t> off(n + 1, 0);
    const int rc = b200_minimize_batch(device, n, seq.data(), len.data(), k, w, fwd.data(), hash.data(), pos.data(),
                                       flag.data(), off.data(), total);
    if (rc == B200_E_NOMEM) throw std::bad_alloc();
    if (rc != B200_OK) throw std::runtime_error(std::string("b200map: ") + b200_last_error());
    std::vector<std::vector<Tuple>> out(n);
    for (size_t i = 0; i < n; ++i) {
        out[i].reserve(off[i + 1] - off[i]);
        for (uint64_t x = off[i]; x < off[i + 1]; ++x) out[i].emplace_back(hash[x], pos[x], flag[x] != 0);
    }
    return out;
}

std::vector<Tuple> KMER::Minimize(const char* sequence, unsigned int sequence_len, unsigned int kmer_len,
                                  unsigned int window_len) {
    std::vector<Tuple> res;
    if (!(sequence_len < kmer_len || window_len == 0))
        res = std::move(MinimizeBatch({{sequence, sequence_len, is_fwd}}, kmer_len, window_len)[0]);
    std::lock_guard<std::mutex> lk(g_mu);
    g_last = res;
    g_last_counted = g_count_frequencies;
    return res;
}

std::unordered_map<unsigned int, int> KMER::GetMinimizerFrequencies() {
    std::lock_guard<std::mutex> lk(g_mu);
    std::unordered_map<unsigned int, int> f;
    if (g_last_counted)
        for (const Tuple& t : g_last) f[std::get<0>(t)]++;   // windows that emitted the hash (:189-192)
    return f;
}

std::set<Tuple> KMER::GetUniqueMinimizers() {
    std::lock_guard<std::mutex> lk(g_mu);
    return std::set<Tuple>(g_last.begin(), g_last.end());
}

void KMER::SetFrequenciesCount(bool set) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_count_frequencies = set;
}

// ---- small host helpers of the reference class (not on the hot path) ---------------------------

std::string KMER::MappKmerBitToString(unsigned int kmer, unsigned int kmer_len) {
    // digits '0'..'3' of the 2-bit codes, most significant base first (reference :44-67)
    std::string s(kmer_len, 'X');
    for (unsigned int i = kmer_len; i-- > 0;) { s[i] = static_cast<char>('0' + (kmer & 3u)); kmer >>= 2; }
    return s;
}

unsigned int KMER::MappSeqCharPointerToBit(const char* seq, unsigned int kmer_len) {
    unsigned int h = 0;
    for (unsigned int i = 0; i < kmer_len; ++i) {
        unsigned int c = 0;   // 'C' and every byte outside ACGT
        if (seq[i] == 'A') c = 1; else if (seq[i] == 'T') c = 2; else if (seq[i] == 'G') c = 3;
        h = (h << 2) | c;
    }
    return h;
}

std::string KMER::ReverseComplement(const std::string& kmer) {
    std::string rc(kmer.rbegin(), kmer.rend());
    for (char& c : rc) {
        if (c == 'A') c = 'T'; else if (c == 'T') c = 'A'; else if (c == 'G') c = 'C'; else if (c == 'C') c = 'G';
    }
    return rc;   // other bytes are left as they are, like the reference's switch without default
}

Tuple KMER::GetTupleWithMinFirst(const std::deque<Tuple>& window) {
    Tuple best{};   // (0, 0, false) when every hash is UINT_MAX
    unsigned int mn = 0xFFFFFFFFu;
    for (const Tuple& t : window)
        if (std::get<0>(t) < mn) { mn = std::get<0>(t); best = t; }
    return best;
}

}  // namespace team
