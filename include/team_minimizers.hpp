// include/team_minimizers.hpp -- drop-in replacement for the reference header
// team_minimizers/team_minimizers.hpp:1-36: same class, same member signatures, sizeof(KMER) == 1.
//
// KMER::Minimize returns one (hash, 1-based position, strand flag) tuple per window exactly as
// team_minimizers.cpp:122-225 does (begin end-minimizers, full windows, end end-minimizers;
// 2-bit hash C=0 A=1 T=2 G=3 with 32-bit truncation; the all-0xFFFFFFFF window yields (0,0,false)),
// computed by the B200 MinimizeBatch kernel. The getters keep the reference's process-global
// semantics (they describe the LAST Minimize call made through any KMER object,
// team_minimizers.cpp:19-22, 128-129) but are guarded by a mutex instead of being a data race.
#ifndef TEAM_MINIMIZERS_HPP
#define TEAM_MINIMIZERS_HPP

#include <deque>
#include <set>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

namespace team {

class KMER {
private:
    bool is_fwd = true;

public:
    // The flag every tuple of this object's Minimize calls carries: the mapper builds KMER(true) for the
    // reference and the reads and KMER(false) for the reverse complement (team_mapper.cpp:417-424, :599).
    KMER(bool is_fwd_);

    // One tuple per window, duplicates kept: w-1 growing prefix windows, the L-k-w+2 full windows, w-1
    // shrinking suffix windows (L-k+w tuples when L >= k+w-2; empty when L < k or w == 0). Runs on the GPU
    // through b200_minimize_batch with n = 1; throws std::runtime_error when no sm_100 device is visible.
    std::vector<std::tuple<unsigned int, unsigned int, bool>> Minimize(
        const char* sequence, unsigned int sequence_len,
        unsigned int kmer_len,
        unsigned int window_len);

    // Host-side helpers of the reference's public surface (team_minimizers.cpp:44-120), kept for callers that
    // use them directly: 2-bit code <-> text, reverse complement of a k-mer string, leftmost strict minimum of
    // a window (the zero tuple when every hash is 0xFFFFFFFF).
    std::string MappKmerBitToString(unsigned int kmer, unsigned int kmer_len);
    unsigned int MappSeqCharPointerToBit(const char* seq, unsigned int kmer_len);
    std::string ReverseComplement(const std::string& kmer);
    std::tuple<unsigned int, unsigned int, bool> GetTupleWithMinFirst(
        const std::deque<std::tuple<unsigned int, unsigned int, bool>>& window);
    // State of the LAST Minimize call of the process (any object): hash -> number of windows that chose it
    // (only while counting is enabled, the default), and the distinct tuples in ascending order. The index
    // build of the mapper reads both right after minimizing the reference (team_mapper.cpp:420-434).
    std::unordered_map<unsigned int, int> GetMinimizerFrequencies();
    std::set<std::tuple<unsigned int, unsigned int, bool>> GetUniqueMinimizers();
    void SetFrequenciesCount(bool set);
};

// Batched form (new): Minimize for many sequences in one trip to the GPU. is_fwd[i] is the flag
// stamped on sequence i's tuples (nullptr = all true).
struct MinimizeJob {
    const char* sequence;
    unsigned int sequence_len;
    bool is_fwd;
};
std::vector<std::vector<std::tuple<unsigned int, unsigned int, bool>>> MinimizeBatch(
    const std::vector<MinimizeJob>& jobs, unsigned int kmer_len, unsigned int window_len, int device = 0);

}  // namespace team

#endif
