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
    KMER(bool is_fwd_);

    std::vector<std::tuple<unsigned int, unsigned int, bool>> Minimize(
        const char* sequence, unsigned int sequence_len,
        unsigned int kmer_len,
        unsigned int window_len);

    std::string MappKmerBitToString(unsigned int kmer, unsigned int kmer_len);
    unsigned int MappSeqCharPointerToBit(const char* seq, unsigned int kmer_len);
    std::string ReverseComplement(const std::string& kmer);
    std::tuple<unsigned int, unsigned int, bool> GetTupleWithMinFirst(
        const std::deque<std::tuple<unsigned int, unsigned int, bool>>& window);
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
