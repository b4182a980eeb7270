// include/team_alignment.hpp -- drop-in replacement for the reference header
// team_alignment/team_alignment.hpp:1-30 (same namespace, enum, signature and defaults), backed
// by the B200 library instead of the CPU matrix fill.
//
// team::Align keeps the reference's observable behaviour (team_alignment.cpp:49-350):
//   * AlignmentType::global / local / semiGlobal = 0 / 1 / 2;
//   * linear gap, integer scores, ties resolved diagonal > left ('I', consumes target) > up
//     ('D', consumes query); a '-' byte makes the gap that consumes it free;
//   * *cigar is ASSIGNED the run-length text ("<count><op>"), the empty path is the 2-byte
//     string "1\0"; *target_begin is 0 for global/semiGlobal and (end column + 1) for local;
//   * an unknown AlignmentType throws std::invalid_argument("Unknown AlignmentType provided.").
// Differences: it needs a visible sm_100 GPU (std::runtime_error otherwise -- there is no CPU
// fallback), and it is one call into the batched C ABI (include/b200map.h) with n = 1, so
// callers that have many pairs should use team::AlignBatch or b200_align_batch directly.
#ifndef TEAM_ALIGNMENT_HPP
#define TEAM_ALIGNMENT_HPP

#include <cstdint>
#include <string>
#include <vector>

namespace team {

enum class AlignmentType {
    global,     // Needleman-Wunsch
    local,      // Smith-Waterman
    semiGlobal  // free end gaps (the reference's "Gotoh" comment is a misnomer: no affine gap)
};

int Align(const char* query, unsigned int query_len,
          const char* target, unsigned int target_len,
          AlignmentType type,
          int match,
          int mismatch,
          int gap,
          std::string* cigar = nullptr,
          unsigned int* target_begin = nullptr);

// Batched form (new): n independent Align() calls sharing type and scores, one trip to the GPU.
// `cigars` / `target_begins` may be nullptr; when given they are resized to n.
struct AlignJob {
    const char* query;
    unsigned int query_len;
    const char* target;
    unsigned int target_len;
};
std::vector<int> AlignBatch(const std::vector<AlignJob>& jobs, AlignmentType type, int match, int mismatch,
                            int gap, std::vector<std::string>* cigars = nullptr,
                            std::vector<unsigned int>* target_begins = nullptr, int device = 0);

}  // namespace team

#endif
