// team_alignment.cpp -- C++ drop-in wrappers over the C ABI (include/b200map.h).
// Mirrors the interface of reference team_alignment/team_alignment.hpp:14-23; the arithmetic
// lives in the CUDA kernels, this file only marshals arguments and translates error codes back
// into the exceptions the reference throws (team_alignment.cpp:73, :347).
#include "team_alignment.hpp"

#include <stdexcept>

#include "b200map.h"

namespace team {

static void raise(int rc) {
    const std::string msg = b200_last_error();
    if (rc == B200_E_TYPE) throw std::invalid_argument("Unknown AlignmentType provided.");
    if (rc == B200_E_NOMEM) throw std::bad_alloc();
    throw std::runtime_error("b200map: " + msg);
}

std::vector<int> AlignBatch(const std::vector<AlignJob>& jobs, AlignmentType type, int match, int mismatch,
                            int gap, std::vector<std::string>* cigars, std::vector<unsigned int>* target_begins,
                            int device) {
    const size_t n = jobs.size();
    std::vector<const char*> q(n), t(n);
    std::vector<uint32_t> ql(n), tl(n);
    uint64_t bound = 0;
    for (size_t i = 0; i < n; ++i) {
        q[i] = jobs[i].query; t[i] = jobs[i].target;
        ql[i] = jobs[i].query_len; tl[i] = jobs[i].target_len;
        bound += 2ull * ((uint64_t)ql[i] + tl[i]) + 2;
    }
    std::vector<int32_t> score(n);
    std::vector<uint32_t> tb(n);
    std::vector<char> cig;
    std::vector<uint64_t> off;
    if (cigars) { cig.resize(bound); off.resize(n + 1); }
    const int rc = b200_align_batch(device, n, q.data(), ql.data(), t.data(), tl.data(), static_cast<int>(type), match,
                                    mismatch, gap, score.data(), tb.data(), cigars ? cig.data() : nullptr,
                                    cigars ? off.data() : nullptr, cigars ? bound : 0);
    if (rc != B200_OK) raise(rc);
    if (cigars) {
        cigars->resize(n);
        for (size_t i = 0; i < n; ++i) (*cigars)[i].assign(cig.data() + off[i], cig.data() + off[i + 1]);  // may hold a NUL
    }
    if (target_begins) target_begins->assign(tb.begin(), tb.end());
    return std::vector<int>(score.begin(), score.end());
}

int Align(const char* query, unsigned int query_len, const char* target, unsigned int target_len,
          AlignmentType type, int match, int mismatch, int gap, std::string* cigar, unsigned int* target_begin) {
    const std::vector<AlignJob> one{{query, query_len, target, target_len}};
    std::vector<std::string> cg;
    std::vector<unsigned int> tb;
    const std::vector<int> s = AlignBatch(one, type, match, mismatch, gap, cigar ? &cg : nullptr, &tb);
    if (cigar) *cigar = cg[0];
    if (target_begin) *target_begin = tb[0];
    return s[0];
}

}  // namespace team
