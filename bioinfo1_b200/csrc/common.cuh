// common.cuh -- shared device/host definitions for the B200 alignment + minimizer kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// Direction matrix layout (HBM), shared by every fill kernel and the walker:
//   one 32-bit word holds the 2-bit codes of kRowsPerWord consecutive rows at one column;
//   word(rowblock rb, column j) = dirs[dir_off + rb * pitch + (j-1)], pitch = roundup4(T).
//   code: 0 diagonal ('M'), 1 left ('I', consumes target), 2 up ('D', consumes query),
//         3 stop (local only: cell score == 0, reference team_alignment.cpp:202).
constexpr int kRowsPerWord = 16;

struct PairDesc {
    uint64_t q_off;    // byte offset of the query in the query buffer
    uint64_t t_off;    // byte offset of the target in the target buffer
    uint64_t dir_off;  // word offset of this pair's direction matrix inside the wave buffer
    uint64_t run_off;  // element offset of this pair's run slots inside the run scratch
    uint64_t qpk_off;  // word offset of the query's 2-bit copy (16 bases per word) in the packed query buffer
    uint64_t tpk_off;  // same for the target
    uint32_t Q, T;
    uint32_t pitch;    // words per row block
    uint32_t klass;    // bits 0-7 kernel class (kClass*), short class: bits 8-12 lane, bit 16 half
};
static_assert(sizeof(PairDesc) == 64, "PairDesc layout");

constexpr uint32_t kClassGeneric = 0;  // align_fill_generic.cuh layout: word(rb, j) = dirs[dir_off + rb*pitch + j-1]
constexpr uint32_t kClassShort = 1;    // align_fill_short.cuh layout, pitch = column count of the 64-pair group
constexpr uint32_t kClassLong = 2;     // align_fill_long.cuh layout: 2 words per (32-row block, column), pitch even
constexpr uint32_t kClassLong16 = 3;   // align_fill_long16.cuh layout: 4 words per (64-row block, slot), slot = column-1 + (row half)

constexpr uint8_t kFlagDash = 1;     // pair contains a '-' byte (free gap, team_alignment.cpp:25-28)
constexpr uint8_t kFlagNonACGT = 2;  // pair contains a byte outside "ACGT"

// Geometry of the fill kernels' register blocks and direction layouts (the planner sizes waves with them).
constexpr int kShortRows = 32;                       // align_fill_short.cuh: rows per register block
constexpr int kShortThreads = 64;                    //   2 warps per CTA, every warp independent
constexpr int kLongRows = 32;                        // align_fill_long.cuh: rows per lane (stripe = 32 lanes x 32 rows)
constexpr int kL16LaneRows = 64;                     // align_fill_long16.cuh: query rows per lane (two blocks of 32)
constexpr int kL16Stripe = kL16LaneRows * kWarp;     //   2048 rows per stripe
constexpr int kL16Chunk = 64;                        //   steps between progress publications / polls / re-centring

struct ShortGroup {   // one per 64-pair group of the short class (one warp's worth of work)
    uint64_t dir_off;  // word offset of the group's direction block inside the wave buffer
    uint32_t cols;     // Tg: columns per row block = max T over the group
    uint32_t pad;
};

struct Scores {
    int match, mismatch, gap;
};

__host__ __device__ inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
__host__ __device__ inline uint64_t div_up64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

}  // namespace b200
