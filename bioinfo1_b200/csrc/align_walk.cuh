// align_walk.cuh -- traceback over the 2-bit direction matrix and CIGAR emission.
//
// Reference semantics restated (team_alignment/team_alignment.cpp):
//   walk :123-138 (global), :202-217 (local: stop at the first cell whose score is 0),
//   :287-302 (semiGlobal), border parents :83-92 (column 0 -> up 'D', row 0 -> left 'I'),
//   semiGlobal tail pad :306-315, run-length encoding :145-160, empty path => "1\0".
//
// Two kernels, one thread per pair:
//   walk_kernel  follows the directions from the end cell and records the runs it meets
//                (in walk order, i.e. reversed) as packed (count << 2 | op) words, plus the
//                byte length of the final CIGAR text;
//   emit_kernel  (after an exclusive scan of the lengths) prints the runs back to front
//                into the pair's slice of the CIGAR buffer.
#pragma once
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ uint32_t dec_digits(uint32_t v) {
    uint32_t d = 1;
    while (v >= 10) { v /= 10; ++d; }
    return d;
}

template <int TYPE>
__global__ void __launch_bounds__(128)
walk_kernel(const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ work, uint32_t n_work,
            const uint32_t* __restrict__ dirs, const uint32_t* __restrict__ end_i,
            const uint32_t* __restrict__ end_j, uint32_t* __restrict__ runs,
            uint32_t* __restrict__ n_runs, uint32_t* __restrict__ cigar_len) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_work) return;
    const uint32_t p = work[w];
    const PairDesc pd = pairs[p];
    const uint32_t Q = pd.Q, T = pd.T;
    uint32_t i = end_i[p], j = end_j[p];
    uint32_t* out = runs + pd.run_off;
    uint32_t nr = 0, bytes = 0;
    uint32_t cur_op = 3, cur_n = 0;

    auto push = [&](uint32_t op, uint32_t cnt) {
        if (op == cur_op) { cur_n += cnt; return; }
        if (cur_n) { out[nr++] = (cur_n << 2) | cur_op; bytes += dec_digits(cur_n) + 1; }
        cur_op = op; cur_n = cnt;
    };

    if (TYPE == 2) {  // the pad is the tail of the text, so it is the first thing a backward walk meets
        if (i == Q && j < T) push(1, T - j);
        else if (j == T && i < Q) push(2, Q - i);
    }
    const uint32_t* base = dirs + pd.dir_off;
    const bool is_short = (pd.klass & 0xffu) == kClassShort;
    const bool is_long = (pd.klass & 0xffu) == kClassLong;
    const uint32_t s_lane = (pd.klass >> 8) & 31u, s_shift = ((pd.klass >> 16) & 1u) * 16u;
    // cache of the last direction word: an 'up' move usually stays inside the same word
    uint32_t cw = 0;
    uint64_t cw_at = ~0ull;
    for (;;) {
        if (TYPE == 1) {
            if (i == 0 || j == 0) break;  // border score is 0
        } else {
            if (i == 0) { if (j) push(1, j); break; }   // row 0: parents point left
            if (j == 0) { push(2, i); break; }          // column 0: parents point up
        }
        uint64_t at;
        uint32_t sh;
        if (is_short) {   // see align_fill_short.cuh
            const uint32_t b = (i - 1) >> 5, rr = (i - 1) & 31u;
            at = (((uint64_t)b * pd.pitch + (j - 1)) * 32 + s_lane) * 4 + (rr >> 3);
            sh = s_shift + 2 * (7 - (rr & 7u));
        } else if (is_long) {   // see align_fill_long.cuh
            const uint32_t b = (i - 1) >> 5, rr = (i - 1) & 31u;
            at = ((uint64_t)b * pd.pitch + (j - 1)) * 2 + (rr >> 4);
            sh = 2 * (15 - (rr & 15u));
        } else {
            const uint32_t rb = (i - 1) / kRowsPerWord, r = (i - 1) % kRowsPerWord;
            at = (uint64_t)rb * pd.pitch + (j - 1);
            sh = 2 * r;
        }
        if (at != cw_at) { cw = __ldg(base + at); cw_at = at; }
        uint32_t code = (cw >> sh) & 3u;
        if (is_short || is_long) code = (code == 3u) ? 3u : 2u - code;   // stored as tag: 2 diag, 1 left, 0 up
        if (TYPE == 1 && code == 3) break;
        push(code, 1);
        if (code == 0) { --i; --j; }
        else if (code == 1) { --j; }
        else { --i; }
    }
    if (cur_n) { out[nr++] = (cur_n << 2) | cur_op; bytes += dec_digits(cur_n) + 1; }
    if (nr == 0) bytes = 2;  // "1\0"
    n_runs[p] = nr;
    cigar_len[p] = bytes;
}

// Score-only batches still owe the caller target_begin (reference :119-121, :197-199, :283-285).
__global__ void target_begin_kernel(uint32_t n, int type, const uint32_t* __restrict__ end_j,
                                    uint32_t* __restrict__ target_begin) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) target_begin[p] = (type == 1) ? end_j[p] + 1 : 0;
}

__global__ void __launch_bounds__(128)
emit_kernel(const PairDesc* __restrict__ pairs, uint32_t n, const uint32_t* __restrict__ runs,
            const uint32_t* __restrict__ n_runs, const uint64_t* __restrict__ cigar_off,
            char* __restrict__ cigar) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t nr = n_runs[p];
    char* dst = cigar + cigar_off[p];
    if (nr == 0) { dst[0] = '1'; dst[1] = '\0'; return; }
    const uint32_t* src = runs + pairs[p].run_off;
    for (uint32_t k = nr; k-- > 0;) {
        const uint32_t rw = src[k];
        uint32_t cnt = rw >> 2;
        const uint32_t op = rw & 3u;
        const uint32_t nd = dec_digits(cnt);
        for (uint32_t d = nd; d-- > 0;) { dst[d] = (char)('0' + cnt % 10); cnt /= 10; }
        dst[nd] = (op == 0) ? 'M' : (op == 1 ? 'I' : 'D');
        dst += nd + 1;
    }
}

}  // namespace b200
