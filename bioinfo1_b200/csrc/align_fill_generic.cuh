// align_fill_generic.cuh -- the catch-all DP fill: int32 scores, raw byte compare, any
// lengths, any integer scores, optional '-' free-gap semantics.
//
// Reference semantics restated (team_alignment/team_alignment.cpp):
//   border init :62-92, recurrence + tie order (diagonal, left, up) :104-114,
//   local clamp + first-max-in-row-major-order :185-192, semi-global end cell :265-278.
//
// Mapping: one warp per pair. A stripe is 32 lanes x 16 rows; lane L owns rows
// [s*512 + 16L + 1, s*512 + 16L + 16] and at step t works on column j = t - L + 1, so the
// warp sweeps anti-diagonal wavefronts. The up/diagonal dependency between neighbouring
// lanes travels by __shfl_up_sync; between stripes the bottom row is parked in a per-warp
// global scratch row. Each lane emits one 32-bit direction word (16 rows x 2 bit) per
// column and stores four columns at a time as one 16-byte vector.
#pragma once
#include "common.cuh"

namespace b200 {

template <int TYPE, bool DASH>
__device__ __forceinline__ void
align_pair_generic(const uint8_t* __restrict__ q, const uint8_t* __restrict__ t, const PairDesc& pd,
                   uint32_t p, const Scores& sc, uint32_t* __restrict__ dirs, int32_t* my_bnd,
                   int32_t* __restrict__ score, uint32_t* __restrict__ end_i,
                   uint32_t* __restrict__ end_j) {
    constexpr int R = kRowsPerWord;
    constexpr int STRIPE = R * kWarp;
    const int lane = threadIdx.x & 31;
    const int init = (TYPE == 0) ? sc.gap : 0;
    const uint32_t Q = pd.Q, T = pd.T;
    {

        if (Q == 0 || T == 0) {
            // no inner cell: the end cell is on the border (reference :117-118, :96-99, :265-278)
            if (lane == 0) {
                int sco = 0;
                uint32_t ei = 0, ej = 0;
                if (TYPE == 0) { ei = Q; ej = T; sco = (int)((Q + T) * (uint32_t)init); }
                else if (TYPE == 2) { ei = 0; ej = T; }  // first candidate (0,T) wins every tie at 0
                score[p] = sco; end_i[p] = ei; end_j[p] = ej;
            }
            return;
        }

        const uint32_t n_stripes = div_up(Q, STRIPE);
        const uint32_t lq = ((Q - 1) / R) % kWarp;  // lane and register that own row Q
        const uint32_t rq = (Q - 1) % R;

        // running end-cell candidates
        int best = INT_MIN; uint32_t bi = 0, bj = 0;        // local: (max, smallest i, smallest j)
        int colbest = INT_MIN; uint32_t coli = 0;           // semi: last column, smallest i
        int rowbest = INT_MIN; uint32_t rowj = 0;           // semi: last row, smallest j
        int final_h = 0;                                    // global: H(Q,T)
        if (TYPE == 2) {
            if (lane == 0) { colbest = 0; coli = 0; }       // H(0,T) = 0 is the first candidate
            if (lane == (int)lq) { rowbest = 0; rowj = 0; } // H(Q,0) = 0
        }

        for (uint32_t s = 0; s < n_stripes; ++s) {
            const uint32_t rows_here = min((uint32_t)STRIPE, Q - s * STRIPE);
            const uint32_t lanes_used = div_up(rows_here, R);
            const bool last_stripe = (s + 1 == n_stripes);
            const uint32_t i0 = s * STRIPE + lane * R;  // rows i0+1 .. i0+R
            const bool lane_on = (uint32_t)lane < lanes_used;
            const uint32_t rows_valid = lane_on ? min((uint32_t)R, Q - i0) : 0;

            int qc[R], H[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t i = i0 + 1 + r;
                qc[r] = (i <= Q) ? (int)q[i - 1] : 0x100;   // 0x100 never equals a byte
                H[r] = (int)(i * (uint32_t)init);           // column 0 (:83-86)
            }
            int up_prev = (int)(i0 * (uint32_t)init);       // H(i0, 0)
            uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
            uint32_t* drow = dirs ? dirs + pd.dir_off + (uint64_t)(s * kWarp + lane) * pd.pitch : nullptr;

            const uint32_t steps = T + lanes_used - 1;
            for (uint32_t st = 0; st < steps; ++st) {
                const int j = (int)st - lane + 1;
                int from_above = __shfl_up_sync(kFull, H[R - 1], 1);
                const bool active = lane_on && j >= 1 && j <= (int)T;
                if (lane == 0 && active)
                    from_above = (s == 0) ? (int)((uint32_t)j * (uint32_t)init) : my_bnd[j];
                if (active) {
                    const int tc = t[j - 1];
                    const int gl = (DASH && tc == '-') ? 0 : sc.gap;    // indel(target[j-1]) :105
                    int up = from_above, dg = up_prev;
                    uint32_t word = 0;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int gu = (DASH && qc[r] == '-') ? 0 : sc.gap;  // indel(query[i-1]) :106
                        int h = dg + ((qc[r] == tc) ? sc.match : sc.mismatch);
                        uint32_t code = 0;
                        const int l = H[r] + gl;
                        const int u = up + gu;
                        if (l > h) { h = l; code = 1; }
                        if (u > h) { h = u; code = 2; }
                        if (TYPE == 1) {
                            if (h < 0) h = 0;
                            if (h == 0) code = 3;
                            const uint32_t i = i0 + 1 + r;
                            if ((uint32_t)r < rows_valid && (h > best || (h == best && i < bi))) {
                                best = h; bi = i; bj = (uint32_t)j;
                            }
                        }
                        dg = H[r];
                        H[r] = h;
                        up = h;
                        word |= code << (2 * r);
                    }
                    up_prev = from_above;
                    if (lane == kWarp - 1 && !last_stripe) my_bnd[j] = H[R - 1];
                    if (TYPE == 2 && last_stripe && lane == (int)lq) {
                        int hq = H[0];
#pragma unroll
                        for (int r = 1; r < R; ++r) if ((uint32_t)r == rq) hq = H[r];
                        if (hq > rowbest) { rowbest = hq; rowj = (uint32_t)j; }
                    }
                    if (drow) {
                        w0 = w1; w1 = w2; w2 = w3; w3 = word;
                        const uint32_t c = (uint32_t)(j - 1);
                        if ((c & 3u) == 3u) {
                            *reinterpret_cast<uint4*>(drow + (c & ~3u)) = make_uint4(w0, w1, w2, w3);
                        } else if (c + 1 == T) {
                            const uint32_t k = c & 3u;  // k+1 trailing words are valid
                            uint32_t* dst = drow + (c & ~3u);
                            if (k == 0) { dst[0] = w3; }
                            else if (k == 1) { dst[0] = w2; dst[1] = w3; }
                            else { dst[0] = w1; dst[1] = w2; dst[2] = w3; }
                        }
                    }
                }
            }
            // every active lane now holds column T of its rows in H[]
            if (TYPE == 0 && last_stripe && lane == (int)lq) {
                int hq = H[0];
#pragma unroll
                for (int r = 1; r < R; ++r) if ((uint32_t)r == rq) hq = H[r];
                final_h = hq;
            }
            if (TYPE == 2) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if ((uint32_t)r < rows_valid && H[r] > colbest) { colbest = H[r]; coli = i0 + 1 + r; }
            }
            __syncwarp();
        }

        // warp-level reduction of the end-cell candidates with the reference's tie rules
        if (TYPE == 0) {
            final_h = __shfl_sync(kFull, final_h, (int)lq);
            if (lane == 0) { score[p] = final_h; end_i[p] = Q; end_j[p] = T; }
        } else if (TYPE == 1) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int ob = __shfl_xor_sync(kFull, best, o);
                const uint32_t oi = __shfl_xor_sync(kFull, bi, o);
                const uint32_t oj = __shfl_xor_sync(kFull, bj, o);
                if (ob > best || (ob == best && (oi < bi || (oi == bi && oj < bj)))) { best = ob; bi = oi; bj = oj; }
            }
            if (lane == 0) { score[p] = best; end_i[p] = bi; end_j[p] = bj; }
        } else {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int ob = __shfl_xor_sync(kFull, colbest, o);
                const uint32_t oi = __shfl_xor_sync(kFull, coli, o);
                if (ob > colbest || (ob == colbest && oi < coli)) { colbest = ob; coli = oi; }
            }
            rowbest = __shfl_sync(kFull, rowbest, (int)lq);
            rowj = __shfl_sync(kFull, rowj, (int)lq);
            if (lane == 0) {
                if (rowbest > colbest) { score[p] = rowbest; end_i[p] = Q; end_j[p] = rowj; }
                else { score[p] = colbest; end_i[p] = coli; end_j[p] = T; }
            }
        }
    }
}

// kClassGeneric pairs only; bit 0 of flags[p] says whether the pair contains a '-' byte
// (set by classify_kernel), which selects the free-gap variant of the recurrence.
template <int TYPE>
__global__ void __launch_bounds__(128)
fill_generic_kernel(const uint8_t* __restrict__ qbuf, const uint8_t* __restrict__ tbuf,
                    const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ work,
                    uint32_t n_work, uint32_t* __restrict__ work_counter,
                    const uint8_t* __restrict__ flags, uint8_t want_mask, uint8_t want_value, Scores sc,
                    uint32_t* __restrict__ dirs, int32_t* bnd, uint32_t bnd_pitch,
                    int32_t* __restrict__ score, uint32_t* __restrict__ end_i,
                    uint32_t* __restrict__ end_j) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int32_t* my_bnd = bnd + (size_t)warp_global * bnd_pitch;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(work_counter, 1u);
        w = __shfl_sync(kFull, w, 0);
        if (w >= n_work) break;
        const uint32_t p = work[w];
        const uint8_t f = flags[p];
        if ((f & want_mask) != want_value) continue;   // another kernel owns this pair
        const PairDesc pd = pairs[p];
        const uint8_t* q = qbuf + pd.q_off;
        const uint8_t* t = tbuf + pd.t_off;
        if (f & kFlagDash) align_pair_generic<TYPE, true>(q, t, pd, p, sc, dirs, my_bnd, score, end_i, end_j);
        else align_pair_generic<TYPE, false>(q, t, pd, p, sc, dirs, my_bnd, score, end_i, end_j);
    }
}

// One warp per listed pair: flags[p] bit0 = contains '-', bit1 = contains a byte outside "ACGT".
__global__ void __launch_bounds__(256)
classify_kernel(const uint8_t* __restrict__ qbuf, const uint8_t* __restrict__ tbuf,
                const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ work, uint32_t n_work,
                uint8_t* __restrict__ flags) {
    const uint32_t wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wi >= n_work) return;
    const uint32_t p = work[wi];
    const PairDesc pd = pairs[p];
    bool dash = false, other = false;
    for (int which = 0; which < 2; ++which) {
        const uint8_t* s = which ? tbuf + pd.t_off : qbuf + pd.q_off;
        const uint32_t len = which ? pd.T : pd.Q;
        for (uint32_t k = lane; k < len; k += kWarp) {
            const uint8_t c = s[k];
            dash |= (c == '-');
            other |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
        }
    }
    const unsigned d = __ballot_sync(kFull, dash), o = __ballot_sync(kFull, other);
    if (lane == 0) flags[p] = (d ? kFlagDash : 0) | (o ? kFlagNonACGT : 0);
}

}  // namespace b200
