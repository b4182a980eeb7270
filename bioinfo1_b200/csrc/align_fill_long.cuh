// align_fill_long.cuh -- K2: the long-pair DP fill (BASELINE configs 4/5: 8-10 kb reads).
//
// One WARP per STRIPE of a pair. A stripe is 32 lanes x 32 rows = 1024 query rows; lane L owns rows
// [s*1024 + 32L + 1, s*1024 + 32L + 32] and at step t works on column j = t - L + 1, so the warp
// sweeps anti-diagonal wavefronts: the dependency on the lane above travels by __shfl_up_sync,
// the dependency on the stripe above through a boundary row in global memory (L2) plus a progress
// counter, so the stripes of one long pair run concurrently on different warps (and SMs).
//
// Arithmetic: the int32 version of align_fill_short.cuh's tagged moving frame (bit-exact
// restatement of team_alignment.cpp:104-114). Cell value Y = 4*H - 4*gap*j + 1; candidates
//     diagonal Y(i-1,j-1) + 4(s-gap)+1 -> tag 2,  left Y(i,j-1) -> tag 1,  up Y(i-1,j) + 4gap-1 -> tag 0
// so one signed max yields maximum, tie order (diagonal > left > up) and direction. Per cell:
//     PRMT (substitution byte -> sign-extended int32), VIADDMNMX x2, LOP3, IMAD x2
// = 3 alu-pipe + 2 fma-pipe + 1 PRMT issue slots, with no 16-bit range limit.
//
// Direction layout (klass kClassLong): 32 rows x 1 column = two 32-bit words;
//   word k of (block rb32, column j) = dirs[dir_off + (rb32 * pitch + (j-1)) * 2 + k], pitch even;
//   row 16k+x of the block sits at bits [2*(15-x), 2*(15-x)+1] as TAG (2 diag, 1 left, 0 up).
// Each lane stores two columns at a time as one 16-byte vector.
//
// Eligibility (align_plan.cu): pure ACGT content (run-time flags; others fall back to the generic
// kernel), |4(s-gap)+1| <= 127. Local alignments take a second, one-stripe pass (locate_long_kernel).
#pragma once
#include "align_fill_short.cuh"
#include "common.cuh"

namespace b200 {


struct LongConsts {
    uint32_t tab_diff, tab_mis;   // byte tables as in ShortConsts
    int cu;                       // 4*gap - 1
    uint32_t mask, one, four;     // ~3, 1, 4 passed through the constant bank (see ShortConsts)
    int gap, init;
};

__host__ inline LongConsts make_long_consts(const Scores& sc, int type) {
    LongConsts k;
    const int sm = 4 * (sc.match - sc.gap) + 1, sx = 4 * (sc.mismatch - sc.gap) + 1;
    k.tab_diff = ((uint32_t)(uint8_t)(int8_t)sm) ^ ((uint32_t)(uint8_t)(int8_t)sx);
    k.tab_mis = ((uint32_t)(uint8_t)(int8_t)sx) * 0x01010101u;
    k.cu = 4 * sc.gap - 1;
    k.mask = 0xfffffffcu; k.one = 1u; k.four = 4u;
    k.gap = sc.gap;
    k.init = (type == 0) ? sc.gap : 0;
    return k;
}

// Work unit = one STRIPE of one pair. Stripes are handed out in (pair, stripe) order from an atomic
// counter, so a long pair is swept by as many warps as it has stripes, pipelined: stripe s+1 trails
// stripe s by about one chunk of columns. The bottom row of every stripe is parked in its own global
// row (never reused inside a run, so there is no overwrite hazard) together with a progress counter
// the consumer polls once per chunk. Forward progress: a stripe only ever waits for the stripe
// before it in the hand-out order, which is already running on a resident warp.
struct StripeResult {      // per stripe, reduced per pair by finalize_long_kernel
    int colbest; uint32_t coli;     // semi: best of the last column inside this stripe (H units), smallest i
    int rowbest; uint32_t rowj;     // semi: best of row Q (last stripe only), smallest j
    int final_h; uint32_t pad;      // global: H(Q,T) (last stripe only)
};
// local: colbest = the stripe's maximum. fill_long_kernel leaves the cell to locate_long_kernel; fill_long16_kernel
// finds it during the fill and says so in `pad`, with the cell in (coli, rowj).
constexpr uint32_t kStripeLocated = 0x80000000u;

constexpr int kLongChunk = 64;     // columns between progress publications / polls

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int TYPE>
__global__ void __launch_bounds__(128)
fill_long_kernel(const uint32_t* __restrict__ qpk, const uint32_t* __restrict__ tpk,
                 const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ work, uint32_t n_work,
                 const uint32_t* __restrict__ task_off, const uint64_t* __restrict__ bnd_off,
                 uint32_t* __restrict__ work_counter, const uint8_t* __restrict__ flags, LongConsts K,
                 uint32_t* __restrict__ dirs, int32_t* bnd, uint32_t* progress,
                 StripeResult* __restrict__ results, uint32_t* __restrict__ stall_flag) {
    constexpr int R = kLongRows;
    constexpr int STRIPE = R * kWarp;
    const int lane = threadIdx.x & 31;
    const uint32_t MASK = K.mask, ONE = K.one, FOUR = K.four;
    const int frame = 4 * (K.init - K.gap);   // border row 0 in the moving frame: Y(0,j) = frame*j + 1
    const uint32_t n_tasks = task_off[n_work];

    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(work_counter, 1u);
        task = __shfl_sync(kFull, task, 0);
        if (task >= n_tasks) break;
        // which pair owns this stripe: last k with task_off[k] <= task
        uint32_t lo = 0, hi = n_work;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (task_off[mid] <= task) lo = mid; else hi = mid; }
        const uint32_t k = lo, s = task - task_off[k];
        const uint32_t p = work[k];
        if (flags[p]) continue;   // not pure ACGT: the generic kernel owns the whole pair
        const PairDesc pd = pairs[p];
        const uint32_t Q = pd.Q, T = pd.T;
        const uint32_t* qw = qpk + pd.qpk_off;
        const uint32_t* tw_base = tpk + pd.tpk_off;
        const uint32_t n_stripes = div_up(Q, STRIPE);
        const uint32_t lq = ((Q - 1) / R) % kWarp, rq = (Q - 1) % R;   // lane / register of row Q
        const uint32_t row_pitch = T + 4;
        int32_t* row_out = bnd + bnd_off[k] + (uint64_t)s * row_pitch;
        const int32_t* row_in = row_out - row_pitch;             // written by stripe s-1 (valid when s > 0)
        uint32_t* prog_out = progress + task;
        const uint32_t* prog_in = progress + task - 1;

        const uint32_t rows_here = min((uint32_t)STRIPE, Q - s * STRIPE);
        const uint32_t lanes_used = div_up(rows_here, R);
        const bool last_stripe = (s + 1 == n_stripes);
        const uint32_t i0 = s * STRIPE + lane * R;
        const bool lane_on = (uint32_t)lane < lanes_used;
        const uint32_t rows_valid = lane_on ? min((uint32_t)R, Q - i0) : 0;

        int colbest = INT_MIN; uint32_t coli = 0;
        int rowbest = INT_MIN; uint32_t rowj = 0;
        int lbest = INT_MIN;                         // local: best H seen by this lane
        if (TYPE == 2) {
            if (s == 0 && lane == 0) { colbest = 0; coli = 0; }               // H(0,T) = 0 comes first
            if (last_stripe && lane == (int)lq) { rowbest = 0; rowj = 0; }    // H(Q,0) = 0
        }

        uint32_t sel[R];
        int Y[R];
        {
            const uint32_t q0 = lane_on ? qw[i0 >> 4] : 0u, q1 = lane_on ? qw[(i0 >> 4) + 1] : 0u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t c = ((r < 16 ? q0 : q1) >> (2 * (r & 15))) & 3u;
                sel[r] = c * 0x1111u + 0x8880u;   // byte0 = tab[c], bytes 1..3 = its sign
                Y[r] = 4 * (int)((i0 + 1 + r) * (uint32_t)K.init) + 1;   // column 0, frame 0
            }
        }
        int up_prev = 4 * (int)(i0 * (uint32_t)K.init) + 1;   // Y(i0, 0)
        uint32_t tw = 0, tw_next = lane_on ? tw_base[0] : 0u;
        uint32_t dw0 = 0, dw1 = 0;                            // direction words of the previous (even) column
        uint32_t* drow = dirs ? dirs + pd.dir_off + (uint64_t)(s * kWarp + lane) * pd.pitch * 2 : nullptr;

        // Boundary row of the stripe above: the 64 values a chunk needs are fetched ONE CHUNK EARLY,
        // coalesced, two per lane, and handed to lane 0 by a shuffle at each step -- no load latency
        // is ever on the critical path of a step.
        int nxt0 = 0, nxt1 = 0, cur0 = 0, cur1 = 0;
        auto wait_for = [&](uint32_t need) {   // stripe above has published at least `need` columns
            if (lane == 0) {
                uint32_t spins = 0;
                while (ld_acquire(prog_in) < need) {
                    __nanosleep(64);
                    if (++spins > (1u << 25)) { atomicExch(stall_flag, 1u); break; }   // never hang the device
                }
            }
            __syncwarp();
        };
        auto fetch = [&](uint32_t first_col) {   // columns first_col + lane and first_col + 32 + lane
            const uint32_t c0 = first_col + lane, c1 = c0 + 32;
            nxt0 = c0 <= T ? __ldcg(row_in + c0) : 0;
            nxt1 = c1 <= T ? __ldcg(row_in + c1) : 0;
        };
        const uint32_t steps = T + lanes_used - 1;
        if (s > 0) { wait_for(min(T, (uint32_t)kLongChunk)); fetch(1); }
        for (uint32_t st0 = 0; st0 < steps; st0 += kLongChunk) {
            const uint32_t st1 = min(steps, st0 + kLongChunk);
            cur0 = nxt0; cur1 = nxt1;
            if (s > 0 && st0 + kLongChunk < T) {   // lane 0 still has columns beyond this chunk
                wait_for(min(T, st0 + 2 * kLongChunk));
                fetch(st0 + kLongChunk + 1);
            }
#pragma unroll 2
            for (uint32_t st = st0; st < st1; ++st) {
                const int j = (int)st - lane + 1;
                const int above = __shfl_up_sync(kFull, Y[R - 1], 1);
                const uint32_t src = st - st0;
                const int bval = (s == 0) ? frame * (int)(st + 1) + 1
                                          : __shfl_sync(kFull, src < 32 ? cur0 : cur1, (int)(src & 31u));
                const int from_above = lane == 0 ? bval : above;
                const bool active = lane_on && j >= 1 && j <= (int)T;
                if (active) {
                    if (((j - 1) & 15) == 0) { tw = tw_next; tw_next = tw_base[((j - 1) >> 4) + 1]; }
                    const uint32_t c = tw & 3u;
                    tw >>= 2;
                    const uint32_t tab = K.tab_mis ^ (K.tab_diff << (8 * c));
                    int up = from_above, dg = up_prev;
                    up_prev = from_above;
                    uint32_t accZ = 0, accY = 0, w0 = 0, w1 = 0;
                    const int clampv = 3 - 4 * K.gap * j;   // local: H = 0 with the stop tag, in this column's frame
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int S = (int)prmt(tab, 0u, sel[r]);
                        const int m1 = __viaddmax_s32(dg, S, Y[r]);
                        int Z = __viaddmax_s32(up, K.cu, m1);
                        if (TYPE == 1) Z = max(Z, clampv);   // clamp at 0 (team_alignment.cpp:185), tag 3 = stop
                        dg = Y[r];
                        Y[r] = (int)lop3_and_or((uint32_t)Z, MASK, ONE);
                        up = Y[r];
                        accZ = accZ * FOUR + (uint32_t)Z;
                        accY = accY * FOUR + (uint32_t)Y[r];
                        if (r == 15) { w0 = accZ - accY + 0x55555555u; accZ = 0; accY = 0; }
                        if (r == 31) { w1 = accZ - accY + 0x55555555u; }
                    }
                    if (lane == kWarp - 1 && !last_stripe) __stcg(row_out + j, Y[R - 1]);
                    if (TYPE == 1) {   // running maximum (value only; the cell is located by a second pass)
                        int cm;
                        if (rows_valid == (uint32_t)R) {
                            int t8[8];
#pragma unroll
                            for (int q4 = 0; q4 < 8; ++q4) t8[q4] = max(max(Y[4 * q4], Y[4 * q4 + 1]), max(Y[4 * q4 + 2], Y[4 * q4 + 3]));
                            cm = max(max(max(t8[0], t8[1]), max(t8[2], t8[3])), max(max(t8[4], t8[5]), max(t8[6], t8[7])));
                        } else {
                            cm = INT_MIN;
#pragma unroll
                            for (int r = 0; r < R; ++r) if ((uint32_t)r < rows_valid) cm = max(cm, Y[r]);
                        }
                        lbest = max(lbest, ((cm - 1) >> 2) + K.gap * j);
                    }
                    if (TYPE == 2 && last_stripe && lane == (int)lq) {
                        int yq = Y[0];
#pragma unroll
                        for (int r = 1; r < R; ++r) if ((uint32_t)r == rq) yq = Y[r];
                        const int hq = ((yq - 1) >> 2) + K.gap * j;   // back to H units
                        if (hq > rowbest) { rowbest = hq; rowj = (uint32_t)j; }
                    }
                    if (drow) {
                        const uint32_t cidx = (uint32_t)(j - 1);
                        if (cidx & 1u) {
                            __stcs(reinterpret_cast<uint4*>(drow + (uint64_t)(cidx - 1) * 2), make_uint4(dw0, dw1, w0, w1));
                        } else if (cidx + 1 == T) {
                            __stcs(reinterpret_cast<uint2*>(drow + (uint64_t)cidx * 2), make_uint2(w0, w1));
                        } else { dw0 = w0; dw1 = w1; }
                    }
                }
            }
            if (!last_stripe && lane == kWarp - 1) {
                // columns 1 .. st1-31 of the bottom row are written (all by this lane): publish them
                const int done = (int)st1 - (kWarp - 1);
                st_release(prog_out, (uint32_t)max(0, min(done, (int)T)));
            }
        }
        // column T of this lane's rows is in Y[] (frame T): back to H units for the end-cell rules
        int final_h = 0;
        if (TYPE == 0 && last_stripe && lane == (int)lq) {
            int yq = Y[0];
#pragma unroll
            for (int r = 1; r < R; ++r) if ((uint32_t)r == rq) yq = Y[r];
            final_h = ((yq - 1) >> 2) + K.gap * (int)T;
        }
        if (TYPE == 2) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int h = ((Y[r] - 1) >> 2) + K.gap * (int)T;
                if ((uint32_t)r < rows_valid && h > colbest) { colbest = h; coli = i0 + 1 + r; }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int ob = __shfl_xor_sync(kFull, colbest, o);
                const uint32_t oi = __shfl_xor_sync(kFull, coli, o);
                if (ob > colbest || (ob == colbest && oi < coli)) { colbest = ob; coli = oi; }
            }
        }
        if (TYPE == 1) {
#pragma unroll
            for (int o = 16; o; o >>= 1) lbest = max(lbest, __shfl_xor_sync(kFull, lbest, o));
            colbest = lbest;   // stripe maximum travels in the colbest slot
        }
        final_h = __shfl_sync(kFull, final_h, (int)lq);
        rowbest = __shfl_sync(kFull, rowbest, (int)lq);
        rowj = __shfl_sync(kFull, rowj, (int)lq);
        if (lane == 0) results[task] = StripeResult{colbest, coli, rowbest, rowj, final_h, 0u};
    }
}

// Combine the per-stripe candidates of one pair (wave position k, pair p) with the reference's tie rules
// (team_alignment.cpp:117-118, :265-278). `results` is read through L2 (__ldcg): the caller may be the last
// stripe's warp of a still running fill kernel, and the entries were written by other SMs.
template <int TYPE>
__device__ __forceinline__ void finalize_pair(uint32_t p, uint32_t Q, uint32_t T, uint32_t t0, uint32_t t1,
                                              const StripeResult* results, int32_t* score, uint32_t* end_i, uint32_t* end_j) {
    const int2* r2 = reinterpret_cast<const int2*>(results);   // an entry = three int2: {colbest, coli}, {rowbest, rowj}, {final_h, pad}
    if (TYPE == 1) {
        int M = INT_MIN; uint32_t sfirst = 0;
        for (uint32_t t = t0; t < t1; ++t) { const int cb = __ldcg(r2 + 3 * t).x; if (cb > M) { M = cb; sfirst = t - t0; } }
        score[p] = M;
        // M == 0: the reference's strict '>' scan keeps the very first cell (team_alignment.cpp:186-192)
        if (M <= 0) { end_i[p] = 1; end_j[p] = 1; }
        else if ((uint32_t)__ldcg(r2 + 3 * (t0 + sfirst) + 2).y & kStripeLocated) {   // the first stripe with the maximum knows its first cell
            end_i[p] = (uint32_t)__ldcg(r2 + 3 * (t0 + sfirst)).y;
            end_j[p] = (uint32_t)__ldcg(r2 + 3 * (t0 + sfirst) + 1).y;
        }
        else { end_i[p] = 0x80000000u | sfirst; end_j[p] = 0; }   // to be resolved by the locate kernel
        return;
    }
    if (TYPE == 0) { score[p] = __ldcg(r2 + 3 * (t1 - 1) + 2).x; end_i[p] = Q; end_j[p] = T; return; }
    int colbest = INT_MIN; uint32_t coli = 0;
    for (uint32_t t = t0; t < t1; ++t) {   // stripes in row order, strict '>' keeps the smallest i
        const int2 c = __ldcg(r2 + 3 * t);
        if (c.x > colbest) { colbest = c.x; coli = (uint32_t)c.y; }
    }
    const int2 last = __ldcg(r2 + 3 * (t1 - 1) + 1);   // {rowbest, rowj} of the last stripe
    if (last.x > colbest) { score[p] = last.x; end_i[p] = Q; end_j[p] = (uint32_t)last.y; }
    else { score[p] = colbest; end_i[p] = coli; end_j[p] = T; }
}

// One thread per pair of the long class: finalize_pair, plus the pairs without inner cells.
template <int TYPE>
__global__ void finalize_long_kernel(const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ work,
                                     uint32_t n_work, const uint32_t* __restrict__ task_off,
                                     const uint8_t* __restrict__ flags, const StripeResult* results,
                                     int init, int32_t* __restrict__ score, uint32_t* __restrict__ end_i,
                                     uint32_t* __restrict__ end_j, int only_empty, uint32_t* __restrict__ ready) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_work) return;
    const uint32_t p = work[k];
    if (flags[p]) return;
    const uint32_t Q = pairs[p].Q, T = pairs[p].T;
    if (Q == 0 || T == 0) {
        if (TYPE == 0) { score[p] = (int)((Q + T) * (uint32_t)init); end_i[p] = Q; end_j[p] = T; }
        else if (TYPE == 1) { score[p] = 0; end_i[p] = 0; end_j[p] = 0; }
        else { score[p] = 0; end_i[p] = 0; end_j[p] = T; }
        if (ready) ready[k] = 1u;   // nothing to fill: a concurrent walker may take the pair at once
        return;
    }
    if (only_empty) return;
    finalize_pair<TYPE>(p, Q, T, task_off[k], task_off[k + 1], results, score, end_i, end_j);
}

// Local alignments, second pass: fill_long_kernel<1> only tracks the VALUE of the maximum; this
// kernel re-sweeps the first stripe that attains it (one warp per pair, no direction output, the
// stripe's top boundary row is still in the boundary buffer) and finds the first cell in row-major
// order whose score equals the maximum -- exactly the cell the reference's running strict '>'
// comparison keeps (team_alignment.cpp:186-192).
__global__ void __launch_bounds__(128)
locate_long_kernel(const uint32_t* __restrict__ qpk, const uint32_t* __restrict__ tpk,
                   const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ work, uint32_t n_work,
                   const uint64_t* __restrict__ bnd_off, const uint8_t* __restrict__ flags, LongConsts K,
                   const int32_t* __restrict__ bnd, const int32_t* __restrict__ score,
                   uint32_t* __restrict__ end_i, uint32_t* __restrict__ end_j) {
    constexpr int R = kLongRows;
    constexpr int STRIPE = R * kWarp;
    const int lane = threadIdx.x & 31;
    const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n_work) return;
    const uint32_t p = work[k];
    if (flags[p]) return;
    const uint32_t marker = end_i[p];
    if (!(marker & 0x80000000u)) return;
    const uint32_t s = marker & 0x7fffffffu;
    const int M = score[p];
    const PairDesc pd = pairs[p];
    const uint32_t Q = pd.Q, T = pd.T;
    const uint32_t* qw = qpk + pd.qpk_off;
    const uint32_t* tw_base = tpk + pd.tpk_off;
    const int32_t* row_in = bnd + bnd_off[k] + (uint64_t)s * (T + 4) - (T + 4);
    const uint32_t MASK = K.mask, ONE = K.one;
    const int frame = -4 * K.gap;   // local: init = 0

    const uint32_t rows_here = min((uint32_t)STRIPE, Q - s * STRIPE);
    const uint32_t lanes_used = div_up(rows_here, R);
    const uint32_t i0 = s * STRIPE + lane * R;
    const bool lane_on = (uint32_t)lane < lanes_used;
    const uint32_t rows_valid = lane_on ? min((uint32_t)R, Q - i0) : 0;
    uint32_t sel[R];
    int Y[R];
    {
        const uint32_t q0 = lane_on ? qw[i0 >> 4] : 0u, q1 = lane_on ? qw[(i0 >> 4) + 1] : 0u;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t c = ((r < 16 ? q0 : q1) >> (2 * (r & 15))) & 3u;
            sel[r] = c * 0x1111u + 0x8880u;
            Y[r] = 1;
        }
    }
    int up_prev = 1;
    uint32_t tw = 0, tw_next = lane_on ? tw_base[0] : 0u;
    uint32_t bi = 0xffffffffu, bj = 0;
    int nxt0 = 0, nxt1 = 0, cur0 = 0, cur1 = 0;
    auto fetch = [&](uint32_t first_col) {
        const uint32_t c0 = first_col + lane, c1 = c0 + 32;
        nxt0 = c0 <= T ? __ldcg(row_in + c0) : 0;
        nxt1 = c1 <= T ? __ldcg(row_in + c1) : 0;
    };
    const uint32_t steps = T + lanes_used - 1;
    if (s > 0) fetch(1);
    for (uint32_t st0 = 0; st0 < steps; st0 += kLongChunk) {
        const uint32_t st1 = min(steps, st0 + kLongChunk);
        cur0 = nxt0; cur1 = nxt1;
        if (s > 0 && st0 + kLongChunk < T) fetch(st0 + kLongChunk + 1);
        for (uint32_t st = st0; st < st1; ++st) {
            const int j = (int)st - lane + 1;
            const int above = __shfl_up_sync(kFull, Y[R - 1], 1);
            const uint32_t src = st - st0;
            const int bval = (s == 0) ? frame * (int)(st + 1) + 1
                                      : __shfl_sync(kFull, src < 32 ? cur0 : cur1, (int)(src & 31u));
            const int from_above = lane == 0 ? bval : above;
            const bool active = lane_on && j >= 1 && j <= (int)T;
            if (active) {
                if (((j - 1) & 15) == 0) { tw = tw_next; tw_next = tw_base[((j - 1) >> 4) + 1]; }
                const uint32_t c = tw & 3u;
                tw >>= 2;
                const uint32_t tab = K.tab_mis ^ (K.tab_diff << (8 * c));
                int up = from_above, dg = up_prev;
                up_prev = from_above;
                const int clampv = 3 - 4 * K.gap * j;
                const int targetY = 4 * M - 4 * K.gap * j + 1;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int S = (int)prmt(tab, 0u, sel[r]);
                    const int m1 = __viaddmax_s32(dg, S, Y[r]);
                    int Z = __viaddmax_s32(up, K.cu, m1);
                    Z = max(Z, clampv);
                    dg = Y[r];
                    Y[r] = (int)lop3_and_or((uint32_t)Z, MASK, ONE);
                    up = Y[r];
                    const uint32_t i = i0 + 1 + r;
                    if (Y[r] == targetY && (uint32_t)r < rows_valid && i < bi) { bi = i; bj = (uint32_t)j; }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const uint32_t oi = __shfl_xor_sync(kFull, bi, o), oj = __shfl_xor_sync(kFull, bj, o);
        if (oi < bi) { bi = oi; bj = oj; }
    }
    if (lane == 0) { end_i[p] = bi; end_j[p] = bj; }
}

}  // namespace b200
