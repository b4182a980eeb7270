// align_fill_long16.cuh -- K3: the long-pair DP fill on packed int16x2 (BASELINE configs 4/5).
//
// Same stripe dataflow as align_fill_long.cuh (one WARP per stripe, stripes of one pair pipelined
// across warps/SMs through boundary rows + acquire/release progress counters), but every register
// carries TWO cells, as in align_fill_short.cuh: a lane owns 64 consecutive query rows, rows
// [64L+1, 64L+32] in the low halves of its 32 registers and rows [64L+33, 64L+64] in the high halves,
// and the high block trails the low block by exactly one column. At warp step st
//     low  block of lane L works on column jlo = st - 2L + 1
//     high block of lane L works on column jhi = st - 2L          (= jlo - 1)
// so the row above the low block is the high block of lane L-1 as it stood after step st-1 (one
// __shfl_up_sync, as before) and the row above the high block is the lane's own low row 31 before
// this step's update (no shuffle at all). A stripe is 32 lanes x 64 rows = 2048 rows.
//
// Arithmetic per register (two cells): PRMT (both substitution terms from two per-column byte tables),
// VIADDMNMX.S16x2 x2, LOP3, IMAD x2 -- the instruction count align_fill_long.cuh spends on ONE cell.
// It is the same tagged moving frame (bit-exact restatement of team_alignment.cpp:104-114, ties
// diagonal > left > up), Y = 4*H - 4*gap*j + 1, except that 16 bits cannot hold Y of an 8 kb read.
// Each (lane, half) block therefore keeps a private int32 base B (a multiple of 4, so the tag bits
// survive) with register value = Y - B:
//   * neighbouring cells differ by a bounded amount (gap <= dH <= max(gap, s_max - gap, 0) per step,
//     see long16_scores_ok in align_plan.cu), so inside a 32-row block and over one 64-column chunk the
//     values stay within a few thousand of each other;
//   * at every chunk start a block re-centres itself on its row 0 (32 VIADD.16x2 per 64 columns);
//   * values that cross blocks (the row above) are shifted by the difference of the two bases, in
//     modular 16-bit arithmetic, which is exact whenever the true result fits -- and it does, because
//     the two rows are adjacent;
//   * the boundary rows between stripes and every score that leaves the kernel are true int32.
// Nothing is approximated: as long as no half leaves int16 (guaranteed by the eligibility bound) the
// results are those of the int32 kernel, bit for bit.
//
// Direction layout (klass kClassLong16): one uint4 per (lane block of 64 rows, slot), slot = step - 2L:
//   word k covers rows 8k..8k+7 of both blocks; low 16 bits = low block at column slot+1, high 16 bits =
//   high block at column slot; row 8k in the top two bits of the half; stored value is the TAG
//   (2 diagonal, 1 left, 0 up, 3 local stop). word = dirs[dir_off + ((i-1 >> 6) * pitch + slot) * 4 + k].
#pragma once
#include "align_fill_long.cuh"
#include "align_fill_short.cuh"
#include "common.cuh"

namespace b200 {


// Y[r] for a warp-uniform r without 31 selects: a jump on r.
#define B200_PICK4(k) case k: v = Y[k]; break; case k + 1: v = Y[k + 1]; break; case k + 2: v = Y[k + 2]; break; case k + 3: v = Y[k + 3]; break;
__device__ __forceinline__ uint32_t pick_uniform(const uint32_t (&Y)[32], uint32_t r) {
    uint32_t v = 0;
    switch (r) {
        B200_PICK4(0) B200_PICK4(4) B200_PICK4(8) B200_PICK4(12) B200_PICK4(16) B200_PICK4(20) B200_PICK4(24) B200_PICK4(28)
        default: break;
    }
    return v;
}
#undef B200_PICK4

// The same search restricted to the first vlo / vhi rows of the two blocks (the lanes that hold the end of the query):
// {max of the low halves, max of the high halves, first row of each}. A real call on a copy of the registers: rare,
// and inlined it costs the fill kernel 16 registers -- the room a traceback CTA needs next to it.
__device__ __noinline__ uint4 masked_max_rows(const uint32_t* Y, uint32_t vlo, uint32_t vhi) {
    int mlo = INT_MIN, mhi = INT_MIN;
    for (uint32_t r = 0; r < 32; ++r) {
        if (r < vlo) mlo = max(mlo, half_lo(Y[r]));
        if (r < vhi) mhi = max(mhi, half_hi(Y[r]));
    }
    uint32_t rlo = 32, rhi = 32;
    for (uint32_t r = 32; r-- > 0;) {
        if (r < vlo && half_lo(Y[r]) == mlo) rlo = r;
        if (r < vhi && half_hi(Y[r]) == mhi) rhi = r;
    }
    return make_uint4((uint32_t)mlo, (uint32_t)mhi, rlo, rhi);
}

// One stripe sweep: directions, boundary row, progress, end-cell candidates. Local alignments keep, per
// 32-row block, the first cell in row-major order that attains the block's maximum (team_alignment.cpp:186-192:
// the reference's running strict '>' keeps exactly that cell); the rows of a column are only searched when the
// column maximum reaches the warp-wide running maximum `thr` (refreshed every chunk) -- a cell below it cannot
// be the stripe's maximum -- so off-diagonal blocks pay nothing and no second pass is needed.
template <int TYPE, int SUB>   // SUB as in align_fill_short.cuh: 1 = substitution term from the shared-memory table
struct Sweep16 {
    uint32_t tab_at;     // SUB = 1: shared address of the lane's column of the substitution table
    // inputs
    const uint32_t* qw; const uint32_t* tw_base;
    uint32_t Q, T, s, lanes_used;
    bool last_stripe;
    const int32_t* row_in; int32_t* row_out;
    const uint32_t* prog_in; uint32_t* prog_out; uint32_t* stall_flag;
    const uint32_t* lb_in; uint32_t* lb_out;   // local: running maximum handed down the stripes of a pair (biased by 2^31, 0 = none yet)
    uint32_t* drow;      // this lane's direction row (or nullptr)
    // outputs
    int colbest; uint32_t coli; int rowbest; uint32_t rowj; int final_h;
    int lbest_lo, lbest_hi;                       // local: maximum of the lane's low / high block ...
    uint32_t bi_lo, bj_lo, bi_hi, bj_hi;          // ... and the first cell (row-major) that attains it

    __device__ __forceinline__ void run(const ShortConsts& K, int lane) {
        constexpr int R = 32;
        const uint32_t MASK = K.mask, ONE = K.one, FOUR = K.four;
        const int gap = K.gap, init = K.init;
        const int frame = 4 * (init - gap);   // border row 0 in the moving frame: Y(0,j) = frame*j + 1
        const uint32_t lq = ((Q - 1) >> 6) & 31u, hq = ((Q - 1) >> 5) & 1u, rq = (Q - 1) & 31u;   // lane / half / register of row Q
        const uint32_t i0 = s * kL16Stripe + (uint32_t)lane * kL16LaneRows;   // low rows i0+1..i0+32, high rows i0+33..i0+64
        const bool lane_on = (uint32_t)lane < lanes_used;
        const uint32_t vlo = lane_on ? min(32u, Q - i0) : 0u;
        const uint32_t vhi = (lane_on && Q - i0 > 32u) ? min(32u, Q - i0 - 32u) : 0u;
        const bool full = (vlo == 32u) && (vhi == 32u);

        colbest = INT_MIN; coli = 0; rowbest = INT_MIN; rowj = 0; final_h = 0;
        lbest_lo = INT_MIN; lbest_hi = INT_MIN; bi_lo = 0xffffffffu; bj_lo = 0; bi_hi = 0xffffffffu; bj_hi = 0;
        int thr = INT_MIN;   // local: lower bound of the stripe's maximum
        if (TYPE == 2) {
            if (s == 0 && lane == 0) { colbest = 0; coli = 0; }               // H(0,T) = 0 comes first
            if (last_stripe && lane == (int)lq) { rowbest = 0; rowj = 0; }    // H(Q,0) = 0
        }

        uint32_t sel[R], Y[R];
        int Blo = 4 * (int)(i0 * (uint32_t)init), Bhi = 4 * (int)((i0 + 32u) * (uint32_t)init);
        {
            const uint32_t w = i0 >> 4;
            const uint32_t a0 = lane_on ? qw[w] : 0u, a1 = lane_on ? qw[w + 1] : 0u;
            const uint32_t b0 = lane_on ? qw[w + 2] : 0u, b1 = lane_on ? qw[w + 3] : 0u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t ca = ((r < 16 ? a0 : a1) >> (2 * (r & 15))) & 3u;
                const uint32_t cb = ((r < 16 ? b0 : b1) >> (2 * (r & 15))) & 3u;
                sel[r] = SUB ? tab_at + subst_row_part(ca, cb)
                             : ca | ((8u + ca) << 4) | ((4u + cb) << 8) | ((12u + cb) << 12);
                Y[r] = dup16(4 * (r + 1) * init + 1);   // column 0, relative to the block's base
            }
        }
        uint32_t up_prev = dup16(1);                      // Y(i0, 0) - Blo (the high half is set by the first step)
        uint32_t tw = 0, tw_next = lane_on ? tw_base[0] : 0u;
        uint32_t tabB = 0;                                // the previous step's low table = this step's high table
        uint32_t c_prev = 0;                              // SUB = 1: the previous step's column code (the high block's column)

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
        const uint32_t steps = T + 2 * lanes_used;   // lane L runs steps 2L .. T + 2L
        if (s > 0) { wait_for(min(T, (uint32_t)kL16Chunk)); fetch(1); }
        for (uint32_t st0 = 0; st0 < steps; st0 += kL16Chunk) {
            const uint32_t st1 = min(steps, st0 + kL16Chunk);
            cur0 = nxt0; cur1 = nxt1;
            if (s > 0 && st0 + kL16Chunk < T) {   // lane 0 still has columns beyond this chunk
                wait_for(min(T, st0 + 2 * kL16Chunk));
                fetch(st0 + kL16Chunk + 1);
            }
            // re-centre a running block pair on its row 0
            if (lane_on && st0 > 2u * lane && st0 <= T + 2u * lane) {
                const int dlo = half_lo(Y[0]) & ~3, dhi = half_hi(Y[0]) & ~3;
                const uint32_t nd = pack16(-dlo, -dhi);
#pragma unroll
                for (int r = 0; r < R; ++r) Y[r] = __vadd2(Y[r], nd);
                up_prev = __vadd2(up_prev, nd);
                Blo += dlo; Bhi += dhi;
            }
            if (TYPE == 1) {
                // lower bound of the pair's maximum: this stripe's own running maximum, and what the stripes above have
                // seen (they are at least a chunk ahead and closer to the alignment that is still on its way down here --
                // without it the rising scores below the diagonal would pass for maxima on every step). Any stale value is
                // a valid bound.
                thr = __reduce_max_sync(kFull, max(lbest_lo, lbest_hi));
                if (s > 0) thr = max(thr, (int)(__ldcg(lb_in) ^ 0x80000000u));
            }
            const int Bab = __shfl_up_sync(kFull, Bhi, 1);
            const uint32_t conv = pack16(Bab - Blo, Blo - Bhi);   // base shifts for the two rows above
#pragma unroll 1
            for (uint32_t st = st0; st < st1; ++st) {
                const int jlo = (int)st - 2 * lane + 1;
                const uint32_t above = __shfl_up_sync(kFull, Y[R - 1], 1);
                const uint32_t src = st - st0;
                const int bval = (s == 0) ? frame * (int)(st + 1) + 1
                                          : __shfl_sync(kFull, src < 32 ? cur0 : cur1, (int)(src & 31u));
                // row above: low half <- high row 31 of the lane above, high half <- own low row 31 (both pre-update)
                uint32_t fa = __vadd2(prmt(above, Y[R - 1], 0x5432u), conv);
                if (lane == 0) fa = (fa & 0xffff0000u) | ((uint32_t)(bval - Blo) & 0xffffu);
                const bool active = lane_on && jlo >= 1 && jlo <= (int)T + 1;
                if (active) {
                    const bool alo = jlo <= (int)T, ahi = jlo >= 2;   // which blocks work on a real column
                    if (!alo) {
                        // the low block is done (this step only the high block has a column): take what the
                        // end-cell rules need from column T before the registers are reused
                        if (TYPE == 2) {
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const int h = ((half_lo(Y[r]) + Blo - 1) >> 2) + gap * (int)T;
                                if ((uint32_t)r < vlo && h > colbest) { colbest = h; coli = i0 + 1 + r; }
                            }
                        }
                        if (TYPE == 0 && last_stripe && lane == (int)lq && hq == 0u)
                            final_h = ((half_lo(pick_any(Y, rq)) + Blo - 1) >> 2) + gap * (int)T;
                    }
                    if (((jlo - 1) & 15) == 0) { tw = tw_next; tw_next = tw_base[((jlo - 1) >> 4) + 1]; }
                    const uint32_t c = tw & 3u;
                    tw >>= 2;
                    const uint32_t tabA = SUB ? subst_col_part(c, c_prev) : K.tab_mis ^ (K.tab_diff << (8 * c));
                    if (SUB) tabB = K.one32;
                    uint32_t up = fa, dg = up_prev;
                    up_prev = fa;
                    uint32_t clampv = 0;
                    if (TYPE == 1) {   // H = 0 with the stop tag, in each block's frame; saturated, it can never win
                        const int cl = max(3 - 4 * gap * jlo - Blo, -32768);
                        const int ch = max(3 - 4 * gap * (jlo - 1) - Bhi, -32768);
                        clampv = pack16(cl, ch);
                    }
                    uint32_t accZ = 0, accY = 0, w[4];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const uint32_t S = SUB ? subst_lookup(sel[r], tabA, tabB) : prmt(tabA, tabB, sel[r]);
                        const uint32_t m1 = __viaddmax_s16x2(dg, S, Y[r]);
                        uint32_t Z = __viaddmax_s16x2(up, K.cu, m1);
                        if (TYPE == 1) Z = __vmaxs2(Z, clampv);   // clamp at 0 (team_alignment.cpp:185), tag 3 = stop
                        dg = Y[r];
                        Y[r] = lop3_and_or(Z, MASK, ONE);
                        up = Y[r];
                        accZ = accZ * FOUR + Z;
                        accY = accY * FOUR + Y[r];
                        if ((r & 7) == 7) { w[r >> 3] = accZ - accY + 0x55555555u; accZ = 0; accY = 0; }
                    }
                    if (SUB) c_prev = c; else tabB = tabA;
                    if (lane == kWarp - 1 && !last_stripe && ahi) __stcg(row_out + (jlo - 1), half_hi(Y[R - 1]) + Bhi);
                    if (drow) __stcs(reinterpret_cast<uint4*>(drow + (uint64_t)(jlo - 1) * 4), make_uint4(w[0], w[1], w[2], w[3]));
                    if (TYPE == 1) {
                        const uint32_t cm = max_tree16(Y);
                        int mlo = half_lo(cm), mhi = half_hi(cm);   // column maxima as register halves (rows past Q included)
                        int hlo = (alo && vlo) ? ((mlo + Blo - 1) >> 2) + gap * jlo : INT_MIN;
                        int hhi = (ahi && vhi) ? ((mhi + Bhi - 1) >> 2) + gap * (jlo - 1) : INT_MIN;
                        // a new block maximum, or a tie that may sit on a smaller row than the current holder -- and not below `thr`
                        if ((hlo >= thr && (hlo > lbest_lo || (hlo == lbest_lo && bi_lo > i0 + 1))) ||
                            (hhi >= thr && (hhi > lbest_hi || (hhi == lbest_hi && bi_hi > i0 + 33)))) {
                            uint32_t rlo = R, rhi = R;   // smallest row of each block that holds the block's column maximum
                            if (full && (cm & 0xffffu) != 0x8000u && (cm >> 16) != 0x8000u) {   // (-32768 has no packed negative)
                                const uint32_t km = first_rows_of_max(Y, cm);
                                rlo = 31u - (uint32_t)half_lo(km); rhi = 31u - (uint32_t)half_hi(km);
                            } else {   // rows past Q may be in the tree: redo with masks (the lanes that hold the end of the query)
                                uint32_t ycopy[R];
#pragma unroll
                                for (int r = 0; r < R; ++r) ycopy[r] = Y[r];
                                const uint4 mr = masked_max_rows(ycopy, vlo, vhi);
                                mlo = (int)mr.x; mhi = (int)mr.y; rlo = mr.z; rhi = mr.w;
                                hlo = (alo && vlo) ? ((mlo + Blo - 1) >> 2) + gap * jlo : INT_MIN;
                                hhi = (ahi && vhi) ? ((mhi + Bhi - 1) >> 2) + gap * (jlo - 1) : INT_MIN;
                            }
                            if (hlo != INT_MIN && (hlo > lbest_lo || (hlo == lbest_lo && i0 + 1 + rlo < bi_lo))) {
                                lbest_lo = hlo; bi_lo = i0 + 1 + rlo; bj_lo = (uint32_t)jlo;
                            }
                            if (hhi != INT_MIN && (hhi > lbest_hi || (hhi == lbest_hi && i0 + 33 + rhi < bi_hi))) {
                                lbest_hi = hhi; bi_hi = i0 + 33 + rhi; bj_hi = (uint32_t)(jlo - 1);
                            }
                        }
                    }
                    if (TYPE == 2 && last_stripe) {   // row Q, every column (smallest j wins ties)
                        const uint32_t yq = pick_uniform(Y, rq);
                        if (lane == (int)lq) {
                            const int jq = hq ? jlo - 1 : jlo;
                            if (jq >= 1 && jq <= (int)T) {
                                const int h = (((hq ? half_hi(yq) + Bhi : half_lo(yq) + Blo) - 1) >> 2) + gap * jq;
                                if (h > rowbest) { rowbest = h; rowj = (uint32_t)jq; }
                            }
                        }
                    }
                    if (!ahi) {
                        // first step of the lane: the high block had no column yet, put its column-0 values back
#pragma unroll
                        for (int r = 0; r < R; ++r) Y[r] = (Y[r] & 0xffffu) | ((uint32_t)(4 * (r + 1) * init + 1) << 16);
                    }
                }
            }
            if (!last_stripe && lane == kWarp - 1) {
                // the high block of lane 31 has finished columns 1 .. st1-63 of the stripe's bottom row: publish them
                const int done = (int)st1 - (2 * (kWarp - 1) + 1);
                if (TYPE == 1) __stcg(lb_out, (uint32_t)thr ^ 0x80000000u);
                st_release(prog_out, (uint32_t)max(0, min(done, (int)T)));
            }
        }
        // the high blocks now hold column T (frame T)
        if (TYPE == 0 && last_stripe && lane == (int)lq && hq == 1u)
            final_h = ((half_hi(pick_any(Y, rq)) + Bhi - 1) >> 2) + gap * (int)T;
        if (TYPE == 2) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int h = ((half_hi(Y[r]) + Bhi - 1) >> 2) + gap * (int)T;
                if ((uint32_t)r < vhi && h > colbest) { colbest = h; coli = i0 + 33 + r; }
            }
        }
    }
};

// 152 registers: three fill CTAs and one traceback CTA (walk_tile_wait_kernel, 128 threads x 40) fit an SM together.
template <int TYPE, int SUB>
__global__ void __maxnreg__(152)
fill_long16_kernel(const uint32_t* __restrict__ qpk, const uint32_t* __restrict__ tpk,
                   const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ work, uint32_t n_work,
                   const uint32_t* __restrict__ task_off, const uint64_t* __restrict__ bnd_off,
                   uint32_t* __restrict__ work_counter, const uint8_t* __restrict__ flags, ShortConsts K,
                   uint32_t* __restrict__ dirs, int32_t* bnd, uint32_t* progress, uint32_t lb_offset,
                   StripeResult* results, uint32_t* __restrict__ stall_flag,
                   // concurrent walk (all null otherwise): the warp that finishes a pair's last outstanding stripe
                   // finalises the pair and raises its ready flag, so a walker can start while the fill goes on
                   uint32_t* pair_done, uint32_t* ready, int32_t* score, uint32_t* end_i, uint32_t* end_j) {
    const int lane = threadIdx.x & 31;
    const uint32_t n_tasks = task_off[n_work];
    uint32_t tab_at = 0;
    if (SUB) {
        __shared__ uint32_t subst_tab[kSubstWords];
        subst_table_fill(subst_tab, K, threadIdx.x, blockDim.x);
        __syncthreads();
        tab_at = (uint32_t)__cvta_generic_to_shared(subst_tab) + (uint32_t)lane * 4u;
    }
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
        Sweep16<TYPE, SUB> sw;
        sw.tab_at = tab_at;
        sw.qw = qpk + pd.qpk_off; sw.tw_base = tpk + pd.tpk_off;
        sw.Q = pd.Q; sw.T = pd.T; sw.s = s;
        const uint32_t n_stripes = div_up(pd.Q, kL16Stripe);
        sw.lanes_used = div_up(min((uint32_t)kL16Stripe, pd.Q - s * kL16Stripe), kL16LaneRows);
        sw.last_stripe = (s + 1 == n_stripes);
        const uint32_t row_pitch = pd.T + 4;
        sw.row_out = bnd + bnd_off[k] + (uint64_t)s * row_pitch;
        sw.row_in = sw.row_out - row_pitch;             // written by stripe s-1 (valid when s > 0)
        sw.prog_out = progress + task; sw.prog_in = progress + task - 1; sw.stall_flag = stall_flag;
        sw.lb_out = progress + lb_offset + task; sw.lb_in = sw.lb_out - 1;   // second half of the progress buffer
        sw.drow = dirs ? dirs + pd.dir_off + (uint64_t)(s * kWarp + lane) * pd.pitch * 4 : nullptr;
        sw.run(K, lane);

        int colbest = sw.colbest; uint32_t coli = sw.coli;
        int rowbest_l = 0; uint32_t rowj_l = 0, located = 0;
        if (TYPE == 2) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int ob = __shfl_xor_sync(kFull, colbest, o);
                const uint32_t oi = __shfl_xor_sync(kFull, coli, o);
                if (ob > colbest || (ob == colbest && oi < coli)) { colbest = ob; coli = oi; }
            }
        }
        if (TYPE == 1) {
            // stripe maximum and the first cell in row-major order that attains it: (value, row, column) travel in
            // the colbest / coli / rowj slots, kStripeLocated tells finalize_pair that no locate pass is needed
            int m = sw.lbest_lo; uint32_t bi = sw.bi_lo, bj = sw.bj_lo;
            if (sw.lbest_hi > m) { m = sw.lbest_hi; bi = sw.bi_hi; bj = sw.bj_hi; }   // a tie keeps the low block (smaller rows)
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int om = __shfl_xor_sync(kFull, m, o);
                const uint32_t oi = __shfl_xor_sync(kFull, bi, o), oj = __shfl_xor_sync(kFull, bj, o);
                if (om > m || (om == m && oi < bi)) { m = om; bi = oi; bj = oj; }
            }
            colbest = m; coli = bi; rowj_l = bj; located = kStripeLocated;
        }
        const uint32_t lq = ((pd.Q - 1) >> 6) & 31u;
        const int final_h = __shfl_sync(kFull, sw.final_h, (int)lq);
        const int rowbest = TYPE == 1 ? rowbest_l : __shfl_sync(kFull, sw.rowbest, (int)lq);
        const uint32_t rowj = TYPE == 1 ? rowj_l : __shfl_sync(kFull, sw.rowj, (int)lq);
        if (lane == 0) results[task] = StripeResult{colbest, coli, rowbest, rowj, final_h, located};
        if (ready != nullptr) {
            // Every lane has written direction words of this stripe: each lane's own stores are made visible
            // device-wide by its own fence, and the warp barrier orders all of them before lane 0's count / flag
            // (shuffles are not memory barriers; a walker acquires `ready` and then reads those words).
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                __threadfence();                                           // this stripe's result before the count
                if (atomicAdd(pair_done + k, 1u) + 1 == n_stripes) {       // every stripe of the pair has reported
                    __threadfence();
                    finalize_pair<TYPE>(p, pd.Q, pd.T, task_off[k], task_off[k] + n_stripes, results, score, end_i, end_j);
                    __threadfence();
                    st_release(ready + k, 1u);
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace b200
