// align_walk.cuh -- traceback over the 2-bit direction matrix and CIGAR emission.
//
// Reference semantics restated (team_alignment/team_alignment.cpp):
//   walk :123-138 (global), :202-217 (local: stop at the first cell whose score is 0),
//   :287-302 (semiGlobal), border parents :83-92 (column 0 -> up 'D', row 0 -> left 'I'),
//   semiGlobal tail pad :306-315, run-length encoding :145-160, empty path => "1\0".
//
// Kernels:
//   walk_kernel       one THREAD per pair (short class: the 64 pairs of a group share their
//                     direction lines, so neighbouring lanes coalesce): follows the directions
//                     from the end cell and records the runs it meets (in walk order, i.e.
//                     reversed) as packed (count << 2 | op) words, plus the byte length of the
//                     final CIGAR text;
//   walk_tile_kernel  one WARP per pair (long / generic classes): a traceback is a chain of
//                     dependent loads, ~1 us each from HBM, so a 16 k-step path costs ~15 ms
//                     however many pairs run in parallel. The warp instead fetches the whole
//                     (row block x 32 columns) tile around the current cell with one coalesced
//                     load into shared memory (128-512 contiguous bytes in the fill kernels'
//                     layouts), prefetches the three tiles the path can enter next into L2, and
//                     walks inside the tile from shared memory: one HBM round trip per ~32 steps;
//   emit_kernel       (after an exclusive scan of the lengths) prints the runs back to front
//                     into the pair's slice of the CIGAR buffer.
#pragma once
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t dec_digits(uint32_t v) {
    uint32_t d = 1;
    while (v >= 10) { v /= 10; ++d; }
    return d;
}

// `spread` (power of two, 1..32): only every spread-th lane of a warp owns a pair. Small batches of long
// pairs are latency-bound pointer chases with divergent paths; spreading them over more warps and SMs
// shortens the critical path, large batches use every lane (spread = 1).
template <int TYPE>
__global__ void __launch_bounds__(128)
walk_kernel(const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ work, uint32_t n_work, uint32_t spread,
            const uint8_t* __restrict__ skip_flags, const uint32_t* __restrict__ dirs, const uint32_t* __restrict__ end_i,
            const uint32_t* __restrict__ end_j, uint32_t* __restrict__ runs,
            uint32_t* __restrict__ n_runs, uint32_t* __restrict__ cigar_len) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid % spread) return;
    const uint32_t w = tid / spread;
    if (w >= n_work) return;
    const uint32_t p = work[w];
    if (skip_flags && skip_flags[p]) return;   // not pure ACGT: filled and walked by the repair pass (align_run.cu)
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
    const uint32_t klass = pd.klass & 0xffu;
    const uint32_t s_lane = (pd.klass >> 8) & 31u, s_shift = ((pd.klass >> 16) & 1u) * 16u;
    const uint64_t pitch = pd.pitch;
    // cache of the last direction word: an 'up' move usually stays inside the same word
    uint32_t cw = 0;
    uint64_t cw_at = ~0ull;
    while (i != 0 && j != 0) {
        uint64_t at;
        uint32_t sh;
        const uint32_t i1 = i - 1, j1 = j - 1;
        if (klass == kClassLong) {            // align_fill_long.cuh: 2 words per (32-row block, column)
            at = ((uint64_t)(i1 >> 5) * pitch + j1) * 2 + ((i1 >> 4) & 1u);
            sh = 2 * (15 - (i1 & 15u));
        } else if (klass == kClassLong16) {   // align_fill_long16.cuh: 4 words per (64-row block, slot)
            const uint32_t half = (i1 >> 5) & 1u;
            at = ((uint64_t)(i1 >> 6) * pitch + j1 + half) * 4 + ((i1 >> 3) & 3u);
            sh = 16 * half + 2 * (7 - (i1 & 7u));
        } else if (klass == kClassShort) {    // align_fill_short.cuh: uint4 per (32-row block, column, lane)
            at = (((uint64_t)(i1 >> 5) * pitch + j1) * 32 + s_lane) * 4 + ((i1 >> 3) & 3u);
            sh = s_shift + 2 * (7 - (i1 & 7u));
        } else {                              // align_fill_generic.cuh: 1 word per (16-row block, column)
            at = (uint64_t)(i1 / kRowsPerWord) * pitch + j1;
            sh = 2 * (i1 % kRowsPerWord);
        }
        if (at != cw_at) { cw = __ldg(base + at); cw_at = at; }
        uint32_t code = (cw >> sh) & 3u;
        if (klass != kClassGeneric) code = (code == 3u) ? 3u : 2u - code;   // stored as tag: 2 diag, 1 left, 0 up
        if (TYPE == 1 && code == 3) break;
        if (code == cur_op) ++cur_n;
        else {
            if (cur_n) { out[nr++] = (cur_n << 2) | cur_op; bytes += dec_digits(cur_n) + 1; }
            cur_op = code; cur_n = 1;
        }
        i -= (code != 1u);   // diagonal and up consume a query row
        j -= (code != 2u);   // diagonal and left consume a target column
    }
    if (TYPE != 1) {   // the borders: row 0 points left ('I'), column 0 points up ('D') (reference :83-92)
        if (i == 0 && j) push(1, j);
        else if (j == 0 && i) push(2, i);
    }
    if (cur_n) { out[nr++] = (cur_n << 2) | cur_op; bytes += dec_digits(cur_n) + 1; }
    if (nr == 0) bytes = 2;  // "1\0"
    n_runs[p] = nr;
    cigar_len[p] = bytes;
}

// Streaming mode of the short-pair fill (ShortStream in align_fill_short.cuh): the gate in front of a wave's traceback.
// One thread waits until every group of the wave has been counted as done by the (still running) persistent fill; the
// traceback kernel follows it in stream order. A kernel of its own, one slot wide, on purpose: traceback CTAs that
// waited themselves would fill every free slot of the machine and keep the next wave's pack kernel -- which the fill is
// waiting for -- from becoming resident.
__global__ void wave_gate_kernel(const uint32_t* counter, uint32_t target, uint32_t* stall_flag) {
    uint32_t spins = 0;
    while (ld_acquire_u32(counter) < target) {
        __nanosleep(400);
        if (++spins > (1u << 22)) { atomicExch(stall_flag, 1u); break; }   // never hang the device
    }
}

// ---- warp-per-pair tile walker ------------------------------------------------------------
struct RunWriter {   // the run list under construction (every lane tracks it, lane 0 stores)
    uint32_t* out;
    uint32_t nr, bytes, cur_op, cur_n;
    bool writer;
    __device__ __forceinline__ void flush() {
        if (cur_n) { if (writer) out[nr] = (cur_n << 2) | cur_op; ++nr; bytes += dec_digits(cur_n) + 1; }
    }
    __device__ __forceinline__ void push(uint32_t op, uint32_t cnt) {
        if (op == cur_op) { cur_n += cnt; return; }
        flush();
        cur_op = op; cur_n = cnt;
    }
};

// Tile geometry of a direction layout: 2^RL rows per row block, WPC words per column slot.
template <uint32_t KLASS> struct TileShape;
template <> struct TileShape<kClassGeneric> { static constexpr uint32_t RL = 4, WPC = 1; };
template <> struct TileShape<kClassLong>    { static constexpr uint32_t RL = 5, WPC = 2; };
template <> struct TileShape<kClassLong16>  { static constexpr uint32_t RL = 6, WPC = 4; };

constexpr uint32_t kTileSlots = 128;                    // column slots per tile
constexpr uint32_t kTileBlocks = 2;                     // row blocks per tile: the current one and the one above
constexpr int kWalkTileWords = kTileBlocks * kTileSlots * 4;   // shared-memory words per warp (largest layout)

// Walks from (i, j) until a border (or, local, a stop cell) is reached.
// The tile is anchored with the current cell in its bottom-right corner (row blocks rb-1..rb, the 128 slots
// ending at the cell's slot): a mostly diagonal path gets 64-128 steps out of one HBM round trip.
// Inside the tile the warp advances by whole RUNS: lane l looks at the cell l steps further along the
// current direction (diagonal first), a ballot counts how far the run goes, and the run is pushed at once --
// a CIGAR is run-length encoded anyway, and ONT-like paths average ~10 cells per run.
// COHERENT: tile loads go through L2 (ld.cg) instead of the read-only path -- for the walker that runs while the
// fill kernel is still writing other pairs' matrices.
template <int TYPE, uint32_t KLASS, bool COHERENT>
__device__ __forceinline__ void walk_tiles(const PairDesc& pd, const uint32_t* __restrict__ base, uint32_t& i, uint32_t& j,
                                           uint32_t* tile, RunWriter& rw, int lane) {
    constexpr uint32_t RL = TileShape<KLASS>::RL, WPC = TileShape<KLASS>::WPC;
    const uint32_t pitch = pd.pitch;
    const uint32_t n_rb = (pd.Q + (1u << RL) - 1) >> RL;
    // tag/code of cell (i1, j1) (0-based) if it lies inside the loaded tile, else 0xff
    uint32_t rb_base = 0, slot_lo = 0;
    auto peek = [&](uint32_t i1, uint32_t j1) -> uint32_t {
        const uint32_t blk = (i1 >> RL) - rb_base;
        uint32_t slot, widx, sh;
        if (KLASS == kClassGeneric) { slot = j1; widx = 0; sh = 2 * (i1 & 15u); }
        else if (KLASS == kClassLong) { slot = j1; widx = (i1 >> 4) & 1u; sh = 2 * (15 - (i1 & 15u)); }
        else { const uint32_t half = (i1 >> 5) & 1u; slot = j1 + half; widx = (i1 >> 3) & 3u; sh = 16 * half + 2 * (7 - (i1 & 7u)); }
        const uint32_t srel = slot - slot_lo;
        if (blk >= kTileBlocks || srel >= kTileSlots) return 0xffu;
        uint32_t code = (tile[(blk * kTileSlots + srel) * WPC + widx] >> sh) & 3u;
        if (KLASS != kClassGeneric) code = (code == 3u) ? 3u : 2u - code;   // stored as tag: 2 diag, 1 left, 0 up
        return code;   // 0 diagonal, 1 left, 2 up, 3 stop
    };
    bool stop = false;
    while (i != 0 && j != 0 && !stop) {
        {   // (re)load the tile around the current cell
            const uint32_t i1 = i - 1;
            const uint32_t rb = i1 >> RL;
            const uint32_t slot = j - 1 + (KLASS == kClassLong16 ? (i1 >> 5) & 1u : 0u);
            rb_base = rb > 0 ? rb - 1 : 0;
            // (one spare slot to the right: in the long16 layout an 'up' move out of a low block lands one slot further)
            slot_lo = slot + 1 >= kTileSlots - 1 ? slot + 1 - (kTileSlots - 1) : 0;
            __syncwarp();
#pragma unroll
            for (uint32_t b = 0; b < kTileBlocks; ++b) {
#pragma unroll
                for (uint32_t q = 0; q < kTileSlots / 32; ++q) {
                    const uint32_t srel = q * 32 + lane, sl = slot_lo + srel;
                    if (rb_base + b < n_rb && sl < pitch) {
                        const uint32_t* src = base + ((uint64_t)(rb_base + b) * pitch + sl) * WPC;
                        uint32_t* dst = tile + (b * kTileSlots + srel) * WPC;
                        if (COHERENT) {
                            if (WPC == 1) dst[0] = __ldcg(src);
                            else if (WPC == 2) *reinterpret_cast<uint2*>(dst) = __ldcg(reinterpret_cast<const uint2*>(src));
                            else *reinterpret_cast<uint4*>(dst) = __ldcg(reinterpret_cast<const uint4*>(src));
                        } else {
                            if (WPC == 1) dst[0] = __ldg(src);
                            else if (WPC == 2) *reinterpret_cast<uint2*>(dst) = __ldg(reinterpret_cast<const uint2*>(src));
                            else *reinterpret_cast<uint4*>(dst) = __ldg(reinterpret_cast<const uint4*>(src));
                        }
                    }
                }
            }
            __syncwarp();
        }
        for (;;) {
            if (i == 0 || j == 0) break;
            // diagonal run: lane l tests cell (i - l, j - l)
            const uint32_t l = (uint32_t)lane;
            uint32_t code = (i > l && j > l) ? peek(i - 1 - l, j - 1 - l) : 0xfeu;
            const uint32_t c0 = __shfl_sync(kFull, code, 0);
            if (c0 == 0xffu) break;                          // current cell outside the tile: reload
            if (TYPE == 1 && c0 == 3u) { stop = true; break; }
            uint32_t run;
            if (c0 == 0u) {
                run = __ffs(~__ballot_sync(kFull, code == 0u)) - 1;   // ballot has bit 0 set; all 32 set -> ffs(0) = 0 -> wraps
                if (run > 32u) run = 32u;
                i -= run; j -= run;
            } else if (c0 == 1u) {                           // left run along row i: lane l tests (i, j - l)
                code = (j > l) ? peek(i - 1, j - 1 - l) : 0xfeu;
                run = __ffs(~__ballot_sync(kFull, code == 1u)) - 1;
                if (run > 32u) run = 32u;
                j -= run;
            } else {                                         // up run along column j: lane l tests (i - l, j)
                code = (i > l) ? peek(i - 1 - l, j - 1) : 0xfeu;
                run = __ffs(~__ballot_sync(kFull, code == 2u)) - 1;
                if (run > 32u) run = 32u;
                i -= run;
            }
            rw.push(c0, run);
        }
    }
}

// The walk of one pair by one warp (tile by tile), from its end cell to a border / stop cell.
template <int TYPE, bool COHERENT>
__device__ __forceinline__ void walk_pair_tiles(const PairDesc& pd, uint32_t p, const uint32_t* __restrict__ dirs,
                                                uint32_t i, uint32_t j, uint32_t* tile, uint32_t* __restrict__ runs,
                                                uint32_t* __restrict__ n_runs, uint32_t* __restrict__ cigar_len, int lane) {
    const uint32_t klass = pd.klass & 0xffu;
    const uint32_t Q = pd.Q, T = pd.T;
    RunWriter rw{runs + pd.run_off, 0, 0, 3, 0, lane == 0};
    if (TYPE == 2) {  // the pad is the tail of the text, so it is the first thing a backward walk meets
        if (i == Q && j < T) rw.push(1, T - j);
        else if (j == T && i < Q) rw.push(2, Q - i);
    }
    const uint32_t* base = dirs + pd.dir_off;
    if (klass == kClassLong16) walk_tiles<TYPE, kClassLong16, COHERENT>(pd, base, i, j, tile, rw, lane);
    else if (klass == kClassLong) walk_tiles<TYPE, kClassLong, COHERENT>(pd, base, i, j, tile, rw, lane);
    else walk_tiles<TYPE, kClassGeneric, COHERENT>(pd, base, i, j, tile, rw, lane);
    if (TYPE != 1) {   // the borders: row 0 points left ('I'), column 0 points up ('D') (reference :83-92)
        if (i == 0 && j) rw.push(1, j);
        else if (j == 0 && i) rw.push(2, i);
    }
    rw.flush();
    if (lane == 0) {
        n_runs[p] = rw.nr;
        cigar_len[p] = rw.nr == 0 ? 2u : rw.bytes;  // empty path: "1\0"
    }
}

template <int TYPE>
__global__ void __launch_bounds__(128)
walk_tile_kernel(const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ work, uint32_t n_work,
                 const uint8_t* __restrict__ skip_flags, const uint32_t* __restrict__ dirs, const uint32_t* __restrict__ end_i,
                 const uint32_t* __restrict__ end_j, uint32_t* __restrict__ runs,
                 uint32_t* __restrict__ n_runs, uint32_t* __restrict__ cigar_len) {
    __shared__ __align__(16) uint32_t tiles[4][kWalkTileWords];
    const int lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_work) return;
    const uint32_t p = work[w];
    if (skip_flags && skip_flags[p]) return;         // the repair pass's
    const PairDesc pd = pairs[p];
    if ((pd.klass & 0xffu) == kClassShort) return;   // walk_kernel's
    walk_pair_tiles<TYPE, false>(pd, p, dirs, end_i[p], end_j[p], tiles[threadIdx.x >> 5], runs, n_runs, cigar_len, lane);
}

// Persistent walkers for the wave that is still being filled: warps take pairs in work order (largest first, the
// order the fill hands its stripes out in), wait for the pair's ready flag -- raised by the fill warp that finished
// the pair's last stripe -- and walk it while the fill goes on with the other pairs. Pairs the 2-bit fill does not
// own (flags != 0) are skipped; they are walked after their fallback fill (the repair pass in align_run.cu).
template <int TYPE>
__global__ void __launch_bounds__(128)
walk_tile_wait_kernel(const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ work, uint32_t n_work,
                      const uint8_t* __restrict__ flags, uint32_t* __restrict__ counter, const uint32_t* ready,
                      uint32_t* __restrict__ stall_flag, const uint32_t* __restrict__ dirs, const uint32_t* end_i,
                      const uint32_t* end_j, uint32_t* __restrict__ runs, uint32_t* __restrict__ n_runs,
                      uint32_t* __restrict__ cigar_len) {
    __shared__ __align__(16) uint32_t tiles[4][kWalkTileWords];
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(counter, 1u);
        w = __shfl_sync(kFull, w, 0);
        if (w >= n_work) break;
        const uint32_t p = work[w];
        if (flags[p]) continue;
        const PairDesc pd = pairs[p];
        uint32_t ok = 1;
        if (lane == 0) {
            uint32_t spins = 0;
            while (ld_acquire_u32(ready + w) == 0u) {
                __nanosleep(256);
                if (++spins > (1u << 24)) { atomicExch(stall_flag, 1u); ok = 0; break; }   // never hang the device
            }
        }
        ok = __shfl_sync(kFull, ok, 0);
        if (!ok) break;
        walk_pair_tiles<TYPE, true>(pd, p, dirs, __ldcg(end_i + p), __ldcg(end_j + p), tiles[threadIdx.x >> 5], runs, n_runs,
                                    cigar_len, lane);
    }
}

// Score-only batches still owe the caller target_begin (reference :119-121, :197-199, :283-285).
__global__ void target_begin_kernel(uint32_t first, uint32_t n, int type, const uint32_t* __restrict__ end_j,
                                    uint32_t* __restrict__ target_begin) {
    const uint32_t p = first + blockIdx.x * blockDim.x + threadIdx.x;   // pairs [first, n)
    if (p < n) target_begin[p] = (type == 1) ? end_j[p] + 1 : 0;
}

__global__ void add_offset_kernel(uint64_t* __restrict__ v, uint32_t n, const uint64_t* __restrict__ base) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] += *base;
}

// One thread per pair (short pairs: a few runs each).
__global__ void __launch_bounds__(128)
emit_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t n, const uint32_t* __restrict__ runs,
            const uint32_t* __restrict__ n_runs, const uint64_t* __restrict__ cigar_off,
            char* __restrict__ cigar) {
    const uint32_t p = first + blockIdx.x * blockDim.x + threadIdx.x;   // pairs [first, n)
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

// One warp per pair (long pairs: thousands of runs each): 32 runs at a time, back to front; a warp scan of
// the runs' text lengths gives every lane its place.
__global__ void __launch_bounds__(128)
emit_warp_kernel(const PairDesc* __restrict__ pairs, uint32_t first, uint32_t n, const uint32_t* __restrict__ runs,
                 const uint32_t* __restrict__ n_runs, const uint64_t* __restrict__ cigar_off,
                 char* __restrict__ cigar) {
    const uint32_t p = first + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);   // pairs [first, n)
    const int lane = threadIdx.x & 31;
    if (p >= n) return;
    const uint32_t nr = n_runs[p];
    char* dst = cigar + cigar_off[p];
    if (nr == 0) { if (lane == 0) { dst[0] = '1'; dst[1] = '\0'; } return; }
    const uint32_t* src = runs + pairs[p].run_off;
    uint32_t at = 0;   // bytes already placed (warp-uniform)
    for (uint32_t base = 0; base < nr; base += kWarp) {
        const uint32_t t = base + lane;              // t-th run of the text = run nr-1-t of the walk
        uint32_t cnt = 0, op = 0, len = 0;
        if (t < nr) { const uint32_t rw = src[nr - 1 - t]; cnt = rw >> 2; op = rw & 3u; len = dec_digits(cnt) + 1; }
        uint32_t incl = len;
#pragma unroll
        for (int o = 1; o < kWarp; o <<= 1) { const uint32_t u = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += u; }
        if (t < nr) {
            char* d = dst + at + incl - len;
            const uint32_t nd = len - 1;
            for (uint32_t k = nd; k-- > 0;) { d[k] = (char)('0' + cnt % 10); cnt /= 10; }
            d[nd] = (op == 0) ? 'M' : (op == 1 ? 'I' : 'D');
        }
        at += __shfl_sync(kFull, incl, kWarp - 1);
    }
}

}  // namespace b200
