// mapper.cuh -- device side of the per-read mapping pipeline that feeds Align
// (reference team_mapper.cpp): reverse complement :49-63, minimizer index :412-477,
// remove_duplicates :28-45, seed lookup :627-638 / :716-729, FindLIS chaining :283-316,
// strand choice and region :639-656. Everything here is integer/byte work, HBM- or latency-bound.
#pragma once
#include "common.cuh"

namespace b200 {

// ---- reverse complement (A<->T, C<->G, every other byte unchanged: the reference's switch has no default)
__global__ void revcomp_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t len) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    uint8_t c = src[len - 1 - i];
    c = (c == 'A') ? 'T' : (c == 'T') ? 'A' : (c == 'G') ? 'C' : (c == 'C') ? 'G' : c;
    dst[i] = c;
}

// ---- index: (hash, pos) tuples of one strand -> 64-bit sort keys hash<<32 | pos
__global__ void make_keys_kernel(const uint32_t* __restrict__ hash, const uint32_t* __restrict__ pos, uint64_t n,
                                 uint64_t* __restrict__ keys) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((uint64_t)hash[i] << 32) | pos[i];
}

// frequency ranking of the forward hashes: ascending order of (~count << 32 | hash) = count descending, hash ascending
__global__ void freq_keys_kernel(const uint32_t* __restrict__ hash, const uint32_t* __restrict__ count, uint32_t n,
                                 uint64_t* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((uint64_t)(0xffffffffu - count[i]) << 32) | hash[i];
}
__global__ void low_words_kernel(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)keys[i];
}

// drop keys whose hash is in the (sorted) ban list; flags feed a stream compaction
__global__ void ban_flag_kernel(const uint64_t* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ banned,
                                uint32_t n_banned, uint8_t* __restrict__ keep) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t h = (uint32_t)(keys[i] >> 32);
    uint32_t lo = 0, hi = n_banned;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (banned[mid] < h) lo = mid + 1; else hi = mid; }
    keep[i] = !(lo < n_banned && banned[lo] == h);
}

__device__ __forceinline__ uint32_t upper_read(const uint64_t* __restrict__ off, uint32_t n, uint64_t g) {
    // last r with off[r] <= g
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (off[mid] <= g) lo = mid; else hi = mid; }
    return lo;
}

// ---- remove_duplicates (:28-45): keep the first occurrence of every (hash, pos, flag) tuple of a read.
// In the begin + full-window sections the minimizer position never moves backwards, so duplicates are
// adjacent; the <= w-1 end-section tuples are compared with the tuples shortly before them. The zero
// tuple (0,0,false) of an all-0xFFFFFFFF window can recur anywhere: its first occurrence is found with
// an atomicMin per read and resolved by dedup_sentinel_kernel.
__global__ void dedup_flag_kernel(const uint32_t* __restrict__ hash, const uint32_t* __restrict__ pos,
                                  const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ read_off,
                                  uint32_t n_reads, uint32_t k, uint32_t w, uint8_t* __restrict__ keep,
                                  uint32_t* __restrict__ first_sentinel) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= out_off[n_reads]) return;
    const uint32_t r = upper_read(out_off, n_reads, g);
    const uint64_t o = g - out_off[r];
    const uint32_t p = pos[g], h = hash[g];
    if (p == 0) { atomicMin(first_sentinel + r, (uint32_t)o); keep[g] = 2; return; }
    const uint64_t L = read_off[r + 1] - read_off[r];
    const uint64_t nk = L - k + 1;
    const uint64_t full = nk >= w ? nk - w + 1 : 0;
    const uint64_t sec12 = (uint64_t)(w - 1) + full;
    bool kp = true;
    if (o < sec12) {
        kp = (o == 0) || pos[g - 1] != p || hash[g - 1] != h;
    } else {
        const uint64_t back_max = 2ull * w + (o - sec12);
        const uint64_t back = o < back_max ? o : back_max;
        for (uint64_t b = 1; b <= back; ++b)
            if (pos[g - b] == p && hash[g - b] == h) { kp = false; break; }
    }
    keep[g] = kp ? 1 : 0;
}

__global__ void dedup_sentinel_kernel(const uint64_t* __restrict__ out_off, uint32_t n_reads,
                                      const uint32_t* __restrict__ first_sentinel, uint8_t* __restrict__ keep) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= out_off[n_reads] || keep[g] != 2) return;
    const uint32_t r = upper_read(out_off, n_reads, g);
    keep[g] = (g - out_off[r] == first_sentinel[r]) ? 1 : 0;
}

// scatter kept tuples to their compacted slot; record where each read's slice starts
__global__ void dedup_scatter_kernel(const uint32_t* __restrict__ hash, const uint32_t* __restrict__ pos,
                                     const uint8_t* __restrict__ keep, const uint32_t* __restrict__ slot, uint64_t n,
                                     uint32_t* __restrict__ dhash, uint32_t* __restrict__ dpos) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n || !keep[g]) return;
    dhash[slot[g]] = hash[g];
    dpos[slot[g]] = pos[g];
}

// value of an exclusive-scan array at each read boundary: out[r] = scan[off[r]] (out[n] = total)
__global__ void gather_offsets_kernel(const uint32_t* __restrict__ scan, const uint64_t* __restrict__ off, uint32_t n,
                                      uint64_t n_items, uint32_t total, uint32_t* __restrict__ out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    out[r] = (off[r] >= n_items) ? total : scan[off[r]];
}
__global__ void gather_offsets32_kernel(const uint32_t* __restrict__ scan, const uint32_t* __restrict__ off, uint32_t n,
                                        uint32_t n_items, uint32_t total, uint32_t* __restrict__ out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    out[r] = (off[r] >= n_items) ? total : scan[off[r]];
}

// ---- seed lookup: one thread per de-duplicated read minimizer; index = sorted unique hash<<32|pos keys
__device__ __forceinline__ uint64_t lower_bound64(const uint64_t* __restrict__ a, uint64_t n, uint64_t v) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void seed_count_kernel(const uint32_t* __restrict__ dhash, uint32_t n_min,
                                  const uint64_t* __restrict__ kf, uint64_t nf, const uint64_t* __restrict__ kr,
                                  uint64_t nr, int rev_requires_fwd, uint32_t* __restrict__ cnt_f,
                                  uint32_t* __restrict__ cnt_r, uint64_t* __restrict__ lo_f, uint64_t* __restrict__ lo_r) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_min) return;
    const uint64_t h = dhash[m];
    const bool top = (h == 0xffffffffull);   // (h + 1) << 32 would wrap
    const uint64_t a = lower_bound64(kf, nf, h << 32), b = top ? nf : lower_bound64(kf, nf, (h + 1) << 32);
    const uint64_t c = lower_bound64(kr, nr, h << 32), d = top ? nr : lower_bound64(kr, nr, (h + 1) << 32);
    const uint32_t cf = (uint32_t)(b - a);
    // FASTA input (:630-637) only consults the reverse index for hashes present in the forward index
    const uint32_t cr = (rev_requires_fwd && cf == 0) ? 0u : (uint32_t)(d - c);
    cnt_f[m] = cf; cnt_r[m] = cr; lo_f[m] = a; lo_r[m] = c;
}

__global__ void seed_emit_kernel(const uint32_t* __restrict__ dpos, uint32_t n_min, const uint64_t* __restrict__ keys,
                                 const uint32_t* __restrict__ cnt, const uint64_t* __restrict__ lo,
                                 const uint32_t* __restrict__ moff, uint32_t* __restrict__ mf, uint32_t* __restrict__ ms) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_min) return;
    const uint32_t c = cnt[m], o = moff[m], fp = dpos[m];
    for (uint32_t x = 0; x < c; ++x) { mf[o + x] = fp; ms[o + x] = (uint32_t)keys[lo[m] + x]; }   // ref positions ascending
}

// ---- FindLIS (:283-316): one warp per (read, strand). lis[i] = 1 + max lis[j] over j < i with
//   s[i] > s[j], f[i] != f[j], f[i]-f[j] < 5000, s[i]-s[j] < 5000 (unsigned), earliest such j kept;
// chain end = first index of the maximum; only its length and its first/last match leave the kernel.
struct ChainResult {
    uint32_t len, first_f, first_s, last_f, last_s, pad;
};

// Matches are taken 32 at a time (one per lane). For a block [i0, i0+32) every lane first scans the
// finished prefix j < i0 on its own: the prefix is fetched 32 entries at a time (coalesced) and handed
// round by shuffle, so there is no reduction per match; then the 32 in-block predecessors are handed
// round in index order. The distance test f[i]-f[j] < 5000 makes most of a long read's prefix
// irrelevant: with pm[j] = max f[0..j] (one scan up front), every j with pm[j] <= min_block(f[i]) - 5000
// fails the test for the whole block and is skipped -- exact whatever the order of f, and f is in fact
// ascending except for the last w-1 minimizers of a read.
__global__ void __launch_bounds__(128)
chain_kernel(const uint32_t* __restrict__ mf, const uint32_t* __restrict__ ms, const uint32_t* __restrict__ roff,
             uint32_t n_reads, uint32_t* lis, int32_t* prev, uint32_t* pmax, ChainResult* __restrict__ out) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    const uint32_t a = roff[r], n = roff[r + 1] - a;
    if (n == 0) { if (lane == 0) out[r] = ChainResult{0, 0, 0, 0, 0, 0}; return; }
    const uint32_t* f = mf + a;
    const uint32_t* s = ms + a;
    uint32_t* L = lis + a;
    int32_t* P = prev + a;
    uint32_t* PM = pmax + a;
    {   // prefix maxima of f
        uint32_t carry = 0;
        for (uint32_t c = 0; c < n; c += kWarp) {
            uint32_t v = c + lane < n ? __ldg(f + c + lane) : 0u;
#pragma unroll
            for (int o = 1; o < kWarp; o <<= 1) { const uint32_t u = __shfl_up_sync(kFull, v, o); if (lane >= o) v = max(v, u); }
            v = max(v, carry);
            if (c + lane < n) __stcg(PM + c + lane, v);
            carry = __shfl_sync(kFull, v, kWarp - 1);
        }
        __syncwarp();
    }
    uint32_t mx = 0, mi = 0xffffffffu;   // running first maximum (std::max_element)
    uint32_t jstart = 0, last_min = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += kWarp) {
        const uint32_t i = i0 + lane;
        const bool valid = i < n;
        const uint32_t fi = valid ? __ldg(f + i) : 0u, si = valid ? __ldg(s + i) : 0u;
        uint32_t fmin = valid ? fi : 0xffffffffu;
#pragma unroll
        for (int o = 16; o; o >>= 1) fmin = min(fmin, __shfl_xor_sync(kFull, fmin, o));
        if (fmin < last_min) jstart = 0;
        last_min = fmin;
        if (fmin >= 5000u) {   // skip the prefix that is out of reach of every match of this block
            const uint32_t thr = fmin - 5000u;
            for (;;) {
                const uint32_t idx = jstart + lane;
                const bool skip = idx < i0 && __ldcg(PM + idx) <= thr;
                const uint32_t b = __ballot_sync(kFull, skip);
                if (b == kFull) { jstart += kWarp; continue; }
                jstart += __ffs(~b) - 1;
                break;
            }
        }
        uint32_t best = 0, bj = 0xffffffffu;
        for (uint32_t c = jstart; c < i0; c += kWarp) {
            const uint32_t idx = c + lane;
            const uint32_t cf = idx < i0 ? __ldg(f + idx) : 0u, cs = idx < i0 ? __ldg(s + idx) : 0u;
            const uint32_t cl = idx < i0 ? __ldcg(L + idx) : 0u;
            const uint32_t cnt = min((uint32_t)kWarp, i0 - c);
            for (uint32_t t = 0; t < cnt; ++t) {
                const uint32_t fj = __shfl_sync(kFull, cf, (int)t), sj = __shfl_sync(kFull, cs, (int)t);
                const uint32_t lj = __shfl_sync(kFull, cl, (int)t);
                if (si > sj && fi != fj && (fi - fj) < 5000u && (si - sj) < 5000u && lj + 1 > best) { best = lj + 1; bj = c + t; }
            }
        }
        const uint32_t in_block = min((uint32_t)kWarp, n - i0);
        for (uint32_t jj = 0; jj + 1 < in_block; ++jj) {
            // lane jj has seen every j < i0 + jj: its value is final
            const uint32_t fj = __shfl_sync(kFull, fi, (int)jj), sj = __shfl_sync(kFull, si, (int)jj);
            const uint32_t lj = __shfl_sync(kFull, best ? best : 1u, (int)jj);
            if ((uint32_t)lane > jj && si > sj && fi != fj && (fi - fj) < 5000u && (si - sj) < 5000u && lj + 1 > best) {
                best = lj + 1; bj = i0 + jj;
            }
        }
        const uint32_t li = best ? best : 1u;
        if (valid) {
            __stcg(L + i, li);
            P[i] = best ? (int32_t)bj : -1;
            if (li > mx) { mx = li; mi = i; }   // per lane: indices ascend, so the first maximum is kept
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const uint32_t ob = __shfl_xor_sync(kFull, mx, o), oi = __shfl_xor_sync(kFull, mi, o);
        if (ob > mx || (ob == mx && oi < mi)) { mx = ob; mi = oi; }
    }
    if (lane == 0) {
        uint32_t i = mi;
        while (P[i] >= 0) i = (uint32_t)P[i];
        out[r] = ChainResult{mx, f[i], s[i], f[mi], s[mi], 0};
    }
}

// ---- strand choice + region (:639-656): forward wins ties; q/t ranges are inclusive, 0-based
struct Region {
    uint32_t mapped, fwd, q_begin, q_end, t_begin, t_end;
};

__global__ void region_kernel(const ChainResult* __restrict__ cf, const ChainResult* __restrict__ cr, uint32_t n_reads,
                              uint32_t k, const uint64_t* __restrict__ read_off, uint64_t ref_len,
                              Region* __restrict__ out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const bool fwd = cf[r].len >= cr[r].len;
    const ChainResult c = fwd ? cf[r] : cr[r];
    Region g{0, fwd ? 1u : 0u, 0, 0, 0, 0};
    if (c.len) {
        g.q_begin = c.first_f - 1; g.q_end = c.last_f + k - 2;
        g.t_begin = c.first_s - 1; g.t_end = c.last_s + k - 2;
        const uint64_t L = read_off[r + 1] - read_off[r];
        // a position-0 sentinel match would make the reference index out of bounds (UB there): drop the read
        g.mapped = (g.q_begin <= g.q_end && g.q_end < L && g.t_begin <= g.t_end && g.t_end < ref_len) ? 1u : 0u;
    }
    out[r] = g;
}

}  // namespace b200
