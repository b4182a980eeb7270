// minimize.cuh -- batched (k,w) minimizer extraction, one output tuple per window exactly
// as team::KMER::Minimize emits them (team_minimizers/team_minimizers.cpp:122-225):
//   section 1 (:146-170)  s = 1..w-1   leftmost min over k-mers [0, s-1]
//   section 2 (:173-194)  e = w-1..n-1 leftmost min over k-mers [e-w+1, e]
//   section 3 (:197-222)  s = 1..min(w-1,n) leftmost min over k-mers [n-s, n-1]
// hash = 2-bit shift-in, C=0 A=1 T=2 G=3, other bytes 0, 32-bit truncation (:70-86);
// a window whose minimum is 0xFFFFFFFF yields the zero tuple (:106-120).
//
// HBM-bound by design: L bytes in, 9 bytes per window out -- so the kernel is written to spend as few
// instructions per window as it can (the first version spent ~80 and ran at a quarter of the copy rate):
//   * a CTA owns one 2048-aligned slice of the GLOBAL output index space (cut at sequence ends), so that
//     every thread's chunk of 8 consecutive tuples is 32-byte aligned in the hash/pos arrays and 8-byte
//     aligned in the flag array: five vector stores per 8 tuples;
//   * the bases the slice needs are staged once into shared memory as 2-bit codes (16 per word, first
//     base most significant), four bases at a time in SWAR form;
//   * a thread pulls three packed words, lines them up with two funnel shifts and gets each of the
//     8+w-1 k-mer hashes its windows touch with one more funnel shift and a shift;
//   * the leftmost minima of the 8 overlapping windows share partial minima (pairs, then quads, for w = 5).
// Chunks that touch section 1 or 3, a sequence end or a slice edge take a per-tuple path.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kMinThreads = 256;
constexpr int kMinTile = 2048;   // output tuples per CTA (8 per thread)
constexpr int kMinMaxW = 8;      // largest window length with a register fast path

struct MinTile {
    uint32_t seq;    // sequence index
    uint32_t first;  // first output slot (within the sequence) of this tile; it ends at the next multiple of
                     // kMinTile in the global output index space, or at the end of the sequence
};

__device__ __forceinline__ uint32_t base_code(uint32_t c) {
    // C=0 A=1 T=2 G=3, everything else 0 (reference :73-78, operator[] default)
    return (c == 'A') ? 1u : (c == 'T') ? 2u : (c == 'G') ? 3u : 0u;
}

__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// four ASCII bases (first base in the low byte) -> 8 bits, first base in the two most significant bits
__device__ __forceinline__ uint32_t code4(uint32_t v) {
    const uint32_t c4 = (v >> 1) & 0x03030303u;                               // A=0 C=1 T=2 G=3 per byte
    const uint32_t sel = (c4 & 0x3u) | ((c4 >> 4) & 0x30u) | ((c4 >> 8) & 0x300u) | ((c4 >> 12) & 0x3000u);
    if (prmt_b32(0x47544341u, 0u, sel) == v) {                                  // re-encoding "ACTG"[code] gives the bytes back
        const uint32_t m4 = c4 ^ ((~c4 >> 1) & 0x01010101u);                    // -> C=0 A=1 T=2 G=3
        return (m4 * 0x40100401u) >> 24;                                        // b0<<6 | b1<<4 | b2<<2 | b3
    }
    uint32_t c8 = 0;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) c8 = (c8 << 2) | base_code((v >> (8 * bb)) & 0xffu);
    return c8;
}

// leftmost strict minimum: b (the right-hand candidate) only wins when strictly smaller
__device__ __forceinline__ void lmin(uint32_t& v, uint32_t& at, uint32_t bv, uint32_t bat) {
    const bool take = bv < v;
    v = take ? bv : v;
    at = take ? bat : at;
}

// W = compile-time window length for the register fast path (1..kMinMaxW), or 0: every tuple takes the
// per-tuple path (any w; also used when the output pointers are not 16-byte aligned).
template <int W>
__global__ void __launch_bounds__(kMinThreads)
minimize_kernel(const uint8_t* __restrict__ buf, const uint64_t* __restrict__ off,
                const uint64_t* __restrict__ out_off, const uint8_t* __restrict__ is_fwd,
                const MinTile* __restrict__ tiles, uint32_t k, uint32_t w,
                uint32_t* __restrict__ hash, uint32_t* __restrict__ pos, uint8_t* __restrict__ flag) {
    extern __shared__ uint32_t codes[];
    const MinTile tl = tiles[blockIdx.x];
    const uint64_t s_off = off[tl.seq];
    const uint32_t L = (uint32_t)(off[tl.seq + 1] - s_off);
    const uint8_t* seq = buf + s_off;
    const uint64_t n = (uint64_t)L - k + 1;                 // k-mers inside the sequence (L >= k here)
    const uint64_t full = n >= w ? n - w + 1 : 0;
    const uint64_t tail = n < (uint64_t)w - 1 ? n : (uint64_t)w - 1;
    const uint64_t total = (uint64_t)(w - 1) + full + tail;
    const uint64_t obase = out_off[tl.seq];
    const uint64_t o0 = tl.first;
    const uint64_t g0 = obase + o0;                          // global index of the tile's first tuple
    const uint64_t o1 = min(total, o0 + ((uint64_t)kMinTile - (g0 & (uint64_t)(kMinTile - 1))));

    // k-mer index range this tile can touch: [x0, x1)
    const uint64_t x0 = o0 > 2ull * w ? o0 - 2ull * w : 0;
    const uint64_t x1 = o1 + 1;                              // sections 1/2 use k-mers <= slot index
    const uint32_t nx = (uint32_t)(x1 - x0);                 // <= kMinTile + 2w + 1
    const uint32_t kk = k < 16 ? k : 16;                     // bases that survive in the 32-bit hash
    const uint32_t hshift = 32 - 2 * kk;                     // (k = 0 never reaches the shifts below)
    const uint32_t nwords = (nx + k - 1 + 15) / 16 + 2;      // two spare words: hashes read one and two words ahead

    // 16 bases -> one packed word, from aligned 32-bit words realigned by funnel shifts; bytes past the
    // end of the sequence read as code 0
    for (uint32_t wi = threadIdx.x; wi < nwords; wi += blockDim.x) {
        const uint64_t b0 = x0 + (uint64_t)wi * 16;
        uint32_t word = 0;
        if (b0 < L) {
            const uint32_t nb = (uint32_t)min((uint64_t)16, L - b0);
            const uintptr_t addr = reinterpret_cast<uintptr_t>(seq + b0);
            const uint32_t* aw = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
            const uint32_t shb = (uint32_t)(addr & 3u) * 8u;
            uint32_t raw[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) raw[q] = ((uint32_t)q * 4u < (uint32_t)(addr & 3u) + nb) ? __ldg(aw + q) : 0u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t v = __funnelshift_r(raw[q], raw[q + 1], shb);   // bases 4q..4q+3, first base in the low byte
                if (nb < 4u * q + 4u) {                                  // ragged end of the sequence
                    const uint32_t valid = nb > 4u * q ? nb - 4u * q : 0u;
                    v = valid ? (v & (0xffffffffu >> (8u * (4u - valid)))) : 0u;
                }
                word |= code4(v) << (8 * (3 - q));
            }
        }
        codes[wi] = word;
    }
    __syncthreads();

    const uint32_t fl = is_fwd[tl.seq] ? 1u : 0u;
    // hash of k-mer x (absolute index): its last kk bases, first of them at base x + k - kk
    auto hash_at = [&](uint64_t x) -> uint32_t {
        if (kk == 0) return 0u;
        const uint32_t b = (uint32_t)(x - x0) + (k - kk);
        const uint32_t top = __funnelshift_l(codes[(b >> 4) + 1], codes[b >> 4], (b & 15u) * 2);
        return top >> hshift;
    };
    auto slow_one = [&](uint64_t o) {
        uint64_t a, b;  // window of k-mer indices [a, b]
        if (o < (uint64_t)w - 1) { a = 0; b = o; }
        else if (o < (uint64_t)w - 1 + full) { a = o - (w - 1); b = o; }
        else { const uint64_t s = o - ((uint64_t)w - 1 + full) + 1; a = n - s; b = n - 1; }
        uint32_t mn = 0xffffffffu, mpos = 0;
        for (uint64_t x = a; x <= b; ++x) {
            const uint32_t h = hash_at(x);
            if (h < mn) { mn = h; mpos = (uint32_t)x + 1; }
        }
        const bool none = (mpos == 0);                       // every hash was 0xFFFFFFFF
        hash[obase + o] = none ? 0u : mn;
        pos[obase + o] = mpos;
        flag[obase + o] = none ? 0 : (uint8_t)fl;
    };

    const uint64_t gbase = g0 & ~(uint64_t)(kMinTile - 1);
    for (uint32_t chunk = threadIdx.x; chunk < kMinTile / 8; chunk += blockDim.x) {
        const uint64_t gc = gbase + 8ull * chunk;
        const uint64_t lo = max(gc, g0), hi = min(gc + 8, obase + o1);
        if (lo >= hi) continue;
        bool fast = false;
        uint64_t oc = 0;
        if (W > 0) {
            oc = gc - obase;   // valid when gc >= obase, which lo == gc implies
            fast = lo == gc && hi == gc + 8 && kk != 0 && oc >= (uint64_t)(W - 1) && oc + 8 <= (uint64_t)(W - 1) + full;
        }
        if (!fast) {
            for (uint64_t g = lo; g < hi; ++g) slow_one(g - obase);
            continue;
        }
        if (W > 0) {
            constexpr int NH = 8 + (W > 0 ? W : 1) - 1;       // hashes the 8 windows touch
            const uint64_t xa = oc - (W - 1);                 // first k-mer of the first window
            const uint32_t b = (uint32_t)(xa - x0) + (k - kk);
            const uint32_t w0 = codes[b >> 4], w1 = codes[(b >> 4) + 1], w2 = codes[(b >> 4) + 2];
            const uint32_t sh0 = (b & 15u) * 2;
            const uint32_t A0 = __funnelshift_l(w1, w0, sh0), A1 = __funnelshift_l(w2, w1, sh0);   // 32 bases from b on
            uint32_t h[NH];
#pragma unroll
            for (int t = 0; t < NH; ++t) h[t] = __funnelshift_l(A1, A0, 2 * t) >> hshift;
            uint32_t mv[8], ma[8];
            if (W == 5) {   // shared partial minima: pairs, quads, then the fifth k-mer
                uint32_t pv[11], pa[11];
#pragma unroll
                for (int t = 0; t < 11; ++t) { pv[t] = h[t]; pa[t] = t; lmin(pv[t], pa[t], h[t + 1], t + 1); }
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    mv[t] = pv[t]; ma[t] = pa[t];
                    lmin(mv[t], ma[t], pv[t + 2], pa[t + 2]);
                    lmin(mv[t], ma[t], h[t + 4], t + 4);
                }
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    mv[t] = h[t]; ma[t] = t;
#pragma unroll
                    for (int u = 1; u < (W > 0 ? W : 1); ++u) lmin(mv[t], ma[t], h[t + u], t + u);
                }
            }
            const uint32_t p0 = (uint32_t)xa + 1;             // reported positions are 1-based k-mer indices
            uint32_t ho[8], po[8];
            uint32_t f0 = 0, f1 = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const bool none = mv[t] == 0xffffffffu;       // every hash of the window was 0xFFFFFFFF: the zero tuple
                ho[t] = none ? 0u : mv[t];
                po[t] = none ? 0u : p0 + ma[t];
                const uint32_t fb = none ? 0u : fl;
                if (t < 4) f0 |= fb << (8 * t); else f1 |= fb << (8 * (t - 4));
            }
            uint4* hp = reinterpret_cast<uint4*>(hash + gc);
            uint4* pp = reinterpret_cast<uint4*>(pos + gc);
            __stcs(hp, make_uint4(ho[0], ho[1], ho[2], ho[3]));
            __stcs(hp + 1, make_uint4(ho[4], ho[5], ho[6], ho[7]));
            __stcs(pp, make_uint4(po[0], po[1], po[2], po[3]));
            __stcs(pp + 1, make_uint4(po[4], po[5], po[6], po[7]));
            __stcs(reinterpret_cast<uint2*>(flag + gc), make_uint2(f0, f1));
        }
    }
}

}  // namespace b200
