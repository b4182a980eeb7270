// minimize.cuh -- batched (k,w) minimizer extraction, one output tuple per window exactly
// as team::KMER::Minimize emits them (team_minimizers/team_minimizers.cpp:122-225):
//   section 1 (:146-170)  s = 1..w-1   leftmost min over k-mers [0, s-1]
//   section 2 (:173-194)  e = w-1..n-1 leftmost min over k-mers [e-w+1, e]
//   section 3 (:197-222)  s = 1..min(w-1,n) leftmost min over k-mers [n-s, n-1]
// hash = 2-bit shift-in, C=0 A=1 T=2 G=3, other bytes 0, 32-bit truncation (:70-86);
// a window whose minimum is 0xFFFFFFFF yields the zero tuple (:106-120).
//
// HBM-bound by design: L bytes in, 9 bytes per window out -- so the kernel is written to spend as few
// instructions per window as it can and to keep loads in flight while it computes:
//   * a CTA owns one 8192-aligned slice of the GLOBAL output index space (cut at sequence ends), so that
//     every thread's chunk of 8 consecutive tuples is 32-byte aligned in the hash/pos arrays and 8-byte
//     aligned in the flag array: five vector stores per 8 tuples;
//   * every WARP stages the bases of its own eighth of the slice into its own piece of shared memory as 2-bit
//     codes (16 per word, first base most significant): three aligned 128-bit loads in flight per lane, sixteen
//     bases validated and converted at a time in SWAR form; no CTA-wide barrier, so the warps of an SM drift
//     apart and one warp's load latency is another's compute time;
//   * a thread pulls three packed words, lines them up with two funnel shifts and gets each of the
//     8+w-1 k-mer hashes its windows touch with one more funnel shift and a shift;
//   * the leftmost minimum of a window is a chain of three-input minima for the value and, for the position,
//     the count of leading candidates that differ from it (fused add-min per candidate, a multiply-add chain).
// Chunks that touch section 1 or 3, a sequence end or a slice edge take a per-tuple path.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kMinThreads = 256;
constexpr int kMinTile = 8192;   // output tuples per CTA (four chunks of 8 per thread)
constexpr int kMinMaxW = 8;      // largest window length with a register fast path
constexpr int kMinWarpTuples = kMinTile / (kMinThreads / 32);   // slice positions per warp
constexpr int kMinStageDepth = 3;  // 16-byte chunks a lane has in flight while staging

// One CTA's work, precomputed by the host so that the kernel starts with ONE broadcast load instead of a
// chain of dependent ones (tile -> sequence offsets -> bases). The tile covers tuples [o0, o1) of a sequence;
// it ends at the next multiple of kMinTile in the global output index space, or at the end of the sequence.
struct MinTile {
    uint64_t src;    // byte offset of the sequence in the packed buffer
    uint64_t gout;   // global output index of tuple o0 (= out_off[seq] + o0)
    uint32_t L;      // sequence length
    uint32_t o0;     // first tuple of the tile, within the sequence
    uint32_t fwd;    // strand flag stamped on the tuples
    uint32_t pad;
};
static_assert(sizeof(MinTile) == 32, "MinTile layout");

__device__ __forceinline__ uint32_t base_code(uint32_t c) {
    // C=0 A=1 T=2 G=3, everything else 0 (reference :73-78, operator[] default)
    return (c == 'A') ? 1u : (c == 'T') ? 2u : (c == 'G') ? 3u : 0u;
}

// Four ASCII bases (first base in the low byte) -> 8 bits, first base in the two most significant bits.
// With h = v >> 1: the letters' codes are (h & 3) = A 0, C 1, T 2, G 3, and a byte is one of "ACGT" exactly when it
// equals (0x41 | (v & 6)) ^ (0x11 if bit 2 is set and bit 1 is not, i.e. T) -- checked over all 256 byte values.
__device__ __forceinline__ uint32_t acgt_mismatch4(uint32_t v, uint32_t h) {    // 0 iff all four bytes are in "ACGT"
    const uint32_t t = (v >> 2) & ~h & 0x01010101u;
    return ((0x41414141u | (v & 0x06060606u)) ^ (t * 0x11u)) ^ v;
}
__device__ __forceinline__ uint32_t pack4_valid(uint32_t h) {                   // h = v >> 1 of four valid letters
    const uint32_t c4 = h & 0x03030303u;
    const uint32_t m4 = c4 ^ ((~c4 >> 1) & 0x01010101u);                        // -> C=0 A=1 T=2 G=3
    return (m4 * 0x40100401u) >> 24;                                            // b0<<6 | b1<<4 | b2<<2 | b3
}
__device__ __forceinline__ uint32_t code4(uint32_t v) {
    const uint32_t h = v >> 1;
    if (acgt_mismatch4(v, h) == 0) return pack4_valid(h);
    uint32_t c8 = 0;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) c8 = (c8 << 2) | base_code((v >> (8 * bb)) & 0xffu);
    return c8;
}
// sixteen bases at once: one validity test for the whole 16-byte chunk
__device__ __forceinline__ uint32_t code16(const uint4& v) {
    const uint32_t hx = v.x >> 1, hy = v.y >> 1, hz = v.z >> 1, hw = v.w >> 1;
    if ((acgt_mismatch4(v.x, hx) | acgt_mismatch4(v.y, hy) | acgt_mismatch4(v.z, hz) | acgt_mismatch4(v.w, hw)) == 0)
        return (pack4_valid(hx) << 24) | (pack4_valid(hy) << 16) | (pack4_valid(hz) << 8) | pack4_valid(hw);
    return (code4(v.x) << 24) | (code4(v.y) << 16) | (code4(v.z) << 8) | code4(v.w);
}

// W = compile-time window length for the register fast path (1..kMinMaxW), or 0: every tuple takes the
// per-tuple path (any w; also used when the output pointers are not 16-byte aligned).
// K16: k >= 16, the only case in which a hash can be 0xFFFFFFFF (the register path then checks for the zero tuple).
template <int W, bool K16>
__global__ void __launch_bounds__(kMinThreads, 8)
minimize_kernel(const uint8_t* __restrict__ buf, const MinTile* __restrict__ tiles, uint32_t k, uint32_t w,
                uint64_t buf_bytes, uint32_t warp_words, uint32_t* __restrict__ hash, uint32_t* __restrict__ pos,
                uint8_t* __restrict__ flag) {
    extern __shared__ uint32_t codes_all[];
    const MinTile tl = tiles[blockIdx.x];
    const uint32_t L = tl.L;
    const uint32_t n = L - k + 1;                            // k-mers inside the sequence (L >= k here)
    const uint32_t full = n >= w ? n - w + 1 : 0;
    const uint32_t tail = n < w - 1 ? n : w - 1;
    const uint32_t total = (w - 1) + full + tail;            // (sequences are shorter than 2^32 - 16: no wrap)
    const uint32_t g_tile = (uint32_t)(tl.gout & (uint64_t)(kMinTile - 1));   // where the tile starts inside its aligned slice
    const uint32_t o1_tile = min(total - tl.o0, (uint32_t)kMinTile - g_tile) + tl.o0;
    hash += tl.gout - g_tile; pos += tl.gout - g_tile; flag += tl.gout - g_tile;  // slice-relative outputs: index g_in + (o - o0)
    // Every WARP stages and processes its own eighth of the slice (kMinWarpTuples slice positions, with the few bases
    // of overlap its windows need): no CTA-wide barrier between staging and use, so the warps of an SM drift apart and
    // one warp's load latency is another's compute time (with CTA-wide staging ncu showed the warps parked at the barrier
    // and on the loads: 6.7 + 8.4 stall cycles per issued instruction).
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t sp_lo = max(g_tile, (uint32_t)warp * kMinWarpTuples);
    const uint32_t sp_hi = min(g_tile + (o1_tile - tl.o0), (uint32_t)(warp + 1) * kMinWarpTuples);
    if (sp_lo >= sp_hi) return;
    const uint32_t o0 = tl.o0 + (sp_lo - g_tile), o1 = tl.o0 + (sp_hi - g_tile), g_in = sp_lo;   // the warp's tuples / first slice position
    uint32_t* codes = codes_all + (size_t)warp * warp_words;
    // k-mer index range this tile can touch: [x0, x1)
    const uint32_t x0 = o0 > 2 * w ? o0 - 2 * w : 0;
    const uint32_t nx = o1 + 1 - x0;                         // sections 1/2 use k-mers <= slot index; <= kMinTile + 2w + 1
    const uint32_t kk = k < 16 ? k : 16;                     // bases that survive in the 32-bit hash
    const uint32_t hshift = 32 - 2 * kk;                     // (k = 0 never reaches the shifts below)
    // Staging works on ALIGNED 16-byte chunks of the input (one 128-bit load and four SWAR conversions per
    // packed word, no realignment): packed word wi holds the 16 bytes at chunk origin + 16 wi, so base x0 sits
    // `mis` bases into word 0 and every base index below is shifted by that much.
    const uintptr_t addr0 = reinterpret_cast<uintptr_t>(buf) + tl.src + x0;
    const uint32_t mis = (uint32_t)(addr0 & 15u);
    const uint8_t* origin = reinterpret_cast<const uint8_t*>(addr0 - mis);
    const int32_t cb0 = (int32_t)x0 - (int32_t)mis;              // sequence index of the first byte of word 0 (>= -15)
    const uint32_t nwords = (mis + nx + k - 1 + 15) / 16 + 2;    // two spare words: hashes read one and two words ahead
    const uint64_t room = buf_bytes - tl.src;                    // bytes from the sequence start to the end of the buffer
    // (all of a lane's loads are issued before the first conversion: up to kMinStageDepth 16-byte chunks in flight)
    for (uint32_t w0i = 0; w0i < nwords; w0i += 32u * kMinStageDepth) {
        uint4 v[kMinStageDepth];
        bool whole[kMinStageDepth];
#pragma unroll
        for (int q = 0; q < kMinStageDepth; ++q) {
            const uint32_t wi = w0i + 32u * q + lane;
            const int64_t cb = (int64_t)cb0 + 16ll * wi;         // sequence index of this chunk's first byte
            whole[q] = wi < nwords && cb >= 0 && cb + 16 <= (int64_t)L && (uint64_t)(cb + 16) <= room;
            v[q] = whole[q] ? __ldg(reinterpret_cast<const uint4*>(origin + 16ull * wi)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int q = 0; q < kMinStageDepth; ++q) {
            const uint32_t wi = w0i + 32u * q + lane;
            if (wi >= nwords) continue;
            uint32_t word = 0;
            if (whole[q]) {
                word = code16(v[q]);
            } else {                                             // first / last chunk of the sequence: bytes outside it are code 0
                const int64_t cb = (int64_t)cb0 + 16ll * wi;
                if (cb < (int64_t)L) {
                    const uint8_t* src = origin + 16ull * wi;
#pragma unroll 1
                    for (int bb = 0; bb < 16; ++bb) {
                        const int64_t at = cb + bb;
                        const uint32_t c = (at >= 0 && at < (int64_t)L) ? (uint32_t)src[bb] : 0u;
                        word = (word << 2) | base_code(c);
                    }
                }
            }
            codes[wi] = word;
        }
    }
    __syncwarp();

    const uint32_t fl = tl.fwd;
    const uint32_t bshift = (k - kk) + mis - x0;             // base index in the staged words of k-mer x's first surviving base: x + bshift
    // hash of k-mer x (index within the sequence): its last kk bases
    auto hash_at = [&](uint32_t x) -> uint32_t {
        if (kk == 0) return 0u;
        const uint32_t b = x + bshift;
        const uint32_t top = __funnelshift_l(codes[(b >> 4) + 1], codes[b >> 4], (b & 15u) * 2);
        return top >> hshift;
    };
    auto slow_one = [&](uint32_t o) {
        uint32_t a, b;  // window of k-mer indices [a, b]
        if (o < w - 1) { a = 0; b = o; }
        else if (o < w - 1 + full) { a = o - (w - 1); b = o; }
        else { const uint32_t s = o - (w - 1 + full) + 1; a = n - s; b = n - 1; }
        uint32_t mn = 0xffffffffu, mpos = 0;
        for (uint32_t x = a; x <= b; ++x) {
            const uint32_t h = hash_at(x);
            if (h < mn) { mn = h; mpos = x + 1; }
        }
        const bool none = (mpos == 0);                       // every hash was 0xFFFFFFFF
        const uint32_t at = g_in + (o - o0);
        hash[at] = none ? 0u : mn;
        pos[at] = mpos;
        flag[at] = none ? 0 : (uint8_t)fl;
    };

    // chunk c of the slice = slice positions [8c, 8c+8) = tuples o0 - g_in + 8c ... of the sequence
    const uint32_t c_first = g_in >> 3, c_last = (g_in + (o1 - o0) + 7) >> 3;   // chunks that intersect the tile
    for (uint32_t chunk = c_first + lane; chunk < c_last; chunk += kWarp) {
        const uint32_t sp = 8 * chunk;                        // slice position of the chunk
        const uint32_t lo = max(sp, g_in), hi = min(sp + 8, g_in + (o1 - o0));
        const uint32_t oc = sp - g_in + o0;                   // tuple index of the chunk's first slot (wraps if sp < g_in: then lo != sp)
        bool fast = false;
        if (W > 0) fast = lo == sp && hi == sp + 8 && kk != 0 && oc >= (uint32_t)(W - 1) && oc + 8 <= (uint32_t)(W - 1) + full;
        if (!fast) {
            for (uint32_t q = lo; q < hi; ++q) slow_one(q - g_in + o0);
            continue;
        }
        if (W > 0) {
            constexpr int NH = 8 + (W > 0 ? W : 1) - 1;       // hashes the 8 windows touch
            const uint32_t xa = oc - (W - 1);                 // first k-mer of the first window
            const uint32_t b = xa + bshift;
            const uint32_t w0 = codes[b >> 4], w1 = codes[(b >> 4) + 1], w2 = codes[(b >> 4) + 2];
            const uint32_t sh0 = (b & 15u) * 2;
            const uint32_t A0 = __funnelshift_l(w1, w0, sh0), A1 = __funnelshift_l(w2, w1, sh0);   // 32 bases from b on
            uint32_t h[NH];
#pragma unroll
            for (int t = 0; t < NH; ++t) h[t] = __funnelshift_l(A1, A0, 2 * t) >> hshift;
            // Leftmost strict minimum of each of the 8 windows, without compare-and-select chains (they cost three
            // alu-pipe instructions per candidate): the VALUE is a chain of three-input minima; the POSITION is the
            // number of leading candidates that differ from it -- with n_u = min(h_u - m, 1) (0 where candidate u attains
            // the minimum, one fused add-min each) the offset of the first zero is n_0 (1 + n_1 (1 + n_2 ( ... ))), a
            // Horner chain of multiply-adds on the otherwise idle fma pipe.
            const uint32_t p0 = xa + 1;                       // reported positions are 1-based k-mer indices
            uint32_t ho[8], po[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                uint32_t m = h[t];
                int u = 1;
#pragma unroll
                for (; u + 1 < W; u += 2) m = __vimin3_u32(m, h[t + u], h[t + u + 1]);
                if (u < W) m = min(m, h[t + u]);
                const uint32_t negm = 0u - m;
                uint32_t P = 0;
#pragma unroll
                for (int v = W - 2; v >= 0; --v) {
                    const uint32_t nv = __viaddmin_u32(h[t + v], negm, 1u);   // h >= m: the difference does not wrap
                    P = nv * P + nv;
                }
                ho[t] = m;
                po[t] = p0 + t + P;
            }
            uint32_t f0 = fl * 0x01010101u, f1 = f0;
            if (K16) {   // only a 16-base hash can be 0xFFFFFFFF: a window of nothing else yields the zero tuple (:106-120)
                f0 = 0; f1 = 0;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const bool none = ho[t] == 0xffffffffu;
                    ho[t] = none ? 0u : ho[t];
                    po[t] = none ? 0u : po[t];
                    const uint32_t fb = none ? 0u : fl;
                    if (t < 4) f0 |= fb << (8 * t); else f1 |= fb << (8 * (t - 4));
                }
            }
            uint4* hp = reinterpret_cast<uint4*>(hash + sp);
            uint4* pp = reinterpret_cast<uint4*>(pos + sp);
            __stcs(hp, make_uint4(ho[0], ho[1], ho[2], ho[3]));
            __stcs(hp + 1, make_uint4(ho[4], ho[5], ho[6], ho[7]));
            __stcs(pp, make_uint4(po[0], po[1], po[2], po[3]));
            __stcs(pp + 1, make_uint4(po[4], po[5], po[6], po[7]));
            __stcs(reinterpret_cast<uint2*>(flag + sp), make_uint2(f0, f1));
        }
    }
}

}  // namespace b200
