// minimize.cuh -- batched (k,w) minimizer extraction, one output tuple per window exactly
// as team::KMER::Minimize emits them (team_minimizers/team_minimizers.cpp:122-225):
//   section 1 (:146-170)  s = 1..w-1   leftmost min over k-mers [0, s-1]
//   section 2 (:173-194)  e = w-1..n-1 leftmost min over k-mers [e-w+1, e]
//   section 3 (:197-222)  s = 1..min(w-1,n) leftmost min over k-mers [n-s, n-1]
// hash = 2-bit shift-in, C=0 A=1 T=2 G=3, other bytes 0, 32-bit truncation (:70-86);
// a window whose minimum is 0xFFFFFFFF yields the zero tuple (:106-120).
//
// HBM-bound: L bytes in, 9 bytes per window out. One CTA owns a tile of kTile consecutive
// output slots of one sequence: it stages the bases it needs into shared memory as 2-bit
// codes (16 per word, first base most significant), derives every k-mer hash of the tile
// with one funnel shift, then each thread takes the leftmost minimum of its window. The
// output slot of a window is a closed-form function of its index, so stores are dense and
// coalesced; the three output arrays are structure-of-arrays.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kMinThreads = 256;
constexpr int kMinTile = 2048;  // output tuples per CTA

struct MinTile {
    uint32_t seq;    // sequence index
    uint32_t first;  // first output slot (within the sequence) of this tile
};

__device__ __forceinline__ uint32_t base_code(uint32_t c) {
    // C=0 A=1 T=2 G=3, everything else 0 (reference :73-78, operator[] default)
    return (c == 'A') ? 1u : (c == 'T') ? 2u : (c == 'G') ? 3u : 0u;
}

// smem: codes[] packed words, then hashes[]
template <int DUMMY = 0>
__global__ void __launch_bounds__(kMinThreads)
minimize_kernel(const uint8_t* __restrict__ buf, const uint64_t* __restrict__ off,
                const uint64_t* __restrict__ out_off, const uint8_t* __restrict__ is_fwd,
                const MinTile* __restrict__ tiles, uint32_t k, uint32_t w,
                uint32_t* __restrict__ hash, uint32_t* __restrict__ pos, uint8_t* __restrict__ flag) {
    extern __shared__ uint32_t smem[];
    const MinTile tl = tiles[blockIdx.x];
    const uint64_t s_off = off[tl.seq];
    const uint32_t L = (uint32_t)(off[tl.seq + 1] - s_off);
    const uint8_t* seq = buf + s_off;
    const uint64_t n = (uint64_t)L - k + 1;                 // k-mers inside the sequence (L >= k here)
    const uint64_t full = n >= w ? n - w + 1 : 0;
    const uint64_t tail = n < (uint64_t)w - 1 ? n : (uint64_t)w - 1;
    const uint64_t total = (uint64_t)(w - 1) + full + tail;
    const uint64_t o0 = tl.first;
    const uint64_t o1 = min(o0 + (uint64_t)kMinTile, total);

    // k-mer index range this tile can touch: [x0, x1)
    const uint64_t x0 = o0 > 2ull * w ? o0 - 2ull * w : 0;
    const uint64_t x1 = o1 + 1;                              // sections 1/2 use k-mers <= slot index
    const uint32_t nx = (uint32_t)(x1 - x0);                 // <= kMinTile + 2w + 1
    // bases needed: [x0, x1 + k - 1); keep the packed words 16-base aligned relative to x0
    const uint32_t kk = k < 16 ? k : 16;                     // bases that survive in the 32-bit hash
    const uint32_t nb = nx + k - 1;
    const uint32_t nwords = (nb + 15) / 16 + 1;
    uint32_t* codes = smem;
    uint32_t* hs = smem + nwords;

    // 16 bases -> one packed word. Bytes come in as aligned 32-bit words realigned by funnel shifts
    // (five loads per 16 bases instead of sixteen byte loads); out-of-range bytes read as code 0.
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
                const uint32_t valid = nb > 4u * q ? min(4u, nb - 4u * q) : 0u;
                if (valid < 4u) v = valid ? (v & (0xffffffffu >> (8u * (4u - valid)))) : 0u;
                uint32_t c8 = 0;   // first base in the two most significant bits
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) c8 = (c8 << 2) | base_code((v >> (8 * bb)) & 0xffu);
                word |= c8 << (8 * (3 - q));
            }
        }
        codes[wi] = word;
    }
    __syncthreads();
    for (uint32_t xi = threadIdx.x; xi < nx; xi += blockDim.x) {
        // k-mer x = x0 + xi covers bases [x, x+k); its hash keeps the last kk of them
        const uint32_t b = xi + (k - kk);                    // first surviving base, relative to x0
        const uint32_t hi = codes[b >> 4], lo = codes[(b >> 4) + 1];
        const uint32_t sh = (b & 15u) * 2;
        const uint32_t top = __funnelshift_l(lo, hi, sh);    // 16 bases starting at b
        hs[xi] = kk == 16 ? top : (kk == 0 ? 0u : (top >> (32 - 2 * kk)));
    }
    __syncthreads();

    const uint8_t fl = is_fwd[tl.seq] ? 1 : 0;
    const uint64_t obase = out_off[tl.seq];
    for (uint64_t o = o0 + threadIdx.x; o < o1; o += blockDim.x) {
        uint64_t a, b;  // window of k-mer indices [a, b]
        if (o < (uint64_t)w - 1) { a = 0; b = o; }
        else if (o < (uint64_t)w - 1 + full) { a = o - (w - 1); b = o; }
        else { const uint64_t s = o - ((uint64_t)w - 1 + full) + 1; a = n - s; b = n - 1; }
        uint32_t mn = 0xffffffffu, mpos = 0;
        for (uint64_t x = a; x <= b; ++x) {
            const uint32_t h = hs[(uint32_t)(x - x0)];
            if (h < mn) { mn = h; mpos = (uint32_t)x + 1; }
        }
        const bool none = (mpos == 0);                       // every hash was 0xFFFFFFFF
        hash[obase + o] = none ? 0u : mn;
        pos[obase + o] = mpos;
        flag[obase + o] = none ? 0 : fl;
    }
}

}  // namespace b200
