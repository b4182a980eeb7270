// host_pack.hpp -- 2-bit packing of a sequence on the HOST, in the layout pack_kernel (align_fill_short.cuh) writes on the
// device: word w holds bases 16w .. 16w+15, base k at bits [2k, 2k+1], code = (byte >> 1) & 3 (A 0, C 1, T 2, G 3), bases
// past the end of the sequence are code 0. Used by the pointer-array entry point: its gather pass touches every byte on
// the host anyway, so it writes 2-bit words instead of copies and the upload shrinks from Q bytes to 4 (Q/16 + 2).
// A byte outside "ACGT" makes the sequence `flagged` (bit 1: a '-', bit 2: any other byte), exactly as pack_kernel flags
// it: such pairs go to the byte-compare kernel, which reads the raw bytes.
// SSE2 only (baseline x86-64): no compile flags, no dispatch.
#pragma once
#include <emmintrin.h>

#include <cstdint>
#include <cstring>

namespace b200 {

// 16 bytes -> one packed word; *bad |= non-zero when a byte is not one of "ACGT".
static inline uint32_t host_pack16(const char* s, int* bad) {
    const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
    const __m128i b1 = _mm_set1_epi8(1), b3 = _mm_set1_epi8(3), b6 = _mm_set1_epi8(6);
    const __m128i h = _mm_srli_epi16(v, 1);                                   // (bits cross byte borders: masked below)
    // a byte is one of "ACGT" exactly when it equals (0x41 | (v & 6)) ^ (0x11 if bit 2 set and bit 1 clear)  [minimize.cuh]
    const __m128i t = _mm_and_si128(_mm_andnot_si128(h, _mm_srli_epi16(v, 2)), b1);        // (v >> 2) & ~(v >> 1) & 1
    const __m128i t11 = _mm_or_si128(t, _mm_slli_epi16(t, 4));                            // t * 0x11 (t is 0 / 1 per byte)
    const __m128i expect = _mm_xor_si128(_mm_or_si128(_mm_set1_epi8(0x41), _mm_and_si128(v, b6)), t11);
    *bad |= _mm_movemask_epi8(_mm_cmpeq_epi8(expect, v)) ^ 0xffff;
    const __m128i c = _mm_and_si128(h, b3);                                   // codes, one per byte
    const __m128i y = _mm_or_si128(c, _mm_srli_epi16(c, 6));                  // low byte of each 16-bit lane: c0 | c1 << 2
    const __m128i z = _mm_or_si128(y, _mm_srli_epi32(y, 12));                 // low byte of each 32-bit lane: 4 codes
    const __m128i lo = _mm_and_si128(z, _mm_set1_epi32(0xff));
    const __m128i p16 = _mm_packs_epi32(lo, lo);                              // 4 x (0..255) -> 16-bit lanes
    const __m128i p8 = _mm_packus_epi16(p16, p16);                            // -> bytes 0..3
    return (uint32_t)_mm_cvtsi128_si32(p8);
}

// Packs `len` bases into out[0 .. len/16 + 1] (the allocation of one sequence in the packed buffers; words past
// ceil(len/16) are zeroed). Returns the sequence's flag byte (0 = pure ACGT).
static inline uint8_t host_pack_sequence(const char* s, uint32_t len, uint32_t* out) {
    const uint32_t n_words = len / 16 + 2;
    int bad = 0;
    uint32_t w = 0;
    for (; (w + 1) * 16 <= len; ++w) out[w] = host_pack16(s + 16 * w, &bad);
    if (w * 16 < len) {                                                       // last, partial word: pad with 'A' (code 0)
        char tmp[16];
        std::memset(tmp, 'A', sizeof tmp);
        std::memcpy(tmp, s + 16 * w, len - 16 * w);
        out[w] = host_pack16(tmp, &bad);
        ++w;
    }
    for (; w < n_words; ++w) out[w] = 0;
    if (!bad) return 0;
    uint8_t flag = 2;                                                         // kFlagNonACGT
    for (uint32_t i = 0; i < len; ++i) if (s[i] == '-') { flag |= 1; break; } // kFlagDash
    return flag;
}

}  // namespace b200
