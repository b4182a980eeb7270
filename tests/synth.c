/* tests/synth.c -- seeded synthetic inputs of the BASELINE.json shapes (TEST / BENCH INFRASTRUCTURE, not product).
 *
 * One generator written once in C so that the same seeds give the same bytes from every caller (pytest,
 * bench.py on any rank, tools/), with no dependence on a numpy or libstdc++ distribution: all randomness
 * comes from splitmix64 streams keyed by (seed, item index), so any rank can generate any subset of a data
 * set without generating the rest, and ranges can be filled by several threads at once.
 *
 *   synth_dna          uniform ACGT (the 4.6 Mbp reference of configs 3/4: seed 1)
 *   synth_ont_spans    log-normal reference spans of ONT-like reads (mean 8 kb, clipped)
 *   synth_ont_reads    reads = substrings of the reference (random start and strand) with indel-heavy errors
 *   synth_pairs        fixed-length pairs: target uniform, query = target with errors, cut / padded to the length
 *
 * Build: gcc -O2 -shared -fPIC -o tests/libsynth.so tests/synth.c -lm   (done by __graft_entry__.build()).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static inline uint64_t sm64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t stream(uint64_t seed, uint64_t what, uint64_t index) {
    uint64_t s = seed * 0xD1342543DE82EF95ull + what * 0x2545F4914F6CDD1Dull;
    sm64(&s);
    s ^= index * 0x9E3779B97F4A7C15ull;
    sm64(&s);
    return s;
}
static inline double u01(uint64_t* s) { return (double)(sm64(s) >> 11) * (1.0 / 9007199254740992.0); }

static const uint8_t kBase[4] = {'A', 'C', 'G', 'T'};
static inline uint8_t comp(uint8_t c) { return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c; }

/* bases [first, first+count) of the uniform ACGT stream `seed` (32 bases per 64-bit draw, blocks of 32) */
void synth_dna(uint64_t seed, uint64_t first, uint64_t count, uint8_t* out) {
    uint64_t i = first;
    const uint64_t end = first + count;
    while (i < end) {
        const uint64_t blk = i >> 5;
        uint64_t s = stream(seed, 1, blk);
        const uint64_t bits = sm64(&s);
        const uint64_t stop = ((blk + 1) << 5) < end ? ((blk + 1) << 5) : end;
        for (; i < stop; ++i) out[i - first] = kBase[(bits >> (2 * (i & 31))) & 3];
    }
}

/* reference span of read i: log-normal with the given mean (sigma 0.5), clipped to [lo, hi] */
void synth_ont_spans(uint64_t seed, uint64_t i0, uint64_t i1, double mean, uint32_t lo, uint32_t hi, uint32_t* span) {
    const double sigma = 0.5, mu = log(mean) - 0.5 * sigma * sigma;
    for (uint64_t i = i0; i < i1; ++i) {
        uint64_t s = stream(seed, 2, i);
        const double u1 = u01(&s), u2 = u01(&s);
        const double z = sqrt(-2.0 * log(u1 > 1e-300 ? u1 : 1e-300)) * cos(6.283185307179586 * u2);
        double L = exp(mu + sigma * z);
        if (L < lo) L = lo;
        if (L > hi) L = hi;
        span[i - i0] = (uint32_t)L;
    }
}

/* errors applied base by base: with probability del the base is dropped, ins a random base is emitted before
 * it, sub it is replaced by a different base. Returns the output length; out may be NULL (dry run). */
static uint32_t mutate(uint64_t* s, const uint8_t* src, uint32_t n, int revcomp, double sub, double ins, double del,
                       uint8_t* out) {
    uint32_t m = 0;
    for (uint32_t x = 0; x < n; ++x) {
        uint8_t b = revcomp ? comp(src[n - 1 - x]) : src[x];
        const uint64_t r64 = sm64(s);
        const double r = (double)(r64 >> 11) * (1.0 / 9007199254740992.0);
        if (r < del) continue;
        if (r < del + ins) {
            if (out) out[m] = kBase[r64 & 3];
            ++m;
        } else if (r < del + ins + sub) {
            const uint32_t code = b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3;
            b = kBase[(code + 1 + (r64 & 0xff) % 3) & 3];
        }
        if (out) out[m] = b;
        ++m;
    }
    return m;
}

/* read i: start uniform in [0, ref_len - span], strand = bit, then errors. len_out[i - i0] receives the read
 * length; when out_buf is not NULL the bases go to out_buf + off[i - i0] (off from a prefix sum of a dry run). */
void synth_ont_reads(uint64_t seed, uint64_t i0, uint64_t i1, const uint8_t* ref, uint64_t ref_len, const uint32_t* span,
                     double sub, double ins, double del, uint32_t* len_out, uint8_t* out_buf, const uint64_t* off,
                     uint64_t* start_out, uint8_t* strand_out) {
    for (uint64_t i = i0; i < i1; ++i) {
        uint64_t s = stream(seed, 3, i);
        uint32_t sp = span[i - i0];
        if (sp > ref_len) sp = (uint32_t)ref_len;
        const uint64_t start = sm64(&s) % (ref_len - sp + 1);
        const int rc = (int)(sm64(&s) & 1);
        const uint32_t m = mutate(&s, ref + start, sp, rc, sub, ins, del, out_buf ? out_buf + off[i - i0] : NULL);
        if (len_out) len_out[i - i0] = m;
        if (start_out) start_out[i - i0] = start;
        if (strand_out) strand_out[i - i0] = (uint8_t)!rc;
    }
}

/* pairs [i0, i1): target = L uniform bases, query = target with errors, cut or padded (fresh bases) to exactly L;
 * written at (i - i0) * L in qbuf / tbuf */
void synth_pairs(uint64_t seed, uint64_t i0, uint64_t i1, uint32_t L, double sub, double ins, double del, uint8_t* qbuf,
                 uint8_t* tbuf, uint8_t* scratch /* 2L + 8 bytes */) {
    for (uint64_t i = i0; i < i1; ++i) {
        uint8_t* t = tbuf + (i - i0) * (uint64_t)L;
        uint8_t* q = qbuf + (i - i0) * (uint64_t)L;
        uint64_t s = stream(seed, 4, i);
        for (uint32_t x = 0; x < L; x += 32) {
            const uint64_t bits = sm64(&s);
            for (uint32_t y = x; y < L && y < x + 32; ++y) t[y] = kBase[(bits >> (2 * (y - x))) & 3];
        }
        const uint32_t m = mutate(&s, t, L, 0, sub, ins, del, scratch);
        for (uint32_t x = 0; x < L; ++x) q[x] = x < m ? scratch[x] : kBase[sm64(&s) & 3];
    }
}
