/* oracle/oracle_minimize.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of team::KMER::Minimize (reference team_minimizers/team_minimizers.cpp
 * :122-225) with the helper semantics of MappSeqCharPointerToBit (:70-86) and
 * GetTupleWithMinFirst (:106-120).
 */
#include "oracle.h"
#include <stdlib.h>

/* :73-78 -- C=0 A=1 T=2 G=3, every other byte maps to 0 (operator[] default) */
static inline uint32_t code_of(const char* s, uint64_t at, uint32_t len) {
    if (at >= len) return 0; /* NUL padding past the end (M4) */
    switch (s[at]) { case 'A': return 1; case 'T': return 2; case 'G': return 3; default: return 0; }
}

/* :80-83 -- 32-bit shift-in; for k > 16 only the last 16 bases survive */
static uint32_t kmer_hash(const char* s, uint64_t at, uint32_t k, uint32_t len) {
    uint32_t h = 0;
    for (uint32_t x = 0; x < k; ++x) h = (h << 2) | code_of(s, at + x, len);
    return h;
}

typedef struct { uint32_t hash, pos; uint8_t flag; } tuple_t;

/* :106-120 -- leftmost strict minimum below UINT_MAX, else the zero tuple */
static tuple_t window_min(const uint32_t* h, uint64_t a, uint64_t b, int is_fwd) {
    tuple_t best = {0, 0, 0};
    uint32_t mn = 0xFFFFFFFFu;
    for (uint64_t x = a; x <= b; ++x)
        if (h[x] < mn) { mn = h[x]; best.hash = h[x]; best.pos = (uint32_t)(x + 1); best.flag = is_fwd ? 1 : 0; }
    return best;
}

int64_t oracle_minimize(const char* seq, uint32_t len, uint32_t k, uint32_t w, int is_fwd,
                        uint32_t* hash, uint32_t* pos, uint8_t* flag, uint64_t cap) {
    if (len < k || w == 0) return 0;                       /* :140 */
    const uint64_t n = (uint64_t)len - k + 1;              /* k-mers inside the sequence */
    const uint64_t n_ext = n > (uint64_t)w - 1 ? n : (uint64_t)w - 1;   /* begin section may overrun */
    const uint64_t full = n >= w ? n - w + 1 : 0;
    const uint64_t tail = n < (uint64_t)w - 1 ? n : (uint64_t)w - 1;     /* :198 break */
    const uint64_t total = (uint64_t)(w - 1) + full + tail;
    if (cap < total) return (int64_t)total;

    uint32_t* h = (uint32_t*)malloc((n_ext ? n_ext : 1) * sizeof(uint32_t));
    if (!h) return -3;
    for (uint64_t x = 0; x < n_ext; ++x) h[x] = kmer_hash(seq, x, k, len);

    uint64_t o = 0;
    tuple_t m;
    for (uint64_t s = 1; s + 1 <= w; ++s) {                /* begin end-minimizers :146-170 */
        m = window_min(h, 0, s - 1, is_fwd);
        hash[o] = m.hash; pos[o] = m.pos; flag[o] = m.flag; ++o;
    }
    for (uint64_t e = (uint64_t)w - 1; e < n; ++e) {       /* full windows :173-194 */
        m = window_min(h, e + 1 - w, e, is_fwd);
        hash[o] = m.hash; pos[o] = m.pos; flag[o] = m.flag; ++o;
    }
    for (uint64_t s = 1; s <= tail; ++s) {                 /* end end-minimizers :197-222 */
        m = window_min(h, n - s, n - 1, is_fwd);
        hash[o] = m.hash; pos[o] = m.pos; flag[o] = m.flag; ++o;
    }
    free(h);
    return (int64_t)o;
}

/* Packed batch: tuples of sequence i at out_off[i] .. out_off[i+1]. Returns the tuple total, -2 when cap is too
 * small (out_off is still filled), -3 out of memory. The sequence buffer is read only inside each sequence. */
int64_t oracle_minimize_batch(uint64_t n, const char* buf, const uint64_t* off, uint32_t k, uint32_t w, int is_fwd,
                              uint32_t* hash, uint32_t* pos, uint8_t* flag, uint64_t cap, uint64_t* out_off) {
    uint64_t at = 0;
    int over = 0;
    out_off[0] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t len = (uint32_t)(off[i + 1] - off[i]);
        const int64_t cnt = oracle_minimize(buf + off[i], len, k, w, is_fwd, 0, 0, 0, 0);
        if (cnt < 0) return cnt;
        if (at + (uint64_t)cnt > cap) over = 1;
        if (!over && cnt) {
            const int64_t c2 = oracle_minimize(buf + off[i], len, k, w, is_fwd, hash + at, pos + at, flag + at, (uint64_t)cnt);
            if (c2 != cnt) return -3;
        }
        at += (uint64_t)cnt;
        out_off[i + 1] = at;
    }
    return over ? -2 : (int64_t)at;
}
