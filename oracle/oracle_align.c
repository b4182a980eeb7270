/* oracle/oracle_align.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of team::Align (reference team_alignment/team_alignment.cpp:49-350)
 * with two rolling score rows and one byte of trace per cell instead of the
 * reference's 8-byte cells, so 10k x 10k pairs fit in 100 MB.
 */
#include "oracle.h"
#include <limits.h>
#include <stdlib.h>
#include <string.h>

enum { P_DIAG = 0, P_LEFT = 1, P_UP = 2, POSITIVE = 4 };

/* wrap-around int32 add: the reference's `int` overflow is UB, we pick 2's complement */
static inline int32_t add32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t mul32(uint32_t a, int32_t b) { return (int32_t)(a * (uint32_t)b); }

/* team_alignment.cpp:20-28 */
static inline int32_t subst(char a, char b, int m, int x) { return a == b ? m : x; }
static inline int32_t indel(char c, int g) { return c == '-' ? 0 : g; }

static uint64_t put_run(char* dst, uint64_t at, uint64_t cap, uint64_t count, char op, int* overflow) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + count % 10); count /= 10; } while (count);
    if (at + (uint64_t)n + 1 > cap) { *overflow = 1; return at + (uint64_t)n + 1; }
    while (n) dst[at++] = tmp[--n];
    dst[at++] = op;
    return at;
}

int oracle_align(const char* q, uint32_t Q, const char* t, uint32_t T, int type,
                 int match, int mismatch, int gap, int want_cigar,
                 int32_t* score, uint32_t* target_begin,
                 char* cigar_buf, uint64_t cigar_cap, uint64_t* cigar_len) {
    if (type < 0 || type > 2) return -1;                 /* :62-74 */
    const int32_t init = (type == 0) ? gap : 0;
    const uint64_t W = (uint64_t)T + 1;

    int32_t* prev = (int32_t*)malloc(W * sizeof(int32_t));
    int32_t* cur = (int32_t*)malloc(W * sizeof(int32_t));
    int32_t* lastcol = (int32_t*)malloc(((uint64_t)Q + 1) * sizeof(int32_t));
    uint8_t* tr = NULL;
    if (want_cigar && Q && T) tr = (uint8_t*)malloc((uint64_t)Q * T);
    if (!prev || !cur || !lastcol || (want_cigar && Q && T && !tr)) {
        free(prev); free(cur); free(lastcol); free(tr);
        return -3;
    }

    for (uint32_t j = 0; j <= T; ++j) prev[j] = mul32(j, init);   /* row 0, :89-92 */
    lastcol[0] = prev[T];

    int32_t best = INT_MIN;                                         /* :96 */
    uint32_t gi = 0, gj = 0;
    for (uint32_t i = 1; i <= Q; ++i) {
        cur[0] = mul32(i, init);                                    /* column 0, :83-86 */
        const char qc = q[i - 1];
        const int32_t gu = indel(qc, gap);
        uint8_t* trow = tr ? tr + (uint64_t)(i - 1) * T : NULL;
        for (uint32_t j = 1; j <= T; ++j) {
            const char tc = t[j - 1];
            int32_t h = add32(prev[j - 1], subst(qc, tc, match, mismatch));   /* :104 */
            uint8_t p = P_DIAG;
            const int32_t l = add32(cur[j - 1], indel(tc, gap));               /* :105 */
            const int32_t u = add32(prev[j], gu);                               /* :106 */
            if (l > h) { h = l; p = P_LEFT; }                                   /* :109-114 */
            if (u > h) { h = u; p = P_UP; }
            if (type == 1) {
                if (h < 0) h = 0;                                               /* :185 */
                if (h > best) { best = h; gi = i; gj = j; }                     /* :186-192 */
            }
            cur[j] = h;
            if (trow) trow[j - 1] = (uint8_t)(p | (h > 0 ? POSITIVE : 0));
        }
        lastcol[i] = cur[T];
        int32_t* sw = prev; prev = cur; cur = sw;
    }
    /* prev now holds row Q */
    int32_t result;
    if (type == 0) {                                                /* :117-121 */
        gi = Q; gj = T; result = prev[T];
        if (target_begin) *target_begin = 0;
    } else if (type == 1) {
        /* untouched (gi,gj) = (0,0) when the matrix has no inner cell */
        result = (Q && T) ? best : 0;
        if (target_begin) *target_begin = gj + 1;                   /* :197-199 */
    } else {
        int32_t mx = INT_MIN;                                       /* :265-278 */
        for (uint32_t i = 0; i <= Q; ++i)
            if (lastcol[i] > mx) { mx = lastcol[i]; gi = i; gj = T; }
        for (uint32_t j = 0; j <= T; ++j)
            if (prev[j] > mx) { mx = prev[j]; gi = Q; gj = j; }
        result = mx;
        if (target_begin) *target_begin = 0;                        /* :283-285 */
    }
    *score = result;

    int rc = 0;
    if (want_cigar) {
        /* walk backwards collecting ops (:123-138, :202-217, :287-302) */
        uint64_t cap_ops = (uint64_t)Q + T + 1, n_ops = 0;
        char* ops = (char*)malloc(cap_ops);
        if (!ops) { rc = -3; goto done; }
        uint32_t i = gi, j = gj;
        if (type == 1) {
            /* while (cost > 0): border cells cost 0 */
            while (i > 0 && j > 0) {
                const uint8_t c = tr[(uint64_t)(i - 1) * T + (j - 1)];
                if (!(c & POSITIVE)) break;
                const uint8_t p = c & 3;
                if (p == P_DIAG) { ops[n_ops++] = 'M'; --i; --j; }
                else if (p == P_LEFT) { ops[n_ops++] = 'I'; --j; }
                else { ops[n_ops++] = 'D'; --i; }
            }
        } else {
            while (i > 0 || j > 0) {
                /* border parents: column 0 -> UP ('D'), row 0 -> LEFT ('I'), :83-92 */
                uint8_t p;
                if (i == 0) p = P_LEFT;
                else if (j == 0) p = P_UP;
                else p = tr[(uint64_t)(i - 1) * T + (j - 1)] & 3;
                if (p == P_DIAG) { ops[n_ops++] = 'M'; --i; --j; }
                else if (p == P_LEFT) { ops[n_ops++] = 'I'; --j; }
                else { ops[n_ops++] = 'D'; --i; }
            }
        }
        /* RLE in forward order (:145-160); semiGlobal tail pad (:306-315) is one extra run source */
        uint64_t out = 0; int overflow = 0;
        char pad_op = 0; uint64_t pad_n = 0;
        if (type == 2 && (gj != T || gi != Q)) {
            if (gi == Q) { pad_op = 'I'; pad_n = T - gj; }
            else if (gj == T) { pad_op = 'D'; pad_n = Q - gi; }
        }
        if (n_ops == 0 && pad_n == 0) {
            /* result[0] of an empty std::string is its NUL terminator: "1\0" (:145-159) */
            if (cigar_cap < 2) overflow = 1; else { cigar_buf[0] = '1'; cigar_buf[1] = '\0'; }
            out = 2;
        } else {
            char run_op = 0; uint64_t run_n = 0;
            for (uint64_t k = n_ops; k-- > 0;) {
                if (ops[k] == run_op) ++run_n;
                else { if (run_n) out = put_run(cigar_buf, out, cigar_cap, run_n, run_op, &overflow);
                       run_op = ops[k]; run_n = 1; }
            }
            if (pad_n) {
                if (pad_op == run_op) run_n += pad_n;
                else { if (run_n) out = put_run(cigar_buf, out, cigar_cap, run_n, run_op, &overflow);
                       run_op = pad_op; run_n = pad_n; }
            }
            if (run_n) out = put_run(cigar_buf, out, cigar_cap, run_n, run_op, &overflow);
        }
        free(ops);
        if (cigar_len) *cigar_len = out;
        if (overflow) rc = -2;
    } else if (cigar_len) {
        *cigar_len = 0;
    }
done:
    free(prev); free(cur); free(lastcol); free(tr);
    return rc;
}

/* Packed batch, single thread (bench.py's CPU arm runs one call per host thread). */
int64_t oracle_align_batch(uint64_t n, const char* qbuf, const uint64_t* qoff, const char* tbuf,
                           const uint64_t* toff, int type, int match, int mismatch, int gap, int want_cigar,
                           int32_t* score, uint32_t* target_begin, uint64_t* cigar_bytes) {
    uint64_t bytes = 0, cap = 0;
    char* buf = NULL;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t ql = qoff[i + 1] - qoff[i], tl = toff[i + 1] - toff[i];
        if (want_cigar && 2 * (ql + tl) + 16 > cap) {
            cap = 2 * (ql + tl) + 16;
            free(buf);
            buf = (char*)malloc(cap);
            if (!buf) return -3;
        }
        uint64_t len = 0;
        int rc = oracle_align(qbuf + qoff[i], (uint32_t)ql, tbuf + toff[i], (uint32_t)tl, type, match, mismatch,
                              gap, want_cigar, &score[i], &target_begin[i], buf, cap, &len);
        if (rc) { free(buf); return rc; }
        bytes += len;
    }
    free(buf);
    if (cigar_bytes) *cigar_bytes = bytes;
    return (int64_t)n;
}

/* Same batch with every output kept: CIGAR texts are written back to back into cigar_buf and cigar_off[i] /
 * cigar_off[i+1] bracket pair i's bytes (relative to this call's first byte). Returns n, -2 when cigar_cap is
 * too small, or oracle_align's error code. Re-entrant (tests run one call per host thread on disjoint slices). */
int64_t oracle_align_batch_cigar(uint64_t n, const char* qbuf, const uint64_t* qoff, const char* tbuf,
                                 const uint64_t* toff, int type, int match, int mismatch, int gap,
                                 int32_t* score, uint32_t* target_begin, char* cigar_buf, uint64_t cigar_cap,
                                 uint64_t* cigar_off) {
    uint64_t at = 0;
    cigar_off[0] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t ql = qoff[i + 1] - qoff[i], tl = toff[i + 1] - toff[i];
        uint64_t len = 0;
        int rc = oracle_align(qbuf + qoff[i], (uint32_t)ql, tbuf + toff[i], (uint32_t)tl, type, match, mismatch,
                              gap, 1, &score[i], &target_begin[i], cigar_buf + at, cigar_cap - at, &len);
        if (rc) return rc;
        at += len;
        cigar_off[i + 1] = at;
    }
    return (int64_t)n;
}
