/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference hot path, used solely as the
 * parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg. Nothing under bioinfo1_b200/ may include, link or call this.
 *
 * Pinned against: the golden table in tests/golden/ (generated from the
 * unmodified reference by tools/make_golden.py) and, when oracle/_ref/libref.so
 * is present, live differential tests against the reference itself.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* type: 0 global, 1 local, 2 semiGlobal (team_alignment.hpp:8-12).
 * Returns 0, -1 for an unknown type (the reference throws invalid_argument,
 * team_alignment.cpp:73), -2 if cigar_cap is too small, -3 out of memory. */
int oracle_align(const char* q, uint32_t ql, const char* t, uint32_t tl, int type,
                 int match, int mismatch, int gap, int want_cigar,
                 int32_t* score, uint32_t* target_begin,
                 char* cigar_buf, uint64_t cigar_cap, uint64_t* cigar_len);

int64_t oracle_align_batch(uint64_t n, const char* qbuf, const uint64_t* qoff, const char* tbuf,
                           const uint64_t* toff, int type, int match, int mismatch, int gap, int want_cigar,
                           int32_t* score, uint32_t* target_begin, uint64_t* cigar_bytes);

/* every output kept: texts back to back in cigar_buf, bracketed by cigar_off[i], cigar_off[i+1] (n+1 entries) */
int64_t oracle_align_batch_cigar(uint64_t n, const char* qbuf, const uint64_t* qoff, const char* tbuf,
                                 const uint64_t* toff, int type, int match, int mismatch, int gap,
                                 int32_t* score, uint32_t* target_begin, char* cigar_buf, uint64_t cigar_cap,
                                 uint64_t* cigar_off);

/* Returns the tuple count; fills the arrays only when cap >= count.
 * Bytes at index >= len are treated as code 0 (what the reference reads from a
 * NUL-padded buffer, team_minimizers.cpp:146-152). */
int64_t oracle_minimize(const char* seq, uint32_t len, uint32_t k, uint32_t w, int is_fwd,
                        uint32_t* hash, uint32_t* pos, uint8_t* flag, uint64_t cap);

int64_t oracle_minimize_batch(uint64_t n, const char* buf, const uint64_t* off, uint32_t k, uint32_t w, int is_fwd,
                              uint32_t* hash, uint32_t* pos, uint8_t* flag, uint64_t cap, uint64_t* out_off);

#ifdef __cplusplus
}
#endif
#endif
