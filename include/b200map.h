/* include/b200map.h -- the drop-in boundary: a plain C ABI over the B200 (sm_100a) kernels.
 *
 * Everything the reference's hot path does on the CPU is reachable through these entry
 * points; there is no CPU fallback behind them (a missing GPU is an error code, never a
 * silent slow path).
 *
 * What each entry point replaces in the reference (AnamarijaKic/bioinfo1):
 *   b200_align_*      n x team::Align(query, query_len, target, target_len, type, match,
 *                     mismatch, gap, cigar*, target_begin*)
 *                       decl  team_alignment/team_alignment.hpp:14-23
 *                       impl  team_alignment/team_alignment.cpp:49-350
 *                       calls team_mapper.cpp:666-678 and :755-767 (one per mapped read)
 *   b200_minimize_*   n x team::KMER(is_fwd).Minimize(sequence, len, kmer_len, window_len)
 *                       decl  team_minimizers/team_minimizers.hpp:20-23
 *                       impl  team_minimizers/team_minimizers.cpp:122-225
 *                       calls team_mapper.cpp:417-427 (index), :606 and :713 (per read)
 *
 * Conventions
 *   - plain pointers and sizes only; no STL, no exceptions, no globals cross this line;
 *   - every function returns B200_OK (0) or a negative B200_E_* code, with a thread-local
 *     message available from b200_last_error();
 *   - "packed" = all sequences concatenated in one byte buffer `buf` with an offsets array
 *     `off` of n+1 entries (sequence i is buf[off[i] .. off[i+1]));
 *   - CIGAR bytes may contain NUL (the reference's empty-path CIGAR is the 2-byte string
 *     "1\0", team_alignment.cpp:145-159): lengths always come from `cigar_off`, never strlen;
 *   - AlignmentType values are the reference enum's (team_alignment.hpp:8-12).
 */
#ifndef B200MAP_H
#define B200MAP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    B200_OK = 0,
    B200_E_TYPE = -1,  /* unknown AlignmentType (reference: std::invalid_argument, team_alignment.cpp:73) */
    B200_E_NOMEM = -2, /* host or device allocation failed */
    B200_E_CUDA = -3,  /* CUDA runtime error (message has the detail) */
    B200_E_CAP = -4,   /* an output buffer is too small; nothing partial is reported as success */
    B200_E_ARG = -5,   /* null/inconsistent argument */
    B200_E_NOGPU = -6  /* no usable sm_100 device */
};

enum { B200_GLOBAL = 0, B200_LOCAL = 1, B200_SEMIGLOBAL = 2 };

typedef struct b200_ctx b200_ctx;               /* one per (host thread, device): streams + workspaces */
typedef struct b200_align_plan b200_align_plan; /* shape-only schedule of one alignment batch */
typedef struct b200_min_plan b200_min_plan;     /* shape-only schedule of one minimizer batch */

const char* b200_last_error(void);
int b200_device_count(void);
int b200_version(void); /* ABI version, currently 1 */

int b200_ctx_create(int device, b200_ctx** out);
void b200_ctx_destroy(b200_ctx* ctx);
/* Tunables (all optional): "dir_budget_bytes" (HBM given to 2-bit direction storage; a batch that needs more
 * is cut into waves of half of it, two in flight), "force_generic" (1 = route every pair through the int32
 * byte-compare kernel), "long16" (0 = long pairs stay on the int32 stripe kernel), "overlap_waves" (0 = waves
 * run one after the other on the caller's stream), "concurrent_walk" (0 = long pairs are walked after their wave's
 * fill instead of next to it), "chunk_pairs" (host-API pipeline chunk), "taper_tail" (host pipeline of uniform short
 * batches: 1 = shrinking tail of waves, 2 = last wave cut in two), "stream_fill" (1 = one persistent fill launch fed
 * by an upload watermark), "host_pack" (pointer-array entry point: 1 = the gather pass packs to 2 bits on the host),
 * "subst_lds" (substitution term from a shared-memory table instead of PRMT: bit 0 short-pair kernel, bit 1 long-pair
 * kernel), "fill_pipe" (0 = the short-pair kernel without software-pipelined columns), "profile" (1 = bracket kernels
 * with events: the "*_ns" counters), "reset_counters". Defaults are what measured best; the
 * environment variables B200_SUBST_LDS, B200_FILL_PIPE, B200_TAPER_TAIL, B200_STREAM_FILL, B200_HOST_PACK and
 * B200_CONCURRENT_WALK preset those six for every context a process creates (tuning runs). Returns B200_E_ARG for an
 * unknown key. */
int b200_ctx_set_option(b200_ctx* ctx, const char* key, int64_t value);
/* Counters since the context was created (or the last "reset_counters"): "kernel_launches", "h2d_bytes",
 * "d2h_bytes", "alu_slots_per_cell_pair_x10" (of the short-pair kernel variant in use); with "profile" on also
 * "fill_ns", "walk_ns", "emit_ns", "other_ns" and "*_launches". */
int64_t b200_ctx_get_counter(b200_ctx* ctx, const char* key);

/* ------------------------------------------------------------------ alignment ---- */

/* Reference-shaped batch: arrays of n independent Align() argument tuples sharing type and
 * scores. `target_begin`, `cigar_buf`/`cigar_off` may be NULL (score only, like passing
 * nullptr to the reference). `cigar_off` has n+1 entries. Uses (and lazily creates) a
 * per-thread default context for `device`. */
int b200_align_batch(int device, size_t n,
                     const char* const* query, const uint32_t* query_len,
                     const char* const* target, const uint32_t* target_len,
                     int type, int match, int mismatch, int gap,
                     int32_t* score, uint32_t* target_begin,
                     char* cigar_buf, uint64_t* cigar_off, uint64_t cigar_cap);

/* Same, with packed HOST buffers (pinned or pageable). Work is pipelined in chunks so that
 * the host->device copy of chunk c+1 overlaps the kernels of chunk c. */
int b200_align_batch_packed(b200_ctx* ctx, size_t n,
                            const char* q_buf, const uint64_t* q_off,
                            const char* t_buf, const uint64_t* t_off,
                            int type, int match, int mismatch, int gap,
                            int32_t* score, uint32_t* target_begin,
                            char* cigar_buf, uint64_t* cigar_off, uint64_t cigar_cap);

/* Device-resident path. The plan is built from lengths only (host offsets), owns the
 * per-pair descriptors and the wave schedule, and can be run any number of times on
 * sequences already in HBM. All `d_*` pointers are device memory; `stream` is a
 * cudaStream_t used as given (0 = the CUDA default stream). `d_target_begin`, `d_cigar`,
 * `d_cigar_off` may be NULL when the plan was created with want_cigar = 0.
 * On return the work is enqueued and, if want_cigar, the total CIGAR byte count has been
 * checked against cigar_cap (that check synchronises `stream` once).
 * `d_q_buf` / `d_t_buf`: the 2-bit packing pass reads whole aligned 32-bit words, so both buffers
 * must be readable from the 4-byte boundary at or below their first byte to the 4-byte boundary at
 * or above their last byte (any cudaMalloc'ed or framework-allocated buffer is; a sub-allocation
 * cut at an odd byte inside a larger buffer is too, as long as the neighbouring bytes are mapped). */
int b200_align_plan_create(b200_ctx* ctx, size_t n, const uint64_t* q_off, const uint64_t* t_off,
                           int type, int match, int mismatch, int gap, int want_cigar,
                           b200_align_plan** out);
void b200_align_plan_destroy(b200_align_plan* plan);
uint64_t b200_align_plan_cells(const b200_align_plan* plan);       /* sum of Q*T */
uint64_t b200_align_plan_cigar_bound(const b200_align_plan* plan); /* worst-case CIGAR bytes */
int b200_align_plan_run(b200_align_plan* plan, const char* d_q_buf, const char* d_t_buf,
                        int32_t* d_score, uint32_t* d_target_begin,
                        char* d_cigar, uint64_t* d_cigar_off, uint64_t cigar_cap, void* stream);

/* ------------------------------------------------------------------ minimizers ---- */

/* Number of tuples Minimize() returns for a sequence of length len
 * (team_minimizers.cpp:140-222): 0 if len < k or w == 0, else
 * (w-1) + max(0, n-w+1) + min(w-1, n) with n = len-k+1. */
uint64_t b200_minimize_count(uint32_t len, uint32_t k, uint32_t w);

/* Reference-shaped batch: n x KMER(is_fwd[i]).Minimize(seq[i], len[i], k, w). Outputs are
 * structure-of-arrays, tuple i of sequence s at index out_off[s] + i. `out_off` (n+1
 * entries) is written by the call; `cap` is the capacity of hash/pos/flag in tuples.
 * Bytes past a sequence's end count as code 0 (what the reference reads from a NUL-padded
 * buffer when len < k+w-2, team_minimizers.cpp:146-152). */
int b200_minimize_batch(int device, size_t n, const char* const* seq, const uint32_t* len,
                        uint32_t k, uint32_t w, const uint8_t* is_fwd,
                        uint32_t* hash, uint32_t* pos, uint8_t* flag, uint64_t* out_off, uint64_t cap);

int b200_minimize_batch_packed(b200_ctx* ctx, size_t n, const char* buf, const uint64_t* off,
                               uint32_t k, uint32_t w, const uint8_t* is_fwd,
                               uint32_t* hash, uint32_t* pos, uint8_t* flag, uint64_t* out_off,
                               uint64_t cap);

int b200_min_plan_create(b200_ctx* ctx, size_t n, const uint64_t* off, uint32_t k, uint32_t w,
                         const uint8_t* is_fwd, b200_min_plan** out);
void b200_min_plan_destroy(b200_min_plan* plan);
uint64_t b200_min_plan_tuples(const b200_min_plan* plan);       /* total tuples = out_off[n] */
const uint64_t* b200_min_plan_out_off(const b200_min_plan* plan); /* host array, n+1 entries */
int b200_min_plan_run(b200_min_plan* plan, const char* d_buf, uint32_t* d_hash, uint32_t* d_pos,
                      uint8_t* d_flag, void* stream);

/* ------------------------------------------------------------------ mapper ("next" rows) ---- */

/* The reference's minimizer index (team_mapper.cpp:412-477): Minimize(reference) and
 * Minimize(reverse complement), per strand the distinct (hash, position) pairs, optionally without
 * the `f`-fraction most frequent forward hashes. Built on the GPU (MinimizeBatch + radix sort +
 * unique); the reference and its reverse complement stay resident in HBM for the Align step.
 * With f > 0 ties at the ban cut are broken by (count desc, hash asc): the reference's own order
 * there is implementation-defined (std::sort over unordered_map iteration order, :437-450). */
typedef struct b200_index b200_index;
int b200_index_build(b200_ctx* ctx, const char* ref, uint64_t ref_len, uint32_t k, uint32_t w, double f,
                     b200_index** out);
void b200_index_destroy(b200_index* index);
/* what[0..5] = tuples fwd, tuples rev, distinct (hash,pos) fwd, distinct rev, indexed fwd, indexed rev */
int b200_index_stats(const b200_index* index, uint64_t what[6]);

typedef struct b200_mapping {
    uint32_t mapped;     /* 0: no chain found (the reference prints nothing for the read) */
    uint32_t strand_fwd; /* 1 '+', 0 '-' */
    uint32_t q_begin, q_end; /* inclusive, 0-based, in the read (team_mapper.cpp:653-654) */
    uint32_t t_begin, t_end; /* inclusive, 0-based, in the chosen strand's coordinates (:655-656) */
    int32_t score;
    uint32_t target_begin;
} b200_mapping;

/* The per-read loop of the reference mapper (team_mapper.cpp:596-700 for FASTA input, :710-790 for
 * FASTQ input) for a packed batch of reads: Minimize -> remove_duplicates -> seed lookup -> FindLIS
 * on both strands -> region -> Align. `fastq_semantics` selects the seed-lookup flavour (:716-729
 * when non-zero; the FASTA flavour :627-638 consults the reverse index only for hashes that are also
 * in the forward index). `cigar_off` (n+1) / `cigar_buf` may be NULL when want_cigar is 0; unmapped
 * reads get an empty CIGAR slice. */
int b200_map_batch(b200_ctx* ctx, const b200_index* index, size_t n, const char* reads_buf,
                   const uint64_t* reads_off, int fastq_semantics, int type, int match, int mismatch, int gap,
                   int want_cigar, b200_mapping* out, char* cigar_buf, uint64_t* cigar_off, uint64_t cigar_cap);

#ifdef __cplusplus
}
#endif
#endif /* B200MAP_H */
