// internal.hpp -- host-side definitions shared by the translation units behind include/b200map.h:
//   ctx.cu           contexts, options, counters, error text
//   align_plan.cu    shape-only planning of an alignment batch (classes, waves, descriptors)
//   align_run.cu     kernel launches of a planned batch (wave loop, emit, repair pass)
//   align_host.cu    host-buffer entry points (upload pipeline, pointer-array gather)
//   minimize_api.cu  minimizer plans and entry points
//   mapper_api.cu    index build and the per-read mapping batch
// Nothing here is part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/b200map.h"
#include "common.cuh"

using b200::PairDesc;
using b200::Scores;

// ------------------------------------------------------------------ errors ----
int b200_fail(int code, const std::string& msg);   // records the thread-local message, returns `code`
#define fail b200_fail
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            const int code__ = (e__ == cudaErrorMemoryAllocation) ? B200_E_NOMEM : B200_E_CUDA; \
            return fail(code__, std::string(#call) + ": " + cudaGetErrorString(e__));         \
        }                                                                                     \
    } while (0)
#define TRY(expr)                \
    do {                         \
        int rc__ = (expr);       \
        if (rc__ != B200_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------ context ----
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return B200_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return fail(B200_E_NOMEM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed");
        }
        cap = want;
        return B200_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostBuf {  // pinned staging
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return B200_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        if (cudaMallocHost(&p, bytes + bytes / 8 + 256) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return fail(B200_E_NOMEM, "cudaMallocHost failed");
        }
        cap = bytes + bytes / 8 + 256;
        return B200_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Per-wave workspaces. Two slots: consecutive waves of a run alternate between them (and between two
// streams), so the latency-bound traceback walk of wave k overlaps the ALU-bound fill of wave k+1.
struct WaveSlot {
    DevBuf dirs, bnd, bnd_short, progress, stripe_res, counter, fix_work, pair_state;
    cudaEvent_t done = nullptr;
    // concurrent walk of the wave that is being filled (long class): its own stream, fork / join events
    cudaStream_t walk_stream = nullptr;
    cudaEvent_t pre_event = nullptr, walk_event = nullptr;
    bool walk_inflight = false;
};

struct b200_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    // workspaces shared by every plan run on this context (one run at a time per context)
    WaveSlot slot[2];
    cudaStream_t aux_stream = nullptr;      // second wave stream of a run
    cudaStream_t emit_stream = nullptr;     // host path: scan / emit / download of finished waves while later ones run
    cudaStream_t pack_stream = nullptr;     // host path: 2-bit packing of a wave as soon as its bytes have landed
    cudaStream_t fill_stream = nullptr;     // host path, streaming mode: the one persistent fill launch of a run
    cudaEvent_t fill_event = nullptr;
    DevBuf stream_state;                    // streaming mode: watermark + per-wave completion counters
    // Host pipeline of uniform short batches: 0 = one fill launch per wave (default), 1 = ONE persistent fill launch
    // fed by an upload watermark (ShortStream). Measured on config 2: 7.13 against 7.20 ms per step -- the persistent
    // launch gives up a quarter of the fill's occupancy so that the small kernels can run next to it, and what is left
    // after the last byte of the upload is one group's latency either way; off by default, kept tested.
    int64_t stream_fill = 0;
    std::vector<cudaEvent_t> pack_done;     // one event per wave of the current run
    std::vector<cudaEvent_t> wave_done;     // one event per wave of the current run
    // B200_TRACE=2: device timeline of a run (events with timing, printed relative to the first)
    struct TlMark { std::string what; cudaEvent_t e; };
    std::vector<TlMark> timeline;
    cudaEvent_t fork_event = nullptr;
    int64_t overlap_waves = 1;              // 0 = all waves on the caller's stream, one after the other
    int64_t concurrent_walk = 1;            // 0 = a wave's pairs are walked after its fill kernel has finished
    DevBuf qpk, tpk, end_i, end_j, runs, n_runs, cigar_len, scan_tmp, flags, total;
    DevBuf wave_flagged;                    // per wave of a run: pairs planned for a 2-bit kernel that are not pure ACGT
    HostBuf h_small;                        // pinned landing zone of the small read-backs of a run
    // staging for the host-buffer entry points
    DevBuf d_q, d_t, d_score, d_tb, d_cigar, d_cigar_off, d_seq, d_hash, d_pos, d_flag;
    HostBuf h_q, h_t, h_off;
    HostBuf h_out_small, h_out_cigar;       // pointer-array entry point: pinned landing zone of the results
    HostBuf h_qpk, h_tpk, h_flags, h_wave_cnt;   // ... and of the 2-bit words / flags its gather pass packs on the host
    int64_t host_pack = 0;                  // 1 = the pointer-array gather of a uniform short batch packs to 2 bits on the host
    // options
    int64_t dir_budget_bytes = 48ll << 30;
    int64_t force_generic = 0;
    int64_t long16 = 1;                     // 0 = long pairs stay on the int32 kernel (align_fill_long.cuh)
    int64_t fill_pipe = 1;                  // 1 = software-pipelined columns in the 2-bit fill kernels (K1)
    int64_t subst_lds = 2;                  // substitution term from the shared-memory table instead of PRMT: bit 0 = K1, bit 1 = K3
    int64_t chunk_pairs = 0;
    // host pipeline of uniform batches: end with a few shrinking waves. Measured on config 2: 7.2 ms per step against
    // 6.9 ms with equal waves (a small wave still takes 0.35-0.6 ms, and the extra waves queue behind each other); off.
    int64_t taper_tail = 0;
    b200_align_plan* host_plan = nullptr;   // recycled by the host-buffer entry points
    b200_align_plan* map_plan = nullptr;    // recycled by b200_map_batch (its device buffers keep their capacity)
    b200_min_plan* map_min_plan = nullptr;
    DevBuf map_buf[27];                     // b200_map_batch scratch (grow-only; cudaMalloc/cudaFree per call cost more than the kernels)
    cudaStream_t copy_stream = nullptr;     // uploads of the host-buffer entry points (overlap with kernels)
    std::vector<cudaEvent_t> copy_events;
    int64_t profile = 0;   // 1 = bracket kernels with CUDA events (adds a sync per run)
    // counters
    int64_t kernel_launches = 0, h2d_bytes = 0, d2h_bytes = 0;
    // per-kind device time, filled only when profile == 1: 0 fill, 1 walk, 2 emit, 3 other
    double kind_us[4] = {0, 0, 0, 0};
    int64_t kind_launches[4] = {0, 0, 0, 0};
    struct Span { int kind; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
};

void prof_begin(b200_ctx* c, cudaStream_t st, int kind);   // event brackets around a launch (no-ops unless ctx->profile)
void prof_end(b200_ctx* c, cudaStream_t st);
void prof_collect(b200_ctx* c, cudaStream_t st);
void tl_mark(b200_ctx* c, cudaStream_t st, const std::string& what);   // B200_TRACE=2 device timeline
void tl_dump(b200_ctx* c);
int set_device(const b200_ctx* c);
int default_ctx(int device, b200_ctx** out);   // per-thread default contexts of the reference-shaped entry points
void ctx_sync_all_streams(b200_ctx* c);        // error paths: nothing of a failed run may still be in flight

#include <chrono>
#include <cstdio>
#include <cstdlib>
struct PhaseTrace {   // B200_TRACE=1 prints host-side phase timings of the host-buffer entry points
    bool on;
    std::chrono::steady_clock::time_point t0;
    std::string line;
    PhaseTrace() : on(std::getenv("B200_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        line += std::string(what) + "=" + std::to_string(std::chrono::duration<double, std::milli>(t1 - t0).count()) + "ms ";
        t0 = t1;
    }
    ~PhaseTrace() { if (on) std::fprintf(stderr, "[b200 trace] %s\n", line.c_str()); }
};

// ------------------------------------------------------------------ align plan ----
// A wave is a slice of the work order whose direction matrices fit the HBM budget together and
// that is served by one fill kernel class.
struct Wave {
    uint32_t klass;
    uint32_t first, count;   // range in the work order
    uint32_t first_group;    // short class: index of the first ShortGroup
    uint64_t dir_words;
};

// Uniform plans: the waves are equal chunks of `u_groups_per_wave` 64-pair groups over groups [0, first_group),
// followed by up to 16 explicitly placed waves (the host pipeline's shrinking tail).
struct UniformTail {
    uint32_t first_group;   // first group of the tail (= number of groups when there is no tail)
    uint32_t n;             // waves in the tail
    uint32_t start[16];     // first group of each tail wave
};

struct b200_align_plan {
    b200_ctx* ctx = nullptr;
    size_t n = 0;
    int type = 0;
    Scores sc{};
    bool want_cigar = false;
    bool long16 = false;               // long class runs on the packed int16x2 kernel (align_fill_long16.cuh)
    uint64_t cells = 0, cigar_bound = 0, run_slots = 0, q_bytes = 0, t_bytes = 0, qpk_words = 0, tpk_words = 0;
    uint32_t max_T = 0, max_Q = 0, max_T_short = 0, max_Q_short = 0;
    size_t n_short = 0, n_long = 0;   // the work order is [short..., long..., generic...]
    std::vector<Wave> waves;
    std::vector<PairDesc> h_pairs;     // kept for the non-ACGT fallback (content is only known at run time)
    std::vector<uint32_t> h_order;
    bool patched = false;              // d_pairs currently holds run-specific fallback descriptors
    bool host_packed = false;          // the 2-bit copies, flags and per-wave flag counts of this run come from the host
    std::vector<cudaEvent_t> wave_events;   // optional, per wave: "this wave's sequence bytes are resident" (host pipeline)
    bool uniform = false;              // every pair has the same (Q,T): descriptors were built on the device
    uint32_t uQ = 0, uT = 0;
    uint64_t u_groups_per_wave = 1;
    UniformTail u_tail{0, 0, {0}};
    uint64_t u_qbase = 0, u_tbase = 0;
    DevBuf d_pairs, d_work, d_groups, d_task_off, d_bnd_off;
    uint64_t max_long_bnd_words = 0;   // boundary rows of the largest long wave
    uint32_t max_long_tasks = 0;

    void reset() {
        max_long_bnd_words = 0; max_long_tasks = 0;
        n = 0; cells = cigar_bound = run_slots = q_bytes = t_bytes = qpk_words = tpk_words = 0;
        max_T = max_Q = max_T_short = max_Q_short = 0; n_short = n_long = 0;
        waves.clear(); h_pairs.clear(); h_order.clear(); patched = false; uniform = false; wave_events.clear();
        host_packed = false;
    }
};

bool short_scores_ok(const b200::Scores& sc, int type);
uint64_t generic_dir_words(uint32_t Q, uint32_t T);
uint64_t wave_budget_words(const b200_ctx* ctx);
void materialize_uniform_host(b200_align_plan* p);
int plan_build(b200_align_plan* p, b200_ctx* ctx, size_t n, const uint64_t* q_off, const uint64_t* t_off, bool rebase,
               bool sync, int type, int match, int mismatch, int gap, int want_cigar, size_t chunk_pairs,
               const std::vector<uint32_t>* tail_groups);
int plan_finish(b200_align_plan* p, b200_ctx* ctx, bool short_scores, bool sync);

// Host destinations of a run made through the host-buffer entry point. When given (and the batch is a uniform
// one cut into several waves), results are downloaded while later waves still run: scores and target_begin
// wave by wave, CIGAR offsets and text in two groups (all waves but the last, then the last).
struct HostOut {
    int32_t* score; uint32_t* target_begin; char* cigar; uint64_t* cigar_off; uint64_t cigar_cap;
    bool done = false;   // set by the run when it has issued (and completed) every download itself
};
int plan_run_impl(b200_align_plan* p, const char* d_q_buf, const char* d_t_buf, int32_t* d_score,
                  uint32_t* d_target_begin, char* d_cigar, uint64_t* d_cigar_off, uint64_t cigar_cap,
                  void* stream, HostOut* ho);
void align_kernels_configure();                             // shared-memory carve-out preferences (once per context)
int align_short_round_pairs(b200_ctx* c, int type, size_t* out);   // pairs one full round of the thread-per-pair fill takes

// ------------------------------------------------------------------ minimizer plan ----
struct b200_min_plan {
    b200_ctx* ctx = nullptr;
    size_t n = 0;
    uint32_t k = 0, w = 0;
    uint64_t tuples = 0;
    size_t n_tiles = 0;
    size_t smem_bytes = 0;
    uint32_t warp_words = 0;   // staged 2-bit words per warp
    uint64_t buf_bytes = 0;   // bytes of the packed sequence buffer the plan was made for (= off[n])
    std::vector<uint64_t> out_off;
    DevBuf d_off, d_out_off, d_fwd, d_tiles;
};

int min_plan_build(b200_min_plan* p, b200_ctx* ctx, size_t n, const uint64_t* off, uint32_t k, uint32_t w,
                   const uint8_t* is_fwd);
