// capi.cu -- host side of the C ABI declared in include/b200map.h: contexts, shape-only
// plans (wave schedule + per-pair descriptors), kernel launches, and the host-buffer
// convenience entry points. No CPU implementation of the hot path lives here: without a
// usable device every entry point returns B200_E_NOGPU.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "../../include/b200map.h"
#include "align_fill_generic.cuh"
#include "align_fill_long.cuh"
#include "align_fill_long16.cuh"
#include "align_fill_short.cuh"
#include "align_walk.cuh"
#include "common.cuh"
#include "mapper.cuh"
#include "minimize.cuh"

using namespace b200;

// ------------------------------------------------------------------ errors ----
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
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

extern "C" const char* b200_last_error(void) { return g_err.c_str(); }
extern "C" int b200_version(void) { return 1; }
extern "C" int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

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
    // options
    int64_t dir_budget_bytes = 48ll << 30;
    int64_t force_generic = 0;
    int64_t long16 = 1;                     // 0 = long pairs stay on the int32 kernel (align_fill_long.cuh)
    int64_t chunk_pairs = 0;
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

// Event brackets around a kernel launch on stream `st` (no-ops unless ctx->profile).
static void prof_begin(b200_ctx* c, cudaStream_t st, int kind) {
    if (!c->profile) return;
    b200_ctx::Span sp{kind, nullptr, nullptr};
    for (cudaEvent_t* e : {&sp.a, &sp.b}) {
        if (!c->event_pool.empty()) { *e = c->event_pool.back(); c->event_pool.pop_back(); }
        else cudaEventCreate(e);
    }
    cudaEventRecord(sp.a, st);
    c->spans.push_back(sp);
}
static void prof_end(b200_ctx* c, cudaStream_t st) {
    if (!c->profile) return;
    cudaEventRecord(c->spans.back().b, st);
}
static void prof_collect(b200_ctx* c, cudaStream_t st) {
    if (!c->profile) return;
    cudaStreamSynchronize(st);
    for (auto& sp : c->spans) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { c->kind_us[sp.kind] += ms * 1e3; c->kind_launches[sp.kind]++; }
        c->event_pool.push_back(sp.a); c->event_pool.push_back(sp.b);
    }
    c->spans.clear();
}

static void tl_mark(b200_ctx* c, cudaStream_t st, const std::string& what) {
    static const bool on = std::getenv("B200_TRACE") && std::atoi(std::getenv("B200_TRACE")) >= 2;
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    c->timeline.push_back({what, e});
}
static void tl_dump(b200_ctx* c) {
    if (c->timeline.empty()) return;
    cudaDeviceSynchronize();
    std::string line = "[b200 timeline ms]";
    for (auto& m : c->timeline) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->timeline[0].e, m.e);
        line += " " + m.what + "=" + std::to_string(ms).substr(0, 5);
    }
    std::fprintf(stderr, "%s\n", line.c_str());
    for (auto& m : c->timeline) cudaEventDestroy(m.e);
    c->timeline.clear();
}

static int set_device(const b200_ctx* c) {
    CU(cudaSetDevice(c->device));
    return B200_OK;
}

extern "C" int b200_ctx_create(int device, b200_ctx** out) {
    if (!out) return fail(B200_E_ARG, "b200_ctx_create: out is null");
    *out = nullptr;
    int n = b200_device_count();
    if (n <= 0) return fail(B200_E_NOGPU, "no CUDA device visible (this library has no CPU fallback)");
    if (device < 0 || device >= n) return fail(B200_E_ARG, "device index out of range");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(B200_E_NOGPU, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                      "; kernels are built for sm_100a only");
    b200_ctx* c = new (std::nothrow) b200_ctx();
    if (!c) return fail(B200_E_NOMEM, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    CU(cudaSetDevice(device));
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    // Kernels only share an SM when they agree on its shared-memory carve-out: the long-pair fills use no shared
    // memory, the tile walkers 16 KB, and with the default preferences a walker CTA did not become resident until
    // the fill running on the SM was over (measured: the "concurrent" walkers finished 2.9 ms after the fill; launched
    // first, they kept the fill out instead). Same explicit preference on all of them.
    {
        const int pct = 20;
        cudaFuncSetAttribute(fill_long16_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(fill_long16_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(fill_long16_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(fill_long_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(fill_long_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(fill_long_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(walk_tile_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(walk_tile_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(walk_tile_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(walk_tile_wait_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(walk_tile_wait_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(walk_tile_wait_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaGetLastError();
    }
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
        c->dir_budget_bytes = std::max<int64_t>(1ll << 30, (int64_t)(free_b / 3));
    *out = c;
    return B200_OK;
}

extern "C" void b200_align_plan_destroy(b200_align_plan* p);
extern "C" void b200_min_plan_destroy(b200_min_plan* p);
extern "C" void b200_ctx_destroy(b200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->host_plan) { b200_align_plan_destroy(c->host_plan); c->host_plan = nullptr; }
    if (c->map_plan) { b200_align_plan_destroy(c->map_plan); c->map_plan = nullptr; }
    if (c->map_min_plan) { b200_min_plan_destroy(c->map_min_plan); c->map_min_plan = nullptr; }
    for (DevBuf& b : c->map_buf) b.release();
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (auto e : c->copy_events) cudaEventDestroy(e);
    for (auto& sp : c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->emit_stream) cudaStreamDestroy(c->emit_stream);
    if (c->pack_stream) cudaStreamDestroy(c->pack_stream);
    for (cudaEvent_t e : c->pack_done) cudaEventDestroy(e);
    for (auto e : c->wave_done) cudaEventDestroy(e);
    if (c->fork_event) cudaEventDestroy(c->fork_event);
    for (WaveSlot& w : c->slot) {
        for (DevBuf* b : {&w.dirs, &w.bnd, &w.bnd_short, &w.progress, &w.stripe_res, &w.counter, &w.fix_work, &w.pair_state}) b->release();
        if (w.done) cudaEventDestroy(w.done);
        if (w.walk_stream) cudaStreamDestroy(w.walk_stream);
        if (w.pre_event) cudaEventDestroy(w.pre_event);
        if (w.walk_event) cudaEventDestroy(w.walk_event);
    }
    for (DevBuf* b : {&c->qpk, &c->tpk, &c->end_i, &c->end_j, &c->runs, &c->n_runs,
                      &c->cigar_len, &c->scan_tmp, &c->flags, &c->total, &c->wave_flagged, &c->d_q, &c->d_t, &c->d_score, &c->d_tb,
                      &c->d_cigar, &c->d_cigar_off, &c->d_seq, &c->d_hash, &c->d_pos, &c->d_flag})
        b->release();
    for (HostBuf* b : {&c->h_q, &c->h_t, &c->h_off, &c->h_small}) b->release();
    delete c;
}

extern "C" int b200_ctx_set_option(b200_ctx* c, const char* key, int64_t value) {
    if (!c || !key) return fail(B200_E_ARG, "null argument");
    const std::string k(key);
    if (k == "dir_budget_bytes") c->dir_budget_bytes = std::max<int64_t>(value, 1 << 20);
    else if (k == "force_generic") c->force_generic = value;
    else if (k == "long16") c->long16 = value;
    else if (k == "overlap_waves") c->overlap_waves = value;
    else if (k == "concurrent_walk") c->concurrent_walk = value;
    else if (k == "chunk_pairs") c->chunk_pairs = value;
    else if (k == "profile") c->profile = value;
    else if (k == "reset_counters") {
        c->kernel_launches = c->h2d_bytes = c->d2h_bytes = 0;
        for (int i = 0; i < 4; ++i) { c->kind_us[i] = 0; c->kind_launches[i] = 0; }
    }
    else return fail(B200_E_ARG, "unknown option " + k);
    return B200_OK;
}

extern "C" int64_t b200_ctx_get_counter(b200_ctx* c, const char* key) {
    if (!c || !key) return -1;
    const std::string k(key);
    if (k == "kernel_launches") return c->kernel_launches;
    if (k == "h2d_bytes") return c->h2d_bytes;
    if (k == "d2h_bytes") return c->d2h_bytes;
    static const char* kinds[4] = {"fill", "walk", "emit", "other"};
    for (int i = 0; i < 4; ++i) {
        if (k == std::string(kinds[i]) + "_ns") return (int64_t)(c->kind_us[i] * 1e3);
        if (k == std::string(kinds[i]) + "_launches") return c->kind_launches[i];
    }
    return -1;
}

// per-thread default contexts for the reference-shaped entry points
static int default_ctx(int device, b200_ctx** out) {
    struct Holder {
        std::vector<b200_ctx*> v;
        ~Holder() { for (auto* c : v) b200_ctx_destroy(c); }
    };
    static thread_local Holder h;
    if (device < 0) return fail(B200_E_ARG, "negative device index");
    if ((size_t)device >= h.v.size()) h.v.resize(device + 1, nullptr);
    if (!h.v[device]) TRY(b200_ctx_create(device, &h.v[device]));
    *out = h.v[device];
    return B200_OK;
}

// ------------------------------------------------------------------ align plan ----
// A wave is a slice of the work order whose direction matrices fit the HBM budget together and
// that is served by one fill kernel class.
struct Wave {
    uint32_t klass;
    uint32_t first, count;   // range in the work order
    uint32_t first_group;    // short class: index of the first ShortGroup
    uint64_t dir_words;
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
    std::vector<cudaEvent_t> wave_events;   // optional, per wave: "this wave's sequence bytes are resident" (host pipeline)
    bool uniform = false;              // every pair has the same (Q,T): descriptors were built on the device
    uint32_t uQ = 0, uT = 0;
    uint64_t u_groups_per_wave = 1;
    uint64_t u_qbase = 0, u_tbase = 0;
    DevBuf d_pairs, d_work, d_groups, d_task_off, d_bnd_off;
    uint64_t max_long_bnd_words = 0;   // boundary rows of the largest long wave
    uint32_t max_long_tasks = 0;

    void reset() {
        max_long_bnd_words = 0; max_long_tasks = 0;
        n = 0; cells = cigar_bound = run_slots = q_bytes = t_bytes = qpk_words = tpk_words = 0;
        max_T = max_Q = max_T_short = max_Q_short = 0; n_short = n_long = 0;
        waves.clear(); h_pairs.clear(); h_order.clear(); patched = false; uniform = false; wave_events.clear();
    }
};

// Uniform batches (every pair the same Q x T, e.g. fixed-length short reads): the descriptors are an
// affine function of the pair index, so they are generated on the device instead of being built on
// the host and copied (48 B per pair).
__global__ void build_uniform_plan_kernel(uint32_t n, uint32_t Q, uint32_t T, uint64_t q_base, uint64_t t_base,
                                          uint64_t words_per_group, uint32_t groups_per_wave,
                                          PairDesc* __restrict__ pairs,
                                          uint32_t* __restrict__ work, ShortGroup* __restrict__ groups) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PairDesc d;
    d.q_off = q_base + (uint64_t)i * Q;
    d.t_off = t_base + (uint64_t)i * T;
    d.dir_off = (uint64_t)((i >> 6) % groups_per_wave) * words_per_group;   // relative to the wave's buffer
    d.run_off = (uint64_t)i * ((uint64_t)Q + T + 1);
    d.qpk_off = (uint64_t)i * (Q / 16 + 2);
    d.tpk_off = (uint64_t)i * (T / 16 + 2);
    d.Q = Q; d.T = T; d.pitch = T;
    const uint32_t slot = i & 63u;
    d.klass = kClassShort | ((slot >> 1) << 8) | ((slot & 1u) << 16);
    pairs[i] = d;
    work[i] = i;
    if (slot == 0) groups[i >> 6] = ShortGroup{d.dir_off, T, 0};
}

static void materialize_uniform_host(b200_align_plan* p) {   // only needed by the non-ACGT fallback
    if (!p->uniform || !p->h_pairs.empty()) return;
    const uint64_t wpg = p->want_cigar ? (uint64_t)div_up(p->uQ, kShortRows) * p->uT * 128 : 0;
    p->h_pairs.resize(p->n);
    p->h_order.resize(p->n);
    for (size_t i = 0; i < p->n; ++i) {
        PairDesc& d = p->h_pairs[i];
        d.q_off = p->u_qbase + i * p->uQ; d.t_off = p->u_tbase + i * p->uT;
        d.dir_off = ((i >> 6) % p->u_groups_per_wave) * wpg; d.run_off = i * ((uint64_t)p->uQ + p->uT + 1);
        d.qpk_off = i * (uint64_t)(p->uQ / 16 + 2); d.tpk_off = i * (uint64_t)(p->uT / 16 + 2);
        d.Q = p->uQ; d.T = p->uT; d.pitch = p->uT;
        const uint32_t slot = (uint32_t)(i & 63u);
        d.klass = kClassShort | ((slot >> 1) << 8) | ((slot & 1u) << 16);
        p->h_order[i] = (uint32_t)i;
    }
}

extern "C" void b200_align_plan_destroy(b200_align_plan* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    p->d_pairs.release();
    p->d_work.release();
    p->d_groups.release();
    p->d_task_off.release();
    p->d_bnd_off.release();
    delete p;
}
extern "C" uint64_t b200_align_plan_cells(const b200_align_plan* p) { return p ? p->cells : 0; }
extern "C" uint64_t b200_align_plan_cigar_bound(const b200_align_plan* p) { return p ? p->cigar_bound : 0; }

static inline uint64_t generic_dir_words(uint32_t Q, uint32_t T) {
    return (uint64_t)div_up(Q, kRowsPerWord) * ((T + 3u) & ~3u);
}

// Can the int16 tagged kernel (align_fill_short.cuh) represent every value of this pair?
static bool short_scores_ok(const Scores& sc, int type) {
    (void)type;
    auto fits8 = [](int v) { return v >= -128 && v <= 127; };
    const long sm = 4l * ((long)sc.match - sc.gap) + 1, sx = 4l * ((long)sc.mismatch - sc.gap) + 1;
    return fits8((int)sm) && fits8((int)sx) && std::abs((long)sc.gap) < 4000 && std::abs((long)sc.match) < 4000 &&
           std::abs((long)sc.mismatch) < 4000;
}
static inline uint64_t long_dir_words(uint32_t Q, uint32_t T) {
    return (uint64_t)div_up(Q, kLongRows) * ((T + 1u) & ~1u) * 2;
}
// The int32 tagged wavefront kernel (align_fill_long.cuh): global / semiGlobal, int8 table entries.
static bool long_scores_ok(const Scores& sc, int type) {
    (void)type;
    auto fits8 = [](long v) { return v >= -128 && v <= 127; };
    const long sm = 4l * ((long)sc.match - sc.gap) + 1, sx = 4l * ((long)sc.mismatch - sc.gap) + 1;
    return fits8(sm) && fits8(sx) && std::abs((long)sc.gap) < (1 << 20) && std::abs((long)sc.match) < (1 << 20) &&
           std::abs((long)sc.mismatch) < (1 << 20);
}
// The packed int16x2 wavefront kernel (align_fill_long16.cuh) keeps every 32-row block relative to a private
// base that is re-centred every kL16Chunk columns, so what must fit in 16 bits is the spread of a block plus
// its drift over one chunk. With g = gap, U = max(g, s_max - g, 0): neighbouring cells obey g <= dH <= U in
// both directions (induction over team_alignment.cpp:104-114; the local clamp only tightens it), hence in the
// moving frame Y = 4H - 4gj + 1 a vertical step changes Y by at most Dv = 4 max(|g|,|U|) + 3 and a horizontal
// step by 0 .. Dh = 4 (U - g) + 3. The induction starts at the borders, whose own step is `init`: it needs
// g <= init <= U, which holds for global (init = g) and, for semiGlobal/local (init = 0), only while g <= 0 --
// with a positive gap score the cells next to a zero border grow by g per row and the bound is gone.
static bool long16_scores_ok(const Scores& sc, int type) {
    if (!long_scores_ok(sc, type)) return false;
    if (type != 0 && sc.gap > 0) return false;
    const long g = sc.gap, smax = std::max(sc.match, sc.mismatch);
    const long U = std::max({g, smax - g, 0l});
    const long Dv = 4 * std::max(std::labs(g), std::labs(U)) + 3, Dh = 4 * (U - g) + 3;
    return 34 * Dv + (kL16Chunk + 4) * Dh + 4 * std::labs(g) + 160 <= 30000;
}
static inline uint32_t long16_pitch(uint32_t T) { return (T + 2u) & ~1u; }   // slots 0..T, even
static inline uint64_t long16_dir_words(uint32_t Q, uint32_t T) {
    return (uint64_t)div_up(Q, kL16LaneRows) * long16_pitch(T) * 4;
}
static bool short_pair_ok(const Scores& sc, uint32_t Q, uint32_t T) {
    const long mx = std::max({std::abs((long)sc.match), std::abs((long)sc.mismatch), std::abs((long)sc.gap), 1l});
    return Q <= 4096 && T <= 4096 && 4l * (((long)Q + T + 2) * mx + std::abs((long)sc.gap) * T + 4) <= 32767;
}

static int plan_finish(b200_align_plan* p, b200_ctx* ctx, bool short_scores, bool sync);

// A class whose direction matrices fit the budget runs as ONE wave (small waves under-fill the machine and
// every wave pays its own tail). Otherwise it is cut into waves of half the budget, two of which are in
// flight at a time (see b200_align_plan_run): the walk of wave k then overlaps the fill of wave k+1.
static inline uint64_t wave_budget_words(const b200_ctx* ctx) { return std::max<uint64_t>((uint64_t)ctx->dir_budget_bytes / 4, 1 << 16); }
static inline uint64_t wave_cap_words(uint64_t budget_words, uint64_t class_total_words) {
    return class_total_words <= budget_words ? budget_words : std::max<uint64_t>(budget_words / 2, 1 << 15);
}

// Fills `p` (fresh or recycled) for a batch. `rebase`: offsets are taken relative to q_off[0] /
// t_off[0] (the host entry points copy only the referenced byte range to the device).
static int plan_build(b200_align_plan* p, b200_ctx* ctx, size_t n, const uint64_t* q_off, const uint64_t* t_off,
                      bool rebase, bool sync, int type, int match, int mismatch, int gap, int want_cigar,
                      size_t chunk_pairs = 0) {
    if (type < 0 || type > 2) return fail(B200_E_TYPE, "Unknown AlignmentType provided.");
    if (n > 0xfffffff0ull) return fail(B200_E_ARG, "batch too large");
    TRY(set_device(ctx));
    p->reset();
    p->ctx = ctx; p->n = n; p->type = type; p->sc = Scores{match, mismatch, gap};
    p->want_cigar = want_cigar != 0;
    const uint64_t qb = (rebase && n) ? q_off[0] : 0, tb = (rebase && n) ? t_off[0] : 0;
    p->q_bytes = n ? q_off[n] - qb : 0;
    p->t_bytes = n ? t_off[n] - tb : 0;
    const bool short_scores = !ctx->force_generic && short_scores_ok(p->sc, type);
    const uint64_t budget_words = wave_budget_words(ctx);

    // ---- uniform fast path -------------------------------------------------------------------
    if (n >= 8192 && short_scores) {
        const uint64_t Q0 = q_off[1] - q_off[0], T0 = t_off[1] - t_off[0];
        bool uni = q_off[1] >= q_off[0] && t_off[1] >= t_off[0] && Q0 <= 4096 && T0 <= 4096 &&
                   short_pair_ok(p->sc, (uint32_t)Q0, (uint32_t)T0);
        // The scan reads 16 bytes per pair from host memory and sits between the start of the upload and the first
        // wave (1 M pairs: 1.2 ms on one core): no early exit inside a block so the compiler can vectorise it, and
        // large batches are split over a few threads.
        auto scan = [&](size_t a, size_t b) -> bool {   // pairs [a, b), a >= 1
            uint64_t bad = 0;
            for (size_t i = a; i < b; ++i)
                bad |= ((q_off[i + 1] - q_off[i]) ^ Q0) | ((t_off[i + 1] - t_off[i]) ^ T0);
            return bad == 0;
        };
        if (uni && n > (1u << 18)) {
            constexpr int kThreads = 4;
            bool ok[kThreads] = {true, true, true, true};
            std::thread th[kThreads - 1];
            for (int t = 1; t < kThreads; ++t)
                th[t - 1] = std::thread([&, t] { ok[t] = scan(std::max<size_t>(1, n * t / kThreads), n * (t + 1) / kThreads); });
            ok[0] = scan(1, n / kThreads);
            for (auto& x : th) x.join();
            uni = ok[0] && ok[1] && ok[2] && ok[3];
        } else if (uni) {
            for (size_t i0 = 1; uni && i0 < n; i0 += 65536) uni = scan(i0, std::min(n, i0 + 65536));
        }
        const uint64_t n_groups = div_up64(n, 64);
        const uint64_t wpg = p->want_cigar ? (uint64_t)div_up((uint32_t)Q0, kShortRows) * T0 * 128 : 0;
        // uniform batches may be cut into equal chunks (whole 64-pair groups) so that the host entry point can
        // overlap the upload of chunk c+1 with the kernels of chunk c
        uint64_t groups_per_wave = chunk_pairs ? std::max<uint64_t>(1, chunk_pairs / 64) : n_groups;
        if (!chunk_pairs && wpg) groups_per_wave = std::max<uint64_t>(1, std::min(n_groups, wave_cap_words(budget_words, n_groups * wpg) / wpg));
        const uint64_t wave_words_u = std::min(n_groups, groups_per_wave) * wpg;
        if (uni && wave_words_u <= (groups_per_wave < n_groups ? std::max<uint64_t>(budget_words / 2, 1 << 15) : budget_words)) {
            p->uniform = true; p->uQ = (uint32_t)Q0; p->uT = (uint32_t)T0; p->u_groups_per_wave = groups_per_wave;
            p->u_qbase = q_off[0] - qb; p->u_tbase = t_off[0] - tb;
            p->run_slots = n * (Q0 + T0 + 1);
            p->qpk_words = n * (Q0 / 16 + 2); p->tpk_words = n * (T0 / 16 + 2);
            p->cells = n * Q0 * T0;
            p->cigar_bound = n * std::max<uint64_t>(2, 2 * (Q0 + T0));
            p->max_T = p->max_T_short = (uint32_t)T0;
            p->max_Q = p->max_Q_short = (uint32_t)Q0;
            p->n_short = n;
            for (uint64_t g0 = 0; g0 < n_groups; g0 += groups_per_wave) {
                const uint64_t g1 = std::min(n_groups, g0 + groups_per_wave);
                const uint64_t first = g0 * 64, last = std::min<uint64_t>(n, g1 * 64);
                p->waves.push_back(Wave{kClassShort, (uint32_t)first, (uint32_t)(last - first), (uint32_t)g0, (g1 - g0) * wpg});
            }
            TRY(p->d_pairs.ensure(n * sizeof(PairDesc)));
            TRY(p->d_work.ensure(n * sizeof(uint32_t)));
            TRY(p->d_groups.ensure(n_groups * sizeof(ShortGroup)));
            build_uniform_plan_kernel<<<(unsigned)div_up64(n, 256), 256, 0, ctx->stream>>>(
                (uint32_t)n, p->uQ, p->uT, p->u_qbase, p->u_tbase, wpg, (uint32_t)groups_per_wave, p->d_pairs.as<PairDesc>(),
                p->d_work.as<uint32_t>(), p->d_groups.as<ShortGroup>());
            ctx->kernel_launches++;
            CU(cudaGetLastError());
            if (sync) CU(cudaStreamSynchronize(ctx->stream));   // the run may use a different stream
            return B200_OK;
        }
    }

    std::vector<PairDesc>& pairs = p->h_pairs;
    pairs.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const uint64_t ql = q_off[i + 1] - q_off[i], tl = t_off[i + 1] - t_off[i];
        if (q_off[i + 1] < q_off[i] || t_off[i + 1] < t_off[i] || ql > 0x3fffffffull || tl > 0x3fffffffull)
            return fail(B200_E_ARG, "offsets must be non-decreasing and sequences shorter than 2^30");
        PairDesc& d = pairs[i];
        d.q_off = q_off[i] - qb; d.t_off = t_off[i] - tb;
        d.Q = (uint32_t)ql; d.T = (uint32_t)tl;
    }
    return plan_finish(p, ctx, short_scores, sync);
}

// Second half of planning, shared with the mapper (which supplies explicit sub-ranges of the read and
// reference buffers): h_pairs[i].{q_off,t_off,Q,T} are set; classify, order, cut into waves, upload.
static int plan_finish(b200_align_plan* p, b200_ctx* ctx, bool short_scores, bool sync) {
    const size_t n = p->n;
    const int type = p->type;
    const uint64_t budget_words = wave_budget_words(ctx);
    std::vector<PairDesc>& pairs = p->h_pairs;
    std::vector<uint32_t> short_list, long_list, generic_list;
    const bool long_scores = !ctx->force_generic && long_scores_ok(p->sc, type);
    p->long16 = long_scores && ctx->long16 && long16_scores_ok(p->sc, type);
    for (size_t i = 0; i < n; ++i) {
        PairDesc& d = pairs[i];
        const uint64_t ql = d.Q, tl = d.T;
        d.pitch = (d.T + 3u) & ~3u;
        d.klass = kClassGeneric;
        d.dir_off = 0;
        d.run_off = p->run_slots;
        p->run_slots += ql + tl + 1;
        d.qpk_off = p->qpk_words; p->qpk_words += ql / 16 + 2;   // one spare word: kernels prefetch one word ahead
        d.tpk_off = p->tpk_words; p->tpk_words += tl / 16 + 2;
        p->cells += ql * tl;
        p->cigar_bound += std::max<uint64_t>(2, 2 * (ql + tl));
        p->max_T = std::max(p->max_T, d.T);
        p->max_Q = std::max(p->max_Q, d.Q);
        if (short_scores && short_pair_ok(p->sc, d.Q, d.T)) short_list.push_back((uint32_t)i);
        else if (long_scores) long_list.push_back((uint32_t)i);
        else generic_list.push_back((uint32_t)i);
    }
    // thread-per-pair only pays off when there are enough pairs to occupy the machine
    if (short_list.size() < 8192) {
        std::vector<uint32_t>& dst = long_scores ? long_list : generic_list;
        dst.insert(dst.end(), short_list.begin(), short_list.end());
        short_list.clear();
    }
    auto cells_of = [&](uint32_t a) { return (uint64_t)pairs[a].Q * pairs[a].T; };
    auto is_sorted_desc = [&](const std::vector<uint32_t>& v, auto key) {
        for (size_t k = 1; k < v.size(); ++k) if (key(v[k - 1]) < key(v[k])) return false;
        return true;
    };
    // short class: neighbours in a 64-pair group should have the same block count and column count
    auto short_key = [&](uint32_t a) { return ((uint64_t)div_up(pairs[a].Q, kShortRows) << 40) | ((uint64_t)pairs[a].T << 20) | pairs[a].Q; };
    if (!is_sorted_desc(short_list, short_key))
        std::stable_sort(short_list.begin(), short_list.end(), [&](uint32_t a, uint32_t b) { return short_key(a) > short_key(b); });
    // warp-per-pair classes: largest first so the dynamic scheduler's tail is made of small pairs
    for (std::vector<uint32_t>* lst : {&long_list, &generic_list})
        if (!is_sorted_desc(*lst, cells_of))
            std::stable_sort(lst->begin(), lst->end(), [&](uint32_t a, uint32_t b) { return cells_of(a) > cells_of(b); });
    p->n_long = long_list.size();

    std::vector<uint32_t>& order = p->h_order;
    order.reserve(n);
    std::vector<ShortGroup> groups;
    p->n_short = short_list.size();
    {   // short waves, in whole groups
        Wave cur{kClassShort, 0, 0, 0, 0};
        uint64_t class_total = 0;
        if (p->want_cigar)
            for (size_t g0 = 0; g0 < short_list.size(); g0 += 64) {   // sorted: the group's first pair has its largest block and column counts
                uint32_t Qg = 0, Tg = 0;
                for (size_t k = g0; k < std::min(short_list.size(), g0 + 64); ++k) { Qg = std::max(Qg, pairs[short_list[k]].Q); Tg = std::max(Tg, pairs[short_list[k]].T); }
                class_total += (uint64_t)div_up(Qg, kShortRows) * Tg * 128;
            }
        const uint64_t cap_words = wave_cap_words(budget_words, class_total);
        for (size_t g0 = 0; g0 < short_list.size(); g0 += 64) {
            const size_t g1 = std::min(short_list.size(), g0 + 64);
            uint32_t Qg = 0, Tg = 0;
            for (size_t k = g0; k < g1; ++k) { Qg = std::max(Qg, pairs[short_list[k]].Q); Tg = std::max(Tg, pairs[short_list[k]].T); }
            p->max_T_short = std::max(p->max_T_short, Tg);
            p->max_Q_short = std::max(p->max_Q_short, Qg);
            const uint64_t words = p->want_cigar ? (uint64_t)div_up(Qg, kShortRows) * Tg * 128 : 0;
            if (cur.count && cur.dir_words + words > cap_words) {
                p->waves.push_back(cur);
                cur = Wave{kClassShort, (uint32_t)order.size(), 0, (uint32_t)groups.size(), 0};
            }
            groups.push_back(ShortGroup{cur.dir_words, Tg, 0});
            for (size_t k = g0; k < g1; ++k) {
                PairDesc& d = pairs[short_list[k]];
                const uint32_t slot = (uint32_t)(k - g0);
                d.klass = kClassShort | ((slot >> 1) << 8) | ((slot & 1u) << 16);
                d.dir_off = cur.dir_words;
                d.pitch = Tg;
                order.push_back(short_list[k]);
            }
            cur.dir_words += words;
            cur.count += (uint32_t)(g1 - g0);
        }
        if (cur.count) p->waves.push_back(cur);
    }
    for (int pass = 0; pass < 2; ++pass) {   // warp-per-pair waves: long class, then generic
        const uint32_t klass = pass == 0 ? kClassLong : kClassGeneric;
        Wave cur{klass, (uint32_t)order.size(), 0, 0, 0};
        auto words_of = [&](const PairDesc& d) -> uint64_t {
            return !p->want_cigar ? 0 : (pass == 0 ? (p->long16 ? long16_dir_words(d.Q, d.T) : long_dir_words(d.Q, d.T))
                                                   : generic_dir_words(d.Q, d.T));
        };
        uint64_t class_total = 0;
        for (uint32_t idx : (pass == 0 ? long_list : generic_list)) class_total += (words_of(pairs[idx]) + 3) & ~3ull;
        const uint64_t cap_words = wave_cap_words(budget_words, class_total);
        for (uint32_t idx : (pass == 0 ? long_list : generic_list)) {
            PairDesc& d = pairs[idx];
            const uint64_t words = words_of(d);
            if (cur.count && cur.dir_words + words > cap_words) {
                p->waves.push_back(cur);
                cur = Wave{klass, (uint32_t)order.size(), 0, 0, 0};
            }
            d.klass = (pass == 0 && p->long16) ? kClassLong16 : klass;
            if (pass == 0) d.pitch = p->long16 ? long16_pitch(d.T) : ((d.T + 1u) & ~1u);
            d.dir_off = cur.dir_words;
            cur.dir_words += (words + 3) & ~3ull;
            ++cur.count;
            order.push_back(idx);
        }
        if (cur.count) p->waves.push_back(cur);
    }

    // long class: stripe hand-out tables, one (count+1)-entry slice per wave, at [wave.first + wave#]
    std::vector<uint32_t> task_off;
    std::vector<uint64_t> bnd_off;
    for (Wave& wv : p->waves) {
        if (wv.klass != kClassLong) continue;
        wv.first_group = (uint32_t)task_off.size();   // reused as the slice start
        uint32_t t = 0; uint64_t b = 0;
        for (uint32_t w = wv.first; w < wv.first + wv.count; ++w) {
            const PairDesc& d = pairs[order[w]];
            task_off.push_back(t); bnd_off.push_back(b);
            const uint32_t ns = (d.Q && d.T) ? div_up(d.Q, p->long16 ? kL16Stripe : kLongRows * kWarp) : 0;
            t += ns; b += (uint64_t)ns * (d.T + 4);
        }
        task_off.push_back(t); bnd_off.push_back(b);
        p->max_long_tasks = std::max(p->max_long_tasks, t);
        p->max_long_bnd_words = std::max(p->max_long_bnd_words, b);
    }
    int rc = p->d_pairs.ensure(std::max<size_t>(1, n) * sizeof(PairDesc));
    if (rc == B200_OK) rc = p->d_work.ensure(std::max<size_t>(1, n) * sizeof(uint32_t));
    if (rc == B200_OK && !task_off.empty()) {
        rc = p->d_task_off.ensure(task_off.size() * 4);
        if (rc == B200_OK) rc = p->d_bnd_off.ensure(bnd_off.size() * 8);
        if (rc == B200_OK) {
            cudaError_t e = cudaMemcpyAsync(p->d_task_off.p, task_off.data(), task_off.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_bnd_off.p, bnd_off.data(), bnd_off.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
            if (e != cudaSuccess) rc = fail(B200_E_CUDA, std::string("plan upload: ") + cudaGetErrorString(e));
        }
    }
    if (rc == B200_OK) rc = p->d_groups.ensure(std::max<size_t>(1, groups.size()) * sizeof(ShortGroup));
    if (rc == B200_OK && n) {
        cudaError_t e = cudaMemcpyAsync(p->d_pairs.p, pairs.data(), n * sizeof(PairDesc), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_work.p, order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && !groups.empty())
            e = cudaMemcpyAsync(p->d_groups.p, groups.data(), groups.size() * sizeof(ShortGroup), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(B200_E_CUDA, std::string("plan upload: ") + cudaGetErrorString(e));
        ctx->h2d_bytes += n * (sizeof(PairDesc) + sizeof(uint32_t));
    }
    return rc;
}

extern "C" int b200_align_plan_create(b200_ctx* ctx, size_t n, const uint64_t* q_off, const uint64_t* t_off,
                                      int type, int match, int mismatch, int gap, int want_cigar,
                                      b200_align_plan** out) {
    if (!ctx || !out || (n && (!q_off || !t_off))) return fail(B200_E_ARG, "b200_align_plan_create: null argument");
    *out = nullptr;
    b200_align_plan* p = new (std::nothrow) b200_align_plan();
    if (!p) return fail(B200_E_NOMEM, "out of host memory");
    p->ctx = ctx;
    const int rc = plan_build(p, ctx, n, q_off, t_off, false, true, type, match, mismatch, gap, want_cigar);
    if (rc != B200_OK) { b200_align_plan_destroy(p); return rc; }
    *out = p;
    return B200_OK;
}

struct U32ToU64 {
    __host__ __device__ uint64_t operator()(const uint32_t& v) const { return (uint64_t)v; }
};

struct RunBufs {   // per-wave device pointers shared by the launch helpers
    const uint8_t *dq, *dt;
    uint32_t* dirs;
    int32_t* score;
    cudaStream_t st;
    WaveSlot* ws;
};

static int launch_fill_generic(b200_align_plan* p, const uint32_t* d_work, uint32_t count, const RunBufs& rb) {
    b200_ctx* c = p->ctx;
    if (!count) return B200_OK;
    const int n_blocks = (int)std::min<uint64_t>((uint64_t)c->sm_count * 8, div_up64(count, 4));
    TRY(rb.ws->bnd.ensure((size_t)n_blocks * 4 * (size_t)(p->max_T + 8) * sizeof(int32_t)));
    CU(cudaMemsetAsync(rb.ws->counter.p, 0, 64, rb.st));
    prof_begin(c, rb.st, 0);
#define GEN(TY) fill_generic_kernel<TY><<<n_blocks, 128, 0, rb.st>>>(rb.dq, rb.dt, p->d_pairs.as<PairDesc>(), d_work, count, \
        rb.ws->counter.as<uint32_t>(), c->flags.as<uint8_t>(), (uint8_t)0, (uint8_t)0, p->sc, rb.dirs, rb.ws->bnd.as<int32_t>(),     \
        p->max_T + 8, rb.score, c->end_i.as<uint32_t>(), c->end_j.as<uint32_t>())
    switch (p->type) { case 0: GEN(0); break; case 1: GEN(1); break; default: GEN(2); break; }
#undef GEN
    prof_end(c, rb.st);
    c->kernel_launches++;
    return B200_OK;
}

static int launch_fill_short(b200_align_plan* p, const Wave& wv, const RunBufs& rb) {
    b200_ctx* c = p->ctx;
    const uint32_t n_groups = (wv.count + 63) / 64;
    int per_sm = 0;
    if (p->type == 0) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_short_kernel<0>, kShortThreads, 0));
    else if (p->type == 1) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_short_kernel<1>, kShortThreads, 0));
    else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_short_kernel<2>, kShortThreads, 0));
    per_sm = std::max(per_sm, 1);
    const int n_blocks = (int)std::min<uint64_t>((uint64_t)c->sm_count * per_sm, div_up64(n_groups, kShortThreads / 32));
    const uint32_t bnd_cols = p->max_T_short + 4;
    TRY(rb.ws->bnd_short.ensure((size_t)n_blocks * (kShortThreads / 32) * bnd_cols * 32 * sizeof(uint32_t)));
    CU(cudaMemsetAsync(rb.ws->counter.p, 0, 64, rb.st));
    const ShortConsts K = make_short_consts(p->sc, p->type);
    prof_begin(c, rb.st, 0);
#define SHORTK(TY) fill_short_kernel<TY><<<n_blocks, kShortThreads, 0, rb.st>>>(                                            \
        c->qpk.as<uint32_t>(), c->tpk.as<uint32_t>(), p->d_pairs.as<PairDesc>(), p->d_work.as<uint32_t>() + wv.first,     \
        wv.count, p->d_groups.as<ShortGroup>() + wv.first_group, rb.ws->counter.as<uint32_t>(), c->flags.as<uint8_t>(), K, \
        rb.dirs, rb.ws->bnd_short.as<uint32_t>(), bnd_cols, rb.score, c->end_i.as<uint32_t>(), c->end_j.as<uint32_t>())
    switch (p->type) { case 0: SHORTK(0); break; case 1: SHORTK(1); break; default: SHORTK(2); break; }
#undef SHORTK
    prof_end(c, rb.st);
    c->kernel_launches++;
    return B200_OK;
}

static int launch_fill_long(b200_align_plan* p, const Wave& wv, const RunBufs& rb) {
    b200_ctx* c = p->ctx;
    int per_sm = 0;
    if (p->long16) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_long16_kernel<1>, 128, 0));
    else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_long_kernel<0>, 128, 0));
    per_sm = std::max(per_sm, 1);
    const uint32_t* d_task_off = p->d_task_off.as<uint32_t>() + wv.first_group;
    const uint64_t* d_bnd_off = p->d_bnd_off.as<uint64_t>() + wv.first_group;
    // every CTA must be resident: stripes wait (poll) on the stripe handed out just before them
    const int n_blocks = (int)std::min<uint64_t>((uint64_t)c->sm_count * per_sm, div_up64(std::max(p->max_long_tasks, 1u), 4));
    TRY(rb.ws->bnd.ensure((p->max_long_bnd_words + 8) * sizeof(int32_t)));
    const size_t prog_words = (size_t)p->max_long_tasks + 8;   // progress counters, then (long16, local) the running maxima
    TRY(rb.ws->progress.ensure(2 * prog_words * 4));
    TRY(rb.ws->stripe_res.ensure(((size_t)p->max_long_tasks + 8) * sizeof(StripeResult)));
    CU(cudaMemsetAsync(rb.ws->counter.p, 0, 64, rb.st));
    CU(cudaMemsetAsync(rb.ws->counter.as<uint32_t>() + 24, 0, 4, rb.st));
    CU(cudaMemsetAsync(rb.ws->progress.p, 0, 2 * prog_words * 4, rb.st));
    const uint32_t* d_work = p->d_work.as<uint32_t>() + wv.first;
    if (p->long16) {
        const ShortConsts K16 = make_short_consts(p->sc, p->type);
        // Waves with CIGARs: the walkers start with the fill and take each pair as soon as its last stripe is in
        // (the traceback of the longest pairs no longer trails the whole wave). Local alignments too: the fill
        // itself keeps the first maximum, so the end cell is known when the last stripe reports. The fill is capped
        // at 152 registers so that three of its CTAs leave room on an SM for a walker CTA.
        WaveSlot& ws = *rb.ws;
        const bool cw = c->concurrent_walk && !c->profile && p->want_cigar && rb.dirs != nullptr;
        uint32_t *d_done = nullptr, *d_ready = nullptr;
        if (cw) {
            TRY(ws.pair_state.ensure(((size_t)wv.count + 8) * 8));
            d_done = ws.pair_state.as<uint32_t>();
            d_ready = d_done + wv.count + 4;
            CU(cudaMemsetAsync(ws.pair_state.p, 0, ((size_t)wv.count + 8) * 8, rb.st));
            CU(cudaMemsetAsync(ws.counter.as<uint32_t>() + 28, 0, 4, rb.st));
            if (!ws.walk_stream) CU(cudaStreamCreateWithFlags(&ws.walk_stream, cudaStreamNonBlocking));
            if (!ws.pre_event) CU(cudaEventCreateWithFlags(&ws.pre_event, cudaEventDisableTiming));
            if (!ws.walk_event) CU(cudaEventCreateWithFlags(&ws.walk_event, cudaEventDisableTiming));
        }
        prof_begin(c, rb.st, 0);
#define LONG16K(TY)                                                                                                    \
    if (cw) {   /* pairs without inner cells have nothing to wait for: result and ready flag now */                   \
        finalize_long_kernel<TY><<<(unsigned)div_up64(wv.count, 128), 128, 0, rb.st>>>(p->d_pairs.as<PairDesc>(), d_work, \
            wv.count, d_task_off, c->flags.as<uint8_t>(), ws.stripe_res.as<StripeResult>(), K16.init, rb.score,        \
            c->end_i.as<uint32_t>(), c->end_j.as<uint32_t>(), 1, d_ready);                                             \
        CU(cudaEventRecord(ws.pre_event, rb.st));                                                                      \
        CU(cudaStreamWaitEvent(ws.walk_stream, ws.pre_event, 0));                                                      \
    }                                                                                                                  \
    fill_long16_kernel<TY><<<n_blocks, 128, 0, rb.st>>>(c->qpk.as<uint32_t>(), c->tpk.as<uint32_t>(), p->d_pairs.as<PairDesc>(), \
        d_work, wv.count, d_task_off, d_bnd_off, rb.ws->counter.as<uint32_t>(), c->flags.as<uint8_t>(), K16, rb.dirs,  \
        rb.ws->bnd.as<int32_t>(), rb.ws->progress.as<uint32_t>(), (uint32_t)prog_words,                                \
        rb.ws->stripe_res.as<StripeResult>(),                                                                          \
        rb.ws->counter.as<uint32_t>() + 24, d_done, d_ready, cw ? rb.score : nullptr,                                  \
        cw ? c->end_i.as<uint32_t>() : nullptr, cw ? c->end_j.as<uint32_t>() : nullptr);                               \
    if (cw) {                                                                                                          \
        walk_tile_wait_kernel<TY><<<(unsigned)c->sm_count, 128, 0, ws.walk_stream>>>(p->d_pairs.as<PairDesc>(), d_work, \
            wv.count, c->flags.as<uint8_t>(), ws.counter.as<uint32_t>() + 28, d_ready, ws.counter.as<uint32_t>() + 24, \
            rb.dirs, c->end_i.as<uint32_t>(), c->end_j.as<uint32_t>(), c->runs.as<uint32_t>(), c->n_runs.as<uint32_t>(), \
            c->cigar_len.as<uint32_t>());                                                                              \
        CU(cudaEventRecord(ws.walk_event, ws.walk_stream));                                                            \
        ws.walk_inflight = true;                                                                                       \
        c->kernel_launches++;                                                                                          \
    }                                                                                                                  \
    if (!cw)                                                                                                           \
    finalize_long_kernel<TY><<<(unsigned)div_up64(wv.count, 128), 128, 0, rb.st>>>(p->d_pairs.as<PairDesc>(), d_work,  \
        wv.count, d_task_off, c->flags.as<uint8_t>(), rb.ws->stripe_res.as<StripeResult>(), K16.init, rb.score,        \
        c->end_i.as<uint32_t>(), c->end_j.as<uint32_t>(), 0, nullptr)
        if (p->type == 0) { LONG16K(0); } else if (p->type == 2) { LONG16K(2); } else { LONG16K(1); }
#undef LONG16K
        prof_end(c, rb.st);
        c->kernel_launches += 2;
        return B200_OK;
    }
    const LongConsts K = make_long_consts(p->sc, p->type);
    prof_begin(c, rb.st, 0);
#define LONGK(TY)                                                                                                      \
    fill_long_kernel<TY><<<n_blocks, 128, 0, rb.st>>>(c->qpk.as<uint32_t>(), c->tpk.as<uint32_t>(), p->d_pairs.as<PairDesc>(), \
        d_work, wv.count, d_task_off, d_bnd_off, rb.ws->counter.as<uint32_t>(), c->flags.as<uint8_t>(), K, rb.dirs,        \
        rb.ws->bnd.as<int32_t>(), rb.ws->progress.as<uint32_t>(), rb.ws->stripe_res.as<StripeResult>(),                            \
        rb.ws->counter.as<uint32_t>() + 24);                                                                               \
    finalize_long_kernel<TY><<<(unsigned)div_up64(wv.count, 128), 128, 0, rb.st>>>(p->d_pairs.as<PairDesc>(), d_work,  \
        wv.count, d_task_off, c->flags.as<uint8_t>(), rb.ws->stripe_res.as<StripeResult>(), K.init, rb.score,              \
        c->end_i.as<uint32_t>(), c->end_j.as<uint32_t>(), 0, nullptr)
    if (p->type == 0) { LONGK(0); } else if (p->type == 2) { LONGK(2); } else {
        LONGK(1);
        locate_long_kernel<<<(unsigned)div_up64((uint64_t)wv.count * 32, 128), 128, 0, rb.st>>>(
            c->qpk.as<uint32_t>(), c->tpk.as<uint32_t>(), p->d_pairs.as<PairDesc>(), d_work, wv.count, d_bnd_off,
            c->flags.as<uint8_t>(), K, rb.ws->bnd.as<int32_t>(), rb.score, c->end_i.as<uint32_t>(), c->end_j.as<uint32_t>());
        c->kernel_launches++;
    }
#undef LONGK
    prof_end(c, rb.st);
    c->kernel_launches += 2;
    return B200_OK;
}

static int launch_walk(b200_align_plan* p, uint32_t wave_klass, const uint32_t* d_work, uint32_t count, const RunBufs& rb,
                       const uint8_t* skip_flags) {
    b200_ctx* c = p->ctx;
    if (!count) return B200_OK;
    prof_begin(c, rb.st, 1);
    if (wave_klass == kClassShort) {
        // thread per pair; few pairs: give each its own quarter-warp or warp so the divergent pointer chases do not serialise
        uint32_t spread = 1;
        while (spread < 32 && (uint64_t)count * spread * 2 <= (uint64_t)c->sm_count * 512) spread *= 2;
        const unsigned wb = (unsigned)div_up64((uint64_t)count * spread, 128);
#define WALK(TY) walk_kernel<TY><<<wb, 128, 0, rb.st>>>(p->d_pairs.as<PairDesc>(), d_work, count, spread, skip_flags, rb.dirs, c->end_i.as<uint32_t>(), \
        c->end_j.as<uint32_t>(), c->runs.as<uint32_t>(), c->n_runs.as<uint32_t>(), c->cigar_len.as<uint32_t>())
        switch (p->type) { case 0: WALK(0); break; case 1: WALK(1); break; default: WALK(2); break; }
#undef WALK
    } else {
        // warp per pair, tile by tile (long and generic layouts)
        const unsigned wb = (unsigned)div_up64((uint64_t)count * 32, 128);
#define WALK(TY) walk_tile_kernel<TY><<<wb, 128, 0, rb.st>>>(p->d_pairs.as<PairDesc>(), d_work, count, skip_flags, rb.dirs, c->end_i.as<uint32_t>(), \
        c->end_j.as<uint32_t>(), c->runs.as<uint32_t>(), c->n_runs.as<uint32_t>(), c->cigar_len.as<uint32_t>())
        switch (p->type) { case 0: WALK(0); break; case 1: WALK(1); break; default: WALK(2); break; }
#undef WALK
    }
    prof_end(c, rb.st);
    c->kernel_launches++;
    return B200_OK;
}

// Host destinations of a run made through the host-buffer entry point. When given (and the batch is a uniform
// one cut into several waves), results are downloaded while later waves still run: scores and target_begin
// wave by wave, CIGAR offsets and text in two groups (all waves but the last, then the last).
struct HostOut {
    int32_t* score; uint32_t* target_begin; char* cigar; uint64_t* cigar_off; uint64_t cigar_cap;
    bool done = false;   // set by the run when it has issued (and completed) every download itself
};

static int plan_run_impl(b200_align_plan* p, const char* d_q_buf, const char* d_t_buf, int32_t* d_score,
                         uint32_t* d_target_begin, char* d_cigar, uint64_t* d_cigar_off, uint64_t cigar_cap,
                         void* stream, HostOut* ho);

extern "C" int b200_align_plan_run(b200_align_plan* p, const char* d_q_buf, const char* d_t_buf,
                                   int32_t* d_score, uint32_t* d_target_begin, char* d_cigar,
                                   uint64_t* d_cigar_off, uint64_t cigar_cap, void* stream) {
    return plan_run_impl(p, d_q_buf, d_t_buf, d_score, d_target_begin, d_cigar, d_cigar_off, cigar_cap, stream, nullptr);
}

static int plan_run_impl(b200_align_plan* p, const char* d_q_buf, const char* d_t_buf, int32_t* d_score,
                         uint32_t* d_target_begin, char* d_cigar, uint64_t* d_cigar_off, uint64_t cigar_cap,
                         void* stream, HostOut* ho) {
    if (!p) return fail(B200_E_ARG, "null plan");
    b200_ctx* c = p->ctx;
    const size_t n = p->n;
    if (p->want_cigar && (!d_cigar_off || (!d_cigar && cigar_cap))) return fail(B200_E_ARG, "plan wants CIGARs but no buffers given");
    if (!d_score) return fail(B200_E_ARG, "d_score is null");
    TRY(set_device(c));
    cudaStream_t st = (cudaStream_t)stream;   // used as given: 0 is the CUDA default stream
    if (n == 0) {
        if (d_cigar_off) CU(cudaMemsetAsync(d_cigar_off, 0, sizeof(uint64_t), st));
        return B200_OK;
    }
    uint64_t max_dir_words = 0;
    for (const Wave& w : p->waves) max_dir_words = std::max(max_dir_words, w.dir_words);
    // Two waves in flight when there are several: wave k runs on stream (k & 1) with workspace slot (k & 1).
    // Profiling runs stay on one stream so the event brackets time each kernel alone.
    const bool overlap = p->waves.size() > 1 && c->overlap_waves && !c->profile;
    const int n_slots = overlap ? 2 : 1;
    if (overlap && !c->aux_stream) CU(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    if (!c->fork_event) CU(cudaEventCreateWithFlags(&c->fork_event, cudaEventDisableTiming));
    for (int k = 0; k < n_slots; ++k) {
        WaveSlot& ws = c->slot[k];
        ws.walk_inflight = false;   // (a run that failed between a fill and its join must not leak into this one)
        TRY(ws.counter.ensure(128));
        if (!ws.done) CU(cudaEventCreateWithFlags(&ws.done, cudaEventDisableTiming));
        if (p->want_cigar) TRY(ws.dirs.ensure(std::max<uint64_t>(max_dir_words, 4) * 4 + 64));
    }
    TRY(c->flags.ensure(n + 8));
    TRY(c->end_i.ensure(n * 4));
    TRY(c->end_j.ensure(n * 4));
    if (p->want_cigar) {
        TRY(c->runs.ensure(p->run_slots * 4));
        TRY(c->n_runs.ensure(n * 4));
        TRY(c->cigar_len.ensure(n * 4));
    }

    const size_t n_packed = p->n_short + p->n_long;   // classes that read the 2-bit copies
    if (n_packed) {
        TRY(c->qpk.ensure((p->qpk_words + p->max_Q / 16 + 72) * 4));
        TRY(c->tpk.ensure((p->tpk_words + p->max_T / 16 + 72) * 4));
        CU(cudaMemsetAsync(c->flags.p, 0, n + 4, st));
    }
    if (p->patched) {   // a previous run (other content) left fallback descriptors behind
        materialize_uniform_host(p);
        CU(cudaMemcpyAsync(p->d_pairs.p, p->h_pairs.data(), n * sizeof(PairDesc), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        p->patched = false;
    }
    // Pairs planned for a 2-bit kernel that turn out not to be pure ACGT are content, not plan: pack_kernel flags and
    // counts them (one counter per wave), every kernel of the wave skips them, and the counters are read back together
    // with the first read-back the run needs anyway (no host round trip per wave). Flagged pairs are rare; when there
    // are any, the repair pass below gives them to the generic kernel after everything else has finished.
    const size_t n_waves = p->waves.size();
    TRY(c->wave_flagged.ensure(n_waves * 4 + 16));
    TRY(c->h_small.ensure((n_waves + 8) * 8));
    uint32_t* h_flagged = c->h_small.as<uint32_t>() + 4;   // [n_waves], after the 8-byte slot of the CIGAR total
    std::memset(c->h_small.p, 0, (n_waves + 8) * 8);
    CU(cudaMemsetAsync(c->wave_flagged.p, 0, n_waves * 4 + 16, st));
    if (overlap) {   // the second stream starts after everything already queued on the caller's stream
        CU(cudaEventRecord(c->fork_event, st));
        CU(cudaStreamWaitEvent(c->aux_stream, c->fork_event, 0));
    }
    // pipelined download (see HostOut): pair index == work position in a uniform plan, so a wave is a contiguous slice
    const bool piped = ho && overlap && p->uniform && p->waves.size() >= 3 && d_target_begin && ho->target_begin;
    if (piped) {
        if (!c->emit_stream) CU(cudaStreamCreateWithFlags(&c->emit_stream, cudaStreamNonBlocking));
        while (c->wave_done.size() < p->waves.size()) {
            cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->wave_done.push_back(e);
        }
    }

    // When the host entry point pipelines the upload, the packing of a wave does not queue behind the traceback of the
    // wave two before it (same stream): it runs on a stream of its own as soon as the wave's bytes are resident, and
    // the wave's stream waits for it. The 2-bit copies and flags of different waves are disjoint.
    const bool pack_ahead = overlap && !p->wave_events.empty();
    if (pack_ahead) {
        if (!c->pack_stream) CU(cudaStreamCreateWithFlags(&c->pack_stream, cudaStreamNonBlocking));
        while (c->pack_done.size() < n_waves) {
            cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->pack_done.push_back(e);
        }
        CU(cudaStreamWaitEvent(c->pack_stream, c->fork_event, 0));   // after the memsets queued on the caller's stream
    }

    if (!ho) tl_mark(c, st, "start");
    // A wave = classify (+ 2-bit pack) -> fill -> traceback walk, in order on its stream; when the host entry
    // point pipelines the upload, wave k first waits for the event that marks its bytes as resident.
    for (size_t k = 0; k < p->waves.size(); ++k) {
        const Wave& wv = p->waves[k];
        WaveSlot& ws = c->slot[overlap ? (k & 1) : 0];
        cudaStream_t wst = (overlap && (k & 1)) ? c->aux_stream : st;
        RunBufs rb{reinterpret_cast<const uint8_t*>(d_q_buf), reinterpret_cast<const uint8_t*>(d_t_buf), nullptr, d_score, wst, &ws};
        uint32_t* d_nflag = c->wave_flagged.as<uint32_t>() + k;
        const uint32_t* work = p->d_work.as<uint32_t>() + wv.first;
        cudaStream_t pst = pack_ahead ? c->pack_stream : wst;
        if (k < p->wave_events.size() && p->wave_events[k]) CU(cudaStreamWaitEvent(pst, p->wave_events[k], 0));
        prof_begin(c, pst, 3);
        if (wv.klass != kClassGeneric) {
            const uint32_t wpp = std::max(1u, div_up(wv.klass == kClassLong ? std::max(p->max_Q, p->max_T)
                                                                           : std::max(p->max_Q_short, p->max_T_short), 16));
            dim3 grid((unsigned)div_up64((uint64_t)wv.count * wpp, 256), 2);
            pack_kernel<<<grid, 256, 0, pst>>>(rb.dq, rb.dt, p->d_pairs.as<PairDesc>(), work, wv.count, wpp,
                                               c->flags.as<uint8_t>(), c->qpk.as<uint32_t>(), c->tpk.as<uint32_t>(), d_nflag);
        } else {
            classify_kernel<<<(unsigned)div_up64((uint64_t)wv.count * 32, 256), 256, 0, pst>>>(
                rb.dq, rb.dt, p->d_pairs.as<PairDesc>(), work, wv.count, c->flags.as<uint8_t>());
        }
        prof_end(c, pst);
        c->kernel_launches++;
        tl_mark(c, pst, "pack" + std::to_string(k));
        if (pack_ahead) {
            CU(cudaEventRecord(c->pack_done[k], pst));
            CU(cudaStreamWaitEvent(wst, c->pack_done[k], 0));
        }
        rb.dirs = p->want_cigar ? ws.dirs.as<uint32_t>() : nullptr;
        // in a wave of a 2-bit class a non-zero flag means "not this wave's pair"; in a generic wave the flags only
        // describe the content (classify_kernel) and every pair is the wave's own
        const uint8_t* skip = wv.klass != kClassGeneric ? c->flags.as<uint8_t>() : nullptr;
        if (wv.klass == kClassShort) TRY(launch_fill_short(p, wv, rb));
        else if (wv.klass != kClassGeneric) TRY(launch_fill_long(p, wv, rb));
        else TRY(launch_fill_generic(p, work, wv.count, rb));
        tl_mark(c, wst, "fill" + std::to_string(k));
        if (ws.walk_inflight) {
            // the wave's pairs are being walked next to the fill; the wave's stream waits for the walkers so that
            // everything after it sees every pair walked
            ws.walk_inflight = false;
            CU(cudaStreamWaitEvent(wst, ws.walk_event, 0));
            tl_mark(c, wst, "cwalk" + std::to_string(k));
        } else if (p->want_cigar) TRY(launch_walk(p, wv.klass, work, wv.count, rb, skip));
        tl_mark(c, wst, "walk" + std::to_string(k));
        if (piped) {
            const uint32_t a = wv.first, b = wv.first + wv.count;
            target_begin_kernel<<<(unsigned)div_up64(wv.count, 256), 256, 0, wst>>>(a, b, p->type, c->end_j.as<uint32_t>(), d_target_begin);
            c->kernel_launches++;
            CU(cudaMemcpyAsync(ho->score + a, d_score + a, (size_t)wv.count * 4, cudaMemcpyDeviceToHost, wst));
            CU(cudaMemcpyAsync(ho->target_begin + a, d_target_begin + a, (size_t)wv.count * 4, cudaMemcpyDeviceToHost, wst));
            c->d2h_bytes += (uint64_t)wv.count * 8;
            CU(cudaEventRecord(c->wave_done[k], wst));
        }
        if (overlap) CU(cudaEventRecord(ws.done, wst));
    }
    CU(cudaGetLastError());
    if (overlap) CU(cudaStreamWaitEvent(st, c->slot[1].done, 0));   // join: the rest runs on the caller's stream

    if (d_target_begin && !piped) {
        target_begin_kernel<<<(unsigned)div_up64(n, 256), 256, 0, st>>>(0u, (uint32_t)n, p->type, c->end_j.as<uint32_t>(), d_target_begin);
        c->kernel_launches++;
    }
    // The per-wave counts of flagged pairs (see above) ride on the read-back that is needed anyway: the copy of waves
    // [k0, k1) is queued on a stream that is ordered after their pack kernels, the caller's next synchronise lands it.
    auto queue_flag_counts = [&](size_t k0, size_t k1, cudaStream_t es) -> int {
        if (n_packed && k1 > k0)
            CU(cudaMemcpyAsync(h_flagged + k0, c->wave_flagged.as<uint32_t>() + k0, (k1 - k0) * 4, cudaMemcpyDeviceToHost, es));
        return B200_OK;
    };
    auto any_flagged = [&](size_t k0, size_t k1) { for (size_t k = k0; k < k1; ++k) if (h_flagged[k]) return true; return false; };
    // CIGAR offsets (scan of the text lengths) and text, for pairs [a, b) on stream es (device scalar d_cigar_off[a] is
    // final by then). Returns the byte count up to pair b through *total_out. If one of the waves [fk0, fk1) counted a
    // flagged pair, *flagged is set and nothing is emitted: the texts of those pairs do not exist yet.
    uint64_t* h_total = c->h_small.as<uint64_t>();
    auto scan_emit = [&](uint32_t a, uint32_t b, cudaStream_t es, uint64_t* total_out, size_t fk0, size_t fk1, bool* flagged) -> int {
        cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> in(c->cigar_len.as<uint32_t>() + a, U32ToU64());
        size_t tmp_bytes = 0;
        CU(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, in, d_cigar_off + a + 1, (int)(b - a), es));
        TRY(c->scan_tmp.ensure(tmp_bytes));
        if (a == 0) CU(cudaMemsetAsync(d_cigar_off, 0, sizeof(uint64_t), es));
        prof_begin(c, es, 3);
        CU(cub::DeviceScan::InclusiveSum(c->scan_tmp.p, tmp_bytes, in, d_cigar_off + a + 1, (int)(b - a), es));
        if (a) add_offset_kernel<<<(unsigned)div_up64(b - a, 256), 256, 0, es>>>(d_cigar_off + a + 1, b - a, d_cigar_off + a);
        prof_end(c, es);
        c->kernel_launches += 2 + (a ? 1 : 0);
        CU(cudaMemcpyAsync(h_total, d_cigar_off + b, sizeof(uint64_t), cudaMemcpyDeviceToHost, es));
        TRY(queue_flag_counts(fk0, fk1, es));
        CU(cudaStreamSynchronize(es));
        if (any_flagged(fk0, fk1)) { *flagged = true; return B200_OK; }
        const uint64_t total = *h_total;
        *total_out = total;
        if (total > cigar_cap)
            return fail(B200_E_CAP, "CIGAR buffer too small: need " + std::to_string(total) + " bytes, have " + std::to_string(cigar_cap));
        prof_begin(c, es, 2);
        // many short pairs: a thread each; fewer, longer pairs (thousands of runs): a warp each
        if (p->n_short * 2 >= n)
            emit_kernel<<<(unsigned)div_up64(b - a, 128), 128, 0, es>>>(p->d_pairs.as<PairDesc>(), a, b, c->runs.as<uint32_t>(), c->n_runs.as<uint32_t>(), d_cigar_off, d_cigar);
        else
            emit_warp_kernel<<<(unsigned)div_up64((uint64_t)(b - a) * 32, 128), 128, 0, es>>>(p->d_pairs.as<PairDesc>(), a, b, c->runs.as<uint32_t>(), c->n_runs.as<uint32_t>(), d_cigar_off, d_cigar);
        prof_end(c, es);
        c->kernel_launches++;
        return B200_OK;
    };
    bool flagged = false;
    if (p->want_cigar && piped) {
        // Offsets, texts and their download go group by group on the emit stream while later waves still run: two
        // waves at a time (one of each wave stream), the last wave alone -- so that what is left to do after the last
        // wave is its own share and not the download of everything before it.
        const size_t nw = n_waves;
        std::vector<size_t> gb{0};
        for (size_t k = 2; k + 1 < nw; k += 2) gb.push_back(k);
        gb.push_back(nw - 1);
        gb.push_back(nw);
        cudaStream_t es = c->emit_stream;
        uint64_t done_bytes = 0;
        for (size_t g = 0; g + 1 < gb.size() && !flagged; ++g) {
            const size_t k0 = gb[g], k1 = gb[g + 1];
            const uint32_t a = p->waves[k0].first, b = k1 < nw ? p->waves[k1].first : (uint32_t)n;
            CU(cudaStreamWaitEvent(es, c->wave_done[k1 - 1], 0));
            if (k1 >= 2) CU(cudaStreamWaitEvent(es, c->wave_done[k1 - 2], 0));   // the other wave stream
            tl_mark(c, es, "ready" + std::to_string(g));
            uint64_t total = 0;
            TRY(scan_emit(a, b, es, &total, k0, k1, &flagged));
            if (flagged) break;
            if (ho->cigar_cap < total) return fail(B200_E_CAP, "CIGAR buffer too small: need " + std::to_string(total));
            if (a == 0) CU(cudaMemcpyAsync(ho->cigar_off, d_cigar_off, ((size_t)b + 1) * 8, cudaMemcpyDeviceToHost, es));
            else CU(cudaMemcpyAsync(ho->cigar_off + a + 1, d_cigar_off + a + 1, ((size_t)b - a) * 8, cudaMemcpyDeviceToHost, es));
            if (total > done_bytes) CU(cudaMemcpyAsync(ho->cigar + done_bytes, d_cigar + done_bytes, total - done_bytes, cudaMemcpyDeviceToHost, es));
            done_bytes = total;
            tl_mark(c, es, "down" + std::to_string(g));
        }
        if (!flagged) {
            CU(cudaStreamSynchronize(es));
            c->d2h_bytes += ((uint64_t)n + 1) * 8 + done_bytes;
            ho->done = true;
        }
    } else if (p->want_cigar) {
        uint64_t total = 0;
        TRY(scan_emit(0, (uint32_t)n, st, &total, 0, n_waves, &flagged));
    } else if (n_packed) {
        TRY(queue_flag_counts(0, n_waves, st));
        CU(cudaStreamSynchronize(st));
        flagged = any_flagged(0, n_waves);
    }
    if (flagged) {
        // Repair pass (rare): the flagged pairs go to the generic kernel once everything queued so far has finished,
        // in chunks whose direction matrices fit the wave budget (they live in the first slot's buffer, free by then);
        // then target_begin, offsets and texts are redone for the whole batch and the caller downloads all of it.
        CU(cudaStreamSynchronize(st));
        if (c->aux_stream) CU(cudaStreamSynchronize(c->aux_stream));
        if (c->emit_stream) CU(cudaStreamSynchronize(c->emit_stream));
        if (c->pack_stream) CU(cudaStreamSynchronize(c->pack_stream));
        materialize_uniform_host(p);
        std::vector<uint8_t> h_flags(n);
        CU(cudaMemcpy(h_flags.data(), c->flags.p, n, cudaMemcpyDeviceToHost));
        std::vector<PairDesc> patched = p->h_pairs;
        std::vector<uint32_t> fix;
        std::vector<size_t> chunk_start{0};
        const uint64_t budget = wave_budget_words(c);
        uint64_t words = 0, max_words = 4;
        for (const Wave& wv : p->waves) {
            if (wv.klass == kClassGeneric) continue;
            for (uint32_t w = wv.first; w < wv.first + wv.count; ++w) {
                const uint32_t idx = p->h_order[w];
                if (!h_flags[idx]) continue;
                PairDesc& d = patched[idx];
                const uint64_t need = p->want_cigar ? generic_dir_words(d.Q, d.T) : 0;
                if (words && ((words + 3) & ~3ull) + need > budget) { chunk_start.push_back(fix.size()); words = 0; }
                d.klass = kClassGeneric;
                d.pitch = (d.T + 3u) & ~3u;
                d.dir_off = (words + 3) & ~3ull;
                words = d.dir_off + need;
                max_words = std::max(max_words, words);
                fix.push_back(idx);
            }
        }
        chunk_start.push_back(fix.size());
        WaveSlot& ws = c->slot[0];
        if (p->want_cigar) TRY(ws.dirs.ensure(max_words * 4 + 64));
        TRY(ws.fix_work.ensure(std::max<size_t>(fix.size(), 1) * 4));
        CU(cudaMemcpyAsync(p->d_pairs.p, patched.data(), n * sizeof(PairDesc), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ws.fix_work.p, fix.data(), fix.size() * 4, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));   // `patched` and `fix` are pageable
        p->patched = true;
        RunBufs rb{reinterpret_cast<const uint8_t*>(d_q_buf), reinterpret_cast<const uint8_t*>(d_t_buf),
                   p->want_cigar ? ws.dirs.as<uint32_t>() : nullptr, d_score, st, &ws};
        for (size_t ch = 0; ch + 1 < chunk_start.size(); ++ch) {
            const uint32_t* d_fix = ws.fix_work.as<uint32_t>() + chunk_start[ch];
            const uint32_t cnt = (uint32_t)(chunk_start[ch + 1] - chunk_start[ch]);
            TRY(launch_fill_generic(p, d_fix, cnt, rb));
            if (p->want_cigar) TRY(launch_walk(p, kClassGeneric, d_fix, cnt, rb, nullptr));
        }
        if (d_target_begin) {
            target_begin_kernel<<<(unsigned)div_up64(n, 256), 256, 0, st>>>(0u, (uint32_t)n, p->type, c->end_j.as<uint32_t>(), d_target_begin);
            c->kernel_launches++;
        }
        if (p->want_cigar) {
            uint64_t total = 0;
            bool again = false;
            TRY(scan_emit(0, (uint32_t)n, st, &total, 0, 0, &again));
        }
    }
    CU(cudaGetLastError());
    if (p->n_long) {
        uint32_t stalled[2] = {0, 0};
        for (int k = 0; k < n_slots; ++k)
            CU(cudaMemcpyAsync(&stalled[k], c->slot[k].counter.as<uint32_t>() + 24, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (stalled[0] | stalled[1]) return fail(B200_E_CUDA, "long-pair kernel: a stripe gave up waiting for its predecessor");
    }
    prof_collect(c, st);
    if (!ho) tl_dump(c);
    return B200_OK;
}

// ------------------------------------------------------------------ align, host buffers ----
#include <chrono>
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

extern "C" int b200_align_batch_packed(b200_ctx* c, size_t n, const char* q_buf, const uint64_t* q_off,
                                       const char* t_buf, const uint64_t* t_off, int type, int match,
                                       int mismatch, int gap, int32_t* score, uint32_t* target_begin,
                                       char* cigar_buf, uint64_t* cigar_off, uint64_t cigar_cap) {
    if (!c) return fail(B200_E_ARG, "null context");
    if (type < 0 || type > 2) return fail(B200_E_TYPE, "Unknown AlignmentType provided.");
    if (n && (!q_off || !t_off || !score)) return fail(B200_E_ARG, "null argument");
    const bool want_cigar = cigar_off != nullptr;
    if (want_cigar && !cigar_buf && cigar_cap) return fail(B200_E_ARG, "cigar_buf is null");
    if (n == 0) { if (cigar_off) cigar_off[0] = 0; return B200_OK; }
    TRY(set_device(c));
    // rebase offsets so that only the referenced byte ranges are copied
    const uint64_t q0 = q_off[0], q1 = q_off[n], t0 = t_off[0], t1 = t_off[n];
    if (((q1 > q0) && !q_buf) || ((t1 > t0) && !t_buf)) return fail(B200_E_ARG, "null sequence buffer");
    PhaseTrace tr;
    // Start the sequence upload first, in kPipe byte slices on a separate copy stream with an event after
    // each slice: the host-side planning below overlaps the DMA, and the plan's waves (for uniform batches,
    // equal chunks of pairs) start as soon as the slice holding their last byte has landed.
    TRY(c->d_q.ensure(q1 - q0 + 64));
    TRY(c->d_t.ensure(t1 - t0 + 64));
    cudaStream_t st = c->stream;
    constexpr int kPipe = 16;
    if (!c->copy_stream) CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    while (c->copy_events.size() < (size_t)kPipe) {
        cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->copy_events.push_back(e);
    }
    const uint64_t qn = q1 - q0, tn = t1 - t0;
    tl_mark(c, c->copy_stream, "start");
    for (int s = 0; s < kPipe; ++s) {
        const uint64_t qa = qn * s / kPipe, qb = qn * (s + 1) / kPipe, ta = tn * s / kPipe, tb = tn * (s + 1) / kPipe;
        if (qb > qa) CU(cudaMemcpyAsync(c->d_q.as<char>() + qa, q_buf + q0 + qa, qb - qa, cudaMemcpyHostToDevice, c->copy_stream));
        if (tb > ta) CU(cudaMemcpyAsync(c->d_t.as<char>() + ta, t_buf + t0 + ta, tb - ta, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(c->copy_events[s], c->copy_stream));
        if (s == 0 || s == kPipe / 2 - 1 || s == kPipe - 1) tl_mark(c, c->copy_stream, "h2d" + std::to_string(s));
    }
    c->h2d_bytes += qn + tn;
    tr.mark("enqueue-h2d");
    if (!c->host_plan) {
        c->host_plan = new (std::nothrow) b200_align_plan();
        if (!c->host_plan) return fail(B200_E_NOMEM, "out of host memory");
        c->host_plan->ctx = c;
    }
    b200_align_plan* plan = c->host_plan;   // recycled: its device and host buffers keep their capacity
    // Uniform short batches are cut into chunks of whole ROUNDS of the thread-per-pair kernel (every resident warp
    // takes one 64-pair group per round): any other size leaves a partly empty last round in every chunk, and the
    // smaller the last chunk, the less work is left when the last byte of the upload lands.
    size_t chunk_pairs = (size_t)c->chunk_pairs;
    if (chunk_pairs == 0) {
        int per_sm = 0;
        if (type == 0) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_short_kernel<0>, kShortThreads, 0));
        else if (type == 1) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_short_kernel<1>, kShortThreads, 0));
        else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_short_kernel<2>, kShortThreads, 0));
        const size_t round_pairs = (size_t)c->sm_count * std::max(per_sm, 1) * (kShortThreads / 32) * 64;
        const size_t rounds_per_chunk = std::max<size_t>(1, div_up64(div_up64(n, round_pairs), 16));   // at most 16 chunks
        chunk_pairs = round_pairs * rounds_per_chunk;
    }
    TRY(plan_build(plan, c, n, q_off, t_off, true, false, type, match, mismatch, gap, want_cigar ? 1 : 0, chunk_pairs));
    // which upload slice does each wave have to wait for
    plan->wave_events.assign(plan->waves.size(), c->copy_events[kPipe - 1]);
    if (plan->uniform) {
        for (size_t k = 0; k < plan->waves.size(); ++k) {
            const uint64_t last_pair = (uint64_t)plan->waves[k].first + plan->waves[k].count;   // exclusive
            const uint64_t qe = last_pair * plan->uQ, te = last_pair * plan->uT;                // bytes needed (exclusive)
            int need = 0;
            while (need < kPipe - 1 && (qn * (need + 1) / kPipe < qe || tn * (need + 1) / kPipe < te)) ++need;
            plan->wave_events[k] = c->copy_events[need];
        }
    }
    tr.mark("plan");

    const uint64_t dev_cigar_cap = want_cigar ? std::min<uint64_t>(plan->cigar_bound, std::max<uint64_t>(cigar_cap, 2)) : 0;
    TRY(c->d_score.ensure(n * 4));
    TRY(c->d_tb.ensure(n * 4));
    if (want_cigar) { TRY(c->d_cigar.ensure(dev_cigar_cap + 16)); TRY(c->d_cigar_off.ensure((n + 1) * 8)); }
    if (tr.on) { cudaStreamSynchronize(st); tr.mark("alloc+h2d"); }
    HostOut ho{score, target_begin, cigar_buf, cigar_off, cigar_cap};
    TRY(plan_run_impl(plan, c->d_q.as<char>(), c->d_t.as<char>(), c->d_score.as<int32_t>(),
                      c->d_tb.as<uint32_t>(), want_cigar ? c->d_cigar.as<char>() : nullptr,
                      want_cigar ? c->d_cigar_off.as<uint64_t>() : nullptr, dev_cigar_cap, st, want_cigar ? &ho : nullptr));
    if (tr.on) { cudaStreamSynchronize(st); tr.mark("run"); }
    if (ho.done) {   // everything was downloaded while the last wave ran
        CU(cudaStreamSynchronize(st));
        tr.mark("d2h");
        tl_dump(c);
        return B200_OK;
    }
    CU(cudaMemcpyAsync(score, c->d_score.p, n * 4, cudaMemcpyDeviceToHost, st));
    c->d2h_bytes += n * 4;
    if (target_begin) { CU(cudaMemcpyAsync(target_begin, c->d_tb.p, n * 4, cudaMemcpyDeviceToHost, st)); c->d2h_bytes += n * 4; }
    if (want_cigar) {
        CU(cudaMemcpyAsync(cigar_off, c->d_cigar_off.p, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const uint64_t total = cigar_off[n];
        if (total > cigar_cap) return fail(B200_E_CAP, "CIGAR buffer too small: need " + std::to_string(total));
        if (total) CU(cudaMemcpyAsync(cigar_buf, c->d_cigar.p, total, cudaMemcpyDeviceToHost, st));
        c->d2h_bytes += (n + 1) * 8 + total;
    }
    CU(cudaStreamSynchronize(st));
    tr.mark("d2h");
    return B200_OK;
}

extern "C" int b200_align_batch(int device, size_t n, const char* const* query, const uint32_t* query_len,
                                const char* const* target, const uint32_t* target_len, int type, int match,
                                int mismatch, int gap, int32_t* score, uint32_t* target_begin,
                                char* cigar_buf, uint64_t* cigar_off, uint64_t cigar_cap) {
    if (type < 0 || type > 2) return fail(B200_E_TYPE, "Unknown AlignmentType provided.");
    if (n && (!query || !query_len || !target || !target_len || !score)) return fail(B200_E_ARG, "null argument");
    b200_ctx* c = nullptr;
    TRY(default_ctx(device, &c));
    if (n == 0) { if (cigar_off) cigar_off[0] = 0; return B200_OK; }
    TRY(set_device(c));
    uint64_t qtot = 0, ttot = 0;
    for (size_t i = 0; i < n; ++i) { qtot += query_len[i]; ttot += target_len[i]; }
    TRY(c->h_q.ensure(qtot + 1));
    TRY(c->h_t.ensure(ttot + 1));
    TRY(c->h_off.ensure(2 * (n + 1) * sizeof(uint64_t)));
    uint64_t* qo = c->h_off.as<uint64_t>();
    uint64_t* to = qo + (n + 1);
    uint64_t qa = 0, ta = 0;
    for (size_t i = 0; i < n; ++i) {
        if ((query_len[i] && !query[i]) || (target_len[i] && !target[i])) return fail(B200_E_ARG, "null sequence pointer");
        qo[i] = qa; to[i] = ta;
        if (query_len[i]) std::memcpy(c->h_q.as<char>() + qa, query[i], query_len[i]);
        if (target_len[i]) std::memcpy(c->h_t.as<char>() + ta, target[i], target_len[i]);
        qa += query_len[i]; ta += target_len[i];
    }
    qo[n] = qa; to[n] = ta;
    return b200_align_batch_packed(c, n, c->h_q.as<char>(), qo, c->h_t.as<char>(), to, type, match, mismatch, gap,
                                   score, target_begin, cigar_buf, cigar_off, cigar_cap);
}

// ------------------------------------------------------------------ minimizers ----
extern "C" uint64_t b200_minimize_count(uint32_t len, uint32_t k, uint32_t w) {
    if (len < k || w == 0) return 0;
    const uint64_t n = (uint64_t)len - k + 1;
    const uint64_t full = n >= w ? n - w + 1 : 0;
    const uint64_t tail = n < (uint64_t)w - 1 ? n : (uint64_t)w - 1;
    return (uint64_t)(w - 1) + full + tail;
}

struct b200_min_plan {
    b200_ctx* ctx = nullptr;
    size_t n = 0;
    uint32_t k = 0, w = 0;
    uint64_t tuples = 0;
    size_t n_tiles = 0;
    size_t smem_bytes = 0;
    uint64_t buf_bytes = 0;   // bytes of the packed sequence buffer the plan was made for (= off[n])
    std::vector<uint64_t> out_off;
    DevBuf d_off, d_out_off, d_fwd, d_tiles;
};

extern "C" void b200_min_plan_destroy(b200_min_plan* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    p->d_off.release(); p->d_out_off.release(); p->d_fwd.release(); p->d_tiles.release();
    delete p;
}
extern "C" uint64_t b200_min_plan_tuples(const b200_min_plan* p) { return p ? p->tuples : 0; }
extern "C" const uint64_t* b200_min_plan_out_off(const b200_min_plan* p) { return p ? p->out_off.data() : nullptr; }

// Fills `p` (fresh or recycled: its device buffers keep their capacity) for a batch.
static int min_plan_build(b200_min_plan* p, b200_ctx* ctx, size_t n, const uint64_t* off, uint32_t k, uint32_t w,
                          const uint8_t* is_fwd) {
    TRY(set_device(ctx));
    p->ctx = ctx; p->n = n; p->k = k; p->w = w;
    p->out_off.assign(n + 1, 0);
    std::vector<MinTile> tiles;
    std::vector<uint8_t> fwd(std::max<size_t>(n, 1), 1);
    for (size_t i = 0; i < n; ++i) {
        if (off[i + 1] < off[i] || off[i + 1] - off[i] > 0xfffffff0ull) return fail(B200_E_ARG, "bad offsets");
        const uint64_t cnt = b200_minimize_count((uint32_t)(off[i + 1] - off[i]), k, w);
        p->out_off[i + 1] = p->out_off[i] + cnt;
        // tiles end at multiples of kMinTile in the GLOBAL output index space (vector stores need the alignment)
        if (is_fwd) fwd[i] = is_fwd[i] ? 1 : 0;
        for (uint64_t f = 0; f < cnt;) {
            tiles.push_back(MinTile{off[i], p->out_off[i] + f, (uint32_t)(off[i + 1] - off[i]), (uint32_t)f, fwd[i], 0u});
            f += kMinTile - ((p->out_off[i] + f) & (uint64_t)(kMinTile - 1));
        }
    }
    p->tuples = p->out_off[n];
    p->n_tiles = tiles.size();
    // shared memory: the packed 2-bit codes of every base the tile can touch (+ spare words, see the kernel)
    const uint64_t nx = (uint64_t)kMinTile + 2ull * w + 1;
    const uint64_t nwords = (15 + nx + k - 1 + 15) / 16 + 3;
    p->smem_bytes = (size_t)(nwords * 4);
    p->buf_bytes = n ? off[n] : 0;
    if (p->smem_bytes > 200 * 1024) return fail(B200_E_ARG, "window/k-mer length too large for the shared-memory tile");
    TRY(p->d_off.ensure((n + 1) * 8));
    TRY(p->d_out_off.ensure((n + 1) * 8));
    TRY(p->d_fwd.ensure(std::max<size_t>(n, 1)));
    TRY(p->d_tiles.ensure(std::max<size_t>(tiles.size(), 1) * sizeof(MinTile)));
    if (n) CU(cudaMemcpyAsync(p->d_off.p, off, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(p->d_out_off.p, p->out_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (n) CU(cudaMemcpyAsync(p->d_fwd.p, fwd.data(), n, cudaMemcpyHostToDevice, ctx->stream));
    if (!tiles.empty()) CU(cudaMemcpyAsync(p->d_tiles.p, tiles.data(), tiles.size() * sizeof(MinTile), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));   // the sources are stack/heap temporaries
    return B200_OK;
}

extern "C" int b200_min_plan_create(b200_ctx* ctx, size_t n, const uint64_t* off, uint32_t k, uint32_t w,
                                    const uint8_t* is_fwd, b200_min_plan** out) {
    if (!ctx || !out || (n && !off)) return fail(B200_E_ARG, "b200_min_plan_create: null argument");
    *out = nullptr;
    b200_min_plan* p = new (std::nothrow) b200_min_plan();
    if (!p) return fail(B200_E_NOMEM, "out of host memory");
    p->ctx = ctx;
    const int rc = min_plan_build(p, ctx, n, off, k, w, is_fwd);
    if (rc != B200_OK) { b200_min_plan_destroy(p); return rc; }
    *out = p;
    return B200_OK;
}

extern "C" int b200_min_plan_run(b200_min_plan* p, const char* d_buf, uint32_t* d_hash, uint32_t* d_pos,
                                 uint8_t* d_flag, void* stream) {
    if (!p) return fail(B200_E_ARG, "null plan");
    if (p->n_tiles == 0) return B200_OK;
    if (!d_buf || !d_hash || !d_pos || !d_flag) return fail(B200_E_ARG, "null device buffer");
    b200_ctx* c = p->ctx;
    TRY(set_device(c));
    cudaStream_t st = (cudaStream_t)stream;   // used as given: 0 is the CUDA default stream
    // register fast path for the common window lengths; it stores 16-byte vectors, so the output arrays must be aligned
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_hash) | reinterpret_cast<uintptr_t>(d_pos)) & 15u) == 0 &&
                         (reinterpret_cast<uintptr_t>(d_flag) & 7u) == 0;
    const uint32_t W = (aligned && p->w >= 1 && p->w <= (uint32_t)kMinMaxW) ? p->w : 0;
#define MINK(WW)                                                                                                         \
    case WW:                                                                                                             \
        if (p->smem_bytes > 48 * 1024)                                                                                   \
            CU(cudaFuncSetAttribute(minimize_kernel<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes)); \
        minimize_kernel<WW><<<(unsigned)p->n_tiles, kMinThreads, p->smem_bytes, st>>>(                                   \
            reinterpret_cast<const uint8_t*>(d_buf), p->d_tiles.as<MinTile>(), p->k, p->w, p->buf_bytes, d_hash,         \
            d_pos, d_flag);                                                                                              \
        break;
    switch (W) { MINK(1) MINK(2) MINK(3) MINK(4) MINK(5) MINK(6) MINK(7) MINK(8) default: MINK(0) }
#undef MINK
    c->kernel_launches++;
    CU(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_minimize_batch_packed(b200_ctx* c, size_t n, const char* buf, const uint64_t* off, uint32_t k,
                                          uint32_t w, const uint8_t* is_fwd, uint32_t* hash, uint32_t* pos,
                                          uint8_t* flag, uint64_t* out_off, uint64_t cap) {
    if (!c) return fail(B200_E_ARG, "null context");
    if (n && (!off || !out_off)) return fail(B200_E_ARG, "null argument");
    if (n == 0) { if (out_off) out_off[0] = 0; return B200_OK; }
    TRY(set_device(c));
    const uint64_t b0 = off[0], b1 = off[n];
    std::vector<uint64_t> o(n + 1);
    for (size_t i = 0; i <= n; ++i) o[i] = off[i] - b0;
    b200_min_plan* plan = nullptr;
    TRY(b200_min_plan_create(c, n, o.data(), k, w, is_fwd, &plan));
    struct Guard { b200_min_plan* p; ~Guard() { b200_min_plan_destroy(p); } } guard{plan};
    std::memcpy(out_off, plan->out_off.data(), (n + 1) * 8);
    const uint64_t tot = plan->tuples;
    if (tot > cap) return fail(B200_E_CAP, "minimizer output needs " + std::to_string(tot) + " tuples");
    if (tot == 0) return B200_OK;
    if (!hash || !pos || !flag || !buf) return fail(B200_E_ARG, "null buffer");
    TRY(c->d_seq.ensure(b1 - b0 + 64));
    TRY(c->d_hash.ensure(tot * 4));
    TRY(c->d_pos.ensure(tot * 4));
    TRY(c->d_flag.ensure(tot));
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->d_seq.p, buf + b0, b1 - b0, cudaMemcpyHostToDevice, st));
    c->h2d_bytes += b1 - b0;
    TRY(b200_min_plan_run(plan, c->d_seq.as<char>(), c->d_hash.as<uint32_t>(), c->d_pos.as<uint32_t>(), c->d_flag.as<uint8_t>(), st));
    CU(cudaMemcpyAsync(hash, c->d_hash.p, tot * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(pos, c->d_pos.p, tot * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(flag, c->d_flag.p, tot, cudaMemcpyDeviceToHost, st));
    c->d2h_bytes += tot * 9;
    CU(cudaStreamSynchronize(st));
    return B200_OK;
}

extern "C" int b200_minimize_batch(int device, size_t n, const char* const* seq, const uint32_t* len, uint32_t k,
                                   uint32_t w, const uint8_t* is_fwd, uint32_t* hash, uint32_t* pos, uint8_t* flag,
                                   uint64_t* out_off, uint64_t cap) {
    if (n && (!seq || !len || !out_off)) return fail(B200_E_ARG, "null argument");
    b200_ctx* c = nullptr;
    TRY(default_ctx(device, &c));
    if (n == 0) { if (out_off) out_off[0] = 0; return B200_OK; }
    uint64_t tot = 0;
    for (size_t i = 0; i < n; ++i) tot += len[i];
    TRY(c->h_q.ensure(tot + 1));
    TRY(c->h_off.ensure((n + 1) * 8));
    uint64_t* o = c->h_off.as<uint64_t>();
    uint64_t a = 0;
    for (size_t i = 0; i < n; ++i) {
        if (len[i] && !seq[i]) return fail(B200_E_ARG, "null sequence pointer");
        o[i] = a;
        if (len[i]) std::memcpy(c->h_q.as<char>() + a, seq[i], len[i]);
        a += len[i];
    }
    o[n] = a;
    return b200_minimize_batch_packed(c, n, c->h_q.as<char>(), o, k, w, is_fwd, hash, pos, flag, out_off, cap);
}


// ------------------------------------------------------------------ mapper ("next" rows) ----
struct b200_index {
    b200_ctx* ctx = nullptr;
    uint64_t ref_len = 0;
    uint32_t k = 0, w = 0;
    double f = 0;
    DevBuf d_ref;              // [reference | reverse complement], 2 * ref_len bytes
    DevBuf keys_fwd, keys_rev; // sorted distinct hash<<32|pos
    uint64_t n_fwd = 0, n_rev = 0;
    uint64_t stats[6] = {0, 0, 0, 0, 0, 0};
};

extern "C" void b200_index_destroy(b200_index* ix) {
    if (!ix) return;
    cudaSetDevice(ix->ctx->device);
    ix->d_ref.release(); ix->keys_fwd.release(); ix->keys_rev.release();
    delete ix;
}
extern "C" int b200_index_stats(const b200_index* ix, uint64_t what[6]) {
    if (!ix || !what) return fail(B200_E_ARG, "null argument");
    for (int i = 0; i < 6; ++i) what[i] = ix->stats[i];
    return B200_OK;
}

struct U8ToU32 {
    __host__ __device__ uint32_t operator()(const uint8_t& v) const { return v ? 1u : 0u; }
};

// exclusive scan of 0/1 byte flags into uint32 slots; returns the total through *total
static int scan_flags(b200_ctx* c, const uint8_t* d_flags, uint64_t n, uint32_t* d_slot, uint32_t* total, cudaStream_t st) {
    *total = 0;
    if (n == 0) return B200_OK;
    if (n > 0x7fffffffull) return fail(B200_E_ARG, "too many items for one scan");
    cub::TransformInputIterator<uint32_t, U8ToU32, const uint8_t*> in(d_flags, U8ToU32());
    size_t tmp = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, d_slot, (int)n, st));
    TRY(c->scan_tmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(c->scan_tmp.p, tmp, in, d_slot, (int)n, st));
    c->kernel_launches += 2;
    uint32_t last_slot = 0; uint8_t last_flag = 0;
    CU(cudaMemcpyAsync(&last_slot, d_slot + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&last_flag, d_flags + (n - 1), 1, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *total = last_slot + (last_flag ? 1u : 0u);
    return B200_OK;
}
static int scan_u32(b200_ctx* c, const uint32_t* d_in, uint64_t n, uint32_t* d_out, uint32_t* total, cudaStream_t st) {
    *total = 0;
    if (n == 0) return B200_OK;
    if (n > 0x7fffffffull) return fail(B200_E_ARG, "too many items for one scan");
    size_t tmp = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_in, d_out, (int)n, st));
    TRY(c->scan_tmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(c->scan_tmp.p, tmp, d_in, d_out, (int)n, st));
    c->kernel_launches += 2;
    uint32_t a = 0, b = 0;
    CU(cudaMemcpyAsync(&a, d_out + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&b, d_in + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *total = a + b;
    return B200_OK;
}

// sort + unique one strand's keys; result left in `dst`
static int sort_unique_keys(b200_ctx* c, uint64_t* d_keys, uint64_t n, DevBuf& tmp_keys, DevBuf& dst, uint64_t* n_out,
                            cudaStream_t st) {
    *n_out = 0;
    if (n == 0) return dst.ensure(8);
    TRY(tmp_keys.ensure(n * 8));
    TRY(dst.ensure(n * 8));
    size_t tb = 0;
    CU(cub::DeviceRadixSort::SortKeys(nullptr, tb, d_keys, tmp_keys.as<uint64_t>(), (int)n, 0, 64, st));
    TRY(c->scan_tmp.ensure(tb));
    CU(cub::DeviceRadixSort::SortKeys(c->scan_tmp.p, tb, d_keys, tmp_keys.as<uint64_t>(), (int)n, 0, 64, st));
    TRY(c->total.ensure(16));
    size_t ub = 0;
    CU(cub::DeviceSelect::Unique(nullptr, ub, tmp_keys.as<uint64_t>(), dst.as<uint64_t>(), c->total.as<uint64_t>(), (int)n, st));
    TRY(c->scan_tmp.ensure(ub));
    CU(cub::DeviceSelect::Unique(c->scan_tmp.p, ub, tmp_keys.as<uint64_t>(), dst.as<uint64_t>(), c->total.as<uint64_t>(), (int)n, st));
    c->kernel_launches += 4;
    uint64_t cnt = 0;
    CU(cudaMemcpyAsync(&cnt, c->total.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_out = cnt;
    return B200_OK;
}

extern "C" int b200_index_build(b200_ctx* c, const char* ref, uint64_t ref_len, uint32_t k, uint32_t w, double f,
                                b200_index** out) {
    if (!c || !out || (ref_len && !ref)) return fail(B200_E_ARG, "b200_index_build: null argument");
    *out = nullptr;
    if (ref_len > 0x7ffffff0ull) return fail(B200_E_ARG, "reference longer than 2^31 is not supported");
    TRY(set_device(c));
    cudaStream_t st = c->stream;
    b200_index* ix = new (std::nothrow) b200_index();
    if (!ix) return fail(B200_E_NOMEM, "out of host memory");
    ix->ctx = c; ix->ref_len = ref_len; ix->k = k; ix->w = w; ix->f = f;
    struct Guard { b200_index* p; bool ok = false; ~Guard() { if (!ok) b200_index_destroy(p); } } guard{ix};
    TRY(ix->d_ref.ensure(2 * ref_len + 64));
    uint8_t* d_fwd = ix->d_ref.as<uint8_t>();
    uint8_t* d_rc = d_fwd + ref_len;
    if (ref_len) {
        CU(cudaMemcpyAsync(d_fwd, ref, ref_len, cudaMemcpyHostToDevice, st));
        c->h2d_bytes += ref_len;
        revcomp_kernel<<<(unsigned)div_up64(ref_len, 256), 256, 0, st>>>(d_fwd, d_rc, ref_len);
        c->kernel_launches++;
    }
    // MinimizeBatch over the two strands (flags true / false like KMER ref(true), ref_rev(false), :417-427)
    const uint64_t off[3] = {0, ref_len, 2 * ref_len};
    const uint8_t fw[2] = {1, 0};
    b200_min_plan* mp = nullptr;
    TRY(b200_min_plan_create(c, 2, off, k, w, fw, &mp));
    struct MG { b200_min_plan* p; ~MG() { b200_min_plan_destroy(p); } } mg{mp};
    const uint64_t tot = mp->tuples, n1 = mp->out_off[1];
    ix->stats[0] = n1; ix->stats[1] = tot - n1;
    DevBuf d_hash, d_pos, d_flag, d_keys, d_tmp;
    struct BG { std::vector<DevBuf*> v; ~BG() { for (auto* b : v) b->release(); } } bg{{&d_hash, &d_pos, &d_flag, &d_keys, &d_tmp}};
    TRY(d_hash.ensure(std::max<uint64_t>(tot, 1) * 4));
    TRY(d_pos.ensure(std::max<uint64_t>(tot, 1) * 4));
    TRY(d_flag.ensure(std::max<uint64_t>(tot, 1)));
    TRY(d_keys.ensure(std::max<uint64_t>(tot, 1) * 8));
    if (tot) {
        TRY(b200_min_plan_run(mp, ix->d_ref.as<char>(), d_hash.as<uint32_t>(), d_pos.as<uint32_t>(), d_flag.as<uint8_t>(), st));
        make_keys_kernel<<<(unsigned)div_up64(tot, 256), 256, 0, st>>>(d_hash.as<uint32_t>(), d_pos.as<uint32_t>(), tot, d_keys.as<uint64_t>());
        c->kernel_launches++;
    }
    // frequency filter: the top int(f * |distinct reverse tuples|) forward hashes by window count, applied to both
    // strands (team_mapper.cpp:433-434, :447-450, :467-470)
    TRY(sort_unique_keys(c, d_keys.as<uint64_t>(), n1, d_tmp, ix->keys_fwd, &ix->n_fwd, st));
    TRY(sort_unique_keys(c, d_keys.as<uint64_t>() + n1, tot - n1, d_tmp, ix->keys_rev, &ix->n_rev, st));
    ix->stats[2] = ix->n_fwd; ix->stats[3] = ix->n_rev;
    const long n_ban_want = (f > 0 && n1) ? (long)(f * (double)ix->n_rev) : 0;
    if (n_ban_want > 0) {
        // on the device: sort the forward hashes, run-length encode them into (hash, window count), order by
        // (count desc, hash asc) through one more radix sort on (~count << 32 | hash), keep the first n_ban
        DevBuf d_hs, d_uh, d_uc, d_fk, d_fk2, d_ban, d_keep;
        struct BG2 { std::vector<DevBuf*> v; ~BG2() { for (auto* b : v) b->release(); } } bg2{{&d_hs, &d_uh, &d_uc, &d_fk, &d_fk2, &d_ban, &d_keep}};
        TRY(d_hs.ensure(n1 * 4)); TRY(d_uh.ensure(n1 * 4)); TRY(d_uc.ensure(n1 * 4));
        TRY(c->total.ensure(16));
        size_t tb = 0;
        CU(cub::DeviceRadixSort::SortKeys(nullptr, tb, d_hash.as<uint32_t>(), d_hs.as<uint32_t>(), (int)n1, 0, 32, st));
        TRY(c->scan_tmp.ensure(tb));
        CU(cub::DeviceRadixSort::SortKeys(c->scan_tmp.p, tb, d_hash.as<uint32_t>(), d_hs.as<uint32_t>(), (int)n1, 0, 32, st));
        CU(cub::DeviceRunLengthEncode::Encode(nullptr, tb, d_hs.as<uint32_t>(), d_uh.as<uint32_t>(), d_uc.as<uint32_t>(), c->total.as<uint32_t>(), (int)n1, st));
        TRY(c->scan_tmp.ensure(tb));
        CU(cub::DeviceRunLengthEncode::Encode(c->scan_tmp.p, tb, d_hs.as<uint32_t>(), d_uh.as<uint32_t>(), d_uc.as<uint32_t>(), c->total.as<uint32_t>(), (int)n1, st));
        uint32_t n_distinct = 0;
        CU(cudaMemcpyAsync(&n_distinct, c->total.p, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const uint32_t n_ban = (uint32_t)std::min<long>(n_ban_want, (long)n_distinct);
        TRY(d_fk.ensure((size_t)n_distinct * 8)); TRY(d_fk2.ensure((size_t)n_distinct * 8)); TRY(d_ban.ensure((size_t)n_ban * 4 + 4));
        freq_keys_kernel<<<(unsigned)div_up64(n_distinct, 256), 256, 0, st>>>(d_uh.as<uint32_t>(), d_uc.as<uint32_t>(), n_distinct, d_fk.as<uint64_t>());
        CU(cub::DeviceRadixSort::SortKeys(nullptr, tb, d_fk.as<uint64_t>(), d_fk2.as<uint64_t>(), (int)n_distinct, 0, 64, st));
        TRY(c->scan_tmp.ensure(tb));
        CU(cub::DeviceRadixSort::SortKeys(c->scan_tmp.p, tb, d_fk.as<uint64_t>(), d_fk2.as<uint64_t>(), (int)n_distinct, 0, 64, st));
        low_words_kernel<<<(unsigned)div_up64(n_ban, 256), 256, 0, st>>>(d_fk2.as<uint64_t>(), n_ban, d_uh.as<uint32_t>());
        CU(cub::DeviceRadixSort::SortKeys(nullptr, tb, d_uh.as<uint32_t>(), d_ban.as<uint32_t>(), (int)n_ban, 0, 32, st));
        TRY(c->scan_tmp.ensure(tb));
        CU(cub::DeviceRadixSort::SortKeys(c->scan_tmp.p, tb, d_uh.as<uint32_t>(), d_ban.as<uint32_t>(), (int)n_ban, 0, 32, st));
        c->kernel_launches += 10;
        for (int strand = 0; strand < 2; ++strand) {
            DevBuf& keys = strand ? ix->keys_rev : ix->keys_fwd;
            uint64_t& cnt = strand ? ix->n_rev : ix->n_fwd;
            if (!cnt) continue;
            TRY(d_keep.ensure(cnt));
            TRY(d_tmp.ensure(cnt * 8));
            ban_flag_kernel<<<(unsigned)div_up64(cnt, 256), 256, 0, st>>>(keys.as<uint64_t>(), cnt, d_ban.as<uint32_t>(), n_ban, d_keep.as<uint8_t>());
            size_t fb = 0;
            CU(cub::DeviceSelect::Flagged(nullptr, fb, keys.as<uint64_t>(), d_keep.as<uint8_t>(), d_tmp.as<uint64_t>(), c->total.as<uint64_t>(), (int)cnt, st));
            TRY(c->scan_tmp.ensure(fb));
            CU(cub::DeviceSelect::Flagged(c->scan_tmp.p, fb, keys.as<uint64_t>(), d_keep.as<uint8_t>(), d_tmp.as<uint64_t>(), c->total.as<uint64_t>(), (int)cnt, st));
            uint64_t kept = 0;
            CU(cudaMemcpyAsync(&kept, c->total.p, 8, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(keys.p, d_tmp.p, cnt * 8, cudaMemcpyDeviceToDevice, st));
            CU(cudaStreamSynchronize(st));
            cnt = kept;
            c->kernel_launches += 3;
        }
    }
    ix->stats[4] = ix->n_fwd; ix->stats[5] = ix->n_rev;
    CU(cudaStreamSynchronize(st));
    guard.ok = true;
    *out = ix;
    return B200_OK;
}

extern "C" int b200_map_batch(b200_ctx* c, const b200_index* ix, size_t n, const char* reads_buf, const uint64_t* reads_off,
                              int fastq_semantics, int type, int match, int mismatch, int gap, int want_cigar,
                              b200_mapping* out, char* cigar_buf, uint64_t* cigar_off, uint64_t cigar_cap) {
    if (!c || !ix || (n && (!reads_off || !out))) return fail(B200_E_ARG, "b200_map_batch: null argument");
    if (type < 0 || type > 2) return fail(B200_E_TYPE, "Unknown AlignmentType provided.");
    if (want_cigar && (!cigar_off || (!cigar_buf && cigar_cap))) return fail(B200_E_ARG, "CIGARs requested but no buffers given");
    if (n == 0) { if (cigar_off) cigar_off[0] = 0; return B200_OK; }
    if (n > 0x7ffffff0ull) return fail(B200_E_ARG, "batch too large");
    TRY(set_device(c));
    cudaStream_t st = c->stream;
    const uint64_t r0 = reads_off[0], r1 = reads_off[n];
    if (r1 > r0 && !reads_buf) return fail(B200_E_ARG, "null reads buffer");
    const uint32_t k = ix->k, w = ix->w;
    const uint32_t nr = (uint32_t)n;
    PhaseTrace tr;

    // ---- upload reads, Minimize (read strand flag true, :599/:712)
    std::vector<uint64_t> off(n + 1);
    for (size_t i = 0; i <= n; ++i) off[i] = reads_off[i] - r0;
    TRY(c->d_q.ensure(r1 - r0 + 64));
    if (r1 > r0) CU(cudaMemcpyAsync(c->d_q.p, reads_buf + r0, r1 - r0, cudaMemcpyHostToDevice, st));
    c->h2d_bytes += r1 - r0;
    if (!c->map_min_plan) {
        c->map_min_plan = new (std::nothrow) b200_min_plan();
        if (!c->map_min_plan) return fail(B200_E_NOMEM, "out of host memory");
        c->map_min_plan->ctx = c;
    }
    b200_min_plan* mp = c->map_min_plan;
    TRY(min_plan_build(mp, c, n, off.data(), k, w, nullptr));
    const uint64_t tot = mp->tuples;
    for (size_t i = 0; i < n; ++i) out[i] = b200_mapping{0, 1, 0, 0, 0, 0, 0, 0};
    if (cigar_off) for (size_t i = 0; i <= n; ++i) cigar_off[i] = 0;
    if (tot == 0) return B200_OK;
    if (tot > 0x7ffffff0ull) return fail(B200_E_ARG, "too many minimizers in one batch: split the reads");

    DevBuf* mb_ = c->map_buf;
    DevBuf &b_hash = mb_[0], &b_pos = mb_[1], &b_flag = mb_[2], &b_keep = mb_[3], &b_slot = mb_[4], &b_dhash = mb_[5], &b_dpos = mb_[6],
           &b_first = mb_[7], &b_doff = mb_[8], &b_cf = mb_[9], &b_cr = mb_[10], &b_lof = mb_[11], &b_lor = mb_[12], &b_mof = mb_[13],
           &b_mor = mb_[14], &b_mff = mb_[15], &b_mfs = mb_[16], &b_mrf = mb_[17], &b_mrs = mb_[18], &b_rof = mb_[19], &b_ror = mb_[20],
           &b_lis = mb_[21], &b_prev = mb_[22], &b_chf = mb_[23], &b_chr = mb_[24], &b_reg = mb_[25], &b_pm = mb_[26];
    TRY(b_hash.ensure(tot * 4)); TRY(b_pos.ensure(tot * 4)); TRY(b_flag.ensure(tot));
    TRY(b200_min_plan_run(mp, c->d_q.as<char>(), b_hash.as<uint32_t>(), b_pos.as<uint32_t>(), b_flag.as<uint8_t>(), st));
    if (tr.on) { cudaStreamSynchronize(st); tr.mark("upload+minimize"); }

    // ---- remove_duplicates (:28-45)
    TRY(b_keep.ensure(tot)); TRY(b_slot.ensure(tot * 4)); TRY(b_first.ensure((n + 1) * 4));
    CU(cudaMemsetAsync(b_first.p, 0xff, (n + 1) * 4, st));
    const unsigned tb = (unsigned)div_up64(tot, 256);
    dedup_flag_kernel<<<tb, 256, 0, st>>>(b_hash.as<uint32_t>(), b_pos.as<uint32_t>(), mp->d_out_off.as<uint64_t>(),
                                          mp->d_off.as<uint64_t>(), nr, k, w, b_keep.as<uint8_t>(), b_first.as<uint32_t>());
    dedup_sentinel_kernel<<<tb, 256, 0, st>>>(mp->d_out_off.as<uint64_t>(), nr, b_first.as<uint32_t>(), b_keep.as<uint8_t>());
    c->kernel_launches += 2;
    uint32_t n_min = 0;
    TRY(scan_flags(c, b_keep.as<uint8_t>(), tot, b_slot.as<uint32_t>(), &n_min, st));
    if (n_min == 0) return B200_OK;
    TRY(b_dhash.ensure((size_t)n_min * 4)); TRY(b_dpos.ensure((size_t)n_min * 4)); TRY(b_doff.ensure((n + 1) * 4));
    dedup_scatter_kernel<<<tb, 256, 0, st>>>(b_hash.as<uint32_t>(), b_pos.as<uint32_t>(), b_keep.as<uint8_t>(),
                                             b_slot.as<uint32_t>(), tot, b_dhash.as<uint32_t>(), b_dpos.as<uint32_t>());
    gather_offsets_kernel<<<(unsigned)div_up64(n + 1, 256), 256, 0, st>>>(b_slot.as<uint32_t>(), mp->d_out_off.as<uint64_t>(), nr,
                                                                          tot, n_min, b_doff.as<uint32_t>());
    c->kernel_launches += 2;
    if (tr.on) { cudaStreamSynchronize(st); tr.mark("dedup"); }

    // ---- seed lookup against both strands of the index
    TRY(b_cf.ensure((size_t)n_min * 4)); TRY(b_cr.ensure((size_t)n_min * 4));
    TRY(b_lof.ensure((size_t)n_min * 8)); TRY(b_lor.ensure((size_t)n_min * 8));
    TRY(b_mof.ensure((size_t)n_min * 4)); TRY(b_mor.ensure((size_t)n_min * 4));
    const unsigned mb = (unsigned)div_up64(n_min, 256);
    seed_count_kernel<<<mb, 256, 0, st>>>(b_dhash.as<uint32_t>(), n_min, ix->keys_fwd.as<uint64_t>(), ix->n_fwd,
                                          ix->keys_rev.as<uint64_t>(), ix->n_rev, fastq_semantics ? 0 : 1, b_cf.as<uint32_t>(),
                                          b_cr.as<uint32_t>(), b_lof.as<uint64_t>(), b_lor.as<uint64_t>());
    c->kernel_launches++;
    uint32_t n_mf = 0, n_mr = 0;
    TRY(scan_u32(c, b_cf.as<uint32_t>(), n_min, b_mof.as<uint32_t>(), &n_mf, st));
    TRY(scan_u32(c, b_cr.as<uint32_t>(), n_min, b_mor.as<uint32_t>(), &n_mr, st));
    TRY(b_mff.ensure(std::max<size_t>(n_mf, 1) * 4)); TRY(b_mfs.ensure(std::max<size_t>(n_mf, 1) * 4));
    TRY(b_mrf.ensure(std::max<size_t>(n_mr, 1) * 4)); TRY(b_mrs.ensure(std::max<size_t>(n_mr, 1) * 4));
    TRY(b_rof.ensure((n + 1) * 4)); TRY(b_ror.ensure((n + 1) * 4));
    seed_emit_kernel<<<mb, 256, 0, st>>>(b_dpos.as<uint32_t>(), n_min, ix->keys_fwd.as<uint64_t>(), b_cf.as<uint32_t>(),
                                         b_lof.as<uint64_t>(), b_mof.as<uint32_t>(), b_mff.as<uint32_t>(), b_mfs.as<uint32_t>());
    seed_emit_kernel<<<mb, 256, 0, st>>>(b_dpos.as<uint32_t>(), n_min, ix->keys_rev.as<uint64_t>(), b_cr.as<uint32_t>(),
                                         b_lor.as<uint64_t>(), b_mor.as<uint32_t>(), b_mrf.as<uint32_t>(), b_mrs.as<uint32_t>());
    const unsigned rb1 = (unsigned)div_up64(n + 1, 256);
    gather_offsets32_kernel<<<rb1, 256, 0, st>>>(b_mof.as<uint32_t>(), b_doff.as<uint32_t>(), nr, n_min, n_mf, b_rof.as<uint32_t>());
    gather_offsets32_kernel<<<rb1, 256, 0, st>>>(b_mor.as<uint32_t>(), b_doff.as<uint32_t>(), nr, n_min, n_mr, b_ror.as<uint32_t>());
    c->kernel_launches += 4;
    if (tr.on) { cudaStreamSynchronize(st); tr.mark("seeds"); }

    // ---- chaining (FindLIS) per strand, strand choice, region
    const size_t n_mmax = std::max<size_t>(std::max(n_mf, n_mr), 1);
    TRY(b_lis.ensure(n_mmax * 4)); TRY(b_prev.ensure(n_mmax * 4)); TRY(b_pm.ensure(n_mmax * 4));
    TRY(b_chf.ensure(n * sizeof(ChainResult))); TRY(b_chr.ensure(n * sizeof(ChainResult))); TRY(b_reg.ensure(n * sizeof(Region)));
    const unsigned cb = (unsigned)div_up64(n * 32, 128);
    chain_kernel<<<cb, 128, 0, st>>>(b_mff.as<uint32_t>(), b_mfs.as<uint32_t>(), b_rof.as<uint32_t>(), nr, b_lis.as<uint32_t>(),
                                     b_prev.as<int32_t>(), b_pm.as<uint32_t>(), b_chf.as<ChainResult>());
    chain_kernel<<<cb, 128, 0, st>>>(b_mrf.as<uint32_t>(), b_mrs.as<uint32_t>(), b_ror.as<uint32_t>(), nr, b_lis.as<uint32_t>(),
                                     b_prev.as<int32_t>(), b_pm.as<uint32_t>(), b_chr.as<ChainResult>());
    region_kernel<<<(unsigned)div_up64(n, 256), 256, 0, st>>>(b_chf.as<ChainResult>(), b_chr.as<ChainResult>(), nr, k,
                                                              mp->d_off.as<uint64_t>(), ix->ref_len, b_reg.as<Region>());
    c->kernel_launches += 3;
    std::vector<Region> reg(n);
    CU(cudaMemcpyAsync(reg.data(), b_reg.p, n * sizeof(Region), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    c->d2h_bytes += n * sizeof(Region);
    tr.mark("chain+region");

    // ---- Align on the found regions (:666-678): explicit sub-ranges of the read and reference buffers
    std::vector<uint32_t> mapped;
    for (size_t i = 0; i < n; ++i) if (reg[i].mapped) mapped.push_back((uint32_t)i);
    const size_t nm = mapped.size();
    if (nm == 0) return B200_OK;
    if (!c->map_plan) {
        c->map_plan = new (std::nothrow) b200_align_plan();
        if (!c->map_plan) return fail(B200_E_NOMEM, "out of host memory");
    }
    b200_align_plan* plan = c->map_plan;
    plan->reset();
    plan->ctx = c; plan->n = nm; plan->type = type; plan->sc = Scores{match, mismatch, gap}; plan->want_cigar = want_cigar != 0;
    plan->h_pairs.resize(nm);
    for (size_t j = 0; j < nm; ++j) {
        const Region& g = reg[mapped[j]];
        PairDesc& d = plan->h_pairs[j];
        d.q_off = off[mapped[j]] + g.q_begin;
        d.t_off = (g.fwd ? 0 : ix->ref_len) + g.t_begin;
        d.Q = g.q_end - g.q_begin + 1;
        d.T = g.t_end - g.t_begin + 1;
    }
    TRY(plan_finish(plan, c, !c->force_generic && short_scores_ok(plan->sc, type), true));
    tr.mark("align-plan");
    const uint64_t dev_cap = want_cigar ? plan->cigar_bound : 0;
    TRY(c->d_score.ensure(nm * 4)); TRY(c->d_tb.ensure(nm * 4));
    if (want_cigar) { TRY(c->d_cigar.ensure(dev_cap + 16)); TRY(c->d_cigar_off.ensure((nm + 1) * 8)); }
    TRY(b200_align_plan_run(plan, c->d_q.as<char>(), ix->d_ref.as<char>(), c->d_score.as<int32_t>(), c->d_tb.as<uint32_t>(),
                            want_cigar ? c->d_cigar.as<char>() : nullptr, want_cigar ? c->d_cigar_off.as<uint64_t>() : nullptr,
                            dev_cap, st));
    if (tr.on) { cudaStreamSynchronize(st); tr.mark("align-run"); }
    std::vector<int32_t> sc(nm);
    std::vector<uint32_t> tbg(nm);
    std::vector<uint64_t> coff(nm + 1, 0);
    CU(cudaMemcpyAsync(sc.data(), c->d_score.p, nm * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(tbg.data(), c->d_tb.p, nm * 4, cudaMemcpyDeviceToHost, st));
    if (want_cigar) CU(cudaMemcpyAsync(coff.data(), c->d_cigar_off.p, (nm + 1) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::vector<char> cig;
    if (want_cigar) {
        if (coff[nm] > cigar_cap) return fail(B200_E_CAP, "CIGAR buffer too small: need " + std::to_string(coff[nm]));
        cig.resize(coff[nm] + 1);
        if (coff[nm]) CU(cudaMemcpyAsync(cig.data(), c->d_cigar.p, coff[nm], cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    c->d2h_bytes += nm * 8 + (want_cigar ? (nm + 1) * 8 + coff[nm] : 0);
    uint64_t at = 0;
    size_t j = 0;
    for (size_t i = 0; i < n; ++i) {
        if (cigar_off) cigar_off[i] = at;
        if (j < nm && mapped[j] == i) {
            const Region& g = reg[i];
            out[i] = b200_mapping{1, g.fwd, g.q_begin, g.q_end, g.t_begin, g.t_end, sc[j], tbg[j]};
            if (want_cigar) {
                const uint64_t len = coff[j + 1] - coff[j];
                std::memcpy(cigar_buf + at, cig.data() + coff[j], len);
                at += len;
            }
            ++j;
        }
    }
    if (cigar_off) cigar_off[n] = at;
    tr.mark("d2h+assemble");
    return B200_OK;
}
