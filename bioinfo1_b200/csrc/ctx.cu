// ctx.cu -- contexts (streams + grow-only workspaces), options, counters and the error text of the C ABI
// declared in include/b200map.h. No CPU implementation of the hot path lives in this library: without a
// usable device every entry point returns B200_E_NOGPU.
#include <cstdio>
#include <cstdlib>
#include <new>
#include <utility>

#include "internal.hpp"

using namespace b200;

static thread_local std::string g_err;
int b200_fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

extern "C" const char* b200_last_error(void) { return g_err.c_str(); }
extern "C" int b200_version(void) { return 1; }
extern "C" int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void prof_begin(b200_ctx* c, cudaStream_t st, int kind) {
    if (!c->profile) return;
    b200_ctx::Span sp{kind, nullptr, nullptr};
    for (cudaEvent_t* e : {&sp.a, &sp.b}) {
        if (!c->event_pool.empty()) { *e = c->event_pool.back(); c->event_pool.pop_back(); }
        else cudaEventCreate(e);
    }
    cudaEventRecord(sp.a, st);
    c->spans.push_back(sp);
}
void prof_end(b200_ctx* c, cudaStream_t st) {
    if (!c->profile) return;
    cudaEventRecord(c->spans.back().b, st);
}
void prof_collect(b200_ctx* c, cudaStream_t st) {
    if (!c->profile) return;
    cudaStreamSynchronize(st);
    for (auto& sp : c->spans) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { c->kind_us[sp.kind] += ms * 1e3; c->kind_launches[sp.kind]++; }
        c->event_pool.push_back(sp.a); c->event_pool.push_back(sp.b);
    }
    c->spans.clear();
}

void tl_mark(b200_ctx* c, cudaStream_t st, const std::string& what) {
    static const bool on = std::getenv("B200_TRACE") && std::atoi(std::getenv("B200_TRACE")) >= 2;
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    c->timeline.push_back({what, e});
}
void tl_dump(b200_ctx* c) {
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

int set_device(const b200_ctx* c) {
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
    CU(cudaSetDevice(device));
    b200_ctx* c = new (std::nothrow) b200_ctx();
    if (!c) return fail(B200_E_NOMEM, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    {
        const cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            return fail(B200_E_CUDA, std::string("cudaStreamCreateWithFlags: ") + cudaGetErrorString(e));
        }
    }
    align_kernels_configure();
    // tuning runs: defaults of six options from the environment (listed in include/b200map.h)
    for (auto kv : {std::pair<const char*, int64_t*>{"B200_SUBST_LDS", &c->subst_lds}, {"B200_TAPER_TAIL", &c->taper_tail},
                    {"B200_CONCURRENT_WALK", &c->concurrent_walk}, {"B200_STREAM_FILL", &c->stream_fill}, {"B200_HOST_PACK", &c->host_pack}, {"B200_FILL_PIPE", &c->fill_pipe}})
        if (const char* e = std::getenv(kv.first)) *kv.second = std::atoll(e);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
        c->dir_budget_bytes = std::max<int64_t>(1ll << 30, (int64_t)(free_b / 3));
    *out = c;
    return B200_OK;
}

// Error paths of a run: work of the failed call must not still be in flight (or landing in the caller's
// host arrays) when the entry point returns, and the next call reuses these workspaces.
void ctx_sync_all_streams(b200_ctx* c) {
    cudaSetDevice(c->device);
    for (cudaStream_t s : {c->stream, c->aux_stream, c->emit_stream, c->pack_stream, c->copy_stream, c->fill_stream,
                           c->slot[0].walk_stream, c->slot[1].walk_stream})
        if (s) cudaStreamSynchronize(s);
    cudaGetLastError();
}

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
    if (c->fill_stream) cudaStreamDestroy(c->fill_stream);
    if (c->fill_event) cudaEventDestroy(c->fill_event);
    c->stream_state.release();
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
    for (HostBuf* b : {&c->h_q, &c->h_t, &c->h_off, &c->h_small, &c->h_out_small, &c->h_out_cigar, &c->h_qpk, &c->h_tpk,
                       &c->h_flags, &c->h_wave_cnt})
        b->release();
    delete c;
}

extern "C" int b200_ctx_set_option(b200_ctx* c, const char* key, int64_t value) {
    if (!c || !key) return fail(B200_E_ARG, "null argument");
    const std::string k(key);
    if (k == "dir_budget_bytes") c->dir_budget_bytes = std::max<int64_t>(value, 1 << 20);
    else if (k == "force_generic") c->force_generic = value;
    else if (k == "long16") c->long16 = value;
    else if (k == "subst_lds") c->subst_lds = value;
    else if (k == "fill_pipe") c->fill_pipe = value;
    else if (k == "overlap_waves") c->overlap_waves = value;
    else if (k == "concurrent_walk") c->concurrent_walk = value;
    else if (k == "chunk_pairs") c->chunk_pairs = value;
    else if (k == "taper_tail") c->taper_tail = value;
    else if (k == "host_pack") c->host_pack = value;
    else if (k == "stream_fill") c->stream_fill = value;
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
    // alu-pipe issue slots the fill kernels spend per register (= two cells), times ten: PRMT counts double
    if (k == "alu_slots_per_cell_pair_x10") return (c->subst_lds & 1) ? 30 : 50;
    static const char* kinds[4] = {"fill", "walk", "emit", "other"};
    for (int i = 0; i < 4; ++i) {
        if (k == std::string(kinds[i]) + "_ns") return (int64_t)(c->kind_us[i] * 1e3);
        if (k == std::string(kinds[i]) + "_launches") return c->kind_launches[i];
    }
    return -1;
}

// per-thread default contexts for the reference-shaped entry points
int default_ctx(int device, b200_ctx** out) {
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
