// align_run.cu -- runs a planned alignment batch: per wave 2-bit pack -> DP fill -> traceback walk, then
// CIGAR offsets (scan) and text; the repair pass for pairs that turn out not to be pure ACGT.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "align_fill_generic.cuh"
#include "align_fill_long.cuh"
#include "align_fill_long16.cuh"
#include "align_fill_short.cuh"
#include "align_walk.cuh"
#include "internal.hpp"

using namespace b200;

// (subst_lds: bit 0 = the short-pair kernel, bit 1 = the long-pair kernel)
static inline int fill_variant(const b200_ctx* c) { return ((c->subst_lds & 1) ? 1 : 0) | (c->fill_pipe ? 2 : 0); }

// f.template operator()<TYPE, VARIANT>() for the run-time (type, variant)
template <class F>
static int dispatch_short(int type, int variant, F&& f) {
#define ROW(TY) switch (variant) { case 0: return f.template operator()<TY, 0>(); case 1: return f.template operator()<TY, 1>(); \
                                   case 2: return f.template operator()<TY, 2>(); default: return f.template operator()<TY, 3>(); }
    if (type == 0) { ROW(0) } else if (type == 1) { ROW(1) } else { ROW(2) }
#undef ROW
}

struct ShortOccupancy {
    int* per_sm;
    template <int TY, int SB> int operator()() const {
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, fill_short_kernel<TY, SB>, kShortThreads, 0));
        return B200_OK;
    }
};
static int short_blocks_per_sm(int type, int variant, int* per_sm) {
    *per_sm = 0;
    TRY(dispatch_short(type, variant, ShortOccupancy{per_sm}));
    *per_sm = std::max(*per_sm, 1);
    return B200_OK;
}

// One ROUND of the thread-per-pair fill: every resident warp takes one 64-pair group.
int align_short_round_pairs(b200_ctx* c, int type, size_t* out) {
    int per_sm = 0;
    TRY(short_blocks_per_sm(type, fill_variant(c), &per_sm));
    *out = (size_t)c->sm_count * per_sm * (kShortThreads / 32) * 64;
    return B200_OK;
}

// Inclusive scan of the per-pair CIGAR byte counts into 64-bit offsets (out[i] = base + len[0] + ... + len[i]), in three
// small launches: tile sums, scan of the tile sums, tile rescan. Our own kernels rather than a library scan because they
// run next to the persistent fill of the streaming mode: kernels only share an SM when they agree on its shared-memory
// carve-out, and a library kernel with kilobytes of shared memory does not become resident next to the fill -- it then
// waits at the head of the hardware queue and keeps the pack kernels the fill is waiting for from starting.
constexpr int kScanThreads = 256, kScanPerThread = 16, kScanTile = kScanThreads * kScanPerThread;

__device__ __forceinline__ uint64_t block_exclusive_scan_u64(uint64_t v, uint64_t* total_out) {
    __shared__ uint64_t warp_sum[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint64_t u = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    uint64_t before = 0, total = 0;
#pragma unroll
    for (int k = 0; k < kScanThreads / 32; ++k) { const uint64_t t = warp_sum[k]; if (k < warp) before += t; total += t; }
    __syncthreads();
    if (total_out) *total_out = total;
    return before + incl - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_tile_sums_kernel(const uint32_t* __restrict__ len, uint32_t n, uint64_t* __restrict__ tile_sum) {
    const uint32_t i0 = blockIdx.x * kScanTile + threadIdx.x * kScanPerThread;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) if (i0 + k < n) s += len[i0 + k];
    uint64_t total;
    block_exclusive_scan_u64(s, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// one CTA: tile sums -> exclusive tile bases (+ *base when given); a thread owns a run of consecutive tiles
__global__ void __launch_bounds__(kScanThreads)
scan_tile_bases_kernel(uint64_t* __restrict__ tile_sum, uint32_t n_tiles, const uint64_t* __restrict__ base) {
    const uint32_t per = (n_tiles + kScanThreads - 1) / kScanThreads, i0 = threadIdx.x * per;
    uint64_t s = 0;
    for (uint32_t k = 0; k < per; ++k) if (i0 + k < n_tiles) s += tile_sum[i0 + k];
    uint64_t run = block_exclusive_scan_u64(s, nullptr) + (base ? *base : 0);
    for (uint32_t k = 0; k < per; ++k)
        if (i0 + k < n_tiles) { const uint64_t v = tile_sum[i0 + k]; tile_sum[i0 + k] = run; run += v; }
}

__global__ void __launch_bounds__(kScanThreads)
scan_write_kernel(const uint32_t* __restrict__ len, uint32_t n, const uint64_t* __restrict__ tile_base, uint64_t* __restrict__ out) {
    const uint32_t i0 = blockIdx.x * kScanTile + threadIdx.x * kScanPerThread;
    uint32_t v[kScanPerThread];
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) { v[k] = i0 + k < n ? len[i0 + k] : 0u; s += v[k]; }
    uint64_t run = block_exclusive_scan_u64(s, nullptr) + tile_base[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) { run += v[k]; if (i0 + k < n) out[i0 + k] = run; }
}


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

struct ShortLaunch {
    b200_align_plan* p; const Wave& wv; const RunBufs& rb; int n_blocks; uint32_t bnd_cols; const ShortConsts& K; const ShortStream& sc;
    template <int TY, int SB> int operator()() const {
        b200_ctx* c = p->ctx;
        fill_short_kernel<TY, SB><<<n_blocks, kShortThreads, 0, rb.st>>>(
            c->qpk.as<uint32_t>(), c->tpk.as<uint32_t>(), p->d_pairs.as<PairDesc>(), p->d_work.as<uint32_t>() + wv.first,
            wv.count, p->d_groups.as<ShortGroup>() + wv.first_group, rb.ws->counter.as<uint32_t>(), c->flags.as<uint8_t>(), K,
            rb.dirs, rb.ws->bnd_short.as<uint32_t>(), bnd_cols, rb.score, c->end_i.as<uint32_t>(), c->end_j.as<uint32_t>(), sc);
        return B200_OK;
    }
};

// One-thread kernel that publishes the streaming watermark (stream order puts it after the slice's pack kernel).
__global__ void publish_watermark_kernel(uint32_t* watermark, uint32_t pairs_ready) {
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(watermark), "r"(pairs_ready) : "memory");
}

// `sctl` (streaming mode, see ShortStream): the launch serves every wave of the run and leaves two CTA slots per SM
// free (with one, measured, the pipeline stalls: a 256-thread scan CTA does not fit 8 K registers), so that the pack, traceback, scan and emit kernels it waits for and feeds can always become resident next to it
// (a fill that owned every slot would spin on a watermark nobody can raise).
static int launch_fill_short(b200_align_plan* p, const Wave& wv, const RunBufs& rb, const ShortStream* sctl = nullptr) {
    b200_ctx* c = p->ctx;
    const uint32_t n_groups = (wv.count + 63) / 64;
    int per_sm = 0;
    TRY(short_blocks_per_sm(p->type, fill_variant(c), &per_sm));
    if (sctl) per_sm = std::max(1, per_sm - 2);   // two 64-thread slots = 16 K registers stay free for the small kernels
    const int n_blocks = (int)std::min<uint64_t>((uint64_t)c->sm_count * per_sm, div_up64(n_groups, kShortThreads / 32));
    const uint32_t bnd_cols = p->max_T_short + 4;
    TRY(rb.ws->bnd_short.ensure((size_t)n_blocks * (kShortThreads / 32) * bnd_cols * 32 * sizeof(uint32_t)));
    CU(cudaMemsetAsync(rb.ws->counter.p, 0, 64, rb.st));
    const ShortConsts K = make_short_consts(p->sc, p->type);
    ShortStream sc{};
    if (sctl) sc = *sctl;
    prof_begin(c, rb.st, 0);
    TRY(dispatch_short(p->type, fill_variant(c), ShortLaunch{p, wv, rb, n_blocks, bnd_cols, K, sc}));
    prof_end(c, rb.st);
    c->kernel_launches++;
    return B200_OK;
}

static int launch_fill_long(b200_align_plan* p, const Wave& wv, const RunBufs& rb) {
    b200_ctx* c = p->ctx;
    int per_sm = 0;
    const int v16 = (c->subst_lds & 2) ? 1 : 0;   // variant of the packed long-pair kernel: substitution term by PRMT / shared table
    if (p->long16) {
        if (v16 == 0) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_long16_kernel<1, 0>, 128, 0));
        else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_long16_kernel<1, 1>, 128, 0));
    }
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
#define LONG16K(TY) if (v16 == 0) { LONG16K2(TY, 0); } else { LONG16K2(TY, 1); }
#define LONG16K2(TY, SB)                                                                                               \
    if (cw) {   /* pairs without inner cells have nothing to wait for: result and ready flag now */                   \
        finalize_long_kernel<TY><<<(unsigned)div_up64(wv.count, 128), 128, 0, rb.st>>>(p->d_pairs.as<PairDesc>(), d_work, \
            wv.count, d_task_off, c->flags.as<uint8_t>(), ws.stripe_res.as<StripeResult>(), K16.init, rb.score,        \
            c->end_i.as<uint32_t>(), c->end_j.as<uint32_t>(), 1, d_ready);                                             \
        CU(cudaEventRecord(ws.pre_event, rb.st));                                                                      \
        CU(cudaStreamWaitEvent(ws.walk_stream, ws.pre_event, 0));                                                      \
    }                                                                                                                  \
    fill_long16_kernel<TY, SB><<<n_blocks, 128, 0, rb.st>>>(c->qpk.as<uint32_t>(), c->tpk.as<uint32_t>(), p->d_pairs.as<PairDesc>(), \
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
#undef LONG16K2
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

extern "C" int b200_align_plan_run(b200_align_plan* p, const char* d_q_buf, const char* d_t_buf,
                                   int32_t* d_score, uint32_t* d_target_begin, char* d_cigar,
                                   uint64_t* d_cigar_off, uint64_t cigar_cap, void* stream) {
    return plan_run_impl(p, d_q_buf, d_t_buf, d_score, d_target_begin, d_cigar, d_cigar_off, cigar_cap, stream, nullptr);
}

static int plan_run_body(b200_align_plan* p, const char* d_q_buf, const char* d_t_buf, int32_t* d_score,
                         uint32_t* d_target_begin, char* d_cigar, uint64_t* d_cigar_off, uint64_t cigar_cap,
                         void* stream, HostOut* ho);

// One cleanup path: whatever fails after the first launch, nothing of the run is left in flight when the
// entry point returns (async downloads into the caller's arrays included), so the next call may reuse the
// context's workspaces.
int plan_run_impl(b200_align_plan* p, const char* d_q_buf, const char* d_t_buf, int32_t* d_score,
                  uint32_t* d_target_begin, char* d_cigar, uint64_t* d_cigar_off, uint64_t cigar_cap,
                  void* stream, HostOut* ho) {
    const int rc = plan_run_body(p, d_q_buf, d_t_buf, d_score, d_target_begin, d_cigar, d_cigar_off, cigar_cap, stream, ho);
    if (rc != B200_OK && p && p->ctx) {
        const std::string msg = b200_last_error();
        ctx_sync_all_streams(p->ctx);
        if (stream) cudaStreamSynchronize((cudaStream_t)stream);
        cudaGetLastError();
        b200_fail(rc, msg);
    }
    return rc;
}

static int plan_run_body(b200_align_plan* p, const char* d_q_buf, const char* d_t_buf, int32_t* d_score,
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
        if (!p->host_packed) CU(cudaMemsetAsync(c->flags.p, 0, n + 4, st));   // (host-packed runs upload the flags)
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
    if (!p->host_packed) CU(cudaMemsetAsync(c->wave_flagged.p, 0, n_waves * 4 + 16, st));   // (... and the per-wave counts)
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

    // Streaming fill (uniform short batches through the host pipeline, see ShortStream in align_fill_short.cuh): one
    // persistent fill launch for all waves, fed by a watermark; the per-wave kernels left are pack + publish (pack
    // stream) and the traceback, which waits on the device for its wave's groups. All waves' direction matrices are
    // live at once, so it needs the whole batch to fit the budget.
    uint64_t all_dir_words = 0;
    for (const Wave& w : p->waves) all_dir_words += (w.dir_words + 3) & ~3ull;
    const bool streaming = piped && pack_ahead && c->stream_fill && p->want_cigar && n_waves <= 32 &&
                           all_dir_words <= wave_budget_words(c) && p->n_short == n;
    ShortStream sctl{};
    uint32_t* d_stream = nullptr;   // [0] watermark, [1 .. n_waves] groups done per wave
    const bool check_stall = p->n_long != 0 || streaming;   // kernels that wait on the device raise a flag instead of hanging
    if (streaming) {
        if (!c->fill_stream) CU(cudaStreamCreateWithFlags(&c->fill_stream, cudaStreamNonBlocking));
        if (!c->fill_event) CU(cudaEventCreateWithFlags(&c->fill_event, cudaEventDisableTiming));
        WaveSlot& ws0 = c->slot[0];
        TRY(ws0.dirs.ensure(std::max<uint64_t>(all_dir_words, 4) * 4 + 64));
        TRY(c->stream_state.ensure(64 * 4));
        TRY(c->scan_tmp.ensure((size_t)div_up64(n, kScanTile) * 8 + 64));   // no allocation once the fill is spinning
        d_stream = c->stream_state.as<uint32_t>();
        CU(cudaMemsetAsync(d_stream, 0, 64 * 4, st));
        CU(cudaMemsetAsync(ws0.counter.as<uint32_t>() + 24, 0, 4, st));
        CU(cudaMemsetAsync(c->slot[1].counter.as<uint32_t>() + 24, 0, 4, st));
        sctl.watermark = d_stream; sctl.wave_done = d_stream + 1; sctl.stall_flag = ws0.counter.as<uint32_t>() + 24;
        sctl.n_waves = (uint32_t)n_waves;
        uint64_t base = 0;
        for (size_t k = 0; k < n_waves; ++k) {
            sctl.group_start[k] = p->waves[k].first / 64;
            sctl.dir_base[k] = base;
            base += (p->waves[k].dir_words + 3) & ~3ull;
        }
        sctl.group_start[n_waves] = (uint32_t)div_up64(n, 64);
        // the fill goes first (it must own its share of every SM before the small kernels come), on a stream of its own;
        // every other stream of the run starts after the counters have been zeroed
        CU(cudaEventRecord(c->fork_event, st));
        CU(cudaStreamWaitEvent(c->fill_stream, c->fork_event, 0));
        CU(cudaStreamWaitEvent(c->pack_stream, c->fork_event, 0));
        CU(cudaStreamWaitEvent(c->aux_stream, c->fork_event, 0));
        Wave all{kClassShort, 0, (uint32_t)n, 0, all_dir_words};
        RunBufs rbf{reinterpret_cast<const uint8_t*>(d_q_buf), reinterpret_cast<const uint8_t*>(d_t_buf), ws0.dirs.as<uint32_t>(),
                    d_score, c->fill_stream, &ws0};
        tl_mark(c, c->fill_stream, "fill-start");
        TRY(launch_fill_short(p, all, rbf, &sctl));
        tl_mark(c, c->fill_stream, "fill-end");
        CU(cudaEventRecord(c->fill_event, c->fill_stream));
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
        if (p->host_packed) {
            // the wave's 2-bit words and flags were packed by the host's gather pass and came with its upload slice
        } else if (wv.klass != kClassGeneric) {
            const uint32_t wpp = std::max(1u, div_up(wv.klass == kClassLong ? std::max(p->max_Q, p->max_T)
                                                                           : std::max(p->max_Q_short, p->max_T_short), 16));
            dim3 grid((unsigned)div_up64((uint64_t)wv.count * wpp, 128), 2);   // (128-thread CTAs: they fit one free fill slot)
            pack_kernel<<<grid, 128, 0, pst>>>(rb.dq, rb.dt, p->d_pairs.as<PairDesc>(), work, wv.count, wpp,
                                               c->flags.as<uint8_t>(), c->qpk.as<uint32_t>(), c->tpk.as<uint32_t>(), d_nflag);
        } else {
            classify_kernel<<<(unsigned)div_up64((uint64_t)wv.count * 32, 256), 256, 0, pst>>>(
                rb.dq, rb.dt, p->d_pairs.as<PairDesc>(), work, wv.count, c->flags.as<uint8_t>());
        }
        prof_end(c, pst);
        if (!p->host_packed) c->kernel_launches++;
        tl_mark(c, pst, "pack" + std::to_string(k));
        if (streaming) {
            publish_watermark_kernel<<<1, 1, 0, pst>>>(d_stream, wv.first + wv.count);
            c->kernel_launches++;
        }
        if (pack_ahead) {
            CU(cudaEventRecord(c->pack_done[k], pst));
            CU(cudaStreamWaitEvent(wst, c->pack_done[k], 0));
        }
        rb.dirs = p->want_cigar ? ws.dirs.as<uint32_t>() : nullptr;
        if (streaming) rb.dirs = c->slot[0].dirs.as<uint32_t>() + sctl.dir_base[k];
        // in a wave of a 2-bit class a non-zero flag means "not this wave's pair"; in a generic wave the flags only
        // describe the content (classify_kernel) and every pair is the wave's own
        const uint8_t* skip = wv.klass != kClassGeneric ? c->flags.as<uint8_t>() : nullptr;
        if (streaming) {}   // the persistent launch above fills every wave
        else if (wv.klass == kClassShort) TRY(launch_fill_short(p, wv, rb));
        else if (wv.klass != kClassGeneric) TRY(launch_fill_long(p, wv, rb));
        else TRY(launch_fill_generic(p, work, wv.count, rb));
        tl_mark(c, wst, "fill" + std::to_string(k));
        if (ws.walk_inflight) {
            // the wave's pairs are being walked next to the fill; the wave's stream waits for the walkers so that
            // everything after it sees every pair walked
            ws.walk_inflight = false;
            CU(cudaStreamWaitEvent(wst, ws.walk_event, 0));
            tl_mark(c, wst, "cwalk" + std::to_string(k));
        } else if (streaming) {
            wave_gate_kernel<<<1, 1, 0, wst>>>(d_stream + 1 + k, (wv.count + 63) / 64, c->slot[0].counter.as<uint32_t>() + 24);
            c->kernel_launches++;
            tl_mark(c, wst, "gate" + std::to_string(k));
            TRY(launch_walk(p, wv.klass, work, wv.count, rb, skip));
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
    if (streaming) CU(cudaStreamWaitEvent(st, c->fill_event, 0));

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
        const uint32_t cnt = b - a, n_tiles = (uint32_t)div_up64(cnt, kScanTile);
        TRY(c->scan_tmp.ensure((size_t)n_tiles * 8 + 64));
        if (a == 0) CU(cudaMemsetAsync(d_cigar_off, 0, sizeof(uint64_t), es));
        prof_begin(c, es, 3);
        scan_tile_sums_kernel<<<n_tiles, kScanThreads, 0, es>>>(c->cigar_len.as<uint32_t>() + a, cnt, c->scan_tmp.as<uint64_t>());
        scan_tile_bases_kernel<<<1, kScanThreads, 0, es>>>(c->scan_tmp.as<uint64_t>(), n_tiles, a ? d_cigar_off + a : nullptr);
        scan_write_kernel<<<n_tiles, kScanThreads, 0, es>>>(c->cigar_len.as<uint32_t>() + a, cnt, c->scan_tmp.as<uint64_t>(), d_cigar_off + a + 1);
        prof_end(c, es);
        c->kernel_launches += 3;
        CU(cudaMemcpyAsync(h_total, d_cigar_off + b, sizeof(uint64_t), cudaMemcpyDeviceToHost, es));
        TRY(queue_flag_counts(fk0, fk1, es));
        // a stripe or walker that gave up waiting leaves stale run counts behind: look at the stall flags with the same
        // read-back, before anything is emitted from them
        uint32_t* h_stall = c->h_small.as<uint32_t>() + 2;
        if (check_stall)
            for (int k = 0; k < n_slots; ++k)
                CU(cudaMemcpyAsync(h_stall + k, c->slot[k].counter.as<uint32_t>() + 24, 4, cudaMemcpyDeviceToHost, es));
        CU(cudaStreamSynchronize(es));
        if (check_stall && (h_stall[0] | h_stall[1]) && streaming && std::getenv("B200_TRACE")) {
            // diagnostic: where the streaming pipeline stood when a warp gave up
            cudaDeviceSynchronize();
            uint32_t hs[64] = {0}, cnt[32] = {0};
            cudaMemcpy(hs, d_stream, sizeof hs, cudaMemcpyDeviceToHost);
            cudaMemcpy(cnt, c->slot[0].counter.p, sizeof cnt, cudaMemcpyDeviceToHost);
            std::fprintf(stderr, "[b200 stall] watermark=%u groups_taken=%u waves=%zu:", hs[0], cnt[0], n_waves);
            for (size_t k = 0; k < n_waves; ++k) std::fprintf(stderr, " w%zu[first %u count %u done %u/%u]", k, p->waves[k].first, p->waves[k].count, hs[1 + k], (p->waves[k].count + 63) / 64);
            std::fprintf(stderr, "\n");
            tl_dump(c);
        }
        if (check_stall && (h_stall[0] | h_stall[1]))
            return fail(B200_E_CUDA, "fill kernel: a warp gave up waiting (for the stripe above it, or for the upload watermark)");
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
    if (check_stall) {
        uint32_t stalled[2] = {0, 0};
        for (int k = 0; k < n_slots; ++k)
            CU(cudaMemcpyAsync(&stalled[k], c->slot[k].counter.as<uint32_t>() + 24, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (stalled[0] | stalled[1])
            return fail(B200_E_CUDA, "fill kernel: a warp gave up waiting (for the stripe above it, or for the upload watermark)");
    }
    prof_collect(c, st);
    if (!ho) tl_dump(c);
    return B200_OK;
}

// Kernels only share an SM when they agree on its shared-memory carve-out: the long-pair fills use no shared
// memory, the tile walkers 16 KB, and with the default preferences a walker CTA did not become resident until
// the fill running on the SM was over (measured: the "concurrent" walkers finished 2.9 ms after the fill; launched
// first, they kept the fill out instead). Same explicit preference on all of them.
void align_kernels_configure() {
    const int pct = 30;   // 68 KB: three fill CTAs with their 8 KB substitution tables + a 16 KB walker CTA
#define K3ATTR(SB) cudaFuncSetAttribute(fill_long16_kernel<0, SB>, cudaFuncAttributePreferredSharedMemoryCarveout, pct); \
                   cudaFuncSetAttribute(fill_long16_kernel<1, SB>, cudaFuncAttributePreferredSharedMemoryCarveout, pct); \
                   cudaFuncSetAttribute(fill_long16_kernel<2, SB>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    K3ATTR(0) K3ATTR(1)
#undef K3ATTR
    cudaFuncSetAttribute(fill_long_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fill_long_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fill_long_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(walk_tile_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(walk_tile_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(walk_tile_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(walk_tile_wait_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(walk_tile_wait_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(walk_tile_wait_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    // CUDA loads kernels lazily, and loading one may have to wait for the device to go idle. The streaming mode keeps a
    // persistent fill kernel spinning on a watermark that only later launches (pack, publish, gate, traceback, scan, emit)
    // can raise: a first-time launch among them would wait for the fill, and the fill for it. Everything the pipeline
    // launches is therefore loaded here, once per context, before anything runs.
    cudaFuncAttributes fa;
#define PRELOAD(K) cudaFuncGetAttributes(&fa, K)
#define PRELOAD_SHORT(SB) PRELOAD((fill_short_kernel<0, SB>)); PRELOAD((fill_short_kernel<1, SB>)); PRELOAD((fill_short_kernel<2, SB>));
    PRELOAD_SHORT(0) PRELOAD_SHORT(1) PRELOAD_SHORT(2) PRELOAD_SHORT(3)
#undef PRELOAD_SHORT
    PRELOAD(pack_kernel); PRELOAD(classify_kernel); PRELOAD(publish_watermark_kernel); PRELOAD(wave_gate_kernel);
    PRELOAD(walk_kernel<0>); PRELOAD(walk_kernel<1>); PRELOAD(walk_kernel<2>);
    PRELOAD(target_begin_kernel); PRELOAD(scan_tile_sums_kernel); PRELOAD(scan_tile_bases_kernel); PRELOAD(scan_write_kernel);
    PRELOAD(emit_kernel); PRELOAD(emit_warp_kernel);
#undef PRELOAD
    cudaGetLastError();
}

// Kernel variant of a context: bit 0 substitution term from the shared-memory table, bit 1 software-pipelined columns.
