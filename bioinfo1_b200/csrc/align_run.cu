// align_run.cu -- runs a planned alignment batch: per wave 2-bit pack -> DP fill -> traceback walk, then
// CIGAR offsets (scan) and text; the repair pass for pairs that turn out not to be pure ACGT.
#include <algorithm>
#include <cstring>

#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "align_fill_generic.cuh"
#include "align_fill_long.cuh"
#include "align_fill_long16.cuh"
#include "align_fill_short.cuh"
#include "align_walk.cuh"
#include "internal.hpp"

using namespace b200;

// Kernels only share an SM when they agree on its shared-memory carve-out: the long-pair fills use no shared
// memory, the tile walkers 16 KB, and with the default preferences a walker CTA did not become resident until
// the fill running on the SM was over (measured: the "concurrent" walkers finished 2.9 ms after the fill; launched
// first, they kept the fill out instead). Same explicit preference on all of them.
void align_kernels_configure() {
    const int pct = 30;   // 68 KB: three fill CTAs with their 8 KB substitution tables + a 16 KB walker CTA
    cudaFuncSetAttribute(fill_long16_kernel<0, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fill_long16_kernel<1, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fill_long16_kernel<2, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fill_long16_kernel<0, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fill_long16_kernel<1, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(fill_long16_kernel<2, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
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

static int short_blocks_per_sm(int type, bool lds, int* per_sm) {
    *per_sm = 0;
#define OCC(TY, SB) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, fill_short_kernel<TY, SB>, kShortThreads, 0))
    if (lds) { if (type == 0) OCC(0, 1); else if (type == 1) OCC(1, 1); else OCC(2, 1); }
    else { if (type == 0) OCC(0, 0); else if (type == 1) OCC(1, 0); else OCC(2, 0); }
#undef OCC
    *per_sm = std::max(*per_sm, 1);
    return B200_OK;
}

// One ROUND of the thread-per-pair fill: every resident warp takes one 64-pair group.
int align_short_round_pairs(b200_ctx* c, int type, size_t* out) {
    int per_sm = 0;
    TRY(short_blocks_per_sm(type, c->subst_lds != 0, &per_sm));
    *out = (size_t)c->sm_count * per_sm * (kShortThreads / 32) * 64;
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
    TRY(short_blocks_per_sm(p->type, c->subst_lds != 0, &per_sm));
    const int n_blocks = (int)std::min<uint64_t>((uint64_t)c->sm_count * per_sm, div_up64(n_groups, kShortThreads / 32));
    const uint32_t bnd_cols = p->max_T_short + 4;
    TRY(rb.ws->bnd_short.ensure((size_t)n_blocks * (kShortThreads / 32) * bnd_cols * 32 * sizeof(uint32_t)));
    CU(cudaMemsetAsync(rb.ws->counter.p, 0, 64, rb.st));
    const ShortConsts K = make_short_consts(p->sc, p->type);
    prof_begin(c, rb.st, 0);
#define SHORTK(TY) SHORTK2(TY, 0)
#define SHORTK1(TY) SHORTK2(TY, 1)
#define SHORTK2(TY, SB) fill_short_kernel<TY, SB><<<n_blocks, kShortThreads, 0, rb.st>>>(                                         \
        c->qpk.as<uint32_t>(), c->tpk.as<uint32_t>(), p->d_pairs.as<PairDesc>(), p->d_work.as<uint32_t>() + wv.first,     \
        wv.count, p->d_groups.as<ShortGroup>() + wv.first_group, rb.ws->counter.as<uint32_t>(), c->flags.as<uint8_t>(), K, \
        rb.dirs, rb.ws->bnd_short.as<uint32_t>(), bnd_cols, rb.score, c->end_i.as<uint32_t>(), c->end_j.as<uint32_t>())
    if (c->subst_lds) { switch (p->type) { case 0: SHORTK1(0); break; case 1: SHORTK1(1); break; default: SHORTK1(2); break; } }
    else { switch (p->type) { case 0: SHORTK(0); break; case 1: SHORTK(1); break; default: SHORTK(2); break; } }
#undef SHORTK
#undef SHORTK1
#undef SHORTK2
    prof_end(c, rb.st);
    c->kernel_launches++;
    return B200_OK;
}

static int launch_fill_long(b200_align_plan* p, const Wave& wv, const RunBufs& rb) {
    b200_ctx* c = p->ctx;
    int per_sm = 0;
    if (p->long16 && c->subst_lds) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_long16_kernel<1, 1>, 128, 0));
    else if (p->long16) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_long16_kernel<1, 0>, 128, 0));
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
#define LONG16K(TY) if (c->subst_lds) { LONG16K2(TY, 1); } else { LONG16K2(TY, 0); }
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
        // a stripe or walker that gave up waiting leaves stale run counts behind: look at the stall flags with the same
        // read-back, before anything is emitted from them
        uint32_t* h_stall = c->h_small.as<uint32_t>() + 2;
        if (p->n_long)
            for (int k = 0; k < n_slots; ++k)
                CU(cudaMemcpyAsync(h_stall + k, c->slot[k].counter.as<uint32_t>() + 24, 4, cudaMemcpyDeviceToHost, es));
        CU(cudaStreamSynchronize(es));
        if (p->n_long && (h_stall[0] | h_stall[1]))
            return fail(B200_E_CUDA, "long-pair kernel: a stripe gave up waiting for its predecessor");
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
