// minimize_api.cu -- minimizer plans (host-built tile records) and the MinimizeBatch entry points.
#include <algorithm>
#include <cstring>
#include <new>

#include "internal.hpp"
#include "minimize.cuh"

using namespace b200;

extern "C" uint64_t b200_minimize_count(uint32_t len, uint32_t k, uint32_t w) {
    if (len < k || w == 0) return 0;
    const uint64_t n = (uint64_t)len - k + 1;
    const uint64_t full = n >= w ? n - w + 1 : 0;
    const uint64_t tail = n < (uint64_t)w - 1 ? n : (uint64_t)w - 1;
    return (uint64_t)(w - 1) + full + tail;
}

extern "C" void b200_min_plan_destroy(b200_min_plan* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    p->d_off.release(); p->d_out_off.release(); p->d_fwd.release(); p->d_tiles.release();
    delete p;
}
extern "C" uint64_t b200_min_plan_tuples(const b200_min_plan* p) { return p ? p->tuples : 0; }
extern "C" const uint64_t* b200_min_plan_out_off(const b200_min_plan* p) { return p ? p->out_off.data() : nullptr; }

// Fills `p` (fresh or recycled: its device buffers keep their capacity) for a batch.
int min_plan_build(b200_min_plan* p, b200_ctx* ctx, size_t n, const uint64_t* off, uint32_t k, uint32_t w,
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
    // (per warp: its share of the slice plus the overlap its windows need, see the kernel)
    const uint64_t nx = (uint64_t)kMinWarpTuples + 2ull * w + 1;
    const uint64_t nwords = (15 + nx + k - 1 + 15) / 16 + 3;
    p->warp_words = (uint32_t)nwords;
    p->smem_bytes = (size_t)(nwords * 4) * (kMinThreads / 32);
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
#define MINK2(WW, KK)                                                                                                    \
        if (p->smem_bytes > 48 * 1024)                                                                                   \
            CU(cudaFuncSetAttribute(minimize_kernel<WW, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes)); \
        minimize_kernel<WW, KK><<<(unsigned)p->n_tiles, kMinThreads, p->smem_bytes, st>>>(                               \
            reinterpret_cast<const uint8_t*>(d_buf), p->d_tiles.as<MinTile>(), p->k, p->w, p->buf_bytes, p->warp_words,  \
            d_hash, d_pos, d_flag);
#define MINK(WW)                                                                                                         \
    case WW:                                                                                                             \
        if (p->k >= 16) { MINK2(WW, true) } else { MINK2(WW, false) }                                                    \
        break;
    switch (W) { MINK(1) MINK(2) MINK(3) MINK(4) MINK(5) MINK(6) MINK(7) MINK(8) default: MINK(0) }
#undef MINK
#undef MINK2
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
