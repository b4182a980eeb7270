// align_host.cu -- the host-buffer entry points of the alignment ABI: upload pipeline over packed host
// buffers, and the reference-shaped pointer-array form on top of it.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <new>
#include <thread>
#include <vector>

#include "internal.hpp"

using namespace b200;

// Where the packed sequence bytes of a host-buffer call come from. The packed entry point's bytes are simply there;
// the pointer-array entry point gathers them into pinned staging with a few threads while earlier slices are already
// on their way to the device, and `wait` blocks until a prefix of the bytes is in place.
struct HostSource {
    const char* q = nullptr;
    const char* t = nullptr;
    std::function<int(uint64_t q_end, uint64_t t_end)> wait;   // optional; returns B200_OK or the producer's error
};

static int align_batch_host(b200_ctx* c, size_t n, const HostSource& src, const uint64_t* q_off, const uint64_t* t_off,
                            int type, int match, int mismatch, int gap, int32_t* score, uint32_t* target_begin,
                            char* cigar_buf, uint64_t* cigar_off, uint64_t cigar_cap) {
    const bool want_cigar = cigar_off != nullptr;
    TRY(set_device(c));
    // rebase offsets so that only the referenced byte ranges are copied
    const uint64_t q0 = q_off[0], q1 = q_off[n], t0 = t_off[0], t1 = t_off[n];
    if (((q1 > q0) && !src.q) || ((t1 > t0) && !src.t)) return fail(B200_E_ARG, "null sequence buffer");
    PhaseTrace tr;
    // Start the sequence upload first, in byte slices on a separate copy stream with an event after each slice: the
    // host-side planning below overlaps the DMA, and the plan's waves start as soon as the slice holding their last
    // byte has landed. The first half goes in kHead equal slices before the plan exists; the second half is cut where
    // the plan's waves end (uniform batches), so that no wave waits for bytes it does not need.
    TRY(c->d_q.ensure(q1 - q0 + 64));
    TRY(c->d_t.ensure(t1 - t0 + 64));
    cudaStream_t st = c->stream;
    constexpr int kHead = 8, kSlices = 16;
    if (!c->copy_stream) CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    const uint64_t qn = q1 - q0, tn = t1 - t0;
    struct Cut { uint64_t q_end, t_end; };
    std::vector<Cut> cuts;   // cuts[s] = bytes resident once copy_events[s] has fired
    auto upload_to = [&](uint64_t q_end, uint64_t t_end) -> int {
        const uint64_t qa = cuts.empty() ? 0 : cuts.back().q_end, ta = cuts.empty() ? 0 : cuts.back().t_end;
        q_end = std::min(std::max(q_end, qa), qn); t_end = std::min(std::max(t_end, ta), tn);
        if (src.wait) TRY(src.wait(q_end, t_end));
        if (q_end > qa) CU(cudaMemcpyAsync(c->d_q.as<char>() + qa, src.q + q0 + qa, q_end - qa, cudaMemcpyHostToDevice, c->copy_stream));
        if (t_end > ta) CU(cudaMemcpyAsync(c->d_t.as<char>() + ta, src.t + t0 + ta, t_end - ta, cudaMemcpyHostToDevice, c->copy_stream));
        if (c->copy_events.size() <= cuts.size()) {
            cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->copy_events.push_back(e);
        }
        CU(cudaEventRecord(c->copy_events[cuts.size()], c->copy_stream));
        cuts.push_back(Cut{q_end, t_end});
        return B200_OK;
    };
    tl_mark(c, c->copy_stream, "start");
    for (int s = 0; s < kHead; ++s) {
        TRY(upload_to(qn * (s + 1) / kSlices, tn * (s + 1) / kSlices));
        if (s == 0 || s == kHead - 1) tl_mark(c, c->copy_stream, "h2d" + std::to_string(s));
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
    // takes one 64-pair group per round): any other size leaves a partly empty last round in every chunk.
    // The batch ends with a few shrinking waves: a wave of one warp per SM sub-partition fills in a quarter of the
    // time a full round takes (its warps do not queue for the alu pipe), and each wave before it is sized so that its
    // fill is over when the next, smaller wave's bytes have landed (fill time / upload time of a wave is ~0.7) -- what
    // is left to do after the last byte of the upload is the smallest wave's work.
    size_t chunk_pairs = (size_t)c->chunk_pairs;
    std::vector<uint32_t> tail;
    if (chunk_pairs == 0) {
        size_t round_pairs = 0;
        TRY(align_short_round_pairs(c, type, &round_pairs));
        if (c->taper_tail == 1) {
            std::vector<uint32_t> rev;
            for (double g = (double)c->sm_count * 4; g < (double)round_pairs / 64 && rev.size() < 8; g *= 1.38) rev.push_back((uint32_t)g);
            tail.assign(rev.rbegin(), rev.rend());
        } else if (c->taper_tail == 2) {
            // only the last wave cut in two: whatever does not fill whole rounds, halved
            const uint64_t groups = div_up64(n, 64), round_groups = round_pairs / 64;
            uint64_t last = groups % round_groups;
            if (last == 0) last = round_groups;
            if (groups > last && last >= 128) { tail.push_back((uint32_t)(last - last / 2)); tail.push_back((uint32_t)(last / 2)); }
        }
        const size_t rounds_per_chunk = std::max<size_t>(1, div_up64(div_up64(n, round_pairs), 16));   // at most 16 chunks
        chunk_pairs = round_pairs * rounds_per_chunk;
    }
    TRY(plan_build(plan, c, n, q_off, t_off, true, false, type, match, mismatch, gap, want_cigar ? 1 : 0, chunk_pairs, &tail));
    tr.mark("plan");
    // second half of the upload, and which slice each wave has to wait for
    plan->wave_events.assign(plan->waves.size(), nullptr);
    if (plan->uniform) {
        for (size_t k = 0; k < plan->waves.size(); ++k) {
            const uint64_t last_pair = (uint64_t)plan->waves[k].first + plan->waves[k].count;   // exclusive
            const uint64_t qe = last_pair * plan->uQ, te = last_pair * plan->uT;                // bytes needed (exclusive)
            if (qe > cuts.back().q_end || te > cuts.back().t_end) {
                TRY(upload_to(qe, te));
                tl_mark(c, c->copy_stream, "h2dw" + std::to_string(k));
            }
            size_t need = 0;
            while (need + 1 < cuts.size() && (cuts[need].q_end < qe || cuts[need].t_end < te)) ++need;
            plan->wave_events[k] = c->copy_events[need];
        }
        if (cuts.back().q_end < qn || cuts.back().t_end < tn) TRY(upload_to(qn, tn));
    } else {
        for (int s = kHead; s < kSlices; ++s) TRY(upload_to(qn * (s + 1) / kSlices, tn * (s + 1) / kSlices));
        plan->wave_events.assign(plan->waves.size(), c->copy_events[cuts.size() - 1]);
    }
    tl_mark(c, c->copy_stream, "h2d-end");

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

extern "C" int b200_align_batch_packed(b200_ctx* c, size_t n, const char* q_buf, const uint64_t* q_off,
                                       const char* t_buf, const uint64_t* t_off, int type, int match,
                                       int mismatch, int gap, int32_t* score, uint32_t* target_begin,
                                       char* cigar_buf, uint64_t* cigar_off, uint64_t cigar_cap) {
    if (!c) return fail(B200_E_ARG, "null context");
    if (type < 0 || type > 2) return fail(B200_E_TYPE, "Unknown AlignmentType provided.");
    if (n && (!q_off || !t_off || !score)) return fail(B200_E_ARG, "null argument");
    if (cigar_off && !cigar_buf && cigar_cap) return fail(B200_E_ARG, "cigar_buf is null");
    if (n == 0) { if (cigar_off) cigar_off[0] = 0; return B200_OK; }
    HostSource src;
    src.q = q_buf; src.t = t_buf;
    const int rc = align_batch_host(c, n, src, q_off, t_off, type, match, mismatch, gap, score, target_begin, cigar_buf,
                                    cigar_off, cigar_cap);
    if (rc != B200_OK) {   // the upload of a call that failed half-way must not outlive it
        const std::string msg = b200_last_error();
        ctx_sync_all_streams(c);
        b200_fail(rc, msg);
    }
    return rc;
}

// ---- pointer arrays -> packed pinned staging, gathered by a few threads while the upload is already running ----
namespace {

unsigned host_threads() {
    static const unsigned n = [] {
        const char* e = std::getenv("B200_HOST_THREADS");
        unsigned v = e ? (unsigned)std::atoi(e) : 0;
        if (v == 0) v = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        return std::max(1u, std::min(v, 64u));
    }();
    return n;
}

// Gathers n (pointer, length) sequences of two arrays into two packed buffers with their n+1 offsets. Pairs are cut
// into blocks taken in order from a shared counter, so a prefix of the packed bytes is complete early and grows
// steadily; wait_pairs(p) returns once every pair below p is in place.
// memcpy for the gather: short sequences (a 150-base read) in 16-byte steps with an overlapping last step -- a
// library call per read costs more than the copy itself -- and long runs through memcpy.
inline void copy_bytes(char* d, const char* s, uint64_t n) {
    if (n >= 16 && n <= 512) {
        uint64_t k = 0;
        for (; k + 16 <= n; k += 16) _mm_storeu_si128(reinterpret_cast<__m128i*>(d + k), _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + k)));
        if (k < n) _mm_storeu_si128(reinterpret_cast<__m128i*>(d + n - 16), _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + n - 16)));
    } else if (n >= 4096) {
        // long runs: streaming stores (the staging buffer is written once and read by the DMA engine, never by this core:
        // a cached store would first read every destination line -- a third more memory traffic)
        while ((reinterpret_cast<uintptr_t>(d) & 15u) && n) { *d++ = *s++; --n; }
        uint64_t k = 0;
        for (; k + 64 <= n; k += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + k)), b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + k + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + k + 32)), e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + k + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + k), a); _mm_stream_si128(reinterpret_cast<__m128i*>(d + k + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + k + 32), c); _mm_stream_si128(reinterpret_cast<__m128i*>(d + k + 48), e);
        }
        if (k < n) std::memcpy(d + k, s + k, n - k);
        _mm_sfence();
    } else {
        std::memcpy(d, s, n);
    }
}

struct Gatherer {
    size_t n = 0;
    const char* const* src[2] = {nullptr, nullptr};
    const uint32_t* len[2] = {nullptr, nullptr};
    char* dst[2] = {nullptr, nullptr};
    uint64_t* off[2] = {nullptr, nullptr};
    static constexpr size_t kBlock = 2048;
    size_t n_blocks = 0;
    std::vector<std::thread> workers;
    std::vector<std::atomic<uint8_t>> done;
    std::atomic<size_t> next{0};
    std::atomic<int> bad{0};
    size_t prefix = 0;   // blocks [0, prefix) are known to be done (consumer side only)

    void offsets(unsigned T) {   // parallel prefix sum of the lengths
        std::vector<uint64_t> part(2 * (size_t)T, 0);
        auto range = [&](unsigned t) { return std::pair<size_t, size_t>(n * t / T, n * (t + 1) / T); };
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                auto [a, b] = range(t);
                for (int w = 0; w < 2; ++w) { uint64_t s = 0; for (size_t i = a; i < b; ++i) s += len[w][i]; part[2 * t + w] = s; }
            });
        for (auto& x : th) x.join();
        th.clear();
        uint64_t base[2] = {0, 0};
        std::vector<uint64_t> start(2 * (size_t)T);
        for (unsigned t = 0; t < T; ++t)
            for (int w = 0; w < 2; ++w) { start[2 * t + w] = base[w]; base[w] += part[2 * t + w]; }
        off[0][n] = base[0]; off[1][n] = base[1];
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                auto [a, b] = range(t);
                for (int w = 0; w < 2; ++w) { uint64_t s = start[2 * t + w]; for (size_t i = a; i < b; ++i) { off[w][i] = s; s += len[w][i]; } }
            });
        for (auto& x : th) x.join();
    }
    void start(unsigned T) {
        n_blocks = (n + kBlock - 1) / kBlock;
        done = std::vector<std::atomic<uint8_t>>(n_blocks);
        for (auto& d : done) d.store(0, std::memory_order_relaxed);
        for (unsigned t = 0; t < T; ++t)
            workers.emplace_back([this] {
                for (;;) {
                    const size_t b = next.fetch_add(1, std::memory_order_relaxed);
                    if (b >= n_blocks) return;
                    const size_t a = b * kBlock, e = std::min(n, a + kBlock);
                    for (int w = 0; w < 2; ++w)
                        for (size_t i = a; i < e;) {
                            const uint32_t l0 = len[w][i];
                            if (l0 == 0) { ++i; continue; }
                            const char* s0 = src[w][i];
                            if (!s0) { bad.store(1, std::memory_order_relaxed); ++i; continue; }
                            // sequences that already lie back to back in the caller's memory (one arena, a packed
                            // buffer behind the pointers) are copied as one run
                            uint64_t run = l0;
                            size_t j = i + 1;
                            while (j < e && (len[w][j] == 0 || src[w][j] == s0 + run)) { run += len[w][j]; ++j; }
                            copy_bytes(dst[w] + off[w][i], s0, run);
                            i = j;
                        }
                    done[b].store(1, std::memory_order_release);
                }
            });
    }
    int wait_pairs(size_t p) {   // every pair below p gathered
        const size_t need = std::min(n_blocks, (p + kBlock - 1) / kBlock);
        while (prefix < need) {
            while (!done[prefix].load(std::memory_order_acquire)) std::this_thread::yield();
            ++prefix;
        }
        return bad.load(std::memory_order_relaxed) ? fail(B200_E_ARG, "null sequence pointer") : B200_OK;
    }
    void join() { for (auto& w : workers) if (w.joinable()) w.join(); workers.clear(); }
    ~Gatherer() { join(); }
};

}  // namespace

extern "C" int b200_align_batch(int device, size_t n, const char* const* query, const uint32_t* query_len,
                                const char* const* target, const uint32_t* target_len, int type, int match,
                                int mismatch, int gap, int32_t* score, uint32_t* target_begin,
                                char* cigar_buf, uint64_t* cigar_off, uint64_t cigar_cap) {
    if (type < 0 || type > 2) return fail(B200_E_TYPE, "Unknown AlignmentType provided.");
    if (n && (!query || !query_len || !target || !target_len || !score)) return fail(B200_E_ARG, "null argument");
    if (cigar_off && !cigar_buf && cigar_cap) return fail(B200_E_ARG, "cigar_buf is null");
    b200_ctx* c = nullptr;
    TRY(default_ctx(device, &c));
    if (n == 0) { if (cigar_off) cigar_off[0] = 0; return B200_OK; }
    TRY(set_device(c));
    TRY(c->h_off.ensure(2 * (n + 1) * sizeof(uint64_t)));
    Gatherer g;
    g.n = n;
    g.src[0] = query; g.src[1] = target; g.len[0] = query_len; g.len[1] = target_len;
    g.off[0] = c->h_off.as<uint64_t>(); g.off[1] = g.off[0] + (n + 1);
    // small batches: one thread, no pipeline to feed
    const unsigned T = n >= 4 * Gatherer::kBlock ? host_threads() : 1u;
    PhaseTrace ptr_trace;
    g.offsets(T);
    ptr_trace.mark("ptr:offsets");
    TRY(c->h_q.ensure(g.off[0][n] + 1));
    TRY(c->h_t.ensure(g.off[1][n] + 1));
    g.dst[0] = c->h_q.as<char>(); g.dst[1] = c->h_t.as<char>();
    g.start(T);
    HostSource src;
    src.q = g.dst[0]; src.t = g.dst[1];
    const uint64_t* qo = g.off[0];
    const uint64_t* to = g.off[1];
    src.wait = [&g, qo, to, n](uint64_t q_end, uint64_t t_end) -> int {
        // the first pair whose bytes start at or beyond both ends: everything below it is needed
        const size_t pq = (size_t)(std::lower_bound(qo, qo + n + 1, q_end) - qo);
        const size_t pt = (size_t)(std::lower_bound(to, to + n + 1, t_end) - to);
        return g.wait_pairs(std::min(n, std::max(pq, pt)));
    };
    // The caller's result arrays are ordinary (pageable) memory: a device-to-host copy into them goes through the
    // driver's bounce buffers and blocks the calling thread each time, which serialises the wave pipeline. Large batches
    // land in pinned staging instead (downloads overlap the later waves) and are copied out with a few threads at the end.
    const bool stage_out = n >= 4 * Gatherer::kBlock;
    int32_t* s_score = score; uint32_t* s_tb = target_begin; uint64_t* s_off = cigar_off; char* s_cig = cigar_buf;
    uint64_t s_cap = cigar_cap;
    if (stage_out) {
        TRY(c->h_out_small.ensure(n * 8 + (n + 1) * 8 + 64));
        s_score = c->h_out_small.as<int32_t>();
        s_tb = target_begin ? reinterpret_cast<uint32_t*>(s_score + n) : nullptr;
        if (cigar_off) {
            s_off = reinterpret_cast<uint64_t*>(c->h_out_small.as<char>() + n * 8);
            s_cap = std::min<uint64_t>(cigar_cap, 2 * (qo[n] + to[n]) + 2 * n + 16);   // no CIGAR is longer than 2 (Q + T) + 2
            TRY(c->h_out_cigar.ensure(s_cap + 16));
            s_cig = c->h_out_cigar.as<char>();
        }
    }
    ptr_trace.mark("ptr:start-gather");
    int rc = align_batch_host(c, n, src, qo, to, type, match, mismatch, gap, s_score, s_tb, s_cig, s_off, s_cap);
    g.join();
    ptr_trace.mark("ptr:align");
    if (rc != B200_OK) {
        const std::string msg = b200_last_error();
        ctx_sync_all_streams(c);   // no copy may still read the staging buffers the next call refills
        b200_fail(rc, msg);
        return rc;
    }
    if (stage_out) {
        struct Part { char* d; const char* s; uint64_t n; };
        std::vector<Part> parts{{reinterpret_cast<char*>(score), reinterpret_cast<const char*>(s_score), n * 4}};
        if (target_begin) parts.push_back({reinterpret_cast<char*>(target_begin), reinterpret_cast<const char*>(s_tb), n * 4});
        if (cigar_off) {
            parts.push_back({reinterpret_cast<char*>(cigar_off), reinterpret_cast<const char*>(s_off), (n + 1) * 8});
            parts.push_back({cigar_buf, s_cig, s_off[n]});
        }
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; ++t)
            th.emplace_back([&, t] { for (const Part& p : parts) std::memcpy(p.d + p.n * t / T, p.s + p.n * t / T, p.n * (t + 1) / T - p.n * t / T); });
        for (const Part& p : parts) std::memcpy(p.d, p.s, p.n / T);
        for (auto& x : th) x.join();
        ptr_trace.mark("ptr:copy-out");
    }
    return rc;
}
