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

#include "host_pack.hpp"
#include "internal.hpp"

using namespace b200;

// Where the packed sequence bytes of a host-buffer call come from. The packed entry point's bytes are simply there;
// the pointer-array entry point gathers them into pinned staging with a few threads while earlier slices are already
// on their way to the device, and `wait` blocks until a prefix of the bytes is in place.
struct HostSource {
    const char* q = nullptr;
    const char* t = nullptr;
    std::function<int(uint64_t q_end, uint64_t t_end)> wait;   // optional; returns B200_OK or the producer's error
    // Host-packed mode (uniform short batches through the pointer-array entry point): the producer does not copy the
    // bytes, it packs them to 2 bits (host_pack.hpp). What goes to the device per slice is then the 2-bit words of whole
    // pairs (straight into the context's packed buffers), one flag byte per pair, and the raw bytes of the few pairs that
    // are not pure ACGT (the byte-compare kernel of the repair pass reads those).
    bool packed = false;
    uint32_t Q = 0, T = 0;                      // the uniform lengths
    const uint32_t* qpk = nullptr;              // n x (Q/16 + 2) words
    const uint32_t* tpk = nullptr;              // n x (T/16 + 2) words
    const uint8_t* flags = nullptr;             // n bytes
    const char* const* q_ptr = nullptr;         // the caller's sequences (raw bytes of flagged pairs are uploaded from them)
    const char* const* t_ptr = nullptr;
};

// The plan of a host-buffer call (recycled in the context). Uniform short batches are cut into chunks of whole ROUNDS
// of the thread-per-pair kernel (every resident warp takes one 64-pair group per round): any other size leaves a partly
// empty last round in every chunk. (`taper_tail`: two measured alternatives for the end of the batch)
static int host_plan_build(b200_ctx* c, size_t n, const uint64_t* q_off, const uint64_t* t_off, int type, int match,
                           int mismatch, int gap, bool want_cigar) {
    if (!c->host_plan) {
        c->host_plan = new (std::nothrow) b200_align_plan();
        if (!c->host_plan) return fail(B200_E_NOMEM, "out of host memory");
        c->host_plan->ctx = c;
    }
    b200_align_plan* plan = c->host_plan;   // recycled: its device and host buffers keep their capacity
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
    return plan_build(plan, c, n, q_off, t_off, true, false, type, match, mismatch, gap, want_cigar ? 1 : 0, chunk_pairs, &tail);
}

// `plan_ready`: the context's host plan has already been built for exactly this batch (the pointer-array entry point
// plans before it decides how to gather).
static int align_batch_host(b200_ctx* c, size_t n, const HostSource& src, const uint64_t* q_off, const uint64_t* t_off,
                            int type, int match, int mismatch, int gap, int32_t* score, uint32_t* target_begin,
                            char* cigar_buf, uint64_t* cigar_off, uint64_t cigar_cap, bool plan_ready = false) {
    const bool want_cigar = cigar_off != nullptr;
    TRY(set_device(c));
    // rebase offsets so that only the referenced byte ranges are copied
    const uint64_t q0 = q_off[0], q1 = q_off[n], t0 = t_off[0], t1 = t_off[n];
    if (!src.packed && (((q1 > q0) && !src.q) || ((t1 > t0) && !src.t))) return fail(B200_E_ARG, "null sequence buffer");
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
    const uint32_t wq = src.Q / 16 + 2, wt = src.T / 16 + 2;   // packed words per sequence (host-packed mode)
    if (src.packed) {   // the run must find these buffers at their final size: it would otherwise reallocate them after the upload
        TRY(c->qpk.ensure(((uint64_t)n * wq + src.Q / 16 + 72) * 4));
        TRY(c->tpk.ensure(((uint64_t)n * wt + src.T / 16 + 72) * 4));
        TRY(c->flags.ensure(n + 8));
        TRY(c->wave_flagged.ensure(64 * 4 + 16));
        TRY(c->h_wave_cnt.ensure(64 * 4));
    }
    uint64_t pairs_up = 0;   // host-packed mode: pairs uploaded so far
    auto upload_to = [&](uint64_t q_end, uint64_t t_end) -> int {
        const uint64_t qa = cuts.empty() ? 0 : cuts.back().q_end, ta = cuts.empty() ? 0 : cuts.back().t_end;
        q_end = std::min(std::max(q_end, qa), qn); t_end = std::min(std::max(t_end, ta), tn);
        if (src.packed) {
            // whole pairs only: the slice ends at the last pair that lies below both byte marks
            uint64_t pe = std::min<uint64_t>(n, std::min(src.Q ? q_end / src.Q : n, src.T ? t_end / src.T : n));
            pe = std::max(pe, pairs_up);
            q_end = pe * src.Q; t_end = pe * src.T;
            if (src.wait) TRY(src.wait(q_end, t_end));
            if (pe > pairs_up) {
                const uint64_t a = pairs_up, cnt = pe - a;
                CU(cudaMemcpyAsync(c->qpk.as<uint32_t>() + a * wq, src.qpk + a * wq, cnt * wq * 4, cudaMemcpyHostToDevice, c->copy_stream));
                CU(cudaMemcpyAsync(c->tpk.as<uint32_t>() + a * wt, src.tpk + a * wt, cnt * wt * 4, cudaMemcpyHostToDevice, c->copy_stream));
                CU(cudaMemcpyAsync(c->flags.as<uint8_t>() + a, src.flags + a, cnt, cudaMemcpyHostToDevice, c->copy_stream));
                c->h2d_bytes += cnt * (wq + wt) * 4 + cnt;
                for (uint64_t i = a; i < pe; ++i) {   // rare: the raw bytes of a pair the 2-bit kernels cannot take
                    if (!src.flags[i]) continue;
                    if (src.Q) CU(cudaMemcpyAsync(c->d_q.as<char>() + i * src.Q, src.q_ptr[i], src.Q, cudaMemcpyHostToDevice, c->copy_stream));
                    if (src.T) CU(cudaMemcpyAsync(c->d_t.as<char>() + i * src.T, src.t_ptr[i], src.T, cudaMemcpyHostToDevice, c->copy_stream));
                    c->h2d_bytes += src.Q + src.T;
                }
                pairs_up = pe;
            }
        } else {
            if (src.wait) TRY(src.wait(q_end, t_end));
            if (q_end > qa) CU(cudaMemcpyAsync(c->d_q.as<char>() + qa, src.q + q0 + qa, q_end - qa, cudaMemcpyHostToDevice, c->copy_stream));
            if (t_end > ta) CU(cudaMemcpyAsync(c->d_t.as<char>() + ta, src.t + t0 + ta, t_end - ta, cudaMemcpyHostToDevice, c->copy_stream));
        }
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
    if (!src.packed) c->h2d_bytes += qn + tn;
    tr.mark("enqueue-h2d");
    if (!plan_ready) TRY(host_plan_build(c, n, q_off, t_off, type, match, mismatch, gap, want_cigar));
    b200_align_plan* plan = c->host_plan;
    tr.mark("plan");
    if (src.packed) {
        // the producer packed for the uniform thread-per-pair plan; anything else cannot use what it made
        if (!plan->uniform || plan->uQ != src.Q || plan->uT != src.T || plan->waves.size() > 64)
            return fail(B200_E_ARG, "internal: host-packed upload without a uniform plan");
        plan->host_packed = true;
    }
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
            if (src.packed) {
                // how many of the wave's pairs the host flagged: the run reads this count back instead of one made by
                // pack_kernel; it travels behind the wave's slice, and the wave waits for an event of its own after it
                if (src.wait) TRY(src.wait(qe, te));
                uint32_t cnt = 0;
                for (uint64_t i = plan->waves[k].first; i < last_pair; ++i) cnt += src.flags[i] != 0;
                c->h_wave_cnt.as<uint32_t>()[k] = cnt;
                CU(cudaMemcpyAsync(c->wave_flagged.as<uint32_t>() + k, c->h_wave_cnt.as<uint32_t>() + k, 4, cudaMemcpyHostToDevice, c->copy_stream));
                if (c->copy_events.size() <= cuts.size()) {
                    cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                    c->copy_events.push_back(e);
                }
                CU(cudaEventRecord(c->copy_events[cuts.size()], c->copy_stream));
                plan->wave_events[k] = c->copy_events[cuts.size()];
                cuts.push_back(cuts.back());
                continue;
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
    // packed mode (uniform lengths Q / T): instead of copying, every sequence is packed to 2 bits (host_pack.hpp) into
    // its wq / wt words, and a pair's flag byte says whether both sequences were pure ACGT
    bool packed = false;
    uint32_t len_q = 0, len_t = 0, wq = 0, wt = 0;
    uint32_t* pk[2] = {nullptr, nullptr};
    uint8_t* flags = nullptr;

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
                    if (packed) {
                        for (size_t i = a; i < e; ++i) {
                            if ((len_q && !src[0][i]) || (len_t && !src[1][i])) { bad.store(1, std::memory_order_relaxed); flags[i] = 0; continue; }
                            const uint8_t fq = host_pack_sequence(src[0][i], len_q, pk[0] + i * wq);
                            const uint8_t ft = host_pack_sequence(src[1][i], len_t, pk[1] + i * wt);
                            flags[i] = fq | ft;
                        }
                    } else
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
    const uint64_t* qo = g.off[0];
    const uint64_t* to = g.off[1];
    // Option `host_pack`: the plan comes first -- it says whether the batch is a uniform one for the thread-per-pair
    // kernel -- and then the gather pass packs to 2 bits instead of copying (the upload shrinks from Q + T bytes per pair
    // to 4 (Q/16 + T/16 + 4)). Off by default: on the 16-thread host of the B200 boxes both passes run at the same
    // ~35 GB/s of sequence bytes (12.8 ms against 12.3 ms per config-2 step), i.e. the gather is bound by the host's
    // cores, not by what it writes or by the upload that follows; a host with more cores per GPU is where it would pay.
    bool plan_ready = false;
    if (c->host_pack && n >= 4 * Gatherer::kBlock) {
        TRY(host_plan_build(c, n, qo, to, type, match, mismatch, gap, cigar_off != nullptr));
        plan_ready = true;
    }
    const b200_align_plan* plan = c->host_plan;
    HostSource src;
    if (plan_ready && plan->uniform && plan->waves.size() <= 64) {
        g.packed = true; g.len_q = plan->uQ; g.len_t = plan->uT; g.wq = g.len_q / 16 + 2; g.wt = g.len_t / 16 + 2;
        TRY(c->h_qpk.ensure((uint64_t)n * g.wq * 4 + 64));
        TRY(c->h_tpk.ensure((uint64_t)n * g.wt * 4 + 64));
        TRY(c->h_flags.ensure(n + 64));
        g.pk[0] = c->h_qpk.as<uint32_t>(); g.pk[1] = c->h_tpk.as<uint32_t>(); g.flags = c->h_flags.as<uint8_t>();
        src.packed = true; src.Q = g.len_q; src.T = g.len_t; src.qpk = g.pk[0]; src.tpk = g.pk[1]; src.flags = g.flags;
        src.q_ptr = query; src.t_ptr = target;
    } else {
        TRY(c->h_q.ensure(g.off[0][n] + 1));
        TRY(c->h_t.ensure(g.off[1][n] + 1));
        g.dst[0] = c->h_q.as<char>(); g.dst[1] = c->h_t.as<char>();
        src.q = g.dst[0]; src.t = g.dst[1];
    }
    g.start(T);
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
    int rc = align_batch_host(c, n, src, qo, to, type, match, mismatch, gap, s_score, s_tb, s_cig, s_off, s_cap, plan_ready);
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
