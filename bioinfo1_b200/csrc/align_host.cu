// align_host.cu -- the host-buffer entry points of the alignment ABI: upload pipeline over packed host
// buffers, and the reference-shaped pointer-array form on top of it.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "internal.hpp"

using namespace b200;

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
        size_t round_pairs = 0;
        TRY(align_short_round_pairs(c, type, &round_pairs));
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
