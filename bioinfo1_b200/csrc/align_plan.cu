// align_plan.cu -- shape-only planning of an alignment batch: which fill kernel serves each pair, the work
// order, the waves (slices whose direction matrices fit the HBM budget together) and the per-pair descriptors.
#include <algorithm>
#include <cstdlib>
#include <new>
#include <thread>

#include "internal.hpp"

using namespace b200;

// Uniform batches (every pair the same Q x T, e.g. fixed-length short reads): the descriptors are an
// affine function of the pair index, so they are generated on the device instead of being built on
// the host and copied (48 B per pair).
__device__ __host__ inline uint32_t uniform_group_in_wave(uint32_t g, uint32_t groups_per_wave, const UniformTail& tail) {
    if (g < tail.first_group) return g % groups_per_wave;
    uint32_t k = 0;
    while (k + 1 < tail.n && g >= tail.start[k + 1]) ++k;
    return g - tail.start[k];
}

__global__ void build_uniform_plan_kernel(uint32_t n, uint32_t Q, uint32_t T, uint64_t q_base, uint64_t t_base,
                                          uint64_t words_per_group, uint32_t groups_per_wave, UniformTail tail,
                                          PairDesc* __restrict__ pairs,
                                          uint32_t* __restrict__ work, ShortGroup* __restrict__ groups) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PairDesc d;
    d.q_off = q_base + (uint64_t)i * Q;
    d.t_off = t_base + (uint64_t)i * T;
    d.dir_off = (uint64_t)uniform_group_in_wave(i >> 6, groups_per_wave, tail) * words_per_group;   // relative to the wave's buffer
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

void materialize_uniform_host(b200_align_plan* p) {   // only needed by the non-ACGT fallback
    if (!p->uniform || !p->h_pairs.empty()) return;
    const uint64_t wpg = p->want_cigar ? (uint64_t)div_up(p->uQ, kShortRows) * p->uT * 128 : 0;
    p->h_pairs.resize(p->n);
    p->h_order.resize(p->n);
    for (size_t i = 0; i < p->n; ++i) {
        PairDesc& d = p->h_pairs[i];
        d.q_off = p->u_qbase + i * p->uQ; d.t_off = p->u_tbase + i * p->uT;
        d.dir_off = uniform_group_in_wave((uint32_t)(i >> 6), (uint32_t)p->u_groups_per_wave, p->u_tail) * wpg; d.run_off = i * ((uint64_t)p->uQ + p->uT + 1);
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

uint64_t generic_dir_words(uint32_t Q, uint32_t T) {
    return (uint64_t)div_up(Q, kRowsPerWord) * ((T + 3u) & ~3u);
}

// Can the int16 tagged kernel (align_fill_short.cuh) represent every value of this pair?
bool short_scores_ok(const Scores& sc, int type) {
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

// A class whose direction matrices fit the budget runs as ONE wave (small waves under-fill the machine and
// every wave pays its own tail). Otherwise it is cut into waves of half the budget, two of which are in
// flight at a time (see b200_align_plan_run): the walk of wave k then overlaps the fill of wave k+1.
uint64_t wave_budget_words(const b200_ctx* ctx) { return std::max<uint64_t>((uint64_t)ctx->dir_budget_bytes / 4, 1 << 16); }
static inline uint64_t wave_cap_words(uint64_t budget_words, uint64_t class_total_words) {
    return class_total_words <= budget_words ? budget_words : std::max<uint64_t>(budget_words / 2, 1 << 15);
}

// Fills `p` (fresh or recycled) for a batch. `rebase`: offsets are taken relative to q_off[0] /
// t_off[0] (the host entry points copy only the referenced byte range to the device).
int plan_build(b200_align_plan* p, b200_ctx* ctx, size_t n, const uint64_t* q_off, const uint64_t* t_off,
               bool rebase, bool sync, int type, int match, int mismatch, int gap, int want_cigar,
               size_t chunk_pairs, const std::vector<uint32_t>* tail_groups) {
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
        // ... and end with a tail of explicitly sized, shrinking waves (sizes in groups, none larger than a regular
        // chunk): what is left to do when the last byte of the upload lands is then the smallest wave's work
        UniformTail tail{(uint32_t)n_groups, 0, {0}};
        if (chunk_pairs && tail_groups && !tail_groups->empty()) {
            uint64_t sum = 0;
            size_t first = tail_groups->size();
            while (first > 0 && tail.n < 16 && sum + (*tail_groups)[first - 1] <= n_groups && (*tail_groups)[first - 1] > 0 &&
                   (*tail_groups)[first - 1] <= groups_per_wave) {
                sum += (*tail_groups)[--first];
                ++tail.n;
            }
            tail.first_group = (uint32_t)(n_groups - sum);
            uint32_t at = tail.first_group;
            for (uint32_t k = 0; k < tail.n; ++k) { tail.start[k] = at; at += (*tail_groups)[first + k]; }
        }
        const uint64_t wave_words_u = std::min(n_groups, groups_per_wave) * wpg;
        if (uni && wave_words_u <= (groups_per_wave < n_groups ? std::max<uint64_t>(budget_words / 2, 1 << 15) : budget_words)) {
            p->uniform = true; p->uQ = (uint32_t)Q0; p->uT = (uint32_t)T0; p->u_groups_per_wave = groups_per_wave;
            p->u_tail = tail;
            p->u_qbase = q_off[0] - qb; p->u_tbase = t_off[0] - tb;
            p->run_slots = n * (Q0 + T0 + 1);
            p->qpk_words = n * (Q0 / 16 + 2); p->tpk_words = n * (T0 / 16 + 2);
            p->cells = n * Q0 * T0;
            p->cigar_bound = n * std::max<uint64_t>(2, 2 * (Q0 + T0));
            p->max_T = p->max_T_short = (uint32_t)T0;
            p->max_Q = p->max_Q_short = (uint32_t)Q0;
            p->n_short = n;
            auto add_wave = [&](uint64_t g0, uint64_t g1) {
                const uint64_t first = g0 * 64, last = std::min<uint64_t>(n, g1 * 64);
                p->waves.push_back(Wave{kClassShort, (uint32_t)first, (uint32_t)(last - first), (uint32_t)g0, (g1 - g0) * wpg});
            };
            for (uint64_t g0 = 0; g0 < tail.first_group; g0 += groups_per_wave) add_wave(g0, std::min<uint64_t>(tail.first_group, g0 + groups_per_wave));
            for (uint32_t k = 0; k < tail.n; ++k) add_wave(tail.start[k], k + 1 < tail.n ? tail.start[k + 1] : n_groups);
            TRY(p->d_pairs.ensure(n * sizeof(PairDesc)));
            TRY(p->d_work.ensure(n * sizeof(uint32_t)));
            TRY(p->d_groups.ensure(n_groups * sizeof(ShortGroup)));
            build_uniform_plan_kernel<<<(unsigned)div_up64(n, 256), 256, 0, ctx->stream>>>(
                (uint32_t)n, p->uQ, p->uT, p->u_qbase, p->u_tbase, wpg, (uint32_t)groups_per_wave, tail, p->d_pairs.as<PairDesc>(),
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
int plan_finish(b200_align_plan* p, b200_ctx* ctx, bool short_scores, bool sync) {
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
    const int rc = plan_build(p, ctx, n, q_off, t_off, false, true, type, match, mismatch, gap, want_cigar, 0, nullptr);
    if (rc != B200_OK) { b200_align_plan_destroy(p); return rc; }
    *out = p;
    return B200_OK;
}
