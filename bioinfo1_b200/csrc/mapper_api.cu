// mapper_api.cu -- the "next" rows of the hot path: minimizer index build (radix sort + unique on the GPU)
// and the per-read mapping batch (Minimize -> de-dup -> seeds -> chains -> region -> Align).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "internal.hpp"
#include "mapper.cuh"

using namespace b200;

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
struct U32Widen {
    __host__ __device__ uint64_t operator()(const uint32_t& v) const { return (uint64_t)v; }
};

// Exclusive scan of per-minimizer match counts into uint32 slots. The total is summed in 64 bits first: with
// f = 0 or a repetitive reference the matches of one batch can pass 2^32, the 32-bit offsets would wrap and the
// emit / chain kernels would write out of bounds (the reference is merely slow there). Such a batch is refused.
static int scan_u32(b200_ctx* c, const uint32_t* d_in, uint64_t n, uint32_t* d_out, uint32_t* total, cudaStream_t st) {
    *total = 0;
    if (n == 0) return B200_OK;
    if (n > 0x7fffffffull) return fail(B200_E_ARG, "too many items for one scan");
    TRY(c->total.ensure(16));
    cub::TransformInputIterator<uint64_t, U32Widen, const uint32_t*> wide(d_in, U32Widen());
    size_t tmp = 0;
    CU(cub::DeviceReduce::Sum(nullptr, tmp, wide, c->total.as<uint64_t>(), (int)n, st));
    TRY(c->scan_tmp.ensure(tmp));
    CU(cub::DeviceReduce::Sum(c->scan_tmp.p, tmp, wide, c->total.as<uint64_t>(), (int)n, st));
    uint64_t sum = 0;
    CU(cudaMemcpyAsync(&sum, c->total.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (sum > 0x7fffffffull)
        return fail(B200_E_ARG, "seed matches of this batch (" + std::to_string(sum) + ") exceed 2^31: split the reads into smaller batches");
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_in, d_out, (int)n, st));
    TRY(c->scan_tmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(c->scan_tmp.p, tmp, d_in, d_out, (int)n, st));
    c->kernel_launches += 3;
    *total = (uint32_t)sum;
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
