// b200_mapper.cpp -- the <team>_mapper command line over the B200 library (drop-in for the
// executable built from reference team_mapper.cpp: same positional arguments, options and PAF
// columns; see main() there, :328-795).
//
//   b200_mapper [options] <reference.fasta> <fragments.fasta|fastq>
//
// Differences that are deliberate and documented:
//   * -s statistics go to stderr (the assignment text and BASELINE.json ask for stderr; the reference
//     prints them to stdout, team_mapper.cpp:186-225), so stdout carries PAF only;
//   * --gpus N shards the reads over N devices (the reference's "-t threads" was never implemented;
//     -t is accepted and ignored); gzip-compressed input files are inflated through zlib, as the reference's are;
//   * with -f > 0 ties at the frequency cut are broken by (count desc, hash asc) -- the reference's
//     order there is implementation-defined.
// All arithmetic of the path (minimizers, index, seeds, chaining, alignment) runs on the GPU through
// include/b200map.h; this file is argument parsing, FASTA/FASTQ text, statistics and PAF printing.
//
// Host pipeline: the input files are mapped into memory and parsed in one pass (memchr over the mapping, the
// sequences land back to back in ONE packed buffer with offsets -- exactly what b200_map_batch takes, so a batch is
// a slice of it and nothing is copied per batch) while, on other threads, every device creates its contexts and
// builds its copy of the index. Batches of reads are then taken from ONE queue by all workers of all devices
// (two contexts / host threads per device), which balances itself whatever the read lengths; each worker formats
// the PAF text of its batch from the batch's own result buffers, and the main thread writes finished batches in
// input order while later ones are still being mapped.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <mutex>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "b200map.h"

namespace {

constexpr const char* kProgram = "toolForGenomeAllignment";   // the reference's PROGRAM_NAME, kept for scripts
constexpr const char* kVersion = "3.1.0";

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void usage(std::ostream& os) {
    os << "\nUsage: " << kProgram << " [options] <file1> <file2>\n"
       << "  file1: reference genome, FASTA. file2: fragments, FASTA or FASTQ.\n"
       << "Options:\n"
       << "\t  -a TYPE    alignment type: global, local, semiGlobal (default: global)\n"
       << "\t  -m MATCH   match score (default: 1)\n"
       << "\t  -n MISMATCH mismatch score (default: -1)\n"
       << "\t  -g GAP     gap score (default: -1)\n"
       << "\t  -k KMER    k-mer length of the minimizers (default: 15)\n"
       << "\t  -w WINDOW  window length of the minimizers (default: 5)\n"
       << "\t  -f FRACTION fraction of most frequent minimizers to ignore (default: 0.001)\n"
       << "\t  -c         print the CIGAR string (cg:Z: tag)\n"
       << "\t  -s         print statistics of both files (stderr)\n"
       << "\t  --gpus N   shard the fragments over N GPUs (default: 1)\n"
       << "\t  -h, --help, --version\n";
}

// ---- input files: read into memory once, parsed in place ----------------------------------------------------------
// Large host buffers (file contents, packed sequences) come from 2 MB-aligned allocations advised to use huge pages:
// a gigabyte filled through 4 KB page faults takes the process's memory-map lock a quarter of a million times, and the
// CUDA contexts starting up on the other threads want that lock too (measured on 8 GPUs: the parse went from 1.1 s to
// 5.6 s with a plain mmap of the file).
struct HugeBuf {
    char* p = nullptr;
    size_t n = 0, cap = 0;
    bool reserve(size_t want) {
        if (want <= cap) return true;
        const size_t two_mb = (size_t)2 << 20, sz = (want + two_mb - 1) / two_mb * two_mb;
        void* q = nullptr;
        if (posix_memalign(&q, two_mb, sz) != 0) return false;
        madvise(q, sz, MADV_HUGEPAGE);
        if (n) std::memcpy(q, p, n);
        std::free(p);
        p = static_cast<char*>(q); cap = sz;
        return true;
    }
    bool append(const char* b, const char* e) {
        const size_t k = (size_t)(e - b);
        if (n + k > cap && !reserve(std::max(n + k, cap + cap / 2))) return false;
        std::memcpy(p + n, b, k);
        n += k;
        return true;
    }
    const char* data() const { return p; }
    size_t size() const { return n; }
};

struct MappedFile {
    const char* p = nullptr;
    size_t n = 0;
    HugeBuf bytes;
    bool open(const std::string& path) {
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { ::close(fd); return false; }
        const size_t sz = (size_t)st.st_size;
        if (!bytes.reserve(sz + 64)) { ::close(fd); return false; }
        while (bytes.n < sz) {
            const ssize_t got = ::read(fd, bytes.p + bytes.n, std::min<size_t>(sz - bytes.n, (size_t)64 << 20));
            if (got < 0) { ::close(fd); return false; }
            if (got == 0) break;
            bytes.n += (size_t)got;
        }
        ::close(fd);
        p = bytes.p; n = bytes.n;
        // gzip input (the reference reads it through bioparser + zlib, CMakeLists.txt:28-30): inflated into memory once
        if (n >= 2 && (unsigned char)p[0] == 0x1f && (unsigned char)p[1] == 0x8b) return inflate_all(path);
        return true;
    }
    bool inflate_all(const std::string& path) {
        gzFile g = gzopen(path.c_str(), "rb");
        if (!g) return false;
        gzbuffer(g, 1 << 20);
        HugeBuf out;
        if (!out.reserve(4 * n + 64)) { gzclose(g); return false; }
        std::vector<char> chunk(1 << 22);
        for (;;) {
            const int got = gzread(g, chunk.data(), (unsigned)chunk.size());
            if (got < 0) { gzclose(g); return false; }
            if (got == 0) break;
            if (!out.append(chunk.data(), chunk.data() + got)) { gzclose(g); return false; }
        }
        gzclose(g);
        std::free(bytes.p);
        bytes = out;
        p = bytes.p; n = bytes.n;
        return true;
    }
};

// Sequences of a file, packed: sequence i is buf[off[i] .. off[i+1]); names point into the file's bytes.
struct SeqSet {
    std::vector<std::string_view> names;
    HugeBuf buf;
    std::vector<uint64_t> off{0};
    size_t size() const { return names.size(); }
    uint64_t len(size_t i) const { return off[i + 1] - off[i]; }
};

// one line [b, e) without its '\n' / '\r'; advances `at` past the newline
inline bool next_line(const char* p, size_t n, size_t& at, const char*& b, const char*& e) {
    if (at >= n) return false;
    b = p + at;
    const char* nl = static_cast<const char*>(std::memchr(b, '\n', n - at));
    e = nl ? nl : p + n;
    at = (size_t)(e - p) + (nl ? 1 : 0);
    while (e > b && e[-1] == '\r') --e;
    return true;
}
inline std::string_view first_token(const char* b, const char* e) {   // header without its marker, up to the first blank
    const char* s = b + 1;
    const char* t = s;
    while (t < e && *t != ' ' && *t != '\t') ++t;
    return std::string_view(s, (size_t)(t - s));
}

bool parse_fasta(const MappedFile& f, SeqSet& out) {
    out = SeqSet();
    if (!out.buf.reserve(f.n + 64)) return false;
    size_t at = 0;
    const char *b, *e;
    bool have = false;
    while (next_line(f.p, f.n, at, b, e)) {
        if (b == e) continue;
        if (*b == '>') {
            if (have) out.off.push_back(out.buf.size());
            out.names.push_back(first_token(b, e));
            have = true;
        } else if (!have) return false;
        else if (!out.buf.append(b, e)) return false;
    }
    if (have) out.off.push_back(out.buf.size());
    return !out.names.empty();
}

bool parse_fastq(const MappedFile& f, SeqSet& out) {
    out = SeqSet();
    if (!out.buf.reserve(f.n / 2 + 64)) return false;
    size_t at = 0;
    const char *b, *e, *sb, *se, *pb, *pe, *qb, *qe;
    while (next_line(f.p, f.n, at, b, e)) {
        if (b == e) continue;
        if (*b != '@') return false;
        if (!next_line(f.p, f.n, at, sb, se) || !next_line(f.p, f.n, at, pb, pe) || !next_line(f.p, f.n, at, qb, qe)) return false;
        if (pb == pe || *pb != '+' || qe - qb != se - sb) return false;
        out.names.push_back(first_token(b, e));
        if (!out.buf.append(sb, se)) return false;
        out.off.push_back(out.buf.size());
    }
    return !out.names.empty();
}

void basic_stats(const char* kind, const SeqSet& recs) {   // reference :186-225 / :229-280
    size_t total = 0, mx = 0, mn = SIZE_MAX;
    std::vector<size_t> lens;
    for (size_t i = 0; i < recs.size(); ++i) {
        const size_t l = recs.len(i);
        std::cerr << "Sequence" << kind << " name: " << recs.names[i] << "\nLength of sequence: " << l << "\n";
        lens.push_back(l);
        total += l; mx = std::max(mx, l); mn = std::min(mn, l);
    }
    std::cerr << "Total number of sequences: " << recs.size() << "\nAverage length of sequences: " << total / recs.size()
              << "\nMaximal length of sequence: " << mx << "\nMinimal length of sequence: " << mn << "\n";
    std::sort(lens.begin(), lens.end(), std::greater<size_t>());
    size_t cum = 0;
    for (size_t l : lens) { cum += l; if (cum >= total / 2) { std::cerr << "N50 length: " << l << "\n"; break; } }
}

// distinct minimizers / singleton fraction of one sequence (reference :481-525, :610-624), via MinimizeBatch
void minimizer_stats(int device, const char* seq, size_t seq_len, bool fwd, uint32_t k, uint32_t w, const char* label) {
    const uint64_t cnt = b200_minimize_count((uint32_t)seq_len, k, w);
    std::vector<uint32_t> hash(cnt ? cnt : 1), pos(cnt ? cnt : 1);
    std::vector<uint8_t> flag(cnt ? cnt : 1);
    uint64_t off[2] = {0, 0};
    const uint32_t len = (uint32_t)seq_len;
    const uint8_t fl = fwd ? 1 : 0;
    if (b200_minimize_batch(device, 1, &seq, &len, k, w, &fl, hash.data(), pos.data(), flag.data(), off, cnt) != B200_OK) return;
    std::unordered_map<uint32_t, int> freq;
    for (uint64_t i = 0; i < cnt; ++i) freq[hash[i]]++;
    size_t singles = 0;
    for (const auto& kv : freq) singles += kv.second == 1;
    std::cerr << "Number of distinct minimizers for " << label << ": " << freq.size() << "\n";
    if (!freq.empty()) std::cerr << "Fraction of singletons on " << label << ": " << (double)singles / freq.size() << "\n";
}

struct Options {
    int type = B200_GLOBAL, match = 1, mismatch = -1, gap = -1;
    uint32_t k = 15, w = 5;
    double f = 0.001;
    bool cigar = false, stats = false;
    int gpus = 1;
    std::string file1, file2;
};

// ---- PAF text (reference :685-698) -----------------------------------------------------------------------------
inline void put_u64(std::string& s, uint64_t v) {
    char tmp[24];
    const auto r = std::to_chars(tmp, tmp + sizeof tmp, v);
    s.append(tmp, (size_t)(r.ptr - tmp));
}
inline void put_i64(std::string& s, int64_t v) {
    char tmp[24];
    const auto r = std::to_chars(tmp, tmp + sizeof tmp, v);
    s.append(tmp, (size_t)(r.ptr - tmp));
}

struct Batch {
    size_t lo = 0, hi = 0;        // reads [lo, hi)
    std::string paf;              // the batch's PAF lines, input order
    bool done = false;
};

// Everything the workers share.
struct Job {
    const Options* o = nullptr;
    const SeqSet* reads = nullptr;
    std::string_view ref_name;
    const char* ref = nullptr;
    uint64_t ref_len = 0;
    bool fastq = false;
    std::vector<Batch> batches;
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    std::mutex mu;                // guards `done` flags and `err`
    std::condition_variable cv;
    std::string err;
    bool ready = false;           // batches are final (the fragments file has been parsed)
    bool trace = false;
    uint64_t max_bases = 0;       // largest batch, for the result buffers
    size_t max_reads = 0;
};

void format_batch(const Job& J, const Batch& b, const b200_mapping* m, const char* cig, const uint64_t* coff, std::string& out) {
    const Options& o = *J.o;
    const uint64_t RL = J.ref_len;
    out.clear();
    for (size_t r = b.lo; r < b.hi; ++r) {
        const b200_mapping& x = m[r - b.lo];
        if (!x.mapped) continue;
        const uint64_t ts = x.strand_fwd ? x.t_begin : RL - x.t_end - 1, te = x.strand_fwd ? (uint64_t)x.t_end + 1 : RL - x.t_begin;
        out.append(J.reads->names[r]); out += '\t'; put_u64(out, J.reads->len(r)); out += '\t';
        put_u64(out, x.q_begin); out += '\t'; put_u64(out, (uint64_t)x.q_end + 1); out += '\t';
        out += x.strand_fwd ? '+' : '-'; out += '\t'; out.append(J.ref_name); out += '\t'; put_u64(out, RL); out += '\t';
        put_u64(out, ts); out += '\t'; put_u64(out, te); out += '\t'; put_i64(out, x.score);
        out += '\t'; put_u64(out, (uint64_t)x.q_end - x.q_begin + 1); out += "\t60";
        if (o.cigar) { out += "\tcg:Z:"; out.append(cig + coff[r - b.lo], (size_t)(coff[r - b.lo + 1] - coff[r - b.lo])); }
        out += '\n';
    }
}

// One worker = one context (own streams and workspaces) on `device`, the device's index shared. Takes batches from
// the job's queue until it is empty. b200_map_batch is a blocking call on one context; with two workers per device
// one batch's seeding, chaining, planning and downloads overlap the other's alignment kernels.
void worker(Job& J, int device, int wi, b200_ctx* ctx, const b200_index* ix) {
    const Options& o = *J.o;
    std::vector<b200_mapping> m(J.max_reads ? J.max_reads : 1);
    const uint64_t cap = o.cigar ? 4 * J.max_bases + 64 * J.max_reads + 64 : 0;
    std::vector<char> cig(cap ? cap : 1);
    std::vector<uint64_t> coff(J.max_reads + 1, 0);
    std::string text;
    for (;;) {
        const size_t bi = J.next.fetch_add(1);
        if (bi >= J.batches.size() || J.failed.load()) return;
        Batch& b = J.batches[bi];
        const size_t n = b.hi - b.lo;
        const double t0 = now_s();
        const int e = b200_map_batch(ctx, ix, n, J.reads->buf.data(), J.reads->off.data() + b.lo, J.fastq ? 1 : 0, o.type, o.match,
                                     o.mismatch, o.gap, o.cigar ? 1 : 0, m.data(), o.cigar ? cig.data() : nullptr,
                                     o.cigar ? coff.data() : nullptr, cap);
        if (e != B200_OK) {
            // the reference logs and skips a read whose Align throws (:680-683); a batch failure is fatal here
            std::lock_guard<std::mutex> g(J.mu);
            if (J.err.empty()) J.err = std::string("ERROR: Exception during Align: ") + b200_last_error();
            J.failed.store(1);
            J.cv.notify_all();
            return;
        }
        const double t1 = now_s();
        format_batch(J, b, m.data(), cig.data(), coff.data(), text);
        {
            std::lock_guard<std::mutex> g(J.mu);
            b.paf.swap(text);
            b.done = true;
        }
        J.cv.notify_all();
        if (J.trace)
            std::fprintf(stderr, "[b200_mapper trace] gpu %d worker %d: batch of %zu reads mapped in %.3f s, formatted in %.3f s\n", device, wi, n,
                         t1 - t0, now_s() - t1);
    }
}

// One device: two contexts at most, one index; runs its workers to the end of the queue.
void device_main(Job& J, int device, int n_workers) {
    const Options& o = *J.o;
    const double t0 = now_s();
    std::vector<b200_ctx*> ctxs((size_t)n_workers, nullptr);
    auto fail_with = [&](const std::string& what) {
        std::lock_guard<std::mutex> g(J.mu);
        if (J.err.empty()) J.err = what;
        J.failed.store(1);
        J.cv.notify_all();
    };
    if (b200_ctx_create(device, &ctxs[0]) != B200_OK) { fail_with(b200_last_error()); return; }
    // the second context is created next to the index build (its start-up is host and allocator time)
    std::thread second_ctx;
    if (n_workers > 1) second_ctx = std::thread([&] { if (b200_ctx_create(device, &ctxs[1]) != B200_OK) ctxs[1] = nullptr; });
    const double t1 = now_s();
    b200_index* ix = nullptr;
    const int rc = b200_index_build(ctxs[0], J.ref, J.ref_len, o.k, o.w, o.f, &ix);
    if (second_ctx.joinable()) second_ctx.join();
    if (rc != B200_OK) { fail_with(b200_last_error()); return; }
    if (J.trace) std::fprintf(stderr, "[b200_mapper trace] gpu %d: context %.3f s, index %.3f s\n", device, t1 - t0, now_s() - t1);
    {   // the fragments are parsed on the main thread meanwhile
        std::unique_lock<std::mutex> lk(J.mu);
        J.cv.wait(lk, [&] { return J.ready || J.failed.load(); });
    }
    if (J.failed.load()) return;
    std::vector<std::thread> th;
    for (int wi = 1; wi < n_workers; ++wi)
        if (ctxs[(size_t)wi]) th.emplace_back(worker, std::ref(J), device, wi, ctxs[(size_t)wi], ix);   // one worker is still correct
    worker(J, device, 0, ctxs[0], ix);
    for (auto& t : th) t.join();
    // The index and the contexts (tens of GB of device workspace) are deliberately not destroyed: the process is
    // about to exit, and returning that memory piece by piece costs about a second.
}

}  // namespace

int main(int argc, char** argv) {
    const bool trace = std::getenv("B200_TRACE") != nullptr;   // phase wall clocks on stderr
    const double t_start = now_s();
    Options o;
    if (argc < 2) { std::cerr << "Error: Not enough arguments\n"; usage(std::cout); return 1; }
    const std::string a1 = argv[1];
    if (a1 == "-h" || a1 == "--help") { usage(std::cout); return 0; }
    if (a1 == "--version") { std::cout << kProgram << " v" << kVersion << std::endl; return 0; }
    if (argc < 3) { std::cerr << "Error: Expected two input files\n"; return 1; }
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        const bool has_val = i + 1 < argc;
        if (a == "-a" && has_val) {
            const std::string v = argv[++i];
            if (v == "global") o.type = B200_GLOBAL; else if (v == "local") o.type = B200_LOCAL;
            else if (v == "semiGlobal") o.type = B200_SEMIGLOBAL;
            else { std::cerr << "Error: Expected Alignment type: global, local, semiGlobal\n"; usage(std::cout); return 1; }
        } else if (a == "-m" && has_val) o.match = std::atoi(argv[++i]);
        else if (a == "-n" && has_val) o.mismatch = std::atoi(argv[++i]);
        else if (a == "-g" && has_val) o.gap = std::atoi(argv[++i]);
        else if (a == "-k" && has_val) o.k = (uint32_t)std::atoi(argv[++i]);
        else if (a == "-w" && has_val) o.w = (uint32_t)std::atoi(argv[++i]);
        else if (a == "-f" && has_val) o.f = std::atof(argv[++i]);
        else if (a == "-t" && has_val) ++i;   // accepted for compatibility with the assignment text, unused
        else if (a == "--gpus" && has_val) o.gpus = std::max(1, std::atoi(argv[++i]));
        else if (a == "-c") o.cigar = true;
        else if (a == "-s") o.stats = true;
        else if (o.file1.empty()) o.file1 = a;
        else if (o.file2.empty()) o.file2 = a;
        else { std::cerr << "Unknown or extra argument: " << a << "\n"; usage(std::cout); return 1; }
    }
    if (o.file1.empty() || o.file2.empty()) { std::cerr << "Error: Two input files are required.\n"; usage(std::cout); return 1; }

    MappedFile f1, f2;
    SeqSet refs, reads;
    if (!f1.open(o.file1) || !parse_fasta(f1, refs)) { std::cerr << "Given reference file is not in FASTA format! \n"; return 1; }
    // only the first sequence is the reference (:415)
    const char* ref = refs.buf.data();
    const uint64_t ref_len = refs.len(0);
    // Everything that touches CUDA happens on other threads from here on: initialising the driver on a box with eight
    // GPUs takes seconds by itself (measured: 4-5 s of the first 8-GPU run's 10 s), and the input files can be read and
    // parsed meanwhile. The starter thread counts the devices, clamps --gpus and starts one thread per device; the devices
    // create their contexts and indexes, then wait for the batches. A second context per device (its start-up runs
    // next to the index build) only when the device's share of the input is several batches.
    // (B200_MAPPER_BATCH_READS / B200_MAPPER_WORKERS: test knobs.)
    Job J;
    J.o = &o; J.reads = &reads; J.ref_name = refs.names[0]; J.ref = ref; J.ref_len = ref_len; J.trace = trace;
    struct stat st2;
    if (stat(o.file2.c_str(), &st2) != 0 || !S_ISREG(st2.st_mode)) { std::cerr << "Given file is not in FASTA or FASTQ format! \n"; return 1; }
    const size_t f2_bytes = (size_t)st2.st_size;
    const char* env_batch = std::getenv("B200_MAPPER_BATCH_READS");
    const char* env_workers = std::getenv("B200_MAPPER_WORKERS");
    const int gpus_asked = o.gpus;
    int workers_per_device = 1;
    bool no_device = false;
    double t_cuda_ready = 0;
    std::vector<std::thread> devs;
    // The driver initialises every VISIBLE device, used or not (about 0.6 s per B200 on the 8-GPU boxes): unless the user
    // has set it, the process only sees the first --gpus devices.
    if (!std::getenv("CUDA_VISIBLE_DEVICES")) {
        std::string vis;
        for (int g = 0; g < gpus_asked; ++g) vis += (g ? "," : "") + std::to_string(g);
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 0);
    }
    std::thread starter([&] {
        const int count = b200_device_count();
        t_cuda_ready = now_s();
        if (count <= 0) { no_device = true; return; }
        o.gpus = std::min(gpus_asked, count);
        // (a second context per device only pays when the device gets several batches: ~400 MB of input each)
        workers_per_device = env_workers && std::atoi(env_workers) > 0 ? std::atoi(env_workers)
                             : (f2_bytes / (size_t)o.gpus >= ((size_t)384 << 20) ? 2 : 1);
        for (int g = 0; g < o.gpus; ++g) devs.emplace_back(device_main, std::ref(J), g, workers_per_device);
    });
    auto abort_devices = [&] {
        if (starter.joinable()) starter.join();
        { std::lock_guard<std::mutex> g(J.mu); J.failed.store(1); }
        J.cv.notify_all();
        for (auto& t : devs) t.join();
    };

    bool fastq = f2.open(o.file2) && parse_fastq(f2, reads);   // FASTQ first, FASTA on failure (:533-556)
    if (!fastq && !(f2.p && parse_fasta(f2, reads))) {
        abort_devices();
        std::cerr << "Given file is not in FASTA or FASTQ format! \n";
        std::fflush(stderr);
        std::_Exit(1);
    }
    J.fastq = fastq;
    const double t_parsed = now_s();
    starter.join();   // (o.gpus is final from here on)
    if (no_device) { std::cerr << "Error: no CUDA device visible (this mapper has no CPU fallback)\n"; std::fflush(stderr); std::_Exit(1); }

    if (o.stats) {
        std::cerr << "Basic statistic for reference genome\n------------------------------------\n";
        basic_stats("FASTA", refs);
        minimizer_stats(0, ref, ref_len, true, o.k, o.w, "forward strand");
        std::cerr << "\nBasic statistic for fragments of genome\n------------------------------------\n";
        basic_stats(fastq ? "FASTQ" : "FASTA", reads);
    }

    // Batches of bounded size: small inputs go through one context in large batches, larger ones in batches of 8 k
    // reads taken from one queue by every worker of every device.
    const size_t n_reads = reads.size();
    const bool small = f2_bytes < ((size_t)192 << 20) && o.gpus == 1;
    const size_t max_reads = env_batch && std::atol(env_batch) > 0 ? (size_t)std::atol(env_batch) : (small ? 65536 : 8192);
    const uint64_t max_bases = small ? (256ull << 20) : (64ull << 20);
    for (size_t i = 0; i < n_reads;) {
        size_t j = i;
        uint64_t bases = 0;
        while (j < n_reads && j - i < max_reads && bases < max_bases) bases += reads.len(j++);
        Batch b; b.lo = i; b.hi = j;
        J.batches.push_back(std::move(b));
        J.max_bases = std::max(J.max_bases, bases);
        J.max_reads = std::max(J.max_reads, j - i);
        i = j;
    }
    { std::lock_guard<std::mutex> g(J.mu); J.ready = true; }
    J.cv.notify_all();

    // writer: finished batches in input order, while later ones are still being mapped. Straight write(2) calls of whole
    // batches: stdio's 4 KB buffer turned 300 MB of PAF into 73 000 system calls (1.3 s).
    std::fflush(stdout);
    auto write_all = [](const char* p, size_t n) {
        while (n) {
            const ssize_t w = ::write(STDOUT_FILENO, p, n);
            if (w <= 0) return false;
            p += w; n -= (size_t)w;
        }
        return true;
    };
    size_t written = 0;
    bool failed = false;
    while (written < J.batches.size()) {
        std::string text;
        {
            std::unique_lock<std::mutex> lk(J.mu);
            J.cv.wait(lk, [&] { return J.batches[written].done || J.failed.load(); });
            if (!J.batches[written].done) { failed = true; break; }
            text.swap(J.batches[written].paf);
        }
        if (!write_all(text.data(), text.size())) { failed = true; if (J.err.empty()) J.err = "write to stdout failed"; J.failed.store(1); J.cv.notify_all(); break; }
        ++written;
    }
    for (auto& t : devs) t.join();
    if (failed || J.failed.load()) { std::cerr << J.err << std::endl; std::fflush(stdout); std::_Exit(1); }
    const double t_done = now_s();
    if (trace)
        std::fprintf(stderr, "[b200_mapper trace] read + parse files %.3f s (CUDA driver up after %.3f s), index + map + write PAF %.3f s (%zu batches, %d devices x %d workers)\n",
                     t_parsed - t_start, t_cuda_ready - t_start, t_done - t_parsed, J.batches.size(), o.gpus, workers_per_device);
    std::fflush(stdout);
    std::fflush(stderr);
    std::_Exit(0);   // skip the CUDA runtime's teardown of the (large) device allocations
}
