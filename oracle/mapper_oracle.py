"""oracle/mapper_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference mapper's per-read pipeline (reference team_mapper.cpp):
index build :412-477, remove_duplicates :28-45, seed lookup :627-638 (FASTA input) / :716-729
(FASTQ input), FindLIS :283-316, strand choice + region :639-656 / :731-745, Align + PAF :666-698.
Minimize / Align come from oracle/liboracle.so. Pinned by tests/golden/mapper/*.paf, which the
unmodified reference mapper printed (tools/make_mapper_golden.py).

Only f = 0 is restated exactly: with f > 0 the reference's ban list depends on std::sort (unstable)
over unordered_map iteration order (team_mapper.cpp:437-450), which is implementation-defined; this
oracle (and the GPU mapper) break those ties by (count desc, hash asc).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))

COMP = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp(s: bytes) -> bytes:
    return s.translate(COMP)[::-1]   # other bytes unchanged (switch without default, :49-63)


def read_fasta(path):
    recs, name, chunks = [], None, []
    with open(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if not line:
                continue
            if line[:1] == b">":
                if name is not None:
                    recs.append((name, b"".join(chunks)))
                name, chunks = line[1:].split()[0].decode() if line[1:].split() else "", []
            else:
                chunks.append(line)
    if name is not None:
        recs.append((name, b"".join(chunks)))
    return recs


def read_fastq(path):
    recs = []
    with open(path, "rb") as f:
        lines = [ln.rstrip(b"\r\n") for ln in f]
    lines = [ln for ln in lines if ln != b""] if False else lines
    i = 0
    while i < len(lines):
        if not lines[i]:
            i += 1
            continue
        if lines[i][:1] != b"@" or i + 3 >= len(lines) + 0 and False:
            raise ValueError("not FASTQ")
        if i + 3 >= len(lines) + 1 or lines[i + 2][:1] != b"+":
            raise ValueError("not FASTQ")
        recs.append((lines[i][1:].split()[0].decode(), lines[i + 1]))
        i += 4
    if not recs:
        raise ValueError("not FASTQ")
    return recs


def load_reads(path):
    """FASTQ first, FASTA on failure -- the reference sniffs by exception (:533-556)."""
    try:
        return read_fastq(path), True
    except Exception:
        return read_fasta(path), False


def find_lis(matches):
    """:283-316, including the unsigned-wrap distance tests and earliest-predecessor rule."""
    n = len(matches)
    if n == 0:
        return []
    lis, prev = [1] * n, [-1] * n
    M = 1 << 32
    for i in range(1, n):
        fi, si = matches[i]
        for j in range(i):
            fj, sj = matches[j]
            if si > sj and lis[i] < lis[j] + 1 and fi != fj and ((fi - fj) % M) < 5000 and ((si - sj) % M) < 5000:
                lis[i] = lis[j] + 1
                prev[i] = j
    mi = lis.index(max(lis))
    out, i = [], mi
    while True:
        out.append(matches[i])
        if prev[i] == -1:
            break
        i = prev[i]
    return out[::-1]


class Index:
    def __init__(self, oracle, ref: bytes, k, w, f=0.0):
        self.ref, self.rc = ref, revcomp(ref)
        hf, pf, _ = oracle.minimize(self.ref, k, w, True)
        hr, pr, _ = oracle.minimize(self.rc, k, w, False)
        banned_f, banned_r = set(), set()
        if f > 0:
            # threshold from the REVERSE strand's distinct tuples for both (process-global quirk, :433-434),
            # and the reverse ban list is drawn from the FORWARD frequency table (:469)
            n_ban = int(f * len(set(zip(hr.tolist(), pr.tolist()))))
            freq = {}
            for h in hf.tolist():
                freq[h] = freq.get(h, 0) + 1
            freq_r = {}
            for h in hr.tolist():
                freq_r[h] = freq_r.get(h, 0) + 1
            top = sorted(freq.items(), key=lambda kv: (-kv[1], kv[0]))
            banned_f = {h for h, _ in top[:min(n_ban, len(top))]}
            banned_r = {h for h, _ in top[:min(n_ban, len(freq_r), len(top))]}
        self.fwd, self.rev = {}, {}
        for h, p in zip(hf.tolist(), pf.tolist()):
            if h not in banned_f:
                self.fwd.setdefault(h, set()).add(p)
        for h, p in zip(hr.tolist(), pr.tolist()):
            if h not in banned_r:
                self.rev.setdefault(h, set()).add(p)
        self.fwd = {h: sorted(v) for h, v in self.fwd.items()}
        self.rev = {h: sorted(v) for h, v in self.rev.items()}


def map_read(oracle, idx: Index, read: bytes, k, w, typ, m, x, g, want_cigar, fasta_path):
    """-> None or dict(q_begin, q_end, strand, t_begin, t_end, score, cigar)"""
    h, p, fl = oracle.minimize(read, k, w, True)
    seen, mins = set(), []
    for t in zip(h.tolist(), p.tolist(), fl.tolist()):
        if t not in seen:
            seen.add(t)
            mins.append(t)
    mf, mr = [], []
    for hh, fp, _ in mins:
        if fasta_path:
            if hh in idx.fwd:          # :630-637: the reverse index is only consulted for hashes in the forward one
                mf += [(fp, rp) for rp in idx.fwd[hh]]
                mr += [(fp, rp) for rp in idx.rev.get(hh, [])]
        else:
            mf += [(fp, rp) for rp in idx.fwd.get(hh, [])]
            mr += [(fp, rp) for rp in idx.rev.get(hh, [])]
    cf, cr = find_lis(mf), find_lis(mr)
    fwd = len(cf) >= len(cr)
    chain = cf if fwd else cr
    if not chain:
        return None
    qb, qe = chain[0][0] - 1, chain[-1][0] + k - 2
    tb, te = chain[0][1] - 1, chain[-1][1] + k - 2
    target = idx.ref if fwd else idx.rc
    score, _, cigar = oracle.align(read[qb:qe + 1], target[tb:te + 1], typ, m, x, g, want_cigar)
    return dict(q_begin=qb, q_end=qe, fwd=fwd, t_begin=tb, t_end=te, score=score, cigar=cigar)


def paf_line(name, read_len, ref_name, ref_len, r, want_cigar):
    if r["fwd"]:
        ts, te, strand = r["t_begin"], r["t_end"] + 1, "+"
    else:
        ts, te, strand = ref_len - r["t_end"] - 1, ref_len - r["t_begin"], "-"
    cols = [name, read_len, r["q_begin"], r["q_end"] + 1, strand, ref_name, ref_len, ts, te, r["score"],
            r["q_end"] - r["q_begin"] + 1, 60]
    line = "\t".join(str(c) for c in cols)
    if want_cigar:
        line += "\tcg:Z:" + r["cigar"].decode("latin-1")
    return line


def run_mapper(oracle, ref_path, reads_path, typ=0, m=1, x=-1, g=-1, k=15, w=5, f=0.001, want_cigar=False):
    refs = read_fasta(ref_path)
    ref_name, ref = refs[0]                       # only the first reference sequence is used (:415)
    reads, is_fastq = load_reads(reads_path)
    idx = Index(oracle, ref, k, w, f)
    out = []
    for name, seq in reads:
        r = map_read(oracle, idx, seq, k, w, typ, m, x, g, want_cigar, fasta_path=not is_fastq)
        if r is not None:
            out.append(paf_line(name, len(seq), ref_name, len(ref), r, want_cigar))
    return out


def parse_mapper_argv(argv):
    """The reference's hand-rolled option loop (:357-391) for the options the tests use."""
    o = dict(typ=0, m=1, x=-1, g=-1, k=15, w=5, f=0.001, want_cigar=False)
    files, i = [], 0
    names = {"global": 0, "local": 1, "semiGlobal": 2}
    while i < len(argv):
        a = argv[i]
        if a == "-a":
            o["typ"] = names[argv[i + 1]]; i += 1
        elif a in ("-m", "-n", "-g", "-k", "-w"):
            o[{"-m": "m", "-n": "x", "-g": "g", "-k": "k", "-w": "w"}[a]] = int(argv[i + 1]); i += 1
        elif a == "-f":
            o["f"] = float(argv[i + 1]); i += 1
        elif a == "-c":
            o["want_cigar"] = True
        elif a == "-s":
            pass
        else:
            files.append(a)
        i += 1
    return o, files
