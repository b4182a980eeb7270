"""Seeded synthetic sequence generators shared by tests and bench.py (numpy only, so the
same seeds give the same bytes everywhere)."""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_CODE = np.zeros(256, dtype=np.int64)
_CODE[ACGT] = np.arange(4)


def random_dna(rng, n):
    return ACGT[rng.integers(0, 4, size=n)]


def mutate(rng, seq, sub=0.0, ins=0.0, dele=0.0):
    """Independent per-base substitution / insertion-before / deletion events."""
    n = len(seq)
    r = rng.random(n)
    keep = r >= dele
    is_ins = (r >= dele) & (r < dele + ins)
    is_sub = (r >= dele + ins) & (r < dele + ins + sub)
    out = seq.copy()
    if is_sub.any():
        idx = _CODE[seq[is_sub]]                       # substitute with a *different* base
        out[is_sub] = ACGT[(idx + rng.integers(1, 4, size=idx.size)) % 4]
    counts = keep.astype(np.int64) + is_ins.astype(np.int64)
    total = int(counts.sum())
    res = np.empty(total, dtype=np.uint8)
    pos = np.cumsum(counts) - counts
    # inserted base goes first, then the (possibly substituted) original
    ins_idx = pos[is_ins]
    res[ins_idx] = ACGT[rng.integers(0, 4, size=len(ins_idx))]
    keep_idx = (pos + is_ins.astype(np.int64))[keep]
    res[keep_idx] = out[keep]
    return res


def fixed_len(rng, seq, n):
    """Truncate or pad (random bases) to exactly n."""
    if len(seq) >= n:
        return seq[:n]
    return np.concatenate([seq, random_dna(rng, n - len(seq))])


def short_pairs(seed, n, length=150, sub=0.05, indel=0.01):
    """BASELINE config 2 shape: target = uniform ACGT, query = target with 5% substitutions
    + 1% indels, both exactly `length` long. Returns (qbuf, qoff, tbuf, toff)."""
    rng = np.random.default_rng(seed)
    tbuf = random_dna(rng, n * length)
    # vectorised mutation over the whole batch, then re-cut into fixed-length queries
    qbuf = tbuf.copy()
    r = rng.random(n * length)
    s = r < sub
    qbuf[s] = ACGT[rng.integers(0, 4, size=int(s.sum()))]
    # indels: shift the tail of a read by one base at ~indel rate (cheap but real indels)
    q2 = qbuf.reshape(n, length).copy()
    n_ev = rng.poisson(indel * length, size=n)
    cols = np.arange(length, dtype=np.int32)[None, :]
    for rnd in range(int(n_ev.max()) if n else 0):
        rows = np.nonzero(n_ev > rnd)[0]
        for c0 in range(0, len(rows), 1 << 17):       # bounded temporaries
            rr = rows[c0:c0 + (1 << 17)]
            p = rng.integers(0, length - 1, size=len(rr)).astype(np.int32)[:, None]
            is_del = rng.random(len(rr)) < 0.5
            fill = ACGT[rng.integers(0, 4, size=len(rr))]
            sub = q2[rr]
            # deletion at p: everything right of p moves left, a fresh base enters at the end
            idx_del = np.minimum(cols + (cols >= p), length - 1)
            d = np.take_along_axis(sub, idx_del, axis=1)
            d[:, -1] = fill
            # insertion at p: everything from p moves right, a fresh base lands at p
            idx_ins = cols - (cols > p)
            i_ = np.take_along_axis(sub, idx_ins, axis=1)
            i_[np.arange(len(rr)), p[:, 0]] = fill
            q2[rr] = np.where(is_del[:, None], d, i_)
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(length))
    return np.append(q2.reshape(-1), np.uint8(0)), off, np.append(tbuf, np.uint8(0)), off.copy()


def ont_like_pairs(seed, n, mean_len=8000, err=0.12, min_len=1000, max_len=40000, fixed=None):
    """Long noisy pairs: target uniform ACGT of log-normal length, query = target at `err`
    error split 40% del / 40% ins / 20% sub (indel-heavy). Returns lists of uint8 arrays."""
    rng = np.random.default_rng(seed)
    qs, ts = [], []
    for _ in range(n):
        if fixed:
            L = fixed
        else:
            L = int(np.clip(rng.lognormal(np.log(mean_len) - 0.125, 0.5), min_len, max_len))
        t = random_dna(rng, L)
        q = mutate(rng, t, sub=err * 0.2, ins=err * 0.4, dele=err * 0.4)
        qs.append(q)
        ts.append(t)
    return qs, ts


def pack_arrays(arrs):
    off = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        off[1:] = np.cumsum([len(a) for a in arrs], dtype=np.uint64)
    buf = np.concatenate(list(arrs) + [np.zeros(1, dtype=np.uint8)]) if arrs else np.zeros(1, dtype=np.uint8)
    return buf.astype(np.uint8), off
