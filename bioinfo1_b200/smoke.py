"""smoke(): one small invocation of the hot path on cuda:0, checked against the CPU oracle.
(The oracle is test infrastructure: this file and tests/ are the only importers.)"""
import os
import sys

import numpy as np


def run():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests"))
    from cpu_checkers import load_oracle
    import seqgen
    from . import capi

    oracle = load_oracle()
    ctx = capi.Context(0)
    qb, qo, tb, to = seqgen.short_pairs(1, 256)
    for typ in (0, 1, 2):
        score, tbeg, cig, coff = ctx.align_packed(qb, qo, tb, to, typ)
        for k in range(0, 256, 5):
            q = qb[int(qo[k]):int(qo[k + 1])].tobytes()
            t = tb[int(to[k]):int(to[k + 1])].tobytes()
            exp = oracle.align(q, t, typ)
            got = (int(score[k]), int(tbeg[k]), cig[int(coff[k]):int(coff[k + 1])].tobytes())
            assert got == exp, (typ, k, got, exp)
    # a uniform batch large enough for the thread-per-pair kernel (K1), a long pair for the stripe kernel (K3)
    qb, qo, tb, to = seqgen.short_pairs(3, 8192)
    score, tbeg, cig, coff = ctx.align_packed(qb, qo, tb, to, 0)
    for k in range(0, 8192, 257):
        exp = oracle.align(qb[int(qo[k]):int(qo[k + 1])].tobytes(), tb[int(to[k]):int(to[k + 1])].tobytes(), 0)
        assert (int(score[k]), int(tbeg[k]), cig[int(coff[k]):int(coff[k + 1])].tobytes()) == exp, ("K1", k)
    lq, lt = seqgen.ont_like_pairs(4, 2, mean_len=3000, min_len=2500, max_len=3500)
    for (got, q, t) in zip(ctx.align([x.tobytes() for x in lq], [x.tobytes() for x in lt], 2), lq, lt):
        assert got == oracle.align(q.tobytes(), t.tobytes(), 2), "K3"
    rng = np.random.default_rng(2)
    seq = seqgen.random_dna(rng, 5000).tobytes()
    (h, p, f), = ctx.minimize([seq], 15, 5)
    eh, ep, ef = oracle.minimize(seq, 15, 5, True)
    assert np.array_equal(h, eh) and np.array_equal(p, ep) and np.array_equal(f, ef)
    print(f"smoke ok: 3 x 256 + 8192 short alignments, 2 long ones + {len(h)} minimizer tuples bit-exact vs oracle; "
          f"kernel launches = {ctx.counter('kernel_launches')}")
    ctx.close()
