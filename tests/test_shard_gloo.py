"""CPU (gloo, world_size 2) test of the multi-rank host logic: partition by cost, per-rank work on
the shard, gather back into input order. The per-rank 'engine' here is the CPU oracle standing in
for the GPU library (test infrastructure) -- the logic under test is bioinfo1_b200/shard.py."""
import os
import random
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, pairs, out_q):
    import torch.distributed as dist
    from bioinfo1_b200 import shard
    from cpu_checkers import load_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [len(q) * len(t) for q, t in pairs]
    mine = shard.partition(costs, world)[rank]
    oracle = load_oracle()
    local = [oracle.align(pairs[i][0], pairs[i][1], 2) for i in mine]
    full = shard.gather_in_order(mine, local, len(pairs))
    slowest = shard.max_over_ranks(10.0 + rank)
    if rank == 0:
        out_q.put((full, slowest, [len(x) for x in shard.partition(costs, world)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_single_process():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_checkers import load_oracle
    rng = random.Random(4)
    pairs = [(bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 60))),
              bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 60)))) for _ in range(41)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, slowest, sizes = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    oracle = load_oracle()
    assert full == [oracle.align(a, b, 2) for a, b in pairs]
    assert slowest == 11.0
    assert sum(sizes) == len(pairs) and min(sizes) > 0


def test_partition_balances_cost():
    from bioinfo1_b200 import shard
    rng = np.random.default_rng(0)
    costs = rng.lognormal(10, 1, size=1000)
    parts = shard.partition(costs, 8)
    loads = np.array([costs[p].sum() for p in parts])
    assert sorted(np.concatenate(parts).tolist()) == list(range(1000))
    assert loads.max() / loads.mean() < 1.02
