"""CPU, world_size 2 over gloo: the N > 1 host logic — table broadcast and read sharding (mates kept together).
The per-shard compute stand-in is the CPU oracle; what is checked is that shard results concatenate to the
unsharded result and that every rank ends up with the same table."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import rcx_testlib as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, prefix, k, resq):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from rcorrector_b200.sharding import broadcast_table, shard_batch, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    codes = counts = None
    if rank == 0:
        codes, counts = T.load_table_arrays(prefix + ".tab")
    t_codes, t_counts = broadcast_table(codes, counts)
    table = T.OracleTable(k)
    table.put(t_codes.numpy().view(np.uint64), t_counts.numpy())
    seq = np.fromfile(prefix + ".seq", dtype=np.uint8)
    qual = np.fromfile(prefix + ".qual", dtype=np.uint8)
    off = np.fromfile(prefix + ".off", dtype=np.uint32)
    n = len(off) - 1
    lo, hi = shard_bounds(n, world, rank, True)
    assert lo % 2 == 0 and (hi % 2 == 0 or rank == world - 1)
    s, q, o = shard_batch(seq, qual, off, lo, hi)
    b = T.Batch(None, s, q, o)
    so, corr, l, m, h = T.oracle_correct(table, b, k, T.INTERLEAVED, 0.01, 48, 4, threads=2)
    resq.put((rank, lo, hi, len(table), so.tobytes(), corr.tobytes(), l.tobytes(), m.tobytes(), h.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_rank(workdir):
    k = 23
    d = T.synth(workdir, "shard", k=k, paired=True, seed=41, ntx=15, reads=1501, len=100, err=0.015)
    prefix = d["bin"]
    # single-rank result
    seq = np.fromfile(prefix + ".seq", dtype=np.uint8)
    qual = np.fromfile(prefix + ".qual", dtype=np.uint8)
    off = np.fromfile(prefix + ".off", dtype=np.uint32)
    codes, counts = T.load_table_arrays(d["tab"])
    table = T.OracleTable(k)
    table.put(codes, counts)
    full = T.oracle_correct(table, T.Batch(None, seq, qual, off), k, T.INTERLEAVED, 0.01, 48, 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, prefix, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == len(off) - 1
    assert res[0][3] == res[1][3] == len(table)
    for j, name in enumerate(["seq", "corr", "l", "m", "h"]):
        cat = res[0][4 + j] + res[1][4 + j]
        assert cat == np.asarray(full[j]).tobytes(), name


@pytest.mark.parametrize("n,world,inter", [(10, 4, True), (7, 2, False), (0, 8, True), (2, 8, True), (1001, 8, False)])
def test_shard_bounds_partition(n, world, inter):
    sys.path.insert(0, ROOT)
    from rcorrector_b200.sharding import shard_bounds

    prev = 0
    for r in range(world):
        lo, hi = shard_bounds(n, world, r, inter)
        assert lo == prev and hi >= lo
        if inter and r < world - 1:
            assert hi % 2 == 0
        prev = hi
    assert prev == n
