"""World-size-2 (and 3) CPU tests of the multi-rank host logic over gloo: column shards + one integer all-reduce give
the whole-alignment vector bit for bit; --dir round-robin covers every locus once and rows come back in sorted order.
The per-shard numbers come from the CPU oracle here (no GPU in this tier); the same plumbing carries the kernels'
vectors on the GPU box (bench.py, tests/test_gpu_parity.py::test_column_shards_add_up)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from polyfasta_b200 import parallel, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, L, q):
    from oracle import c_oracle as co
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c0, c1 = parallel.shard_columns(L, world, rank)
        mat = np.ascontiguousarray(synth.text_matrix(7, n, L, 200000, 100000, c0, c1))
        pops = [list(range(n)), list(range(0, n, 2))]
        vec = []
        cds = []
        for rows in pops:
            st = co.site_stats(mat, rows)
            vec += [st["S"], st["H"]] + st["sfs"]
        t = torch.tensor(vec, dtype=torch.int64)
        dist.all_reduce(t)   # what cli.sharded_alignment_rows does when the fused NVLink exchange is not available
        # the CLI's own planning: sorted paths -> units -> owner ranks -> gathered rows
        loci = sorted("file%d.fa" % i for i in range(1, 12))
        units = parallel.plan_units(loci, [100] * len(loci), 1 << 20, 2, 1 << 30)
        mine = [(ui, ["row for " + name for name in ps]) for ui, (kind, ps) in enumerate(units)
                if parallel.chunk_owner(ui, world) == rank]
        rows = parallel.gather_rows(mine)
        if rank == 0:
            q.put((t.tolist(), rows))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_column_shards_allreduce_gloo(world):
    from oracle import c_oracle as co
    n, L = 24, 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, L, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, rows = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mat = np.ascontiguousarray(synth.text_matrix(7, n, L, 200000, 100000))
    want = []
    for r in [list(range(n)), list(range(0, n, 2))]:
        st = co.site_stats(mat, r)
        want += [st["S"], st["H"]] + st["sfs"]
    assert got == want
    loci = sorted("file%d.fa" % i for i in range(1, 12))
    assert [r for _, part in rows for r in part] == ["row for " + name for name in loci]
    assert [i for i, _ in rows] == list(range(6))


def test_plan_units():
    paths = ["a", "b", "BIG", "c", "d", "e"]
    sizes = [10, 10, 5000, 10, 10, 10]
    assert parallel.plan_units(paths, sizes, 1000, 2, 1 << 30) == [("chunk", ["a", "b"]), ("all", ["BIG"]), ("chunk", ["c", "d"]), ("chunk", ["e"])]
    assert parallel.plan_units(paths, sizes, None, 512, 25) == [("chunk", ["a", "b"]), ("chunk", ["BIG"]), ("chunk", ["c", "d"]), ("chunk", ["e"])]
    assert parallel.plan_units([], [], 1000, 2, 100) == []
    assert [parallel.chunk_owner(j, 3) for j in range(5)] == [0, 1, 2, 0, 1]


def test_shard_columns_properties():
    for total in (0, 1, 2, 3, 10, 999, 1000, 3001, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                b, e = parallel.shard_columns(total, world, r)
                assert b == prev and b % 3 == 0 and e >= b
                prev = e
            assert prev == total
