"""GPU tests of the in-kernel NVLink exchange (pfa_xchg): K2 / K4 fused with the sum over column shards.

The ranks of one job normally live in one process per GPU (tests/test_gpu_xchg.py::test_two_processes_two_gpus, skipped on a
one-GPU box); the exchange protocol itself -- slot rotation, flags, zeroing, vectors of changing length -- is also driven
with several ranks inside ONE process on ONE device (every rank its own context, stream and symmetric buffer), which
runs on any GPU box.  Everything is compared bit for bit with the unsharded scan and with the C oracle."""
import os
import socket

import numpy as np
import pytest
import torch

import polyfasta_b200 as pf
from polyfasta_b200 import api, parallel
from oracle import c_oracle as co
from test_gpu_parity import _random_text, _upper

pytestmark = pytest.mark.gpu


def _whole(ctx, text, pops):
    a = pf.Alignment.from_rows(ctx, text)
    a.set_pops(pops)
    s, c = a.site_stats(), a.cds_stats()
    a.free()
    return np.concatenate([np.array([x["S"], x["H"]] + x["sfs"], dtype=np.int64) for x in s]), np.concatenate([x["raw"] for x in c])


def test_single_rank_exchange_equals_plain_scan():
    """world = 1: the fused epilogue (self push, flag, copy) must reproduce pfa_site_stats / pfa_cds_stats, launch after
    launch (both slots, vectors that grow and shrink as the populations change)"""
    ctx = pf.Context(0)
    rng = np.random.default_rng(11)
    n, L = 300, 2002
    text = _random_text(rng, n, L, p_junk=0.02)
    x = parallel.connect_exchange(ctx, 4096)
    for pops in ([list(range(n))], [list(range(n)), list(range(0, n, 2)), list(range(7, 90))], [list(range(20))], [list(range(n))]):
        want_s, want_c = _whole(ctx, text, pops)
        a = pf.Alignment.from_rows(ctx, text)
        a.set_pops(pops)
        out_s = torch.full((a.site_len(),), -1, dtype=torch.int64, device="cuda")
        out_c = torch.full((api.PFA_CDS_LEN * len(pops),), -1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        for _ in range(3):
            a.site_stats_xchg(x, out_s.data_ptr())
            a.cds_stats_xchg(x, out_c.data_ptr())
            ctx.sync()
            assert np.array_equal(out_s.cpu().numpy(), want_s)
            assert np.array_equal(out_c.cpu().numpy(), want_c)
        a.free()
    assert not x.timed_out()
    x.close()
    ctx.close()


@pytest.mark.parametrize("world,L", [(2, 4099), (3, 4099), (8, 4099), (2, 5), (4, 10), (8, 20)])
def test_ranks_in_one_process(world, L):
    """`world` ranks on one device, each with its own context / stream / buffer, column shards of one alignment:
    every rank must end up with the whole alignment's vector.  L = 4099: --cds with a trailing partial codon on the last
    shard; L < 3 * world: all ranks but the last hold EMPTY shards (the exchange-only launch)"""
    rng = np.random.default_rng(100 + world)
    n = 260
    text = _random_text(rng, n, L, p_junk=0.02)
    up = _upper(text)
    pops = [list(range(n)), list(range(1, n, 3))]
    ctxs = [pf.Context(0) for _ in range(world)]
    want_s, want_c = _whole(ctxs[0], text, pops)
    oracle = co.site_stats(up, pops[1])
    off1 = 2 + n // 2
    assert (int(want_s[off1]), int(want_s[off1 + 1])) == (oracle["S"], oracle["H"])
    wc = co.cds_stats(up, pops[1])
    assert [int(x) for x in want_c[api.PFA_CDS_LEN: api.PFA_CDS_LEN + 6]] == [wc[k] for k in ("nstops", "missing", "S_s", "H_s", "S_n", "H_n")]
    if L < 3 * world:
        assert parallel.shard_columns(L, world, 0) == (0, 0)
    xs = [api.Exchange(c, 2048) for c in ctxs]
    api.Exchange.connect_local(xs)
    shards = []
    for r in range(world):
        c0, c1 = parallel.shard_columns(L, world, r)
        a = pf.Alignment.from_rows(ctxs[r], text, c0, c1)
        a.set_pops(pops)
        shards.append(a)
    out_s = [torch.full((shards[0].site_len(),), -1, dtype=torch.int64, device="cuda") for _ in range(world)]
    out_c = [torch.full((api.PFA_CDS_LEN * len(pops),), -1, dtype=torch.int64, device="cuda") for _ in range(world)]
    torch.cuda.synchronize()
    for it in range(4):
        order = range(world) if it % 2 == 0 else reversed(range(world))   # the launch order must not matter
        for r in order:
            shards[r].site_stats_xchg(xs[r], out_s[r].data_ptr())
        for r in range(world):
            shards[r].cds_stats_xchg(xs[r], out_c[r].data_ptr())
        for c in ctxs:
            c.sync()
        for r in range(world):
            assert np.array_equal(out_s[r].cpu().numpy(), want_s), (it, r)
            assert np.array_equal(out_c[r].cpu().numpy(), want_c), (it, r)
    # K2 + K4 with ONE exchange (what --cds needs per alignment): [site vector | codon vectors]
    out_b = [torch.full((len(want_s) + len(want_c),), -1, dtype=torch.int64, device="cuda") for _ in range(world)]
    torch.cuda.synchronize()
    for it in range(2):
        for r in range(world):
            shards[r].site_cds_stats_xchg(xs[r], out_b[r].data_ptr())
        for c in ctxs:
            c.sync()
        for r in range(world):
            assert np.array_equal(out_b[r].cpu().numpy(), np.concatenate([want_s, want_c])), (it, r)
    # K3 over the column shards: the pairwise sums of every population, summed over the ranks, equal the whole alignment's
    whole = pf.Alignment.from_rows(ctxs[0], text)
    whole.set_pops(pops)
    want_pw = whole.pairwise()
    whole.free()
    assert want_pw[1] == co.pairwise_sum(up, pops[1])
    out_p = [torch.full((len(pops),), -1, dtype=torch.int64, device="cuda") for _ in range(world)]
    torch.cuda.synchronize()
    for r in range(world):
        shards[r].pairwise_xchg(xs[r], out_p[r].data_ptr())
    for c in ctxs:
        c.sync()
    for r in range(world):
        assert out_p[r].cpu().tolist() == want_pw, r
    # a vector produced by another kernel, summed in place
    bufs = [torch.arange(5, dtype=torch.int64, device="cuda") * (r + 1) for r in range(world)]
    torch.cuda.synchronize()
    for r in range(world):
        xs[r].allreduce(bufs[r].data_ptr(), 5)
    for c in ctxs:
        c.sync()
    tot = sum(range(1, world + 1))
    for r in range(world):
        assert bufs[r].cpu().tolist() == [i * tot for i in range(5)]
    assert not any(x.timed_out() for x in xs)
    for a in shards:
        a.free()
    for x in xs:
        x.close()
    for c in ctxs:
        c.close()


def test_missing_rank_times_out_instead_of_hanging():
    ctx = pf.Context(0)
    ctx2 = pf.Context(0)
    xs = [api.Exchange(ctx, 64), api.Exchange(ctx2, 64)]
    api.Exchange.connect_local(xs)
    buf = torch.ones(4, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    xs[0].set_timeout_ms(300)
    xs[0].allreduce(buf.data_ptr(), 4)   # rank 1 never calls
    assert xs[0].timed_out()
    assert not xs[0].timed_out()   # reported once, then cleared
    for x in xs:
        x.close()
    ctx.close()
    ctx2.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mp_worker(rank, world, port, n, L, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = pf.Context(rank)
        rng = np.random.default_rng(2024)
        text = _random_text(rng, n, L, p_junk=0.02)
        pops = [list(range(n)), list(range(0, n, 2))]
        x = parallel.connect_exchange(ctx, 4096)
        c0, c1 = parallel.shard_columns(L, world, rank)
        a = pf.Alignment.from_rows(ctx, text, c0, c1)
        a.set_pops(pops)
        out_s = torch.full((a.site_len(),), -1, dtype=torch.int64, device="cuda")
        out_c = torch.full((api.PFA_CDS_LEN * len(pops),), -1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        res = []
        for _ in range(5):
            a.site_stats_xchg(x, out_s.data_ptr())
            a.cds_stats_xchg(x, out_c.data_ptr())
            ctx.sync()
            res.append((out_s.cpu().numpy().copy(), out_c.cpu().numpy().copy()))
        bad = x.timed_out()
        dist.barrier()
        a.free()
        x.close()
        q.put((rank, bad, res))
    finally:
        dist.destroy_process_group()


def test_two_processes_two_gpus():
    """the real thing: one process per GPU, buffers mapped through CUDA IPC, reductions over NVLink"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world, n, L = 2, 500, 30001
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_mp_worker, args=(r, world, port, n, L, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
    ctx = pf.default_context(0)
    rng = np.random.default_rng(2024)
    text = _random_text(rng, n, L, p_junk=0.02)
    want_s, want_c = _whole(ctx, text, [list(range(n)), list(range(0, n, 2))])
    for rank, bad, res in got:
        assert not bad
        for s, c in res:
            assert np.array_equal(s, want_s) and np.array_equal(c, want_c), rank


def _cli_worker(rank, world, port, argv_list, shard_min, q):
    import io
    from contextlib import redirect_stdout
    # PFA_BIG_FILE_MIN: the two large files are parsed in place (mapped), so rank 0 parses them once and the other rank adopts
    # the exported row layout (RankGroup.parse_once)
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank), "MASTER_ADDR": "127.0.0.1",
                       "MASTER_PORT": str(port), "PFA_BIG_FILE_MIN": "1000000"})
    from polyfasta_b200 import cli
    cli.SHARD_MIN_BYTES = shard_min
    outs = []
    for argv in argv_list:
        buf = io.StringIO()
        with redirect_stdout(buf):
            cli.main(argv)
        outs.append(buf.getvalue())
    q.put((rank, outs))


def test_cli_under_two_ranks(tmp_path):
    """the drop-in CLI launched as one process per GPU: --dir loci round-robin over the ranks, a large file column-sharded
    with the exchange fused into K2 / K4; rank 0 prints exactly what one process prints"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import io
    import shutil
    from contextlib import redirect_stdout
    import torch.multiprocessing as mp
    from conftest import GOLDEN
    from polyfasta_b200 import cli
    d = tmp_path / "loci"
    shutil.copytree(os.path.join(GOLDEN, "example_theta_0.01"), d)
    rng = np.random.default_rng(8)
    n, L = 60, 30000
    text = _random_text(rng, n, L, p_junk=0.002)
    with open(d / "file5_big.fa", "wb") as f:      # sorts between file5.fa and file6.fa
        for i in range(n):
            f.write((">indiv%d\n" % i).encode() + text[i].tobytes() + b"\n")
    # a second large file whose length is not a multiple of 3: the trailing partial codon belongs to the LAST rank's shard
    text2 = _random_text(rng, 45, 30001, p_junk=0.001)
    with open(tmp_path / "odd_big.fa", "wb") as f:
        for i in range(45):
            f.write((">pop%d_%d\n" % (i % 2, i)).encode() + text2[i].tobytes() + b"\n")
    argvs = [["-d", str(d), "-p", "indiv1,indiv2,nobody", "--jc"], ["-d", str(d), "--cds", "--jc"],
             ["-f", str(d / "file5_big.fa"), "--cds", "-p", "indiv1,indiv3"],
             ["-f", str(tmp_path / "odd_big.fa"), "--cds", "--jc", "-p", "pop0,pop1"]]
    want = []
    for argv in argvs:
        buf = io.StringIO()
        with redirect_stdout(buf):
            cli.main(argv)
        want.append(buf.getvalue())
    assert "file5_big.fa" in want[0] and want[2].count("\n") == 3
    assert "not a multiple of 3: odd_big.fa" in want[3] and want[3].count("\n") == 4
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_cli_worker, args=(r, 2, port, argvs, 1_000_000, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(60)
    assert got[0] == want
    assert all(o == "" for o in got[1])
