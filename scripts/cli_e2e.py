"""GPU-box probe of the drop-in CLI on one large FASTA file (C3-like: CDS, two populations): where does the wall time go?
usage: python scripts/cli_e2e.py [n] [L]"""
import io, os, sys, time
from contextlib import redirect_stdout
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import polyfasta_b200 as pf
from polyfasta_b200 import api, cli

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300_000
ctx = pf.default_context(0)
d = torch.empty((n, L), dtype=torch.uint8, device="cuda")
api.synth_text_device(ctx, d.data_ptr(), L, n, 3, 50000, 10000, 0, L)
ctx.sync()
mat = d.cpu().numpy()
del d
path = "/tmp/c3_like.fa"
t0 = time.perf_counter()
wrap = int(os.environ.get("WRAP", "0"))   # 0: one line per sequence; else that many bases per line
with open(path, "wb") as f:
    for i in range(n):
        f.write((">pop%d_indiv%d\n" % (1 if i < n // 2 else 2, i)).encode())
        if wrap and L % wrap == 0:
            f.write(np.concatenate([mat[i].reshape(-1, wrap), np.full((L // wrap, 1), 10, np.uint8)], axis=1).tobytes())
        else:
            f.write(mat[i].tobytes())
            f.write(b"\n")
print("wrote %s (%.2f GB) in %.1f s" % (path, os.path.getsize(path) / 1e9, time.perf_counter() - t0))
for rep in range(2):
    t0 = time.perf_counter(); fa = pf.Fasta.from_file(path); t1 = time.perf_counter()
    aln = pf.Alignment.from_fasta(ctx, fa); ctx.sync(); t2 = time.perf_counter()
    aln.set_pops([list(range(n // 2)), list(range(n // 2, n))])
    s = aln.site_stats(); c = aln.cds_stats(); t3 = time.perf_counter()
    aln.free(); fa.close()
    print("rep %d: parse %.1f ms (%.2f GB/s), upload %.1f ms (%.1f GB/s) %s, set_pops+K2+K4 %.1f ms" %
          (rep, (t1 - t0) * 1e3, n * L / (t1 - t0) / 1e9, (t2 - t1) * 1e3, n * L / (t2 - t1) / 1e9, ctx.ingest_stats(), (t3 - t2) * 1e3))
for args in (["-f", path, "--cds", "--jc", "-p", "pop1,pop2"], ["-f", path, "--jc", "-p", "pop1,pop2"]):
    buf = io.StringIO()
    t0 = time.perf_counter()
    with redirect_stdout(buf):
        cli.main(args)
    dt = time.perf_counter() - t0
    print("CLI %s: %.1f ms -> %.2e bases/s\n%s" % (" ".join(args[2:]), dt * 1e3, n * L / dt, buf.getvalue().strip()))
if os.environ.get("CLI_PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile()
    buf = io.StringIO()
    pr.enable()
    with redirect_stdout(buf):
        cli.main(["-f", path, "--cds", "--jc", "-p", "pop1,pop2"])
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
    pr = cProfile.Profile()
    pr.enable()
    with redirect_stdout(buf):
        cli.main(["-f", path, "--jc", "-p", "pop1,pop2"])
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
os.remove(path)
