"""GPU-box report over the five BASELINE.json configs (C1..C5): what each costs through this implementation.
C4 is bench.py's line; this script covers C1, C2, C3 and C5 and prints one JSON object per config.
usage: python scripts/configs_report.py [--c5-loci N]"""
import argparse, io, json, os, subprocess, sys, tempfile, time
from contextlib import redirect_stdout
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import polyfasta_b200 as pf
from polyfasta_b200 import api, cli, synth

ap = argparse.ArgumentParser()
ap.add_argument("--c5-loci", type=int, default=4000)
ap.add_argument("--c3-sites", type=int, default=3_000_000)
args = ap.parse_args()
EX = os.path.join(ROOT, "tests", "golden", "example_theta_0.01")


def cli_time(argv, reps=3):
    best, text = 1e9, ""
    for _ in range(reps):
        buf = io.StringIO()
        t0 = time.perf_counter()
        with redirect_stdout(buf):
            cli.main(argv)
        best = min(best, time.perf_counter() - t0)
        text = buf.getvalue()
    return best, text


def emit(name, **kw):
    print(json.dumps({"config": name, **kw}), flush=True)


# ---- C1 / C2: the shipped example loci through the drop-in CLI (in-process, warm; and one cold subprocess) ----
t, out = cli_time(["-f", os.path.join(EX, "file1.fa")])
t0 = time.perf_counter()
p = subprocess.run([sys.executable, os.path.join(ROOT, "PolyFastA.py"), "-f", os.path.join(EX, "file1.fa")], capture_output=True, text=True)
cold = time.perf_counter() - t0
emit("C1 file1.fa (20 x 1000)", warm_ms=t * 1e3, cold_process_s=cold, row=out.strip().split("\n")[-1], same_as_cold=p.stdout == out)
t, out = cli_time(["-d", EX, "-p", "indiv1,indiv2", "--jc"])
emit("C2 --dir example_theta_0.01 (10 loci) -p indiv1,indiv2 --jc", warm_ms=t * 1e3, rows=len(out.strip().split("\n")) - 1,
     first_row=out.strip().split("\n")[1])
t, out = cli_time(["-d", EX, "-p", "pop1,pop2", "--jc"])
emit("C2 literal -p pop1,pop2 (no header matches)", warm_ms=t * 1e3, rows=len(out.strip().split("\n")) - 1, first_row=out.strip().split("\n")[1])

# ---- C3: synthetic in-frame CDS alignment 2,000 x 3 Mb, two populations, --cds --jc ----
ctx = pf.default_context(0)
n, L = 2000, args.c3_sites
d = torch.empty((n, L), dtype=torch.uint8, device="cuda")
api.synth_text_device(ctx, d.data_ptr(), L, n, 3, 50000, 10000, 0, L)
ctx.sync()
h = torch.empty((n, L), dtype=torch.uint8, pin_memory=True)
h.copy_(d)
torch.cuda.synchronize()
del d
torch.cuda.empty_cache()
pops = [list(range(n // 2)), list(range(n // 2, n))]
best = {}
for rep in range(3):
    t0 = time.perf_counter()
    a = pf.Alignment.from_host_ptr(ctx, h.data_ptr(), n, L, L)
    ctx.sync()
    t1 = time.perf_counter()
    a.set_pops(pops)
    s = a.site_stats()
    c = a.cds_stats()
    t2 = time.perf_counter()
    rows = cli.format_rows(ctx, "c3.fa", L, True, True, [("pop1", n // 2), ("pop2", n // 2)], s, c)
    t3 = time.perf_counter()
    a.free()
    for k, v in (("upload_ms", t1 - t0), ("scans_ms", t2 - t1), ("finalise_rows_ms", t3 - t2), ("total_ms", t3 - t0)):
        best[k] = min(best.get(k, 1e9), v * 1e3)
emit("C3 synthetic CDS 2000 x %d from pinned host text, 2 populations, --cds --jc" % L, **best, bases_per_s=n * L / (best["total_ms"] * 1e-3),
     ingest=ctx.ingest_stats(), row=rows[0])
del h

# ---- C5: --dir batch of loci (100 x 5 kb each) through the CLI ----
nl = args.c5_loci
tmp = tempfile.mkdtemp()
base = np.ascontiguousarray(synth.text_matrix(5, 100, 5000))
for i in range(nl):
    m = base.copy()
    m[:, (i * 7) % 5000] = np.frombuffer(b"ACGT", dtype=np.uint8)[(np.arange(100) + i) % 4]   # every locus differs a little
    with open(os.path.join(tmp, "locus%06d.fa" % i), "wb") as f:
        f.write(b"".join(b">pop%d_ind%d\n" % (1 + r % 2, r) + m[r].tobytes() + b"\n" for r in range(100)))
for extra, label in (([], "all rows"), (["-p", "pop1,pop2"], "-p pop1,pop2"), (["--cds"], "--cds (batched codon scan K4b)"),
                     (["--cds", "-p", "pop1,pop2"], "--cds -p pop1,pop2")):
    t, out = cli_time(["-d", tmp, "--jc", "-s"] + extra, reps=2)
    emit("C5 --dir %d loci of 100 x 5000, %s" % (nl, label), wall_s=t, us_per_locus=t / nl * 1e6, bases_per_s=nl * 5e5 / t,
         rows=len(out.strip().split("\n")), first_row=out.strip().split("\n")[-1])
