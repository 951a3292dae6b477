"""GPU-box probe: the C5 shape (loci of 100 x 5 kb) through the batched path, from in-memory rows and from FASTA files."""
import os, sys, time, tempfile, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import polyfasta_b200 as pf
from polyfasta_b200 import api
from oracle import c_oracle as co

ctx = pf.Context(0)
nloci = int(os.environ.get("NLOCI", "2000"))
mats = [co.synth_text(5 + i, 100, 5000) for i in range(nloci)]
batch = api.Batch(ctx)
for rep in range(3):
    batch.clear()
    t0 = time.perf_counter()
    for m in mats[:1000]:
        batch.add_rows(m)
    t1 = time.perf_counter()
    batch.run(True)
    t2 = time.perf_counter()
    print("batch of 1000 loci from rows: add %.1f ms, run %.1f ms -> %.1f us/locus, %.2e bases/s" %
          ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) / 1000 * 1e6, 1000 * 5e5 / (t2 - t0)))
want = co.site_stats(mats[7])
got = batch.result(7)
assert (got["S"], got["H"]) == (want["S"], want["H"])
# files -> CLI
d = tempfile.mkdtemp()
for i, m in enumerate(mats):
    with open(os.path.join(d, "locus%05d.fa" % i), "wb") as f:
        for r in range(100):
            f.write(b">pop%d_ind%d\n" % (1 + r % 2, r) + m[r].tobytes() + b"\n")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for extra in ([], ["-p", "pop1,pop2"]):
    t0 = time.perf_counter()
    p = subprocess.run([sys.executable, os.path.join(root, "PolyFastA.py"), "-d", d, "--jc", "-s"] + extra, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    print("CLI -d (%d loci) %s: %.2f s wall -> %.1f us/locus, %.2e bases/s ; rows=%d rc=%d" %
          (nloci, " ".join(extra), dt, dt / nloci * 1e6, nloci * 5e5 / dt, len(p.stdout.strip().split("\n")), p.returncode))
    if p.returncode:
        print(p.stderr[-2000:])
print(p.stdout.split("\n")[0])
