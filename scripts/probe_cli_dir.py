import os, sys, time, tempfile, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle as co
nloci = int(os.environ.get("NLOCI", "3000"))
d = tempfile.mkdtemp()
m = co.synth_text(5, 100, 5000)
for i in range(nloci):
    with open(os.path.join(d, "locus%05d.fa" % i), "wb") as f:
        for r in range(100):
            f.write(b">pop%d_ind%d\n" % (1 + r % 2, r) + m[r].tobytes() + b"\n")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
env = dict(os.environ, POLYFASTA_TIMING="1")
for extra in ([], ["-p", "pop1,pop2"], []):
    t0 = time.perf_counter()
    p = subprocess.run([sys.executable, os.path.join(root, "PolyFastA.py"), "-d", d, "--jc", "-s"] + extra, capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    print("CLI -d (%d loci) %s: %.2f s wall -> %.1f us/locus ; rows=%d rc=%d" % (nloci, " ".join(extra), dt, dt / nloci * 1e6, len(p.stdout.strip().split("\n")), p.returncode))
    print(p.stderr[-1500:])
t0 = time.perf_counter(); subprocess.run([sys.executable, "-c", "import polyfasta_b200; polyfasta_b200.default_context(0)"], env=dict(os.environ, PYTHONPATH=root)); print("python + ctx startup: %.2f s" % (time.perf_counter() - t0))
