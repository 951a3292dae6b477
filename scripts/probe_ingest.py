"""upload (ingest) throughput of one large pinned text matrix under the plain and the hybrid path"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polyfasta_b200 as pf
from polyfasta_b200 import api
n, cols = 10000, int(os.environ.get("COLS", 600000))
ctx = pf.Context(0)
ld = (cols + 255) // 256 * 256
d = torch.empty((n, ld), dtype=torch.uint8, device="cuda")
api.synth_text_device(ctx, d.data_ptr(), ld, n, 4, 50000, 10000, 0, cols)
ctx.sync()
h = torch.empty((n, ld), dtype=torch.uint8, pin_memory=True)
h.copy_(d); torch.cuda.synchronize(); del d
def run(label, env):
    for k in ("PFA_INGEST_HYBRID", "PFA_INGEST_CHUNK_MB", "PFA_HOST_THREADS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        a = pf.Alignment.from_host_ptr(ctx, h.data_ptr(), n, cols, ld)
        ctx.sync()
        best = min(best, time.perf_counter() - t0)
        st = ctx.ingest_stats()
        a.free()
    print("%-28s %.1f ms  %.1f GB/s text  %s" % (label, best * 1e3, n * cols / best / 1e9, st), flush=True)
run("plain", {"PFA_INGEST_HYBRID": "0"})
for t in (16, 15, 12, 8):
    run("hybrid threads=%d" % t, {"PFA_HOST_THREADS": str(t)})
for mb in (16, 32, 128):
    run("hybrid 14 thr chunk %d MB" % mb, {"PFA_HOST_THREADS": "14", "PFA_INGEST_CHUNK_MB": str(mb)})
run("plain again", {"PFA_INGEST_HYBRID": "0"})
# the same alignment with 1 % gaps / N: packed with a validity bitmap on AVX-512 VBMI hosts, raw otherwise
import numpy as np
hv = h.numpy()
rng = np.random.default_rng(0)
idx_r = rng.integers(0, n, 6_000_000); idx_c = rng.integers(0, cols, 6_000_000)
hv[idx_r, idx_c] = np.frombuffer(b"-N", dtype=np.uint8)[rng.integers(0, 2, 6_000_000)]
run("gaps+N: plain", {"PFA_INGEST_HYBRID": "0"})
run("gaps+N: hybrid", {})
