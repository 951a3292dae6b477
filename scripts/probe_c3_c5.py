"""GPU-box probe: K4/K2 on the C3 shape, K3 at n=2000, and the per-locus overhead of the C5 shape.  Not part of the product."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import polyfasta_b200 as pf
from polyfasta_b200 import synth

ctx = pf.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)

def timed(fn, reps=5):
    fn(); stream.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream); stream.synchronize()
    return a.elapsed_time(b) / reps

with torch.cuda.stream(stream):
    n, L = 2000, 3_000_000
    aln = pf.Alignment.synthetic(ctx, n, L, 3)
    pops = [list(range(0, 1000)), list(range(1000, 2000))]
    for label, p in (("1 pop (all rows)", None), ("2 pops", pops)):
        aln.set_pops(p)
        k = aln.num_pops
        out_s = torch.zeros(aln.site_len(), dtype=torch.int64, device="cuda")
        out_c = torch.zeros(k * 71, dtype=torch.int64, device="cuda")
        ms2 = timed(lambda: aln.site_stats_device(out_s.data_ptr()))
        ms4 = timed(lambda: aln.cds_stats_device(out_c.data_ptr()))
        gb = n * L * 0.25 / 1e9
        print("C3 %s: K2 %.3f ms (%.0f GB/s)  K4 %.3f ms (%.0f GB/s)  -> %.2e bases/s for K2+K4" %
              (label, ms2, gb / ms2 * 1e3, ms4, gb / ms4 * 1e3, n * L / ((ms2 + ms4) * 1e-3)))
    aln.free()
    if "--c3-only" in sys.argv:
        sys.exit(0)
    # K3 on a 2000 x 100 kb slice
    a = pf.Alignment.synthetic(ctx, 2000, 100_000, 3)
    out_p = torch.zeros(1, dtype=torch.int64, device="cuda")
    D = torch.zeros((2000, 2000), dtype=torch.int32, device="cuda")
    from polyfasta_b200._lib import lib, check
    import ctypes
    f = lambda: check(lib().pfa_pairwise_device(a.handle, ctypes.c_void_p(out_p.data_ptr()), ctypes.c_void_p(D.data_ptr())), ctx.handle)
    ms = timed(f, 3)
    pairs = 2000 * 1999 / 2
    print("K3 2000 x 100kb: %.2f ms -> %.2e pair-words/s ; sum=%d H/2=%d" % (ms, pairs * (100_000 / 32) / (ms * 1e-3), int(out_p[0]), a.site_stats()[0]["H"] // 2))
    a.free()
    # C5 shape: many loci of 100 x 5 kb, one after the other through the API
    mats = [np.ascontiguousarray(synth.text_matrix(5 + i, 100, 5000)) for i in range(50)]
    t = time.perf_counter()
    reps = 4
    for r in range(reps):
        for m in mats:
            al = pf.Alignment.from_rows(ctx, m)
            s = al.site_stats()
            al.free()
    dt = (time.perf_counter() - t) / (reps * len(mats))
    print("C5 locus (100 x 5kb) through from_rows + site_stats + free: %.1f us per locus -> %.2e bases/s" % (dt * 1e6, 100 * 5000 / dt))
