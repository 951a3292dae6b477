"""K2 time for several alignment shapes under PFA_SITE_ITER_MAX (lanes per site vs chunks per lane)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polyfasta_b200 as pf
ctx = pf.Context(0)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
def timed(fn, reps=10):
    fn(); stream.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): fn()
    b.record(stream); stream.synchronize()
    return a.elapsed_time(b) / reps
with torch.cuda.stream(stream):
    for n, L in ((100, 40_000_000), (500, 10_000_000), (2000, 3_000_000), (5000, 2_000_000), (10000, 1_000_000)):
        aln = pf.Alignment.synthetic(ctx, n, L, 3)
        out = torch.zeros(aln.site_len(), dtype=torch.int64, device="cuda")
        res = []
        for im in (5,):
            os.environ["PFA_SITE_ITER_MAX"] = str(im)
            ms = timed(lambda: aln.site_stats_device(out.data_ptr()))
            res.append("iter<=%d: %.3f ms %4.0f GB/s" % (im, ms, aln.packed_bytes / 3 * 2 / ms / 1e6))
        print("n=%5d L=%9d (%.2f GB planes): %s" % (n, L, aln.packed_bytes / 3 * 2 / 1e9, " | ".join(res)), flush=True)
        aln.free()
