import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polyfasta_b200 as pf
ctx = pf.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
def timed(fn, reps=20):
    fn(); stream.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream); stream.synchronize()
    return a.elapsed_time(b) / reps
with torch.cuda.stream(stream):
    for n, L in [(10000, 2_000_000), (2000, 3_000_000)]:
        aln = pf.Alignment.synthetic(ctx, n, L, 4)
        aln.poke_gaps(4, 100)
        out = torch.zeros(aln.site_len(), dtype=torch.int64, device="cuda")
        for bits in (0, 1, 2, 3, 4, 6):
            os.environ["PFA_PROBE_BITS"] = str(bits)
            print("%d x %d 100 ppm bits=%d (1: no cell fetch, 2: no gap arithmetic, 4: whole v in flagged blocks): K2 %.3f" % (n, L, bits, timed(lambda: aln.site_stats_device(out.data_ptr()))), flush=True)
        aln.free()
