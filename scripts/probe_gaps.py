"""GPU-box probe: the site scan on a C4-shaped shard with sparse gaps, validity flags on / off, slot variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polyfasta_b200 as pf

ctx = pf.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)


def timed(fn, reps=10):
    fn(); stream.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream); stream.synchronize()
    return a.elapsed_time(b) / reps


shapes = [(10000, 2_000_000), (2000, 3_000_000)] if "--c3" in sys.argv else [(10000, 2_000_000)]
with torch.cuda.stream(stream):
    for n, L in shapes:
        for ppm in (0, 1, 100, 1000):
            aln = pf.Alignment.synthetic(ctx, n, L, 4)
            if ppm:
                aln.poke_gaps(4, ppm)
            out = torch.zeros(aln.site_len(), dtype=torch.int64, device="cuda")
            cds = torch.zeros(71, dtype=torch.int64, device="cuda")
            res = []
            for flag in ("1", "0"):
                os.environ["PFA_VFLAG"] = flag
                res.append("%s K2 %.3f" % ("sparse" if flag == "1" else "dense-v", timed(lambda: aln.site_stats_device(out.data_ptr()))))
                kn = ctx.last_kernel
                if n <= 12000:
                    res.append("K4 %.3f" % timed(lambda: aln.cds_stats_device(cds.data_ptr())))
            print("%d x %d gaps %4d ppm: %s  [%s]" % (n, L, ppm, " | ".join(res), kn[-40:]), flush=True)
            aln.free()
