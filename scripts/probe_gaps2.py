"""GPU-box probe: where does the validity-aware site scan lose time?  Clean shard through the pure kernel, the same shard through
the validity-aware kernel with no flag set (PFA_PROBE_SPARSE_V), then 1 / 10 / 100 gaps per 10^6 bases."""
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


with torch.cuda.stream(stream):
    for n, L in [(10000, 2_000_000), (2000, 3_000_000)]:
        for ppm, probe in ((0, False), (0, True), (1, False), (10, False), (100, False)):
            aln = pf.Alignment.synthetic(ctx, n, L, 4)
            if ppm:
                aln.poke_gaps(4, ppm)
            os.environ.pop("PFA_PROBE_SPARSE_V", None)
            if probe:
                os.environ["PFA_PROBE_SPARSE_V"] = "1"
            out = torch.zeros(aln.site_len(), dtype=torch.int64, device="cuda")
            cds = torch.zeros(71, dtype=torch.int64, device="cuda")
            k2 = timed(lambda: aln.site_stats_device(out.data_ptr()))
            kn = ctx.last_kernel
            k4 = timed(lambda: aln.cds_stats_device(cds.data_ptr()))
            print("%d x %d gaps %4d ppm%s: K2 %.3f  K4 %.3f  [%s]" % (n, L, ppm, " (validity-aware kernel forced)" if probe else "", k2, k4, kn[:60]), flush=True)
            aln.free()
