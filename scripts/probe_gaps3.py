"""GPU-box probe: the validity-aware kernels on a CLEAN shard with parts switched off (PFA_PROBE_BITS)."""
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
        out = torch.zeros(aln.site_len(), dtype=torch.int64, device="cuda")
        cds = torch.zeros(71, dtype=torch.int64, device="cuda")
        for probe, bits in ((False, 0), (True, 0), (True, 4), (True, 7), (False, 0)):
            os.environ.pop("PFA_PROBE_SPARSE_V", None)
            if probe:
                os.environ["PFA_PROBE_SPARSE_V"] = "1"
            os.environ["PFA_PROBE_BITS"] = str(bits)
            k2 = timed(lambda: aln.site_stats_device(out.data_ptr()))
            kn = ctx.last_kernel
            k4 = timed(lambda: aln.cds_stats_device(cds.data_ptr()))
            print("%d x %d %s bits=%d: K2 %.3f  K4 %.3f  [%s]" % (n, L, "validity-aware" if probe else "pure", bits, k2, k4, kn[:58]), flush=True)
        aln.free()
