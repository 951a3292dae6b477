"""A/B of the register-resident K2 and its TMA variant (per-warp shared-memory rings fed by cp.async.bulk)"""
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
    for n, L in ((10000, 4_000_000), (9000, 4_000_000), (8300, 4_000_000), (20000, 2_000_000), (14000, 3_000_000)):
        aln = pf.Alignment.synthetic(ctx, n, L, 3)
        for hv in (False, True):
            aln.force_validity(hv)
            for pops in (None, [list(range(n // 2)), list(range(n // 2, n))]):
                aln.set_pops(pops)
                out = torch.zeros(aln.site_len(), dtype=torch.int64, device="cuda")
                res = []
                for label, env in (("reg", {"PFA_SITE_TMA": "0"}), ("tma", {"PFA_SITE_TMA": "1"}),
                                   ("tma 2 slots", {"PFA_SITE_TMA": "2"})):
                    os.environ.update(env)
                    ms = min(timed(lambda: aln.site_stats_device(out.data_ptr())) for _ in range(2))
                    ref = out.cpu().clone() if label == "reg" else ref
                    assert torch.equal(out.cpu(), ref), label
                    res.append("%s %.3f ms %4.0f GB/s" % (label, ms, aln.packed_bytes / 3 * (3 if hv else 2) / ms / 1e6))
                print("n=%5d L=%8d planes=%d pops=%d: %s" % (n, L, 3 if hv else 2, 2 if pops else 1, " | ".join(res)), flush=True)
        aln.free()
