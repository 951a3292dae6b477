"""K2 time over alignment shapes: register-resident kernel vs the TMA variant (slots per warp, passes per slot)"""
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
CFG = [("reg", {"PFA_SITE_TMA": "0"}), ("default", {})] + [("tma m x%d" % f, {"PFA_SITE_TMA": "1", "PFA_SITE_TMA_M": "x%d" % f, "PFA_SITE_TMA_MIN_LPS": "1"}) for f in (1, 2, 3)]
with torch.cuda.stream(stream):
    for n, L in ((20, 100_000_000), (100, 40_000_000), (300, 10_000_000), (500, 10_000_000), (1000, 6_000_000), (2000, 3_000_000), (3000, 2_000_000), (4000, 2_000_000), (6000, 2_000_000), (10000, 2_000_000), (16000, 1_000_000)):
        aln = pf.Alignment.synthetic(ctx, n, L, 3)
        out = torch.zeros(aln.site_len(), dtype=torch.int64, device="cuda")
        wq = (n + 127) // 128
        lps = 1
        while lps < 32 and (wq + lps - 1) // lps > 5: lps *= 2
        m0 = max(1, 2560 // ((32 // lps) * wq * 16))   # passes per slot for ~2.5 KB per plane
        res, ref = [], None
        for label, env in CFG:
            for k in ("PFA_SITE_TMA", "PFA_SITE_TMA_M", "PFA_SITE_TMA_MIN_LPS"): os.environ.pop(k, None)
            env = dict(env)
            if env.get("PFA_SITE_TMA_M", "").startswith("x"): env["PFA_SITE_TMA_M"] = str(int(env["PFA_SITE_TMA_M"][1:]) * m0)
            os.environ.update(env)
            ms = min(timed(lambda: aln.site_stats_device(out.data_ptr())) for _ in range(2))
            cur = out.cpu().clone()
            ref = cur if ref is None else ref
            assert torch.equal(cur, ref), label
            res.append("%s %.3f" % (label, ms))
        print("n=%5d L=%9d (%.2f GB, m0=%d): %s" % (n, L, aln.packed_bytes / 3 * 2 / 1e9, m0, " | ".join(res)), flush=True)
        aln.free()
