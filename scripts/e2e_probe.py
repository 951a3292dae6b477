"""GPU-box probe: where does the end-to-end (host text -> result) time go?  Not part of the product."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polyfasta_b200 as pf
from polyfasta_b200 import api

n, cols = 10000, int(os.environ.get("COLS", "1000000"))
ld = (cols + 255) // 256 * 256
ctx = pf.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
with torch.cuda.stream(stream):
    d_text = torch.empty((n, ld), dtype=torch.uint8, device="cuda")
    api.synth_text_device(ctx, d_text.data_ptr(), ld, n, 4, 50000, 10000, 0, cols)
    ctx.sync()
    h_text = torch.empty((n, ld), dtype=torch.uint8, pin_memory=True)
    h_text.copy_(d_text); torch.cuda.synchronize()
    for rep in range(3):
        t = time.perf_counter(); d_text.copy_(h_text, non_blocking=True); torch.cuda.synchronize()
        dt = time.perf_counter() - t
        print("plain contiguous H2D of %.1f GB: %.1f ms = %.1f GB/s" % (n * ld / 1e9, dt * 1e3, n * ld / dt / 1e9))
    # device-resident encode only (K1 from device text)
    for rep in range(2):
        t = time.perf_counter(); a = pf.Alignment.from_device_ptr(ctx, d_text.data_ptr(), n, cols, ld); ctx.sync()
        dt = time.perf_counter() - t
        print("alloc + K1 encode from DEVICE text: %.1f ms (%.1f GB/s of text)" % (dt * 1e3, n * cols / dt / 1e9))
        t = time.perf_counter(); a.free(); print("  free: %.1f ms" % ((time.perf_counter() - t) * 1e3))
    out = torch.zeros(2 + n // 2, dtype=torch.int64, device="cuda")
    for rep in range(3):
        t0 = time.perf_counter()
        a = pf.Alignment.from_host_ptr(ctx, h_text.data_ptr(), n, cols, ld)
        t1 = time.perf_counter()
        a.site_stats_device(out.data_ptr()); stream.synchronize()
        t2 = time.perf_counter()
        a.free()
        t3 = time.perf_counter()
        print("from_host_ptr %.1f ms | K2 %.1f ms | free %.1f ms | total %.1f ms = %.2e bases/s" %
              ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, n * cols / (t3 - t0)))
    t = time.perf_counter(); x = torch.empty(int(3.8e9), dtype=torch.uint8, device="cuda"); torch.cuda.synchronize()
    print("torch alloc 3.8 GB: %.1f ms" % ((time.perf_counter() - t) * 1e3))
