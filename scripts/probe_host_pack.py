"""host packer throughput on this box: pfa_host_pack2_rows over a text matrix, per thread count (GB/s of text read)"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polyfasta_b200 import _lib
L = _lib.lib()
n, cols = 10000, 100000
rng = np.random.default_rng(1)
row = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, cols)]
txt = np.empty((n, cols), dtype=np.uint8)
txt[:] = row
dst = np.zeros((n, cols // 4), dtype=np.uint8)
for T in (1, 2, 4, 8, 12, 16, 24, 32):
    if T > (os.cpu_count() or 1):
        break
    L.pfa_host_pack2_rows(txt.ctypes.data, n, cols, cols, dst.ctypes.data, cols // 4, T)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        d = L.pfa_host_pack2_rows(txt.ctypes.data, n, cols, cols, dst.ctypes.data, cols // 4, T)
    dt = time.perf_counter() - t0
    print("threads %2d: %.1f GB/s of text (dirty rows %d)" % (T, reps * txt.nbytes / dt / 1e9, d), flush=True)
t0 = time.perf_counter()
b = txt.copy()
print("numpy copy: %.1f GB/s read" % (txt.nbytes / (time.perf_counter() - t0) / 1e9))
