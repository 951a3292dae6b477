"""where does the cold start of the CLI go? (fresh process)"""
import time, sys, os
t0 = time.perf_counter()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
t1 = time.perf_counter()
import polyfasta_b200
from polyfasta_b200 import api, _lib
t2 = time.perf_counter()
L = _lib.lib()
t3 = time.perf_counter()
n = L.pfa_device_count()
t4 = time.perf_counter()
ctx = api.Context(0)
t5 = time.perf_counter()
b = api.Batch(ctx)
t6 = time.perf_counter()
f = polyfasta_b200.Fasta.from_file(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/example_theta_0.01/file1.fa"))
a = polyfasta_b200.Alignment.from_fasta(ctx, f)
s = a.site_stats()
t7 = time.perf_counter()
r = ctx.finalize([(20, s[0]["S"], s[0]["H"], 1000, False)])
t8 = time.perf_counter()
print("numpy %.0f ms | package %.0f | dlopen %.0f | device_count (cuInit) %.0f | ctx_create %.0f | batch_create %.0f | parse+upload+K2 %.0f | finalize %.0f | total %.0f ms"
      % tuple(x * 1e3 for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t8 - t7, t8 - t0)))
