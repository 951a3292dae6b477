"""small end-to-end run for compute-sanitizer (memcheck): every kernel family once on tiny inputs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import polyfasta_b200 as pf
ctx = pf.Context(0)
rng = np.random.default_rng(1)
for n, L in [(20, 100), (129, 301), (2100, 96), (4200, 33)]:
    text = np.frombuffer(b"ACGTacgt-N?RY", dtype=np.uint8)[rng.integers(0, 13, (n, L))]
    text = np.where(rng.random((n, L)) < 0.9, np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)][None, :], text).astype(np.uint8)
    a = pf.Alignment.from_rows(ctx, np.ascontiguousarray(text))
    a.set_pops([list(range(n)), list(range(0, n, 2))])
    s = a.site_stats(want_isvar=True); c = a.cds_stats(want_labels=True)
    if n <= 200:
        p = a.pairwise()
    a.free()
    b = pf.api.Batch(ctx); b.add_rows(np.ascontiguousarray(text), [list(range(n)), list(range(1, n, 2))]); b.add_rows(np.ascontiguousarray(text[:7, :40])); b.run(True); b.close()
a = pf.Alignment.synthetic(ctx, 300, 2000, 1); a.site_stats(); a.cds_stats(); a.free()
print("sanitize run ok", ctx.finalize([(20, 5, 100, 50, True)]))
