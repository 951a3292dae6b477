"""Small fixed workloads for ncu captures of K4 (C3 shape) and K3.  usage: python scripts/ncu_targets.py k4|k3"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polyfasta_b200 as pf
from polyfasta_b200._lib import lib, check

which = sys.argv[1]
ctx = pf.Context(0)
if which == "k4":
    aln = pf.Alignment.synthetic(ctx, 2000, 3_000_000, 3)
    aln.set_pops([list(range(0, 1000)), list(range(1000, 2000))])
    out = torch.zeros(2 * 71, dtype=torch.int64, device="cuda")
    for _ in range(3):
        aln.cds_stats_device(out.data_ptr())
    ctx.sync()
    print("k4", out[:6].tolist())
else:
    a = pf.Alignment.synthetic(ctx, 2000, 100_000, 3)
    out = torch.zeros(1, dtype=torch.int64, device="cuda")
    D = torch.zeros((2000, 2000), dtype=torch.int32, device="cuda")
    for _ in range(2):
        check(lib().pfa_pairwise_device(a.handle, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(D.data_ptr())), ctx.handle)
    ctx.sync()
    print("k3", int(out[0]))
