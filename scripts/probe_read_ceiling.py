"""read-only bandwidth of the C4 planes for several (CTAs per SM, loads in flight per thread) of pfa_read_probe_kernel"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polyfasta_b200 as pf
ctx = pf.Context(0)
aln = pf.Alignment.synthetic(ctx, 10000, 10_000_000, 4)
gb = aln.packed_bytes / 3 * 2 / 1e9
for bps in (2, 4, 8, 16):
    for un in (4, 8, 16):
        os.environ["PFA_PROBE"] = "%d,%d" % (bps, un)
        ms = min(aln.read_probe(2, 5) for _ in range(3))
        print("CTAs/SM %2d  loads in flight %2d : %.3f ms  %.0f GB/s" % (bps, un, ms, gb / ms * 1e3), flush=True)
os.environ.pop("PFA_PROBE", None)
for st, kb in ((2, 100), (3, 64), (4, 48), (5, 40), (6, 32), (8, 24), (12, 16), (16, 12), (4, 32), (4, 16)):
    os.environ["PFA_PROBE_TMA"] = "%d,%d" % (st, kb)
    ms = min(aln.read_probe(2, 5) for _ in range(3))
    print("TMA bulk: %2d stages x %3d KB : %.3f ms  %.0f GB/s" % (st, kb, ms, gb / ms * 1e3), flush=True)
