"""polyfasta_b200 -- B200-native implementation of PolyFastA's diversity-statistics hot path.

The functions below mirror the reference's own function interface for this path (same names, argument meaning
and return values as PolyFastA.py) so that parity tests read like calls into the reference; every one of them runs
its arithmetic on the GPU through libpolyfasta_b200.so.  There is no CPU fallback."""
import sys

from . import api
from .api import Alignment, Context, Fasta, NotFasta, default_context  # noqa: F401
from ._lib import PolyFastaError  # noqa: F401

__all__ = ["readfasta", "getvarsites", "getsfs", "getvarCDSsites", "polymorphism", "nucleotide_diversity3",
           "Alignment", "Context", "Fasta", "NotFasta", "default_context", "PolyFastaError"]


def readfasta(file, stdin=False):
    """{header: SEQUENCE} or 1 after printing '# file ... is not FASTA!' (PolyFastA.py:227-250)"""
    try:
        f = Fasta.from_bytes(sys.stdin.buffer.read()) if stdin else Fasta.from_file(file)
    except NotFasta:
        print(f"# file {file} is not FASTA!")
        return 1
    try:
        return f.as_dict()
    finally:
        f.close()


def _aln(d):
    return Alignment.from_strings(default_context(), list(d.values()))


def getvarsites(d, seqlen):
    """(pos, var): positions and full columns of every column with > 1 distinct character (PolyFastA.py:252-261).
    The GPU marks the variable columns; the columns themselves are sliced from `d` on the host."""
    if not d or seqlen == 0:
        return [], []
    a = _aln({k: v[:seqlen] for k, v in d.items()})
    try:
        isvar = a.site_stats(want_isvar=True)[0]["isvar"]
    finally:
        a.free()
    pos = [int(p) for p in isvar.nonzero()[0]]
    seqs = list(d.values())
    return pos, [[s[p:p + 1] for s in seqs] for p in pos]


def _columns_to_rows(var):
    return ["".join(col[r] for col in var) for r in range(len(var[0]))]


def getsfs(var):
    """folded SFS of a list of variable columns (PolyFastA.py:274-282); columns with fewer than two alleles among
    A/C/G/T, on which the reference raises IndexError, are skipped"""
    a = Alignment.from_strings(default_context(), _columns_to_rows(var))
    try:
        return a.site_stats()[0]["sfs"]
    finally:
        a.free()


def polymorphism(var, seqlen, jc):
    """(S, pi_site, theta_site, D | "NA") of a list of variable columns (PolyFastA.py:502-520)"""
    if len(var) == 0:
        return 0, 0, 0, "NA"
    ctx = default_context()
    a = Alignment.from_strings(ctx, _columns_to_rows(var))
    try:
        st = a.site_stats()[0]
    finally:
        a.free()
    return ctx.finalize([(st["n"], st["S"], st["H"], seqlen, jc)])[0]


def getvarCDSsites(d, seqlen):
    """(count_syn, S_positions, N_positions, nstops, missing) (PolyFastA.py:284-315); positions ascending"""
    a = _aln({k: v[:seqlen] for k, v in d.items()})
    try:
        c = a.cds_stats(want_labels=True)[0]
    finally:
        a.free()
    lab = c["labels"]
    return (c["ssites"], [int(p) for p in (lab == 1).nonzero()[0]], [int(p) for p in (lab == 2).nonzero()[0]],
            c["nstops"], c["missing"])


def nucleotide_diversity3(haplo):
    """average number of pairwise differences between rows (the reference's dead PolyFastA.py:468-480)"""
    a = Alignment.from_strings(default_context(), ["".join(h) for h in haplo])
    try:
        tot = a.pairwise()[0]
    finally:
        a.free()
    n = len(haplo)
    return tot / (n * (n - 1) // 2)
