"""Pins the C oracle (oracle/c/polyfasta_oracle.c) to the reference's golden vectors and to the
Python oracle.  CPU only."""
import math
import os
import random

import numpy as np

from conftest import GOLDEN, load_golden
from oracle import c_oracle as co
from oracle import polyfasta_oracle as orc
from test_oracle_golden import check_poly, close


def _check(seqs, heads, L, pops, want_pairwise=False):
    mat = co.text_matrix(seqs)
    for key, rec in pops.items():
        rows = None if key == "NA" else orc.pop_rows(heads, key)
        if rec is None:
            continue
        st = co.site_stats(mat, rows, per_site=True)
        assert (st["n"], st["S"], st["H"]) == (rec["n"], rec["S"], rec["H"])
        assert np.flatnonzero(st["isvar"]).tolist() == rec["pos"]
        assert st["sfs"] == (rec["sfs"] if rec["S"] else [0] * (rec["n"] // 2))
        for jc in (0, 1):
            check_poly(co.finalize(st["n"], st["S"], st["H"], L, jc), rec["poly_jc%d" % jc], st["n"], st["S"], st["H"])
        c = rec.get("cds")
        if c is None:
            continue
        cs = co.cds_stats(mat, rows, want_labels=True)
        for k in ("nstops", "missing", "S_s", "H_s", "S_n", "H_n"):
            assert cs[k] == c[k], k
        assert np.flatnonzero(cs["labels"] == 1).tolist() == sorted(c["S_pos"])
        assert np.flatnonzero(cs["labels"] == 2).tolist() == sorted(c["N_pos"])
        assert {str(k): v for k, v in cs["sum3_by_len"].items()} == {k: v for k, v in c["sum3_by_len"].items() if v}
        assert close(cs["ssites"], c["count_syn"]) and close(cs["nsites"], c["nsites"])
        if want_pairwise and st["n"] <= 40:
            assert 2 * co.pairwise_sum(mat, rows) == rec["H"]


def test_example_loci_c():
    for fn, entry in load_golden("kat_examples.json").items():
        with open(os.path.join(GOLDEN, "example_theta_0.01", fn)) as f:
            heads, seqs = orc.parse_fasta(f.read())
        _check(seqs, heads, entry["seqlen"], entry["pops"], want_pairwise=True)


def test_random_cases_c():
    for case in load_golden("random_cases.json"):
        heads, seqs = orc.parse_fasta(case["text"])
        _check(seqs, heads, case["seqlen"], case["pops"], want_pairwise=True)


def test_finalize_cases_c():
    for rec in load_golden("finalize_cases.json"):
        check_poly(co.finalize(rec["n"], rec["S"], rec["H"], rec["seqlen"], rec["jc"]), rec["out"], rec["n"], rec["S"], rec["H"])


def test_c_vs_python_oracle_medium():
    rng = random.Random(5)
    n, L = 37, 3000
    anc = [rng.choice("ACGT") for _ in range(L)]
    rows = []
    for r in range(n):
        s = list(anc)
        for p in range(L):
            u = rng.random()
            if u < 0.03:
                s[p] = rng.choice("ACGT")
            elif u < 0.035:
                s[p] = rng.choice("-N?RY")
        rows.append("".join(s))
    a = orc.site_stats(rows, L)
    b = co.site_stats(co.text_matrix(rows), None, threads=3)
    assert (a["S"], a["H"], a["sfs"]) == (b["S"], b["H"], b["sfs"])
    ca, cb = orc.cds_stats(rows, L), co.cds_stats(co.text_matrix(rows), None, threads=2)
    for k in ("nstops", "missing", "S_s", "H_s", "S_n", "H_n", "sum3_by_len"):
        assert ca[k] == cb[k], k
    assert math.isclose(ca["ssites"], cb["ssites"], rel_tol=1e-13)
