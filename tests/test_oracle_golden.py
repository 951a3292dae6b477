"""Pins the CPU oracle (oracle/polyfasta_oracle.py) to vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import math

import pytest

from oracle import polyfasta_oracle as orc
from conftest import load_golden


def close(a, b, rel=1e-12):
    if isinstance(a, str) or isinstance(b, str):
        return a == b
    return math.isclose(a, b, rel_tol=rel, abs_tol=0.0) or a == b


def check_poly(got, want, n, S, H):
    assert got[0] == want[0]
    assert close(got[1], want[1]) and close(got[2], want[2])
    if isinstance(want[3], str) or isinstance(got[3], str):
        assert got[3] == want[3]
    else:
        # D = (pi - theta)/Dv cancels: bound the error by 1e-12 of the terms, not of the difference
        pi_tot = H / (n * (n - 1))
        scale = max(abs(want[3]), abs(want[3]) * (pi_tot / max(abs(pi_tot - S / sum(1.0 / i for i in range(1, n))), 1e-300)))
        assert abs(got[3] - want[3]) <= 1e-12 * max(scale, 1e-300)


def test_tables_against_reference():
    g = load_golden("codon_classifier.json")
    assert {c: orc.SYN3[c] for c in orc.CODONS} == g["syn3"]
    assert len(orc.SENSE) == 61 and len(set(orc.CLASS.values())) == 23


def test_pair_classifier_exhaustive():
    g = load_golden("codon_classifier.json")["pairs"]
    assert len(g) == 64 * 63
    for key, (S, N) in g.items():
        a, b = key[:3], key[3:]
        lab = orc.classify({a, b})
        assert [i for i in range(3) if lab[i] == 1] == S, key
        assert [i for i in range(3) if lab[i] == 2] == N, key


def test_multi_classifier():
    for cods, S, N in load_golden("codon_classifier.json")["multi"]:
        lab = orc.classify(set(cods))
        assert [i for i in range(3) if lab[i] == 1] == S, cods
        assert [i for i in range(3) if lab[i] == 2] == N, cods


def test_ingest_cases():
    for name, rec in load_golden("ingest_cases.json").items():
        got = orc.parse_fasta(rec["text"])
        if "ret" in rec:
            assert got is None, name
        else:
            assert got is not None, name
            assert got[0] == rec["keys"] and got[1] == rec["seqs"], name


def _check_pop(rows, L, rec, file, key):
    st = orc.site_stats(rows, L)
    assert (st["n"], st["S"], st["H"], st["pos"]) == (rec["n"], rec["S"], rec["H"], rec["pos"])
    assert st["sfs"] == (rec["sfs"] if rec["S"] else [0] * (rec["n"] // 2))   # getsfs needs >= 1 var column
    if rec["sfs_ref"] is not None and rec["S"]:
        assert st["sfs"] == rec["sfs_ref"]
    for jc in (0, 1):
        check_poly(orc.finalize(st["n"], st["S"], st["H"], L, bool(jc)), rec["poly_jc%d" % jc], st["n"], st["S"], st["H"])
    c = rec.get("cds")
    if c is None:
        return
    cs = orc.cds_stats(rows, L)
    assert cs["nstops"] == c["nstops"] and cs["missing"] == c["missing"]
    # the reference appends the 3rd codon position before the 1st (PolyFastA.py:363-411): compare as sets
    assert cs["S_pos"] == sorted(c["S_pos"]) and cs["N_pos"] == sorted(c["N_pos"])
    assert len(set(c["S_pos"])) == len(c["S_pos"]) and len(set(c["N_pos"])) == len(c["N_pos"])
    assert (cs["S_s"], cs["H_s"], cs["S_n"], cs["H_n"]) == (c["S_s"], c["H_s"], c["S_n"], c["H_n"])
    assert {str(k): v for k, v in sorted(cs["sum3_by_len"].items())} == c["sum3_by_len"]
    assert close(cs["ssites"], c["count_syn"], 1e-13) and close(cs["nsites"], c["nsites"], 1e-13)
    assert close(orc.ssites_from_ints(cs["sum3_by_len"]), c["count_syn"], 1e-12)
    for cds in (0, 1):
        for jc in (0, 1):
            want = rec["row_cds%d_jc%d" % (cds, jc)].rstrip("\n")
            got = (orc.cds_row if cds else orc.noncds_row)(file, L, key, rows, bool(jc))
            _rows_equal(got, want)


def _rows_equal(got, want):
    g, w = got.split(","), want.split(",")
    assert len(g) == len(w), (got, want)
    for x, y in zip(g, w):
        if x == y:
            continue
        fx, fy = float(x), float(y)
        assert math.isclose(fx, fy, rel_tol=2e-11), (got, want)   # D columns: see check_poly for the tight bound


def test_example_loci():
    kat = load_golden("kat_examples.json")
    import os
    from conftest import GOLDEN
    for fn, entry in kat.items():
        with open(os.path.join(GOLDEN, "example_theta_0.01", fn)) as f:
            heads, seqs = orc.parse_fasta(f.read())
        assert heads == entry["headers"]
        L = entry["seqlen"]
        for key, rec in entry["pops"].items():
            rows = seqs if key == "NA" else [seqs[i] for i in orc.pop_rows(heads, key)]
            _check_pop(rows, L, rec, fn, key)


def test_random_cases():
    for ci, case in enumerate(load_golden("random_cases.json")):
        heads, seqs = orc.parse_fasta(case["text"])
        L = case["seqlen"]
        for key, rec in case["pops"].items():
            rows = seqs if key == "NA" else [seqs[i] for i in orc.pop_rows(heads, key)]
            if rec is None:
                assert rows == []
                continue
            _check_pop(rows, L, rec, "case%d.fa" % ci, key)
            if len(rows) <= 12 and L <= 40:
                assert 2 * orc.pairwise_sum(rows, L) == rec["H"]


def test_finalize_cases():
    for rec in load_golden("finalize_cases.json"):
        got = orc.finalize(rec["n"], rec["S"], rec["H"], rec["seqlen"], rec["jc"])
        check_poly(got, rec["out"], rec["n"], rec["S"], rec["H"])
