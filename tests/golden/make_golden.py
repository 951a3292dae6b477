#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Everything this script writes is committed; tests never import the reference.  The reference
module is imported as-is (it is import-safe, PolyFastA.py:591-592) and only its own functions
are called: readfasta, getvarsites, getsfs, getvarCDSsites, get_syn_nonsyn_cod_sites,
syncodfreq, polymorphism, print_result.  `-p` results are obtained the way PolyFastA.py:133-134
would (sub-dict + print_result) because PolyFastA.py:126 raises TypeError on Python 3.
"""
import contextlib
import importlib.util
import io
import itertools
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile

REF_DIR = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True

spec = importlib.util.spec_from_file_location("ref_polyfasta", os.path.join(REF_DIR, "PolyFastA.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=None, separators=(",", ":"), sort_keys=True)
        f.write("\n")


def H_of(var):
    """sum over var columns of n^2 - sum_a c_a^2, computed from the reference's own `var`."""
    tot = 0
    for col in var:
        n = len(col)
        tot += n * n - sum(col.count(a) ** 2 for a in set(col))
    return tot


def sfs_of(var):
    """reference getsfs (PolyFastA.py:274-282); None where it raises (fewer than 2 ACGT alleles)."""
    try:
        return ref.getsfs(var)
    except IndexError:
        return None


def sfs_skipping(var):
    """getsfs restricted to the columns on which it is defined (>= 2 alleles containing A/C/G/T)."""
    if not var:
        return []
    n = len(var[0])
    sfs = [0] * int(n / 2)
    for col in var:
        try:
            one = ref.getsfs([col])
        except IndexError:
            continue
        for i, x in enumerate(one):
            sfs[i] += x
    return sfs


def captured(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **k)
    return r, buf.getvalue()


def site_record(d, seqlen, jc_both=True):
    pos, var = ref.getvarsites(d, seqlen)
    rec = {"n": len(d), "S": len(var), "H": H_of(var), "pos": pos, "sfs_ref": sfs_of(var) if var else [],
           "sfs": sfs_skipping(var)}
    for jc in (False, True):
        rec["poly_jc%d" % jc] = list(ref.polymorphism(var, seqlen, jc))
    return rec


def cds_record(d, seqlen):
    pos, var = ref.getvarsites(d, seqlen)
    count_syn, s, n, nstops, missing = ref.getvarCDSsites(d, seqlen)
    rec = {"count_syn": count_syn, "S_pos": s, "N_pos": n, "nstops": nstops, "missing": missing}
    nsites = (seqlen - missing) - count_syn
    rec["nsites"] = nsites
    var_s = [var[pos.index(i)] for i in s if i in pos]
    var_n = [var[pos.index(i)] for i in n if i in pos]
    rec["S_s"], rec["H_s"], rec["S_n"], rec["H_n"] = len(var_s), H_of(var_s), len(var_n), H_of(var_n)
    # integer form of count_syn: sum of 3*syncodfreq grouped by number of unique clean codons
    by_len = {}
    import re
    for cp in range(0, seqlen, 3):
        cods = list(set(d[x][cp:cp + 3] for x in d))
        clean = [c for c in cods if re.match("^[AGTC][AGTC][AGTC]$", c)]
        if clean:
            by_len[len(clean)] = by_len.get(len(clean), 0) + sum(round(3 * ref.syncodfreq(c)) for c in clean)
    rec["sum3_by_len"] = {str(k): v for k, v in sorted(by_len.items())}
    for jc in (False, True):
        try:
            rec["poly_s_jc%d" % jc] = list(ref.polymorphism(var_s, count_syn, jc))
            rec["poly_n_jc%d" % jc] = list(ref.polymorphism(var_n, nsites, jc))
        except Exception as e:  # unreachable per SURVEY Q9, recorded if it ever happens
            rec["poly_err_jc%d" % jc] = repr(e)
    return rec


def rows_text(d, seqlen, cds, jc, file, pop):
    """the row print_result would print to the screen for this (sub-)dict"""
    _, out = captured(ref.print_result, d, seqlen, cds, "", "a", file, pop, True, 0, jc)
    return out


# ----------------------------------------------------------------------------------------------
# 1. the shipped example loci (configs C1 / C2): copy the data files, capture CLI output + KATs
# ----------------------------------------------------------------------------------------------
ex_src = os.path.join(REF_DIR, "example_theta_0.01")
ex_dst = os.path.join(HERE, "example_theta_0.01")
os.makedirs(ex_dst, exist_ok=True)
for fn in sorted(os.listdir(ex_src)):
    shutil.copyfile(os.path.join(ex_src, fn), os.path.join(ex_dst, fn))


def run_cli(args, stdin=None, cwd=None):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONHASHSEED="0")
    # bytes, not text=True: universal newlines would turn the "\r" of the progress line into "\n"
    p = subprocess.run([sys.executable, os.path.join(REF_DIR, "PolyFastA.py")] + args,
                       input=None if stdin is None else stdin.encode(), capture_output=True, env=env, cwd=cwd)
    return {"args": args, "rc": p.returncode, "stdout": p.stdout.decode(), "stdin": stdin is not None}


cli = []
# run from tests/golden so that relative paths in "# file ... is not FASTA!" style rows are stable
cli.append(run_cli(["-d", "example_theta_0.01"], cwd=HERE))
cli.append(run_cli(["-d", "example_theta_0.01", "--jc"], cwd=HERE))
cli.append(run_cli(["-d", "example_theta_0.01", "-s"], cwd=HERE))
cli.append(run_cli(["-f", "example_theta_0.01/file1.fa"], cwd=HERE))
cli.append(run_cli(["-f", "example_theta_0.01/file1.fa", "--jc"], cwd=HERE))
cli.append(run_cli(["-f", "example_theta_0.01/file1.fa", "example_theta_0.01/file7.fa", "-s"], cwd=HERE))
cli.append(run_cli(["-d", "example_theta_0.01", "--cds", "--jc"], cwd=HERE))
cli.append(run_cli(["-d", "example_theta_0.01", "--cds"], cwd=HERE))
cli.append(run_cli(["-d", "example_theta_0.01", "-c", "-s"], cwd=HERE))
with open(os.path.join(ex_src, "file1.fa")) as f:
    file1_text = f.read()
cli.append(run_cli(["--pipe"], stdin=file1_text, cwd=HERE))
cli.append(run_cli(["--pipe", "--name", "locusX"], stdin=file1_text, cwd=HERE))
cli.append(run_cli(["--pipe", "--cds", "--jc", "-n", "geneY"], stdin=file1_text, cwd=HERE))
cli.append(run_cli(["-f", "example_theta_0.01/file1.fa", "-d", "example_theta_0.01"], cwd=HERE))
cli.append(run_cli(["-f", "example_theta_0.01/file1.fa", "-p"], cwd=HERE))
# data problems are rows, not exceptions (PolyFastA.py:135-140, :246-248, :160, :105-106): small fixtures written here
edge = os.path.join(HERE, "edge")
os.makedirs(edge, exist_ok=True)
edge_files = {
    "ragged.fa": ">a\nACGTACGT\n>b\nACGTACG\n>c\nACGTACGT\n",
    "notfasta.txt": "this is not\na fasta file\n",
    "novar_cds.fa": ">a\nATGGCTAAATTT\n>b\nATGGCTAAATTT\n>c\nATGGCTAAATTT\n",
    "novar_cds_partial.fa": ">a\nATGGCTAAATT\n>b\nATGGCTAAATT\n",
    "empty_header_last.fa": ">a\nACGT\n>\nTTTT\n",
}
for fn, text in edge_files.items():
    with open(os.path.join(edge, fn), "w", newline="") as f:
        f.write(text)
cli.append(run_cli(["-f", "edge/ragged.fa"], cwd=HERE))
cli.append(run_cli(["-f", "edge/ragged.fa", "--cds", "--jc"], cwd=HERE))
cli.append(run_cli(["-f", "edge/notfasta.txt"], cwd=HERE))
cli.append(run_cli(["-f", "edge/empty_header_last.fa", "-s"], cwd=HERE))
cli.append(run_cli(["-f", "edge/novar_cds.fa", "--cds"], cwd=HERE))
cli.append(run_cli(["-f", "edge/novar_cds.fa", "--cds", "--jc", "-s"], cwd=HERE))
cli.append(run_cli(["-f", "edge/novar_cds_partial.fa", "--cds"], cwd=HERE))
cli.append(run_cli(["-f", "edge/novar_cds.fa"], cwd=HERE))
cli.append(run_cli(["-f", "example_theta_0.01/file1.fa", "--pipe"], stdin=file1_text, cwd=HERE))
cli.append(run_cli(["-f", "edge/ragged.fa", "edge/notfasta.txt", "example_theta_0.01/file2.fa", "edge/novar_cds.fa"], cwd=HERE))
cli.append(run_cli(["-d", "edge", "-s"], cwd=HERE))
# --out behaviour: capture both the screen text and the file that was written
for src_args, extra in ((["-d", "example_theta_0.01"], []), (["-d", "example_theta_0.01"], ["-s"]), (["-d", "example_theta_0.01"], ["--cds"]),
                        (["-d", "edge"], []), (["-d", "edge"], ["--cds"])):
    with tempfile.TemporaryDirectory() as td:
        outp = os.path.join(td, "out.csv")
        with open(outp, "w") as f:
            f.write("PRE-EXISTING LINE\n")
        r = run_cli(src_args + ["--out", outp] + extra, cwd=HERE)
        r["args"] = src_args + ["--out", "@OUT@"] + extra
        r["stdout"] = r["stdout"].replace(outp, "@OUT@")
        with open(outp) as f:
            r["outfile"] = f.read()
        cli.append(r)
dump("cli_examples.json", cli)

kat = {}
for fn in sorted(os.listdir(ex_src)):
    d = ref.readfasta(os.path.join(ex_src, fn), False)
    seqlen = len(next(iter(d.values())))
    entry = {"seqlen": seqlen, "headers": list(d.keys()), "pops": {}}
    for key in ("NA", "indiv1", "indiv2", "indiv", "indiv3"):
        dg = d if key == "NA" else {k: d[k] for k in d if key in k}
        rec = site_record(dg, seqlen)
        rec["cds"] = cds_record(dg, seqlen)
        for cds in (False, True):
            for jc in (False, True):
                rec["row_cds%d_jc%d" % (cds, jc)] = rows_text(dg, seqlen, cds, jc, fn, key)
        entry["pops"][key] = rec
    kat[fn] = entry
dump("kat_examples.json", kat)

# ----------------------------------------------------------------------------------------------
# 2. ingest semantics (PolyFastA.py:227-250): odd FASTA texts -> readfasta result
# ----------------------------------------------------------------------------------------------
ingest_texts = {
    "plain": ">a\nACGT\n>b\nACGA\n",
    "wrapped": ">a desc here\nAC\nGT\n>b\nACG\nA\n",
    "lowercase": ">a\nacgt\n>b\nAcGn\n",
    "crlf": ">a\r\nACGT\r\n>b\r\nACGA\r\n",
    "dup_header": ">a\nAAAA\n>b\nCCCC\n>a\nGG\nGG\n",
    "pre_header_text": "junk line\nmore junk\n>a\nACGT\n>b\nACGT\n",
    "empty_header_mid": ">a\nACGT\n>\nTTTT\n>b\nACGA\n",
    "empty_header_last": ">a\nACGT\n>\nTTTT\n",
    "no_header": "ACGT\nACGT\n",
    "blank_lines": ">a\n\nAC\n\nGT\n\n>b\nACGA\n\n",
    "inner_spaces": ">a\nAC GT\n>b\nAC-GT\n",
    "trailing_ws": ">a  \nACGT  \n>b\t\nACGA\t\n",
    "no_final_newline": ">a\nACGT\n>b\nACGA",
    "ragged": ">a\nACGT\n>b\nACG\n",
    "header_only": ">a\n>b\n",
    "iupac": ">a\nACRYKMSWN-?.*\n>b\nacrykmswn-?.*\n",
    "gt_in_seq_line": ">a\nAC>GT\n>b\nACGGT\n",
    "header_ws_only": "> \nACGT\n>b\nACGA\n",
    "tabs_inside": ">a\nAC\tGT\n>b\nACGGT\n",
    "empty_file": "",
    "leading_ws_header": " >a\nACGT\n>b\nACGA\n",
}
ingest = {}
for name, text in ingest_texts.items():
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False, newline="") as tf:
        tf.write(text)
        path = tf.name
    try:
        r, out = captured(ref.readfasta, path, False)
        rec = {"text": text, "stdout": out.replace(path, "@PATH@")}
        if isinstance(r, dict):
            rec["keys"] = list(r.keys())
            rec["seqs"] = list(r.values())
        else:
            rec["ret"] = r
    except Exception as e:
        rec = {"text": text, "raises": type(e).__name__}
    os.unlink(path)
    ingest[name] = rec
dump("ingest_cases.json", ingest)

# ----------------------------------------------------------------------------------------------
# 3. codon classifier (PolyFastA.py:319-434) exhaustive pairs + random multi sets; syncodfreq table
# ----------------------------------------------------------------------------------------------
bases = "ACGT"
codons = ["".join(p) for p in itertools.product(bases, repeat=3)]
stops = {"TGA", "TAA", "TAG"}
sense = [c for c in codons if c not in stops]
pairs = {}
for a, b in itertools.permutations(codons, 2):
    S, N = ref.get_syn_nonsyn_cod_sites([[a, b], 0])
    pairs[a + b] = [sorted(S), sorted(N)]
rng = random.Random(20261018)
multi = []
for _ in range(6000):
    k = rng.choice([3, 3, 3, 4, 4, 5, 6, 8, 12, 20, 40, 64])
    cs = rng.sample(codons, min(k, 64))
    if rng.random() < 0.5:
        # clustered sets (few variable positions) are what real data produce
        base = rng.choice(codons)
        cs = sorted({base[:i] + x + base[i + 1:] for i in rng.sample(range(3), rng.choice([1, 2])) for x in bases} | {base})
        rng.shuffle(cs)
        cs = cs[: max(3, rng.randint(3, len(cs)))]
    S, N = ref.get_syn_nonsyn_cod_sites([list(cs), 0])
    multi.append([cs, sorted(S), sorted(N)])
dump("codon_classifier.json", {"pairs": pairs, "multi": multi,
                                "syn3": {c: round(3 * ref.syncodfreq(c)) for c in codons}})

# ----------------------------------------------------------------------------------------------
# 4. random alignments with gaps / N / IUPAC / '?' / lower case / stops: full-path vectors
# ----------------------------------------------------------------------------------------------


def random_alignment(rng, n, L, p_seg, p_junk, junk_alphabet, lower):
    anc = [rng.choice(bases) for _ in range(L)]
    rows = [list(anc) for _ in range(n)]
    for p in range(L):
        if rng.random() < p_seg:
            k = rng.randint(1, max(1, n - 1))
            der = rng.choice([b for b in bases if b != anc[p]])
            for r in rng.sample(range(n), k):
                rows[r][p] = der
            if rng.random() < 0.15:
                der2 = rng.choice([b for b in bases if b not in (anc[p], der)])
                for r in rng.sample(range(n), rng.randint(1, max(1, n // 3))):
                    rows[r][p] = der2
    for r in range(n):
        for p in range(L):
            if rng.random() < p_junk:
                rows[r][p] = rng.choice(junk_alphabet)
    # runs of gaps (indels), whole-column gaps and whole-column N
    if p_junk > 0:
        for _ in range(rng.randint(0, 3)):
            p0 = rng.randrange(L)
            ln = rng.randint(1, 7)
            who = rng.sample(range(n), rng.randint(1, n))
            for r in who:
                for p in range(p0, min(L, p0 + ln)):
                    rows[r][p] = "-"
        if rng.random() < 0.5:
            p = rng.randrange(L)
            ch = rng.choice("-N?R")
            for r in range(n):
                rows[r][p] = ch
    seqs = ["".join(r) for r in rows]
    # readfasta right-strips every line (PolyFastA.py:245): a trailing blank would make the file ragged,
    # which main() rejects before the hot path (PolyFastA.py:112,135-140)
    seqs = [s.rstrip() + "-" * (len(s) - len(s.rstrip())) for s in seqs]
    if lower:
        seqs = ["".join(c.lower() if rng.random() < 0.3 else c for c in s) for s in seqs]
    return seqs


cases = []
rng = random.Random(7)
shapes = [(2, 9), (3, 12), (4, 30), (5, 31), (7, 33), (11, 64), (20, 90), (33, 60), (40, 45), (64, 36), (65, 30),
          (130, 24), (1, 12), (2, 1), (6, 2), (9, 100)]
junks = ["-", "-N", "-N?", "-NRYKM", "-N?RYKMSWBDHV.* X"]
for ci in range(220):
    n, L = shapes[ci % len(shapes)]
    if ci >= 160:
        n, L = rng.randint(2, 48), rng.randint(1, 150)
    p_junk = rng.choice([0, 0, 0.01, 0.05, 0.2])
    seqs = random_alignment(rng, n, L, rng.choice([0.02, 0.1, 0.3, 0.8]), p_junk, rng.choice(junks),
                            lower=rng.random() < 0.3)
    heads = [("pop1_" if i % 3 else "pop2_") + "ind%d" % i for i in range(n)]
    if ci % 5 == 0:
        heads = ["s%d" % i for i in range(n)]
    text = "".join(">%s\n%s\n" % (h, s) for h, s in zip(heads, seqs))
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as tf:
        tf.write(text)
        path = tf.name
    d = ref.readfasta(path, False)
    os.unlink(path)
    assert ref.all_same([len(v) for v in d.values()]) and len(next(iter(d.values()))) == L
    rec = {"text": text, "seqlen": L, "pops": {}}
    for key in ("NA", "pop1", "pop2", "ind1", "s1"):
        dg = d if key == "NA" else {k: d[k] for k in d if key in k}
        if not dg:
            rec["pops"][key] = None
            continue
        pr = site_record(dg, L)
        try:
            pr["cds"] = cds_record(dg, L)
            for cds in (False, True):
                for jc in (False, True):
                    pr["row_cds%d_jc%d" % (cds, jc)] = rows_text(dg, L, cds, jc, "case%d.fa" % ci, key)
        except Exception as e:
            pr["cds_raises"] = type(e).__name__
        rec["pops"][key] = pr
    cases.append(rec)
dump("random_cases.json", cases)

# ----------------------------------------------------------------------------------------------
# 5. finalisation vectors (PolyFastA.py:485-534): polymorphism on synthetic var lists
# ----------------------------------------------------------------------------------------------
fin = []
rng = random.Random(99)
for n in [2, 3, 4, 5, 7, 11, 20, 50, 100, 257, 1000, 4000]:
    for S in [1, 2, 5, 40, 300]:
        var = []
        for _ in range(S):
            k = rng.randint(1, n - 1)
            col = ["A"] * (n - k) + ["C"] * k
            if n >= 4 and rng.random() < 0.2:
                col[0] = "G"
            var.append(col)
        for seqlen in (S, 1000, 12345.678, 333.3333333333333):
            for jc in (False, True):
                r = ref.polymorphism(var, seqlen, jc)
                fin.append({"n": n, "S": S, "H": H_of(var), "seqlen": seqlen, "jc": jc, "out": list(r)})
# JC failure branch (x >= 0.75 -> uncorrected), PolyFastA.py:513-516
var = [["A", "C", "G", "T"]] * 10
for seqlen in (1, 2, 9, 10, 11, 12, 13, 14, 20):
    r = ref.polymorphism(var, seqlen, True)
    fin.append({"n": 4, "S": 10, "H": H_of(var), "seqlen": seqlen, "jc": True, "out": list(r)})
dump("finalize_cases.json", fin)
print("golden vectors written to", HERE)
