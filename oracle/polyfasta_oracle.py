"""CPU oracle for the PolyFastA diversity-statistics hot path  --  TEST INFRASTRUCTURE ONLY.

This is an exact-integer restatement of what the reference (`/root/reference/PolyFastA.py`) computes on
the hot path; it is NOT a copy of the reference and it is NOT part of the product.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import it.
The product (`polyfasta_b200/`) never imports anything under `oracle/`.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function here against vectors
produced by running the unmodified reference in the build container (`tests/golden/make_golden.py`):
the 10 shipped loci (C1/C2), 220 random alignments with gaps/N/IUPAC/lower case/stops, all 64*63
ordered codon pairs, 6000 multi-codon sets, the ingest corner cases and the finalisation vectors.

Restatement.  For one population with n rows and a column p, c_a(p) is the number of
rows showing character a (ANY character is an allele: PolyFastA.py:256-258):

    isvar(p) = [#distinct characters > 1]            S = sum_p isvar(p)                (PolyFastA.py:258)
    H        = sum_p (n^2 - sum_a c_a(p)^2)          pi_tot = H / (n (n-1))            (PolyFastA.py:485-492)
    theta_tot = S / a1,  a1 = sum_{i<n} 1/i                                               (PolyFastA.py:494-497)
    D        = (pi_tot - theta_tot) / sqrt(e1 S + e2 S (S-1))                            (PolyFastA.py:522-534)
"""
import math

BASES = "ACGT"
CODONS = [a + b + c for a in BASES for b in BASES for c in BASES]       # index = 16*i1 + 4*i2 + i3
CODON_INDEX = {c: i for i, c in enumerate(CODONS)}
# standard genetic code in TCAG order (what PolyFastA.py:324-329 and :537-556 encode)
_TCAG = "TCAG"
_AA_TCAG = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
AMINO = {a + b + c: _AA_TCAG[16 * i + 4 * j + k]
         for i, a in enumerate(_TCAG) for j, b in enumerate(_TCAG) for k, c in enumerate(_TCAG)}
STOPS = frozenset(c for c in CODONS if AMINO[c] == "*")                  # TGA TAA TAG (PolyFastA.py:285)
SENSE = [c for c in CODONS if c not in STOPS]


def _family_label(codon):
    """the 3-character class of PolyFastA.py:324-329: amino acid, size of its block among the four
    codons sharing the first two bases, and the IUPAC letter of the third bases of that block."""
    aa = AMINO[codon]
    block = [x for x in BASES if AMINO[codon[:2] + x] == aa]
    if len(block) == 4:
        return aa + "4N"
    if len(block) == 3:
        return aa + "3H"
    if len(block) == 2:
        return aa + ("2Y" if codon[2] in "CT" else "2R")
    return aa + "0G"


CLASS = {c: _family_label(c) for c in SENSE}


def syn3(codon):
    """3 * syncodfreq(codon) (PolyFastA.py:536-557) as an integer 0..4: the number of single-base
    neighbours that code the same amino acid; stop codons count 0."""
    if codon in STOPS:
        return 0
    aa = AMINO[codon]
    return sum(1 for i in range(3) for x in BASES
               if x != codon[i] and AMINO[codon[:i] + x + codon[i + 1:]] == aa)


SYN3 = {c: syn3(c) for c in CODONS}

# ------------------------------------------------------------------------------------------------
# ingest  (PolyFastA.py:227-250)
# ------------------------------------------------------------------------------------------------


def parse_fasta(text):
    """text -> (headers, sequences) in first-seen header order, or None when the reference would
    print '# file ... is not FASTA!' (PolyFastA.py:246-248: the LAST header seen is empty / none seen).
    A line starts a record iff its first character is '>' (:232/:241); the header is the rest of the
    line right-stripped; a repeated header restarts that record but keeps its position (:234/:243);
    lines before the first non-empty header are dropped (:235/:244); sequence lines are right-stripped
    and upper-cased (:236/:245)."""
    order, seqs, head = [], {}, ""
    for line in text.splitlines(True):
        if line[0] == ">":
            head = line[1:].rstrip()
            if head not in seqs:
                order.append(head)
            seqs[head] = []
        elif head:
            seqs[head].append(line.rstrip().upper())
    if head == "":
        return None
    return order, ["".join(seqs[h]) for h in order]


def pop_rows(headers, key):
    """row indices whose header contains `key` (PolyFastA.py:125, substring match)"""
    return [i for i, h in enumerate(headers) if key in h]


# ------------------------------------------------------------------------------------------------
# per-site scan  (PolyFastA.py:252-261, 274-282, 485-497)
# ------------------------------------------------------------------------------------------------


def column_counts(rows, p):
    cnt = {}
    for r in rows:
        ch = r[p:p + 1]
        cnt[ch] = cnt.get(ch, 0) + 1
    return cnt


def site_stats(rows, L, want_sfs=True):
    """-> dict(n, S, H, pos, sfs).  sfs follows getsfs (PolyFastA.py:274-282) on the columns where it
    is defined: bin = (second largest count among alleles that are exactly A/C/G/T) - 1, int(n/2)
    bins; a variable column with fewer than two ACGT alleles makes the reference raise IndexError
    and is skipped here."""
    n = len(rows)
    S, H, pos = 0, 0, []
    sfs = [0] * (n // 2)
    for p in range(L):
        cnt = column_counts(rows, p)
        if len(cnt) > 1:
            S += 1
            pos.append(p)
            H += n * n - sum(c * c for c in cnt.values())
            if want_sfs:
                acgt = sorted((c for a, c in cnt.items() if a in ("A", "C", "G", "T")), reverse=True)
                if len(acgt) >= 2:
                    sfs[acgt[1] - 1] += 1
    return {"n": n, "S": S, "H": H, "pos": pos, "sfs": sfs}


def column_H(rows, p):
    n = len(rows)
    cnt = column_counts(rows, p)
    return (n * n - sum(c * c for c in cnt.values())) if len(cnt) > 1 else 0


def pairwise_sum(rows, L):
    """sum_{i<j} d_ij, d_ij = number of columns whose characters differ (dead nucleotide_diversity3,
    PolyFastA.py:468-480).  Identity: equals H/2."""
    tot = 0
    for i in range(len(rows)):
        for j in range(i + 1, len(rows)):
            tot += sum(1 for p in range(L) if rows[i][p:p + 1] != rows[j][p:p + 1])
    return tot


# ------------------------------------------------------------------------------------------------
# codon classifier  (PolyFastA.py:319-434)   labels: 0 none, 1 synonymous, 2 nonsynonymous
# ------------------------------------------------------------------------------------------------
_L_PAIR = frozenset(("L4N", "L2R"))
_R_PAIR = frozenset(("R4N", "R2R"))
_BAND_FROM = frozenset(("L2R", "R2R"))
_BAND_TO = frozenset(("P4N", "L4N", "R4N", "H2Y", "Q2R"))
_NO_3H_2R = frozenset(("AAG", "AGG", "GAG", "TTG", "ATT", "ATC"))
_SYN_PAIR_SETS = [frozenset(x) for x in (("R2R", "H2Y"), ("L2R", "H2Y"), ("D2Y", "R2R"), ("F2Y", "Q2R"),
                                          ("C2Y", "Q2R"), ("E2R", "S2Y"), ("S2Y", "Q2R"))]
_R4N_PARTNERS = frozenset(("ATG", "ATA", "ACA", "ACG", "AAA", "AAG"))


def pair_labels(a, b):
    """labels of the 3 codon positions for two distinct sense codons (PolyFastA.py:347-414)."""
    lab = [0, 0, 0]
    diff = [i for i in range(3) if a[i] != b[i]]
    ka, kb = CLASS[a], CLASS[b]
    kinds = {ka, kb}
    if len(diff) == 1:
        i = diff[0]
        if ka == kb or (i == 0 and (kinds <= _L_PAIR or kinds <= _R_PAIR)):           # :351-355
            lab[i] = 1
        else:
            lab[i] = 2                                                                   # :358
        return lab
    tails = {ka[1:], kb[1:]}
    if diff[-1] == 2:                                                                    # :363
        third_syn = (
            "4N" in tails or tails == {"2Y"} or tails == {"2R"}                          # :365
            or ("I3H" in kinds and ("2Y" in tails or ("2R" in tails and not {a, b} <= _NO_3H_2R)))  # :370-373
            or ("L2R" in kinds and "W0G" in kinds)                                       # :376
            or ("M0G" in kinds and kinds & {"R2R", "K2R", "L2R"})                        # :381
            or any(kinds <= s for s in _SYN_PAIR_SETS)                                   # :386-393
        )
        lab[2] = 1 if third_syn else 2
    if diff[0] == 0:                                                                     # :398
        first_syn = (
            (kinds & _BAND_FROM and kinds & _BAND_TO)                                    # :401
            or ("L4N" in kinds and ({a, b} & {"TCA", "TCG"}))                            # :404-406 ("T0G" never matches)
            or ("R4N" in kinds and ({a, b} & _R4N_PARTNERS))                             # :407-408
        )
        lab[0] = 1 if first_syn else 2
    if 1 in diff[:2]:                                                                    # :413-414
        lab[1] = 2
    return lab


def multi_labels(cods):
    """labels for >= 3 distinct sense codons (PolyFastA.py:415-432)."""
    lab = [0, 0, 0]
    vcp = [i for i in range(3) if len({c[i] for c in cods}) > 1]
    kinds = [CLASS[c] for c in cods]
    if len(set(kinds)) < len(kinds):
        top = max(kinds.count(k) for k in set(kinds))
        if top >= len({c[vcp[-1]] for c in cods}):                                       # :426
            lab[vcp[-1]] = 1
        for i in vcp[:-1]:                                                               # :428-429
            lab[i] = 2
    else:
        for i in vcp:                                                                    # :432
            lab[i] = 2
    return lab


def classify(clean_codons):
    """clean codons of one column (set of [ACGT]{3} strings, stops allowed) -> labels[3]
    (PolyFastA.py:331-334 drops the stops; fewer than two sense codons label nothing)."""
    g = sorted(c for c in clean_codons if c not in STOPS)
    if len(g) <= 1:
        return [0, 0, 0]
    if len(g) == 2:
        return pair_labels(g[0], g[1])
    return multi_labels(g)


# ------------------------------------------------------------------------------------------------
# CDS scan  (PolyFastA.py:150-180, 284-315)
# ------------------------------------------------------------------------------------------------


def _is_clean(s):
    return len(s) == 3 and all(ch in BASES for ch in s)


def cds_stats(rows, L):
    """-> dict(nstops, missing, sum3_by_len{len: sum}, ssites, nsites, S_pos, N_pos, S_s, H_s, S_n, H_n).
    ssites is the reference's sequential float sum (PolyFastA.py:307); the integer form is sum3_by_len:
    ssites == sum_l sum3_by_len[l] / (3 l) exactly in rationals."""
    n = len(rows)
    nstops = missing = 0
    by_len = {}
    ssites = 0.0
    S_pos, N_pos = [], []
    for cp in range(0, L, 3):
        uniq = {r[cp:cp + 3] for r in rows}
        if uniq & STOPS:
            nstops += 1                                                                  # :293
        clean = {u for u in uniq if _is_clean(u)}                                        # :301
        if not clean:
            missing += 3                                                                 # :305
            continue
        tot3 = sum(SYN3[c] for c in clean)
        by_len[len(clean)] = by_len.get(len(clean), 0) + tot3
        ssites += sum(SYN3[c] / 3 for c in clean) / len(clean)                          # :307
        if len(clean) > 1:
            lab = classify(clean)
            for i in range(3):
                if lab[i] == 1:
                    S_pos.append(cp + i)
                elif lab[i] == 2:
                    N_pos.append(cp + i)
    out = {"nstops": nstops, "missing": missing, "sum3_by_len": by_len, "ssites": ssites,
           "nsites": (L - missing) - ssites, "S_pos": S_pos, "N_pos": N_pos}
    out["S_s"] = len(S_pos)
    out["H_s"] = sum(column_H(rows, p) for p in S_pos)
    out["S_n"] = len(N_pos)
    out["H_n"] = sum(column_H(rows, p) for p in N_pos)
    return out


def ssites_from_ints(by_len):
    """the integer accumulators -> synonymous-site count; the reference's float sum agrees to ~1e-13"""
    return sum(v / (3.0 * k) for k, v in sorted(by_len.items()))


# ------------------------------------------------------------------------------------------------
# finalisation  (PolyFastA.py:485-534)
# ------------------------------------------------------------------------------------------------


def finalize(n, S, H, seqlen, jc):
    """(n, S, H, seqlen) -> (S, pi_site, theta_site, D | 'NA'), or (0, 0, 0, 'NA') when S == 0
    (PolyFastA.py:503-504).  pi_tot = H/(n(n-1)); the remaining arithmetic keeps the reference's
    operation order (:525-533, :512-519)."""
    if S == 0:
        return 0, 0, 0, "NA"
    pi_tot = H / (n * (n - 1))
    a1 = sum(1.0 / i for i in range(1, n))
    a2 = sum(1.0 / (i ** 2) for i in range(1, n))
    th_tot = S / a1
    b1 = (n + 1.0) / (3.0 * (n - 1.0))
    b2 = (2.0 * ((n ** 2.0) + n + 3.0)) / (9.0 * n * (n - 1.0))
    c1 = b1 - (1 / a1)
    c2 = b2 - ((n + 2) / (a1 * n)) + (a2 / (a1 ** 2))
    e1 = c1 / a1
    e2 = c2 / ((a1 ** 2) + a2)
    dv = math.sqrt((e1 * S) + (e2 * S * (S - 1)))
    D = (pi_tot - th_tot) / dv if dv != 0 else "NA"                                      # :508-511
    pi_site = pi_tot / seqlen
    if jc:
        x = 1 - (4. / 3.) * pi_site
        if x > 0:                                                                        # :513-516
            pi_site = -0.75 * math.log(x)
    return S, pi_site, th_tot / seqlen, D


# ------------------------------------------------------------------------------------------------
# rows  (PolyFastA.py:147-225)
# ------------------------------------------------------------------------------------------------


def noncds_row(file, seqlen, pop, rows, jc):
    st = site_stats(rows, seqlen, want_sfs=False)
    if st["S"] == 0:
        return f"{file},{seqlen},{pop},{len(rows)},0,0,0,NA"                             # :187/:189
    f = finalize(st["n"], st["S"], st["H"], seqlen, jc)
    return f"{file},{seqlen},{pop},{st['n']},{f[0]},{f[1]},{f[2]},{f[3]}"                # :196/:198


def cds_row(file, seqlen, pop, rows, jc):
    st = site_stats(rows, seqlen, want_sfs=False)
    c = cds_stats(rows, seqlen)
    n = len(rows)
    ss, ns = c["ssites"], c["nsites"]
    if st["S"] == 0:
        return f"{file},{round(ss, 2)},{round(ns, 2)},{pop},{n},0,0,0,NA,0,0,0,NA,0"      # :160/:162
    a = finalize(n, c["S_s"], c["H_s"], ss, jc)
    b = finalize(n, c["S_n"], c["H_n"], ns, jc)
    return (f"{file},{round(ss, 2)},{round(ns, 2)},{pop},{n},{a[0]},{b[0]},{a[1]},{b[1]},"
            f"{a[2]},{b[2]},{a[3]},{b[3]},{c['nstops']}")                                # :178/:180
