"""ctypes front end of the C oracle (oracle/c/polyfasta_oracle.c)  --  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpolyfasta_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "c", "polyfasta_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"] if force else ["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        L.orc_site_stats.argtypes = [p, i64, i64, p, i64, p, p, p, p, p, ctypes.c_int]
        L.orc_cds_stats.argtypes = [p, i64, i64, p, i64, p, p, p, ctypes.c_int]
        L.orc_pairwise_sum.argtypes = [p, i64, i64, p, i64, p, ctypes.c_int]
        L.orc_pairwise_sum.restype = i64
        L.orc_finalize.argtypes = [i64, i64, i64, ctypes.c_double, ctypes.c_int, p, p]
        L.orc_synth_text.argtypes = [p, i64, i64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, i64, i64, ctypes.c_uint64, ctypes.c_int]
        L.orc_synth_text.restype = None
        _lib = L
    return _lib


def text_matrix(seqs):
    """list of equal-length (upper-cased) str/bytes -> uint8 matrix [n][L]"""
    n = len(seqs)
    L = len(seqs[0]) if n else 0
    m = np.zeros((n, max(L, 1)), dtype=np.uint8)
    for i, s in enumerate(seqs):
        b = s.encode("latin-1") if isinstance(s, str) else s
        if L:
            m[i, :L] = np.frombuffer(b, dtype=np.uint8)
    return m[:, :L] if L else m[:, :0]


def _rows(rows, n):
    r = np.arange(n, dtype=np.int32) if rows is None else np.asarray(rows, dtype=np.int32)
    return np.ascontiguousarray(r)


def site_stats(mat, rows=None, want_sfs=True, per_site=False, threads=0):
    mat = np.ascontiguousarray(mat, dtype=np.uint8) if mat.strides[1] != 1 else mat
    L = mat.shape[1]
    ld = mat.strides[0] if mat.shape[0] > 1 else max(L, 1)
    r = _rows(rows, mat.shape[0])
    n = len(r)
    S, H = ctypes.c_int64(), ctypes.c_int64()
    sfs = np.zeros(max(n // 2, 1), dtype=np.int64)
    isvar = np.zeros(max(L, 1), dtype=np.uint8) if per_site else None
    hsite = np.zeros(max(L, 1), dtype=np.int64) if per_site else None
    lib().orc_site_stats(mat.ctypes.data, ld, L, r.ctypes.data, n, ctypes.byref(S), ctypes.byref(H),
                         sfs.ctypes.data if want_sfs else None,
                         isvar.ctypes.data if per_site else None, hsite.ctypes.data if per_site else None, threads)
    out = {"n": n, "S": S.value, "H": H.value, "sfs": sfs[: n // 2].tolist() if want_sfs else None}
    if per_site:
        out["isvar"], out["hsite"] = isvar[:L], hsite[:L]
    return out


def cds_stats(mat, rows=None, want_labels=False, threads=0):
    L = mat.shape[1]
    ld = mat.strides[0] if mat.shape[0] > 1 else max(L, 1)
    r = _rows(rows, mat.shape[0])
    out = np.zeros(71, dtype=np.int64)
    ss = ctypes.c_double()
    labels = np.zeros(max(L, 1), dtype=np.uint8) if want_labels else None
    lib().orc_cds_stats(mat.ctypes.data, ld, L, r.ctypes.data, len(r), out.ctypes.data, ctypes.byref(ss),
                        labels.ctypes.data if want_labels else None, threads)
    res = {"nstops": int(out[0]), "missing": int(out[1]), "S_s": int(out[2]), "H_s": int(out[3]),
           "S_n": int(out[4]), "H_n": int(out[5]),
           "sum3_by_len": {l: int(out[6 + l]) for l in range(65) if out[6 + l]}, "ssites": ss.value,
           "nsites": (L - int(out[1])) - ss.value}
    if want_labels:
        res["labels"] = labels[:L]
    return res


def pairwise_sum(mat, rows=None, want_matrix=False, threads=0):
    L = mat.shape[1]
    ld = mat.strides[0] if mat.shape[0] > 1 else max(L, 1)
    r = _rows(rows, mat.shape[0])
    d = np.zeros((len(r), len(r)), dtype=np.int32) if want_matrix else None
    tot = lib().orc_pairwise_sum(mat.ctypes.data, ld, L, r.ctypes.data, len(r), d.ctypes.data if want_matrix else None, threads)
    return (tot, d) if want_matrix else tot


def finalize(n, S, H, seqlen, jc):
    out = (ctypes.c_double * 3)()
    na = ctypes.c_int()
    if not lib().orc_finalize(n, S, H, float(seqlen), int(bool(jc)), out, ctypes.byref(na)):
        return 0, 0, 0, "NA"
    return S, out[0], out[1], ("NA" if na.value else out[2])


def synth_text(seed, n, L, p_seg_ppm=50000, tri_ppm=10000, col_begin=0, col_end=None, threads=0):
    """C twin of polyfasta_b200.synth.text_matrix (fast enough for the benchmark's CPU sample)"""
    col_end = L if col_end is None else col_end
    out = np.empty((n, col_end - col_begin), dtype=np.uint8)
    mult = next((p for p in (7919, 7927, 7933, 7937, 7949, 7951, 7963, 7993, 8009, 8011, 8017, 8039) if n % p), 1)
    lib().orc_synth_text(out.ctypes.data, out.strides[0], n, seed, p_seg_ppm, tri_ppm, col_begin, col_end, mult, threads)
    return out


def num_threads():
    return int(lib().orc_num_threads())
