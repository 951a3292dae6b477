"""numpy twin of the device generator (csrc/pfa_synth.cuh): the same deterministic synthetic alignment, as text
on the host, and its closed-form column counts.  Integer arithmetic only, so host and device agree bit for bit."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
_PRIMES = [7919, 7927, 7933, 7937, 7949, 7951, 7963, 7993, 8009, 8011, 8017, 8039]


def mix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def multiplier(n):
    for p in _PRIMES:
        if n % p != 0:
            return p
    return 1


def site_params(seed, sites, n, p_seg_ppm=50000, tri_ppm=10000):
    """vectorised pfa_synth_site_params -> dict of arrays (anc, der1, der2, k1, k2, B) over `sites`"""
    sites = np.asarray(sites, dtype=np.uint64)
    u = np.uint64
    with np.errstate(over="ignore"):
        h1 = mix64(u(seed) ^ mix64(sites))
        anc = (h1 & u(3)).astype(np.int64)
        seg = ((h1 >> u(8)) % u(1000000)) < u(p_seg_ppm)
        if n < 2:
            seg[:] = False
        h2, h3 = mix64(h1 + u(1)), mix64(h1 + u(2))
        nb = max(int(n - 1).bit_length(), 1)
        r = np.maximum(u(max(n - 1, 0)) >> (h2 % u(nb)), u(1))
        k1 = np.where(seg, u(1) + (h2 >> u(8)) % r, u(0)).astype(np.int64)
        o1 = (u(1) + (h2 >> u(40)) % u(3)).astype(np.int64)
        der1 = np.where(seg, (anc + o1) & 3, anc)
        tri = seg & (n >= 3) & (k1 + 2 <= n) & (((h3 >> u(8)) % u(1000000)) < u(tri_ppm))
        r2 = np.minimum(max((n - 1) >> 2, 1), np.maximum(n - 1 - k1, 1)).astype(np.uint64)
        k2 = np.where(tri, u(1) + (h3 >> u(32)) % r2, u(0)).astype(np.int64)
        o2 = 1 + ((o1 - 1) + 1 + (h3 & u(1)).astype(np.int64)) % 3
        der2 = np.where(tri, (anc + o2) & 3, anc)
        B = np.where(seg, mix64(h1 + u(3)) % u(max(n, 1)), u(0)).astype(np.int64)
    return {"anc": anc, "der1": der1, "der2": der2, "k1": k1, "k2": k2, "B": B}


def text_matrix(seed, n, L, p_seg_ppm=50000, tri_ppm=10000, col_begin=0, col_end=None):
    """uint8 [n][cols] upper-case text of columns [col_begin, col_end)"""
    col_end = L if col_end is None else col_end
    sites = np.arange(col_begin, col_end, dtype=np.uint64)
    sp = site_params(seed, sites, n, p_seg_ppm, tri_ppm)
    mult = multiplier(n)
    rows = np.arange(n, dtype=np.int64)[:, None]
    pos = (mult * rows + sp["B"][None, :]) % n
    base = np.where(pos < sp["k1"][None, :], sp["der1"][None, :],
                    np.where(pos >= n - sp["k2"][None, :], sp["der2"][None, :], sp["anc"][None, :]))
    return np.frombuffer(b"ACGT", dtype=np.uint8)[base]


def expected_site_stats(seed, n, L, p_seg_ppm=50000, tri_ppm=10000, col_begin=0, col_end=None, chunk=1 << 22):
    """closed-form (S, H, sfs) of the all-rows population: column counts are (n-k1-k2, k1, k2)"""
    col_end = L if col_end is None else col_end
    S = 0
    H = 0
    sfs = np.zeros(n // 2, dtype=np.int64)
    for c0 in range(col_begin, col_end, chunk):
        sp = site_params(seed, np.arange(c0, min(col_end, c0 + chunk), dtype=np.uint64), n, p_seg_ppm, tri_ppm)
        k1, k2 = sp["k1"], sp["k2"]
        seg = k1 > 0
        c0_ = n - k1 - k2
        S += int(seg.sum())
        h = n * n - (c0_ * c0_ + k1 * k1 + k2 * k2)
        H += int(h[seg].sum())
        cnt = np.sort(np.stack([c0_[seg], k1[seg], k2[seg]]), axis=0)   # ascending; second largest = row 1
        second = cnt[1]
        np.add.at(sfs, second - 1, 1)
    return {"n": n, "S": S, "H": H, "sfs": sfs.tolist()}


def poke_gaps(text, seed, gap_ppm, n=None, col_begin=0):
    """numpy twin of pfa_aln_poke_gaps (pfa_synth_gap_bit): returns a copy of the uint8 text matrix [n][cols] with the same
    cells turned into '-'"""
    text = np.array(text, dtype=np.uint8, copy=True)
    n = text.shape[0] if n is None else n
    cols = text.shape[1]
    nwords = (n + 31) // 32
    u = np.uint64
    with np.errstate(over="ignore"):
        sites = np.arange(col_begin, col_begin + cols, dtype=np.uint64)[:, None]
        w = np.arange(nwords, dtype=np.uint64)[None, :]
        key = (u(seed) * u(0x9E3779B97F4A7C15) + u(0x5851F42D4C957F2D)) ^ mix64(sites * u(1048583) + w)
        h = mix64(key)
        hit = ((h >> u(8)) % u(1000000)) < u(32 * gap_ppm)
        bit = ((h >> u(40)) & u(31)).astype(np.int64)
    si, wi = np.nonzero(hit)
    rows = wi * 32 + bit[si, wi]
    ok = rows < n
    text[rows[ok], si[ok]] = ord("-")
    return text
