// Deterministic synthetic alignment shared by the device generator (pfa_encode.cu) and its numpy twin
// (polyfasta_b200/synth.py).  Integer arithmetic only, so host and device agree bit for bit.
//
// Per site: ancestral base anc (uniform ACGT); the site segregates with probability p_seg_ppm/1e6 and then
// carries a derived base der1 on exactly k1 rows (k1 log-uniform over the octaves of 1..n-1, i.e. roughly the
// neutral 1/k spectrum); with probability tri_ppm/1e6 a second derived base der2 sits on exactly k2 further
// rows.  Rows are placed by the permutation pos(r) = (mult*r + B) mod n (mult prime, gcd(mult, n) = 1):
// der1 where pos < k1, der2 where pos >= n-k2.  Hence the column counts are known in closed form:
// (n-k1-k2, k1, k2) -- which is what the full-size parity tests check the kernels against.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PFA_HD __host__ __device__ __forceinline__
#else
#define PFA_HD inline
#endif

struct pfa_synth_site {
    uint32_t anc, der1, der2;
    uint64_t k1, k2, B;
};

PFA_HD uint64_t pfa_mix64(uint64_t x) {  // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

PFA_HD pfa_synth_site pfa_synth_site_params(uint64_t seed, uint64_t site, uint64_t n, uint32_t p_seg_ppm, uint32_t tri_ppm) {
    pfa_synth_site s;
    const uint64_t h1 = pfa_mix64(seed ^ pfa_mix64(site));
    s.anc = (uint32_t)(h1 & 3u);
    s.der1 = s.der2 = s.anc;
    s.k1 = s.k2 = 0;
    s.B = 0;
    if (n < 2 || ((h1 >> 8) % 1000000ull) >= p_seg_ppm) return s;
    const uint64_t h2 = pfa_mix64(h1 + 1), h3 = pfa_mix64(h1 + 2);
    uint64_t nb = 0;
    for (uint64_t x = n - 1; x; x >>= 1) ++nb;
    uint64_t r = (n - 1) >> (h2 % nb);
    if (r < 1) r = 1;
    s.k1 = 1 + (h2 >> 8) % r;
    const uint32_t o1 = 1u + (uint32_t)((h2 >> 40) % 3ull);
    s.der1 = (s.anc + o1) & 3u;
    if (n >= 3 && s.k1 + 2 <= n && ((h3 >> 8) % 1000000ull) < tri_ppm) {
        uint64_t r2 = (n - 1) >> 2;
        if (r2 < 1) r2 = 1;
        if (r2 > n - 1 - s.k1) r2 = n - 1 - s.k1;
        s.k2 = 1 + (h3 >> 32) % r2;
        const uint32_t o2 = 1u + ((o1 - 1u) + 1u + (uint32_t)(h3 & 1ull)) % 3u;
        s.der2 = (s.anc + o2) & 3u;
    }
    s.B = pfa_mix64(h1 + 3) % n;
    return s;
}

PFA_HD uint32_t pfa_synth_base(const pfa_synth_site& s, uint64_t row, uint64_t n, uint64_t mult) {
    const uint64_t pos = (mult * row + s.B) % n;
    if (pos < s.k1) return s.der1;
    if (pos >= n - s.k2) return s.der2;
    return s.anc;
}

// sparse gaps on top of the synthetic alignment: the 32-row word w of site `site` loses ONE row to '-' with probability
// 32 * gap_ppm / 1e6 (i.e. gap_ppm gaps per million bases); returns the bit of that row or 32 for "none"
PFA_HD uint32_t pfa_synth_gap_bit(uint64_t seed, uint64_t site, uint64_t w, uint32_t gap_ppm) {
    const uint64_t h = pfa_mix64((seed * 0x9E3779B97F4A7C15ull + 0x5851F42D4C957F2Dull) ^ pfa_mix64(site * 1048583ull + w));
    if ((h >> 8) % 1000000ull >= 32ull * gap_ppm) return 32u;
    return (uint32_t)(h >> 40) & 31u;
}

inline uint64_t pfa_synth_multiplier(uint64_t n) {
    static const uint64_t primes[] = {7919, 7927, 7933, 7937, 7949, 7951, 7963, 7993, 8009, 8011, 8017, 8039};
    for (uint64_t p : primes)
        if (n % p != 0) return p;
    return 1;
}
