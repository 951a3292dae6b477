/* CPU oracle (C) for the PolyFastA diversity-statistics hot path  --  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of what the reference computes on the hot path, for sizes the pure-Python
 * oracle (oracle/polyfasta_oracle.py) cannot finish in seconds, and the `cpu_baseline` / `--impl
 * reference` leg of bench.py.  It is not part of the product: nothing under polyfasta_b200/ links,
 * loads or calls it.  Parity status: PINNED - tests/test_oracle_c.py checks it against the golden
 * vectors generated from the unmodified reference (tests/golden/) and against the Python oracle.
 *
 * Input is the alignment as an upper-cased byte matrix text[row*ld + site] (what readfasta,
 * PolyFastA.py:227-250, leaves in memory) plus the row indices of one population (PolyFastA.py:125,133).
 *
 *   orc_site_stats   getvarsites + the sums inside nucleotide_diversity / wattersons_theta / getsfs
 *                    (PolyFastA.py:252-261, 274-282, 485-497)
 *   orc_cds_stats    getvarCDSsites + get_syn_nonsyn_cod_sites + the var-site matching of print_result
 *                    (PolyFastA.py:165-170, 284-434)
 *   orc_pairwise_sum nucleotide_diversity3 (PolyFastA.py:468-480), brute force
 *   orc_finalize     polymorphism / Dvar / jukes_cantor_correction (PolyFastA.py:499-534)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "codon_tables.h"

#define COLBLK 64

static inline int base_code(uint8_t ch) {
    switch (ch) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
}

/* counts of every distinct byte of column p among the given rows -> S/H/SFS contribution.
 * Any byte is an allele (PolyFastA.py:256-258).  Returns 1 when the column is variable. */
static int column_full(const uint8_t* text, int64_t ld, int64_t p, const int32_t* rows, int64_t n,
                       int64_t* h_out, int64_t* second_acgt) {
    int64_t cnt[256];
    memset(cnt, 0, sizeof cnt);
    for (int64_t r = 0; r < n; ++r) cnt[text[(int64_t)rows[r] * ld + p]]++;
    int distinct = 0;
    int64_t sq = 0;
    for (int b = 0; b < 256; ++b) if (cnt[b]) { distinct++; sq += cnt[b] * cnt[b]; }
    *h_out = distinct > 1 ? n * n - sq : 0;
    /* getsfs (PolyFastA.py:279-281): counts of alleles that are exactly A/C/G/T, second largest */
    int64_t a[4] = { cnt['A'], cnt['C'], cnt['G'], cnt['T'] };
    int64_t m1 = 0, m2 = 0; int present = 0;
    for (int i = 0; i < 4; ++i) if (a[i]) { present++; if (a[i] > m1) { m2 = m1; m1 = a[i]; } else if (a[i] > m2) m2 = a[i]; }
    *second_acgt = present >= 2 ? m2 : 0;
    return distinct > 1;
}

int orc_site_stats(const uint8_t* text, int64_t ld, int64_t L, const int32_t* rows, int64_t n,
                   int64_t* S_out, int64_t* H_out, int64_t* sfs /* n/2 bins or NULL */,
                   uint8_t* isvar /* L or NULL */, int64_t* hsite /* L or NULL */, int nthreads) {
    int64_t S = 0, H = 0;
    int64_t nbins = n / 2;
    if (sfs) memset(sfs, 0, sizeof(int64_t) * (size_t)nbins);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    int64_t nblk = (L + COLBLK - 1) / COLBLK;
#pragma omp parallel reduction(+ : S, H)
    {
        int64_t* lsfs = sfs ? (int64_t*)calloc((size_t)(nbins > 0 ? nbins : 1), sizeof(int64_t)) : NULL;
#pragma omp for schedule(dynamic, 16)
        for (int64_t blk = 0; blk < nblk; ++blk) {
            int64_t p0 = blk * COLBLK, w = L - p0 < COLBLK ? L - p0 : COLBLK;
            int32_t cnt[COLBLK][5];
            memset(cnt, 0, sizeof cnt);
            for (int64_t r = 0; r < n; ++r) {
                const uint8_t* row = text + (int64_t)rows[r] * ld + p0;
                for (int64_t c = 0; c < w; ++c) cnt[c][base_code(row[c])]++;
            }
            for (int64_t c = 0; c < w; ++c) {
                int64_t h, second; int var;
                if (cnt[c][4]) {                      /* something that is not ACGT: count bytes exactly */
                    var = column_full(text, ld, p0 + c, rows, n, &h, &second);
                } else {
                    int present = 0; int64_t sq = 0, m1 = 0, m2 = 0;
                    for (int i = 0; i < 4; ++i) {
                        int64_t v = cnt[c][i];
                        if (v) { present++; sq += v * v; if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) m2 = v; }
                    }
                    var = present > 1; h = var ? n * n - sq : 0; second = present >= 2 ? m2 : 0;
                }
                if (var) { S++; H += h; if (lsfs && second >= 1 && second <= nbins) lsfs[second - 1]++; }
                if (isvar) isvar[p0 + c] = (uint8_t)var;
                if (hsite) hsite[p0 + c] = h;
            }
        }
        if (lsfs) {
#pragma omp critical
            for (int64_t i = 0; i < nbins; ++i) sfs[i] += lsfs[i];
            free(lsfs);
        }
    }
    *S_out = S; *H_out = H;
    return 0;
}

/* labels for >= 3 distinct sense codons (PolyFastA.py:415-432); present = 64-bit set of sense codons */
static void multi_labels(uint64_t present, int lab[3]) {
    int nb[3] = {0, 0, 0}, kcount[23];
    uint8_t seen[3][4];
    memset(seen, 0, sizeof seen); memset(kcount, 0, sizeof kcount);
    int ncod = 0, top = 0, nkinds = 0;
    for (int c = 0; c < 64; ++c) if (present >> c & 1) {
        ncod++;
        int b[3] = { c >> 4, (c >> 2) & 3, c & 3 };
        for (int i = 0; i < 3; ++i) if (!seen[i][b[i]]) { seen[i][b[i]] = 1; nb[i]++; }
        int k = ORC_CLASS[c];
        if (kcount[k]++ == 0) nkinds++;
        if (kcount[k] > top) top = kcount[k];
    }
    int last = -1;
    for (int i = 0; i < 3; ++i) if (nb[i] > 1) last = i;
    lab[0] = lab[1] = lab[2] = 0;
    if (nkinds < ncod) {
        for (int i = 0; i < last; ++i) if (nb[i] > 1) lab[i] = 2;
        if (top >= nb[last]) lab[last] = 1;
    } else {
        for (int i = 0; i < 3; ++i) if (nb[i] > 1) lab[i] = 2;
    }
}

/* out[0]=nstops out[1]=missing out[2]=S_s out[3]=H_s out[4]=S_n out[5]=H_n out[6+l]=sum3_by_len[l], l=0..64 */
int orc_cds_stats(const uint8_t* text, int64_t ld, int64_t L, const int32_t* rows, int64_t n,
                  int64_t out[71], double* ssites_out, uint8_t* labels /* L or NULL */, int nthreads) {
    int64_t ncol = (L + 2) / 3;
    memset(out, 0, sizeof(int64_t) * 71);
    if (labels) memset(labels, 0, (size_t)L);
    /* per codon column: the presence set of clean codons, then everything is a function of that set */
    uint64_t* present = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(ncol > 0 ? ncol : 1));
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t cc = 0; cc < ncol; ++cc) {
        int64_t cp = cc * 3;
        uint64_t m = 0;
        if (cp + 3 <= L)
            for (int64_t r = 0; r < n; ++r) {
                const uint8_t* s = text + (int64_t)rows[r] * ld + cp;
                int a = base_code(s[0]), b = base_code(s[1]), c = base_code(s[2]);
                if ((a | b | c) < 4) m |= 1ull << (16 * a + 4 * b + c);
            }
        present[cc] = m;
    }
    const uint64_t stopmask = (1ull << 48) | (1ull << 50) | (1ull << 56);   /* TAA TAG TGA */
    double ssites = 0.0;
    for (int64_t cc = 0; cc < ncol; ++cc) {
        uint64_t m = present[cc];
        int64_t cp = cc * 3;
        if (m & stopmask) out[0]++;                                            /* :293 */
        int len = __builtin_popcountll(m);
        if (len == 0) { out[1] += 3; continue; }                               /* :305 */
        int tot3 = 0; double fsum = 0.0;
        for (int c = 0; c < 64; ++c) if (m >> c & 1) { tot3 += ORC_SYN3[c]; fsum += ORC_SYN3[c] / 3.0; }
        out[6 + len] += tot3;
        ssites += fsum / len;                                                  /* :307 */
        if (len < 2) continue;
        uint64_t g = m & ~stopmask;                                            /* :331 */
        int ng = __builtin_popcountll(g);
        if (ng < 2) continue;
        int lab[3];
        if (ng == 2) {
            int a = __builtin_ctzll(g), b = 63 - __builtin_clzll(g);
            uint8_t pl = ORC_PAIR[a][b];
            lab[0] = pl & 3; lab[1] = (pl >> 2) & 3; lab[2] = (pl >> 4) & 3;
        } else multi_labels(g, lab);
        for (int i = 0; i < 3; ++i) {
            if (!lab[i]) continue;
            int64_t h, second;
            column_full(text, ld, cp + i, rows, n, &h, &second);               /* :169-170: the full column */
            if (lab[i] == 1) { out[2]++; out[3] += h; } else { out[4]++; out[5] += h; }
            if (labels) labels[cp + i] = (uint8_t)lab[i];
        }
    }
    free(present);
    *ssites_out = ssites;
    return 0;
}

int64_t orc_pairwise_sum(const uint8_t* text, int64_t ld, int64_t L, const int32_t* rows, int64_t n,
                         int32_t* dmat /* n*n or NULL */, int nthreads) {
    int64_t tot = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tot)
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t* a = text + (int64_t)rows[i] * ld;
        for (int64_t j = i + 1; j < n; ++j) {
            const uint8_t* b = text + (int64_t)rows[j] * ld;
            int64_t d = 0;
            for (int64_t p = 0; p < L; ++p) d += a[p] != b[p];
            tot += d;
            if (dmat) { dmat[i * n + j] = (int32_t)d; dmat[j * n + i] = (int32_t)d; }
        }
    }
    return tot;
}

/* Neumaier sum of 1/i^k, i = 1..n-1: CPython >= 3.12 sum() is compensated, so the reference's
 * a1/a2 (PolyFastA.py:525-526) are the correctly rounded sums of the rounded terms. */
static double harmonic(int64_t n, int k) {
    double s = 0.0, c = 0.0;
    for (int64_t i = 1; i < n; ++i) {
        double x = k == 1 ? 1.0 / (double)i : 1.0 / ((double)i * (double)i);
        double t = s + x;
        if (fabs(s) >= fabs(x)) c += (s - t) + x; else c += (x - t) + s;
        s = t;
    }
    return s + c;
}

/* out[0]=pi_site out[1]=theta_site out[2]=D ; *is_na set when Dv == 0 (PolyFastA.py:508-511).
 * returns 0 when S == 0 (the row is the no-variation row), 1 otherwise. */
int orc_finalize(int64_t n, int64_t S, int64_t H, double seqlen, int jc, double out[3], int* is_na) {
    out[0] = out[1] = out[2] = 0.0; *is_na = 1;
    if (S == 0) return 0;
    double N = (double)n, ss = (double)S;
    double pi_tot = (double)H / (N * (N - 1.0));
    double a1 = harmonic(n, 1), a2 = harmonic(n, 2);
    double th = ss / a1;
    double b1 = (N + 1.0) / (3.0 * (N - 1.0));
    double b2 = (2.0 * ((N * N) + N + 3.0)) / (9.0 * N * (N - 1.0));
    double c1 = b1 - (1.0 / a1);
    double c2 = b2 - ((N + 2.0) / (a1 * N)) + (a2 / (a1 * a1));
    double e1 = c1 / a1;
    double e2 = c2 / ((a1 * a1) + a2);
    double dv = sqrt((e1 * ss) + (e2 * ss * (ss - 1.0)));
    if (dv != 0.0) { out[2] = (pi_tot - th) / dv; *is_na = 0; }
    double pi_site = pi_tot / seqlen;
    if (jc) { double x = 1.0 - (4.0 / 3.0) * pi_site; if (x > 0.0) pi_site = -0.75 * log(x); }
    out[0] = pi_site; out[1] = th / seqlen;
    return 1;
}

/* C twin of the synthetic alignment used by the benchmark (same integer recipe as polyfasta_b200/synth.py):
 * writes columns [col_begin, col_end) as upper-case text, out[row*ld + (col - col_begin)]. */
static uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

void orc_synth_text(uint8_t* out, int64_t ld, int64_t n, uint64_t seed, uint32_t p_seg_ppm, uint32_t tri_ppm,
                    int64_t col_begin, int64_t col_end, uint64_t mult, int nthreads) {
    static const char B[4] = {'A', 'C', 'G', 'T'};
    uint64_t nb = 0;
    for (uint64_t x = (uint64_t)n - 1; x; x >>= 1) ++nb;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t col = col_begin; col < col_end; ++col) {
        const uint64_t h1 = mix64(seed ^ mix64((uint64_t)col));
        uint32_t anc = (uint32_t)(h1 & 3u), der1 = anc, der2 = anc;
        uint64_t k1 = 0, k2 = 0, Bo = 0;
        if (n >= 2 && ((h1 >> 8) % 1000000ull) < p_seg_ppm) {
            const uint64_t h2 = mix64(h1 + 1), h3 = mix64(h1 + 2);
            uint64_t r = ((uint64_t)n - 1) >> (h2 % nb);
            if (r < 1) r = 1;
            k1 = 1 + (h2 >> 8) % r;
            const uint32_t o1 = 1u + (uint32_t)((h2 >> 40) % 3ull);
            der1 = (anc + o1) & 3u;
            if (n >= 3 && k1 + 2 <= (uint64_t)n && ((h3 >> 8) % 1000000ull) < tri_ppm) {
                uint64_t r2 = ((uint64_t)n - 1) >> 2;
                if (r2 < 1) r2 = 1;
                if (r2 > (uint64_t)n - 1 - k1) r2 = (uint64_t)n - 1 - k1;
                k2 = 1 + (h3 >> 32) % r2;
                der2 = (anc + 1u + ((o1 - 1u) + 1u + (uint32_t)(h3 & 1ull)) % 3u) & 3u;
            }
            Bo = mix64(h1 + 3) % (uint64_t)n;
        }
        uint8_t* dst = out + (col - col_begin);
        if (!k1) {
            for (int64_t r = 0; r < n; ++r) dst[r * ld] = (uint8_t)B[anc];
        } else {
            for (int64_t r = 0; r < n; ++r) {
                const uint64_t pos = (mult * (uint64_t)r + Bo) % (uint64_t)n;
                dst[r * ld] = (uint8_t)B[pos < k1 ? der1 : (pos >= (uint64_t)n - k2 ? der2 : anc)];
            }
        }
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
