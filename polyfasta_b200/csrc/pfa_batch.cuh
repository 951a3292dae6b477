// Device-side layout of a batch of small loci (pfa_batch.cu), shared with the batched codon kernels (pfa_codon.cu keeps them
// next to the constant-memory codon tables).
#pragma once
#include "pfa_common.cuh"

struct PfaLocusDesc {
    long long text_off;   // byte offset of the locus' text matrix in the blob
    long long plane_off;  // uint4 offset of its first site record in each plane
    long long mask_off;   // uint4 offset of its masks ([k][Wq], then the union [Wq])
    long long site_base;  // global index of its first site (exception keys, group prefix)
    long long tile_base;  // first K1b tile
    long long pop_base;   // first (locus, population) slot
    int n, L, ld, Wq, k, pad;
};

struct PfaPopSlot {
    long long n;        // rows in the population
    long long out_off;  // offset of [S, H, sfs...] in the batch result vector
    long long locus;
    double seqlen;
};

#define PFA_BATCH_CTILE 256  // codon columns per tile of the batched codon scan (one warp per tile, never across two loci)

struct PfaBatchArgs {
    const uint4* b0;
    const uint4* b1;
    const uint4* v;
    const uint4* masks;
    const PfaLocusDesc* desc;
    const PfaPopSlot* pops;
    const long long* site_base;  // nloci + 1
    const int* locus_invalid;
    long long* out;
    long long n_sites;
    int nloci;
    // codon scan (K4b): tiles of PFA_BATCH_CTILE codon columns, per (locus, population) PFA_CDS_LEN accumulators, plain row counts
    const long long* ctile_base;  // nloci + 1
    long long n_ctiles;
    long long* cds_out;           // [pops][PFA_CDS_LEN]
    const long long* popn;        // [pops]
};

__device__ __forceinline__ int pfa_find_locus(const long long* __restrict__ base, int nloci, long long x) {
    // largest i with base[i] <= x ; base has nloci + 1 entries
    int lo = 0, hi = nloci;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (base[mid] <= x) lo = mid;
        else hi = mid;
    }
    return lo;
}


// K4b (pfa_codon.cu): codon scan of every locus of the batch; keys / heads: the batch's sorted exception list (may be empty)
int pfa_launch_batch_cds(pfa_ctx* ctx, const PfaBatchArgs& args, int max_Wq, const unsigned long long* keys, long long n_exc,
                         const long long* heads, long long n_heads);
