// K3: pairwise mismatch counts d_ij = #columns whose characters differ between rows i and j.
//
// Replaces the reference's (dead) nucleotide_diversity3 (PolyFastA.py:468-480): sum_{i<j} d_ij / C(n,2); the
// identity sum_{i<j} d_ij = H/2 ties it to the site scan.  Integer work on the INT pipe -- XOR / OR / POPC over
// packed words -- not tensor cores.
//
//  1. pfa_rowmajor_kernel transposes the site-major planes into row-major bit-planes over SITES
//     ([plane][row][word], 32 sites per word) with warp ballots.
//  2. pfa_pairwise_kernel: one 64x64 tile of row pairs per CTA, 4x4 pairs per thread; the K dimension (site words)
//     is staged through shared memory in [plane][word][row] order so that a thread fetches its four rows with one
//     128-bit shared load; per word and pair: mismatch = (a0^b0)|(a1^b1)|(av^bv), POPC, add.
//  3. rows that both show the escape class compare equal in the planes; pfa_pairwise_escape_kernel adds the pairs
//     whose escape BYTES differ from the sorted exception list.
//  4. pfa_pairwise_popsum_kernel reduces the upper triangle per population.
#include "pfa_common.cuh"

#define PW_TILE 64
#define PW_KC 16

__global__ void __launch_bounds__(256) pfa_rowmajor_kernel(const uint32_t* __restrict__ b0, const uint32_t* __restrict__ b1,
                                                           const uint32_t* __restrict__ v, int64_t ns, int Wn, int64_t npad,
                                                           int64_t Wl, uint32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t sw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // word over sites
    const int64_t w = blockIdx.y;                                                      // word over rows
    if (sw * 32 >= ns) return;
    const int64_t site = sw * 32 + lane;
    const uint32_t* planes[3] = {b0, b1, v};
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const uint32_t x = site < ns ? __ldg(planes[p] + site * Wn + w) : 0u;
        uint32_t mine = 0;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const uint32_t word = __ballot_sync(0xffffffffu, (x >> r) & 1u);
            if (lane == r) mine = word;
        }
        const int64_t row = w * 32 + lane;
        out[((int64_t)p * npad + row) * Wl + sw] = mine;
    }
}

__global__ void __launch_bounds__(256) pfa_pairwise_kernel(const uint32_t* __restrict__ rm, int64_t npad, int64_t Wl, int64_t n,
                                                           int32_t* __restrict__ D) {
    // upper-triangular tile index -> (bi, bj), bj >= bi
    const int nt = (int)(npad / PW_TILE);
    int bi = 0, rem = blockIdx.x;
    while (rem >= nt - bi) { rem -= nt - bi; ++bi; }
    const int bj = bi + rem;
    __shared__ __align__(16) uint32_t As[3][PW_KC][PW_TILE];
    __shared__ __align__(16) uint32_t Bs[3][PW_KC][PW_TILE];
    const int tj = threadIdx.x & 15, ti = threadIdx.x >> 4;
    int acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0;
    for (int64_t k0 = 0; k0 < Wl; k0 += PW_KC) {
        // stage 64 rows x PW_KC words x 3 planes of both tiles; consecutive threads read consecutive words of a row
        for (int idx = threadIdx.x; idx < 3 * PW_TILE * PW_KC; idx += 256) {
            const int kw = idx % PW_KC, r = (idx / PW_KC) % PW_TILE, p = idx / (PW_KC * PW_TILE);
            const int64_t k = k0 + kw;
            uint32_t a = 0, b = 0;
            if (k < Wl) {
                a = __ldg(rm + ((int64_t)p * npad + (int64_t)bi * PW_TILE + r) * Wl + k);
                b = __ldg(rm + ((int64_t)p * npad + (int64_t)bj * PW_TILE + r) * Wl + k);
            }
            As[p][kw][r] = a;
            Bs[p][kw][r] = b;
        }
        __syncthreads();
#pragma unroll 4
        for (int kw = 0; kw < PW_KC; ++kw) {
            const uint4 a0 = *reinterpret_cast<const uint4*>(&As[0][kw][ti * 4]);
            const uint4 a1 = *reinterpret_cast<const uint4*>(&As[1][kw][ti * 4]);
            const uint4 av = *reinterpret_cast<const uint4*>(&As[2][kw][ti * 4]);
            const uint4 c0 = *reinterpret_cast<const uint4*>(&Bs[0][kw][tj * 4]);
            const uint4 c1 = *reinterpret_cast<const uint4*>(&Bs[1][kw][tj * 4]);
            const uint4 cv = *reinterpret_cast<const uint4*>(&Bs[2][kw][tj * 4]);
            const uint32_t A0[4] = {a0.x, a0.y, a0.z, a0.w}, A1[4] = {a1.x, a1.y, a1.z, a1.w}, AV[4] = {av.x, av.y, av.z, av.w};
            const uint32_t B0[4] = {c0.x, c0.y, c0.z, c0.w}, B1[4] = {c1.x, c1.y, c1.z, c1.w}, BV[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[i][j] += __popc((A0[i] ^ B0[j]) | (A1[i] ^ B1[j]) | (AV[i] ^ BV[j]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t r = (int64_t)bi * PW_TILE + ti * 4 + i, c = (int64_t)bj * PW_TILE + tj * 4 + j;
            if (r < n && c < n) {
                D[r * n + c] = acc[i][j];
                D[c * n + r] = acc[i][j];
            }
        }
}

// pairs of rows that both carry an escape symbol at a site but different bytes: one more mismatch each
__global__ void __launch_bounds__(256) pfa_pairwise_escape_kernel(const unsigned long long* __restrict__ keys, int64_t n_exc,
                                                                  const int64_t* __restrict__ heads, int64_t n_heads, int64_t n,
                                                                  int32_t* __restrict__ D) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t h = wid; h < n_heads; h += nwarps) {
        const int64_t i0 = heads[h], i1 = (h + 1 < n_heads) ? heads[h + 1] : n_exc;
        const int64_t m = i1 - i0;
        for (int64_t p = lane; p < m * m; p += 32) {
            const int64_t x = p / m, y = p % m;
            if (x >= y) continue;
            const unsigned long long kx = keys[i0 + x], ky = keys[i0 + y];
            if (((kx >> 24) & 0xff) == ((ky >> 24) & 0xff)) continue;
            const int64_t rx = (int64_t)(kx & 0xffffff), ry = (int64_t)(ky & 0xffffff);
            atomicAdd(&D[rx * n + ry], 1);
            atomicAdd(&D[ry * n + rx], 1);
        }
    }
}

__global__ void __launch_bounds__(256) pfa_pairwise_popsum_kernel(const int32_t* __restrict__ D, int64_t n, const uint32_t* __restrict__ masks,
                                                                  int Wn, int k, int64_t* __restrict__ out) {
    const int64_t i = blockIdx.x;
    __shared__ unsigned long long red[8];
    for (int q = 0; q < k; ++q) {
        const uint32_t* mq = masks + (int64_t)q * Wn;
        if (!((mq[i >> 5] >> (i & 31)) & 1u)) continue;  // block-uniform
        unsigned long long s = 0;
        for (int64_t j = i + 1 + threadIdx.x; j < n; j += blockDim.x)
            if ((mq[j >> 5] >> (j & 31)) & 1u) s += (unsigned long long)D[i * n + j];
        for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (int w = 0; w < 8; ++w) t += red[w];
            if (t) atomicAdd(reinterpret_cast<unsigned long long*>(out + q), t);
        }
        __syncthreads();
    }
}

int pfa_launch_pairwise(pfa_aln* a, int64_t* d_out, int32_t* d_matrix) {
    pfa_ctx* ctx = a->ctx;
    PFA_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(int64_t) * (size_t)a->k, ctx->stream));
    const int64_t n = a->n;
    if (n == 0) return PFA_OK;
    int32_t* D = d_matrix;
    bool own = false;
    if (!D) {
        PFA_CUDA(ctx, pfa_dmalloc(ctx, &D, sizeof(int32_t) * (size_t)(n * n)));
        own = true;
    }
    const int Wn = a->Wq * 4;
    const int64_t npad = (int64_t)a->Wq * 128;
    int rc = PFA_OK;
    cudaError_t e = cudaMemsetAsync(D, 0, sizeof(int32_t) * (size_t)(n * n), ctx->stream);
    if (e == cudaSuccess && a->ns > 0) {
        if (!a->rowmajor) {
            a->Wl = pfa_round_up((a->ns + 31) / 32, 4);
            e = pfa_dmalloc(ctx, &a->rowmajor, sizeof(uint32_t) * (size_t)(3 * npad * a->Wl));
            if (e == cudaSuccess) e = cudaMemsetAsync(a->rowmajor, 0, sizeof(uint32_t) * (size_t)(3 * npad * a->Wl), ctx->stream);
            if (e == cudaSuccess) {
                const int64_t sw = (a->ns + 31) / 32;
                dim3 grid((unsigned)((sw + 7) / 8), (unsigned)Wn);
                pfa_rowmajor_kernel<<<grid, 256, 0, ctx->stream>>>((const uint32_t*)a->b0, (const uint32_t*)a->b1, (const uint32_t*)a->v,
                                                                   a->ns, Wn, npad, a->Wl, a->rowmajor);
                ctx->launches++;
                e = cudaGetLastError();
            }
        }
        if (e == cudaSuccess) {
            const int64_t nt = npad / PW_TILE;
            pfa_pairwise_kernel<<<(unsigned)(nt * (nt + 1) / 2), 256, 0, ctx->stream>>>(a->rowmajor, npad, a->Wl, n, D);
            ctx->launches++;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && a->n_exc_sites > 0) {
            int64_t eb = std::min<int64_t>((a->n_exc_sites + 7) / 8, (int64_t)ctx->sm_count * 8);
            pfa_pairwise_escape_kernel<<<(unsigned)eb, 256, 0, ctx->stream>>>(a->exc_keys, a->n_exc, a->exc_heads, a->n_exc_sites, n, D);
            ctx->launches++;
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess) {
        pfa_pairwise_popsum_kernel<<<(unsigned)n, 256, 0, ctx->stream>>>(D, n, (const uint32_t*)a->d_masks, Wn, a->k, d_out);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) rc = pfa_fail(ctx, PFA_ERR_CUDA, "pairwise failed: %s", cudaGetErrorString(e));
    if (own) {
        cudaStreamSynchronize(ctx->stream);
        pfa_dfree(ctx, D);
    }
    return rc;
}
