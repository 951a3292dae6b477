// K5: finalisation in fp64 on the device.
//
// Replaces polymorphism / nucleotide_diversity / wattersons_theta / Dvar / jukes_cantor_correction
// (PolyFastA.py:485-534).  Inputs are the exact integers (n, S, H); pi_tot = H / (n (n-1)) is the closed form
// of N/(N-1) * sum_cols (1 - sum_a (c_a/n)^2) (:486-491).  All arithmetic is IEEE fp64 with explicit
// round-to-nearest intrinsics, i.e. no FMA contraction, in the reference's operation order (:525-533).
// a1 = sum 1/i and a2 = sum 1/i^2 are accumulated as double-double partial sums and rounded once: CPython
// (>= 3.12) sums floats with Neumaier compensation, so the reference's a1/a2 are the rounded exact sums too.
#include "pfa_common.cuh"

struct dd {
    double hi, lo;
};

__device__ __forceinline__ dd dd_add_d(dd a, double x) {  // a + x, error-free transformation (Knuth two-sum)
    const double s = __dadd_rn(a.hi, x);
    const double bb = __dsub_rn(s, a.hi);
    const double e = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(x, bb));
    dd r;
    r.hi = s;
    r.lo = __dadd_rn(a.lo, e);
    return r;
}

__device__ __forceinline__ dd dd_add(dd a, dd b) {
    dd r = dd_add_d(a, b.hi);
    r.lo = __dadd_rn(r.lo, b.lo);
    // renormalise
    const double s = __dadd_rn(r.hi, r.lo);
    r.lo = __dsub_rn(r.lo, __dsub_rn(s, r.hi));
    r.hi = s;
    return r;
}

__global__ void __launch_bounds__(256) pfa_finalize_kernel(const pfa_final_in* __restrict__ in, pfa_final_out* __restrict__ out, int count) {
    __shared__ dd s1[256], s2[256];
    const int e = blockIdx.x;
    if (e >= count) return;
    const pfa_final_in p = in[e];
    const int t = threadIdx.x;
    dd a1 = {0.0, 0.0}, a2 = {0.0, 0.0};
    for (long long i = 1 + t; i < p.n; i += 256) {
        const double di = (double)i;
        a1 = dd_add_d(a1, __ddiv_rn(1.0, di));
        a2 = dd_add_d(a2, __ddiv_rn(1.0, __dmul_rn(di, di)));
    }
    s1[t] = a1;
    s2[t] = a2;
    __syncthreads();
    for (int off = 128; off; off >>= 1) {
        if (t < off) {
            s1[t] = dd_add(s1[t], s1[t + off]);
            s2[t] = dd_add(s2[t], s2[t + off]);
        }
        __syncthreads();
    }
    if (t != 0) return;
    pfa_final_out r;
    r.pi_site = 0.0; r.theta_site = 0.0; r.D = 0.0; r.D_is_NA = 1; r.no_var = 0;
    if (p.S == 0) {  // polymorphism returns (0, 0, 0, "NA") (PolyFastA.py:503-504)
        r.no_var = 1;
        out[e] = r;
        return;
    }
    const double A1 = __dadd_rn(s1[0].hi, s1[0].lo), A2 = __dadd_rn(s2[0].hi, s2[0].lo);
    const double N = (double)p.n, ss = (double)p.S;
    const double pi_tot = __ddiv_rn((double)p.H, __dmul_rn(N, __dsub_rn(N, 1.0)));
    const double th_tot = __ddiv_rn(ss, A1);                                                     // :497
    const double b1 = __ddiv_rn(__dadd_rn(N, 1.0), __dmul_rn(3.0, __dsub_rn(N, 1.0)));           // :527
    const double b2 = __ddiv_rn(__dmul_rn(2.0, __dadd_rn(__dadd_rn(__dmul_rn(N, N), N), 3.0)),
                                __dmul_rn(__dmul_rn(9.0, N), __dsub_rn(N, 1.0)));                // :528
    const double c1 = __dsub_rn(b1, __ddiv_rn(1.0, A1));                                         // :529
    const double A1sq = __dmul_rn(A1, A1);
    const double c2 = __dadd_rn(__dsub_rn(b2, __ddiv_rn(__dadd_rn(N, 2.0), __dmul_rn(A1, N))), __ddiv_rn(A2, A1sq));  // :530
    const double e1 = __ddiv_rn(c1, A1);                                                         // :531
    const double e2 = __ddiv_rn(c2, __dadd_rn(A1sq, A2));                                        // :532
    const double rad = __dadd_rn(__dmul_rn(e1, ss), __dmul_rn(__dmul_rn(e2, ss), __dsub_rn(ss, 1.0)));
    const double dv = __dsqrt_rn(rad);                                                           // :533
    if (dv > 0.0) {  // Dv == 0 -> ZeroDivisionError -> "NA" (:508-511)
        r.D = __ddiv_rn(__dsub_rn(pi_tot, th_tot), dv);
        r.D_is_NA = 0;
    }
    double pi_site = __ddiv_rn(pi_tot, p.seqlen);
    if (p.jc) {  // -0.75*log(1 - 4/3 x); a failing log keeps x (:499-500, :513-516)
        const double x = __dsub_rn(1.0, __dmul_rn(4.0 / 3.0, pi_site));
        if (x > 0.0) pi_site = __dmul_rn(-0.75, log(x));
    }
    r.pi_site = pi_site;
    r.theta_site = __ddiv_rn(th_tot, p.seqlen);                                                  // :519
    out[e] = r;
}

// ssites = sum_l sum3_by_len[l] / (3 l): the exact-integer form of the running float sum of PolyFastA.py:307
__global__ void pfa_ssites_kernel(const int64_t* __restrict__ cds, double* __restrict__ ssites, int count) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const int64_t* row = cds + (int64_t)e * PFA_CDS_LEN;
    double s = 0.0;
    for (int l = 1; l <= 64; ++l)
        if (row[PFA_CDS_SUM3 + l]) s = __dadd_rn(s, __ddiv_rn((double)row[PFA_CDS_SUM3 + l], __dmul_rn(3.0, (double)l)));
    ssites[e] = s;
}

// device-to-device finalisation of `count` entries (used by the batched path: inputs never leave the GPU)
int pfa_launch_finalize(pfa_ctx* ctx, const pfa_final_in* d_in, pfa_final_out* d_out, int count) {
    if (count <= 0) return PFA_OK;
    pfa_finalize_kernel<<<count, 256, 0, ctx->stream>>>(d_in, d_out, count);
    PFA_LAUNCH_CHECK(ctx);
    return PFA_OK;
}

static int ensure_scratch(pfa_ctx* ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return PFA_OK;
    if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    ctx->h_scratch = ctx->d_scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = pfa_round_up((int64_t)bytes, 1 << 16);
    PFA_CUDA(ctx, cudaHostAlloc(&ctx->h_scratch, want, cudaHostAllocDefault));
    PFA_CUDA(ctx, cudaMalloc(&ctx->d_scratch, want));
    ctx->scratch_bytes = want;
    return PFA_OK;
}

extern "C" int pfa_finalize(pfa_ctx* ctx, const pfa_final_in* in, pfa_final_out* out, int count) {
    if (!ctx) return PFA_ERR_ARG;
    if (count <= 0) return PFA_OK;
    if (!in || !out) return pfa_fail(ctx, PFA_ERR_ARG, "pfa_finalize: null buffer");
    PFA_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t in_b = sizeof(pfa_final_in) * (size_t)count, out_b = sizeof(pfa_final_out) * (size_t)count;
    const size_t off = (size_t)pfa_round_up((int64_t)in_b, 256);
    int rc = ensure_scratch(ctx, off + out_b);
    if (rc) return rc;
    memcpy(ctx->h_scratch, in, in_b);
    char* d = static_cast<char*>(ctx->d_scratch);
    char* h = static_cast<char*>(ctx->h_scratch);
    PFA_CUDA(ctx, cudaMemcpyAsync(d, h, in_b, cudaMemcpyHostToDevice, ctx->stream));
    pfa_finalize_kernel<<<count, 256, 0, ctx->stream>>>(reinterpret_cast<const pfa_final_in*>(d),
                                                        reinterpret_cast<pfa_final_out*>(d + off), count);
    PFA_LAUNCH_CHECK(ctx);
    PFA_CUDA(ctx, cudaMemcpyAsync(h + off, d + off, out_b, cudaMemcpyDeviceToHost, ctx->stream));
    PFA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out, h + off, out_b);
    return PFA_OK;
}

extern "C" int pfa_cds_ssites(pfa_ctx* ctx, const int64_t* cds_out, double* ssites, int count) {
    if (!ctx) return PFA_ERR_ARG;
    if (count <= 0) return PFA_OK;
    if (!cds_out || !ssites) return pfa_fail(ctx, PFA_ERR_ARG, "pfa_cds_ssites: null buffer");
    PFA_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t in_b = sizeof(int64_t) * PFA_CDS_LEN * (size_t)count, out_b = sizeof(double) * (size_t)count;
    const size_t off = (size_t)pfa_round_up((int64_t)in_b, 256);
    int rc = ensure_scratch(ctx, off + out_b);
    if (rc) return rc;
    char* d = static_cast<char*>(ctx->d_scratch);
    char* h = static_cast<char*>(ctx->h_scratch);
    memcpy(h, cds_out, in_b);
    PFA_CUDA(ctx, cudaMemcpyAsync(d, h, in_b, cudaMemcpyHostToDevice, ctx->stream));
    pfa_ssites_kernel<<<(count + 63) / 64, 64, 0, ctx->stream>>>(reinterpret_cast<const int64_t*>(d),
                                                                 reinterpret_cast<double*>(d + off), count);
    PFA_LAUNCH_CHECK(ctx);
    PFA_CUDA(ctx, cudaMemcpyAsync(h + off, d + off, out_b, cudaMemcpyDeviceToHost, ctx->stream));
    PFA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(ssites, h + off, out_b);
    return PFA_OK;
}
