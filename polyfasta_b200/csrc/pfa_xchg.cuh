// In-kernel sum of the per-shard integer vectors over NVLink / NVSwitch peer memory (the one
// exchange step of the column-sharded path).  No collective library on the data path: the LAST block of a scan
// kernel pushes the shard's vector into every rank's accumulation buffer with system-scope 64-bit reductions
// (red.global.add.u64 over peer mappings), raises a flag on every rank, waits until all ranks have raised its own,
// and copies the finished sum to the caller's buffer.  One launch = scan + all-reduce.
//
// Symmetric buffer of one rank (cudaMalloc, exported with cudaIpcGetMemHandle, see pfa_xchg.cu):
//   flag[2] (uint32, 128-byte slot each) | acc[2][cap] (int64)
// Exchange number e uses slot e & 1.  A slot is zeroed by its owner during exchange e-1 BEFORE the owner raises
// its flags of e-1; a peer can only add into it after it has seen those flags, so two slots suffice.  Flags count
// up for ever: after exchange e the flag of slot e&1 reads (e/2 + 1) * world.
#pragma once
#include <stdint.h>

#define PFA_XCHG_MAX_RANKS 16
#define PFA_XCHG_FLAG_BYTES 256  // two 128-byte flag slots in front of the accumulators
#define PFA_XCHG_TIMEOUT_NS 4000000000ull  // default; PFA_XCHG_TIMEOUT_MS / pfa_xchg_set_timeout_ms override it per exchange

struct PfaXchgDev {
    int world;  // 0: no exchange (single-GPU launches)
    int rank;
    unsigned int epoch;  // exchanges completed before this launch
    int len;             // int64 words exchanged by this launch (<= cap)
    int zero_len;        // longest vector exchanged so far (>= len): that much of the idle slot is cleared
    int64_t cap;
    unsigned long long timeout_ns;  // how long the last block waits for the other ranks' flags before it gives up
    unsigned long long* partial;  // this rank's vector: the blocks add into it; zero at entry, zeroed again at exit
    unsigned int* ticket;         // block counter, zero at entry and at exit
    unsigned int* status;         // set to 1 when the wait timed out (a rank is missing)
    unsigned long long* stamps;   // [8] %globaltimer at the stages of the last exchange (pfa_xchg_stamps)
    int64_t* out;                 // the reduced vector goes here (device memory of this rank)
    char* base[PFA_XCHG_MAX_RANKS];  // symmetric buffer of every rank as mapped into this process
};

__device__ __forceinline__ unsigned int* pfa_xchg_flag(char* base, int slot) {
    return reinterpret_cast<unsigned int*>(base + 128 * slot);
}
__device__ __forceinline__ unsigned long long* pfa_xchg_acc(char* base, int slot, int64_t cap) {
    return reinterpret_cast<unsigned long long*>(base + PFA_XCHG_FLAG_BYTES) + (int64_t)slot * cap;
}
__device__ __forceinline__ void pfa_red_add_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void pfa_signal_sys(unsigned int* p) {
    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned int pfa_ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pfa_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Called by EVERY thread of EVERY block at the very end of a scan kernel, after the block has added its share into
// x.partial.  Returns in all blocks but the last one to arrive; that one runs the exchange.
__device__ __forceinline__ void pfa_xchg_epilogue(const PfaXchgDev& x) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(x.ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) x.stamps[0] = pfa_globaltimer();
    const int slot = (int)(x.epoch & 1u);
    const int tid = threadIdx.x, nt = blockDim.x;
    // the other slot is idle until every rank has seen this exchange's flags: clear it for exchange epoch+1
    unsigned long long* idle = pfa_xchg_acc(x.base[x.rank], slot ^ 1, x.cap);
    for (int i = tid; i < x.zero_len; i += nt) __stcg(idle + i, 0ull);
    // push this shard's vector into every rank (self included); four independent loads in flight per thread
    for (int i0 = tid; i0 < x.len; i0 += 4 * nt) {
        unsigned long long v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = i0 + u * nt < x.len ? __ldcg(x.partial + i0 + u * nt) : 0ull;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (v[u]) {
                for (int p = 0; p < x.world; ++p) pfa_red_add_sys(pfa_xchg_acc(x.base[p], slot, x.cap) + i0 + u * nt, v[u]);
                __stcg(x.partial + i0 + u * nt, 0ull);
            }
    }
    if (tid == 0) x.stamps[1] = pfa_globaltimer();
    __threadfence_system();
    __syncthreads();
    if (tid == 0) x.stamps[2] = pfa_globaltimer();
    if (tid < x.world) pfa_signal_sys(pfa_xchg_flag(x.base[tid], slot));
    if (tid == 0) {
        const unsigned int want = (x.epoch / 2u + 1u) * (unsigned int)x.world;
        const unsigned int* mine = pfa_xchg_flag(x.base[x.rank], slot);
        const unsigned long long t0 = pfa_globaltimer();
        while ((int)(pfa_ld_acquire_sys(mine) - want) < 0) {
            if (pfa_globaltimer() - t0 > x.timeout_ns) {
                *x.status = 1u;
                break;
            }
        }
        *x.ticket = 0u;
        x.stamps[3] = pfa_globaltimer();
    }
    __syncthreads();
    const unsigned long long* sum = pfa_xchg_acc(x.base[x.rank], slot, x.cap);
    for (int i0 = tid; i0 < x.len; i0 += 4 * nt) {
        unsigned long long v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = i0 + u * nt < x.len ? __ldcg(sum + i0 + u * nt) : 0ull;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + u * nt < x.len) x.out[i0 + u * nt] = (int64_t)v[u];
    }
    __syncthreads();
    if (tid == 0) x.stamps[4] = pfa_globaltimer();
}
