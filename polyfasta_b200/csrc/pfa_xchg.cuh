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
    int wide;            // all blocks of the launch take part in the exchange (they are co-resident): see pfa_xchg_epilogue_wide
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
__device__ __forceinline__ unsigned int pfa_ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pfa_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// The exchange by ALL blocks of the launch (x.wide: the launch has at most one block per SM slot, so they are resident together;
// ranks on distinct GPUs).  With one block doing everything, pushing 5,002 words to 8 ranks took 8.8 us and the whole epilogue
// ~20 us of a 480 us step; here every block pushes, and later copies, 1/gridDim of the vector:
//   barrier A (all blocks have added their share into x.partial)  ->  each block: clear its slice of the idle slot, push its
//   slice to every rank, fence  ->  barrier B: the last block there raises the flags on all ranks  ->  every block waits for
//   this rank's flag and copies its slice of the total to the caller's buffer  ->  the last block out resets the counters.
// x.ticket[0], [2], [3] count the blocks at the three stages; spins give up after x.timeout_ns (status word).
__device__ __forceinline__ void pfa_xchg_epilogue_wide(const PfaXchgDev& x) {
    __shared__ int s_flag;
    const int tid = threadIdx.x, nt = blockDim.x;
    const unsigned G = gridDim.x;
    unsigned int* cnt_a = x.ticket;
    unsigned int* cnt_b = x.ticket + 2;
    unsigned int* cnt_c = x.ticket + 3;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        atomicAdd(cnt_a, 1u);
        const unsigned long long t0 = pfa_globaltimer();
        while (pfa_ld_acquire_gpu(cnt_a) < G) {
            if (pfa_globaltimer() - t0 > x.timeout_ns) {
                *x.status = 1u;
                break;
            }
        }
        if (blockIdx.x == 0) x.stamps[0] = pfa_globaltimer();
    }
    __syncthreads();
    const int slot = (int)(x.epoch & 1u);
    const int per = (x.len + (int)G - 1) / (int)G, zper = (x.zero_len + (int)G - 1) / (int)G;
    const int lo = min(x.len, (int)blockIdx.x * per), hi = min(x.len, lo + per);
    unsigned long long* idle = pfa_xchg_acc(x.base[x.rank], slot ^ 1, x.cap);
    for (int i = (int)blockIdx.x * zper + tid; i < min(x.zero_len, ((int)blockIdx.x + 1) * zper); i += nt) __stcg(idle + i, 0ull);
    for (int i = lo + tid; i < hi; i += nt) {
        const unsigned long long v = __ldcg(x.partial + i);
        if (v) {
            for (int p = 0; p < x.world; ++p) pfa_red_add_sys(pfa_xchg_acc(x.base[p], slot, x.cap) + i, v);
            __stcg(x.partial + i, 0ull);
        }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        if (blockIdx.x == 0) x.stamps[1] = x.stamps[2] = pfa_globaltimer();
        s_flag = atomicAdd(cnt_b, 1u) == G - 1 ? 1 : 0;
        __threadfence();
    }
    __syncthreads();
    if (s_flag && tid < x.world) pfa_signal_sys(pfa_xchg_flag(x.base[tid], slot));  // every block's pushes are out and fenced
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
        if (blockIdx.x == 0) x.stamps[3] = pfa_globaltimer();
    }
    __syncthreads();
    const unsigned long long* sum = pfa_xchg_acc(x.base[x.rank], slot, x.cap);
    for (int i = lo + tid; i < hi; i += nt) x.out[i] = (int64_t)__ldcg(sum + i);
    __syncthreads();
    if (tid == 0) {
        if (blockIdx.x == 0) x.stamps[4] = pfa_globaltimer();
        __threadfence();
        if (atomicAdd(cnt_c, 1u) == G - 1) {  // everybody is past the barriers: zero the counters for the next launch
            *cnt_a = 0u;
            *cnt_b = 0u;
            *cnt_c = 0u;
        }
    }
}

// Called by EVERY thread of EVERY block at the very end of a scan kernel, after the block has added its share into
// x.partial.  Returns in all blocks but the last one to arrive; that one runs the exchange.
__device__ __forceinline__ void pfa_xchg_epilogue(const PfaXchgDev& x) {
    if (x.wide) {
        pfa_xchg_epilogue_wide(x);
        return;
    }
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
