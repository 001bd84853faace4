// Probe: sustained cost of one tcgen05.mma.kind::tf32 (M = 128, K = 8) as a function of how it is issued and
// where its operands live.  One CTA, one issuing thread, kIters MMAs accumulating into the same TMEM tile,
// cycles from the first issue to the arrival of the final tcgen05.commit.
//   mode 0: `if (lane == 0)` issue (compiler wraps every MMA in an ELECT / BRA.U.ANY loop), A and B in smem, N = 64
//   mode 1: converged warp + elect.sync, A and B in smem, N = 64
//   mode 2: like 1, N = 128          mode 3: like 1, N = 256
//   mode 4: like 1, N = 64, A in TMEM (only B is read from shared memory)
//   mode 5: like 4, N = 256
//   mode 6: the head's 3xTF32 pattern, lane == 0 issue: (a_hi,b_hi) (a_hi,b_lo) (a_lo,b_hi) per K step, N = 64
//   mode 7: like 6 with elect.sync in a converged warp
// Every mode runs on 1 CTA and on 148 CTAs (one per SM), with zero and with random operand data.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc_n(uint32_t n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok));
    return ok != 0;
}

constexpr int kIters = 1020;

__global__ void __launch_bounds__(128) rate_probe(int mode, int fill, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) unsigned long long mbar, mbar2;
    __shared__ uint32_t tmem_base;
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (49152 + 32768) / 4; i += 128) reinterpret_cast<float*>(base)[i] = fill ? __uint_as_float(0x3f000000u + ((i * 2654435761u) >> 12 & 0x7fe000u)) : 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar2)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    long long t0 = 0, t1 = 0, t_issue = 0;
    if (warp == 1) {
        const uint32_t n = (mode == 2) ? 128u : ((mode == 3 || mode == 5) ? 256u : 64u);
        const uint32_t idesc = idesc_n(n);
        const uint64_t da = make_desc(smem_u32(base)), db = make_desc(smem_u32(base + 49152));
        const uint32_t a_tm = tmem + 256u;       // A operand columns (zeros are fine: contents do not change the timing)
        t0 = clock64();
        if (mode >= 6) {
            const uint64_t a_hi = da, a_lo = make_desc(smem_u32(base + 24576)), b_hi = db, b_lo = make_desc(smem_u32(base + 49152 + 8192));
            if (mode == 6) {
                if (lane == 0) {
                    for (int i = 0; i < kIters; i += 12) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            mma_ss(tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
                            mma_ss(tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
                            mma_ss(tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
                }
            } else {
                // mode 8: + one tcgen05.commit per 12 MMAs; 9: three commits; 10: A start shifted by 5 rows (640 B);
                // 11: alternate between two accumulators; 12: first MMA of each 12 overwrites (accumulate = 0)
                const uint64_t sh = (mode == 10) ? (uint64_t)(640 >> 4) : 0;
                for (int i = 0; i < kIters; i += 12) {
                    if (elect_one()) {
                        const uint32_t d = tmem + ((mode == 11 && (i / 12 & 1)) ? 64u : 0u);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            mma_ss(d, a_hi + sh + 2 * k, b_hi + 2 * k, idesc, (mode == 12 && k == 0) ? 0u : 1u);
                            mma_ss(d, a_hi + sh + 2 * k, b_lo + 2 * k, idesc, 1u);
                            mma_ss(d, a_lo + sh + 2 * k, b_hi + 2 * k, idesc, 1u);
                        }
                        if (mode == 8 || mode == 9)
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar2)) : "memory");
                        if (mode == 9) {
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar2)) : "memory");
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar2)) : "memory");
                        }
                    }
                    __syncwarp();
                }
                if (elect_one())
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
            }
        } else if (mode == 0) {
            if (lane == 0) {
                for (int i = 0; i < kIters; i += 4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) mma_ss(tmem, da + 2 * k, db + 2 * k, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
            }
        } else {
            for (int i = 0; i < kIters; i += 4) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (mode >= 4) mma_ts(tmem, a_tm + 8 * k, db + 2 * k, idesc, 1u);
                        else mma_ss(tmem, da + 2 * k, db + 2 * k, idesc, 1u);
                    }
                }
                __syncwarp();
            }
            if (elect_one())
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        t_issue = clock64();
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
        } while (!ok);
        t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) { out[0] = t_issue - t0; out[1] = t1 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long* d; long long h[2];
    cudaMalloc(&d, 16);
    const int smem = 49152 + 32768 + 1024;
    cudaFuncSetAttribute(rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[] = {"lane==0 issue, SS, N=64", "elect.sync,    SS, N=64", "elect.sync,    SS, N=128", "elect.sync,    SS, N=256",
                           "elect.sync,    TS, N=64 (A in TMEM)", "elect.sync,    TS, N=256 (A in TMEM)"};
    const char* names8[] = {names[0], names[1], names[2], names[3], names[4], names[5], "3xTF32 pattern, lane==0, SS, N=64", "3xTF32 pattern, elect.sync, SS, N=64",
                            "3xTF32 + 1 commit / 12", "3xTF32 + 3 commits / 12", "3xTF32, A shifted 5 rows", "3xTF32, two accumulators", "3xTF32, overwrite every 12"};
    for (int grid = 148; grid <= 148; grid += 147)
        for (int fill = 1; fill < 2; ++fill)
            for (int mode = 6; mode < 13; ++mode) {
                for (int rep = 0; rep < 2; ++rep) rate_probe<<<grid, 128, smem>>>(mode, fill, d);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("grid %3d %s mode %d  %-40s issue %.1f cyc/MMA   complete %.1f cyc/MMA   (%s)\n", grid, fill ? "random" : "zeros ", mode,
                       names8[mode], (double)h[0] / kIters, (double)h[1] / kIters, cudaGetErrorString(e));
            }
    return 0;
}
