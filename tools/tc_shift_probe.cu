// Probe: can the A operand of tcgen05.mma (K-major, SWIZZLE_128B, rows 128 B apart) start at a row that is
// not a multiple of 8, i.e. at a start address that is only 128-byte aligned?  If it can, a 3x3 convolution tap
// over a row-padded pixel list is the same smem tile read through a descriptor shifted by (4 dy + dx) rows,
// and the tensor-core head needs no per-tap restaging.  Two descriptor variants per shift: base_offset = 0 and
// base_offset = (start >> 7) & 7.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

constexpr int kRows = 160, kPad = 16;

__global__ void __launch_bounds__(128) tc_probe(const float* A, const float* B, float* D, int shift, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* sA = reinterpret_cast<float*>(smem);                      // kRows x 32 floats, swizzled by absolute row
    float* sB = reinterpret_cast<float*>(smem + kRows * 128);
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kRows * 32; i += 128) {
        const int r = i >> 5, k = i & 31, chunk = k >> 2;
        sA[r * 32 + ((chunk ^ (r & 7)) << 2) + (k & 3)] = A[i];
    }
    for (int i = tid; i < 64 * 32; i += 128) {
        const int r = i >> 5, k = i & 31, chunk = k >> 2;
        sB[r * 32 + ((chunk ^ (r & 7)) << 2) + (k & 3)] = B[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(sA) + (uint32_t)((kPad + shift) * 128);
        const uint64_t da = make_desc(a0, mode ? (a0 >> 7) & 7 : 0), db = make_desc(smem_u32(sB), 0);
        for (int k = 0; k < 4; ++k) {
            const uint32_t acc = k > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
        } while (!ok);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
    static float hA[kRows * 32], hB[64 * 32], hD[128 * 64];
    srand(1);
    for (int i = 0; i < kRows * 32; ++i) hA[i] = (float)((rand() % 17) - 8) * 0.25f;
    for (int i = 0; i < 64 * 32; ++i) hB[i] = (float)((rand() % 13) - 6) * 0.5f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    const int smem = kRows * 128 + 64 * 128 + 1024;
    cudaFuncSetAttribute(tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode = 0; mode < 2; ++mode)
        for (int shift = -5; shift <= 8; ++shift) {
            cudaMemset(dD, 0xff, sizeof hD);
            tc_probe<<<1, 128, smem>>>(dA, dB, dD, shift, mode);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
            int bad = 0, bad_rows = 0;
            for (int m = 0; m < 128; ++m) {
                int rb = 0;
                for (int n = 0; n < 64; ++n) {
                    float s = 0;
                    for (int k = 0; k < 32; ++k) s += hA[(m + kPad + shift) * 32 + k] * hB[n * 32 + k];
                    if (!(fabs((double)hD[m * 64 + n] - s) <= 1e-6)) { ++bad; rb = 1; }
                }
                bad_rows += rb;
            }
            printf("mode %d (base_offset %s) shift %+d: err=%s mismatches=%d bad_rows=%d\n", mode, mode ? "(addr>>7)&7" : "0",
                   shift, cudaGetErrorString(e), bad, bad_rows);
        }
    return 0;
}
