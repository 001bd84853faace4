// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (kind::tf32, M=128, N=64, K=32 as 4 k-steps): A[m][k] is
// written with tcgen05.st to lane m, column a_col + k; B is a K-major SWIZZLE_128B tile in shared memory.
// (derived from tc_probe.cu) one tcgen05.mma (kind::tf32, M=128, N=64, K=32 as 4 k-steps) with K-major SWIZZLE_128B operands
// written by ordinary stores, accumulator in TMEM, read back with tcgen05.ld.  Validates the shared-memory
// descriptor / instruction descriptor / TMEM addressing used by the tensor-core head.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // K-major, SWIZZLE_128B: start>>4 | LBO(=1)<<16 | SBO(=1024B>>4)<<32 | version 1 <<46 | layout 2 << 61
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(128) tc_probe(const float* A, const float* B, float* D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* sA = reinterpret_cast<float*>(smem);
    float* sB = reinterpret_cast<float*>(smem + 128 * 128);
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 64 * 32; i += 128) {
        const int r = i >> 5, k = i & 31, chunk = k >> 2;
        sB[r * 32 + ((chunk ^ (r & 7)) << 2) + (k & 3)] = B[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    {   // A -> TMEM columns 64..95: thread (warp w, lane l) owns row 32 w + l
        uint32_t r[32];
        for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(A[(warp * 32 + lane) * 32 + k]);
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 64u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                     "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                     ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                       "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                       "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                       "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(sA)), db = make_desc(smem_u32(sB));
        for (int k = 0; k < 4; ++k) {
            const uint32_t acc = k > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem), "r"(tmem + 64u + 8u * k), "l"(db + 2 * k), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {   // wait for the MMAs
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
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main() {
    float hA[128 * 32], hB[64 * 32], hD[128 * 64], ref[128 * 64];
    srand(1);
    for (int i = 0; i < 128 * 32; ++i) hA[i] = (float)((rand() % 17) - 8) * 0.25f;      // exactly representable in tf32
    for (int i = 0; i < 64 * 32; ++i) hB[i] = (float)((rand() % 13) - 6) * 0.5f;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
        float s = 0; for (int k = 0; k < 32; ++k) s += hA[m * 32 + k] * hB[n * 32 + k]; ref[m * 64 + n] = s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, sizeof hD);
    const int smem = 128 * 128 + 64 * 128 + 1024;
    cudaFuncSetAttribute(tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tc_probe<<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int i = 0; i < 128 * 64; ++i) { double d = fabs((double)hD[i] - ref[i]); if (!(d <= 1e-6)) { if (bad < 5) printf("mismatch at m=%d n=%d got %f want %f\n", i / 64, i % 64, hD[i], ref[i]); ++bad; } if (d > maxerr) maxerr = d; }
    printf("tcgen05 tf32 A-in-TMEM probe: err=%s mismatches=%d maxerr=%g\n", cudaGetErrorString(e), bad, maxerr);
    return 0;
}
