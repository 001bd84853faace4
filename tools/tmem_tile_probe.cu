// Probe: tensor memory as a second read-only scratchpad for the RoIAlign tile kernel.
//
// roi_tile_tma_kernel is bound by shared-memory wavefronts (LSU data pipe 71 % busy, profiles/r2_summary.md).
// Question: can part of the tile be served from TENSOR memory instead?  tcgen05.cp.32x128b.warpx4 copies 32 rows
// x 16 bytes from a SWIZZLE_128B shared-memory tile into 4 columns of all four lane quarters (so any warp can read
// it), and tcgen05.ld.32x32b.x4 hands a lane 4 consecutive columns = the same 4 pixels of its channel an LDS.128
// gives it.  This probe checks (1) the layout, (2) the read rates of LDS.128, of tcgen05.ld and of both at once,
// (3) what the copies cost.  Measured on the B200: layout exact (0 mismatches); LDS.128 128 B/clk/SM, tcgen05.ld.x4
// up to 222, both at once 120 + 134 - the read paths add up; but a tcgen05.cp costs ~75 cycles per instruction
// (32x128b.warpx4, 512 bytes; 128x256b / 128x128b: ~130 cycles) and LDS.128 falls to ~50 B/clk/SM while copies run:
// 64 copies per tile (8 rows) take 5 200 cycles.  The kernel variant (-DRR_T2_TMEM=1) is slower: 0.525 against 0.463 ms.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_tile_probe tools/tmem_tile_probe.cu && /tmp/tmem_tile_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kThreads = 768;
constexpr int kWarps = kThreads / 32;
constexpr int kTW = 32, kTH = 24, kTC = 32;
constexpr int kTileBytes = kTW * kTH * kTC * 4;
constexpr int kTmRows = 8;                       // tile rows kept in tensor memory (8 x 32 pixels = 256 columns)

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw128_desc(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void ldtm4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void ldtm_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: layout check.  mode 1: LDS.128 only (n_lds warps).  mode 2: tcgen05.ld only (n_tm warps).  mode 3: both.
__global__ void __launch_bounds__(kThreads, 1)
probe(int mode, int iters, int n_lds, int n_tm, int n_cp, int* mismatches, float* sink, unsigned long long* cycles, float* dump) {
    extern __shared__ unsigned char raw[];
    __shared__ uint32_t s_tmem;
    __shared__ unsigned long long s_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* tile = raw + ((1024u - (saddr(raw) & 1023u)) & 1023u);
    for (int i = tid; i < kTH * kTC * kTW; i += kThreads) {
        const int x = i & 31, c = (i >> 5) & 31, y = i >> 10;
        const float v = (float)(y * 10000 + c * 100 + x);
        *reinterpret_cast<float*>(tile + (y * kTC + c) * 128 + (((x >> 2) ^ (c & 7)) << 4) + ((x & 3) << 2)) = v;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(&s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(&s_bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    long long t_cp0 = 0, t_cp1 = 0;
    if (tid == 0) {
        t_cp0 = clock64();
        for (int y = 0; y < kTmRows; ++y)
            for (int q = 0; q < 8 && y * 8 + q < n_cp; ++q) {
                const uint64_t d = sw128_desc(saddr(tile) + y * 4096 + q * 16);
                asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem + y * 32 + 4 * q), "l"(d) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(saddr(&s_bar)) : "memory");
    }
    {
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(saddr(&s_bar)), "r"(0) : "memory");
        } while (!ok);
    }
    if (tid == 0) { t_cp1 = clock64(); cycles[gridDim.x * kWarps + blockIdx.x] = (unsigned long long)(t_cp1 - t_cp0); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tq = tmem + ((uint32_t)(32 * (warp & 3)) << 16);

    if (mode == 4) {
    } else if (mode == 6 || mode == 7) {                     // cost of the wider copy shapes (no replication): n_cp copies
        __syncthreads();
        if (tid == 0) {
            const long long t0 = clock64();
            for (int i = 0; i < n_cp; ++i) {
                const uint64_t d = sw128_desc(saddr(tile) + (i % 6) * 4 * 4096 + ((i / 6) & 3) * 32);
                if (mode == 6)
                    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem + (i % 32) * 8), "l"(d) : "memory");
                else
                    asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(tmem + (i % 64) * 4), "l"(d) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(saddr(&s_bar)) : "memory");
            uint32_t ok;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(saddr(&s_bar)), "r"(1) : "memory");
            } while (!ok);
            cycles[gridDim.x * kWarps + blockIdx.x] = (unsigned long long)(clock64() - t0);
        }
    } else if (mode == 0) {
        int bad = 0;
        for (int y = 0; y < kTmRows; ++y)
            for (int q = 0; q < 8; ++q) {
                uint32_t r[4];
                ldtm4(tq + y * 32 + 4 * q, r);
                ldtm_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float want = (float)(y * 10000 + lane * 100 + 4 * q + j);
                    if (__uint_as_float(r[j]) != want) ++bad;
                    if (blockIdx.x == 0 && warp == 5 && y == 1 && q == 2) dump[lane * 4 + j] = __uint_as_float(r[j]);
                }
            }
        if (bad) atomicAdd(mismatches, bad);
    } else {
        const bool do_lds = (mode == 1 || mode == 3 || mode == 5) && warp < n_lds;
        if (mode == 5 && tid == kThreads - 32) {          // a copier thread next to the LDS warps: 64 copies per round, n_cp rounds
            for (int rnd = 0; rnd < n_cp; ++rnd)
                for (int y = 0; y < kTmRows; ++y)
                    for (int q = 0; q < 8; ++q) {
                        const uint64_t d = sw128_desc(saddr(tile) + y * 4096 + q * 16);
                        asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem + y * 32 + 4 * q), "l"(d) : "memory");
                    }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(saddr(&s_bar)) : "memory");
        }
        const bool do_tm = (mode == 2 || mode == 3) && warp >= kWarps - n_tm;
        float acc = 0.f;
        __syncthreads();
        const long long t0 = clock64();
        if (do_lds) {
            const unsigned char* p = tile + lane * 128;
            const int k7 = lane & 7;
            for (int it = 0; it < iters; ++it) {
#pragma unroll 2
                for (int y = kTmRows; y < kTH; ++y) {
                    float4 v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        v[q] = *reinterpret_cast<const float4*>(p + y * 4096 + ((((q + it) & 7) ^ k7) << 4));
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc += (v[q].x + v[q].y) + (v[q].z + v[q].w);
                }
            }
        }
        if (do_tm) {
            for (int it = 0; it < iters; ++it) {
#pragma unroll 2
                for (int y2 = 0; y2 < kTH - kTmRows; ++y2) {       // same number of 512-byte reads as the LDS warps
                    const int y = y2 & (kTmRows - 1);
                    uint32_t r[4][4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) ldtm4(tq + y * 32 + 4 * ((q + it) & 7), r[q]);
                    ldtm_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        acc += (__uint_as_float(r[q][0]) + __uint_as_float(r[q][1])) + (__uint_as_float(r[q][2]) + __uint_as_float(r[q][3]));
                }
            }
        }
        const long long t1 = clock64();
        if (lane == 0) cycles[blockIdx.x * kWarps + warp] = (do_lds || do_tm) ? (unsigned long long)(t1 - t0) : 0ull;
        sink[blockIdx.x * kThreads + tid] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

int main() {
    const int grid = 148, smem = kTileBytes + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int* d_bad; float *d_sink, *d_dump; unsigned long long* d_cyc;
    cudaMalloc(&d_bad, 4); cudaMalloc(&d_sink, grid * kThreads * 4); cudaMalloc(&d_dump, 128 * 4);
    cudaMalloc(&d_cyc, (grid * kWarps + grid) * 8);
    cudaMemset(d_bad, 0, 4); cudaMemset(d_dump, 0, 512);
    probe<<<grid, kThreads, smem>>>(0, 0, 0, 0, 64, d_bad, d_sink, d_cyc, d_dump);
    cudaError_t e = cudaDeviceSynchronize();
    int bad = -1; float dump[128];
    cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(dump, d_dump, 512, cudaMemcpyDeviceToHost);
    unsigned long long cyc[148 * kWarps + 148];
    cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    printf("layout check: %s, mismatches %d of %d; 64 tcgen05.cp + commit + wait: %llu cycles\n", cudaGetErrorString(e), bad,
           grid * kWarps * kTmRows * 8 * 4 * 32, cyc[grid * kWarps]);
    printf("  warp 5, y=1, q=2: lane0 %.0f %.0f %.0f %.0f  lane1 %.0f %.0f  lane31 %.0f (want 10008 10009 10010 10011 / 10108 10109 / 13108)\n",
           dump[0], dump[1], dump[2], dump[3], dump[4], dump[5], dump[124]);
    if (e != cudaSuccess) return 1;
    for (int n_cp : {0, 1, 8, 16, 32, 64}) {               // copy cost: fixed latency or per instruction?
        probe<<<grid, kThreads, smem>>>(4, 0, 0, 0, n_cp, d_bad, d_sink, d_cyc, d_dump);
        cudaDeviceSynchronize();
        cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
        printf("%2d tcgen05.cp + commit + wait: %llu cycles\n", n_cp, cyc[grid * kWarps]);
    }
    for (int mode : {6, 7})
        for (int n_cp : {1, 8, 24, 48}) {
            probe<<<grid, kThreads, smem>>>(mode, 0, 0, 0, n_cp, d_bad, d_sink, d_cyc, d_dump);
            cudaError_t e2 = cudaDeviceSynchronize();
            cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
            printf("%2d tcgen05.cp.%s + commit + wait: %llu cycles (%s)\n", n_cp, mode == 6 ? "128x256b" : "128x128b", cyc[grid * kWarps], cudaGetErrorString(e2));
        }
    const int iters = 2000;
    const double bytes_per_warp = (double)iters * (kTH - kTmRows) * 4 * 512;
    auto run = [&](int mode, int n_lds, int n_tm, const char* what) {
        probe<<<grid, kThreads, smem>>>(mode, iters, n_lds, n_tm, 64, d_bad, d_sink, d_cyc, d_dump);
        cudaDeviceSynchronize();
        cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
        unsigned long long mx_l = 0, mx_t = 0;
        for (int w = 0; w < kWarps; ++w) {
            const bool is_l = (mode == 1 || mode == 3) && w < n_lds;
            if (is_l) mx_l = cyc[w] > mx_l ? cyc[w] : mx_l; else mx_t = cyc[w] > mx_t ? cyc[w] : mx_t;
        }
        const int nl = (mode == 1 || mode == 3) ? n_lds : 0, nt = (mode == 2 || mode == 3) ? n_tm : 0;
        printf("%-34s LDS warps %2d: %7.1f B/clk/SM   tcgen05.ld warps %2d: %7.1f B/clk/SM\n", what, nl,
               nl ? nl * bytes_per_warp / (double)mx_l : 0.0, nt, nt ? nt * bytes_per_warp / (double)mx_t : 0.0);
    };
    run(1, 4, 0, "LDS.128 only");
    run(1, 8, 0, "LDS.128 only");
    run(1, 16, 0, "LDS.128 only");
    run(1, 22, 0, "LDS.128 only");
    run(2, 0, 4, "tcgen05.ld.x4 only");
    run(2, 0, 8, "tcgen05.ld.x4 only");
    run(2, 0, 16, "tcgen05.ld.x4 only");
    run(2, 0, 22, "tcgen05.ld.x4 only");
    for (int rounds : {0, 200, 800}) {                       // LDS.128 rate while a thread issues `rounds` x 64 copies
        probe<<<grid, kThreads, smem>>>(5, iters, 22, 0, rounds, d_bad, d_sink, d_cyc, d_dump);
        cudaDeviceSynchronize();
        cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
        unsigned long long mx = 0;
        for (int w = 0; w < 22; ++w) mx = cyc[w] > mx ? cyc[w] : mx;
        printf("LDS.128, 22 warps, next to %3d x 64 tcgen05.cp: %7.1f B/clk/SM (%llu cycles)\n", rounds, 22 * bytes_per_warp / (double)mx, mx);
    }
    run(3, 16, 8, "both at once");
    run(3, 12, 12, "both at once");
    run(3, 8, 16, "both at once");
    return 0;
}
