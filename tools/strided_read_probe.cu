// Probe: how fast does HBM deliver a feature map [B,256,272,480] fp32 when it is read the way roi_tile_kernel stages it:
// 296 persistent CTAs (2 per SM) x 512 threads, a ticket = 32 channel planes x 24 rows x SEG bytes (SEG = tile width * 4),
// every thread has all its loads of a ticket in flight, tickets handed out by an atomic counter in tile order with the 8
// channel groups of a tile back to back (the kernel's order).  No compute: this is the staging phase alone.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kB = 8, kC = 256, kH = 272, kW = 480, kTH = 24, kTC = 32;

template <int SEG_FLOATS>          // floats per row segment: 32 (128 B), 64 (256 B), 128 (512 B)
__global__ void __launch_bounds__(512, 2) stage_only(const float* __restrict__ feat, int* __restrict__ ticket, float* __restrict__ sink) {
    constexpr int kPerThread = kTC * kTH * SEG_FLOATS / 512;          // loads per thread per ticket (48 at 128 B)
    constexpr int kTilesX = kW / SEG_FLOATS + (kW % SEG_FLOATS ? 1 : 0);
    constexpr int kTilesY = (kH + kTH - 1) / kTH;
    const int n_tickets = kB * kTilesY * kTilesX * (kC / kTC);
    __shared__ int s_t;
    float acc = 0.f;
    for (;;) {
        if (threadIdx.x == 0) s_t = atomicAdd(ticket, 1);
        __syncthreads();
        const int t = s_t;
        __syncthreads();
        if (t >= n_tickets) break;
        const int g = t % (kC / kTC), tile = t / (kC / kTC);
        const int tx = tile % kTilesX, ty = (tile / kTilesX) % kTilesY, img = tile / (kTilesX * kTilesY);
        float v[kPerThread];
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            const int e = threadIdx.x + i * 512;                      // element of the [32][24][SEG] tile
            const int x = e % SEG_FLOATS, y = (e / SEG_FLOATS) % kTH, c = e / (SEG_FLOATS * kTH);
            const int gx = min(tx * SEG_FLOATS + x, kW - 1), gy = min(ty * kTH + y, kH - 1);
            v[i] = __ldg(feat + (((size_t)img * kC + g * kTC + c) * kH + gy) * kW + gx);
        }
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) acc += v[i];
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <int SEG_FLOATS>
void run(const float* feat, int* ticket, float* sink, const char* name) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemset(ticket, 0, 4);
        cudaEventRecord(a);
        stage_only<SEG_FLOATS><<<296, 512>>>(feat, ticket, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
    }
    const double bytes = (double)kB * kC * kH * kW * 4;
    printf("%-28s %.3f ms  %.0f GB/s   (%s)\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t n = (size_t)kB * kC * kH * kW;
    float *feat, *sink; int* ticket;
    cudaMalloc(&feat, n * 4); cudaMalloc(&sink, 4); cudaMalloc(&ticket, 4);
    cudaMemset(feat, 0, n * 4);
    run<32>(feat, ticket, sink, "128-byte row segments");
    run<64>(feat, ticket, sink, "256-byte row segments");
    run<128>(feat, ticket, sink, "512-byte row segments");
    return 0;
}
