// Probe: packed fp32 FMA (fma.rn.f32x2, sm_100+) issue rate vs scalar FFMA.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__global__ void ffma2_kernel(unsigned long long* out, int iters, long long* cycles) {
    unsigned long long c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = (unsigned long long)(i + threadIdx.x) * 0x3f8000003f800000ull;
    unsigned long long a = 0x3f8000013f800001ull, b = 0x3a83126f3a83126full;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma2(c[i], a, b);
    }
    long long t1 = clock64();
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void ffma_kernel(float* out, int iters, long long* cycles) {
    float c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = (float)i;
    float a = threadIdx.x * 1e-3f, b = 1.0001f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], b, a);
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    unsigned long long* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int warps = 4; warps <= 32; warps *= 2) {
        ffma2_kernel<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        ffma2_kernel<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("fma.f32x2 warps/SM=%2d  cycles per instruction per SMSP = %.2f (2 FMAs each)\n", warps, (double)h / ((double)iters * 16 * warps / 4.0));
        ffma_kernel<<<148, warps * 32>>>((float*)out, iters, cyc); cudaDeviceSynchronize();
        ffma_kernel<<<148, warps * 32>>>((float*)out, iters, cyc); cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("ffma      warps/SM=%2d  cycles per instruction per SMSP = %.2f\n", warps, (double)h / ((double)iters * 16 * warps / 4.0));
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
