// Probe: issue rate of legacy mma.sync.m16n8k8 TF32 and of FFMA on this GPU (cycles per warp instruction
// per SM sub-partition), to decide whether a 3xTF32 tensor-core head can beat the FFMA head.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe tools/mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void mma_kernel(float* out, int iters, long long* cycles) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
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
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int warps = 4; warps <= 32; warps *= 2) {
        mma_kernel<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        mma_kernel<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double per_smsp = (double)h / ((double)iters * 8 * warps / 4.0);
        printf("mma.m16n8k8.tf32  warps/SM=%2d  cycles per MMA per SMSP = %.2f  -> %.0f dense TF32 TFLOP/s at 1.9 GHz\n", warps, per_smsp,
               2.0 * 1024 / per_smsp * 4 * 148 * 1.9e9 / 1e12);
        ffma_kernel<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        ffma_kernel<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        per_smsp = (double)h / ((double)iters * 16 * warps / 4.0);
        printf("ffma              warps/SM=%2d  cycles per FFMA per SMSP = %.2f\n", warps, per_smsp);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
