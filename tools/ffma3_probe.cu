// Probe: FFMA with three distinct register sources, as in the head's inner loops:  acc[j][p] += w[j] * x[p]
#include <cstdio>
#include <cuda_runtime.h>
template <int NW, int NX>
__global__ void k(float* out, const float* in, int iters, long long* cycles) {
    float acc[NW][NX], w[NW], x[NX];
#pragma unroll
    for (int j = 0; j < NW; ++j) { w[j] = in[j]; for (int p = 0; p < NX; ++p) acc[j][p] = 0.f; }
#pragma unroll
    for (int p = 0; p < NX; ++p) x[p] = in[8 + p + threadIdx.x % 3];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NW; ++j)
#pragma unroll
            for (int p = 0; p < NX; ++p) acc[j][p] = fmaf(w[j], x[p], acc[j][p]);
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NW; ++j) for (int p = 0; p < NX; ++p) s += acc[j][p];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int NW, int NX> void run(float* out, float* in, long long* cyc, const char* name) {
    long long h; const int iters = 4096;
    for (int warps = 4; warps <= 16; warps *= 2) {
        k<NW, NX><<<148, warps * 32>>>(out, in, iters, cyc); cudaDeviceSynchronize();
        k<NW, NX><<<148, warps * 32>>>(out, in, iters, cyc); cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%s warps/SM=%2d cycles per FFMA per SMSP = %.2f\n", name, warps, (double)h / ((double)iters * NW * NX * warps / 4.0));
    }
}
int main() {
    float *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 256); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 256);
    run<2, 9>(out, in, cyc, "acc[2][9] (conv1/conv2)");
    run<8, 9>(out, in, cyc, "acc[8][9] (conv3)     ");
    run<4, 9>(out, in, cyc, "acc[4][9]             ");
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
