// Micro-benchmark: FP64 latency / per-warp issue interval on B200 at low occupancy
// (1 or 2 warps per SM sub-partition), the regime of the 1,024-instance QP kernel.
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_kernel(double *out, long long *cyc, int iters, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x * 1e-3 + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void shfl_kernel(double *out, long long *cyc, int iters)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31) + 1.0;
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void lds_kernel(double *out, long long *cyc, int iters)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 7 + 1) % 1024;
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) idx = (int)sm[idx];
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = idx;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void rsqrt_kernel(double *out, long long *cyc, int iters)
{
    double x = 2.0 + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) x = rsqrt(x) + 3.0;
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 4096;
    for (int threads : {32, 128, 256, 512, 1024}) {
        printf("--- %d threads/block (%d warps per SM sub-partition), 148 blocks\n", threads, (threads + 127) / 128);
#define RUN(name, kern, ops)                                                       \
    kern; cudaDeviceSynchronize(); kern; cudaDeviceSynchronize();                  \
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);                        \
    printf("%-28s %7.2f cycles per op\n", name, (double)h / iters / (ops));
        RUN("DFMA dependent chain", (dfma_kernel<1><<<148, threads>>>(out, cyc, iters, 0.999, 1e-3)), 1)
        RUN("DFMA 2 chains (per DFMA)", (dfma_kernel<2><<<148, threads>>>(out, cyc, iters, 0.999, 1e-3)), 2)
        RUN("DFMA 4 chains (per DFMA)", (dfma_kernel<4><<<148, threads>>>(out, cyc, iters, 0.999, 1e-3)), 4)
        RUN("DFMA 8 chains (per DFMA)", (dfma_kernel<8><<<148, threads>>>(out, cyc, iters, 0.999, 1e-3)), 8)
        RUN("DFMA 16 chains (per DFMA)", (dfma_kernel<16><<<148, threads>>>(out, cyc, iters, 0.999, 1e-3)), 16)
        RUN("SHFL.64 + DADD dependent", (shfl_kernel<<<148, threads>>>(out, cyc, iters)), 1)
        RUN("LDS.64 + cvt dependent", (lds_kernel<<<148, threads>>>(out, cyc, iters)), 1)
        RUN("rsqrt(double) + DADD dep.", (rsqrt_kernel<<<148, threads>>>(out, cyc, iters)), 1)
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
