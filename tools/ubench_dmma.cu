// Micro-benchmark: FP64 tensor-core MMA (mma.sync.aligned.m8n8k4.f64, DMMA in SASS) on B200 at the occupancy of the
// 1,024-instance QP kernel (1-2 warps per SM sub-partition): dependent-chain latency and per-instruction issue interval
// with 1..9 independent accumulator tiles, beside the DFMA figures for the same arithmetic (one m8n8k4 = 256 FMAs = 8
// warp-wide DFMAs).  Decides whether the W = [B A]' L product and the Gram matrix W W' of the stage factorisation
// (mpcb_qp.cuh) are worth moving onto the tensor cores: north_star asks for that decision to be backed by a measurement.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int TILES>
__global__ void dmma_kernel(double *out, long long *cyc, int iters, double a0, double b0)
{
    double c[TILES][2];
#pragma unroll
    for (int t = 0; t < TILES; t++) { c[t][0] = threadIdx.x * 1e-3 + t; c[t][1] = 1.0; }
    const double a = a0 + threadIdx.x * 1e-6, b = b0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int t = 0; t < TILES; t++) dmma(c[t][0], c[t][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int t = 0; t < TILES; t++) s += c[t][0] + c[t][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

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

int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 2048;
    for (int threads : {32, 128, 256, 512, 1024}) {
        printf("--- %d threads/block (%d warps per SM sub-partition), 148 blocks\n", threads, (threads + 127) / 128);
#define RUN(name, kern, ops)                                                       \
    kern; cudaDeviceSynchronize(); kern; cudaDeviceSynchronize();                  \
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);                        \
    printf("%-40s %7.2f cycles per instruction\n", name, (double)h / iters / (ops));
        RUN("DMMA m8n8k4 dependent chain", (dmma_kernel<1><<<148, threads>>>(out, cyc, iters, 0.999, 1e-3)), 1)
        RUN("DMMA 2 independent tiles", (dmma_kernel<2><<<148, threads>>>(out, cyc, iters, 0.999, 1e-3)), 2)
        RUN("DMMA 3 independent tiles", (dmma_kernel<3><<<148, threads>>>(out, cyc, iters, 0.999, 1e-3)), 3)
        RUN("DMMA 6 independent tiles", (dmma_kernel<6><<<148, threads>>>(out, cyc, iters, 0.999, 1e-3)), 6)
        RUN("DMMA 9 independent tiles", (dmma_kernel<9><<<148, threads>>>(out, cyc, iters, 0.999, 1e-3)), 9)
        RUN("DFMA dependent chain", (dfma_kernel<1><<<148, threads>>>(out, cyc, iters, 0.999, 1e-3)), 1)
        RUN("DFMA 8 chains (per DFMA)", (dfma_kernel<8><<<148, threads>>>(out, cyc, iters, 0.999, 1e-3)), 8)
    }
    printf("one m8n8k4 DMMA = 256 FMA = 8 warp-wide DFMA; break-even issue interval = 8 x (DFMA interval)\n");
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
