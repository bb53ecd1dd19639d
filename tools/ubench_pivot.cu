// Micro-benchmark of the Householder-LQ pivot loop of qp_kernel (mpcb_qp.cuh), isolated:
// cycles per pivot for one warp alone on an SM sub-partition and with several resident warps,
// for the loop variants tried in round 1.  Each variant factorises the same [diag | W] rows, so
// their L factors can be compared.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mpc_blaster_b200/csrc tools/ubench_pivot.cu -o tools/ubench_pivot
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "mpcb_common.cuh"

using namespace mpcb;
constexpr int NX = 17, NU = 6, NZ = 23, NXP = 20;

struct Sm {
    alignas(32) double vrow[2][NXP];
    double Lout[NZ * 24];
    double nrmb[2][4];
};

// V0: the round-1 loop (sigma from the pivot lane's own dot product)
__device__ __forceinline__ void pivots_v0(Sm &sm, double (&w)[NX], double Hd, double dsq, int lane)
{
    const sptr vr0 = sptr_of(sm.vrow[0]);
    const sptr lrow = sptr_of(sm.Lout + (lane < NZ ? lane : 0) * 24);
#pragma unroll 1
    for (int j = 0; j < NZ; j++) {
        const sptr vr = sptr_add(vr0, (j & 1) * NXP);
        const double hdj = warp_shfl(Hd, j), dsj = warp_shfl(dsq, j);
        const bool piv = (lane == j);
        double v[NX];
        sp_row_store<0, NX>(vr, w, piv);
        warp_sync();
        sp_row_load<0, NX>(vr, v);
        double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
        for (int c = 0; c + 3 < NX; c += 4) { d0 += v[c] * w[c]; d1 += v[c + 1] * w[c + 1]; d2 += v[c + 2] * w[c + 2]; d3 += v[c + 3] * w[c + 3]; }
#pragma unroll
        for (int c = NX & ~3; c < NX; c++) d0 += v[c] * w[c];
        const double dot = (d0 + d1) + (d2 + d3);
        const double djj = warp_shfl(dot, j);
        const double s2v = hdj + djj;
        const double rs = fast_rsqrt(s2v);
        const double idjj = fast_rcp(djj);
        const double sig = s2v * rs;
        const double kap = (djj > 0) ? (sig - dsj) * idjj : 0.0;
        const double lij = (lane > j) ? dot * rs : 0.0;
        const double f = lij * kap;
#pragma unroll
        for (int c = 0; c < NX; c++) w[c] -= f * v[c];
        sp_st1<0>(sptr_add(lrow, j), piv ? sig : lij, lane < NZ && lane >= j);
    }
}

// V1..: every lane keeps |w|^2 of its own row up to date (same summation order as the dot), so the
// pivot's sigma / beta chain starts from a shuffle right after the update and runs beside the
// row broadcast + dot instead of after them.
//   SHORT : sigma refined beside rsqrt (sig0 = s2v*y0) instead of after it
//   BCAST : 0 = 128-bit shared loads, 1 = 64-bit shared loads, 2 = shuffles
MPCB_DEV void rsqrt_sig(double x, double &rs, double &sig)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double s0 = x * y;
    const double e = fma(x, -(y * y), 1.0);
    const double c = fma(e, 0.375, 0.5);
    rs = fma(c, y * e, y);
    sig = fma(c, s0 * e, s0);
}
template <bool SHORT, int BCAST, int EXP = 0>
__device__ __forceinline__ void pivots_v1(Sm &sm, double (&w)[NX], double Hd, double dsq, int lane)
{
    const sptr vr0 = sptr_of(sm.vrow[0]);
    const sptr lrow = sptr_of(sm.Lout + (lane < NZ ? lane : 0) * 24);
    double nrm;
    {
        double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
        for (int c = 0; c + 3 < NX; c += 4) { d0 += w[c] * w[c]; d1 += w[c + 1] * w[c + 1]; d2 += w[c + 2] * w[c + 2]; d3 += w[c + 3] * w[c + 3]; }
#pragma unroll
        for (int c = NX & ~3; c < NX; c++) d0 += w[c] * w[c];
        nrm = (d0 + d1) + (d2 + d3);
    }
#pragma unroll 1
    for (int j = 0; j < NZ; j++) {
        const sptr vr = sptr_add(vr0, (j & 1) * NXP);
        const bool piv = (lane == j);
        double v[NX];
        if (BCAST < 2) sp_row_store<0, NX>(vr, w, piv);
        const double hdj = warp_shfl(Hd, j), dsj = warp_shfl(dsq, j);
        const double djj = warp_shfl(nrm, j);
        if (EXP == 2) {
#pragma unroll
            for (int c = 0; c < NX; c++) v[c] = w[c] * 0.5 + hdj;
        } else if (BCAST == 0) { warp_sync(); sp_row_load<0, NX>(vr, v); }
        else if (BCAST == 1) {
            warp_sync();
#pragma unroll
            for (int c = 0; c < NX; c++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[c]) : "r"(vr + 8u * c) : "memory");
        } else {
#pragma unroll
            for (int c = 0; c < NX; c++) v[c] = warp_shfl(w[c], j);
        }
        const double s2v = hdj + djj;
        double rs, sig;
        if (EXP == 1) { rs = s2v * 1e-3; sig = s2v * 0.5; }
        else if (SHORT) rsqrt_sig(s2v, rs, sig);
        else { rs = fast_rsqrt(s2v); sig = s2v * rs; }
        const double idjj = (EXP == 1) ? djj * 1e-4 : fast_rcp(djj);
        const double beta = (djj > 0) ? (SHORT ? (sig - dsj) * (idjj * rs) : ((sig - dsj) * idjj) * rs) : 0.0;
        double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
        for (int c = 0; c + 3 < NX; c += 4) { d0 += v[c] * w[c]; d1 += v[c + 1] * w[c + 1]; d2 += v[c + 2] * w[c + 2]; d3 += v[c + 3] * w[c + 3]; }
#pragma unroll
        for (int c = NX & ~3; c < NX; c++) d0 += v[c] * w[c];
        const double dot = (EXP == 3) ? v[0] * w[0] : (d0 + d1) + (d2 + d3);
        const double f = (lane > j) ? dot * beta : 0.0;
        const double lij = (lane > j) ? dot * rs : 0.0;
        double n0 = 0, n1 = 0, n2 = 0, n3 = 0;
#pragma unroll
        for (int c = 0; c + 3 < NX; c += 4) {
            w[c] -= f * v[c]; w[c + 1] -= f * v[c + 1]; w[c + 2] -= f * v[c + 2]; w[c + 3] -= f * v[c + 3];
            n0 += w[c] * w[c]; n1 += w[c + 1] * w[c + 1]; n2 += w[c + 2] * w[c + 2]; n3 += w[c + 3] * w[c + 3];
        }
#pragma unroll
        for (int c = NX & ~3; c < NX; c++) { w[c] -= f * v[c]; n0 += w[c] * w[c]; }
        nrm = (EXP == 4) ? nrm - f * dot : (n0 + n1) + (n2 + n3);
        sp_st1<0>(sptr_add(lrow, j), piv ? sig : lij, lane < NZ && lane >= j);
    }
}

template <int V>
__global__ void bench_kernel(const double *W0, const double *H0, double *Lall, long long *cyc, int rounds)
{
    __shared__ Sm sm[1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Sm &s = sm[warp];
    const int inst = blockIdx.x * (blockDim.x >> 5) + warp;
    double w0[NX];
    for (int c = 0; c < NX; c++) w0[c] = lane < NZ ? W0[((size_t)(inst % 64) * NZ + lane) * NX + c] : 0.0;
    const double Hd = lane < NZ ? H0[(inst % 64) * NZ + lane] : 1.0;
    const double dsq = sqrt(Hd);
    double acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
        double w[NX];
#pragma unroll
        for (int c = 0; c < NX; c++) w[c] = w0[c] + acc * 1e-300;
        if (V == 0) pivots_v0(s, w, Hd, dsq, lane);
        else if (V == 1) pivots_v1<false, 0>(s, w, Hd, dsq, lane);
        else if (V == 2) pivots_v1<true, 0>(s, w, Hd, dsq, lane);
        else if (V == 3) pivots_v1<true, 1>(s, w, Hd, dsq, lane);
        else if (V == 4) pivots_v1<true, 2>(s, w, Hd, dsq, lane);
        else if (V == 5) pivots_v1<false, 0, 1>(s, w, Hd, dsq, lane);
        else if (V == 6) pivots_v1<false, 0, 2>(s, w, Hd, dsq, lane);
        else if (V == 7) pivots_v1<false, 0, 3>(s, w, Hd, dsq, lane);
        else pivots_v1<false, 0, 4>(s, w, Hd, dsq, lane);
        acc += w[0];
        __syncwarp();
    }
    long long t1 = clock64();
    if (lane < NZ)
        for (int j = 0; j < NZ; j++) Lall[((size_t)inst * NZ + lane) * NZ + j] = (j <= lane) ? s.Lout[lane * 24 + j] : 0.0;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    if (acc == 1.2345) Lall[0] = acc;
}

// V9: two warps per instance, the 17 columns of W split 9 + 8 between them; partial dot products and partial
// |w_j|^2 exchanged through shared memory with one 64-thread barrier per pivot (double-buffered by pivot parity).
struct Sm2 {
    alignas(32) double vrow[2][2][12];   // [warp][parity][own columns]
    double part[2][2][32];               // [parity][warp][lane]  partial dots
    double pn[2][2];                     // [parity][warp]        partial |w_j|^2
    double hd[24], ds[24];
    double Lout[NZ * 24];
};
template <int NC>
__device__ __forceinline__ void pivots_2w(Sm2 &sm, double (&w)[NC], int warp, int lane)
{
    const sptr lrow = sptr_of(sm.Lout + (lane < NZ ? lane : 0) * 24);
#pragma unroll 1
    for (int j = 0; j < NZ; j++) {
        const int par = j & 1;
        const sptr vr = sptr_of(sm.vrow[warp][par]);
        const bool piv = (lane == j);
        double v[NC];
        sp_row_store<0, NC>(vr, w, piv);
        warp_sync();
        sp_row_load<0, NC>(vr, v);
        double d0 = 0, d1 = 0, e0 = 0, e1 = 0;
#pragma unroll
        for (int c = 0; c + 1 < NC; c += 2) { d0 += v[c] * w[c]; d1 += v[c + 1] * w[c + 1]; e0 += v[c] * v[c]; e1 += v[c + 1] * v[c + 1]; }
        if (NC & 1) { d0 += v[NC - 1] * w[NC - 1]; e0 += v[NC - 1] * v[NC - 1]; }
        sm.part[par][warp][lane] = d0 + d1;
        if (lane == 0) sm.pn[par][warp] = e0 + e1;
        asm volatile("bar.sync 1, 64;" ::: "memory");
        const double dot = sm.part[par][0][lane] + sm.part[par][1][lane];
        const double djj = sm.pn[par][0] + sm.pn[par][1];
        const double hdj = sm.hd[j], dsj = sm.ds[j];
        const double s2v = hdj + djj;
        const double rs = fast_rsqrt(s2v);
        const double idjj = fast_rcp(djj);
        const double sig = s2v * rs;
        const double kap = (djj > 0) ? (sig - dsj) * idjj : 0.0;
        const double lij = (lane > j) ? dot * rs : 0.0;
        const double f = lij * kap;
#pragma unroll
        for (int c = 0; c < NC; c++) w[c] -= f * v[c];
        if (warp == 0) sp_st1<0>(sptr_add(lrow, j), piv ? sig : lij, lane < NZ && lane >= j);
    }
}
__global__ void __launch_bounds__(64) bench_kernel2w(const double *W0, const double *H0, double *Lall, long long *cyc, int rounds)
{
    __shared__ Sm2 s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inst = blockIdx.x;
    constexpr int NC0 = 9, NC1 = 8;
    double w0[NC0];
    for (int c = 0; c < NC0; c++) {
        const int col = warp == 0 ? c : NC0 + c;
        w0[c] = (lane < NZ && col < NX) ? W0[((size_t)(inst % 64) * NZ + lane) * NX + col] : 0.0;
    }
    if (warp == 0 && lane < NZ) { s.hd[lane] = H0[(inst % 64) * NZ + lane]; s.ds[lane] = sqrt(s.hd[lane]); }
    __syncthreads();
    double acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
        if (warp == 0) {
            double w[NC0];
#pragma unroll
            for (int c = 0; c < NC0; c++) w[c] = w0[c] + acc * 1e-300;
            pivots_2w<NC0>(s, w, 0, lane);
            acc += w[0];
        } else {
            double w[NC1];
#pragma unroll
            for (int c = 0; c < NC1; c++) w[c] = w0[c] + acc * 1e-300;
            pivots_2w<NC1>(s, w, 1, lane);
            acc += w[0];
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (warp == 0 && lane < NZ)
        for (int j = 0; j < NZ; j++) Lall[((size_t)inst * NZ + lane) * NZ + j] = (j <= lane) ? s.Lout[lane * 24 + j] : 0.0;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    if (acc == 1.2345) Lall[0] = acc;
}

int main()
{
    const int NI = 64;
    std::vector<double> W(NI * NZ * NX), H(NI * NZ);
    srand(1);
    for (auto &x : W) x = (rand() / (double)RAND_MAX - 0.5) * 4.0;
    for (size_t i = 0; i < H.size(); i++) H[i] = (i % 7 == 3) ? 1e12 : (i % 5 == 1 ? 3.3e-7 : 0.05 + rand() / (double)RAND_MAX);
    double *dW, *dH, *dL;
    long long *dc, hc;
    const int maxinst = 148 * 8 * 8;
    cudaMalloc(&dW, W.size() * 8); cudaMalloc(&dH, H.size() * 8); cudaMalloc(&dL, (size_t)maxinst * NZ * NZ * 8); cudaMalloc(&dc, 8);
    cudaMemcpy(dW, W.data(), W.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dH, H.data(), H.size() * 8, cudaMemcpyHostToDevice);
    const int rounds = 200;
    std::vector<double> L0((size_t)64 * NZ * NZ), L1(L0.size());
    for (int bps : {1, 7, 14}) {   // one-warp blocks per SM, like qp_kernel
        for (int v = 0; v < 9; v++) {
            for (int rep = 0; rep < 2; rep++) {
                const dim3 g(148 * bps), b(32);
                switch (v) {
                case 0: bench_kernel<0><<<g, b>>>(dW, dH, dL, dc, rounds); break;
                case 1: bench_kernel<1><<<g, b>>>(dW, dH, dL, dc, rounds); break;
                case 2: bench_kernel<2><<<g, b>>>(dW, dH, dL, dc, rounds); break;
                case 3: bench_kernel<3><<<g, b>>>(dW, dH, dL, dc, rounds); break;
                case 4: bench_kernel<4><<<g, b>>>(dW, dH, dL, dc, rounds); break;
                case 5: bench_kernel<5><<<g, b>>>(dW, dH, dL, dc, rounds); break;
                case 6: bench_kernel<6><<<g, b>>>(dW, dH, dL, dc, rounds); break;
                case 7: bench_kernel<7><<<g, b>>>(dW, dH, dL, dc, rounds); break;
                default: bench_kernel<8><<<g, b>>>(dW, dH, dL, dc, rounds); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
            cudaMemcpy((v ? L1 : L0).data(), dL, L0.size() * 8, cudaMemcpyDeviceToHost);
            double md = 0;
            for (size_t i = 0; i < L0.size(); i++) md = fmax(md, fabs(L0[i] - L1[i]));
            printf("warps/SM %2d  variant %d: %7.1f cycles per pivot   max|L - L_v0| = %.2e\n", bps, v, (double)hc / rounds / NZ, md);
        }
    }
    for (int bps : {1, 7, 14}) {
        for (int rep = 0; rep < 2; rep++) { bench_kernel2w<<<148 * bps, 64>>>(dW, dH, dL, dc, rounds); cudaDeviceSynchronize(); }
        cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(L1.data(), dL, L0.size() * 8, cudaMemcpyDeviceToHost);
        double md = 0;
        for (size_t i = 0; i < L0.size(); i++) md = fmax(md, fabs(L0[i] - L1[i]));
        printf("instances/SM %2d  two warps per instance (columns 9+8): %7.1f cycles per pivot   max|L - L_v0| = %.2e\n", bps, (double)hc / rounds / NZ, md);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
