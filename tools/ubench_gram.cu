// Micro-benchmark of one stage factorisation of qp_kernel (mpcb_qp.cuh), isolated: L L' = diag(Hd) + W W' for one
// instance per warp (lane i owns row i of the 23 x 17 matrix W), cycles per factorisation with 1 / 7 / 8 one-warp
// blocks per SM.  Variants:
//   0  Householder LQ of [sqrt(Hd) | W]               (the loop qp_kernel runs once mu <= MPCB_GRAM_MU)
//   1  Gram matrix, unrolled over the rows + Cholesky  (what qp_kernel runs while mu > MPCB_GRAM_MU; round 1, session 4)
//   2  "window" Gram + the same Cholesky               (NOT in the product -- written at the end of round 1 without GPU
//      time left to run it; a lever for the next round)
// Variant 2: M is symmetric of odd order 23, so the entries {(i, (i - d) mod 23) : d = 0..11} over all lanes i cover every
// unordered pair exactly once -- 12 dot products per lane instead of 23, rolled (small code: variant 1 is 11 KB of SASS
// and a third of its samples are instruction-cache misses).  Results go to shared memory as a packed lower triangle
// (rows padded to even length for 128-bit loads) and every lane reads its row back.
// All variants factorise the same data; the tool prints max|L - L_variant0| so a wrong variant shows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mpc_blaster_b200/csrc tools/ubench_gram.cu -o tools/ubench_gram
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "mpcb_common.cuh"

using namespace mpcb;
constexpr int NX = 17, NU = 6, NZ = 23, NXP = 20, NZP = 24;
constexpr int LDW = (NX + 1) & ~1;       // 18: rows of the W image stay 16-byte aligned
constexpr int ND = NZ / 2 + 1;           // window length (12)
constexpr int MPK = 288;                 // packed lower triangle, rows padded to even length

struct Sm {
    alignas(32) double Wsh[NZ * LDW + 2];
    alignas(32) double Msh[MPK];
    alignas(32) double cb[2 * NZP];
    alignas(32) double vrow[2][NXP];
    double hd[NZP], ds[NZP];
    double Lout[NZ * NZP];
};

// offset of row r in the packed triangle: rows 2a and 2a+1 both take 2a+2 entries
__host__ __device__ constexpr int row_off(int r) { return 2 * (r >> 1) * ((r >> 1) + 1) + (r & 1) * (2 * (r >> 1) + 2); }
static_assert(row_off(NZ - 1) + NZP <= MPK, "packed triangle does not fit");

// ---- variant 0: the product's Householder loop
__device__ __forceinline__ void factor_lq(Sm &sm, double (&w)[NX], double Hd, int lane)
{
    const double dsq = sqrt(Hd);
    if (lane < NZ) { sm.hd[lane] = Hd; sm.ds[lane] = dsq; }
    if (lane >= NZ) {
#pragma unroll
        for (int c = 0; c < NX; c++) w[c] = 0.0;
    }
    const sptr vr0 = sptr_of(sm.vrow[0]);
    const sptr lrow = sptr_of(sm.Lout + (lane < NZ ? lane : 0) * NZP);
#pragma unroll 1
    for (int j = 0; j < NZ; j++) {
        const sptr vr = sptr_add(vr0, (j & 1) * NXP);
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
        const double hdj = sm.hd[j], dsj = sm.ds[j];
        double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
#pragma unroll
        for (int c = 0; c + 3 < NX; c += 4) { e0 += v[c] * v[c]; e1 += v[c + 1] * v[c + 1]; e2 += v[c + 2] * v[c + 2]; e3 += v[c + 3] * v[c + 3]; }
#pragma unroll
        for (int c = NX & ~3; c < NX; c++) e0 += v[c] * v[c];
        const double djj = (e0 + e1) + (e2 + e3);
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

// ---- right-looking Cholesky with row `lane` of M in registers (as in qp_kernel)
__device__ __forceinline__ void cholesky_rows(Sm &sm, double (&m)[NZ + 1], int lane)
{
    const sptr cb0 = sptr_of(sm.cb);
    const sptr cbl = sptr_add(cb0, lane < NZ ? lane : 0);
    const sptr lrow = sptr_of(sm.Lout + (lane < NZ ? lane : 0) * NZP);
    static_for<0, NZ>([&](auto J) {
        constexpr int j = decltype(J)::value;
        constexpr int par = (j & 1) * NZP;
        sp_st1<par>(cbl, m[j], lane < NZ);
        warp_sync();
        double a[NZ + 1];
        static_for<(j & ~1), NZ, 2>([&](auto Cc) {
            constexpr int c = decltype(Cc)::value;
            sp_ld2<par + c>(cb0, a[c], a[c + 1]);
        });
        const double rs = fast_rsqrt(a[j]);
        const double sig = a[j] * rs;
        const double f = m[j] * (rs * rs);
        const double lij = m[j] * rs;
#pragma unroll
        for (int c = j + 1; c < NZ; c++) m[c] -= f * a[c];
        sp_st1<j>(lrow, (lane == j) ? sig : lij, lane < NZ && lane >= j);
    });
}

// ---- variant 1: Gram unrolled over the rows (every lane forms all 23 entries of its row)
__device__ __forceinline__ void factor_gram_rows(Sm &sm, double (&w)[NX], double Hd, int lane)
{
    if (lane >= NZ) {
#pragma unroll
        for (int c = 0; c < NX; c++) w[c] = 0.0;
    }
    sp_row_store<0, NX>(sptr_of(sm.Wsh + (lane < NZ ? lane : 0) * LDW), w, lane < NZ);
    warp_sync();
    const sptr w0 = sptr_of(sm.Wsh);
    double m[NZ + 1];
    m[NZ] = 0.0;
    static_for<0, NZ>([&](auto C) {
        constexpr int c = decltype(C)::value;
        double v[NX];
        sp_row_load<0, NX>(sptr_add(w0, c * LDW), v);
        double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
        for (int i = 0; i + 3 < NX; i += 4) { d0 += v[i] * w[i]; d1 += v[i + 1] * w[i + 1]; d2 += v[i + 2] * w[i + 2]; d3 += v[i + 3] * w[i + 3]; }
#pragma unroll
        for (int i = NX & ~3; i < NX; i++) d0 += v[i] * w[i];
        m[c] = ((d0 + d1) + (d2 + d3)) + (lane == c ? Hd : 0.0);
    });
    cholesky_rows(sm, m, lane);
}

// ---- variant 2: window Gram (12 entries per lane, rolled) through a packed triangle in shared memory
__device__ __forceinline__ void factor_gram_window(Sm &sm, double (&w)[NX], double Hd, int lane)
{
    const int i = lane < NZ ? lane : 0;
    sp_row_store<0, NX>(sptr_of(sm.Wsh + i * LDW), w, lane < NZ);
    warp_sync();
    const sptr w0 = sptr_of(sm.Wsh);
#pragma unroll 2
    for (int d = 0; d < ND; d++) {
        int c = i - d;
        if (c < 0) c += NZ;
        double v[NX];
        sp_row_load<0, NX>(sptr_add(w0, c * LDW), v);
        double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
        for (int q = 0; q + 3 < NX; q += 4) { d0 += v[q] * w[q]; d1 += v[q + 1] * w[q + 1]; d2 += v[q + 2] * w[q + 2]; d3 += v[q + 3] * w[q + 3]; }
#pragma unroll
        for (int q = NX & ~3; q < NX; q++) d0 += v[q] * w[q];
        const double val = ((d0 + d1) + (d2 + d3)) + (d == 0 ? Hd : 0.0);
        const int r = max(i, c), q = min(i, c);
        if (lane < NZ) sm.Msh[row_off(r) + q] = val;
    }
    warp_sync();
    double m[NZ + 1];
    const sptr mrow = sptr_of(sm.Msh + row_off(i));
    static_for<0, NZ, 2>([&](auto Cc) {
        constexpr int c = decltype(Cc)::value;
        sp_ld2<c>(mrow, m[c], m[c + 1]);   // entries beyond the row's own length belong to later rows: never used
    });
    cholesky_rows(sm, m, lane);
}

template <int V>
__global__ void __launch_bounds__(32) bench_kernel(const double *W0, const double *H0, double *Lall, long long *cyc, int rounds)
{
    __shared__ Sm s;
    const int lane = threadIdx.x & 31;
    const int inst = blockIdx.x;
    double w0[NX];
    for (int c = 0; c < NX; c++) w0[c] = lane < NZ ? W0[((size_t)(inst % 64) * NZ + lane) * NX + c] : 0.0;
    const double Hd = lane < NZ ? H0[(inst % 64) * NZ + lane] : 1.0;
    double acc = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
        double w[NX];
#pragma unroll
        for (int c = 0; c < NX; c++) w[c] = w0[c] + acc * 1e-300;
        if (V == 0) factor_lq(s, w, Hd, lane);
        else if (V == 1) factor_gram_rows(s, w, Hd, lane);
        else factor_gram_window(s, w, Hd, lane);
        __syncwarp();
        acc += s.Lout[(lane < NZ ? lane : 0) * NZP];
        __syncwarp();
    }
    const long long t1 = clock64();
    if (lane < NZ)
        for (int j = 0; j < NZ; j++) Lall[((size_t)inst * NZ + lane) * NZ + j] = (j <= lane) ? s.Lout[lane * NZP + j] : 0.0;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    if (acc == 1.2345) Lall[0] = acc;
}

int main()
{
    const int NI = 64;
    std::vector<double> W(NI * NZ * NX), H(NI * NZ);
    srand(1);
    for (auto &x : W) x = (rand() / (double)RAND_MAX - 0.5) * 4.0;
    // early-iteration magnitudes (mu > 1e-4): barrier terms up to ~1e6, the smallest curvature dt*1e-5
    for (size_t i = 0; i < H.size(); i++) H[i] = (i % 7 == 3) ? 1e6 : (i % 5 == 1 ? 3.3e-7 : 0.05 + rand() / (double)RAND_MAX);
    double *dW, *dH, *dL;
    long long *dc, hc;
    const int maxinst = 148 * 8;
    cudaMalloc(&dW, W.size() * 8); cudaMalloc(&dH, H.size() * 8); cudaMalloc(&dL, (size_t)maxinst * NZ * NZ * 8); cudaMalloc(&dc, 8);
    cudaMemcpy(dW, W.data(), W.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dH, H.data(), H.size() * 8, cudaMemcpyHostToDevice);
    const int rounds = 200;
    std::vector<double> L0((size_t)64 * NZ * NZ), L1(L0.size());
    for (int bps : {1, 7, 8}) {
        for (int v = 0; v < 3; v++) {
            for (int rep = 0; rep < 2; rep++) {
                const dim3 g(148 * bps), b(32);
                if (v == 0) bench_kernel<0><<<g, b>>>(dW, dH, dL, dc, rounds);
                else if (v == 1) bench_kernel<1><<<g, b>>>(dW, dH, dL, dc, rounds);
                else bench_kernel<2><<<g, b>>>(dW, dH, dL, dc, rounds);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
            cudaMemcpy((v ? L1 : L0).data(), dL, L0.size() * 8, cudaMemcpyDeviceToHost);
            double md = 0, mx = 0;
            for (size_t i = 0; i < L0.size(); i++) { md = fmax(md, fabs(L0[i] - L1[i])); mx = fmax(mx, fabs(L0[i])); }
            printf("warps/SM %2d  variant %d: %8.1f cycles per stage factorisation   max|L - L_v0| = %.2e (max|L| = %.2e)\n", bps, v,
                   (double)hc / rounds, v ? md : 0.0, mx);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
