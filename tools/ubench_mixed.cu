// Micro-benchmark for the mixed-precision question, GPU part.  The CPU restatement shows that an
// FP32 stage factorisation is numerically viable in the early interior-point iterations only as a Householder LQ
// (FP32 Gram + Cholesky loses positive definiteness, profiles/r02_mixed_precision_viability.txt).  This times, for one
// BLASTER17 stage matrix [diag(sqrt(Hd)) | W] (23 x 17) per warp and at the occupancy of the 1,024-instance kernel:
//   (a) the shipped FP64 early-iteration path: Gram matrix on the tensor cores (36 DMMAs) + unrolled Cholesky,
//   (b) the shipped FP64 Householder-LQ pivot loop,
//   (c) the same pivot loop in FP32 (float rows in registers, float broadcast row, MUFU.RSQ + one Newton step).
// An FP32 mode pays only if (c) is clearly below (a): it would replace (a), not (b).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mpc_blaster_b200/csrc tools/ubench_mixed.cu -o tools/ubench_mixed
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "mpcb_common.cuh"

using namespace mpcb;
constexpr int NX = 17, NZ = 23, NXP = 20, NI = 3, NJ = 3;

struct Sm {
    alignas(32) double vrow[2][NXP];
    alignas(32) float vrowf[2][NXP];
    double hd[24], ds[24];
    float hdf[24], dsf[24];
    alignas(32) double img[432];   // Gram image (packed blocks) / W rows
    double col[2][24];
    double Lout[NZ * 24];
};

// (b) / (c): Householder LQ pivot loop, rows in registers
template <typename T>
__device__ __forceinline__ T t_rsqrt(T x);
template <>
__device__ __forceinline__ double t_rsqrt<double>(double x) { return fast_rsqrt(x); }
template <>
__device__ __forceinline__ float t_rsqrt<float>(float x)
{
    float y = rsqrtf(x);
    return y * (1.5f - 0.5f * x * y * y);
}
template <typename T>
__device__ __forceinline__ void lq_loop(Sm &sm, T (&w)[NX], int lane)
{
    T *vr0 = sizeof(T) == 8 ? (T *)sm.vrow[0] : (T *)sm.vrowf[0];
    const T *hd = sizeof(T) == 8 ? (const T *)sm.hd : (const T *)sm.hdf;
    const T *ds = sizeof(T) == 8 ? (const T *)sm.ds : (const T *)sm.dsf;
#pragma unroll 1
    for (int j = 0; j < NZ; j++) {
        T *vr = vr0 + (j & 1) * NXP;
        if (lane == j) {
#pragma unroll
            for (int c = 0; c < NX; c++) vr[c] = w[c];
        }
        __syncwarp();
        T v[NX];
#pragma unroll
        for (int c = 0; c < NX; c++) v[c] = vr[c];
        T d0 = 0, d1 = 0, d2 = 0, d3 = 0, e0 = 0, e1 = 0, e2 = 0, e3 = 0;
#pragma unroll
        for (int c = 0; c + 3 < NX; c += 4) {
            d0 += v[c] * w[c]; d1 += v[c + 1] * w[c + 1]; d2 += v[c + 2] * w[c + 2]; d3 += v[c + 3] * w[c + 3];
            e0 += v[c] * v[c]; e1 += v[c + 1] * v[c + 1]; e2 += v[c + 2] * v[c + 2]; e3 += v[c + 3] * v[c + 3];
        }
        d0 += v[NX - 1] * w[NX - 1]; e0 += v[NX - 1] * v[NX - 1];
        const T dot = (d0 + d1) + (d2 + d3), djj = (e0 + e1) + (e2 + e3);
        const T s2v = hd[j] + djj;
        const T rs = t_rsqrt<T>(s2v);
        const T sig = s2v * rs;
        const T kap = (djj > T(0)) ? (sig - ds[j]) / djj : T(0);
        const T lij = (lane > j) ? dot * rs : T(0);
        const T f = lij * kap;
#pragma unroll
        for (int c = 0; c < NX; c++) w[c] -= f * v[c];
        if (lane < NZ && lane >= j) sm.Lout[lane * 24 + j] = (double)(lane == j ? sig : lij);
    }
}

// (a): Gram on the tensor cores from W tiles + unrolled Cholesky (the shipped early-iteration path, mpcb_qp.cuh)
__device__ __forceinline__ void gram_chol(Sm &sm, double (&wt)[NI][NJ][2], double Hd, int lane)
{
    const int tg = lane >> 2, tq = lane & 3;
    double mt[NI][NI][2];
    static_for<0, NI>([&](auto I_) { static_for<0, NI>([&](auto J_) { mt[decltype(I_)::value][decltype(J_)::value][0] = 0.0; mt[decltype(I_)::value][decltype(J_)::value][1] = 0.0; }); });
    static_for<0, NJ>([&](auto C_) {
        constexpr int Jc = decltype(C_)::value;
        static_for<0, 2>([&](auto H_) {
            constexpr int h = decltype(H_)::value;
            static_for<0, NI>([&](auto I_) {
                constexpr int I = decltype(I_)::value;
                static_for<0, I + 1>([&](auto J_) { constexpr int J = decltype(J_)::value; warp_dmma(mt[I][J][0], mt[I][J][1], wt[I][Jc][h], wt[J][Jc][h]); });
            });
        });
    });
    auto moff = [](int I) { return 8 * (4 * I * (I + 1) + 2 * I); };
    auto mstride = [](int I) { return 8 * (I + 1) + 2; };
    if (lane < NZ) sm.hd[lane] = Hd;
    __syncwarp();
    static_for<0, NI>([&](auto I_) {
        constexpr int I = decltype(I_)::value;
        const bool rin = 8 * I + tg < NZ;
        const double hdv = sm.hd[rin ? 8 * I + tg : 0];
        if (rin && (tg >> 1) == tq) { if (tg & 1) mt[I][I][1] += hdv; else mt[I][I][0] += hdv; }
        static_for<0, I + 1>([&](auto J_) {
            constexpr int J = decltype(J_)::value;
            sp_st2<0>(sptr_of(sm.img + moff(I) + tg * mstride(I) + 8 * J + 2 * tq), mt[I][J][0], mt[I][J][1], rin);
        });
    });
    __syncwarp();
    double m[NZ + 1];
    {
        const int r = lane < NZ ? lane : 0;
        const sptr mrow = sptr_of(sm.img + moff(r >> 3) + (r & 7) * mstride(r >> 3));
        static_for<0, NZ, 2>([&](auto Cc) { constexpr int c = decltype(Cc)::value; sp_ld2<c>(mrow, m[c], m[c + 1]); });
    }
    const sptr cb0 = sptr_of(sm.col[0]);
    const sptr cbl = sptr_add(cb0, lane < NZ ? lane : 0);
    static_for<0, NZ>([&](auto J) {
        constexpr int j = decltype(J)::value;
        constexpr int par = (j & 1) * 24;
        sp_st1<par>(cbl, m[j], lane < NZ);
        __syncwarp();
        double a[NZ + 1];
        static_for<(j & ~1), NZ, 2>([&](auto Cc) { constexpr int c = decltype(Cc)::value; sp_ld2<par + c>(cb0, a[c], a[c + 1]); });
        const double rs = fast_rsqrt(a[j]);
        const double f = m[j] * (rs * rs);
#pragma unroll
        for (int c = j + 1; c < NZ; c++) m[c] -= f * a[c];
        if (lane >= j && lane < NZ) sm.Lout[lane * 24 + j] = (lane == j) ? a[j] * rs : m[j] * rs;
    });
}

template <int V>
__global__ void bench_kernel(const double *W0, const double *H0, double *Lall, long long *cyc, int rounds)
{
    __shared__ Sm sm;
    const int lane = threadIdx.x & 31, inst = blockIdx.x;
    const int tg = lane >> 2, tq = lane & 3;
    const double *Wi = W0 + (size_t)(inst % 64) * NZ * NX;
    const double Hd = lane < NZ ? H0[(inst % 64) * NZ + lane] : 1.0;
    if (lane < NZ) { sm.hd[lane] = Hd; sm.ds[lane] = sqrt(Hd); sm.hdf[lane] = (float)Hd; sm.dsf[lane] = sqrtf((float)Hd); }
    __syncwarp();
    double acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
        if (V == 0) {
            double wt[NI][NJ][2];
            for (int I = 0; I < NI; I++)
                for (int J = 0; J < NJ; J++)
                    for (int h = 0; h < 2; h++) {
                        const int row = 8 * I + tg, col = 8 * J + 2 * tq + h;
                        wt[I][J][h] = (row < NZ && col < NX) ? Wi[row * NX + col] + acc * 1e-300 : 0.0;
                    }
            gram_chol(sm, wt, Hd, lane);
            acc += sm.Lout[0];
        } else if (V == 1) {
            double w[NX];
            for (int c = 0; c < NX; c++) w[c] = lane < NZ ? Wi[lane * NX + c] + acc * 1e-300 : 0.0;
            lq_loop<double>(sm, w, lane);
            acc += w[0];
        } else {
            float w[NX];
            for (int c = 0; c < NX; c++) w[c] = lane < NZ ? (float)(Wi[lane * NX + c] + acc * 1e-300) : 0.0f;
            lq_loop<float>(sm, w, lane);
            acc += w[0];
        }
        __syncwarp();
    }
    long long t1 = clock64();
    if (lane < NZ)
        for (int j = 0; j < NZ; j++) Lall[((size_t)inst * NZ + lane) * NZ + j] = (j <= lane) ? sm.Lout[lane * 24 + j] : 0.0;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    if (acc == 1.2345) Lall[0] = acc;
}

int main()
{
    const int NINST = 64;
    std::vector<double> W(NINST * NZ * NX), H(NINST * NZ);
    srand(1);
    for (auto &x : W) x = (rand() / (double)RAND_MAX - 0.5) * 4.0;
    for (size_t i = 0; i < H.size(); i++) H[i] = (i % 5 == 1 ? 3.3e-7 : 0.05 + rand() / (double)RAND_MAX);   // early iterations: no huge barrier terms yet
    double *dW, *dH, *dL;
    long long *dc, hc;
    const int maxinst = 148 * 14;
    cudaMalloc(&dW, W.size() * 8); cudaMalloc(&dH, H.size() * 8); cudaMalloc(&dL, (size_t)maxinst * NZ * NZ * 8); cudaMalloc(&dc, 8);
    cudaMemcpy(dW, W.data(), W.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dH, H.data(), H.size() * 8, cudaMemcpyHostToDevice);
    const int rounds = 200;
    std::vector<double> L0((size_t)64 * NZ * NZ), L1(L0.size());
    const char *names[3] = {"(a) FP64 Gram on tensor cores + Cholesky", "(b) FP64 Householder LQ", "(c) FP32 Householder LQ"};
    for (int bps : {1, 7, 14}) {
        for (int v = 0; v < 3; v++) {
            for (int rep = 0; rep < 2; rep++) {
                if (v == 0) bench_kernel<0><<<148 * bps, 32>>>(dW, dH, dL, dc, rounds);
                else if (v == 1) bench_kernel<1><<<148 * bps, 32>>>(dW, dH, dL, dc, rounds);
                else bench_kernel<2><<<148 * bps, 32>>>(dW, dH, dL, dc, rounds);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
            cudaMemcpy((v == 1 ? L0 : L1).data(), dL, L0.size() * 8, cudaMemcpyDeviceToHost);
            if (v == 0) { printf("warps/SM %2d  %-44s %8.0f cycles per stage factorisation\n", bps, names[v], (double)hc / rounds); continue; }
            double md = 0;
            if (v == 2) for (size_t i = 0; i < L0.size(); i++) md = fmax(md, fabs(L0[i] - L1[i]) / (1.0 + fabs(L0[i])));
            printf("warps/SM %2d  %-44s %8.0f cycles per stage factorisation%s", bps, names[v], (double)hc / rounds, v == 2 ? "" : "\n");
            if (v == 2) printf("   max rel |L - L_fp64| = %.1e\n", md);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
