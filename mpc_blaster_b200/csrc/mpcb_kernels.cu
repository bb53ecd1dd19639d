// __global__ wrappers, host batch scheduler and the C ABI (include/mpcb.h) of the
// B200-native batched BLASTER MPC solver.  Built for sm_100a only; there is no CPU path.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "../../include/mpcb.h"
#include "mpcb_common.cuh"
#include "mpcb_linearize.cuh"
#include "mpcb_qp.cuh"
#include "mpcb_qp8.cuh"
#include "mpcb_poc.cuh"

using namespace mpcb;

namespace {

std::atomic<int64_t> g_launches{0};
thread_local std::string g_create_error;  // message of the last failed create / stateless call on this thread

// ------------------------------------------------------------------ kernels
__device__ __forceinline__ const double *param_ptr(const double *p, int p_mode, int inst, int k, int N)
{
    if (p_mode == MPCB_SHARED) return p;
    if (p_mode == MPCB_PER_INSTANCE) return p + (size_t)inst * kNP;
    return p + ((size_t)inst * N + k) * kNP;
}

// K1: one warp per (instance, shooting interval).
template <int NX, int NU>
__global__ void __launch_bounds__(128) linearize_kernel(const __grid_constant__ Params P, const double *__restrict__ X,
                                                        const double *__restrict__ U, const double *__restrict__ p,
                                                        int p_mode, double *__restrict__ ws, int inst0, int B,
                                                        const int32_t *__restrict__ skip)
{
    using L = Layout<NX, NU>;
    const int N = P.N;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= (long long)B * N) return;
    const int li = (int)(gw / N), k = (int)(gw % N);
    const int inst = inst0 + li;
    if (skip && skip[inst]) return;  // SQP: this instance has converged
    const double *Xi = X + (size_t)inst * (N + 1) * NX;
    const double *Ui = U + (size_t)inst * N * NU;
    double *wsk = ws + (size_t)li * L::instance_stride(N) + (size_t)k * L::STAGE;
    linearize_warp<NX, NU, double>(P, Xi + (size_t)k * NX, Ui + (size_t)k * NU, Xi + (size_t)(k + 1) * NX,
                                   param_ptr(p, p_mode, inst, k, N), wsk);
}

// K2: one warp per instance -- IPM/Riccati QP solve + RTI update.
// NSLOT = 2 / MINB = 1: latency variant (stage prefetch, all the registers it wants);
// NSLOT = 1 / MINB = 12: throughput variant the host scheduler picks for chunks of many waves
// (16 KB of shared memory and <= 168 registers per warp -> 12 warps per SM instead of 9).
// STRICT: reference-semantics instantiation (mpcb_config.strict_reference, see qp_solve_warp).
template <int NX, int NU, int WPB, int NSLOT, int MINB, bool STRICT>
__global__ void __launch_bounds__(32 * WPB, MINB) qp_kernel(const __grid_constant__ Params P, double *__restrict__ X,
                                                      double *__restrict__ U, const double *__restrict__ x0,
                                                      const double *__restrict__ yref, int yref_mode,
                                                      double *__restrict__ ws, double *__restrict__ u0,
                                                      int32_t *__restrict__ status, int32_t *__restrict__ iters,
                                                      int inst0, int B, const int32_t *__restrict__ skip)
{
    using L = Layout<NX, NU>;
    __shared__ __align__(16) QpSmem<NX, NU, double, NSLOT> sm_arr[WPB];
    QpSmem<NX, NU, double, NSLOT> *sm = &sm_arr[threadIdx.x >> 5];
    const int N = P.N;
    const int li = blockIdx.x * WPB + (threadIdx.x >> 5);
    if (li >= B) return;
    const int inst = inst0 + li;
    if (skip && skip[inst]) return;  // SQP: this instance has converged
    double *Xi = X + (size_t)inst * (N + 1) * NX;
    double *Ui = U + (size_t)inst * N * NU;
    const double *yr = yref;
    if (yref_mode == MPCB_PER_INSTANCE) yr = yref + (size_t)inst * (NX + NU);
    if (yref_mode == MPCB_PER_STAGE) yr = yref + (size_t)inst * (N + 1) * (NX + NU);
    int it = 0;
    const int st = qp_solve_warp<NX, NU, double, NSLOT, STRICT>(P, *sm, ws + (size_t)li * L::instance_stride(N), Xi, Ui,
                                                         x0 + (size_t)inst * NX, yr, yref_mode == MPCB_PER_STAGE, &it);
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
        if (status) status[inst] = st;
        if (iters) iters[inst] = it;
    }
    __syncwarp();
    if (u0 && lane < NU) u0[(size_t)inst * NU + lane] = Ui[lane];
}

// K2': four instances per warp, eight lanes each (mpcb_qp8.cuh): the throughput variant for chunks of many waves.
// Persistent: the grid is one wave of resident warps, whose groups draw the chunk's instances from `next`.
template <int NX, int NU>
__global__ void __launch_bounds__(32) qp8_kernel(const __grid_constant__ Params P, double *__restrict__ X, double *__restrict__ U,
                                                 const double *__restrict__ x0, const double *__restrict__ yref, int yref_mode,
                                                 double *__restrict__ ws, double *__restrict__ u0, int32_t *__restrict__ status,
                                                 int32_t *__restrict__ iters, int inst0, int B, unsigned *__restrict__ next)
{
    __shared__ Qp8Smem<NX, NU> sm;
    Qp8Batch job;
    job.X = X; job.U = U; job.x0 = x0; job.yref = yref;
    job.yref_stride = yref_mode == MPCB_PER_INSTANCE ? (size_t)(NX + NU) : yref_mode == MPCB_PER_STAGE ? (size_t)(P.N + 1) * (NX + NU) : 0;
    job.yps = yref_mode == MPCB_PER_STAGE;
    job.ws = ws; job.u0 = u0; job.status = status; job.iters = iters; job.inst0 = inst0; job.B = B; job.next = next;
    qp8_solve_queue<NX, NU>(P, sm, job);
}

template <int NX, int NU>
__global__ void plant_kernel(const __grid_constant__ Params P, const double *__restrict__ x, const double *__restrict__ u,
                             const double *__restrict__ p, int p_mode, double *__restrict__ xn, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    plant_step_thread<NX, NU, double>(P, x + (size_t)i * NX, u + (size_t)i * NU,
                                      p_mode == MPCB_SHARED ? p : p + (size_t)i * kNP, xn + (size_t)i * NX);
}

template <int NX, int NU>
__global__ void reset_kernel(const __grid_constant__ Params P, double *__restrict__ X, double *__restrict__ U,
                             const double *__restrict__ x_init, const double *__restrict__ u_init, int u_per_instance,
                             int B)
{
    const int N = P.N;
    const size_t per = (size_t)(N + 1) * NX + (size_t)N * NU;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= per * B) return;
    const int inst = (int)(tid / per);
    const size_t r = tid % per;
    if (r < (size_t)(N + 1) * NX) {
        X[(size_t)inst * (N + 1) * NX + r] = x_init ? x_init[(size_t)inst * NX + r % NX] : 0.0;
    } else {
        const size_t q = r - (size_t)(N + 1) * NX;
        U[(size_t)inst * N * NU + q] = u_init ? u_init[(u_per_instance ? (size_t)inst * NU : 0) + q % NU] : 0.0;
    }
}

// get_cost() [upstream D10]
template <int NX, int NU>
__global__ void cost_kernel(const __grid_constant__ Params P, const double *__restrict__ X, const double *__restrict__ U,
                            const double *__restrict__ yref, int yref_mode, double *__restrict__ cost, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int N = P.N, NY = NX + NU;
    const double *Xi = X + (size_t)i * (N + 1) * NX, *Ui = U + (size_t)i * N * NU;
    const double *yr = yref;
    if (yref_mode == MPCB_PER_INSTANCE) yr = yref + (size_t)i * NY;
    if (yref_mode == MPCB_PER_STAGE) yr = yref + (size_t)i * (N + 1) * NY;
    double acc = 0.0;
    for (int k = 0; k <= N; k++) {
        const double *y = yr + (yref_mode == MPCB_PER_STAGE ? (size_t)k * NY : 0);
        double s = 0.0;
        for (int j = 0; j < NX; j++) {
            const double e = Xi[(size_t)k * NX + j] - y[j];
            s += (k < N ? P.Q[j] : P.Qt[j]) * e * e;
        }
        if (k < N)
            for (int j = 0; j < NU; j++) {
                const double e = Ui[(size_t)k * NU + j] - y[NX + j];
                s += P.R[j] * e * e;
            }
        acc += (k < N ? 0.5 * P.dt : 0.5) * s;
    }
    cost[i] = acc;
}

template <int NX, int NU>
__global__ void debug_copy_kernel(const __grid_constant__ Params P, const double *__restrict__ ws, double *__restrict__ BAt,
                                  double *__restrict__ b, int inst0, int B)
{
    using L = Layout<NX, NU>;
    const int N = P.N;
    const size_t per = (size_t)N * (L::NZ * NX + NX);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= per * B) return;
    const int li = (int)(tid / per);
    const size_t r = tid % per;
    const int k = (int)(r / (L::NZ * NX + NX));
    const int e = (int)(r % (L::NZ * NX + NX));
    const double *wk = ws + (size_t)li * L::instance_stride(N) + (size_t)k * L::STAGE;
    const size_t inst = (size_t)inst0 + li;
    if (e < L::NZ * NX) BAt[(inst * N + k) * (L::NZ * NX) + e] = wk[L::O_BAT + (e / NX) * L::LDB + e % NX];
    else b[(inst * N + k) * NX + (e - L::NZ * NX)] = wk[L::O_B + e - L::NZ * NX];
}

// Test hook: the interior-point iterate the last solve ended with and the QP it solved, read out of the workspace
// records (nothing is recomputed).  One thread per (instance, stage, component); any output may be null.
struct QpDebugOut {
    double *z, *pi, *tl, *tu, *ll, *lu, *lb, *ub, *g;  // [B,N+1,nz] each, pi [B,N+1,nx]
};
template <int NX, int NU>
__global__ void debug_qp_kernel(const __grid_constant__ Params P, const double *__restrict__ ws, QpDebugOut o, int inst0, int B)
{
    using L = Layout<NX, NU>;
    const int N = P.N;
    const size_t per = (size_t)(N + 1) * L::NZ;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= per * B) return;
    const int li = (int)(tid / per);
    const int k = (int)((tid % per) / L::NZ), j = (int)(tid % L::NZ);
    const double *wk = ws + (size_t)li * L::instance_stride(N) + (size_t)k * L::STAGE;
    const size_t inst = (size_t)inst0 + li;
    const size_t oz = (inst * (N + 1) + k) * L::NZ + j;
    const VarKind vk = var_kind<NX, NU>(k, j, N);
    if (o.z) o.z[oz] = (vk.var || (k == 0 && j >= NU)) ? wk[L::O_Z + j] : 0.0;
    if (o.tl) o.tl[oz] = vk.hasb ? wk[L::O_TL + j] : 0.0;
    if (o.tu) o.tu[oz] = vk.hasb ? wk[L::O_TU + j] : 0.0;
    if (o.ll) o.ll[oz] = vk.hasb ? wk[L::O_LL + j] : 0.0;
    if (o.lu) o.lu[oz] = vk.hasb ? wk[L::O_LUP + j] : 0.0;
    if (o.lb) o.lb[oz] = vk.hasb ? wk[L::O_LB + j] : -HUGE_VAL;
    if (o.ub) o.ub[oz] = vk.hasb ? wk[L::O_UB + j] : HUGE_VAL;
    if (o.g) o.g[oz] = (vk.var || (k == 0 && j >= NU)) ? wk[L::O_G + j] : 0.0;
    if (o.pi && j < NX) o.pi[(inst * (N + 1) + k) * NX + j] = k >= 1 ? wk[L::O_PI + j] : 0.0;
}

// ---- SQP to convergence (reference options nlp_solver_max_iter / nlp_solver_tol_*,
// acados_ocp_blasterModel.json solver_options).  Per-instance state of one mpcb_solve_sqp call:
struct SqpState {
    double *pi, *ll, *lu;  // multipliers of the last QP: pi[B,N+1,nx], ll / lu[B,N+1,nz]
    int32_t *done;         // 1 once the instance has converged or failed; its iterate is then left alone
    int32_t *status;       // SQP status: 0 converged, 2 iteration cap, 4 QP failure, 1 NaN
    int32_t *sqp_iters;    // QP solves performed
    int32_t *qp_iters;     // interior-point iterations over all of them
    double *res;           // [B,4] last evaluated residuals (stat, eq, ineq, comp); may be null
    double tol[4];
};

// NLP residuals at the stored iterate with the multipliers of the last QP, one warp per instance, right after the
// rollout kernel has re-linearised (BAt, b in the records are those of this iterate): inf-norms of the Lagrangian
// gradient (Gauss-Newton LINEAR_LS cost, x_0 pinned), the dynamics defect (and x_0 - X_0), the bound violation and the
// complementarity products.  Converged instances are flagged done (and u0 is written for them).
template <int NX, int NU>
__global__ void __launch_bounds__(128) nlp_res_kernel(const __grid_constant__ Params P, const double *__restrict__ X,
                                                      const double *__restrict__ U, const double *__restrict__ x0,
                                                      const double *__restrict__ yref, int yref_mode, const double *__restrict__ ws,
                                                      SqpState st, int sqp_it, double *__restrict__ u0, int inst0, int B)
{
    using L = Layout<NX, NU>;
    constexpr int NZ = L::NZ;
    const int N = P.N;
    const int li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (li >= B) return;
    const int inst = inst0 + li, lane = threadIdx.x & 31;
    if (st.done[inst]) return;
    const double *Xi = X + (size_t)inst * (N + 1) * NX, *Ui = U + (size_t)inst * N * NU;
    const double *yr0 = yref;
    if (yref_mode == MPCB_PER_INSTANCE) yr0 = yref + (size_t)inst * NZ;
    if (yref_mode == MPCB_PER_STAGE) yr0 = yref + (size_t)inst * (N + 1) * NZ;
    const double *pi = st.pi + (size_t)inst * (N + 1) * NX;
    const double *ll = st.ll + (size_t)inst * (N + 1) * NZ, *lu = st.lu + (size_t)inst * (N + 1) * NZ;
    double rs = 0.0, re = 0.0, ri = 0.0, rc = 0.0;
    for (int k = 0; k <= N; k++) {
        const double *wk = ws + (size_t)li * L::instance_stride(N) + (size_t)k * L::STAGE;
        const double *yr = yr0 + (yref_mode == MPCB_PER_STAGE ? (size_t)k * NZ : 0);
        const VarKind vk = var_kind<NX, NU>(k, lane, N);
        if (lane < NZ) {
            const double H0 = hess_diag<NX, NU, double>(P, k, lane);
            double y = 0.0, yv = 0.0, lb = 0.0, ub = 0.0;
            if (lane < NU) {
                if (k < N) { y = Ui[(size_t)k * NU + lane]; yv = yr[NX + lane]; lb = P.lbu[lane]; ub = P.ubu[lane]; }
            } else {
                y = Xi[(size_t)k * NX + lane - NU]; yv = yr[lane - NU]; lb = P.lbx[lane - NU]; ub = P.ubx[lane - NU];
            }
            if (vk.var) {
                double r = H0 * (y - yv);
                if (k < N)
                    for (int c = 0; c < NX; c++) r += wk[L::O_BAT + lane * L::LDB + c] * pi[(size_t)(k + 1) * NX + c];
                if (lane >= NU) r -= pi[(size_t)k * NX + lane - NU];
                if (vk.hasb) r += lu[(size_t)k * NZ + lane] - ll[(size_t)k * NZ + lane];
                rs = fmax(rs, fabs(r));
            }
            if (vk.hasb) {
                ri = fmax(ri, fmax(lb - y, y - ub));
                rc = fmax(rc, fmax(fabs(ll[(size_t)k * NZ + lane] * (y - lb)), fabs(lu[(size_t)k * NZ + lane] * (ub - y))));
            }
        }
        if (lane < NX) {
            if (k < N) re = fmax(re, fabs(wk[L::O_B + lane]));
            if (k == 0) re = fmax(re, fabs(x0[(size_t)inst * NX + lane] - Xi[lane]));
        }
    }
    rs = warp_max(rs); re = warp_max(re); ri = warp_max(ri); rc = warp_max(rc);
    const bool nan = !(rs == rs) || !(re == re);
    const bool conv = !nan && rs <= st.tol[0] && re <= st.tol[1] && ri <= st.tol[2] && rc <= st.tol[3];
    if (lane == 0) {
        if (st.res) { double *o = st.res + (size_t)inst * 4; o[0] = rs; o[1] = re; o[2] = ri; o[3] = rc; }
        if (conv) { st.done[inst] = 1; st.status[inst] = ST_OK; }
        else if (nan) { st.done[inst] = 1; st.status[inst] = ST_NAN; }
    }
    if ((conv || nan) && u0 && lane < NU) u0[(size_t)inst * NU + lane] = Ui[lane];
}

// After the QP kernel of an SQP iteration: keep the multipliers for the next residual evaluation and account for the
// solve.  A QP that failed leaves the iterate untouched (or, with strict_reference, only the iteration cap still applies
// the step), so repeating it would change nothing: the instance ends with the QP-failure status, as acados' SQP does.
template <int NX, int NU>
__global__ void sqp_book_kernel(const __grid_constant__ Params P, const double *__restrict__ ws, SqpState st,
                                const int32_t *__restrict__ qp_status, const int32_t *__restrict__ qp_iters, int inst0, int B)
{
    using L = Layout<NX, NU>;
    const int N = P.N;
    const size_t per = (size_t)(N + 1) * L::NZ;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= per * B) return;
    const int li = (int)(tid / per);
    const int k = (int)((tid % per) / L::NZ), j = (int)(tid % L::NZ);
    const size_t inst = (size_t)inst0 + li;
    if (st.done[inst]) return;  // flagged before this QP (the failure flag below is set by this kernel: racing readers only skip a dead copy)
    const double *wk = ws + (size_t)li * L::instance_stride(N) + (size_t)k * L::STAGE;
    const VarKind vk = var_kind<NX, NU>(k, j, N);
    st.ll[(inst * (N + 1) + k) * L::NZ + j] = vk.hasb ? wk[L::O_LL + j] : 0.0;
    st.lu[(inst * (N + 1) + k) * L::NZ + j] = vk.hasb ? wk[L::O_LUP + j] : 0.0;
    if (j < NX) st.pi[(inst * (N + 1) + k) * NX + j] = k >= 1 ? wk[L::O_PI + j] : 0.0;
    if (k == 0 && j == 0) {
        const int q = qp_status[inst];
        st.sqp_iters[inst] += 1;
        st.qp_iters[inst] += qp_iters[inst];
        const bool applied = q == ST_OK || (P.strict && q == ST_MAXITER);
        if (!applied) { st.status[inst] = q == ST_NAN ? ST_NAN : ST_QPFAIL; __threadfence(); st.done[inst] = 1; }
    }
}

__global__ void fill_i32_kernel(int32_t *a, int32_t v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

// Horizon shift of the stored iterate (the usual RTI companion; the reference's
// scripts never shift): X_k <- X_{k+1}, U_k <- U_{k+1}, the last stage is repeated.
template <int NX, int NU>
__global__ void shift_kernel(const __grid_constant__ Params P, double *__restrict__ X, double *__restrict__ U, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int N = P.N;
    double *Xi = X + (size_t)i * (N + 1) * NX, *Ui = U + (size_t)i * N * NU;
    for (int k = 0; k < N; k++)
        for (int j = 0; j < NX; j++) Xi[(size_t)k * NX + j] = Xi[(size_t)(k + 1) * NX + j];
    for (int k = 0; k + 1 < N; k++)
        for (int j = 0; j < NU; j++) Ui[(size_t)k * NU + j] = Ui[(size_t)(k + 1) * NU + j];
}

// closed-loop bookkeeping: x <- xnext, count failures / iterations
template <int NX>
__global__ void loop_book_kernel(double *__restrict__ x, const double *__restrict__ xn, const int32_t *__restrict__ status,
                                 const int32_t *__restrict__ iters, int32_t *__restrict__ n_fail,
                                 int32_t *__restrict__ iters_sum, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    for (int j = 0; j < NX; j++) x[(size_t)i * NX + j] = xn[(size_t)i * NX + j];
    if (n_fail && status[i] != 0) n_fail[i] += 1;
    if (iters_sum) iters_sum[i] += iters[i];
}

// reference mavros_blaster_sim.py:27-30,91-100
__global__ void command_map_kernel(const double *__restrict__ x, const double *__restrict__ u0, int nx, int nu,
                                   double *__restrict__ quat, double *__restrict__ thrust, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    if (quat) euler_to_quat<double>(x[(size_t)i * nx + 3], x[(size_t)i * nx + 4], x[(size_t)i * nx + 5], quat + (size_t)i * 4);
    if (thrust) {
        const double *u = u0 + (size_t)i * nu;
        const double avg = 2.3 * (0.25 * (u[0] + u[1] + u[2] + u[3])) / 9.81;
        thrust[i] = 0.0014 * avg * avg * avg - 0.0263 * avg * avg + 0.2464 * avg - 0.0286;
    }
}

// Jet point of contact and its Jacobians, one vehicle pose per thread (mpcb_poc.cuh;
// reference Jacobian_POC_Solver.py:234-300).  x17 != nullptr: poses are read from state vectors
// (position x[0:3], Euler angles x[3:6], nozzle angles x[12:14], blastermodel.py:171-190).
__global__ void poc_kernel(const double *__restrict__ euler, const double *__restrict__ motor, const double *__restrict__ position,
                           const double *__restrict__ x17, int B, double V, double drag, int mode, double T_blast,
                           double *__restrict__ poc, double *__restrict__ J_mot, double *__restrict__ J_eul,
                           double *__restrict__ J_pos, double *__restrict__ p25, double *__restrict__ t_flight, int32_t *__restrict__ status)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double e[3], m[2], pos[3];
    if (x17) {
        const double *x = x17 + (size_t)i * 17;
        for (int c = 0; c < 3; c++) { pos[c] = x[c]; e[c] = x[3 + c]; }
        m[0] = x[12]; m[1] = x[13];
    } else {
        for (int c = 0; c < 3; c++) { pos[c] = position[(size_t)i * 3 + c]; e[c] = euler[(size_t)i * 3 + c]; }
        m[0] = motor[(size_t)i * 2]; m[1] = motor[(size_t)i * 2 + 1];
    }
    PocOut o;
    if (mode == POC_MODE_ANALYTIC) poc_analytic(e, m, pos, V, drag, o);
    else poc_reference(e, m, pos, V, drag, o);
    for (int r = 0; r < 3; r++) {
        if (poc) poc[(size_t)i * 3 + r] = o.poc[r];
        if (J_mot) for (int c = 0; c < 2; c++) J_mot[(size_t)i * 6 + r * 2 + c] = o.J[r][c];
        if (J_eul) for (int c = 0; c < 3; c++) J_eul[(size_t)i * 9 + r * 3 + c] = o.J[r][2 + c];
        if (J_pos) for (int c = 0; c < 3; c++) J_pos[(size_t)i * 9 + r * 3 + c] = o.J[r][5 + c];
    }
    if (p25) poc_pack_params(o, T_blast, p25 + (size_t)i * kNP);
    if (t_flight) t_flight[i] = o.t_flight;
    if (status) status[i] = o.status;
}

// FP64 FMA-pipe micro-benchmark: 8 independent DFMA chains per thread.
__global__ void fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace

// ------------------------------------------------------------------ handle
struct mpcb_handle {
    mpcb_config cfg;
    Params P;
    int nx, nu, N, device;
    int max_batch, ws_batch;
    size_t ws_stride;  // doubles per instance
    double *X = nullptr, *U = nullptr, *ws = nullptr;
    double *p_default = nullptr;  // device copy of the default parameter vector
    int32_t *status_scratch = nullptr, *iters_scratch = nullptr;
    double *xn_scratch = nullptr, *u0_scratch = nullptr;
    // host staging for mpcb_solve_host
    double *h_pin = nullptr;
    size_t h_pin_bytes = 0;
    double *d_stage = nullptr;
    size_t d_stage_bytes = 0;
    cudaStream_t own_stream = nullptr;
    int throughput_batch = 1 << 30;  // chunks at least this large use the high-occupancy QP kernel variant
    int qp8_batch = 1 << 30;         // chunks at least this large use the four-instances-per-warp kernel
    int qp8_resident = 1;            // warps of that kernel one wave holds (SMs x resident CTAs per SM)
    unsigned *qp8_next = nullptr;    // its work counter
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};  // around K1 and K2 of the last solve (profiling)
    bool profile = false;
    // state of mpcb_solve_sqp (allocated on first use): multipliers of the last QP and per-instance flags
    double *sqp_pi = nullptr, *sqp_ll = nullptr, *sqp_lu = nullptr;
    int32_t *sqp_done = nullptr, *sqp_status = nullptr, *sqp_count = nullptr, *sqp_qpit = nullptr;
    int last_solve_batch = 0;  // instances whose records the workspace still holds (mpcb_debug_qp)
    std::string err;
};

// Entry points that must run on the handle's GPU make it current for their duration and restore the caller's device
// on every return path (a process driving several GPUs keeps its own current device).
struct DeviceGuard {
    int prev = -1, dev;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int d) : dev(d)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};

namespace {

int fail(mpcb_handle *h, const char *what, cudaError_t e = cudaSuccess)
{
    std::string m = what;
    if (e != cudaSuccess) { m += ": "; m += cudaGetErrorString(e); }
    if (h) h->err = m; else g_create_error = m;
    return -1;
}

#define CK(h, call)                                             \
    do {                                                        \
        cudaError_t e_ = (call);                                \
        if (e_ != cudaSuccess) return fail((h), #call, e_);     \
    } while (0)

void invert3(const double *J, double *Ji)
{
    const double a = J[0], b = J[1], c = J[2], d = J[3], e = J[4], f = J[5], g = J[6], hh = J[7], i = J[8];
    const double det = a * (e * i - f * hh) - b * (d * i - f * g) + c * (d * hh - e * g);
    const double r = 1.0 / det;
    Ji[0] = (e * i - f * hh) * r; Ji[1] = (c * hh - b * i) * r; Ji[2] = (b * f - c * e) * r;
    Ji[3] = (f * g - d * i) * r;  Ji[4] = (a * i - c * g) * r;  Ji[5] = (c * d - a * f) * r;
    Ji[6] = (d * hh - e * g) * r; Ji[7] = (b * g - a * hh) * r; Ji[8] = (a * e - b * d) * r;
}

Params make_params(const mpcb_config &c)
{
    Params P;
    memset(&P, 0, sizeof(P));
    P.variant = c.variant; P.N = c.N; P.dt = c.dt; P.mass = c.mass; P.inv_mass = 1.0 / c.mass;
    memcpy(P.J, c.J, sizeof(P.J));
    invert3(c.J, P.Jinv);
    P.l_x = c.l_x; P.l_y = c.l_y; P.c = c.c;
    // moment map, reference blastermodel.py:95-101
    const double G[3][4] = {{-c.l_y, c.l_y, -c.l_y, c.l_y}, {-c.l_x, c.l_x, c.l_x, -c.l_x}, {-c.c, -c.c, c.c, c.c}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) {
            double a = 0;
            for (int k = 0; k < 3; k++) a += P.Jinv[3 * i + k] * G[k][j];
            P.JinvG[4 * i + j] = a;
        }
    memcpy(P.Q, c.Q, sizeof(P.Q)); memcpy(P.R, c.R, sizeof(P.R)); memcpy(P.Qt, c.Qt, sizeof(P.Qt));
    memcpy(P.lbx, c.lbx, sizeof(P.lbx)); memcpy(P.ubx, c.ubx, sizeof(P.ubx));
    memcpy(P.lbu, c.lbu, sizeof(P.lbu)); memcpy(P.ubu, c.ubu, sizeof(P.ubu));
    P.ipm_max_iter = c.ipm_max_iter; P.ipm_mu0 = c.ipm_mu0; P.ipm_thr0 = c.ipm_thr0;
    P.tol_stat = c.tol_stat; P.tol_eq = c.tol_eq; P.tol_ineq = c.tol_ineq; P.tol_comp = c.tol_comp;
    P.alpha_min = c.alpha_min;
    P.strict = c.strict_reference ? 1 : 0;
    return P;
}

constexpr int kWPB = 1;  // warps (= instances) per CTA of the QP kernel
#ifndef MPCB_TP_MINB
#define MPCB_TP_MINB 12  // resident CTAs per SM the throughput variant of the QP kernel is compiled for
#endif

template <int NX, int NU>
int launch_solve_chunks(mpcb_handle *h, const double *x0, const double *yref, int yref_mode, const double *p, int p_mode,
                        double *u0, int32_t *status, int32_t *iters, int B, cudaStream_t s, const SqpState *sqp = nullptr,
                        int sqp_it = 0, bool sqp_eval_only = false)
{
    // SQP to convergence tests the Lagrangian gradient with the QP's multipliers at 1e-6: that needs the refined solve of
    // the reference-semantics instantiation (the default rule set leaves up to ~1e-3 in the multipliers of active bounds)
    Params P = h->P;
    if (sqp) P.strict = 1;
    static_assert(sizeof(QpSmem<NX, NU, double, 2>) * kWPB <= 48 * 1024, "static shared memory limit");
    using L = Layout<NX, NU>;
    const size_t smem = 0;
    const int32_t *skip = sqp ? sqp->done : nullptr;
    for (int i0 = 0; i0 < B; i0 += h->ws_batch) {
        const int nb = (B - i0 < h->ws_batch) ? B - i0 : h->ws_batch;
        const long long warps = (long long)nb * h->N;
        const unsigned g1 = (unsigned)((warps * 32 + 127) / 128);
        const bool prof = h->profile && i0 == 0;
        if (prof) cudaEventRecord(h->ev[0], s);
        linearize_kernel<NX, NU><<<g1, 128, 0, s>>>(P, h->X, h->U, p, p_mode, h->ws, i0, nb, skip);
        g_launches += 1;
        if (sqp) {
            nlp_res_kernel<NX, NU><<<(unsigned)(((long long)nb * 32 + 127) / 128), 128, 0, s>>>(P, h->X, h->U, x0, yref, yref_mode, h->ws, *sqp,
                                                                                             sqp_it, u0, i0, nb);
            g_launches += 1;
            if (sqp_eval_only) continue;
        }
        if (prof) cudaEventRecord(h->ev[1], s);
        if (P.strict)
            // reference semantics (explicit residual norms, no divergence exit): the one-instance latency kernel, any batch
            qp_kernel<NX, NU, kWPB, 2, 1, true><<<(nb + kWPB - 1) / kWPB, 32 * kWPB, smem, s>>>(P, h->X, h->U, x0, yref, yref_mode,
                                                                                                  h->ws, u0, status, iters, i0, nb, skip);
        else if (nb >= h->qp8_batch) {
            // one wave of resident warps; their groups draw the chunk's instances from the work counter
            const int want = (nb + kGPW - 1) / kGPW;
            const int grid = want < h->qp8_resident ? want : h->qp8_resident;
            CK(h, cudaMemsetAsync(h->qp8_next, 0, sizeof(unsigned), s));
            qp8_kernel<NX, NU><<<grid, 32, 0, s>>>(P, h->X, h->U, x0, yref, yref_mode, h->ws, u0, status, iters, i0, nb, h->qp8_next);
        }
        else if (nb >= h->throughput_batch)
            qp_kernel<NX, NU, kWPB, 1, MPCB_TP_MINB, false><<<(nb + kWPB - 1) / kWPB, 32 * kWPB, smem, s>>>(P, h->X, h->U, x0, yref, yref_mode,
                                                                                                    h->ws, u0, status, iters, i0, nb, skip);
        else
            qp_kernel<NX, NU, kWPB, 2, 1, false><<<(nb + kWPB - 1) / kWPB, 32 * kWPB, smem, s>>>(P, h->X, h->U, x0, yref, yref_mode,
                                                                                                   h->ws, u0, status, iters, i0, nb, skip);
        if (prof) cudaEventRecord(h->ev[2], s);
        g_launches += 1;
        if (sqp) {
            const size_t per = (size_t)(h->N + 1) * L::NZ;
            sqp_book_kernel<NX, NU><<<(unsigned)((per * nb + 255) / 256), 256, 0, s>>>(P, h->ws, *sqp, status, iters, i0, nb);
            g_launches += 1;
        }
    }
    h->last_solve_batch = B <= h->ws_batch ? B : 0;
    CK(h, cudaGetLastError());
    return 0;
}

int check_batch(mpcb_handle *h, int B)
{
    if (!h) return fail(nullptr, "null handle");
    if (B < 0 || B > h->max_batch) return fail(h, "batch size exceeds max_batch");
    return 0;
}

}  // namespace

// ------------------------------------------------------------------ C ABI
extern "C" {

int mpcb_config_default(mpcb_config *cfg, int variant, int N)
{
    if (!cfg || (variant != 17 && variant != 12 && variant != 13) || N < 2) return -1;
    memset(cfg, 0, sizeof(*cfg));
    cfg->variant = variant; cfg->N = N; cfg->dt = 2.0 / 60.0;
    cfg->mass = 9.0;
    cfg->J[0] = 0.50781; cfg->J[4] = 0.47314; cfg->J[8] = 0.72975;
    cfg->l_x = 0.3434; cfg->l_y = 0.3475; cfg->c = 0.03;
    const double Q[17] = {1e3, 1e3, 1e3, 1e3, 1e3, 1e3, 5, 5, 5, 10, 10, 10, 1e-2, 1e-2, 1e3, 1e3, 1e3};
    const double R[6] = {5e-2, 5e-2, 5e-2, 5e-2, 1e-5, 1e-5};
    const double lbx[17] = {-1.5, -1.5, 0, -0.174532925, -0.174532925, -0.349066, -1.0, -1.0, -1.0, -0.0872665, -0.0872665,
                            -0.0872665, -0.174532925, -0.523599, -1.5, -1.5, -2.5};
    const double ubx[17] = {1.5, 1.5, 5.0, 0.174532925, 0.174532925, 0.349066, 1.0, 1.0, 1.0, 0.0872665, 0.0872665,
                            0.0872665, 1.22173, 0.523599, 1.5, 1.5, 2.5};
    const double lbu[6] = {0, 0, 0, 0, -0.0872665, -0.0872665}, ubu[6] = {65, 65, 65, 65, 0.0872665, 0.0872665};
    for (int i = 0; i < 17; i++) { cfg->Q[i] = Q[i]; cfg->Qt[i] = 10 * Q[i]; cfg->lbx[i] = lbx[i]; cfg->ubx[i] = ubx[i]; }
    for (int i = 0; i < 6; i++) { cfg->R[i] = R[i]; cfg->lbu[i] = lbu[i]; cfg->ubu[i] = ubu[i]; }
    cfg->ipm_max_iter = 60; cfg->ipm_mu0 = 1e2; cfg->ipm_thr0 = -0.5;
    cfg->tol_stat = 1e-6; cfg->tol_eq = 1e-8; cfg->tol_ineq = 1e-8; cfg->tol_comp = 1e-8; cfg->alpha_min = 1e-8;
    cfg->dtype = MPCB_F64; cfg->max_batch = 1024; cfg->ws_batch = 0; cfg->device = -1;
    cfg->strict_reference = 0; cfg->throughput_batch = 0; cfg->qp8_batch = 0; cfg->qp8_warps = 0;
    if (variant == 13) {
        // QUAT13: x = [p, q(w,x,y,z), v, omega].  The Euler weights go to the quaternion components; the Euler boxes
        // (10, 10, 20 deg) become boxes on the vector part (sine of half the angle), q_w stays near 1.
        const double hq[3] = {sin(0.174532925 / 2), sin(0.174532925 / 2), sin(0.349066 / 2)};
        double Q13[13], lb13[13], ub13[13];
        for (int i = 0; i < 3; i++) { Q13[i] = Q[i]; lb13[i] = lbx[i]; ub13[i] = ubx[i]; }
        Q13[3] = 1e3; lb13[3] = 0.9; ub13[3] = 1.05;
        for (int i = 0; i < 3; i++) { Q13[4 + i] = 1e3; lb13[4 + i] = -hq[i]; ub13[4 + i] = hq[i]; }
        for (int i = 0; i < 6; i++) { Q13[7 + i] = Q[6 + i]; lb13[7 + i] = lbx[6 + i]; ub13[7 + i] = ubx[6 + i]; }
        for (int i = 0; i < 17; i++) { cfg->Q[i] = cfg->Qt[i] = cfg->lbx[i] = cfg->ubx[i] = 0.0; }
        for (int i = 0; i < 13; i++) { cfg->Q[i] = Q13[i]; cfg->Qt[i] = 10 * Q13[i]; cfg->lbx[i] = lb13[i]; cfg->ubx[i] = ub13[i]; }
    }
    return 0;
}

int mpcb_create(const mpcb_config *cfg, mpcb_handle **out)
{
    if (!cfg || !out) return fail(nullptr, "null argument");
    if (cfg->variant != 17 && cfg->variant != 12 && cfg->variant != 13) return fail(nullptr, "variant must be 17, 12 or 13");
    if (cfg->N < 2 || cfg->N > 4096) return fail(nullptr, "horizon out of range");
    if (cfg->dtype != MPCB_F64)
        return fail(nullptr, "only dtype = MPCB_F64 is implemented (an interior point with active state bounds is not viable in FP32, DESIGN.md)");
    if (cfg->max_batch < 1) return fail(nullptr, "max_batch must be >= 1");
    if (!(cfg->dt > 0) || !(cfg->mass > 0)) return fail(nullptr, "dt and mass must be positive");
    {
        // an interior point needs boxes with an interior: lb < ub strictly (a zero-width box would start with a zero
        // slack and an infinite multiplier -> NaN status on every instance).  Pin a variable with a tiny box instead.
        const int nxv = cfg->variant == 17 ? 17 : cfg->variant == 13 ? 13 : 12, nuv = cfg->variant == 17 ? 6 : 4;
        for (int i = 0; i < nxv; i++)
            if (!(cfg->lbx[i] < cfg->ubx[i])) return fail(nullptr, "state bounds must satisfy lbx < ubx strictly (zero-width or inverted box)");
        for (int i = 0; i < nuv; i++)
            if (!(cfg->lbu[i] < cfg->ubu[i])) return fail(nullptr, "input bounds must satisfy lbu < ubu strictly (zero-width or inverted box)");
        for (int i = 0; i < nxv; i++)
            if (!(cfg->Q[i] >= 0) || !(cfg->Qt[i] >= 0)) return fail(nullptr, "weights must be non-negative");
        for (int i = 0; i < nuv; i++)
            if (!(cfg->R[i] > 0)) return fail(nullptr, "input weights R must be positive (strict convexity of the QP)");
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(nullptr, "no CUDA device: this library has no CPU path", e);
    mpcb_handle *h = new (std::nothrow) mpcb_handle();
    if (!h) return fail(nullptr, "out of host memory");
    h->cfg = *cfg;
    h->P = make_params(*cfg);
    h->nx = cfg->variant == 17 ? 17 : cfg->variant == 13 ? 13 : 12;
    h->nu = cfg->variant == 17 ? 6 : 4;
    h->N = cfg->N;
    h->max_batch = cfg->max_batch;
    if (cfg->device >= 0) {
        h->device = cfg->device;
    } else {
        cudaGetDevice(&h->device);
    }
    DeviceGuard guard(h->device);  // the caller's current device is restored on every return path
    if (guard.err != cudaSuccess) { fail(nullptr, "cudaSetDevice", guard.err); delete h; return -1; }
    h->ws_stride = cfg->variant == 17 ? Layout<17, 6>::instance_stride(h->N)
                   : cfg->variant == 13 ? Layout<13, 4>::instance_stride(h->N) : Layout<12, 4>::instance_stride(h->N);
    int wsb = cfg->ws_batch;
    if (wsb <= 0) {
        // auto: at most ~24 GiB of solver workspace resident at once
        const size_t budget = (size_t)24 << 30;
        size_t cap = budget / (h->ws_stride * sizeof(double));
        if (cap < 1) cap = 1;
        wsb = (size_t)h->max_batch < cap ? h->max_batch : (int)cap;
    }
    if (wsb > h->max_batch) wsb = h->max_batch;
    h->ws_batch = wsb;
    {
        // Kernel selection by chunk size (mpcb_config fields; 0 = the measured defaults, profiles/r02_crossover.txt):
        //  * throughput_batch: the single-buffer variant (12 warps per SM) used to win from 4,096 instances on; since the
        //    latency variant carries P_k in its normal-equations iterations (round 2) it no longer does at any size
        //    measured (BLASTER17: 327 k against 378 k solves/s at 4,096, 346 k against 376 k at 6,144, 371 k against
        //    384 k at 16,384) -- off by default, the field still forces it;
        //  * qp8_batch: the four-instances-per-warp kernel (persistent, groups refilled from a work counter, next
        //    stage's record prefetched into L2) is faster from two of its waves (148 SMs x 7 warps x 4 = 4,144
        //    instances each) on for BLASTER17 (388 k against 379 k at 8,192, 421 k against 384 k at 16,384) and from
        //    4,096 instances on for QUAD12 (704 k against 620 k; at 3,072 the latency variant leads 616 k to 586 k).
        h->throughput_batch = cfg->throughput_batch > 0 ? cfg->throughput_batch : (1 << 30);
        h->qp8_batch = cfg->qp8_batch > 0 ? cfg->qp8_batch : (cfg->variant == 17 ? 8192 : 4096);
    }
    const size_t B = (size_t)h->max_batch;
    const size_t nX = B * (h->N + 1) * h->nx, nU = B * h->N * h->nu;
#define ALLOC(ptr, bytes)                                                                   \
    do {                                                                                    \
        e = cudaMalloc((void **)&(ptr), (bytes));                                           \
        if (e != cudaSuccess) { fail(nullptr, "cudaMalloc " #ptr, e); mpcb_destroy(h); return -1; } \
    } while (0)
    ALLOC(h->X, nX * sizeof(double));
    ALLOC(h->U, nU * sizeof(double));
    ALLOC(h->ws, (size_t)h->ws_batch * h->ws_stride * sizeof(double));
    ALLOC(h->p_default, kNP * sizeof(double));
    ALLOC(h->status_scratch, B * sizeof(int32_t));
    ALLOC(h->iters_scratch, B * sizeof(int32_t));
    ALLOC(h->xn_scratch, B * h->nx * sizeof(double));
    ALLOC(h->u0_scratch, B * h->nu * sizeof(double));
    ALLOC(h->qp8_next, sizeof(unsigned));
#undef ALLOC
    {
        // one wave of the persistent four-instances-per-warp kernel
        int per_sm = 0, sms = 0;
        if (cfg->variant == 17) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qp8_kernel<17, 6>, 32, 0);
        else if (cfg->variant == 13) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qp8_kernel<13, 4>, 32, 0);
        else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qp8_kernel<12, 4>, 32, 0);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        if (e != cudaSuccess || per_sm < 1 || sms < 1) { fail(nullptr, "occupancy query of qp8_kernel", e); mpcb_destroy(h); return -1; }
        h->qp8_resident = per_sm * sms;
        // test hook: a smaller grid makes the groups of a warp go through many instances
        if (cfg->qp8_warps >= 1 && cfg->qp8_warps < h->qp8_resident) h->qp8_resident = cfg->qp8_warps;
    }
    double pd[kNP] = {0};
    pd[24] = 2.2 * 9.81;  // reference blastermodel.py:280-282
    cudaMemcpy(h->p_default, pd, sizeof(pd), cudaMemcpyHostToDevice);
    cudaMemset(h->X, 0, nX * sizeof(double));
    cudaMemset(h->U, 0, nU * sizeof(double));
    cudaMemset(h->ws, 0, (size_t)h->ws_batch * h->ws_stride * sizeof(double));
    e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { fail(nullptr, "cudaStreamCreate", e); mpcb_destroy(h); return -1; }
    for (int i = 0; i < 3; i++) {
        e = cudaEventCreate(&h->ev[i]);
        if (e != cudaSuccess) { fail(nullptr, "cudaEventCreate", e); mpcb_destroy(h); return -1; }
    }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fail(nullptr, "init", e); mpcb_destroy(h); return -1; }
    *out = h;
    return 0;
}

int mpcb_destroy(mpcb_handle *h)
{
    if (!h) return 0;
    DeviceGuard guard(h->device);
    cudaFree(h->sqp_pi); cudaFree(h->sqp_ll); cudaFree(h->sqp_lu);
    cudaFree(h->sqp_done); cudaFree(h->sqp_status); cudaFree(h->sqp_count); cudaFree(h->sqp_qpit);
    cudaFree(h->X); cudaFree(h->U); cudaFree(h->ws); cudaFree(h->p_default);
    cudaFree(h->status_scratch); cudaFree(h->iters_scratch); cudaFree(h->xn_scratch); cudaFree(h->u0_scratch);
    cudaFree(h->d_stage); cudaFree(h->qp8_next);
    if (h->h_pin) cudaFreeHost(h->h_pin);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    for (int i = 0; i < 3; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    delete h;
    return 0;
}

const char *mpcb_last_error(const mpcb_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }
int mpcb_nx(const mpcb_handle *h) { return h ? h->nx : -1; }
int mpcb_nu(const mpcb_handle *h) { return h ? h->nu : -1; }
int mpcb_horizon(const mpcb_handle *h) { return h ? h->N : -1; }
int64_t mpcb_kernel_launches(void) { return g_launches.load(); }

int mpcb_reset(mpcb_handle *h, const double *x_init, const double *u_init, int u_per_instance, int B, void *stream)
{
    if (check_batch(h, B)) return -1;
    if (B == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t per = (size_t)(h->N + 1) * h->nx + (size_t)h->N * h->nu;
    const unsigned grid = (unsigned)((per * B + 255) / 256);
    if (h->nx == 17) reset_kernel<17, 6><<<grid, 256, 0, s>>>(h->P, h->X, h->U, x_init, u_init, u_per_instance, B);
    else if (h->nx == 13) reset_kernel<13, 4><<<grid, 256, 0, s>>>(h->P, h->X, h->U, x_init, u_init, u_per_instance, B);
    else reset_kernel<12, 4><<<grid, 256, 0, s>>>(h->P, h->X, h->U, x_init, u_init, u_per_instance, B);
    g_launches += 1;
    CK(h, cudaGetLastError());
    return 0;
}

int mpcb_solve(mpcb_handle *h, const double *x0, const double *yref, int yref_mode, const double *p, int p_mode,
               double *u0, double *X, double *U, int32_t *status, int32_t *iters, int B, void *stream)
{
    if (check_batch(h, B)) return -1;
    if (!x0 || !yref) return fail(h, "x0 and yref are required");
    if (yref_mode < 0 || yref_mode > 2 || p_mode < 0 || p_mode > 2) return fail(h, "bad yref_mode / p_mode");
    if (B == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (!p) { p = h->p_default; p_mode = MPCB_SHARED; }
    int rc = (h->nx == 17)   ? launch_solve_chunks<17, 6>(h, x0, yref, yref_mode, p, p_mode, u0, status, iters, B, s)
             : (h->nx == 13) ? launch_solve_chunks<13, 4>(h, x0, yref, yref_mode, p, p_mode, u0, status, iters, B, s)
                             : launch_solve_chunks<12, 4>(h, x0, yref, yref_mode, p, p_mode, u0, status, iters, B, s);
    if (rc) return rc;
    if (X) CK(h, cudaMemcpyAsync(X, h->X, (size_t)B * (h->N + 1) * h->nx * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (U) CK(h, cudaMemcpyAsync(U, h->U, (size_t)B * h->N * h->nu * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return 0;
}

int mpcb_solve_host(mpcb_handle *h, const double *x0, const double *yref, int yref_mode, const double *p, int p_mode,
                    double *u0, double *X, double *U, int32_t *status, int32_t *iters, int B)
{
    if (check_batch(h, B)) return -1;
    if (!x0 || !yref) return fail(h, "x0 and yref are required");
    if (yref_mode < 0 || yref_mode > 2 || p_mode < 0 || p_mode > 2) return fail(h, "bad yref_mode / p_mode");
    if (B == 0) return 0;
    DeviceGuard guard(h->device);
    if (guard.err != cudaSuccess) return fail(h, "cudaSetDevice", guard.err);
    const int nx = h->nx, nu = h->nu, N = h->N, ny = nx + nu;
    const size_t n_x0 = (size_t)B * nx;
    const size_t n_y = yref_mode == MPCB_SHARED ? ny : yref_mode == MPCB_PER_INSTANCE ? (size_t)B * ny : (size_t)B * (N + 1) * ny;
    const size_t n_p = !p ? 0 : p_mode == MPCB_SHARED ? kNP : p_mode == MPCB_PER_INSTANCE ? (size_t)B * kNP : (size_t)B * N * kNP;
    const size_t n_u0 = u0 ? (size_t)B * nu : 0, n_X = X ? (size_t)B * (N + 1) * nx : 0, n_U = U ? (size_t)B * N * nu : 0;
    const size_t n_st = (size_t)B;  // int32 status + iters packed behind the doubles
    const size_t in_d = n_x0 + n_y + n_p;
    const size_t out_d = n_u0 + n_X + n_U;
    const size_t bytes = (in_d + out_d) * sizeof(double) + 2 * n_st * sizeof(int32_t);
    if (bytes > h->h_pin_bytes) {
        if (h->h_pin) cudaFreeHost(h->h_pin);
        cudaFree(h->d_stage);
        h->h_pin = nullptr; h->d_stage = nullptr; h->h_pin_bytes = h->d_stage_bytes = 0;
        CK(h, cudaMallocHost((void **)&h->h_pin, bytes));
        CK(h, cudaMalloc((void **)&h->d_stage, bytes));
        h->h_pin_bytes = h->d_stage_bytes = bytes;
    }
    cudaStream_t s = h->own_stream;
    double *hp = h->h_pin, *dp = h->d_stage;
    memcpy(hp, x0, n_x0 * sizeof(double));
    memcpy(hp + n_x0, yref, n_y * sizeof(double));
    if (p) memcpy(hp + n_x0 + n_y, p, n_p * sizeof(double));
    CK(h, cudaMemcpyAsync(dp, hp, in_d * sizeof(double), cudaMemcpyHostToDevice, s));
    double *d_u0 = dp + in_d, *d_X = d_u0 + n_u0, *d_U = d_X + n_X;
    int32_t *d_st = (int32_t *)(d_U + n_U), *d_it = d_st + n_st;
    int rc = mpcb_solve(h, dp, dp + n_x0, yref_mode, p ? dp + n_x0 + n_y : nullptr, p_mode, u0 ? d_u0 : nullptr,
                        X ? d_X : nullptr, U ? d_U : nullptr, d_st, d_it, B, s);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(hp + in_d, dp + in_d, out_d * sizeof(double) + 2 * n_st * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CK(h, cudaStreamSynchronize(s));
    if (u0) memcpy(u0, hp + in_d, n_u0 * sizeof(double));
    if (X) memcpy(X, hp + in_d + n_u0, n_X * sizeof(double));
    if (U) memcpy(U, hp + in_d + n_u0 + n_X, n_U * sizeof(double));
    const int32_t *h_st = (const int32_t *)(hp + in_d + out_d);
    if (status) memcpy(status, h_st, n_st * sizeof(int32_t));
    if (iters) memcpy(iters, h_st + n_st, n_st * sizeof(int32_t));
    return 0;
}

int mpcb_plant_step(mpcb_handle *h, const double *x, const double *u, const double *p, int p_mode, double *xnext,
                    int B, void *stream)
{
    if (check_batch(h, B)) return -1;
    if (!x || !u || !xnext) return fail(h, "null argument");
    if (p_mode != MPCB_SHARED && p_mode != MPCB_PER_INSTANCE) return fail(h, "plant p_mode must be SHARED or PER_INSTANCE");
    if (B == 0) return 0;
    if (!p) { p = h->p_default; p_mode = MPCB_SHARED; }
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (B + 127) / 128;
    if (h->nx == 17) plant_kernel<17, 6><<<grid, 128, 0, s>>>(h->P, x, u, p, p_mode, xnext, B);
    else if (h->nx == 13) plant_kernel<13, 4><<<grid, 128, 0, s>>>(h->P, x, u, p, p_mode, xnext, B);
    else plant_kernel<12, 4><<<grid, 128, 0, s>>>(h->P, x, u, p, p_mode, xnext, B);
    g_launches += 1;
    CK(h, cudaGetLastError());
    return 0;
}

int mpcb_closed_loop(mpcb_handle *h, double *x, const double *yref, int yref_mode, const double *p, int p_mode,
                     int steps, double *u_last, int32_t *n_fail, int32_t *iters_sum, int B, void *stream)
{
    if (check_batch(h, B)) return -1;
    if (!x || !yref || steps < 0) return fail(h, "bad argument");
    if (p && p_mode == MPCB_PER_STAGE) return fail(h, "closed loop takes p SHARED or PER_INSTANCE");
    if (B == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_fail) CK(h, cudaMemsetAsync(n_fail, 0, (size_t)B * sizeof(int32_t), s));
    if (iters_sum) CK(h, cudaMemsetAsync(iters_sum, 0, (size_t)B * sizeof(int32_t), s));
    double *u0 = u_last ? u_last : h->u0_scratch;
    for (int t = 0; t < steps; t++) {
        if (mpcb_solve(h, x, yref, yref_mode, p, p_mode, u0, nullptr, nullptr, h->status_scratch, h->iters_scratch, B, s))
            return -1;
        if (mpcb_plant_step(h, x, u0, p, p ? p_mode : MPCB_SHARED, h->xn_scratch, B, s)) return -1;
        const unsigned grid = (B + 127) / 128;
        if (h->nx == 17)
            loop_book_kernel<17><<<grid, 128, 0, s>>>(x, h->xn_scratch, h->status_scratch, h->iters_scratch, n_fail, iters_sum, B);
        else if (h->nx == 13)
            loop_book_kernel<13><<<grid, 128, 0, s>>>(x, h->xn_scratch, h->status_scratch, h->iters_scratch, n_fail, iters_sum, B);
        else
            loop_book_kernel<12><<<grid, 128, 0, s>>>(x, h->xn_scratch, h->status_scratch, h->iters_scratch, n_fail, iters_sum, B);
        g_launches += 1;
    }
    CK(h, cudaGetLastError());
    return 0;
}

int mpcb_cost(mpcb_handle *h, const double *yref, int yref_mode, double *cost, int B, void *stream)
{
    if (check_batch(h, B)) return -1;
    if (!yref || !cost) return fail(h, "null argument");
    if (B == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (B + 127) / 128;
    if (h->nx == 17) cost_kernel<17, 6><<<grid, 128, 0, s>>>(h->P, h->X, h->U, yref, yref_mode, cost, B);
    else if (h->nx == 13) cost_kernel<13, 4><<<grid, 128, 0, s>>>(h->P, h->X, h->U, yref, yref_mode, cost, B);
    else cost_kernel<12, 4><<<grid, 128, 0, s>>>(h->P, h->X, h->U, yref, yref_mode, cost, B);
    g_launches += 1;
    CK(h, cudaGetLastError());
    return 0;
}

int mpcb_get_iterate(mpcb_handle *h, double *X, double *U, int B, void *stream)
{
    if (check_batch(h, B)) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    if (X) CK(h, cudaMemcpyAsync(X, h->X, (size_t)B * (h->N + 1) * h->nx * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (U) CK(h, cudaMemcpyAsync(U, h->U, (size_t)B * h->N * h->nu * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return 0;
}

int mpcb_set_iterate(mpcb_handle *h, const double *X, const double *U, int B, void *stream)
{
    if (check_batch(h, B)) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    if (X) CK(h, cudaMemcpyAsync(h->X, X, (size_t)B * (h->N + 1) * h->nx * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (U) CK(h, cudaMemcpyAsync(h->U, U, (size_t)B * h->N * h->nu * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return 0;
}

int mpcb_debug_linearize(mpcb_handle *h, const double *p, int p_mode, double *BAt, double *b, int B, void *stream)
{
    if (check_batch(h, B)) return -1;
    if (!BAt || !b) return fail(h, "null argument");
    if (!p) { p = h->p_default; p_mode = MPCB_SHARED; }
    cudaStream_t s = (cudaStream_t)stream;
    for (int i0 = 0; i0 < B; i0 += h->ws_batch) {
        const int nb = (B - i0 < h->ws_batch) ? B - i0 : h->ws_batch;
        const long long warps = (long long)nb * h->N;
        const unsigned g1 = (unsigned)((warps * 32 + 127) / 128);
        if (h->nx == 17) {
            linearize_kernel<17, 6><<<g1, 128, 0, s>>>(h->P, h->X, h->U, p, p_mode, h->ws, i0, nb, nullptr);
            const size_t per = (size_t)h->N * (23 * 17 + 17);
            debug_copy_kernel<17, 6><<<(unsigned)((per * nb + 255) / 256), 256, 0, s>>>(h->P, h->ws, BAt, b, i0, nb);
        } else if (h->nx == 13) {
            linearize_kernel<13, 4><<<g1, 128, 0, s>>>(h->P, h->X, h->U, p, p_mode, h->ws, i0, nb, nullptr);
            const size_t per = (size_t)h->N * (17 * 13 + 13);
            debug_copy_kernel<13, 4><<<(unsigned)((per * nb + 255) / 256), 256, 0, s>>>(h->P, h->ws, BAt, b, i0, nb);
        } else {
            linearize_kernel<12, 4><<<g1, 128, 0, s>>>(h->P, h->X, h->U, p, p_mode, h->ws, i0, nb, nullptr);
            const size_t per = (size_t)h->N * (16 * 12 + 12);
            debug_copy_kernel<12, 4><<<(unsigned)((per * nb + 255) / 256), 256, 0, s>>>(h->P, h->ws, BAt, b, i0, nb);
        }
        g_launches += 2;
    }
    h->last_solve_batch = 0;  // the records no longer belong to the last solve
    CK(h, cudaGetLastError());
    return 0;
}

int mpcb_shift(mpcb_handle *h, int B, void *stream)
{
    if (check_batch(h, B)) return -1;
    if (B == 0) return 0;
    const unsigned grid = (B + 127) / 128;
    if (h->nx == 17) shift_kernel<17, 6><<<grid, 128, 0, (cudaStream_t)stream>>>(h->P, h->X, h->U, B);
    else if (h->nx == 13) shift_kernel<13, 4><<<grid, 128, 0, (cudaStream_t)stream>>>(h->P, h->X, h->U, B);
    else shift_kernel<12, 4><<<grid, 128, 0, (cudaStream_t)stream>>>(h->P, h->X, h->U, B);
    g_launches += 1;
    CK(h, cudaGetLastError());
    return 0;
}

namespace {
int sqp_alloc(mpcb_handle *h)
{
    if (h->sqp_done) return 0;
    const size_t B = (size_t)h->max_batch, s1 = (size_t)(h->N + 1);
    CK(h, cudaMalloc((void **)&h->sqp_pi, B * s1 * h->nx * sizeof(double)));
    CK(h, cudaMalloc((void **)&h->sqp_ll, B * s1 * (h->nx + h->nu) * sizeof(double)));
    CK(h, cudaMalloc((void **)&h->sqp_lu, B * s1 * (h->nx + h->nu) * sizeof(double)));
    CK(h, cudaMalloc((void **)&h->sqp_done, B * sizeof(int32_t)));
    CK(h, cudaMalloc((void **)&h->sqp_status, B * sizeof(int32_t)));
    CK(h, cudaMalloc((void **)&h->sqp_count, B * sizeof(int32_t)));
    CK(h, cudaMalloc((void **)&h->sqp_qpit, B * sizeof(int32_t)));
    return 0;
}
}  // namespace

int mpcb_solve_sqp(mpcb_handle *h, const double *x0, const double *yref, int yref_mode, const double *p, int p_mode,
                   int max_iter, const double *tol, double *u0, double *X, double *U, int32_t *status, int32_t *iters,
                   int32_t *sqp_iters, double *nlp_res, int B, void *stream)
{
    if (check_batch(h, B)) return -1;
    if (max_iter < 1) return fail(h, "max_iter must be >= 1");
    if (!x0 || !yref) return fail(h, "x0 and yref are required");
    if (yref_mode < 0 || yref_mode > 2 || p_mode < 0 || p_mode > 2) return fail(h, "bad yref_mode / p_mode");
    if (B == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (!tol) {
        // fixed number of SQP iterations, no residual test: status / iters are those of the last QP
        for (int it = 0; it < max_iter; it++) {
            const bool last = it + 1 == max_iter;
            if (mpcb_solve(h, x0, yref, yref_mode, p, p_mode, u0, last ? X : nullptr, last ? U : nullptr, status, iters, B, stream))
                return -1;
        }
        if (sqp_iters) { fill_i32_kernel<<<(B + 255) / 256, 256, 0, s>>>(sqp_iters, max_iter, B); g_launches += 1; }
        if (nlp_res) CK(h, cudaMemsetAsync(nlp_res, 0, (size_t)B * 4 * sizeof(double), s));
        CK(h, cudaGetLastError());
        return 0;
    }
    for (int i = 0; i < 4; i++)
        if (!(tol[i] >= 0)) return fail(h, "SQP tolerances must be non-negative");
    if (sqp_alloc(h)) return -1;
    if (!p) { p = h->p_default; p_mode = MPCB_SHARED; }
    const size_t s1 = (size_t)(h->N + 1);
    SqpState st;
    st.pi = h->sqp_pi; st.ll = h->sqp_ll; st.lu = h->sqp_lu;
    st.done = h->sqp_done; st.status = h->sqp_status; st.sqp_iters = h->sqp_count; st.qp_iters = h->sqp_qpit; st.res = nlp_res;
    for (int i = 0; i < 4; i++) st.tol[i] = tol[i];
    // acados' default initial multipliers are zero [upstream D4]
    CK(h, cudaMemsetAsync(st.pi, 0, (size_t)B * s1 * h->nx * sizeof(double), s));
    CK(h, cudaMemsetAsync(st.ll, 0, (size_t)B * s1 * (h->nx + h->nu) * sizeof(double), s));
    CK(h, cudaMemsetAsync(st.lu, 0, (size_t)B * s1 * (h->nx + h->nu) * sizeof(double), s));
    CK(h, cudaMemsetAsync(st.done, 0, (size_t)B * sizeof(int32_t), s));
    CK(h, cudaMemsetAsync(st.sqp_iters, 0, (size_t)B * sizeof(int32_t), s));
    CK(h, cudaMemsetAsync(st.qp_iters, 0, (size_t)B * sizeof(int32_t), s));
    fill_i32_kernel<<<(B + 255) / 256, 256, 0, s>>>(st.status, ST_MAXITER, B);
    g_launches += 1;
    // acados' SQP loop: linearise, evaluate the residuals, stop if converged, else solve the QP and take the full step;
    // after max_iter QPs the status is max-iter (one more linearisation reports the residuals of the final iterate).
    for (int it = 0; it <= max_iter; it++) {
        const bool eval_only = it == max_iter;
        int rc = (h->nx == 17)   ? launch_solve_chunks<17, 6>(h, x0, yref, yref_mode, p, p_mode, h->u0_scratch, h->status_scratch, h->iters_scratch, B, s, &st, it, eval_only)
                 : (h->nx == 13) ? launch_solve_chunks<13, 4>(h, x0, yref, yref_mode, p, p_mode, h->u0_scratch, h->status_scratch, h->iters_scratch, B, s, &st, it, eval_only)
                                 : launch_solve_chunks<12, 4>(h, x0, yref, yref_mode, p, p_mode, h->u0_scratch, h->status_scratch, h->iters_scratch, B, s, &st, it, eval_only);
        if (rc) return rc;
    }
    h->last_solve_batch = 0;
    const size_t nu = h->nu;
    if (u0) CK(h, cudaMemcpyAsync(u0, h->u0_scratch, (size_t)B * nu * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (status) CK(h, cudaMemcpyAsync(status, st.status, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (iters) CK(h, cudaMemcpyAsync(iters, st.qp_iters, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (sqp_iters) CK(h, cudaMemcpyAsync(sqp_iters, st.sqp_iters, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (X) CK(h, cudaMemcpyAsync(X, h->X, (size_t)B * (h->N + 1) * h->nx * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (U) CK(h, cudaMemcpyAsync(U, h->U, (size_t)B * h->N * h->nu * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return 0;
}

int mpcb_debug_qp(mpcb_handle *h, double *z, double *pi, double *tl, double *tu, double *ll, double *lu, double *lb, double *ub,
                  double *g, double *BAt, double *b, int B, void *stream)
{
    if (check_batch(h, B)) return -1;
    if (B == 0) return 0;
    if (B > h->last_solve_batch)
        return fail(h, "mpcb_debug_qp: the workspace does not hold the records of that many instances of the last mpcb_solve "
                       "(call it right after a solve of at most ws_batch instances)");
    cudaStream_t s = (cudaStream_t)stream;
    QpDebugOut o{z, pi, tl, tu, ll, lu, lb, ub, g};
    const int nz = h->nx + h->nu;
    const size_t per = (size_t)(h->N + 1) * nz;
    const unsigned grid = (unsigned)((per * B + 255) / 256);
    const size_t per2 = (size_t)h->N * ((size_t)nz * h->nx + h->nx);
    const unsigned grid2 = (unsigned)((per2 * B + 255) / 256);
    if (h->nx == 17) {
        debug_qp_kernel<17, 6><<<grid, 256, 0, s>>>(h->P, h->ws, o, 0, B);
        if (BAt && b) debug_copy_kernel<17, 6><<<grid2, 256, 0, s>>>(h->P, h->ws, BAt, b, 0, B);
    } else if (h->nx == 13) {
        debug_qp_kernel<13, 4><<<grid, 256, 0, s>>>(h->P, h->ws, o, 0, B);
        if (BAt && b) debug_copy_kernel<13, 4><<<grid2, 256, 0, s>>>(h->P, h->ws, BAt, b, 0, B);
    } else {
        debug_qp_kernel<12, 4><<<grid, 256, 0, s>>>(h->P, h->ws, o, 0, B);
        if (BAt && b) debug_copy_kernel<12, 4><<<grid2, 256, 0, s>>>(h->P, h->ws, BAt, b, 0, B);
    }
    g_launches += 2;
    CK(h, cudaGetLastError());
    return 0;
}

int mpcb_profile(mpcb_handle *h, int enable)
{
    if (!h) return -1;
    h->profile = enable != 0;
    return 0;
}

int mpcb_last_kernel_ms(mpcb_handle *h, float *ms_linearize, float *ms_qp)
{
    if (!h || !h->profile) return fail(h, "profiling not enabled");
    CK(h, cudaEventSynchronize(h->ev[2]));
    if (ms_linearize) CK(h, cudaEventElapsedTime(ms_linearize, h->ev[0], h->ev[1]));
    if (ms_qp) CK(h, cudaEventElapsedTime(ms_qp, h->ev[1], h->ev[2]));
    return 0;
}

int mpcb_fp64_peak(int device, double *tflops)
{
    if (!tflops) return -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    DeviceGuard guard(device >= 0 ? device : dev);
    if (guard.err != cudaSuccess) return -1;
    if (device >= 0) dev = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double *out = nullptr;
    if (cudaMalloc((void **)&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return -1;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(a);
        fp64_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    g_launches += 5;
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(out);
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mpcb_command_map(mpcb_handle *h, const double *x, const double *u0, double *quat, double *thrust, int B, void *stream)
{
    if (check_batch(h, B)) return -1;
    if (!x || (thrust && !u0)) return fail(h, "null argument");
    if (h->nx == 13) return fail(h, "command mapping takes Euler-angle states (variants 17 and 12); a QUAT13 state already carries the attitude quaternion");
    if (B == 0) return 0;
    command_map_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, u0, h->nx, h->nu, quat, thrust, B);
    g_launches += 1;
    CK(h, cudaGetLastError());
    return 0;
}

int mpcb_poc_jacobians(const double *euler, const double *motor, const double *position, const double *x17, int B,
                       double stream_velocity, double drag, int mode, double T_blast, double *poc, double *J_mot, double *J_eul,
                       double *J_pos, double *p25, double *t_flight, int32_t *status, void *stream)
{
    if (B < 0) return fail(nullptr, "negative batch");
    if (!x17 && (!euler || !motor || !position)) return fail(nullptr, "null argument: give euler/motor/position or x17");
    if (mode != POC_MODE_REFERENCE && mode != POC_MODE_ANALYTIC) return fail(nullptr, "mode must be MPCB_POC_REFERENCE or MPCB_POC_ANALYTIC");
    if (!(stream_velocity > 0) || !(drag > 0)) return fail(nullptr, "stream velocity and drag must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, "no CUDA device (this library has no CPU fallback)");
    if (B == 0) return 0;
    poc_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(euler, motor, position, x17, B, stream_velocity, drag, mode, T_blast, poc, J_mot,
                                                               J_eul, J_pos, p25, t_flight, status);
    g_launches += 1;
    if (cudaGetLastError() != cudaSuccess) return fail(nullptr, "poc_kernel launch failed");
    return 0;
}

#ifdef MPCB_PHASE_CLOCKS
// experiment build only (tools/phase_clocks.py)
int mpcb_debug_phase_clocks(unsigned long long *out16, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_phase_clk, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_phase_clk, z, sizeof(z)); }
    return 0;
}
#endif
}  // extern "C"
