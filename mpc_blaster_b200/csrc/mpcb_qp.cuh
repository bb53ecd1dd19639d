// Stage-structured QP solve of one SQP-RTI iteration: Mehrotra predictor-corrector
// interior point on a square-root Riccati factorisation.
// ONE MPC INSTANCE PER WARP.
//
// What the reference runs here: acados SQP_RTI -> HPIPM (qp_solver
// 'PARTIAL_CONDENSING_HPIPM' with qp_solver_cond_N = N, i.e. the full-horizon OCP-QP,
// reference blastermodel.py:274,284; cold start, acados_ocp_blasterModel.json
// qp_solver_warm_start = 0).  Same algorithm class, new design:
//   * lane j < NZ owns component j of the stage variable z_k = [du_k; dx_k] and row j of
//     the stage matrices;
//   * the Riccati factor L_k (L_k L_k' = diag(H_k + barrier) + [B A]' P_{k+1} [B A]) comes
//     from a Householder LQ of [sqrt(diag) | [B A]' L_{k+1}] -- the normal equations are never
//     formed, so active *state* bounds (barrier ~ 1e15) cost eps*sqrt(barrier) instead of
//     eps*barrier (the role of HPIPM's lq_fact);
//   * every stage-serial sweep runs out of shared memory: the next stage's record is
//     fetched with cp.async into the other half of a double buffer while the current stage
//     is being processed, so the recursion never waits on HBM/L2 latency;
//   * everything that is elementwise in the stage index (box slacks/multipliers, step-length
//     ratios, the update) is done in "flat" passes with independent loads, off the serial
//     critical path; the corrector is solved as an increment on the affine backward sweep.
//   Per IPM iteration the stage matrices are read four times and L is written once.
#pragma once
#include "mpcb_common.cuh"

namespace mpcb {

// Complementarity above which the latency variant (NSLOT = 2) factorises the stage matrix in normal-equations form
// (classical recursion on P_k: products on the FP64 tensor cores, NU pivots) instead of the Householder LQ;
// -DMPCB_GRAM_MU=1e30 switches it off (tools/ab.py).  History of the switch, launch time at 1,024 instances: 1e-4 -7.8 %
// (round 1, Gram + 23-pivot Cholesky), 1e-5 -1.3 % more (round 2); with the P form each interior-point iteration moved
// from the LQ saves 20 x 6.7 k cycles, and 3e-6 gives -3.4 % more with every parity assertion of the GPU suite still
// holding at a tolerance of 3e-7 (1e-7 except one 1.0e-7 < d < 3e-7 case; at 1e-5 all hold at 1e-7); 1e-6 fails three
// tests with deviations up to 3.6e-6 (profiles/r02_ab_gram_mu.txt, r02_parity_margins.txt).  The single-buffer throughput
// variant (168-register cap, 12 warps per SM) spills with it and is slower, so it keeps the LQ throughout.
// Also only while mu <= mu0: on an infeasible QP the multipliers diverge (mu climbs towards the 100 * mu0 exit) and the
// normal-equations form loses a pivot before the LQ does -- the instance would end with ST_QPFAIL where the checkers report ST_MINSTEP.
#ifndef MPCB_GRAM_MU
#define MPCB_GRAM_MU 3e-6
#endif
// W = [B A]' Lxx and the Gram matrix W W' of the latency variant on the FP64 tensor cores (mma.sync m8n8k4, DMMA):
// -DMPCB_DMMA=0 restores the CUDA-core products (tools/ab.py).
#ifndef MPCB_DMMA
#define MPCB_DMMA 1
#endif
// NX = 8 n + 1 (BLASTER17): last column of P_{k+1} on the CUDA cores instead of a mostly empty third tile; -DMPCB_TAIL_COLUMN=0: all on the tensor cores
#ifndef MPCB_TAIL_COLUMN
#define MPCB_TAIL_COLUMN 1
#endif
#if defined(MPCB_PHASE_CLOCKS) && !defined(MPCB_HOST_EMU)
// experiment build only (tools/phase_clocks.py): cycles per phase of the one-instance kernel, summed over warps
__device__ unsigned long long g_phase_clk[16];
#define MPCB_PH(i) do { if (lane == 0) { const long long t_ = clock64(); atomicAdd(&g_phase_clk[i], (unsigned long long)(t_ - ph_t)); ph_t = t_; } } while (0)
#define MPCB_PH_COUNT(i) do { if (lane == 0) atomicAdd(&g_phase_clk[i], 1ull); } while (0)
#define MPCB_PH_INIT long long ph_t = clock64()
#else
#define MPCB_PH(i)
#define MPCB_PH_COUNT(i)
#define MPCB_PH_INIT
#endif
constexpr double kMuDiverge = 1e2;  // infeasibility test: mu > kMuDiverge * mu0 (the CPU checkers apply the same test; no feasible instance of the test scenarios exceeds 3 * mu0)

// Per-warp shared memory: two stage-record images (same offsets as the global record) plus
// the data that is carried from one stage of a sweep to the next.
// NSLOT = 2: the next stage's record is prefetched while the current one is processed (latency
// variant, 25 KB per warp); NSLOT = 1: one buffer, fetched at the top of each stage (throughput
// variant for batches of many waves: 16 KB per warp, so more warps per SM hide the latency instead).
template <int NX, int NU, typename T, int NSLOT = 2>
struct QpSmem {
    using L = Layout<NX, NU>;
    // every member is a multiple of 4 elements long, so all of them stay 32-byte aligned
    alignas(32) T slot[NSLOT][L::STAGE];
    unsigned long long mbar[4];  // two mbarriers of the stage-prefetch pipeline (+ padding)
    T Lxx[L::LXX];       // factor of P_{k+1} (backward sweep), [NX][NX]; upper triangle stays zero
    T vrow[2][L::NXP];   // Householder pivot row broadcast (double buffered)
    T Lcol[L::NZ * L::NUP];  // first NU columns of the L_k being factorised, row-major [NZ][NUP]
    T Linv[L::NUP];      // 1/diag(Luu)
    T cPi[L::NXP], cZx[L::NXP], cPv[L::NXP], cDx[L::NXP];  // carried: pi_{k+1}, dx-part of z_{k+1}, p_{k+1}, dx_k
    T cDx2[L::NXP];      // forward sweeps: dx_k alternates between cDx and cDx2 (no sync between reading dx_k and writing dx_{k+1})
    T sT1[L::NXP], sT2[L::NXP], sDz[L::NZP], sRb[L::NXP];
    T hd[L::NZP], ds[L::NZP];  // Hd_k and sqrt(Hd_k) of every row, read by the pivot loop
};

// static description of component j of stage k
struct VarKind {
    bool var;   // is an optimisation variable (not the pinned x_0, not a u at stage N)
    bool hasb;  // has box bounds: lbu/ubu on stages 0..N-1, lbx/ubx on stages 1..N-1 [upstream D2]
};
template <int NX, int NU>
MPCB_DEV VarKind var_kind(int k, int j, int N)
{
    VarKind v;
    if (j < NU) { v.var = k < N; v.hasb = k < N; }
    else if (j < NX + NU) { v.var = k > 0; v.hasb = (k >= 1 && k < N); }
    else { v.var = false; v.hasb = false; }
    return v;
}
// Gauss-Newton LINEAR_LS Hessian diagonal: dt*diag(Q,R) on stages < N, Q_t at N [upstream D1]
template <int NX, int NU, typename T>
MPCB_DEV T hess_diag(const Params &P, int k, int j)
{
    if (j < NU) return (T)(P.dt * P.R[j]);
    if (j < NX + NU) return (k < P.N) ? (T)(P.dt * P.Q[j - NU]) : (T)P.Qt[j - NU];
    return T(1);
}

// Step of the box slacks / multipliers for a Newton step dz (one component):
//   dt_l = dz + r_dl, dt_u = -dz + r_du, dlam = -(r_m + lam*dt)/t
template <typename T>
struct BoxStep { T dtl, dtu, dll, dlu; };

template <typename T>
MPCB_DEV BoxStep<T> box_step(T z, T dz, T lb, T ub, T tl, T tu, T ll, T lu, T rml, T rmu, T itl, T itu)
{
    BoxStep<T> b;
    b.dtl = dz + (z - lb - tl);
    b.dtu = -dz + (ub - z - tu);
    b.dll = -(rml + ll * b.dtl) * itl;
    b.dlu = -(rmu + lu * b.dtu) * itu;
    return b;
}

// largest step keeping v + a*dv >= 0, as its reciprocal (0 when dv >= 0)
template <typename T>
MPCB_DEV T inv_ratio(T dv, T iv) { return dv < T(0) ? -dv * iv : T(0); }

// Inputs of the elementwise box computations for a batch of FBN consecutive stages
// (component `lane` of each stage), loaded together so their latencies overlap.
template <typename T, int FBN>
struct BoxIn {
    bool ok[FBN];   // has bounds (and the stage exists)
    bool var[FBN];  // is an optimisation variable
    T z[FBN], tl[FBN], tu[FBN], ll[FBN], lu[FBN], lb[FBN], ub[FBN], dza[FBN], dz[FBN];
};
template <int NX, int NU, typename T, int FBN, bool WITH_DZ>
MPCB_DEV void load_box(BoxIn<T, FBN> &in, const T *__restrict__ ws, int k0, int kend, int N, int lane)
{
    using L = Layout<NX, NU>;
    MPCB_UNROLL
    for (int u = 0; u < FBN; u++) {
        const int k = k0 + u;
        const VarKind vk = var_kind<NX, NU>(k, lane, N);
        in.ok[u] = (k < kend) && vk.hasb;
        in.var[u] = (k < kend) && vk.var;
        in.z[u] = in.dz[u] = in.dza[u] = in.lb[u] = in.ub[u] = in.ll[u] = in.lu[u] = T(0);
        in.tl[u] = in.tu[u] = T(1);
        const T *wk = ws + (size_t)k * L::STAGE;
        if (in.var[u]) {
            in.z[u] = wk[L::O_Z + lane];
            if (WITH_DZ) in.dz[u] = wk[L::O_DZ + lane];
        }
        if (in.ok[u]) {
            in.tl[u] = wk[L::O_TL + lane]; in.tu[u] = wk[L::O_TU + lane];
            in.ll[u] = wk[L::O_LL + lane]; in.lu[u] = wk[L::O_LUP + lane];
            in.lb[u] = wk[L::O_LB + lane]; in.ub[u] = wk[L::O_UB + lane];
            in.dza[u] = wk[L::O_DZA + lane];
        }
    }
}

// t2 = P r + p; r in sm.sRb, p in sm.cPv, result in sm.sT2.  sm.Lxx holds the factor of P (P = Lxx Lxx', zero upper
// triangle) or -- PFORM, the normal-equations iterations of the latency variant -- the symmetric matrix P itself.
template <int NX, int NU, typename T, int NSLOT>
MPCB_DEV void apply_P(QpSmem<NX, NU, T, NSLOT> &sm, bool pform)
{
    const int lane = lane_id();
    const int c = lane < NX ? lane : 0;
    if (pform) {
        T b0 = sm.cPv[c], b1 = T(0), b2 = T(0), b3 = T(0);
        MPCB_UNROLL
        for (int j = 0; j + 3 < NX; j += 4) {
            b0 += sm.Lxx[c * NX + j] * sm.sRb[j]; b1 += sm.Lxx[c * NX + j + 1] * sm.sRb[j + 1];
            b2 += sm.Lxx[c * NX + j + 2] * sm.sRb[j + 2]; b3 += sm.Lxx[c * NX + j + 3] * sm.sRb[j + 3];
        }
        MPCB_UNROLL
        for (int j = NX & ~3; j < NX; j++) b0 += sm.Lxx[c * NX + j] * sm.sRb[j];
        if (lane < NX) sm.sT2[lane] = (b0 + b1) + (b2 + b3);
        warp_sync();
        return;
    }
    T a0 = T(0), a1 = T(0);
    MPCB_UNROLL
    for (int j = 0; j + 1 < NX; j += 2) {
        a0 += sm.Lxx[j * NX + c] * sm.sRb[j];
        a1 += sm.Lxx[(j + 1) * NX + c] * sm.sRb[j + 1];
    }
    if (NX & 1) a0 += sm.Lxx[(NX - 1) * NX + c] * sm.sRb[NX - 1];
    if (lane < NX) sm.sT1[lane] = a0 + a1;
    warp_sync();
    T b0 = sm.cPv[c], b1 = T(0);
    MPCB_UNROLL
    for (int j = 0; j + 1 < NX; j += 2) {
        b0 += sm.Lxx[c * NX + j] * sm.sT1[j];
        b1 += sm.Lxx[c * NX + j + 1] * sm.sT1[j + 1];
    }
    if (NX & 1) b0 += sm.Lxx[c * NX + NX - 1] * sm.sT1[NX - 1];
    if (lane < NX) sm.sT2[lane] = b0 + b1;
    warp_sync();
}

// acc + sum_c a[c] * b[c], c < N: four accumulator chains in the latency variant (one warp per scheduler: the chains are
// the critical path), two in the register-capped throughput variant (measured: four cost it 1.6 % at 4,096 instances)
template <int N, int CHAINS, typename T, class FA, class FB>
MPCB_DEV T dot_chains(T acc, FA &&a, FB &&b)
{
    if constexpr (CHAINS == 4) {
        T s0 = acc, s1 = T(0), s2 = T(0), s3 = T(0);
        MPCB_UNROLL
        for (int c = 0; c + 3 < N; c += 4) { s0 += a(c) * b(c); s1 += a(c + 1) * b(c + 1); s2 += a(c + 2) * b(c + 2); s3 += a(c + 3) * b(c + 3); }
        MPCB_UNROLL
        for (int c = N & ~3; c < N; c++) s0 += a(c) * b(c);
        return (s0 + s1) + (s2 + s3);
    } else {
        T s0 = acc, s1 = T(0);
        MPCB_UNROLL
        for (int c = 0; c + 1 < N; c += 2) { s0 += a(c) * b(c); s1 += a(c + 1) * b(c + 1); }
        if (N & 1) s0 += a(N - 1) * b(N - 1);
        return s0 + s1;
    }
}

// Forward substitution with the first NU columns of L_k held row-wise in registers:
// on return lanes c < NU hold lvec_c = (Luu^{-1} l_u)_c and lanes NU.. hold l_x - Lxu lvec.
template <int NU, typename T>
MPCB_DEV T fwd_subst(T l, const T *Lu, const T *invd, int nz)
{
    const int lane = lane_id();
    T out = l;
    MPCB_UNROLL
    for (int c = 0; c < NU; c++) {
        const T lc = warp_shfl(out, c) * invd[c];
        if (lane == c) out = lc;
        if (lane > c && lane < nz) out -= Lu[c] * lc;
    }
    return out;
}

// The same substitution without the NU dependent shuffles: l_u goes through shared memory once (`bc`), Luu is read from
// the column-major factor image `lu` (lu[c * NZP + r] = L[r][c], the record layout) and EVERY lane runs the NU-step
// recurrence redundantly in registers -- same operations in the same order as fwd_subst, bit-identical results, but the
// dependent chain is NU x (FMA + MUL) instead of NU x (shuffle + MUL + FMA) (a shuffle of a double is WARPSYNC + 2 SHFL).
template <int NX, int NU, typename T>
MPCB_DEV T fwd_subst_img(T l, const T *Lu, const T *invd, const T *lu, T *bc, int nz)
{
    using L = Layout<NX, NU>;
    const int lane = lane_id();
    if (lane < NU) bc[lane] = l;
    warp_sync();
    T lv[NU];
    MPCB_UNROLL
    for (int c = 0; c < NU; c++) {
        T a = bc[c];
        MPCB_UNROLL
        for (int cc = 0; cc < c; cc++) a -= lu[cc * L::NZP + c] * lv[cc];
        lv[c] = a * invd[c];
    }
    T out = l;
    if (lane >= NU && lane < nz) {
        MPCB_UNROLL
        for (int c = 0; c < NU; c++) out -= Lu[c] * lv[c];
    }
    MPCB_UNROLL
    for (int c = 0; c < NU; c++)
        if (lane == c) out = lv[c];
    return out;
}

// One forward sweep: dz_k = [du_k; dx_k] with du_k = -Luu^{-T}(lvec_k + Lxu' dx_k),
// dx_{k+1} = r_k + [B A] dz_k.  FINAL additionally produces dpi_{k+1} = P_{k+1} dx_{k+1} + p_{k+1}.
// The elementwise box work of the step just computed is folded in (it only depends on dz_k and
// overlaps the latency of the recursion): the reciprocal of the largest admissible step `imax`,
// and for the affine sweep the sums that give mu_aff(alpha) plus the corrector gradient pieces.
template <int NX, int NU, typename T, int NSLOT, bool FINAL>
MPCB_DEV void forward_sweep(const Params &P, QpSmem<NX, NU, T, NSLOT> &sm, StagePipe &pipe, T *__restrict__ ws, T sigmu, T &imax_out,
                            T &acc1_out, T &acc2_out, bool pform = false)
{
    using L = Layout<NX, NU>;
    constexpr int NZ = L::NZ;
    constexpr int O_OUT = FINAL ? L::O_DZ : L::O_DZA;
    // in-lane substitution: latency variant only (measured on the single-buffer throughput variant, 168-register cap: 1.2 % slower)
    constexpr bool kInLane = (NSLOT == 2);
    const int lane = lane_id();
    const int N = P.N;
    // record k: [BAt | Lu | invd | lvec | rb | z tl tu ll lu lb ub]; FINAL: also [dza] and [Lxx | pv] of record k+1
    constexpr int RUN1 = L::O_G;
    constexpr int TOTAL = RUN1 + (FINAL ? L::NZP + L::LXX + L::NXP : 0);
    auto fetch = [&](int k, int half) {
        const T *wk = ws + (size_t)k * L::STAGE;
        T *dst = sm.slot[half];
        pipe_expect(pipe, half, TOTAL);
        pipe_copy(pipe, half, dst, wk, RUN1);
        if (FINAL) {
            pipe_copy(pipe, half, dst + L::O_DZA, wk + L::O_DZA, L::NZP);
            pipe_copy(pipe, half, dst + L::O_LXX, wk + L::STAGE + L::O_LXX, L::LXX + L::NXP);
        }
    };
    if (NSLOT == 2) fetch(0, 0);
    T imax = T(0), acc1 = T(0), acc2 = T(0);
    if (lane < NX) sm.cDx[lane] = T(0);
    for (int k = 0; k < N; k++) {
        T *wk = ws + (size_t)k * L::STAGE;
        const int half = (NSLOT == 2) ? (k & 1) : 0;
        const T *s = sm.slot[half];
        const T *dxc = (k & 1) ? sm.cDx2 : sm.cDx;  // dx_k
        T *dxn_ = (k & 1) ? sm.cDx : sm.cDx2;       // dx_{k+1}
        if (NSLOT == 2) {
            if (k + 1 < N) fetch(k + 1, (k + 1) & 1);
        } else {
            fetch(k, 0);
            if (k + 1 < N) {  // the runs the next stage will fetch: HBM -> L2
                l2_prefetch(wk + L::STAGE, RUN1, lane == 0);
                if (FINAL) {
                    l2_prefetch(wk + L::STAGE + L::O_DZA, L::NZP, lane == 0);
                    l2_prefetch(wk + 2 * L::STAGE + L::O_LXX, L::LXX + L::NXP, lane == 0);
                }
            }
        }
        pipe_wait(pipe, half);
        warp_sync();
        // du = -Luu^{-T} (lvec + Lxu' dx)
        T yy = T(0);
        if (lane < NU) {
            T a0 = s[L::O_LVEC + lane], a1 = T(0), a2 = T(0), a3 = T(0);
            const T *col = s + L::O_LU + lane * L::NZP + NU;
            MPCB_UNROLL
            for (int i = 0; i + 3 < NX; i += 4) {
                a0 += col[i] * dxc[i]; a1 += col[i + 1] * dxc[i + 1];
                a2 += col[i + 2] * dxc[i + 2]; a3 += col[i + 3] * dxc[i + 3];
            }
            MPCB_UNROLL
            for (int i = NX & ~3; i < NX; i++) a0 += col[i] * dxc[i];
            yy = -((a0 + a1) + (a2 + a3));
            if (kInLane) sm.sT1[lane] = yy;
        }
        T duv[NU];
        T dz = T(0);
        if constexpr (kInLane) {
            warp_sync();
            // back substitution du = Luu^{-T} yy, redundantly in every lane (yy through shared memory, Luu from the record
            // image): NU x (FMA + MUL) on the dependent chain instead of NU x (shuffle + MUL + FMA); every lane ends up with
            // all of du_k, so the product below takes it from registers and needs no broadcast of dz_k
            MPCB_UNROLL
            for (int i = NU - 1; i >= 0; i--) {
                T a = sm.sT1[i];
                MPCB_UNROLL
                for (int c = NU - 1; c > i; c--) a -= s[L::O_LU + i * L::NZP + c] * duv[c];
                duv[i] = a * s[L::O_INVD + i];
            }
            if (lane >= NU && lane < NZ) dz = dxc[lane - NU];
            MPCB_UNROLL
            for (int i = 0; i < NU; i++)
                if (lane == i) dz = duv[i];
            if (lane < NZ) wk[O_OUT + lane] = dz;
        } else {
            T du = T(0);
            MPCB_UNROLL
            for (int i = NU - 1; i >= 0; i--) {
                const T dui = warp_shfl(yy, i) * s[L::O_INVD + i];
                if (lane == i) du = dui;
                if (lane < i) yy -= s[L::O_LU + lane * L::NZP + i] * dui;
            }
            if (lane < NU) dz = du;
            else if (lane < NZ) dz = dxc[lane - NU];
            if (lane < NZ) {
                wk[O_OUT + lane] = dz;
                sm.sDz[lane] = dz;
            }
        }
        // box slacks / multipliers along this step (component `lane` of stage k)
        if (var_kind<NX, NU>(k, lane, N).hasb) {
            const T z = s[L::O_Z + lane], tl = s[L::O_TL + lane], tu = s[L::O_TU + lane];
            const T ll = s[L::O_LL + lane], lu = s[L::O_LUP + lane], lb = s[L::O_LB + lane], ub = s[L::O_UB + lane];
            const T itl = fast_rcp(tl), itu = fast_rcp(tu);
            T rml = ll * tl, rmu = lu * tu;
            if (FINAL) {
                const BoxStep<T> a = box_step(z, s[L::O_DZA + lane], lb, ub, tl, tu, ll, lu, rml, rmu, itl, itu);
                rml += a.dll * a.dtl - sigmu;
                rmu += a.dlu * a.dtu - sigmu;
            }
            const BoxStep<T> b = box_step(z, dz, lb, ub, tl, tu, ll, lu, rml, rmu, itl, itu);
            imax = fmax(imax, fmax(fmax(inv_ratio(b.dtl, itl), inv_ratio(b.dtu, itu)),
                                   fmax(inv_ratio(b.dll, fast_rcp(ll)), inv_ratio(b.dlu, fast_rcp(lu)))));
            if (!FINAL) {
                acc1 += ll * b.dtl + tl * b.dll + lu * b.dtu + tu * b.dlu;
                acc2 += b.dll * b.dtl + b.dlu * b.dtu;
                // corrector gradient: (dl_a dt_a - sigma mu)/t_l - (du_a dtu_a - sigma mu)/t_u = c1 - sigma mu * c2
                wk[L::O_C1 + lane] = b.dll * b.dtl * itl - b.dlu * b.dtu * itu;
                wk[L::O_C2 + lane] = itl - itu;
            }
        }
        // dx_{k+1} = rb_k + [B A] dz_k: du_k from this lane's registers (or the broadcast copy of dz_k), dx_k from the buffer read above
        // (dx_{k+1} goes to the other one)
        if (!kInLane) warp_sync();
        if (lane < NX) {
            T a0 = s[L::O_RB + lane], a1 = T(0), a2 = T(0), a3 = T(0);
            auto dzj = [&](auto J_) {
                constexpr int j = decltype(J_)::value;
                if constexpr (!kInLane) return sm.sDz[j];
                else if constexpr (j < NU) return duv[j];
                else return dxc[j - NU];
            };
            static_for<0, (NZ & ~3), 4>([&](auto J_) {
                constexpr int j = decltype(J_)::value;
                a0 += s[L::O_BAT + j * L::LDB + lane] * dzj(IntC<j>{});
                a1 += s[L::O_BAT + (j + 1) * L::LDB + lane] * dzj(IntC<j + 1>{});
                a2 += s[L::O_BAT + (j + 2) * L::LDB + lane] * dzj(IntC<j + 2>{});
                a3 += s[L::O_BAT + (j + 3) * L::LDB + lane] * dzj(IntC<j + 3>{});
            });
            static_for<(NZ & ~3), NZ>([&](auto J_) {
                constexpr int j = decltype(J_)::value;
                a0 += s[L::O_BAT + j * L::LDB + lane] * dzj(IntC<j>{});
            });
            const T dxn = (a0 + a1) + (a2 + a3);
            dxn_[lane] = dxn;
            if (FINAL) {
                // dpi_{k+1} = Lxx_{k+1} (Lxx_{k+1}' dx_{k+1}) + p_{k+1}
                sm.sRb[lane] = dxn;
            }
        }
        warp_sync();
        if (FINAL && pform) {
            // dpi_{k+1} = P_{k+1} dx_{k+1} + p_{k+1} with the matrix P itself in the record (normal-equations iterations)
            const int c = lane < NX ? lane : 0;
            const T *Px = s + L::O_LXX + c * NX;
            T b0 = s[L::O_PV + c], b1 = T(0), b2 = T(0), b3 = T(0);
            MPCB_UNROLL
            for (int j = 0; j + 3 < NX; j += 4) {
                b0 += Px[j] * sm.sRb[j]; b1 += Px[j + 1] * sm.sRb[j + 1];
                b2 += Px[j + 2] * sm.sRb[j + 2]; b3 += Px[j + 3] * sm.sRb[j + 3];
            }
            MPCB_UNROLL
            for (int j = NX & ~3; j < NX; j++) b0 += Px[j] * sm.sRb[j];
            if (lane < NX) wk[L::STAGE + L::O_DPI + lane] = (b0 + b1) + (b2 + b3);
            warp_sync();
        } else if (FINAL) {
            const int c = lane < NX ? lane : 0;
            const T *Lx = s + L::O_LXX;
            T a0 = T(0), a1 = T(0);
            MPCB_UNROLL
            for (int j = 0; j + 1 < NX; j += 2) {
                a0 += Lx[j * NX + c] * sm.sRb[j];
                a1 += Lx[(j + 1) * NX + c] * sm.sRb[j + 1];
            }
            if (NX & 1) a0 += Lx[(NX - 1) * NX + c] * sm.sRb[NX - 1];
            if (lane < NX) sm.sT1[lane] = a0 + a1;
            warp_sync();
            T b0 = s[L::O_PV + c], b1 = T(0);
            MPCB_UNROLL
            for (int j = 0; j + 1 < NX; j += 2) {
                b0 += Lx[c * NX + j] * sm.sT1[j];
                b1 += Lx[c * NX + j + 1] * sm.sT1[j + 1];
            }
            if (NX & 1) b0 += Lx[c * NX + NX - 1] * sm.sT1[NX - 1];
            if (lane < NX) wk[L::STAGE + L::O_DPI + lane] = b0 + b1;
            warp_sync();
        }
    }
    // terminal stage: dz_N = [0; dx_N]
    if (lane < NZ) ws[(size_t)N * L::STAGE + O_OUT + lane] = (lane < NU) ? T(0) : ((N & 1) ? sm.cDx2 : sm.cDx)[lane - NU];
    pipe_fence();
    warp_sync();
    imax_out = warp_max(imax);
    acc1_out = warp_sum(acc1);
    acc2_out = warp_sum(acc2);
}

// Forward sweep of an iterative-refinement increment (STRICT instantiation only; plain global loads, no prefetch
// pipeline): the backward increment sweep left [d lvec_k; d p_k] in the c1 slot of every record; this computes the step
// increment (ddu_k = -Luu^{-T}(d lvec_k + Lxu' ddx_k), ddx_{k+1} = [B A] ddz_k, ddpi_{k+1} = P_{k+1} ddx_{k+1} + d p_{k+1}),
// ADDS it to dz / dpi in the records and redoes the box step-length computation of the corrected step.
template <int NX, int NU, typename T, int NSLOT>
MPCB_DEV void refine_forward(const Params &P, QpSmem<NX, NU, T, NSLOT> &sm, T *__restrict__ ws, T sigmu, T &imax_out, bool pform)
{
    using L = Layout<NX, NU>;
    constexpr int NZ = L::NZ;
    const int lane = lane_id();
    const int N = P.N;
    T imax = T(0);
    if (lane < NX) sm.cDx[lane] = T(0);
    warp_sync();
    for (int k = 0; k < N; k++) {
        T *wk = ws + (size_t)k * L::STAGE;
        T yy = T(0);
        if (lane < NU) {
            T a = wk[L::O_C1 + lane];
            MPCB_UNROLL4
            for (int i = 0; i < NX; i++) a += wk[L::O_LU + lane * L::NZP + NU + i] * sm.cDx[i];
            yy = -a;
        }
        T du = T(0);
        MPCB_UNROLL
        for (int i = NU - 1; i >= 0; i--) {
            const T dui = warp_shfl(yy, i) * wk[L::O_INVD + i];
            if (lane == i) du = dui;
            if (lane < i) yy -= wk[L::O_LU + lane * L::NZP + i] * dui;
        }
        T ddz = T(0);
        if (lane < NU) ddz = du;
        else if (lane < NZ) ddz = sm.cDx[lane - NU];
        T dz = T(0);
        if (lane < NZ) {
            dz = wk[L::O_DZ + lane] + ddz;
            wk[L::O_DZ + lane] = dz;
            sm.sDz[lane] = ddz;
        }
        if (var_kind<NX, NU>(k, lane, N).hasb) {
            const T z = wk[L::O_Z + lane], tl = wk[L::O_TL + lane], tu = wk[L::O_TU + lane];
            const T ll = wk[L::O_LL + lane], lu = wk[L::O_LUP + lane], lb = wk[L::O_LB + lane], ub = wk[L::O_UB + lane];
            const T itl = fast_rcp(tl), itu = fast_rcp(tu);
            T rml = ll * tl, rmu = lu * tu;
            const BoxStep<T> a = box_step(z, wk[L::O_DZA + lane], lb, ub, tl, tu, ll, lu, rml, rmu, itl, itu);
            rml += a.dll * a.dtl - sigmu;
            rmu += a.dlu * a.dtu - sigmu;
            const BoxStep<T> b = box_step(z, dz, lb, ub, tl, tu, ll, lu, rml, rmu, itl, itu);
            imax = fmax(imax, fmax(fmax(inv_ratio(b.dtl, itl), inv_ratio(b.dtu, itu)),
                                   fmax(inv_ratio(b.dll, fast_rcp(ll)), inv_ratio(b.dlu, fast_rcp(lu)))));
        }
        warp_sync();
        // ddx_{k+1} = [B A] ddz_k (the dynamics residual of the first solve is zero to rounding by construction)
        if (lane < NX) {
            T a0 = T(0), a1 = T(0);
            MPCB_UNROLL4
            for (int j = 0; j + 1 < NZ; j += 2) {
                a0 += wk[L::O_BAT + j * L::LDB + lane] * sm.sDz[j];
                a1 += wk[L::O_BAT + (j + 1) * L::LDB + lane] * sm.sDz[j + 1];
            }
            if (NZ & 1) a0 += wk[L::O_BAT + (NZ - 1) * L::LDB + lane] * sm.sDz[NZ - 1];
            const T dxn = a0 + a1;
            sm.cDx[lane] = dxn;
            sm.sRb[lane] = dxn;
        }
        warp_sync();
        // ddpi_{k+1} = Lxx_{k+1} (Lxx_{k+1}' ddx_{k+1}) + d p_{k+1}
        const int c = lane < NX ? lane : 0;
        const T *Lx = wk + L::STAGE + L::O_LXX;
        T t1 = T(0);
        if (!pform) {
            MPCB_UNROLL4
            for (int j = 0; j < NX; j++) t1 += Lx[j * NX + c] * sm.sRb[j];
        }
        if (lane < NX) sm.sT1[lane] = pform ? sm.sRb[lane] : t1;  // P-form: the record holds P itself, one product
        warp_sync();
        T t2 = wk[L::STAGE + L::O_C1 + NU + c];
        MPCB_UNROLL4
        for (int j = 0; j < NX; j++) t2 += Lx[c * NX + j] * sm.sT1[j];
        if (lane < NX) wk[L::STAGE + L::O_DPI + lane] += t2;
        warp_sync();
    }
    if (lane >= NU && lane < NZ) ws[(size_t)N * L::STAGE + L::O_DZ + lane] += sm.cDx[lane - NU];
    pipe_fence();
    warp_sync();
    imax_out = warp_max(imax);
}

// The whole QP solve for one instance.  On return the persistent iterate Xi/Ui has taken
// the full step (FIXED_STEP, step length 1.0: acados_ocp_blasterModel.json globalization /
// nlp_solver_step_length).  Returns the status; *iters_out = IPM iterations.
// STRICT (mpcb_config.strict_reference): the stopping test uses the residual norms of the iterate evaluated explicitly
// in the backward sweep S1 (stationarity included, as HPIPM's test does) instead of their extrapolated values, there is
// no early exit on diverging multipliers, and the last iterate is applied when the iteration cap is reached.
template <int NX, int NU, typename T, int NSLOT, bool STRICT = false>
MPCB_DEV int qp_solve_warp(const Params &P, QpSmem<NX, NU, T, NSLOT> &sm, T *__restrict__ ws, T *__restrict__ Xi,
                           T *__restrict__ Ui, const T *__restrict__ x0, const T *__restrict__ yref, int yps,
                           int *iters_out)
{
    using L = Layout<NX, NU>;
    constexpr int NZ = L::NZ;
    const int lane = lane_id();
    const int N = P.N;
    const T thr0 = (T)P.ipm_thr0, mu0 = (T)P.ipm_mu0;
    const T nb = (T)(2 * NU * N + 2 * NX * (N - 1));
    constexpr int FB = 7;  // stages per batch in the flat update pass (all loads of a batch are in flight together)
    const T H0s = hess_diag<NX, NU, T>(P, 0, lane), H0N = hess_diag<NX, NU, T>(P, N, lane);  // stage / terminal weight of this lane

    for (int idx = lane; idx < NX * NX; idx += 32) sm.Lxx[idx] = T(0);
    StagePipe pipe;
    pipe_init(pipe, sm.mbar);

    // ---------------- F0: QP data and cold start [upstream D8]: z = 0 (dx_0 pinned), pi = 0,
    // t >= slack floor, lam = mu0/t.  Cost gradient and bounds on the increments.
    T eg = T(0), eb = T(0), ed = T(0);
    MPCB_UNROLL4
    for (int k = 0; k <= N; k++) {
        T *wk = ws + (size_t)k * L::STAGE;
        const VarKind vk = var_kind<NX, NU>(k, lane, N);
        const T H0 = (k < N) ? H0s : H0N;
        const T *yr = yref + (yps ? (size_t)k * NZ : 0);
        T y = T(0), g = T(0), lb = T(0), ub = T(0), z = T(0);
        if (lane < NU) {
            if (k < N) {
                y = Ui[(size_t)k * NU + lane];
                g = H0 * (y - yr[NX + lane]);
                lb = (T)P.lbu[lane] - y; ub = (T)P.ubu[lane] - y;
            }
        } else if (lane < NZ) {
            const int i = lane - NU;
            y = Xi[(size_t)k * NX + i];
            g = H0 * (y - yr[i]);
            if (vk.hasb) { lb = (T)P.lbx[i] - y; ub = (T)P.ubx[i] - y; }
            if (k == 0) z = x0[i] - y;  // [upstream D3] dx_0 = x0 - X_0
        }
        T tl = T(1), tu = T(1), ll = T(0), lu = T(0);
        if (vk.hasb) {
            // slack floor: absolute (thr0 >= 0) or the fraction -thr0 of the box width (thr0 < 0)
            const T flo = (thr0 >= T(0)) ? thr0 : -thr0 * (ub - lb);
            tl = fmax(z - lb, flo);
            tu = fmax(ub - z, flo);
            ll = mu0 / tl;
            lu = mu0 / tu;
            ed = fmax(ed, fmax(fabs(z - lb - tl), fabs(ub - z - tu)));
        }
        if (vk.var) eg = fmax(eg, fabs(H0 * z + g - ll + lu));
        if (lane < NZ) {
            wk[L::O_Z + lane] = z; wk[L::O_TL + lane] = tl; wk[L::O_TU + lane] = tu;
            wk[L::O_LL + lane] = ll; wk[L::O_LUP + lane] = lu;
            wk[L::O_LB + lane] = lb; wk[L::O_UB + lane] = ub; wk[L::O_G + lane] = g;
        }
        if (lane < NX) {
            wk[L::O_PI + lane] = T(0);
            if (k < N) {
                T rb = wk[L::O_B + lane];
                if (k == 0)
                    for (int i = 0; i < NX; i++) rb += wk[L::O_BAT + (NU + i) * L::LDB + lane] * (x0[i] - Xi[i]);
                eb = fmax(eb, fabs(rb));
            }
        }
    }
    T est_g = warp_max(eg), est_b = warp_max(eb), est_d = warp_max(ed);
    T comp = mu0, mu = mu0;
    int status = ST_MAXITER, it = 0;
    pipe_fence();  // the QP data written above is fetched by the bulk-copy pipeline below
    warp_sync();
    MPCB_PH_INIT;

    for (it = 0; it < P.ipm_max_iter; it++) {
        if (!(est_g == est_g) || !(est_b == est_b) || !(mu == mu)) { status = ST_NAN; break; }
        if (!STRICT) {
            // diverging multipliers (mean complementarity 100 x its starting value) are the signature of
            // an infeasible QP: stop instead of running to the iteration cap (same status as min-step)
            if (mu > T(kMuDiverge) * mu0) { status = ST_MINSTEP; break; }
            if (est_g <= (T)P.tol_stat && est_b <= (T)P.tol_eq && est_d <= (T)P.tol_ineq && comp <= (T)P.tol_comp) {
                status = ST_OK;
                break;
            }
        }
        // explicit residual norms of this iterate, gathered by S1 as by-products of the right-hand side: dynamics (xb) and
        // bound slacks (xd) always -- est_b / est_d of the next test are these MEASURED values times (1 - alpha), not an
        // extrapolation from the cold start -- stationarity (xg) in the STRICT instantiation only
        T xg = T(0), xb = T(0), xd = T(0);
        // latency variant only (both measured on the single-buffer throughput variant, 168-register cap, 12 warps per SM:
        // tensor-core products 0.4 .. 1 % slower, profiles/r02_ab_dmma_tp.txt; hybrid factorisation +4 % at 4,096 instances,
        // -0.5 % at 6,144, -5.6 % at 16,384, -8 % at 65,536, profiles/r02_ab_tp_hybrid.txt)
        constexpr bool kDmma = (NSLOT == 2) && (MPCB_DMMA != 0);
        constexpr bool kGramFactor = (NSLOT == 2) && (MPCB_GRAM_MU < 1e29);
        // this iteration factorises in normal-equations form (early iterations, see the stage loop) ...
        const bool gram_it = kGramFactor && mu > T(MPCB_GRAM_MU) && mu <= (T)P.ipm_mu0;
        // ... and then carries P_k itself from stage to stage instead of its Cholesky factor (sm.Lxx, the Lxx field of
        // the records): the classical Riccati recursion.  Only the NU input pivots are eliminated; what is left in the
        // state rows is P_k, which the next stage needs only inside [B A]' P_k [B A] -- its factor would be multiplied
        // back together there.  17 of the 23 pivots of the dependent chain (store -> sync -> load -> rsqrt -> FMA) go.
        const bool pform = kDmma && gram_it;
        MPCB_PH(0);
        MPCB_PH_COUNT(11);
        // ================= S1: backward sweep -- residuals, factorisation, affine right-hand side
        // record k: run A = [BAt], run B = [z tl tu ll lu lb ub g pi b]
        constexpr int RUNB = L::O_C1 - L::O_Z;
        T last_sig = T(1);
        auto fetch1 = [&](int k, int half) {
            const T *wk = ws + (size_t)k * L::STAGE;
            pipe_expect(pipe, half, L::BAT + RUNB);
            pipe_copy(pipe, half, sm.slot[half], wk, L::BAT);
            pipe_copy(pipe, half, sm.slot[half] + L::O_Z, wk + L::O_Z, RUNB);
        };
        {
            // terminal stage N: no inputs, no bounds; L_N = sqrt(Q_t), p_N = q_N
            T *wN = ws + (size_t)N * L::STAGE;
            if (NSLOT == 2) fetch1(N - 1, (N - 1) & 1);
            if (lane >= NU && lane < NZ) {
                const int i = lane - NU;
                const T H0 = H0N;
                const T zN = wN[L::O_Z + lane], piN = wN[L::O_PI + i];
                const T q = H0 * zN + wN[L::O_G + lane] - piN;
                if (STRICT) xg = fmax(xg, fabs(q));
                // (the whole row: a P-form iteration leaves a full symmetric matrix behind, the factor form needs the zero upper triangle)
                MPCB_UNROLL
                for (int c = 0; c < NX; c++) sm.Lxx[i * NX + c] = (c == i) ? (pform ? H0 : sqrt(H0)) : T(0);
                sm.cPv[i] = q;
                sm.cPi[i] = piN;
                sm.cZx[i] = zN;
                wN[L::O_PV + i] = q;
            }
            warp_sync();
            for (int idx = lane; idx < NX * NX; idx += 32) wN[L::O_LXX + idx] = sm.Lxx[idx];
        }
        for (int k = N - 1; k >= 0; k--) {
            T *wk = ws + (size_t)k * L::STAGE;
            const int half = (NSLOT == 2) ? (k & 1) : 0;
            const T *s = sm.slot[half];
            if (NSLOT == 2) {
                if (k > 0) fetch1(k - 1, (k - 1) & 1);
            } else {
                fetch1(k, 0);
                if (k > 0) {  // the runs the next stage will fetch: HBM -> L2
                    l2_prefetch(wk - L::STAGE, L::BAT, lane == 0);
                    l2_prefetch(wk - L::STAGE + L::O_Z, RUNB, lane == 0);
                }
            }
            pipe_wait(pipe, half);
            warp_sync();
            const int jr = lane < NZ ? lane : 0;
            T brow[NX];
            MPCB_UNROLL
            for (int c = 0; c < NX; c++) brow[c] = s[L::O_BAT + jr * L::LDB + c];
            const VarKind vk = var_kind<NX, NU>(k, lane, N);
            const T zj = s[L::O_Z + jr];
            T Hd = H0s, q = T(0);
            {
                T ll = T(0), lu = T(0), tl = T(1), tu = T(1);
                if (vk.hasb) { tl = s[L::O_TL + lane]; tu = s[L::O_TU + lane]; ll = s[L::O_LL + lane]; lu = s[L::O_LUP + lane]; }
                if (vk.var) {
                    q = dot_chains<NX, (NSLOT == 2 ? 4 : 2)>(Hd * zj + s[L::O_G + lane] - ll + lu, [&](int c) { return brow[c]; }, [&](int c) { return sm.cPi[c]; });
                    if (lane >= NU) q -= s[L::O_PI + lane - NU];
                    if (STRICT) xg = fmax(xg, fabs(q));
                }
                if (vk.hasb) {
                    const T itl = fast_rcp(tl), itu = fast_rcp(tu);
                    const T rdl = zj - s[L::O_LB + lane] - tl, rdu = s[L::O_UB + lane] - zj - tu;
                    xd = fmax(xd, fmax(fabs(rdl), fabs(rdu)));
                    Hd += ll * itl + lu * itu;
                    // affine right-hand side (r_m = lam*t):  q += lam_l + lam_l r_dl/t_l - lam_u - lam_u r_du/t_u
                    q += (ll + ll * rdl * itl) - (lu + lu * rdu * itu);
                }
            }
            if (kDmma && lane < NZ) sm.hd[lane] = Hd;  // read back per tile after the Gram product (a warp_sync lies between)
            // r_k = b_k + [B A] z_k - dx-part of z_{k+1}
            if (lane < NX) {
                T a0 = s[L::O_B + lane] - sm.cZx[lane], a1 = T(0), a2 = T(0), a3 = T(0);
                MPCB_UNROLL
                for (int j = 0; j + 3 < NZ; j += 4) {
                    a0 += s[L::O_BAT + j * L::LDB + lane] * s[L::O_Z + j];
                    a1 += s[L::O_BAT + (j + 1) * L::LDB + lane] * s[L::O_Z + j + 1];
                    a2 += s[L::O_BAT + (j + 2) * L::LDB + lane] * s[L::O_Z + j + 2];
                    a3 += s[L::O_BAT + (j + 3) * L::LDB + lane] * s[L::O_Z + j + 3];
                }
                MPCB_UNROLL
                for (int j = NZ & ~3; j < NZ; j++) a0 += s[L::O_BAT + j * L::LDB + lane] * s[L::O_Z + j];
                const T rb = (a0 + a1) + (a2 + a3);
                xb = fmax(xb, fabs(rb));
                wk[L::O_RB + lane] = rb;
                sm.sRb[lane] = rb;
            }
            warp_sync();
            // t2 = P_{k+1} r_k + p_{k+1}
            apply_P<NX, NU, T, NSLOT>(sm, pform);
            // carry this stage's pi and dx-part of z to stage k-1 (cPi / cZx were consumed above)
            if (lane < NX) { sm.cPi[lane] = s[L::O_PI + lane]; sm.cZx[lane] = s[L::O_Z + NU + lane]; }
            // affine backward vector before the substitution: l = q + [B A]' t2 (the last use of this lane's row of [B A]')
            T lin;
            {
                lin = dot_chains<NX, (NSLOT == 2 ? 4 : 2)>(q, [&](int c) { return brow[c]; }, [&](int c) { return sm.sT2[c]; });
            }
            // P form, NX = 8 n + 1 (BLASTER17): the last column of P_{k+1} would cost a third column tile and a fifth k-step
            // of mostly padding on the tensor cores (27 of 75 DMMAs); it is taken on the CUDA cores instead -- T[:, NX-1]
            // here, row per lane, while this lane's row of [B A]' is still in registers
            constexpr bool kTail = (NSLOT == 2) && (MPCB_DMMA != 0) && (NX % 8 == 1) && (MPCB_TAIL_COLUMN != 0);
            T t_tail = T(0);
            if (kTail && pform)
                t_tail = dot_chains<NX, 4>(T(0), [&](int c) { return brow[c]; }, [&](int c) { return sm.Lxx[(NX - 1) * NX + c]; });
            // W = [B A]' Lxx_{k+1}
            T w[NX];  // row `lane` of W (CUDA-core product, or read back from the tiles for the Householder loop)
            // Tensor-core form: W as NI x NJ tiles of 8 x 8, accumulated over NS k-steps of 4 rows of Lxx (lower triangular:
            // tiles above the diagonal band are skipped).  Lane 4 g + q supplies A = [B A]'[8 I + g][4 s + q] and
            // B = Lxx[4 s + q][8 J + g] straight from the record image / the factor in shared memory and ends up with
            // W[8 I + g][8 J + 2 q + h], h = 0, 1.  27 DMMAs + 24 loads instead of 153 DFMAs + 153 broadcast loads (BLASTER17).
            constexpr int NI = (NZ + 7) / 8, NJ = (NX + 7) / 8, NS = (NX + 3) / 4;
            const int tg = lane >> 2, tq = lane & 3;
            T wt[NI][NJ][2];
            T baf[2 * NJ][NI];  // P form: [B A]' fragments shared by the two products of the stage
            constexpr int NXT = kTail ? NX - 1 : NX, NJT = (NXT + 7) / 8;  // columns / column tiles the P-form products take on the tensor cores
            if constexpr (kDmma) {
                static_for<0, NI>([&](auto I_) {
                    static_for<0, NJ>([&](auto J_) { wt[decltype(I_)::value][decltype(J_)::value][0] = T(0); wt[decltype(I_)::value][decltype(J_)::value][1] = T(0); });
                });
                auto wprod = [&]() {  // factor form: W = [B A]' Lxx_{k+1}, Lxx lower triangular
                    static_for<0, NS>([&](auto S_) {
                        constexpr int st = decltype(S_)::value;
                        const int kk = 4 * st + tq;       // row of Lxx = column of [B A]' of this lane's fragments
                        const bool kin = kk < NX;
                        T af[NI];
                        static_for<0, NI>([&](auto I_) {
                            constexpr int I = decltype(I_)::value;
                            const int row = 8 * I + tg;
                            const bool in = kin && row < NZ;
                            const T v = s[L::O_BAT + (in ? row * L::LDB + kk : 0)];
                            af[I] = in ? v : T(0);
                        });
                        static_for<0, NJ>([&](auto J_) {
                            constexpr int J = decltype(J_)::value;
                            if constexpr (8 * J <= 4 * st + 3) {  // some row of this k-step reaches the tile's columns (Lxx[j][c] = 0 for c > j)
                                const int col = 8 * J + tg;
                                const bool in = kin && col < NX;
                                const T v = sm.Lxx[in ? kk * NX + col : 0];
                                const T bf = in ? v : T(0);
                                static_for<0, NI>([&](auto I_) { constexpr int I = decltype(I_)::value; warp_dmma(wt[I][J][0], wt[I][J][1], af[I], bf); });
                            }
                        });
                    });
                };
                if (pform) {
                    // P form: the k-steps run over (Jc, h) with k-index q <-> column 8 Jc + 2 q + h of [B A]' -- the assignment the
                    // product M = T [B A] below needs for its B fragments, which are these A fragments: loaded once (baf)
                    static_for<0, NJT>([&](auto C_) {
                        constexpr int Jc = decltype(C_)::value;
                        static_for<0, 2>([&](auto H_) {
                            constexpr int h = decltype(H_)::value;
                            if constexpr (8 * Jc + h < NXT) {
                                const int kk = 8 * Jc + 2 * tq + h;
                                const bool kin = kk < NXT;
                                static_for<0, NI>([&](auto I_) {
                                    constexpr int I = decltype(I_)::value;
                                    const int row = 8 * I + tg;
                                    const bool in = kin && row < NZ;
                                    const T v = s[L::O_BAT + (in ? row * L::LDB + kk : 0)];
                                    baf[2 * Jc + h][I] = in ? v : T(0);
                                });
                                static_for<0, NJT>([&](auto J_) {
                                    constexpr int J = decltype(J_)::value;
                                    const int col = 8 * J + tg;
                                    const bool in = kin && col < NXT;
                                    const T v = sm.Lxx[in ? kk * NX + col : 0];
                                    const T bf = in ? v : T(0);
                                    static_for<0, NI>([&](auto I_) { constexpr int I = decltype(I_)::value; warp_dmma(wt[I][J][0], wt[I][J][1], baf[2 * Jc + h][I], bf); });
                                });
                            }
                        });
                    });
                    if constexpr (kTail) {
                        // row NX-1 of P_{k+1} (the k-index the tiles left out) into the tiles: T[i][c] += [B A]'[i][NX-1] P[NX-1][c]
                        static_for<0, NI>([&](auto I_) {
                            constexpr int I = decltype(I_)::value;
                            const int row = 8 * I + tg;
                            const T a = (row < NZ) ? s[L::O_BAT + (row < NZ ? row : 0) * L::LDB + NX - 1] : T(0);
                            static_for<0, NJT>([&](auto J_) {
                                constexpr int J = decltype(J_)::value;
                                wt[I][J][0] += a * sm.Lxx[(NX - 1) * NX + 8 * J + 2 * tq];
                                wt[I][J][1] += a * sm.Lxx[(NX - 1) * NX + 8 * J + 2 * tq + 1];
                            });
                        });
                    }
                } else {
                    wprod();
                }
            } else {
                MPCB_UNROLL
                for (int c = 0; c < NX; c++) {
                    T a = T(0);
                    MPCB_UNROLL
                    for (int j = c; j < NX; j++) a += brow[j] * sm.Lxx[j * NX + c];
                    w[c] = a;
                }
            }
            T Lu[NU], invd[NU];
            MPCB_PH(1);
            if (gram_it) {
            // Early interior-point iterations (mu > MPCB_GRAM_MU): the normal-equations form, as HPIPM's default
            // Riccati -- M = diag(Hd) + W W' formed row by row (lane i owns row i, the rows of W broadcast from
            // shared memory), then a right-looking Cholesky, fully unrolled so that row i stays in registers: per
            // pivot one shared-memory round (column j, unscaled, diagonal included), one rsqrt, 22-j FMAs.  About
            // half the instructions of the Householder LQ below, but it loses eps*|P| where the LQ loses
            // eps*sqrt|P|, and |P| grows like 1/mu on active state bounds: used on every iteration it changes
            // iteration counts and moves u by up to 8e-6 against the checker; confined to mu > 1e-4 the solutions
            // agree with the all-LQ solver's to the same 1e-11 (emulator, tests/test_kernel_emulation.py).
            {
                constexpr int LDW = (NX + 1) & ~1;  // rows of the W image stay 16-byte aligned
                static_assert(NZ * LDW <= L::STAGE - L::O_C1, "the W image must fit the part of the record image this sweep does not fetch");
                static_assert(2 * L::NZP <= L::NZ * L::NUP, "column buffer must fit sm.Lcol");
                if (!kDmma && lane >= NZ) {
                    MPCB_UNROLL
                    for (int c = 0; c < NX; c++) w[c] = T(0);
                }
                T *Wsh = sm.slot[half] + L::O_C1;
                T m[NZ + 1];
                MPCB_UNROLL
                for (int c = 0; c <= NZ; c++) m[c] = T(0);
                if constexpr (kDmma) {
                    // M = T [B A] with T = [B A]' P_{k+1} in the accumulator registers (the product above).  The contraction runs
                    // over the columns of T, and ANY assignment of columns to k-indices is valid as long as A and B use the same
                    // one: lane (g, q) feeds T[8 I + g][8 Jc + 2 q + h], which it holds, as A, and [B A]'[8 J + g][8 Jc + 2 q + h]
                    // as B -- the fragment it loaded as A for the product above (baf) -- the k-steps run over (Jc, h).
                    // NI (NI + 1) / 2 tiles x the k-steps that reach a column < NX (30 DMMAs, BLASTER17).
                    T mt[NI][NI][2];
                    static_for<0, NI>([&](auto I_) {
                        static_for<0, NI>([&](auto J_) { mt[decltype(I_)::value][decltype(J_)::value][0] = T(0); mt[decltype(I_)::value][decltype(J_)::value][1] = T(0); });
                    });
                    static_for<0, NJT>([&](auto C_) {
                        constexpr int Jc = decltype(C_)::value;
                        static_for<0, 2>([&](auto H_) {
                            constexpr int h = decltype(H_)::value;
                            if constexpr (8 * Jc + h < NXT) {
                                static_for<0, NI>([&](auto I_) {
                                    constexpr int I = decltype(I_)::value;
                                    static_for<0, I + 1>([&](auto J_) {
                                        constexpr int J = decltype(J_)::value;
                                        warp_dmma(mt[I][J][0], mt[I][J][1], wt[I][Jc][h], baf[2 * Jc + h][J]);
                                    });
                                });
                            }
                        });
                    });
                    // tiles -> rows: block I (rows 8 I .. 8 I + 7) is stored with 8 (I + 1) columns at an even, padded stride
                    // (bank-conflict-free 128-bit row reads); the diagonal gets Hd on its way; lane r then reads row r.
                    // Entries right of a block's width belong to the unused upper triangle: whatever is read there never
                    // reaches a stored value (column j of L only takes m[j] of the lanes below the diagonal).
                    auto moff = [](int I) { return 8 * (4 * I * (I + 1) + 2 * I); };      // sum_{i<I} 8 (8 (i + 1) + 2)
                    auto mstride = [](int I) { return 8 * (I + 1) + 2; };
                    static_assert(8 * (4 * (NI - 1) * NI + 2 * (NI - 1)) + ((NZ - 1) % 8) * (8 * NI + 2) + L::NZP <= L::STAGE - L::O_C1,
                                  "the Gram image (and a full-width read of its last row) must fit the free tail of the record image");
                    static_for<0, NI>([&](auto I_) {
                        constexpr int I = decltype(I_)::value;
                        const bool rin = 8 * I + tg < NZ;
                        // Hd of row 8 I + g sits on the diagonal tile in lane q = g / 2, element g % 2
                        const T hdv = sm.hd[rin ? 8 * I + tg : 0];
                        if (rin && (tg >> 1) == tq) { if (tg & 1) mt[I][I][1] += hdv; else mt[I][I][0] += hdv; }
                        static_for<0, I + 1>([&](auto J_) {
                            constexpr int J = decltype(J_)::value;
                            sp_st2<0>(sptr_of(Wsh + moff(I) + tg * mstride(I) + 8 * J + 2 * tq), mt[I][J][0], mt[I][J][1], rin);
                        });
                    });
                    warp_sync();
                    {
                        const int r = lane < NZ ? lane : 0;
                        const sptr mrow = sptr_of(Wsh + moff(r >> 3) + (r & 7) * mstride(r >> 3));
                        static_for<0, NZ, 2>([&](auto Cc) {
                            constexpr int c = decltype(Cc)::value;
                            sp_ld2<c>(mrow, m[c], m[c + 1]);
                        });
                    }
                    if constexpr (kTail) {
                        // the column the tiles left out: M[i][c] += T[i][NX-1] [B A]'[c][NX-1], row per lane
                        static_for<0, NZ>([&](auto Cc) {
                            constexpr int c = decltype(Cc)::value;
                            m[c] += t_tail * s[L::O_BAT + c * L::LDB + NX - 1];
                        });
                    }
                } else {
                sp_row_store<0, NX>(sptr_of(Wsh + (lane < NZ ? lane : 0) * LDW), w, lane < NZ);
                warp_sync();
                const sptr w0 = sptr_of(Wsh);
                auto gram = [&](auto C) {
                    constexpr int c = decltype(C)::value;
                    T v[NX];
                    sp_row_load<0, NX>(sptr_add(w0, c * LDW), v);
                    T d0 = T(0), d1 = T(0), d2 = T(0), d3 = T(0);
                    MPCB_UNROLL
                    for (int i = 0; i + 3 < NX; i += 4) { d0 += v[i] * w[i]; d1 += v[i + 1] * w[i + 1]; d2 += v[i + 2] * w[i + 2]; d3 += v[i + 3] * w[i + 3]; }
                    MPCB_UNROLL
                    for (int i = NX & ~3; i < NX; i++) d0 += v[i] * w[i];
                    m[c] = ((d0 + d1) + (d2 + d3)) + (lane == c ? Hd : T(0));
                };
                static_for<0, NU>(gram);
                if (k > 0) static_for<NU, NZ>(gram);
                }
                const sptr cb0 = sptr_of(sm.Lcol);
                const sptr cbl = sptr_add(cb0, lane < NZ ? lane : 0);
                T sig = T(1);
                auto pivot = [&](auto J) {
                    constexpr int j = decltype(J)::value;
                    constexpr int par = (j & 1) * L::NZP;
                    sp_st1<par>(cbl, m[j], lane < NZ);
                    warp_sync();
                    T a[NZ + 1];
                    static_for<(j & ~1), NZ, 2>([&](auto Cc) {
                        constexpr int c = decltype(Cc)::value;
                        sp_ld2<par + c>(cb0, a[c], a[c + 1]);
                    });
                    const T rs = fast_rsqrt(a[j]);
                    sig = a[j] * rs;
                    const T f = m[j] * (rs * rs);
                    const T lij = m[j] * rs;
                    MPCB_UNROLL
                    for (int c = j + 1; c < NZ; c++) m[c] -= f * a[c];
                    const T val = (lane == j) ? sig : (lane > j ? lij : T(0));
                    if constexpr (j < NU) { Lu[j] = val; invd[j] = rs; }
                    else { if (lane >= j && lane < NZ) sm.Lxx[(lane - NU) * NX + (j - NU)] = val; }
                };
                static_for<0, NU>(pivot);
                if constexpr (kDmma) {
                    // P form: after the input pivots the state rows hold the Schur complement P_k (lower part valid); it
                    // replaces P_{k+1} in shared memory as a full symmetric matrix (the products above are done with it)
                    if (k > 0 && lane >= NU && lane < NZ) {
                        const int i = lane - NU;
                        static_for<0, NX>([&](auto C_) {
                            constexpr int c = decltype(C_)::value;
                            if (c <= i) { sm.Lxx[i * NX + c] = m[NU + c]; sm.Lxx[c * NX + i] = m[NU + c]; }
                        });
                    }
                } else {
                    if (k > 0) static_for<NU, NZ>(pivot);
                }
                last_sig = sig;
            }
            warp_sync();
            MPCB_PH(2);
            MPCB_PH_COUNT(9);
            } else {
            const T dsq = sqrt(Hd);
            if (lane < NZ) { sm.hd[lane] = Hd; sm.ds[lane] = dsq; }  // visible after the first pivot's warp_sync
            if constexpr (kDmma) {
                // the Householder loop works on row `lane` of W in registers: tiles -> W image -> rows
                constexpr int LDW = (NX + 1) & ~1;
                static_assert(NZ * LDW <= L::STAGE - L::O_C1, "the W image must fit the part of the record image this sweep does not fetch");
                T *Wsh = sm.slot[half] + L::O_C1;
                static_for<0, NI>([&](auto I_) {
                    constexpr int I = decltype(I_)::value;
                    static_for<0, NJ>([&](auto J_) {
                        constexpr int J = decltype(J_)::value;
                        const int row = 8 * I + tg;
                        sp_st2<0>(sptr_of(Wsh + (row < NZ ? row : 0) * LDW + 8 * J + 2 * tq), wt[I][J][0], wt[I][J][1], row < NZ && 8 * J + 2 * tq < LDW);
                    });
                });
                warp_sync();
                sp_row_load<0, NX>(sptr_of(Wsh + (lane < NZ ? lane : 0) * LDW), w);
            }
            // Householder LQ of [diag(dsq) | W], one pivot row per step:
            //   sigma^2 = Hd_j + |w_j|^2,  L_ij = (w_i . w_j)/sigma,
            //   w_i -= L_ij * kappa * w_j,  kappa = 1/(sigma + dsq_j) = (sigma - dsq_j)/|w_j|^2
            // (the second form lets 1/|w_j|^2 be computed beside rsqrt(sigma^2) instead of after it;
            // its cancellation error is O(eps |w_i|)).  The loop is deliberately
            // NOT unrolled: its body (~2 KB of SASS) then stays in the L0 instruction cache, which
            // matters at 1-2 resident warps per scheduler.  Column j of L goes to shared memory
            // (sm.Lcol, [NZ][NZ+1]); at stage 0 only the u-block is needed (x_0 is pinned).
            const int jend = (k == 0) ? NU : NZ;
            T sig = T(1);
            if (lane >= NZ) {
                MPCB_UNROLL
                for (int c = 0; c < NX; c++) w[c] = T(0);  // idle lanes carry zero rows: no masks needed below
            }
            const sptr vr0 = sptr_of(sm.vrow[0]);
            // column j of L goes to sm.Lcol ([NZ][NUP], j < NU) or straight into sm.Lxx (j >= NU, lower part)
            const sptr lu_row = sptr_of(sm.Lcol + (lane < NZ ? lane : 0) * L::NUP);
            const sptr lxx_row = sptr_add(sptr_of(sm.Lxx + (lane >= NU && lane < NZ ? lane - NU : 0) * NX), -NU);
            const sptr linv = sptr_of(sm.Linv);
            MPCB_NOUNROLL
            for (int j = 0; j < jend; j++) {
                const sptr vr = sptr_add(vr0, (j & 1) * L::NXP);
                const bool piv = (lane == j);
                T v[NX];
                // shared-memory broadcast of the pivot row: same latency as 34 shuffles, less MIO pressure (measured in round 1)
                sp_row_store<0, NX>(vr, w, piv);
                warp_sync();
                sp_row_load<0, NX>(vr, v);
                T d0 = T(0), d1 = T(0), d2 = T(0), d3 = T(0);
                MPCB_UNROLL
                for (int c = 0; c + 3 < NX; c += 4) {
                    d0 += v[c] * w[c]; d1 += v[c + 1] * w[c + 1]; d2 += v[c + 2] * w[c + 2]; d3 += v[c + 3] * w[c + 3];
                }
                MPCB_UNROLL
                for (int c = NX & ~3; c < NX; c++) d0 += v[c] * w[c];
                const T dot = (d0 + d1) + (d2 + d3);
                // no shuffle in the loop (+3.4 % at 1,024 instances against shuffling Hd_j, sqrt(Hd_j) and the pivot's dot): Hd / sqrt(Hd) of the pivot come from shared memory and every lane
                // forms |w_j|^2 itself from the broadcast row (same summation order as the pivot lane's dot)
                const T hdj = sm.hd[j], dsj = sm.ds[j];
                T djj;
                {
                    T e0 = T(0), e1 = T(0), e2 = T(0), e3 = T(0);
                    MPCB_UNROLL
                    for (int c = 0; c + 3 < NX; c += 4) { e0 += v[c] * v[c]; e1 += v[c + 1] * v[c + 1]; e2 += v[c + 2] * v[c + 2]; e3 += v[c + 3] * v[c + 3]; }
                    MPCB_UNROLL
                    for (int c = NX & ~3; c < NX; c++) e0 += v[c] * v[c];
                    djj = (e0 + e1) + (e2 + e3);
                }
                const T s2v = hdj + djj;
                const T rs = fast_rsqrt(s2v);
                const T idjj = fast_rcp(djj);
                sig = s2v * rs;
                const T kap = (djj > T(0)) ? (sig - dsj) * idjj : T(0);
                const T lij = (lane > j) ? dot * rs : T(0);
                const T f = lij * kap;
                MPCB_UNROLL
                for (int c = 0; c < NX; c++) w[c] -= f * v[c];
                const bool upart = j < NU;
                sp_st1<0>(sptr_add(upart ? lu_row : lxx_row, j), piv ? sig : lij, lane < NZ && (upart || lane >= j));
                sp_st1<0>(sptr_add(linv, j), rs, piv && upart);
            }
            last_sig = sig;
            warp_sync();

            MPCB_UNROLL
            for (int c = 0; c < NU; c++) { Lu[c] = sm.Lcol[(lane < NZ ? lane : 0) * L::NUP + c]; invd[c] = sm.Linv[c]; }
            MPCB_PH(3);
            MPCB_PH_COUNT(10);
            }
            // the factor's input block goes to the (unfetched) Lu slice of the record image for the in-lane substitution
            T l;
            if constexpr (NSLOT == 2) {
                T *luimg = sm.slot[half] + L::O_LU;
                if (lane < NU) {
                    MPCB_UNROLL
                    for (int c = 0; c < NU; c++) luimg[c * L::NZP + lane] = Lu[c];
                }
                l = fwd_subst_img<NX, NU, T>(lin, Lu, invd, luimg, sm.sT1, NZ);
            } else {
                l = fwd_subst<NU, T>(lin, Lu, invd, NZ);
            }
            if (lane < NU) wk[L::O_LVEC + lane] = l;
            else if (lane < NZ) { wk[L::O_PV + lane - NU] = l; sm.cPv[lane - NU] = l; }
            if (lane < NZ) {
                MPCB_UNROLL
                for (int c = 0; c < NU; c++) wk[L::O_LU + c * L::NZP + lane] = Lu[c];
            }
            if (lane < NU) {
                T mine = invd[0];
                MPCB_UNROLL
                for (int c = 1; c < NU; c++) if (lane == c) mine = invd[c];
                wk[L::O_INVD + lane] = mine;
            }
            warp_sync();
            if (k > 0)
                for (int idx = lane; idx < NX * NX; idx += 32) wk[L::O_LXX + idx] = sm.Lxx[idx];
            MPCB_PH(4);
        }
        pipe_fence();  // L, lvec, r_b, p written by this sweep are fetched by the next ones
        warp_sync();
        xb = warp_max(xb);
        xd = warp_max(xd);
        if (STRICT) {
            // reference semantics: test the residuals S1 has just evaluated on this iterate (before using the factorisation)
            xg = warp_max(xg);
            if (!(xg == xg) || !(xb == xb)) { status = ST_NAN; break; }
            if (xg <= (T)P.tol_stat && xb <= (T)P.tol_eq && xd <= (T)P.tol_ineq && comp <= (T)P.tol_comp) { status = ST_OK; break; }
        }
        // a breakdown (NaN) anywhere in the recursion propagates into the last pivot of stage 0
        if (!(last_sig == last_sig) || !(last_sig < T(HUGE_VAL))) { status = ST_QPFAIL; break; }

        // ================= S2: forward sweep, affine step (+ its step length, mu_aff, corrector pieces)
        T a_aff, mu_aff, sigmu;
        {
            T imax, s1, s2;
            forward_sweep<NX, NU, T, NSLOT, false>(P, sm, pipe, ws, T(0), imax, s1, s2);
            a_aff = (imax > T(1)) ? T(1) / imax : T(1);
            mu_aff = (mu * nb + a_aff * s1 + a_aff * a_aff * s2) / nb;
            T sigma = mu_aff / mu;
            sigma = sigma * sigma * sigma;
            sigmu = sigma * mu;
        }
        MPCB_PH(5);

        // ================= S3: backward sweep for the corrector increment (delta form)
        // record k: [BAt | Lu | invd | lvec] and [c1 c2]; pv_k is read-modify-written in global memory.
        // REF = 1 (STRICT only): the same backward sweep for an iterative-refinement increment -- the right-hand side is the
        // residual rho the flat pass below left in the c1 slot (every row, and a terminal-stage part = the increment of p_N).
        auto corrector_sweep = [&](auto REF) {
            constexpr bool REFINE = decltype(REF)::value != 0;
            constexpr int RUN1 = L::O_RB;
            auto fetch3 = [&](int k, int half) {
                const T *wk = ws + (size_t)k * L::STAGE;
                pipe_expect(pipe, half, RUN1 + 2 * L::NZP);
                pipe_copy(pipe, half, sm.slot[half], wk, RUN1);
                pipe_copy(pipe, half, sm.slot[half] + L::O_C1, wk + L::O_C1, 2 * L::NZP);
            };
            if (NSLOT == 2) fetch3(N - 1, (N - 1) & 1);
            if (lane < NX) sm.cPv[lane] = REFINE ? ws[(size_t)N * L::STAGE + L::O_C1 + NU + lane] : T(0);  // increment of p_N
            for (int k = N - 1; k >= 0; k--) {
                T *wk = ws + (size_t)k * L::STAGE;
                const int half = (NSLOT == 2) ? (k & 1) : 0;
                const T *s = sm.slot[half];
                const T pv_old = (lane >= NU && lane < NZ) ? wk[L::O_PV + lane - NU] : T(0);
                if (NSLOT == 2) {
                    if (k > 0) fetch3(k - 1, (k - 1) & 1);
                } else {
                    fetch3(k, 0);
                    if (k > 0) {  // the runs the next stage will fetch: HBM -> L2
                        l2_prefetch(wk - L::STAGE, RUN1, lane == 0);
                        l2_prefetch(wk - L::STAGE + L::O_C1, 2 * L::NZP, lane == 0);
                        l2_prefetch(wk - L::STAGE + L::O_PV, L::NXP, lane == 0);
                    }
                }
                pipe_wait(pipe, half);
                warp_sync();
                const int jr = lane < NZ ? lane : 0;
                const VarKind vk = var_kind<NX, NU>(k, lane, N);
                const T l0 = dot_chains<NX, (NSLOT == 2 ? 4 : 2)>(
                    REFINE ? (lane < NZ ? s[L::O_C1 + jr] : T(0)) : (vk.hasb ? s[L::O_C1 + lane] - sigmu * s[L::O_C2 + lane] : T(0)),
                    [&](int c) { return s[L::O_BAT + jr * L::LDB + c]; }, [&](int c) { return sm.cPv[c]; });
                const T l1 = T(0);
                T Lu[NU], invd[NU];
                MPCB_UNROLL
                for (int c = 0; c < NU; c++) { Lu[c] = s[L::O_LU + c * L::NZP + jr]; invd[c] = s[L::O_INVD + c]; }
                const T l = (NSLOT == 2) ? fwd_subst_img<NX, NU, T>((lane < NZ) ? l0 + l1 : T(0), Lu, invd, s + L::O_LU, sm.sT1, NZ)
                                         : fwd_subst<NU, T>((lane < NZ) ? l0 + l1 : T(0), Lu, invd, NZ);
                warp_sync();
                if (REFINE) {
                    // the increment stays separate ([d lvec; d p_k] replaces rho in the c1 slot): refine_forward adds the
                    // step it yields to dz / dpi, so the rounding of the first solve is corrected, not repeated
                    if (lane < NZ) wk[L::O_C1 + lane] = l;
                    if (lane >= NU && lane < NZ) sm.cPv[lane - NU] = l;
                } else {
                    if (lane < NU) wk[L::O_LVEC + lane] = s[L::O_LVEC + lane] + l;
                    else if (lane < NZ) { wk[L::O_PV + lane - NU] = pv_old + l; sm.cPv[lane - NU] = l; }
                }
                warp_sync();
            }
            pipe_fence();
            warp_sync();
        };
        corrector_sweep(IntC<0>{});
        MPCB_PH(6);

        // ================= S4: forward sweep, full predictor-corrector step (+ dpi, + step length)
        T alpha;
        {
            T imax, d1, d2;
            forward_sweep<NX, NU, T, NSLOT, true>(P, sm, pipe, ws, sigmu, imax, d1, d2, pform);
            // alpha = min(1, max(0.995, 1 - mu_aff) * alpha_max)
            const T tau = fmax(T(0.995), T(1) - mu_aff);
            alpha = (imax > tau) ? tau / imax : T(1);
        }
        MPCB_PH(7);
        if (STRICT) {
            // ================= R: one step of iterative refinement on the step just computed (HPIPM: itref_corr_max).
            // The residual of the stationarity rows of the reduced Newton system,
            //   rho_k = Hd_k dz_k + q_k + [B A]' dpi_{k+1} - [dpi_k]_x,   q_k = the corrector's right-hand side,
            // evaluated from the data in one flat pass (the dynamics rows hold to rounding by construction of the forward
            // sweep), goes through the same factorisation as a delta (S3 with REF = 1) and the forward sweep is redone.
            // With active state bounds the Riccati solve alone leaves the explicit stationarity norm at 1e-5 .. 1e-3
            // (in the multipliers; the primal step is already accurate); with this step the explicit norms reach
            // HPIPM's tolerances after the same number of iterations the extrapolated test predicts.
            MPCB_NOUNROLL
            for (int k = 0; k <= N; k++) {
                T *wk = ws + (size_t)k * L::STAGE;
                const VarKind vk = var_kind<NX, NU>(k, lane, N);
                T rho = T(0);
                if (vk.var) {
                    const T H0 = (k < N) ? H0s : H0N;
                    const T z = wk[L::O_Z + lane], dz = wk[L::O_DZ + lane];
                    T r = H0 * z + wk[L::O_G + lane], Hd = H0, dsum = T(0);
                    if (k < N) {
                        MPCB_UNROLL4
                        for (int c = 0; c < NX; c++) {
                            const T a = wk[L::O_BAT + lane * L::LDB + c];
                            r += a * wk[L::STAGE + L::O_PI + c];
                            dsum += a * wk[L::STAGE + L::O_DPI + c];
                        }
                    }
                    if (lane >= NU) { r -= wk[L::O_PI + lane - NU]; dsum -= wk[L::O_DPI + lane - NU]; }
                    if (vk.hasb) {
                        const T tl = wk[L::O_TL + lane], tu = wk[L::O_TU + lane], ll = wk[L::O_LL + lane], lu = wk[L::O_LUP + lane];
                        const T itl = T(1) / tl, itu = T(1) / tu;
                        const T rdl = z - wk[L::O_LB + lane] - tl, rdu = wk[L::O_UB + lane] - z - tu;
                        Hd += ll * itl + lu * itu;
                        r += ll * rdl * itl - lu * rdu * itu + (wk[L::O_C1 + lane] - sigmu * wk[L::O_C2 + lane]);
                    }
                    rho = Hd * dz + r + dsum;
                }
                if (lane < NZ) { wk[L::O_C1 + lane] = rho; wk[L::O_C2 + lane] = T(0); }
            }
            pipe_fence();
            warp_sync();
            corrector_sweep(IntC<1>{});
            T imax;
            refine_forward<NX, NU, T, NSLOT>(P, sm, ws, sigmu, imax, pform);
            const T tau = fmax(T(0.995), T(1) - mu_aff);
            alpha = (imax > tau) ? tau / imax : T(1);
        }
        // ================= F4b: take the step
        {
            T cmax = T(0), msum = T(0);
            // (ascending: the pass ends on the records the backward sweep S1 of the next iteration starts with; running it
            // last-batch-first, so that it starts on what S4 touched last, measured 1.6 % slower at 1,024 instances)
            for (int k0 = 0; k0 <= N; k0 += FB) {
                BoxIn<T, FB> in;
                load_box<NX, NU, T, FB, true>(in, ws, k0, N + 1, N, lane);
                T pi[FB], dpi[FB];
                MPCB_UNROLL
                for (int u = 0; u < FB; u++) {
                    pi[u] = dpi[u] = T(0);
                    if (k0 + u >= 1 && k0 + u <= N && lane < NX) {
                        const T *wk = ws + (size_t)(k0 + u) * L::STAGE;
                        pi[u] = wk[L::O_PI + lane];
                        dpi[u] = wk[L::O_DPI + lane];
                    }
                }
                MPCB_UNROLL
                for (int u = 0; u < FB; u++) {
                    const int k = k0 + u;
                    if (k > N) continue;
                    T *wk = ws + (size_t)k * L::STAGE;
                    if (in.ok[u]) {
                        T tl = in.tl[u], tu = in.tu[u], ll = in.ll[u], lu = in.lu[u];
                        const T itl = fast_rcp(tl), itu = fast_rcp(tu);
                        const BoxStep<T> a = box_step(in.z[u], in.dza[u], in.lb[u], in.ub[u], tl, tu, ll, lu, ll * tl, lu * tu, itl, itu);
                        const BoxStep<T> b = box_step(in.z[u], in.dz[u], in.lb[u], in.ub[u], tl, tu, ll, lu, ll * tl + a.dll * a.dtl - sigmu,
                                                      lu * tu + a.dlu * a.dtu - sigmu, itl, itu);
                        tl += alpha * b.dtl; tu += alpha * b.dtu; ll += alpha * b.dll; lu += alpha * b.dlu;
                        wk[L::O_TL + lane] = tl; wk[L::O_TU + lane] = tu; wk[L::O_LL + lane] = ll; wk[L::O_LUP + lane] = lu;
                        cmax = fmax(cmax, fmax(ll * tl, lu * tu));
                        msum += ll * tl + lu * tu;
                    }
                    if (in.var[u]) wk[L::O_Z + lane] = in.z[u] + alpha * in.dz[u];
                    if (k >= 1 && lane < NX) wk[L::O_PI + lane] = pi[u] + alpha * dpi[u];
                }
            }
            comp = warp_max(cmax);
            mu = warp_sum(msum) / nb;
        }
        est_g *= (T(1) - alpha);
        est_b = xb * (T(1) - alpha);  // measured on this iterate by S1, then the exact-arithmetic decay of one step
        est_d = xd * (T(1) - alpha);
        pipe_fence();
        warp_sync();
        MPCB_PH(8);
        if (!(alpha >= (T)P.alpha_min)) {
            status = (alpha == alpha) ? ST_MINSTEP : ST_NAN;
            it++;
            break;
        }
    }

    // ---------------- RTI update: X += dx, U += du (full step).  A failed QP (status != 0: NaN,
    // max-iter, min-step = infeasible, breakdown) leaves the iterate untouched, so one bad solve
    // cannot poison the warm start of the following control steps.  STRICT: the last iterate is
    // applied when the iteration cap was hit, as acados does with HPIPM's max-iter return.
    warp_sync();
    const bool take = status == ST_OK || (STRICT && status == ST_MAXITER);
    MPCB_UNROLL4
    for (int k = 0; k <= (take ? N : -1); k++) {
        const T *wk = ws + (size_t)k * L::STAGE;
        if (lane < NU) {
            if (k < N) Ui[(size_t)k * NU + lane] += wk[L::O_Z + lane];
        } else if (lane < NZ) {
            Xi[(size_t)k * NX + lane - NU] += wk[L::O_Z + lane];
        }
    }
    *iters_out = it;
    return status;
}

}  // namespace mpcb
