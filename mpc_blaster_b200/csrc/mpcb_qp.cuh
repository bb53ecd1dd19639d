// Stage-structured QP solve of one SQP-RTI iteration: Mehrotra predictor-corrector
// interior point on a square-root Riccati factorisation.
// ONE MPC INSTANCE PER WARP.
//
// What the reference runs here: acados SQP_RTI -> HPIPM (qp_solver
// 'PARTIAL_CONDENSING_HPIPM' with qp_solver_cond_N = N, i.e. the full-horizon OCP-QP,
// reference blastermodel.py:274,284; cold start, acados_ocp_blasterModel.json
// qp_solver_warm_start = 0).  Same algorithm class, new design:
//   * lane j < NZ owns component j of the stage variable z_k = [du_k; dx_k] and row j of
//     the stage matrices; stage matrices are staged in shared memory, vectors are
//     exchanged through shared memory or shuffles;
//   * the Riccati factor L_k (L_k L_k' = diag(H_k + barrier) + [B A]' P_{k+1} [B A]) is
//     obtained by Householder LQ of [sqrt(diag) | [B A]' L_{k+1}] -- never forming the
//     normal equations, so active *state* bounds (barrier ~ 1e15) cost eps*sqrt(barrier)
//     instead of eps*barrier (the role of HPIPM's lq_fact);
//   * the corrector is solved as a correction on top of the affine backward sweep, so per
//     IPM iteration the stage matrices are read four times and L is written once.
#pragma once
#include "mpcb_common.cuh"

namespace mpcb {

template <int NX, int NU, typename T>
struct QpSmem {
    using L = Layout<NX, NU>;
    T BAt[L::NZ * L::LDB];
    T Lxx[NX * NX];
    T Lu[NU * L::NZP];
    T vrow[2][L::NXP];
    T sPi[L::NZP], sZ[L::NZP], sRb[L::NZP], sT1[L::NZP], sT2[L::NZP], sPv[L::NZP], sDx[L::NZP], sDz[L::NZP];
};

// Everything lane j needs to know about component j of stage k.
template <typename T>
struct StageVar {
    bool var;    // is an optimisation variable (not the pinned x_0, not a u at stage N)
    bool hasb;   // has box bounds
    T H0, g, lb, ub;
};

// Gauss-Newton LINEAR_LS cost and bounds on the increments:
// stage Hessian dt*diag(Q,R), terminal Q_t unscaled [upstream D1]; lbu/ubu on stages
// 0..N-1, lbx/ubx on stages 1..N-1 [upstream D2]; x_0 pinned [upstream D3].
template <int NX, int NU, typename T>
MPCB_DEV StageVar<T> stage_var(const Params &P, int k, int j, const T *__restrict__ Xi, const T *__restrict__ Ui,
                               const T *__restrict__ yref, int yref_per_stage)
{
    constexpr int NZ = NX + NU;
    StageVar<T> s;
    const int N = P.N;
    s.var = false; s.hasb = false; s.H0 = T(1); s.g = T(0); s.lb = T(0); s.ub = T(0);
    if (j >= NZ) return s;
    const T *yr = yref + (yref_per_stage ? (size_t)k * NZ : 0);
    if (j < NU) {
        if (k < N) {
            const T y = Ui[(size_t)k * NU + j];
            const T w = (T)(P.dt * P.R[j]);
            s.var = true; s.hasb = true; s.H0 = w; s.g = w * (y - yr[NX + j]);
            s.lb = (T)P.lbu[j] - y; s.ub = (T)P.ubu[j] - y;
        }
    } else {
        const int i = j - NU;
        const T y = Xi[(size_t)k * NX + i];
        const T w = (k < N) ? (T)(P.dt * P.Q[i]) : (T)P.Qt[i];
        s.var = (k > 0); s.H0 = w; s.g = w * (y - yr[i]);
        if (k >= 1 && k < N) { s.hasb = true; s.lb = (T)P.lbx[i] - y; s.ub = (T)P.ubx[i] - y; }
    }
    return s;
}

// Step of the box slacks / multipliers for Newton step dz (one component).
template <typename T>
struct BoxStep { T dtl, dtu, dll, dlu; };

template <typename T>
MPCB_DEV BoxStep<T> box_step(T z, T dz, T lb, T ub, T tl, T tu, T ll, T lu, T rml, T rmu)
{
    BoxStep<T> b;
    const T rdl = z - lb - tl, rdu = ub - z - tu;
    b.dtl = dz + rdl;
    b.dtu = -dz + rdu;
    b.dll = -(rml + ll * b.dtl) / tl;
    b.dlu = -(rmu + lu * b.dtu) / tu;
    return b;
}

template <typename T>
MPCB_DEV T ratio(T v, T dv) { return dv < T(0) ? -v / dv : T(HUGE_VAL); }

template <int NX, int NU, typename T>
MPCB_DEV void load_BAt(QpSmem<NX, NU, T> &sm, const T *__restrict__ wk)
{
    using L = Layout<NX, NU>;
    const int lane = lane_id();
    for (int idx = lane; idx < L::NZ * NX; idx += 32) sm.BAt[(idx / NX) * L::LDB + (idx % NX)] = wk[L::O_BAT + idx];
}

// t2 = P r + p with P = Lxx Lxx' (Lxx in shared memory); r in sm.sRb, p in sm.sPv, result in sm.sT2.
template <int NX, int NU, typename T>
MPCB_DEV void apply_P(QpSmem<NX, NU, T> &sm)
{
    const int lane = lane_id();
    if (lane < NX) {
        T a = T(0);
        for (int j = lane; j < NX; j++) a += sm.Lxx[j * NX + lane] * sm.sRb[j];
        sm.sT1[lane] = a;
    }
    warp_sync();
    if (lane < NX) {
        T a = sm.sPv[lane];
        for (int c = 0; c <= lane; c++) a += sm.Lxx[lane * NX + c] * sm.sT1[c];
        sm.sT2[lane] = a;
    }
    warp_sync();
}

// One forward sweep (affine: FINAL=false, corrector: FINAL=true).  Returns the largest
// admissible step and, for the affine sweep, the three sums that give mu_aff(alpha).
template <int NX, int NU, typename T, bool FINAL>
MPCB_DEV void forward_sweep(const Params &P, QpSmem<NX, NU, T> &sm, T *__restrict__ ws, const T *__restrict__ Xi,
                            const T *__restrict__ Ui, const T *__restrict__ yref, int yps, T sigmu, T &amax, T &s1,
                            T &s2)
{
    using L = Layout<NX, NU>;
    constexpr int NZ = L::NZ;
    const int lane = lane_id();
    const int N = P.N;
    T amin = T(HUGE_VAL), acc1 = T(0), acc2 = T(0);
    if (lane < NX) sm.sDx[lane] = T(0);
    warp_sync();
    for (int k = 0; k < N; k++) {
        T *wk = ws + (size_t)k * L::STAGE;
        T *wk1 = wk + L::STAGE;
        load_BAt<NX, NU, T>(sm, wk);
        for (int idx = lane; idx < NU * L::NZP; idx += 32) sm.Lu[idx] = wk[L::O_LU + idx];
        if (FINAL)
            for (int idx = lane; idx < NX * NX; idx += 32) sm.Lxx[idx] = wk1[L::O_LXX + idx];
        T invd[NU];
        MPCB_UNROLL
        for (int c = 0; c < NU; c++) invd[c] = wk[L::O_INVD + c];
        warp_sync();
        // du = -Luu^{-T} (lvec + Lxu' dx)
        T yy = T(0);
        if (lane < NU) {
            T a = wk[L::O_LVEC + lane];
            for (int i = 0; i < NX; i++) a += sm.Lu[lane * L::NZP + NU + i] * sm.sDx[i];
            yy = -a;
        }
        T du = T(0);
        MPCB_UNROLL
        for (int i = NU - 1; i >= 0; i--) {
            const T dui = warp_shfl(yy, i) * invd[i];
            if (lane == i) du = dui;
            if (lane < i) yy -= sm.Lu[lane * L::NZP + i] * dui;
        }
        T dz = T(0);
        if (lane < NU) dz = du;
        else if (lane < NZ) dz = sm.sDx[lane - NU];
        if (lane < NZ) {
            wk[(FINAL ? L::O_DZ : L::O_DZA) + lane] = dz;
            sm.sDz[lane] = dz;
        }
        warp_sync();
        // dx_{k+1} = rb_k + [B A] dz_k
        T dxn = T(0);
        if (lane < NX) {
            dxn = wk[L::O_RB + lane];
            for (int j = 0; j < NZ; j++) dxn += sm.BAt[j * L::LDB + lane] * sm.sDz[j];
            sm.sRb[lane] = dxn;  // input of apply_P below
            if (FINAL) sm.sPv[lane] = wk1[L::O_PV + lane];
        }
        // step-length bookkeeping for the bounded components of stage k
        {
            const StageVar<T> sv = stage_var<NX, NU, T>(P, k, lane, Xi, Ui, yref, yps);
            if (sv.hasb) {
                const T z = wk[L::O_Z + lane], tl = wk[L::O_TL + lane], tu = wk[L::O_TU + lane];
                const T ll = wk[L::O_LL + lane], lu = wk[L::O_LUP + lane];
                T rml = ll * tl, rmu = lu * tu;
                if (FINAL) {
                    const T dza = wk[L::O_DZA + lane];
                    const BoxStep<T> a = box_step(z, dza, sv.lb, sv.ub, tl, tu, ll, lu, rml, rmu);
                    rml += a.dll * a.dtl - sigmu;
                    rmu += a.dlu * a.dtu - sigmu;
                }
                const BoxStep<T> b = box_step(z, dz, sv.lb, sv.ub, tl, tu, ll, lu, rml, rmu);
                amin = fmin(amin, fmin(fmin(ratio(tl, b.dtl), ratio(tu, b.dtu)), fmin(ratio(ll, b.dll), ratio(lu, b.dlu))));
                if (!FINAL) {
                    acc1 += ll * b.dtl + tl * b.dll + lu * b.dtu + tu * b.dlu;
                    acc2 += b.dll * b.dtl + b.dlu * b.dtu;
                }
            }
        }
        warp_sync();
        if (FINAL) {
            // dpi_{k+1} = P_{k+1} dx_{k+1} + p_{k+1}
            apply_P<NX, NU, T>(sm);
            if (lane < NX) wk1[L::O_DPI + lane] = sm.sT2[lane];
        }
        if (lane < NX) sm.sDx[lane] = dxn;
        warp_sync();
    }
    // terminal stage: dz_N = [0; dx_N], no bounds
    {
        T *wN = ws + (size_t)N * L::STAGE;
        if (lane < NZ) wN[(FINAL ? L::O_DZ : L::O_DZA) + lane] = (lane < NU) ? T(0) : sm.sDx[lane - NU];
    }
    warp_sync();
    amax = warp_min(amin);
    s1 = warp_sum(acc1);
    s2 = warp_sum(acc2);
}

// Forward substitution with the first NU columns of L_k held row-wise in registers:
// on return lanes c < NU hold lvec_c = (Luu^{-1} l_u)_c and lanes NU.. hold p_k.
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

// The whole QP solve for one instance.  On return the persistent iterate Xi/Ui has taken
// the full step (FIXED_STEP, step length 1.0: acados_ocp_blasterModel.json globalization /
// nlp_solver_step_length).  Returns the status; *iters_out = IPM iterations.
template <int NX, int NU, typename T>
MPCB_DEV int qp_solve_warp(const Params &P, QpSmem<NX, NU, T> &sm, T *__restrict__ ws, T *__restrict__ Xi,
                           T *__restrict__ Ui, const T *__restrict__ x0, const T *__restrict__ yref, int yps,
                           int *iters_out)
{
    using L = Layout<NX, NU>;
    constexpr int NZ = L::NZ;
    const int lane = lane_id();
    const int N = P.N;
    const T thr0 = (T)P.ipm_thr0, mu0 = (T)P.ipm_mu0;
    const T nb = (T)(2 * NU * N + 2 * NX * (N - 1));

    for (int idx = lane; idx < NX * NX; idx += 32) sm.Lxx[idx] = T(0);

    // ---------------- cold start [upstream D8]: z = 0 (dx_0 pinned), pi = 0, t >= thr0, lam = mu0/t
    T eg = T(0), eb = T(0), ed = T(0);
    for (int k = 0; k <= N; k++) {
        T *wk = ws + (size_t)k * L::STAGE;
        const StageVar<T> sv = stage_var<NX, NU, T>(P, k, lane, Xi, Ui, yref, yps);
        T z = T(0);
        if (k == 0 && lane >= NU && lane < NZ) z = x0[lane - NU] - Xi[lane - NU];
        T tl = T(1), tu = T(1), ll = T(0), lu = T(0);
        if (sv.hasb) {
            tl = fmax(z - sv.lb, thr0);
            tu = fmax(sv.ub - z, thr0);
            ll = mu0 / tl;
            lu = mu0 / tu;
            ed = fmax(ed, fmax(fabs(z - sv.lb - tl), fabs(sv.ub - z - tu)));
        }
        if (sv.var) eg = fmax(eg, fabs(sv.H0 * z + sv.g - ll + lu));
        if (lane < NZ) {
            wk[L::O_Z + lane] = z; wk[L::O_TL + lane] = tl; wk[L::O_TU + lane] = tu;
            wk[L::O_LL + lane] = ll; wk[L::O_LUP + lane] = lu;
        }
        if (lane < NX) {
            wk[L::O_PI + lane] = T(0);
            if (k < N) {
                T rb = wk[L::O_B + lane];
                if (k == 0)
                    for (int i = 0; i < NX; i++) rb += wk[L::O_BAT + (NU + i) * NX + lane] * (x0[i] - Xi[i]);
                eb = fmax(eb, fabs(rb));
            }
        }
    }
    T est_g = warp_max(eg), est_b = warp_max(eb), est_d = warp_max(ed);
    T comp = mu0, mu = mu0;
    int status = ST_MAXITER, it = 0;
    warp_sync();

    for (it = 0; it < P.ipm_max_iter; it++) {
        if (!(est_g == est_g) || !(est_b == est_b) || !(mu == mu)) { status = ST_NAN; break; }
        if (est_g <= (T)P.tol_stat && est_b <= (T)P.tol_eq && est_d <= (T)P.tol_ineq && comp <= (T)P.tol_comp) {
            status = ST_OK;
            break;
        }
        int fail = 0;
        // ================= S1: backward sweep -- residuals, factorisation, affine RHS
        {
            T *wN = ws + (size_t)N * L::STAGE;
            const StageVar<T> sv = stage_var<NX, NU, T>(P, N, lane, Xi, Ui, yref, yps);
            T q = T(0);
            if (lane >= NU && lane < NZ) {
                const int i = lane - NU;
                q = sv.H0 * wN[L::O_Z + lane] + sv.g - wN[L::O_PI + i];
                const T d = sqrt(sv.H0);
                for (int c = 0; c < NX; c++)
                    if (c <= i) sm.Lxx[i * NX + c] = (c == i) ? d : T(0);
                sm.sPv[i] = q;
                wN[L::O_PV + i] = q;
            }
            if (lane < NZ) wN[L::O_Q + lane] = q;
            warp_sync();
            for (int idx = lane; idx < NX * NX; idx += 32) wN[L::O_LXX + idx] = sm.Lxx[idx];
        }
        for (int k = N - 1; k >= 0; k--) {
            T *wk = ws + (size_t)k * L::STAGE;
            T *wk1 = wk + L::STAGE;
            load_BAt<NX, NU, T>(sm, wk);
            if (lane < NX) sm.sPi[lane] = wk1[L::O_PI + lane];
            const T zj = (lane < NZ) ? wk[L::O_Z + lane] : T(0);
            if (lane < NZ) sm.sZ[lane] = zj;
            warp_sync();
            T brow[NX];
            MPCB_UNROLL
            for (int c = 0; c < NX; c++) brow[c] = (lane < NZ) ? sm.BAt[lane * L::LDB + c] : T(0);
            const StageVar<T> sv = stage_var<NX, NU, T>(P, k, lane, Xi, Ui, yref, yps);
            T Hd = sv.H0, q = T(0);
            {
                T rg = T(0);
                T ll = T(0), lu = T(0), tl = T(1), tu = T(1);
                if (sv.hasb) {
                    tl = wk[L::O_TL + lane]; tu = wk[L::O_TU + lane]; ll = wk[L::O_LL + lane]; lu = wk[L::O_LUP + lane];
                }
                if (sv.var) {
                    rg = sv.H0 * zj + sv.g - ll + lu;
                    MPCB_UNROLL
                    for (int c = 0; c < NX; c++) rg += brow[c] * sm.sPi[c];
                    if (lane >= NU) rg -= wk[L::O_PI + lane - NU];
                }
                q = rg;
                if (sv.hasb) {
                    const T rdl = zj - sv.lb - tl, rdu = sv.ub - zj - tu;
                    Hd += ll / tl + lu / tu;
                    // affine right-hand side: rm = lam*t
                    q += (ll * tl + ll * rdl) / tl - (lu * tu + lu * rdu) / tu;
                }
            }
            if (lane < NX) {
                T rb = wk[L::O_B + lane] - wk1[L::O_Z + NU + lane];
                for (int j = 0; j < NZ; j++) rb += sm.BAt[j * L::LDB + lane] * sm.sZ[j];
                wk[L::O_RB + lane] = rb;
                sm.sRb[lane] = rb;
            }
            if (lane < NZ) wk[L::O_Q + lane] = q;
            warp_sync();
            // t2 = P_{k+1} rb + p_{k+1}
            apply_P<NX, NU, T>(sm);
            // W = [B A]' Lxx_{k+1}   (row `lane`)
            T w[NX];
            MPCB_UNROLL
            for (int c = 0; c < NX; c++) {
                T a = T(0);
                MPCB_UNROLL
                for (int j = c; j < NX; j++) a += brow[j] * sm.Lxx[j * NX + c];
                w[c] = a;
            }
            const T dsq = sqrt(Hd);
            T Lu[NU], invd[NU];
            // Householder LQ of [diag(dsq) | W]; at stage 0 only the u-block is needed
            const int jend = (k == 0) ? NU : NZ;
            MPCB_UNROLL
            for (int j = 0; j < NU; j++) {
                T *vr = sm.vrow[j & 1];
                if (lane == j) {
                    MPCB_UNROLL
                    for (int c = 0; c < NX; c++) vr[c] = w[c];
                }
                warp_sync();
                T dot = T(0);
                MPCB_UNROLL
                for (int c = 0; c < NX; c++) dot += vr[c] * w[c];
                const T s2v = warp_shfl(Hd, j) + warp_shfl(dot, j);
                const T rs = fast_rsqrt(s2v);
                const T sig = s2v * rs;
                const T v0 = warp_shfl(dsq, j) + sig;
                const T beta = rs * fast_rcp(v0);
                if (!(s2v > T(0)) || !(s2v < T(HUGE_VAL))) fail = 1;
                const T f = (lane > j && lane < NZ) ? beta * dot : T(0);
                MPCB_UNROLL
                for (int c = 0; c < NX; c++) w[c] -= f * vr[c];
                Lu[j] = (lane == j) ? sig : f * v0;
                invd[j] = rs;
            }
            for (int j = NU; j < jend; j++) {
                T *vr = sm.vrow[j & 1];
                if (lane == j) {
                    MPCB_UNROLL
                    for (int c = 0; c < NX; c++) vr[c] = w[c];
                }
                warp_sync();
                T dot = T(0);
                MPCB_UNROLL
                for (int c = 0; c < NX; c++) dot += vr[c] * w[c];
                const T s2v = warp_shfl(Hd, j) + warp_shfl(dot, j);
                const T rs = fast_rsqrt(s2v);
                const T sig = s2v * rs;
                const T v0 = warp_shfl(dsq, j) + sig;
                const T beta = rs * fast_rcp(v0);
                if (!(s2v > T(0)) || !(s2v < T(HUGE_VAL))) fail = 1;
                const T f = (lane > j && lane < NZ) ? beta * dot : T(0);
                MPCB_UNROLL
                for (int c = 0; c < NX; c++) w[c] -= f * vr[c];
                if (lane >= j && lane < NZ) sm.Lxx[(lane - NU) * NX + (j - NU)] = (lane == j) ? sig : f * v0;
            }
            // affine backward vectors: l = q + [B A]' t2
            T l = q;
            MPCB_UNROLL
            for (int c = 0; c < NX; c++) l += brow[c] * sm.sT2[c];
            l = fwd_subst<NU, T>(l, Lu, invd, NZ);
            if (lane < NU) wk[L::O_LVEC + lane] = l;
            else if (lane < NZ) { wk[L::O_PV + lane - NU] = l; sm.sPv[lane - NU] = l; }
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
        }
        if (warp_or(fail)) { status = ST_QPFAIL; break; }

        // ================= S2: forward sweep, affine step
        T a_aff, s1, s2;
        forward_sweep<NX, NU, T, false>(P, sm, ws, Xi, Ui, yref, yps, T(0), a_aff, s1, s2);
        a_aff = fmin(T(1), a_aff);
        const T mu_aff = (mu * nb + a_aff * s1 + a_aff * a_aff * s2) / nb;
        T sigma = mu_aff / mu;
        sigma = sigma * sigma * sigma;
        const T sigmu = sigma * mu;

        // ================= S3: backward sweep for the corrector increment (delta form)
        if (lane < NX) sm.sPv[lane] = T(0);
        warp_sync();
        for (int k = N - 1; k >= 0; k--) {
            T *wk = ws + (size_t)k * L::STAGE;
            load_BAt<NX, NU, T>(sm, wk);
            T Lu[NU], invd[NU];
            MPCB_UNROLL
            for (int c = 0; c < NU; c++) {
                Lu[c] = (lane < NZ) ? wk[L::O_LU + c * L::NZP + lane] : T(0);
                invd[c] = wk[L::O_INVD + c];
            }
            const StageVar<T> sv = stage_var<NX, NU, T>(P, k, lane, Xi, Ui, yref, yps);
            T l = T(0);
            if (sv.hasb) {
                const T z = wk[L::O_Z + lane], tl = wk[L::O_TL + lane], tu = wk[L::O_TU + lane];
                const T ll = wk[L::O_LL + lane], lu = wk[L::O_LUP + lane];
                const BoxStep<T> a = box_step(z, wk[L::O_DZA + lane], sv.lb, sv.ub, tl, tu, ll, lu, ll * tl, lu * tu);
                l = (a.dll * a.dtl - sigmu) / tl - (a.dlu * a.dtu - sigmu) / tu;
            }
            warp_sync();
            if (lane < NZ) {
                for (int c = 0; c < NX; c++) l += sm.BAt[lane * L::LDB + c] * sm.sPv[c];
            }
            l = fwd_subst<NU, T>(l, Lu, invd, NZ);
            warp_sync();
            if (lane < NU) wk[L::O_LVEC + lane] += l;
            else if (lane < NZ) { wk[L::O_PV + lane - NU] += l; sm.sPv[lane - NU] = l; }
            warp_sync();
        }

        // ================= S4: forward sweep, full predictor-corrector step
        T a_max, d1, d2;
        forward_sweep<NX, NU, T, true>(P, sm, ws, Xi, Ui, yref, yps, sigmu, a_max, d1, d2);
        const T alpha = fmin(T(1), fmax(T(0.995), T(1) - mu_aff) * a_max);

        // ================= S5: take the step
        T cmax = T(0), msum = T(0);
        for (int k = 0; k <= N; k++) {
            T *wk = ws + (size_t)k * L::STAGE;
            const StageVar<T> sv = stage_var<NX, NU, T>(P, k, lane, Xi, Ui, yref, yps);
            if (sv.var) {
                const T z = wk[L::O_Z + lane], dz = wk[L::O_DZ + lane];
                if (sv.hasb) {
                    T tl = wk[L::O_TL + lane], tu = wk[L::O_TU + lane], ll = wk[L::O_LL + lane], lu = wk[L::O_LUP + lane];
                    const BoxStep<T> a = box_step(z, wk[L::O_DZA + lane], sv.lb, sv.ub, tl, tu, ll, lu, ll * tl, lu * tu);
                    const BoxStep<T> b = box_step(z, dz, sv.lb, sv.ub, tl, tu, ll, lu, ll * tl + a.dll * a.dtl - sigmu,
                                                  lu * tu + a.dlu * a.dtu - sigmu);
                    tl += alpha * b.dtl; tu += alpha * b.dtu; ll += alpha * b.dll; lu += alpha * b.dlu;
                    wk[L::O_TL + lane] = tl; wk[L::O_TU + lane] = tu; wk[L::O_LL + lane] = ll; wk[L::O_LUP + lane] = lu;
                    cmax = fmax(cmax, fmax(ll * tl, lu * tu));
                    msum += ll * tl + lu * tu;
                }
                wk[L::O_Z + lane] = z + alpha * dz;
            }
            if (k >= 1 && lane < NX) wk[L::O_PI + lane] += alpha * wk[L::O_DPI + lane];
        }
        comp = warp_max(cmax);
        mu = warp_sum(msum) / nb;
        est_g *= (T(1) - alpha);
        est_b *= (T(1) - alpha);
        est_d *= (T(1) - alpha);
        warp_sync();
        if (!(alpha >= (T)P.alpha_min)) {
            status = (alpha == alpha) ? ST_MINSTEP : ST_NAN;
            it++;
            break;
        }
    }

    // ---------------- RTI update: X += dx, U += du (full step)
    warp_sync();
    for (int k = 0; k <= N; k++) {
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
