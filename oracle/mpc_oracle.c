/* C oracle of the BLASTER quadrotor SQP-RTI solve.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED (solver semantics): the reference delegates this arithmetic to
 * acados/HPIPM/BLASFEO/CasADi, none of which is vendored or installable here; the
 * reference ships no tests and no recorded outputs.  This file restates
 *   dynamics   /root/reference/src/scripts/blastermodel.py:93-167,171-210
 *   cost       blastermodel.py:228-257        bounds  blastermodel.py:261-270
 *   options    blastermodel.py:272-287 + acados_ocp_blasterModel.json solver_options
 *   loop body  src/scripts/simulation_blaster.py:56-105
 * with the stage-structured algorithm class acados/HPIPM use (ERK4 + forward
 * sensitivities, Gauss-Newton LINEAR_LS, Mehrotra IPM on a square-root Riccati
 * factorisation).  Conventions that are upstream knowledge are tagged [upstream Dn]
 * (SURVEY.md appendix D).  It is validated against the NumPy oracle (dense-KKT IPM,
 * oracle/blaster_oracle.py), and is what bench.py times as the CPU baseline ("port").
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this.
 * Build: oracle/build_oracle.py  ->  oracle/_build/libmpc_oracle.so
 */
#include <math.h>
#include <stdio.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int variant; /* 17 or 12 */
    int N;
    double dt, mass, J[9], Jinv[9], l_x, l_y, c;
    double Q[17], R[6], Qt[17], lbx[17], ubx[17], lbu[6], ubu[6];
    int ipm_max_iter;
    int rg_mode; /* 0 stationarity residual tracked analytically, 1 recomputed from pi and used in the test, 2 recomputed for the
                    Newton right-hand side while the test uses the extrapolated norms, confirmed explicitly before success (the product's rule) */
    int ric_alg; /* 0 Cholesky of Hd+WW', 1 Householder LQ of [sqrt(Hd) | W] (the checker), 3 experiment: classical recursion on P_k with the pinned directions kept in a side list (ric_factor_split) */
    double ipm_mu0, ipm_thr0, tol_stat, tol_eq, tol_ineq, tol_comp, alpha_min;
    int strict; /* reference semantics (mpcb_config.strict_reference): explicit residual norms in the stopping test (rg_mode 1
                   is then implied), no early exit on diverging multipliers, last iterate applied on max-iter */
    int itref;  /* iterative-refinement steps on the corrector solve (experiment / strict mode; 0 = none) */
    double mixed_mu; /* experiment (tools/mixed_precision_viability.py): > 0 = the stage factorisation (W = [B A]'L, Gram matrix,
                        Cholesky) is done in FP32 while mixed_mu < mu <= mu0; FP64 Householder LQ otherwise */
} orc_problem;

#define GRAV 9.81 /* blastermodel.py:93 */

/* ORC_DEBUG=1 prints the IPM residual history (read once, not in the timed loop) */
static int orc_debug(void)
{
    static int flag = -1;
    if (flag < 0) flag = getenv("ORC_DEBUG") != NULL;
    return flag;
}

/* experiment ric_alg = 3 (tools/split_factor_viability.py): thresholds (squared column norm of a carried column; barrier
 * diagonal of a state) and a histogram of the list length per stage factorisation */
static double orc_huge_tau(void)
{
    static double tau = -1.0;
    if (tau < 0.0) { const char *e = getenv("ORC_HUGE_TAU"); tau = e ? atof(e) : 1e6; }
    return tau;
}
static double orc_huge_tau_d(void)
{
    static double tau = -1.0;
    if (tau < 0.0) { const char *e = getenv("ORC_HUGE_TAU_D"); tau = e ? atof(e) : 1e6; }
    return tau;
}
static double orc_split_mu(void) /* the experiment is used while mu > this (0 = every iteration) */
{
    static double v = -1.0;
    if (v < 0.0) { const char *e = getenv("ORC_SPLIT_MU"); v = e ? atof(e) : 0.0; }
    return v;
}
static int orc_huge_hmax(void) /* capacity of the list of carried columns */
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("ORC_HUGE_HMAX"); v = e ? atoi(e) : 34; if (v > 34) v = 34; }
    return v;
}
static long orc_split_iters[2]; /* interior-point iterations factorised by the LQ / by the split recursion */
static void orc_split_iter(int split)
{
#ifdef _OPENMP
#pragma omp atomic
#endif
    orc_split_iters[split]++;
}
void orc_split_iterations(long *out, int reset)
{
    out[0] = orc_split_iters[0]; out[1] = orc_split_iters[1];
    if (reset) orc_split_iters[0] = orc_split_iters[1] = 0;
}
static long orc_huge_hist[36];
static void orc_huge_count(int nh)
{
#ifdef _OPENMP
#pragma omp atomic
#endif
    orc_huge_hist[nh < 35 ? nh : 35]++;
}
void orc_huge_histogram(long *out, int reset)
{
    for (int i = 0; i < 36; i++) { out[i] = orc_huge_hist[i]; if (reset) orc_huge_hist[i] = 0; }
}

/* blastermodel.py:124,162-167: full 17-state model */
static void orc_f17(const orc_problem *P, const double *x, const double *u, const double *p, double *xd)
{
    const double sf = sin(x[3]), cf = cos(x[3]), st = sin(x[4]), ct = cos(x[4]), sp = sin(x[5]), cp = cos(x[5]);
    const double s1 = sin(x[12]), c1 = cos(x[12]), s2 = sin(x[13]), c2 = cos(x[13]);
    const double *v = x + 6, *om = x + 9;
    const double Tb = p[24], Ts = u[0] + u[1] + u[2] + u[3];
    const double R[3][3] = {{cp * ct, cp * st * sf - sp * cf, cp * st * cf + sp * sf},
                            {sp * ct, sp * st * sf + cp * cf, sp * st * cf - cp * sf},
                            {-st, ct * sf, ct * cf}};
    /* body-frame force: e3*sum(T) + R_gimbal e3 * T_blast, R_gimbal = Ry(a1) Rx(a2) (:143-160) */
    const double w[3] = {Tb * s1 * c2, -Tb * s2, Ts + Tb * c1 * c2};
    const double tt = st / ct;
    const double ed[3] = {om[0] + sf * tt * om[1] + cf * tt * om[2], cf * om[1] - sf * om[2], (sf * om[1] + cf * om[2]) / ct};
    for (int i = 0; i < 3; i++) xd[i] = v[i];
    for (int i = 0; i < 3; i++) xd[3 + i] = ed[i];
    for (int i = 0; i < 3; i++) xd[6 + i] = (R[i][0] * w[0] + R[i][1] * w[1] + R[i][2] * w[2]) / P->mass;
    xd[8] -= GRAV;
    /* moments :95-101 */
    const double M[3] = {(u[1] + u[3] - u[0] - u[2]) * P->l_y, (-u[0] - u[3] + u[1] + u[2]) * P->l_x, (-u[0] - u[1] + u[2] + u[3]) * P->c};
    double Jo[3], cr[3];
    for (int i = 0; i < 3; i++) Jo[i] = P->J[3 * i] * om[0] + P->J[3 * i + 1] * om[1] + P->J[3 * i + 2] * om[2];
    cr[0] = om[1] * Jo[2] - om[2] * Jo[1];
    cr[1] = om[2] * Jo[0] - om[0] * Jo[2];
    cr[2] = om[0] * Jo[1] - om[1] * Jo[0];
    for (int i = 0; i < 3; i++)
        xd[9 + i] = P->Jinv[3 * i] * (M[0] - cr[0]) + P->Jinv[3 * i + 1] * (M[1] - cr[1]) + P->Jinv[3 * i + 2] * (M[2] - cr[2]);
    xd[12] = u[4];
    xd[13] = u[5];
    /* pocdot = J_p v + J_euler etadot + J_angles alphadot (:165), params column-major (:203-210) */
    for (int i = 0; i < 3; i++)
        xd[14 + i] = p[15 + i] * v[0] + p[18 + i] * v[1] + p[21 + i] * v[2]
                     + p[6 + i] * ed[0] + p[9 + i] * ed[1] + p[12 + i] * ed[2]
                     + p[0 + i] * u[4] + p[3 + i] * u[5];
}

static void orc_jac17(const orc_problem *P, const double *x, const double *u, const double *p, double fx[17][17], double fu[17][6])
{
    const double sf = sin(x[3]), cf = cos(x[3]), st = sin(x[4]), ct = cos(x[4]), sp = sin(x[5]), cp = cos(x[5]);
    const double s1 = sin(x[12]), c1 = cos(x[12]), s2 = sin(x[13]), c2 = cos(x[13]);
    const double *om = x + 9;
    const double Tb = p[24], Ts = u[0] + u[1] + u[2] + u[3], m = 1.0 / P->mass;
    const double R[3][3] = {{cp * ct, cp * st * sf - sp * cf, cp * st * cf + sp * sf},
                            {sp * ct, sp * st * sf + cp * cf, sp * st * cf - cp * sf},
                            {-st, ct * sf, ct * cf}};
    const double dRth[3][3] = {{-cp * st, cp * ct * sf, cp * ct * cf}, {-sp * st, sp * ct * sf, sp * ct * cf}, {-ct, -st * sf, -st * cf}};
    const double w[3] = {Tb * s1 * c2, -Tb * s2, Ts + Tb * c1 * c2};
    const double tt = st / ct;
    const double E[3][3] = {{1.0, sf * tt, cf * tt}, {0.0, cf, -sf}, {0.0, sf / ct, cf / ct}};
    const double ed[3] = {om[0] + sf * tt * om[1] + cf * tt * om[2], cf * om[1] - sf * om[2], (sf * om[1] + cf * om[2]) / ct};
    const double dEf[3] = {tt * ed[1], -ed[2] * ct, ed[1] / ct}; /* d(etadot)/dphi   */
    const double dEt[3] = {ed[2] / ct, 0.0, ed[2] * tt};         /* d(etadot)/dtheta */
    memset(fx, 0, sizeof(double) * 17 * 17);
    memset(fu, 0, sizeof(double) * 17 * 6);
    for (int i = 0; i < 3; i++) fx[i][6 + i] = 1.0;
    for (int i = 0; i < 3; i++) {
        fx[3 + i][3] = dEf[i];
        fx[3 + i][4] = dEt[i];
        for (int j = 0; j < 3; j++) fx[3 + i][9 + j] = E[i][j];
    }
    double Rw[3];
    for (int i = 0; i < 3; i++) Rw[i] = R[i][0] * w[0] + R[i][1] * w[1] + R[i][2] * w[2];
    const double dg1[3] = {c1 * c2, 0.0, -s1 * c2}, dg2[3] = {-s1 * s2, -c2, -c1 * s2};
    for (int i = 0; i < 3; i++) {
        fx[6 + i][3] = m * (R[i][2] * w[1] - R[i][1] * w[2]);
        fx[6 + i][4] = m * (dRth[i][0] * w[0] + dRth[i][1] * w[1] + dRth[i][2] * w[2]);
        fx[6 + i][12] = m * Tb * (R[i][0] * dg1[0] + R[i][1] * dg1[1] + R[i][2] * dg1[2]);
        fx[6 + i][13] = m * Tb * (R[i][0] * dg2[0] + R[i][1] * dg2[1] + R[i][2] * dg2[2]);
        for (int j = 0; j < 4; j++) fu[6 + i][j] = m * R[i][2];
    }
    fx[6][5] = -m * Rw[1];
    fx[7][5] = m * Rw[0];
    /* omegadot: -Jinv ([om]x J - [J om]x) */
    double Jo[3], D[3][3];
    for (int i = 0; i < 3; i++) Jo[i] = P->J[3 * i] * om[0] + P->J[3 * i + 1] * om[1] + P->J[3 * i + 2] * om[2];
    const double So[3][3] = {{0, -om[2], om[1]}, {om[2], 0, -om[0]}, {-om[1], om[0], 0}};
    const double SJ[3][3] = {{0, -Jo[2], Jo[1]}, {Jo[2], 0, -Jo[0]}, {-Jo[1], Jo[0], 0}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double a = -SJ[i][j];
            for (int k = 0; k < 3; k++) a += So[i][k] * P->J[3 * k + j];
            D[i][j] = a;
        }
    const double G[3][4] = {{-P->l_y, P->l_y, -P->l_y, P->l_y}, {-P->l_x, P->l_x, P->l_x, -P->l_x}, {-P->c, -P->c, P->c, P->c}};
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
            double a = 0;
            for (int k = 0; k < 3; k++) a += P->Jinv[3 * i + k] * D[k][j];
            fx[9 + i][9 + j] = -a;
        }
        for (int j = 0; j < 4; j++) {
            double a = 0;
            for (int k = 0; k < 3; k++) a += P->Jinv[3 * i + k] * G[k][j];
            fu[9 + i][j] = a;
        }
    }
    fu[12][4] = 1.0;
    fu[13][5] = 1.0;
    for (int i = 0; i < 3; i++) {
        const double Je[3] = {p[6 + i], p[9 + i], p[12 + i]};
        for (int j = 0; j < 3; j++) fx[14 + i][6 + j] = p[15 + i + 3 * j];
        fx[14 + i][3] = Je[0] * dEf[0] + Je[1] * dEf[1] + Je[2] * dEf[2];
        fx[14 + i][4] = Je[0] * dEt[0] + Je[1] * dEt[1] + Je[2] * dEt[2];
        for (int j = 0; j < 3; j++) fx[14 + i][9 + j] = Je[0] * E[0][j] + Je[1] * E[1][j] + Je[2] * E[2][j];
        fu[14 + i][4] = p[0 + i];
        fu[14 + i][5] = p[3 + i];
    }
}

/* QUAT13: the 12-state quadrotor with its attitude as a unit quaternion,
 * x = [p(3), q(w,x,y,z), v(3), omega(3)], u = [T0..T3]; the quaternion algebra is that of the
 * reference's utils/MathUtils.py (quatMultiplication :5-23, quat2Rot :41-54), which no model of
 * the reference uses.  PARITY UNPINNED for the model; tested against the Euler model.
 * Same padded signature as orc_f17 / orc_jac17 (the first 13 states / 4 inputs are used). */
static void orc_f13(const orc_problem *P, const double *x, const double *u, const double *p, double *xd)
{
    const double w = x[3], qx = x[4], qy = x[5], qz = x[6];
    const double *v = x + 7, *om = x + 10;
    const double F = (u[0] + u[1] + u[2] + u[3] + p[24]) / P->mass;
    for (int i = 0; i < 17; i++) xd[i] = 0.0;
    for (int i = 0; i < 3; i++) xd[i] = v[i];
    /* qdot = 1/2 q (x) [0, omega] */
    xd[3] = 0.5 * (-qx * om[0] - qy * om[1] - qz * om[2]);
    xd[4] = 0.5 * (w * om[0] + qy * om[2] - qz * om[1]);
    xd[5] = 0.5 * (w * om[1] - qx * om[2] + qz * om[0]);
    xd[6] = 0.5 * (w * om[2] + qx * om[1] - qy * om[0]);
    /* vdot = R(q) e3 (sum T + T_blast)/M + g: third column of quat2Rot */
    xd[7] = F * 2.0 * (qx * qz + w * qy);
    xd[8] = F * 2.0 * (qy * qz - w * qx);
    xd[9] = F * (2.0 * (w * w + qz * qz) - 1.0) - GRAV;
    const double M[3] = {(u[1] + u[3] - u[0] - u[2]) * P->l_y, (-u[0] - u[3] + u[1] + u[2]) * P->l_x, (-u[0] - u[1] + u[2] + u[3]) * P->c};
    double Jo[3], cr[3];
    for (int i = 0; i < 3; i++) Jo[i] = P->J[3 * i] * om[0] + P->J[3 * i + 1] * om[1] + P->J[3 * i + 2] * om[2];
    cr[0] = om[1] * Jo[2] - om[2] * Jo[1];
    cr[1] = om[2] * Jo[0] - om[0] * Jo[2];
    cr[2] = om[0] * Jo[1] - om[1] * Jo[0];
    for (int i = 0; i < 3; i++)
        xd[10 + i] = P->Jinv[3 * i] * (M[0] - cr[0]) + P->Jinv[3 * i + 1] * (M[1] - cr[1]) + P->Jinv[3 * i + 2] * (M[2] - cr[2]);
}

static void orc_jac13(const orc_problem *P, const double *x, const double *u, const double *p, double fx[17][17], double fu[17][6])
{
    const double w = x[3], qx = x[4], qy = x[5], qz = x[6];
    const double *om = x + 10;
    const double a = om[0], b = om[1], c = om[2];
    const double m = 1.0 / P->mass, F = (u[0] + u[1] + u[2] + u[3] + p[24]) * m;
    memset(fx, 0, sizeof(double) * 17 * 17);
    memset(fu, 0, sizeof(double) * 17 * 6);
    for (int i = 0; i < 3; i++) fx[i][7 + i] = 1.0;
    const double Om[4][4] = {{0, -a, -b, -c}, {a, 0, c, -b}, {b, -c, 0, a}, {c, b, -a, 0}};
    const double Xi[4][3] = {{-qx, -qy, -qz}, {w, -qz, qy}, {qz, w, -qx}, {-qy, qx, w}};
    for (int i = 0; i < 4; i++) {
        for (int j = 0; j < 4; j++) fx[3 + i][3 + j] = 0.5 * Om[i][j];
        for (int j = 0; j < 3; j++) fx[3 + i][10 + j] = 0.5 * Xi[i][j];
    }
    const double dr3[3][4] = {{qy, qz, w, qx}, {-qx, -w, qz, qy}, {2 * w, 0, 0, 2 * qz}};
    const double r3[3] = {2.0 * (qx * qz + w * qy), 2.0 * (qy * qz - w * qx), 2.0 * (w * w + qz * qz) - 1.0};
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 4; j++) fx[7 + i][3 + j] = 2.0 * F * dr3[i][j];
        for (int j = 0; j < 4; j++) fu[7 + i][j] = m * r3[i];
    }
    double Jo[3], D[3][3];
    for (int i = 0; i < 3; i++) Jo[i] = P->J[3 * i] * om[0] + P->J[3 * i + 1] * om[1] + P->J[3 * i + 2] * om[2];
    const double So[3][3] = {{0, -om[2], om[1]}, {om[2], 0, -om[0]}, {-om[1], om[0], 0}};
    const double SJ[3][3] = {{0, -Jo[2], Jo[1]}, {Jo[2], 0, -Jo[0]}, {-Jo[1], Jo[0], 0}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = -SJ[i][j];
            for (int k = 0; k < 3; k++) s += So[i][k] * P->J[3 * k + j];
            D[i][j] = s;
        }
    const double G[3][4] = {{-P->l_y, P->l_y, -P->l_y, P->l_y}, {-P->l_x, P->l_x, P->l_x, -P->l_x}, {-P->c, -P->c, P->c, P->c}};
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += P->Jinv[3 * i + k] * D[k][j];
            fx[10 + i][10 + j] = -s;
        }
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += P->Jinv[3 * i + k] * G[k][j];
            fu[10 + i][j] = s;
        }
    }
}

#define MODEL_F orc_f17
#define MODEL_JAC orc_jac17
#define NX 17
#define NU 6
#define SFX(n) b17_##n
#include "mpc_oracle_body.h"
#undef NX
#undef NU
#undef SFX

#define NX 12
#define NU 4
#define SFX(n) q12_##n
#include "mpc_oracle_body.h"
#undef NX
#undef NU
#undef SFX
#undef MODEL_F
#undef MODEL_JAC

#define MODEL_F orc_f13
#define MODEL_JAC orc_jac13
#define NX 13
#define NU 4
#define SFX(n) q13_##n
#include "mpc_oracle_body.h"
#undef NX
#undef NU
#undef SFX
#undef MODEL_F
#undef MODEL_JAC

#define ORC_NX(P) ((P)->variant == 17 ? 17 : (P)->variant == 13 ? 13 : 12)
#define ORC_NU(P) ((P)->variant == 17 ? 6 : 4)

/* ------------------------------------------------------------------ exported API */
void orc_f(const orc_problem *P, const double *x, const double *u, const double *p, double *xd)
{
    if (P->variant == 17) b17_f(P, x, u, p, xd); else if (P->variant == 13) q13_f(P, x, u, p, xd); else q12_f(P, x, u, p, xd);
}

void orc_rk4_sens(const orc_problem *P, const double *x, const double *u, const double *p, double *xn, double *BAt)
{
    if (P->variant == 17) b17_rk4_sens(P, x, u, p, xn, BAt);
    else if (P->variant == 13) q13_rk4_sens(P, x, u, p, xn, BAt);
    else q12_rk4_sens(P, x, u, p, xn, BAt);
}

void orc_plant_step_batch(const orc_problem *P, const double *x, const double *u, const double *p, int p_per_inst,
                          double *xn, int B, int nthreads)
{
    const int nx = ORC_NX(P), nu = ORC_NU(P);
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int i = 0; i < B; i++) {
        const double *pi = p + (p_per_inst ? (size_t)i * 25 : 0);
        if (P->variant == 17) b17_plant_step(P, x + (size_t)i * nx, u + (size_t)i * nu, pi, xn + (size_t)i * nx);
        else if (P->variant == 13) q13_plant_step(P, x + (size_t)i * nx, u + (size_t)i * nu, pi, xn + (size_t)i * nx);
        else q12_plant_step(P, x + (size_t)i * nx, u + (size_t)i * nu, pi, xn + (size_t)i * nx);
    }
}

/* One SQP_RTI iteration for B independent instances (OpenMP over instances).
 * X[B,(N+1),nx], U[B,N,nu] persistent iterate (in/out); x0[B,nx];
 * yref: yref_mode 0 -> [ny] shared, 1 -> [B,ny], 2 -> [B,N+1,ny];
 * p:    p_mode    0 -> [25] shared, 1 -> [B,25], 2 -> [B,N,25].
 * status[B], iters[B].  Returns 0 or -1 on allocation failure. */
int orc_rti_solve_batch(const orc_problem *P, double *X, double *U, const double *x0, const double *yref, int yref_mode,
                        const double *p, int p_mode, int32_t *status, int32_t *iters, int B, int nthreads)
{
    const int nx = ORC_NX(P), nu = ORC_NU(P), ny = nx + nu, N = P->N;
    const size_t wsz = P->variant == 17 ? b17_ws_doubles(N) : P->variant == 13 ? q13_ws_doubles(N) : q12_ws_doubles(N);
    int fail = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        double *ws = (double *)malloc(wsz * sizeof(double));
        if (!ws) {
#pragma omp atomic write
            fail = 1;
        }
#pragma omp for schedule(dynamic, 4)
        for (int i = 0; i < B; i++) {
            if (!ws) continue;
            const double *yr = yref + (yref_mode == 0 ? 0 : yref_mode == 1 ? (size_t)i * ny : (size_t)i * (N + 1) * ny);
            const double *pp = p + (p_mode == 0 ? 0 : p_mode == 1 ? (size_t)i * 25 : (size_t)i * N * 25);
            int it = 0, st;
            if (P->variant == 17)
                st = b17_rti_solve(P, X + (size_t)i * (N + 1) * nx, U + (size_t)i * N * nu, x0 + (size_t)i * nx, yr, yref_mode == 2,
                                   pp, p_mode == 2, ws, &it);
            else if (P->variant == 13)
                st = q13_rti_solve(P, X + (size_t)i * (N + 1) * nx, U + (size_t)i * N * nu, x0 + (size_t)i * nx, yr, yref_mode == 2,
                                   pp, p_mode == 2, ws, &it);
            else
                st = q12_rti_solve(P, X + (size_t)i * (N + 1) * nx, U + (size_t)i * N * nu, x0 + (size_t)i * nx, yr, yref_mode == 2,
                                   pp, p_mode == 2, ws, &it);
            status[i] = st;
            iters[i] = it;
        }
        free(ws);
    }
    return fail ? -1 : 0;
}

/* SQP to convergence for B independent instances (see sqp_solve in mpc_oracle_body.h).  tol[4] = {stat, eq, ineq, comp};
 * status[B], sqp_iters[B], qp_iters[B], res[B,4]. */
int orc_sqp_solve_batch(const orc_problem *P, double *X, double *U, const double *x0, const double *yref, int yref_mode,
                        const double *p, int p_mode, int max_iter, const double *tol, int32_t *status, int32_t *sqp_iters,
                        int32_t *qp_iters, double *res, int B, int nthreads)
{
    const int nx = ORC_NX(P), nu = ORC_NU(P), ny = nx + nu, N = P->N;
    const size_t wsz = P->variant == 17 ? b17_ws_doubles(N) : P->variant == 13 ? q13_ws_doubles(N) : q12_ws_doubles(N);
    int fail = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        double *ws = (double *)malloc(wsz * sizeof(double));
        if (!ws) {
#pragma omp atomic write
            fail = 1;
        }
#pragma omp for schedule(dynamic, 1)
        for (int i = 0; i < B; i++) {
            if (!ws) continue;
            const double *yr = yref + (yref_mode == 0 ? 0 : yref_mode == 1 ? (size_t)i * ny : (size_t)i * (N + 1) * ny);
            const double *pp = p + (p_mode == 0 ? 0 : p_mode == 1 ? (size_t)i * 25 : (size_t)i * N * 25);
            int si = 0, qi = 0, st;
            double *Xi = X + (size_t)i * (N + 1) * nx, *Ui = U + (size_t)i * N * nu;
            if (P->variant == 17)
                st = b17_sqp_solve(P, Xi, Ui, x0 + (size_t)i * nx, yr, yref_mode == 2, pp, p_mode == 2, max_iter, tol, ws, &si, &qi, res + (size_t)i * 4);
            else if (P->variant == 13)
                st = q13_sqp_solve(P, Xi, Ui, x0 + (size_t)i * nx, yr, yref_mode == 2, pp, p_mode == 2, max_iter, tol, ws, &si, &qi, res + (size_t)i * 4);
            else
                st = q12_sqp_solve(P, Xi, Ui, x0 + (size_t)i * nx, yr, yref_mode == 2, pp, p_mode == 2, max_iter, tol, ws, &si, &qi, res + (size_t)i * 4);
            status[i] = st; sqp_iters[i] = si; qp_iters[i] = qi;
        }
        free(ws);
    }
    return fail ? -1 : 0;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

size_t orc_problem_size(void) { return sizeof(orc_problem); }
