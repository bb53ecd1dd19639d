// Jet point-of-contact (POC) and its Jacobians: the parameter generator that feeds p[0:24] of
// the OCP (SURVEY 8f "next" row 2).  ONE INSTANCE (vehicle pose) PER THREAD -- the work is a few
// thousand scalar flops, so a batch is embarrassingly parallel and needs no shared memory.
//
// What the reference runs here (all under src/scripts/): Jacobian_POC_Solver.py --
//   setInitConditions :153-175 + htm.py:7-36   nozzle pose -> jet initial state [p0, v0]
//   _createIntegrator :59-99                   p' = v, v' = -M_c v + g, ERK4, 10 steps over T
//   _solveRootFindingProblem :115-152          Newton on z(T) = 0, forward-difference slope
//   solveJacobians :234-300                    forward differences, eps = 1e-6
// on the CPU, once, before the control loop (simulation_blaster.py:37-39), although the result
// depends on the pose.  Two modes:
//   POC_MODE_REFERENCE  the reference's algorithm step for step (parity mode);
//   POC_MODE_ANALYTIC   the linear jet ODE in closed form, Newton to machine precision, Jacobians
//                       by the implicit-function theorem: what one would refresh every control
//                       step per vehicle on the device.
#pragma once
#include "mpcb_common.cuh"

namespace mpcb {

constexpr int POC_MODE_REFERENCE = 0, POC_MODE_ANALYTIC = 1;
constexpr double kPocG = 9.81;

struct PocOut {
    double poc[3];
    double J[3][8];  // columns: alpha1, alpha2 | phi, theta, psi | x, y, z
    double t_flight;
    int status;      // 0 ok, 1 NaN, 2 iteration cap
};

MPCB_DEV void mat3_mul(const double (&A)[3][3], const double (&B)[3][3], double (&C)[3][3])
{
    MPCB_UNROLL
    for (int i = 0; i < 3; i++)
        MPCB_UNROLL
        for (int j = 0; j < 3; j++) C[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
}
MPCB_DEV void mat3_vec(const double (&A)[3][3], const double (&v)[3], double (&r)[3])
{
    MPCB_UNROLL
    for (int i = 0; i < 3; i++) r[i] = A[i][0] * v[0] + A[i][1] * v[1] + A[i][2] * v[2];
}

// htm.py:31-36: scipy from_euler('zyx', [psi, theta, phi]) = extrinsic z, y, x = Rx(phi) Ry(theta) Rz(psi).
// d = 0: the rotation; d = 1, 2, 3: its derivative with respect to phi, theta, psi.
MPCB_DEV void poc_rot_w_b(double phi, double theta, double psi, int d, double (&R)[3][3])
{
    double sf, cf, st, ct, sp, cp;
    sincos_(phi, &sf, &cf); sincos_(theta, &st, &ct); sincos_(psi, &sp, &cp);
    double Rx[3][3] = {{1, 0, 0}, {0, cf, -sf}, {0, sf, cf}};
    double Ry[3][3] = {{ct, 0, st}, {0, 1, 0}, {-st, 0, ct}};
    double Rz[3][3] = {{cp, -sp, 0}, {sp, cp, 0}, {0, 0, 1}};
    if (d == 1) { const double D[3][3] = {{0, 0, 0}, {0, -sf, -cf}, {0, cf, -sf}}; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rx[i][j] = D[i][j]; }
    if (d == 2) { const double D[3][3] = {{-st, 0, ct}, {0, 0, 0}, {-ct, 0, -st}}; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Ry[i][j] = D[i][j]; }
    if (d == 3) { const double D[3][3] = {{-sp, -cp, 0}, {cp, -sp, 0}, {0, 0, 0}}; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rz[i][j] = D[i][j]; }
    double t[3][3];
    mat3_mul(Rx, Ry, t);
    mat3_mul(t, Rz, R);
}

// htm.py:7-29: T_b_s2 = hbs1 hs1s2 hs2n.  tb = its translation, zb = third column of its rotation
// (the jet leaves along -z of the nozzle frame); d = 1, 2: derivatives w.r.t. alpha1, alpha2.
MPCB_DEV void poc_nozzle(double a1, double a2, int d, double (&tb)[3], double (&zb)[3])
{
    double s1, c1, s2, c2;
    sincos_(a1, &s1, &c1); sincos_(a2, &s2, &c2);
    if (d == 0) {
        tb[0] = 0.01672 + 0.0425 + c1 * (-0.05322) + s1 * (-0.15946);
        tb[1] = 0.0;
        tb[2] = -0.22937 - s1 * (-0.05322) + c1 * (-0.15946);
        zb[0] = s1 * c2; zb[1] = s2; zb[2] = c1 * c2;
    } else if (d == 1) {
        tb[0] = -s1 * (-0.05322) + c1 * (-0.15946);
        tb[1] = 0.0;
        tb[2] = -c1 * (-0.05322) - s1 * (-0.15946);
        zb[0] = c1 * c2; zb[1] = 0.0; zb[2] = -s1 * c2;
    } else {
        tb[0] = tb[1] = tb[2] = 0.0;
        zb[0] = -s1 * s2; zb[1] = c2; zb[2] = -c1 * s2;
    }
}

// Jacobian_POC_Solver.py:153-175: x0 = [translation of T_w_b T_b_s2, R (0, 0, -V)]
MPCB_DEV void poc_init_conditions(const double *e, const double *m, const double *pos, double V, double (&x0)[6])
{
    double R[3][3], tb[3], zb[3], r[3];
    poc_rot_w_b(e[0], e[1], e[2], 0, R);
    poc_nozzle(m[0], m[1], 0, tb, zb);
    mat3_vec(R, tb, r);
    MPCB_UNROLL
    for (int i = 0; i < 3; i++) x0[i] = pos[i] + r[i];
    mat3_vec(R, zb, r);
    MPCB_UNROLL
    for (int i = 0; i < 3; i++) x0[3 + i] = -V * r[i];
}

// Jacobian_POC_Solver.py:77-99: explicit RK4, 10 uniform steps, of p' = v, v' = -c v + g
MPCB_DEV void poc_flight_rk4(const double (&x0)[6], double T, double c, double (&x)[6])
{
    const double h = T / 10.0;
    MPCB_UNROLL
    for (int i = 0; i < 6; i++) x[i] = x0[i];
    for (int s = 0; s < 10; s++) {
        double k1[6], k2[6], k3[6], k4[6], y[6];
        auto f = [&](const double (&z)[6], double (&k)[6]) {
            k[0] = z[3]; k[1] = z[4]; k[2] = z[5];
            k[3] = -c * z[3]; k[4] = -c * z[4]; k[5] = -c * z[5] - kPocG;
        };
        f(x, k1);
        for (int i = 0; i < 6; i++) y[i] = x[i] + 0.5 * h * k1[i];
        f(y, k2);
        for (int i = 0; i < 6; i++) y[i] = x[i] + 0.5 * h * k2[i];
        f(y, k3);
        for (int i = 0; i < 6; i++) y[i] = x[i] + h * k3[i];
        f(y, k4);
        for (int i = 0; i < 6; i++) x[i] = x[i] + h / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
    }
}

// Jacobian_POC_Solver.py:115-152.  Returns the time of flight; *status: 0 ok, 1 NaN, 2 cap (the
// reference loops without a cap).
MPCB_DEV double poc_time_of_flight(const double (&x0)[6], double c, int *status)
{
    double T = 0.1, x[6];
    for (int it = 0; it < 100; it++) {
        poc_flight_rk4(x0, T, c, x);
        const double f = x[2];
        poc_flight_rk4(x0, T + 1e-5, c, x);
        const double fp = (x[2] - f) / 1e-5;
        T = T - f / fp;
        if (T < 0) T = -T;
        poc_flight_rk4(x0, T, c, x);
        if (!(x[2] == x[2])) { *status = 1; return T; }
        if (fabs(x[2]) <= 1e-3) return T;
    }
    *status = 2;
    return T;
}

MPCB_DEV void poc_reference_point(const double *e, const double *m, const double *pos, double V, double c, double (&poc)[3], double *tf,
                                  int *status)
{
    double x0[6], x[6];
    poc_init_conditions(e, m, pos, V, x0);
    const double T = poc_time_of_flight(x0, c, status);
    poc_flight_rk4(x0, T, c, x);
    poc[0] = x[0]; poc[1] = x[1]; poc[2] = x[2];
    if (tf) *tf = T;
}

// Jacobian_POC_Solver.py:234-300 (each coordinate perturbed on its own: the reference's list call pattern)
MPCB_DEV void poc_reference(const double *e, const double *m, const double *pos, double V, double c, PocOut &o)
{
    const double eps = 1e-6;
    o.status = 0;
    poc_reference_point(e, m, pos, V, c, o.poc, &o.t_flight, &o.status);
    for (int col = 0; col < 8; col++) {
        double ee[3] = {e[0], e[1], e[2]}, mm[2] = {m[0], m[1]}, pp[3] = {pos[0], pos[1], pos[2]}, q[3];
        if (col < 2) mm[col] = mm[col] + eps;
        else if (col < 5) ee[col - 2] = ee[col - 2] + eps;
        else pp[col - 5] = pp[col - 5] + eps;
        poc_reference_point(ee, mm, pp, V, c, q, nullptr, &o.status);
        for (int i = 0; i < 3; i++) o.J[i][col] = (q[i] - o.poc[i]) / eps;
    }
}

// Exact counterpart: v(T) = vinf + (v0 - vinf) e^{-cT}, p(T) = p0 + vinf T + (v0 - vinf) k(T), k = (1 - e^{-cT})/c;
// z(T*) = 0 by Newton with the analytic slope v_z; d POC/d theta = dp/dtheta|_T - v (dz/dtheta|_T) / v_z.
MPCB_DEV void poc_analytic(const double *e, const double *m, const double *pos, double V, double c, PocOut &o)
{
    double x0[6];
    poc_init_conditions(e, m, pos, V, x0);
    const double vinf_z = -kPocG / c;
    double T = x0[2] / fmax(-x0[5], 1e-9);
    T = fmax(T, 1e-6);
    double k = 0, ez = 1;
    o.status = 2;
    for (int it = 0; it < 50; it++) {
        ez = exp(-c * T);
        k = -expm1(-c * T) / c;
        const double z = x0[2] + vinf_z * T + (x0[5] - vinf_z) * k;
        const double vz = vinf_z + (x0[5] - vinf_z) * ez;
        const double dT = -z / vz;
        T += dT;
        if (!(T == T)) { o.status = 1; break; }
        if (fabs(dT) <= 1e-15 * fmax(T, 1.0)) { o.status = 0; break; }
    }
    ez = exp(-c * T);
    k = -expm1(-c * T) / c;
    double v[3];
    const double vinf[3] = {0.0, 0.0, vinf_z};
    for (int i = 0; i < 3; i++) {
        o.poc[i] = x0[i] + vinf[i] * T + (x0[3 + i] - vinf[i]) * k;
        v[i] = vinf[i] + (x0[3 + i] - vinf[i]) * ez;
    }
    o.t_flight = T;
    double R[3][3], tb[3], zb[3];
    poc_rot_w_b(e[0], e[1], e[2], 0, R);
    poc_nozzle(m[0], m[1], 0, tb, zb);
    for (int col = 0; col < 8; col++) {
        double dp0[3], dv0[3], r[3];
        if (col < 2) {  // nozzle angles: p0 = pos + R tb, v0 = -V R zb
            double dtb[3], dzb[3];
            poc_nozzle(m[0], m[1], col + 1, dtb, dzb);
            mat3_vec(R, dtb, dp0);
            mat3_vec(R, dzb, r);
            for (int i = 0; i < 3; i++) dv0[i] = -V * r[i];
        } else if (col < 5) {  // Euler angles
            double dR[3][3];
            poc_rot_w_b(e[0], e[1], e[2], col - 1, dR);
            mat3_vec(dR, tb, dp0);
            mat3_vec(dR, zb, r);
            for (int i = 0; i < 3; i++) dv0[i] = -V * r[i];
        } else {
            for (int i = 0; i < 3; i++) { dp0[i] = (i == col - 5) ? 1.0 : 0.0; dv0[i] = 0.0; }
        }
        double dp[3];
        for (int i = 0; i < 3; i++) dp[i] = dp0[i] + dv0[i] * k;
        const double s = dp[2] / v[2];
        for (int i = 0; i < 3; i++) o.J[i][col] = dp[i] - v[i] * s;
    }
}

// Pack as the reference does before ocp_solver.set(k, 'p', ...) (simulation_blaster.py:67):
// [vec_colmajor(J_mot 3x2), vec_colmajor(J_eul 3x3), vec_colmajor(J_pos 3x3), T_blast]
MPCB_DEV void poc_pack_params(const PocOut &o, double T_blast, double *p25)
{
    int n = 0;
    for (int col = 0; col < 8; col++)
        for (int i = 0; i < 3; i++) p25[n++] = o.J[i][col];
    p25[24] = T_blast;
}

}  // namespace mpcb
