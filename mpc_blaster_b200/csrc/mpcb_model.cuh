// BLASTER quadrotor dynamics and the Jacobian-times-column product, as inlined device
// functions.  Restates the CasADi model of reference blastermodel.py:93-167 (what acados
// evaluates through CasADi-generated C as f and its forward VDE).
//
// State   x = [p(3), phi, theta, psi, v(3), omega(3), alpha1, alpha2, poc(3)]  (blastermodel.py:171-183)
// Input   u = [T0..T3, alpha1_dot, alpha2_dot]                                  (:184-190)
// Params  p = [vec(J_angles 3x2), vec(J_euler 3x3), vec(J_p 3x3), T_blast]      (:203-210, column-major)
// QUAD12 = states 0..11 and inputs 0..3 with the gimbal frozen at alpha1 = alpha2 = 0.
// QUAT13 = QUAD12 with the attitude as a unit quaternion, x = [p(3), q(w,x,y,z), v(3), omega(3)]:
//          qdot = 1/2 q (x) [0, omega], vdot = R(q) e3 (sum T + T_blast)/M + g with the quaternion algebra of
//          reference utils/MathUtils.py (quat_mul / quat_to_rot below).  SURVEY 8a row A9; the reference has no
//          such model, so it is tested against the Euler model and the CPU checkers, not against reference outputs.
#pragma once
#include "mpcb_common.cuh"

namespace mpcb {

constexpr double kGravity = 9.81;  // blastermodel.py:93

// Rotation / quaternion helpers of reference utils/MathUtils.py (Hamilton product :5-23,
// unit inverse :25-39, quat2Rot :41-54) as device functions, q = [w, x, y, z].  The
// reference imports them but never calls them (the OCP uses Euler angles); they are used
// by the command-mapping epilogue (mavros_blaster_sim.py:91-100) and unit-tested.
template <typename T>
MPCB_HD void quat_mul(const T *a, const T *b, T *c)
{
    c[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    c[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    c[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    c[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
template <typename T>
MPCB_HD void quat_inv_unit(const T *q, T *r)
{
    r[0] = q[0]; r[1] = -q[1]; r[2] = -q[2]; r[3] = -q[3];
}
template <typename T>
MPCB_HD void quat_to_rot(const T *e, T *R /* 3x3 row-major */)
{
    R[0] = 2 * (e[0] * e[0] + e[1] * e[1]) - 1; R[1] = 2 * (e[1] * e[2] - e[0] * e[3]); R[2] = 2 * (e[1] * e[3] + e[0] * e[2]);
    R[3] = 2 * (e[1] * e[2] + e[0] * e[3]); R[4] = 2 * (e[0] * e[0] + e[2] * e[2]) - 1; R[5] = 2 * (e[2] * e[3] - e[0] * e[1]);
    R[6] = 2 * (e[1] * e[3] - e[0] * e[2]); R[7] = 2 * (e[2] * e[3] + e[0] * e[1]); R[8] = 2 * (e[0] * e[0] + e[3] * e[3]) - 1;
}
// Euler (phi,theta,psi) -> quaternion of R = Rz(psi) Ry(theta) Rx(phi) (blastermodel.py:122)
template <typename T>
MPCB_HD void euler_to_quat(T phi, T th, T psi, T *q)
{
    T cf = cos(phi / 2), sf = sin(phi / 2), ct = cos(th / 2), st = sin(th / 2), cp = cos(psi / 2), sp = sin(psi / 2);
    q[0] = cp * ct * cf + sp * st * sf;
    q[1] = cp * ct * sf - sp * st * cf;
    q[2] = cp * st * cf + sp * ct * sf;
    q[3] = sp * ct * cf - cp * st * sf;
}

// Trig of the five angles the model needs.  Each of lanes 0..4 evaluates one sincos
// and the results are broadcast, so the warp pays for one sincos instead of five.
template <int NX, typename T>
struct Trig {
    T sf, cf, st, ct, sp, cp, s1, c1, s2, c2;
};

template <int NX, typename T>
MPCB_DEV void eval_trig(const T *xs, Trig<NX, T> &g)
{
    if constexpr (NX == 13) { (void)xs; (void)g; return; }  // quaternion attitude: no trigonometry at all
    const int lane = lane_id();
    T ang = xs[3];
    if (lane == 1) ang = xs[4];
    if (lane == 2) ang = xs[5];
    if (NX == 17) {
        if (lane == 3) ang = xs[12 < NX ? 12 : 0];
        if (lane == 4) ang = xs[13 < NX ? 13 : 0];
    }
    T s, c;
    sincos_(ang, &s, &c);
    g.sf = warp_shfl(s, 0); g.cf = warp_shfl(c, 0);
    g.st = warp_shfl(s, 1); g.ct = warp_shfl(c, 1);
    g.sp = warp_shfl(s, 2); g.cp = warp_shfl(c, 2);
    if (NX == 17) {
        g.s1 = warp_shfl(s, 3); g.c1 = warp_shfl(c, 3);
        g.s2 = warp_shfl(s, 4); g.c2 = warp_shfl(c, 4);
    } else {
        g.s1 = 0; g.c1 = 1; g.s2 = 0; g.c2 = 1;
    }
}

// Thread-local variant (plant step: one instance per thread).
template <int NX, typename T>
MPCB_DEV void eval_trig_local(const T *xs, Trig<NX, T> &g)
{
    if constexpr (NX == 13) { (void)xs; (void)g; return; }
    sincos_(xs[3], &g.sf, &g.cf);
    sincos_(xs[4], &g.st, &g.ct);
    sincos_(xs[5], &g.sp, &g.cp);
    if (NX == 17) {
        sincos_(xs[12 < NX ? 12 : 0], &g.s1, &g.c1);
        sincos_(xs[13 < NX ? 13 : 0], &g.s2, &g.c2);
    } else {
        g.s1 = 0; g.c1 = 1; g.s2 = 0; g.c2 = 1;
    }
}

// Quantities shared by f and its Jacobian at one RK stage point.
template <int NX, typename T>
struct StagePoint {
    T R[3][3];   // Rz(psi) Ry(theta) Rx(phi)
    T w[3];      // body-frame force  e3*sum(T) + R_gimbal e3 * T_blast
    T E[3][3];   // Euler-rate map inv(R_to_omega)
    T ed[3];     // Euler rates
    T tt, ict;
};

template <int NX, typename T>
MPCB_DEV void eval_point(const Trig<NX, T> &g, const T *xs, T Tsum, T Tb, StagePoint<NX, T> &s)
{
    if constexpr (NX == 13) {
        // QUAT13 needs R(q) e3 (third column of quat_to_rot) and the body-z force
        const T w = xs[3], qx = xs[4], qy = xs[5], qz = xs[6];
        s.R[0][2] = 2 * (qx * qz + w * qy);
        s.R[1][2] = 2 * (qy * qz - w * qx);
        s.R[2][2] = 2 * (w * w + qz * qz) - 1;
        s.w[2] = Tsum + Tb;
        return;
    }
    s.R[0][0] = g.cp * g.ct; s.R[0][1] = g.cp * g.st * g.sf - g.sp * g.cf; s.R[0][2] = g.cp * g.st * g.cf + g.sp * g.sf;
    s.R[1][0] = g.sp * g.ct; s.R[1][1] = g.sp * g.st * g.sf + g.cp * g.cf; s.R[1][2] = g.sp * g.st * g.cf - g.cp * g.sf;
    s.R[2][0] = -g.st;       s.R[2][1] = g.ct * g.sf;                      s.R[2][2] = g.ct * g.cf;
    s.w[0] = Tb * g.s1 * g.c2;
    s.w[1] = -Tb * g.s2;
    s.w[2] = Tsum + Tb * g.c1 * g.c2;
    s.ict = T(1) / g.ct;
    s.tt = g.st * s.ict;
    s.E[0][0] = 1; s.E[0][1] = g.sf * s.tt;  s.E[0][2] = g.cf * s.tt;
    s.E[1][0] = 0; s.E[1][1] = g.cf;         s.E[1][2] = -g.sf;
    s.E[2][0] = 0; s.E[2][1] = g.sf * s.ict; s.E[2][2] = g.cf * s.ict;
    const T *om = xs + 9;
    s.ed[0] = om[0] + s.E[0][1] * om[1] + s.E[0][2] * om[2];
    s.ed[1] = s.E[1][1] * om[1] + s.E[1][2] * om[2];
    s.ed[2] = s.E[2][1] * om[1] + s.E[2][2] * om[2];
}

// xdot = f(x,u,p)  (blastermodel.py:124,162-167).  `u` has NU entries, `pp` 25.
template <int NX, int NU, typename T>
MPCB_DEV void eval_f(const Params &P, const StagePoint<NX, T> &s, const T *xs, const T *u, const T *pp, T *xd)
{
    if constexpr (NX == 13) {
        const T w = xs[3], qx = xs[4], qy = xs[5], qz = xs[6];
        const T *v = xs + 7, *om = xs + 10;
        xd[0] = v[0]; xd[1] = v[1]; xd[2] = v[2];
        // qdot = 1/2 q (x) [0, omega]  (quat_mul with a zero scalar part)
        xd[3] = T(0.5) * (-qx * om[0] - qy * om[1] - qz * om[2]);
        xd[4] = T(0.5) * (w * om[0] + qy * om[2] - qz * om[1]);
        xd[5] = T(0.5) * (w * om[1] - qx * om[2] + qz * om[0]);
        xd[6] = T(0.5) * (w * om[2] + qx * om[1] - qy * om[0]);
        const T F = (T)P.inv_mass * s.w[2];
        xd[7] = F * s.R[0][2]; xd[8] = F * s.R[1][2]; xd[9] = F * s.R[2][2] - (T)kGravity;
        T Jo[3], cr[3];
        MPCB_UNROLL
        for (int i = 0; i < 3; i++) Jo[i] = (T)P.J[3 * i] * om[0] + (T)P.J[3 * i + 1] * om[1] + (T)P.J[3 * i + 2] * om[2];
        cr[0] = om[1] * Jo[2] - om[2] * Jo[1];
        cr[1] = om[2] * Jo[0] - om[0] * Jo[2];
        cr[2] = om[0] * Jo[1] - om[1] * Jo[0];
        MPCB_UNROLL
        for (int i = 0; i < 3; i++) {
            T a = -((T)P.Jinv[3 * i] * cr[0] + (T)P.Jinv[3 * i + 1] * cr[1] + (T)P.Jinv[3 * i + 2] * cr[2]);
            MPCB_UNROLL
            for (int j = 0; j < 4; j++) a += (T)P.JinvG[4 * i + j] * u[j];
            xd[(10 + i) < NX ? 10 + i : 0] = a;
        }
        (void)pp;
        return;
    }
    const T *v = xs + 6, *om = xs + 9;
    xd[0] = v[0]; xd[1] = v[1]; xd[2] = v[2];
    xd[3] = s.ed[0]; xd[4] = s.ed[1]; xd[5] = s.ed[2];
    const T m = (T)P.inv_mass;
    MPCB_UNROLL
    for (int i = 0; i < 3; i++) xd[6 + i] = m * (s.R[i][0] * s.w[0] + s.R[i][1] * s.w[1] + s.R[i][2] * s.w[2]);
    xd[8] -= (T)kGravity;
    T Jo[3], cr[3];
    MPCB_UNROLL
    for (int i = 0; i < 3; i++) Jo[i] = (T)P.J[3 * i] * om[0] + (T)P.J[3 * i + 1] * om[1] + (T)P.J[3 * i + 2] * om[2];
    cr[0] = om[1] * Jo[2] - om[2] * Jo[1];
    cr[1] = om[2] * Jo[0] - om[0] * Jo[2];
    cr[2] = om[0] * Jo[1] - om[1] * Jo[0];
    MPCB_UNROLL
    for (int i = 0; i < 3; i++) {
        T a = -((T)P.Jinv[3 * i] * cr[0] + (T)P.Jinv[3 * i + 1] * cr[1] + (T)P.Jinv[3 * i + 2] * cr[2]);
        MPCB_UNROLL
        for (int j = 0; j < 4; j++) a += (T)P.JinvG[4 * i + j] * u[j];
        xd[9 + i] = a;
    }
    if (NX == 17) {
        xd[12 < NX ? 12 : 0] = u[4 < NU ? 4 : 0];
        xd[13 < NX ? 13 : 0] = u[5 < NU ? 5 : 0];
        MPCB_UNROLL
        for (int i = 0; i < 3; i++)
            xd[(14 + i) < NX ? 14 + i : 0] = pp[15 + i] * v[0] + pp[18 + i] * v[1] + pp[21 + i] * v[2]
                                             + pp[6 + i] * s.ed[0] + pp[9 + i] * s.ed[1] + pp[12 + i] * s.ed[2]
                                             + pp[0 + i] * u[4 < NU ? 4 : 0] + pp[3 + i] * u[5 < NU ? 5 : 0];
    }
}

// K = (df/dx) S + (df/du) e_ucol for one sensitivity column S (NX entries);
// ucol = index of the input this column differentiates w.r.t., or -1 for a state column.
// Uses the sparsity of df/dx: rows p <- v, rows alpha <- u only,
// poc rows re-use the Euler-rate rows.
template <int NX, int NU, typename T>
MPCB_DEV void eval_jac_col(const Params &P, const Trig<NX, T> &g, const StagePoint<NX, T> &s, const T *xs, const T *pp,
                           T Tb, const T *S, int ucol, T *K)
{
    if constexpr (NX == 13) {
        const T w = xs[3], qx = xs[4], qy = xs[5], qz = xs[6];
        const T *om = xs + 10;
        const T *Sq = S + 3, *So = S + 10;
        K[0] = S[7]; K[1] = S[8]; K[2] = S[9];
        // d(qdot) = 1/2 (dq (x) [0, omega] + q (x) [0, domega])
        K[3] = T(0.5) * (-om[0] * Sq[1] - om[1] * Sq[2] - om[2] * Sq[3] - qx * So[0] - qy * So[1] - qz * So[2]);
        K[4] = T(0.5) * (om[0] * Sq[0] + om[2] * Sq[2] - om[1] * Sq[3] + w * So[0] - qz * So[1] + qy * So[2]);
        K[5] = T(0.5) * (om[1] * Sq[0] - om[2] * Sq[1] + om[0] * Sq[3] + qz * So[0] + w * So[1] - qx * So[2]);
        K[6] = T(0.5) * (om[2] * Sq[0] + om[1] * Sq[1] - om[0] * Sq[2] - qy * So[0] + qx * So[1] + w * So[2]);
        // d(vdot) = F d(R e3)/dq dq + (R e3 / M) dT
        const T m = (T)P.inv_mass, F2 = 2 * m * s.w[2];
        const T thr = (ucol >= 0 && ucol < 4) ? m : T(0);
        K[7] = F2 * (qy * Sq[0] + qz * Sq[1] + w * Sq[2] + qx * Sq[3]) + thr * s.R[0][2];
        K[8] = F2 * (-qx * Sq[0] - w * Sq[1] + qz * Sq[2] + qy * Sq[3]) + thr * s.R[1][2];
        K[9] = F2 * (2 * w * Sq[0] + 2 * qz * Sq[3]) + thr * s.R[2][2];
        T Jo[3], JS[3], c1v[3], c2v[3];
        MPCB_UNROLL
        for (int i = 0; i < 3; i++) {
            Jo[i] = (T)P.J[3 * i] * om[0] + (T)P.J[3 * i + 1] * om[1] + (T)P.J[3 * i + 2] * om[2];
            JS[i] = (T)P.J[3 * i] * So[0] + (T)P.J[3 * i + 1] * So[1] + (T)P.J[3 * i + 2] * So[2];
        }
        c1v[0] = om[1] * JS[2] - om[2] * JS[1]; c1v[1] = om[2] * JS[0] - om[0] * JS[2]; c1v[2] = om[0] * JS[1] - om[1] * JS[0];
        c2v[0] = Jo[1] * So[2] - Jo[2] * So[1]; c2v[1] = Jo[2] * So[0] - Jo[0] * So[2]; c2v[2] = Jo[0] * So[1] - Jo[1] * So[0];
        MPCB_UNROLL
        for (int i = 0; i < 3; i++) {
            T a = T(0);
            MPCB_UNROLL
            for (int k = 0; k < 3; k++) a -= (T)P.Jinv[3 * i + k] * (c1v[k] - c2v[k]);
            if (ucol >= 0 && ucol < 4) a += (T)P.JinvG[4 * i + ucol];
            K[(10 + i) < NX ? 10 + i : 0] = a;
        }
        (void)g; (void)pp; (void)Tb;
        return;
    }
    const T *om = xs + 9;
    const T m = (T)P.inv_mass;
    K[0] = S[6]; K[1] = S[7]; K[2] = S[8];
    // Euler rates: d/dphi, d/dtheta, d/domega
    {
        const T dEf0 = s.tt * s.ed[1], dEf1 = -s.ed[2] * g.ct, dEf2 = s.ed[1] * s.ict;
        const T dEt0 = s.ed[2] * s.ict, dEt2 = s.ed[2] * s.tt;
        K[3] = dEf0 * S[3] + dEt0 * S[4] + S[9] + s.E[0][1] * S[10] + s.E[0][2] * S[11];
        K[4] = dEf1 * S[3] + s.E[1][1] * S[10] + s.E[1][2] * S[11];
        K[5] = dEf2 * S[3] + dEt2 * S[4] + s.E[2][1] * S[10] + s.E[2][2] * S[11];
    }
    // vdot = m R w + g
    {
        T Rw[3];
        MPCB_UNROLL
        for (int i = 0; i < 3; i++) Rw[i] = s.R[i][0] * s.w[0] + s.R[i][1] * s.w[1] + s.R[i][2] * s.w[2];
        const T dRth[3][3] = {{-g.cp * g.st, g.cp * g.ct * g.sf, g.cp * g.ct * g.cf},
                              {-g.sp * g.st, g.sp * g.ct * g.sf, g.sp * g.ct * g.cf},
                              {-g.ct, -g.st * g.sf, -g.st * g.cf}};
        const T dpsi[3] = {-Rw[1], Rw[0], T(0)};
        const T thr = (ucol >= 0 && ucol < 4) ? m : T(0);
        MPCB_UNROLL
        for (int i = 0; i < 3; i++) {
            T a = m * (s.R[i][2] * s.w[1] - s.R[i][1] * s.w[2]) * S[3];
            a += m * (dRth[i][0] * s.w[0] + dRth[i][1] * s.w[1] + dRth[i][2] * s.w[2]) * S[4];
            a += m * dpsi[i] * S[5];
            if (NX == 17) {
                const T dg1[3] = {g.c1 * g.c2, T(0), -g.s1 * g.c2};
                const T dg2[3] = {-g.s1 * g.s2, -g.c2, -g.c1 * g.s2};
                a += m * Tb * (s.R[i][0] * dg1[0] + s.R[i][2] * dg1[2]) * S[12 < NX ? 12 : 0];
                a += m * Tb * (s.R[i][0] * dg2[0] + s.R[i][1] * dg2[1] + s.R[i][2] * dg2[2]) * S[13 < NX ? 13 : 0];
            }
            a += thr * s.R[i][2];
            K[6 + i] = a;
        }
    }
    // omegadot = Jinv (G T - om x J om):  d/dom = -Jinv ([om]x J - [J om]x)
    {
        T Jo[3], JS[3], c1v[3], c2v[3];
        const T *So = S + 9;
        MPCB_UNROLL
        for (int i = 0; i < 3; i++) {
            Jo[i] = (T)P.J[3 * i] * om[0] + (T)P.J[3 * i + 1] * om[1] + (T)P.J[3 * i + 2] * om[2];
            JS[i] = (T)P.J[3 * i] * So[0] + (T)P.J[3 * i + 1] * So[1] + (T)P.J[3 * i + 2] * So[2];
        }
        // om x (J S) and (J om) x S
        c1v[0] = om[1] * JS[2] - om[2] * JS[1]; c1v[1] = om[2] * JS[0] - om[0] * JS[2]; c1v[2] = om[0] * JS[1] - om[1] * JS[0];
        c2v[0] = Jo[1] * So[2] - Jo[2] * So[1]; c2v[1] = Jo[2] * So[0] - Jo[0] * So[2]; c2v[2] = Jo[0] * So[1] - Jo[1] * So[0];
        MPCB_UNROLL
        for (int i = 0; i < 3; i++) {
            T a = T(0);
            MPCB_UNROLL
            for (int k = 0; k < 3; k++) a -= (T)P.Jinv[3 * i + k] * (c1v[k] - c2v[k]);
            if (ucol >= 0 && ucol < 4) a += (T)P.JinvG[4 * i + ucol];
            K[9 + i] = a;
        }
    }
    if (NX == 17) {
        K[12 < NX ? 12 : 0] = (ucol == 4) ? T(1) : T(0);
        K[13 < NX ? 13 : 0] = (ucol == 5) ? T(1) : T(0);
        MPCB_UNROLL
        for (int i = 0; i < 3; i++) {
            T a = pp[15 + i] * S[6] + pp[18 + i] * S[7] + pp[21 + i] * S[8];
            a += pp[6 + i] * K[3] + pp[9 + i] * K[4] + pp[12 + i] * K[5];
            if (ucol == 4) a += pp[0 + i];
            if (ucol == 5) a += pp[3 + i];
            K[(14 + i) < NX ? 14 + i : 0] = a;
        }
    }
}

}  // namespace mpcb
