// Rollout + forward-sensitivity kernel bodies.
//
// One warp integrates ONE shooting interval (instance i, stage k) with the classic RK4
// tableau applied to the augmented ODE [x; S],  Sdot = (df/dx) S + [df/du 0]  -- what
// acados' ERK integrator does with CasADi's forward VDE (integrator_type='ERK',
// reference blastermodel.py:277; 4 stages x 1 step, acados_ocp_blasterModel.json
// sim_method_num_stages/num_steps).
//
// Lane c < NZ owns sensitivity column c, ordered like the stage variable z = [u; x]
// (c < NU: d/du_c, else d/dx_{c-NU}); lane NZ (and the idle lanes above it) integrate
// the state itself.  The column lives in registers; the stage point is broadcast from the
// state lane by shuffles.  Nothing is staged in shared memory.
#pragma once
#include "mpcb_model.cuh"

namespace mpcb {

template <int NX, int NU, typename T>
MPCB_DEV void linearize_warp(const Params &P, const T *__restrict__ Xk, const T *__restrict__ Uk,
                             const T *__restrict__ Xk1, const T *__restrict__ pp, T *__restrict__ ws_stage)
{
    using L = Layout<NX, NU>;
    constexpr int NZ = L::NZ;
    const int lane = lane_id();
    const bool state_lane = lane >= NZ;
    const int ucol = lane < NU ? lane : -1;
    const int xcol = (lane >= NU && lane < NZ) ? lane - NU : -1;
    const T h = (T)P.dt;

    T u[NU];
    MPCB_UNROLL
    for (int j = 0; j < NU; j++) u[j] = Uk[j];
    const T Tsum = u[0] + u[1] + u[2] + u[3];
    const T Tb = pp[24];

    T x0[NX];  // only meaningful on the state lanes
    MPCB_UNROLL
    for (int i = 0; i < NX; i++) x0[i] = Xk[i];

    T K[NX], acc[NX];
    MPCB_UNROLL
    for (int i = 0; i < NX; i++) {
        K[i] = T(0);
        acc[i] = state_lane ? x0[i] : (i == xcol ? T(1) : T(0));
    }

    MPCB_UNROLL
    for (int s = 0; s < 4; s++) {
        const T a = (s == 0) ? T(0) : (s == 3 ? h : T(0.5) * h);
        const T bw = (s == 0 || s == 3) ? h / T(6) : h / T(3);
        // stage point: x + a*k, broadcast from the state lane
        T xs[NX];
        MPCB_UNROLL
        for (int i = 0; i < NX; i++) xs[i] = warp_shfl(x0[i] + a * K[i], NZ);
        Trig<NX, T> g;
        eval_trig<NX, T>(xs, g);
        StagePoint<NX, T> sp;
        eval_point<NX, T>(g, xs, Tsum, Tb, sp);
        if (state_lane) {
            eval_f<NX, NU, T>(P, sp, xs, u, pp, K);
        } else {
            T S[NX];
            MPCB_UNROLL
            for (int i = 0; i < NX; i++) S[i] = (i == xcol ? T(1) : T(0)) + a * K[i];
            eval_jac_col<NX, NU, T>(P, g, sp, xs, pp, Tb, S, ucol, K);
        }
        MPCB_UNROLL
        for (int i = 0; i < NX; i++) acc[i] += bw * K[i];
    }

    if (lane < NZ) {
        // row `lane` of [B'; A']
        T *row = ws_stage + L::O_BAT + lane * L::LDB;
        MPCB_UNROLL
        for (int i = 0; i < NX; i++) row[i] = acc[i];
    } else if (lane == NZ) {
        // b_k = phi(X_k, U_k) - X_{k+1}
        MPCB_UNROLL
        for (int i = 0; i < NX; i++) ws_stage[L::O_B + i] = acc[i] - Xk1[i];
    }
}

// Plant step = the same RK4 step without sensitivities (AcadosSimSolver built from the
// same OCP, reference blastermodel.py:290, simulation_blaster.py:94-104); one instance
// per thread.
template <int NX, int NU, typename T>
MPCB_DEV void plant_step_thread(const Params &P, const T *__restrict__ x, const T *__restrict__ uin,
                                const T *__restrict__ pp, T *__restrict__ xn)
{
    const T h = (T)P.dt;
    T u[NU], x0[NX], K[NX], acc[NX];
    MPCB_UNROLL
    for (int j = 0; j < NU; j++) u[j] = uin[j];
    const T Tsum = u[0] + u[1] + u[2] + u[3];
    const T Tb = pp[24];
    MPCB_UNROLL
    for (int i = 0; i < NX; i++) { x0[i] = x[i]; acc[i] = x0[i]; K[i] = T(0); }
    MPCB_UNROLL
    for (int s = 0; s < 4; s++) {
        const T a = (s == 0) ? T(0) : (s == 3 ? h : T(0.5) * h);
        const T bw = (s == 0 || s == 3) ? h / T(6) : h / T(3);
        T xs[NX];
        MPCB_UNROLL
        for (int i = 0; i < NX; i++) xs[i] = x0[i] + a * K[i];
        Trig<NX, T> g;
        eval_trig_local<NX, T>(xs, g);
        StagePoint<NX, T> sp;
        eval_point<NX, T>(g, xs, Tsum, Tb, sp);
        eval_f<NX, NU, T>(P, sp, xs, u, pp, K);
        MPCB_UNROLL
        for (int i = 0; i < NX; i++) acc[i] += bw * K[i];
    }
    MPCB_UNROLL
    for (int i = 0; i < NX; i++) xn[i] = acc[i];
}

}  // namespace mpcb
