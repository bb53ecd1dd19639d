/* mpcb.h -- C ABI of the B200-native batched BLASTER MPC solver.
 *
 * Drop-in boundary for the per-control-step optimal-control solve of
 * sml93/mpc_blaster.  In the reference that solve is reached through acados' ctypes
 * wrapper around a generated shared library; the calls it replaces are the loop body of
 * reference src/scripts/simulation_blaster.py:56-105:
 *
 *   ocp_solver.set(0,'lbx'|'ubx',x)      :60-61   -> x0 argument of mpcb_solve
 *   ocp_solver.cost_set(k,'yref',yref)   :63-78   -> yref argument (+ yref_mode)
 *   ocp_solver.set(k,'p',params)         :67-69   -> p argument (+ p_mode)
 *   ocp_solver.solve()                   :80      -> mpcb_solve
 *   ocp_solver.get(0,'u') / get(k,'x')   :87-89   -> u0 / X / U outputs
 *   ocp_solver.get_cost()                :86      -> mpcb_cost
 *   integrator.set/solve/get             :84-104  -> mpcb_plant_step
 *   blasterModel(...) ctor + generateController()  (blastermodel.py:16,214-292) -> mpcb_config / mpcb_create
 *
 * All array arguments are plain pointers, row-major, FP64.  Unless a function name ends in
 * _host, pointers are DEVICE pointers on the handle's GPU and the call is asynchronous on
 * `stream` (a cudaStream_t passed as void*; NULL = default stream).  No allocation happens
 * inside mpcb_solve / mpcb_plant_step / mpcb_cost, so they are CUDA-graph capturable.
 * One handle per stream; handles are independent (one per GPU for multi-GPU sharding).  As with any
 * CUDA library that takes a stream, the caller makes the handle's GPU the current device before a
 * stream-based call (the _host entry point does it itself).
 *
 * Return value: 0 on success, negative on API misuse or CUDA failure (message through
 * mpcb_last_error).  Per-instance solver outcome goes to status[B], mirroring acados'
 * codes: 0 success, 1 NaN, 2 max-iter, 3 min-step, 4 QP failure.
 *
 * Failed solves (status != 0): the stored iterate of that instance is left untouched, so the u0 / X / U returned for it
 * are the PREVIOUS iterate's (stale) values -- check status (mpcb_closed_loop counts such steps in n_fail).  With
 * mpcb_config.strict_reference = 1 a max-iter solve (status 2) does apply its last interior-point iterate, as acados does.
 */
#ifndef MPCB_H
#define MPCB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCB_VARIANT_BLASTER17 17 /* the reference's model: 17 states / 6 inputs / 25 params */
#define MPCB_VARIANT_QUAD12 12    /* states 0..11, inputs 0..3, gimbal frozen (north_star sizing) */
#define MPCB_VARIANT_QUAT13 13    /* QUAD12 with a quaternion attitude: x = [p(3), q(w,x,y,z), v(3), omega(3)], 4 inputs;
                                     quaternion algebra of reference utils/MathUtils.py:5-54 (which no reference model
                                     uses); Q/lbx/ubx entries 3..6 belong to q; the iterate must be initialised with
                                     unit quaternions (mpcb_reset with x_init) -- the all-zero default is not one */
#define MPCB_NP 25
#define MPCB_F64 64
#define MPCB_F32 32

/* yref_mode / p_mode */
#define MPCB_SHARED 0       /* one vector for every instance and stage: yref[ny] / p[25]      */
#define MPCB_PER_INSTANCE 1 /* yref[B,ny] / p[B,25], same on every stage of an instance      */
#define MPCB_PER_STAGE 2    /* yref[B,N+1,ny] (terminal row uses its first nx) / p[B,N,25]    */

/* POD mirror of blasterModel(mass,J,l_x,l_y,N,Tf,c,Q,R,Q_t,blastThruster,statesBound,controlBound)
 * (reference blastermodel.py:16); Q/R/Qt are the diagonals (the reference's weights are
 * diagonal: acados_ocp_blasterModel.json cost.W). */
typedef struct mpcb_config {
    int32_t variant;      /* MPCB_VARIANT_* */
    int32_t N;            /* horizon length */
    double dt;            /* Tf / N */
    double mass, J[9], l_x, l_y, c;
    double Q[17], R[6], Qt[17];
    double lbx[17], ubx[17], lbu[6], ubu[6];
    /* interior-point options (defaults = HPIPM's documented defaults, qp_solver_iter_max from blastermodel.py:279 is 500) */
    int32_t ipm_max_iter;
    double ipm_mu0;       /* initial lam*t of every bound (cold start) */
    double ipm_thr0;      /* initial slack floor: >= 0 absolute, < 0 the fraction -ipm_thr0 of the box width */
    double tol_stat, tol_eq, tol_ineq, tol_comp, alpha_min;
    int32_t dtype;        /* arithmetic type: MPCB_F64 (the reference computes in IEEE double); MPCB_F32 is rejected, see DESIGN.md */
    int32_t max_batch;    /* capacity of the persistent iterate (instances) */
    int32_t ws_batch;     /* instances of solver workspace resident at once (0 = auto) */
    int32_t device;       /* CUDA device ordinal, -1 = current */
    /* Reference-semantics switch.  0 (default): the throughput-oriented stopping rules of DESIGN.md section 2 -- the
     * stationarity norm tracked through its exact-arithmetic decay, the dynamics / bound-slack norms measured on every
     * iterate, early exit (status 3) when the multipliers diverge, a failed QP leaves the iterate untouched.  1: what the
     * reference's stack does -- explicitly evaluated residual norms (stationarity included) in the stopping test, no
     * divergence exit (the interior point runs to ipm_max_iter; the reference sets qp_solver_iter_max = 500,
     * blastermodel.py:279), and the last interior-point iterate is applied when that cap is hit (acados SQP_RTI takes
     * the QP solver's last iterate on max-iter; the reference script ignores the status, simulation_blaster.py:80).
     * Strict solves always run the one-instance-per-warp kernel. */
    int32_t strict_reference;
    /* Kernel selection by workspace-chunk size, 0 = measured defaults: chunks of at least
     * qp8_batch instances (default 8,192 BLASTER17 / 4,096 QUAD12, QUAT13) use the four-instances-per-warp kernel, chunks
     * of at least throughput_batch the single-buffer variant of the one-instance-per-warp kernel (default: never -- the
     * latency variant is faster at every size measured since round 2).  qp8_warps > 0 caps the qp8 grid (test hook). */
    int32_t throughput_batch, qp8_batch, qp8_warps;
} mpcb_config;

typedef struct mpcb_handle mpcb_handle;

/* canonical constants of reference simulation_blaster.py:12-30, dt = 1/30 s */
int mpcb_config_default(mpcb_config *cfg, int variant, int N);

/* Fails (message through mpcb_last_error(NULL)) without a CUDA device, for dtype != MPCB_F64, for boxes without an
 * interior (lbx < ubx and lbu < ubu must hold strictly: pin a variable with a tiny box, not a zero-width one), for
 * negative Q / Qt or non-positive R entries. */
int mpcb_create(const mpcb_config *cfg, mpcb_handle **out);
int mpcb_destroy(mpcb_handle *h);
const char *mpcb_last_error(const mpcb_handle *h); /* h may be NULL: error of the last failed create */

int mpcb_nx(const mpcb_handle *h);
int mpcb_nu(const mpcb_handle *h);
int mpcb_horizon(const mpcb_handle *h);

/* Set the persistent SQP iterate of instances [0,B): X_k = x_init[i] for all k, U_k = u_init
 * (u_per_instance ? u_init[B,nu] : u_init[nu]).  NULL pointers mean zeros (acados' default
 * initial iterate).  Replaces ocp_solver.set(k,'x'|'u',...). */
int mpcb_reset(mpcb_handle *h, const double *x_init, const double *u_init, int u_per_instance, int B, void *stream);

/* One SQP-RTI iteration for instances [0,B): linearise about the stored iterate (RK4 +
 * forward sensitivities), build the Gauss-Newton QP, solve it, take the full step.
 * Outputs (each may be NULL): u0[B,nu], X[B,N+1,nx], U[B,N,nu], status[B], iters[B]. */
int mpcb_solve(mpcb_handle *h, const double *x0, const double *yref, int yref_mode, const double *p, int p_mode,
               double *u0, double *X, double *U, int32_t *status, int32_t *iters, int B, void *stream);

/* Same with HOST buffers: stages inputs through pinned memory, copies to the device,
 * solves, copies the requested outputs back and synchronises.  p may be NULL (default
 * parameters: POC Jacobians 0, T_blast = 2.2*9.81, reference blastermodel.py:280-282). */
int mpcb_solve_host(mpcb_handle *h, const double *x0, const double *yref, int yref_mode, const double *p, int p_mode,
                    double *u0, double *X, double *U, int32_t *status, int32_t *iters, int B);

/* Multi-iteration SQP on the same (x0, yref, p): every iteration re-linearises about the iterate the previous one
 * produced (nlp_solver_type SQP instead of SQP_RTI; the reference's options carry nlp_solver_max_iter = 100 and
 * nlp_solver_tol_stat/eq/ineq/comp = 1e-6 but select SQP_RTI: blastermodel.py:278, acados_ocp_blasterModel.json
 * solver_options).
 *   tol != NULL: tol[4] = {stat, eq, ineq, comp}.  Per instance, acados' SQP loop: linearise; evaluate the NLP residuals
 *     (inf-norms of the Lagrangian gradient with the last QP's multipliers, of the dynamics defect, of the bound
 *     violation and of the complementarity products); stop when all four are within tolerance; otherwise solve the QP and
 *     take the full step; at most max_iter QPs.  Finished instances are skipped by the kernels.  The QPs of this mode are
 *     always solved with the reference-semantics rule set (strict_reference: explicit residual norms and one step of
 *     iterative refinement), whatever the handle's setting: the Lagrangian-gradient test needs the QP's multipliers to
 *     1e-6, which the default rule set does not deliver for active bounds.  status[B]: 0 converged,
 *     2 max_iter reached, 4 a QP failed (its own status was not 0), 1 NaN; iters[B]: interior-point iterations summed
 *     over the QPs; sqp_iters[B]: QPs solved; nlp_res[B,4]: the last evaluated residuals (the final iterate's).
 *   tol == NULL: exactly max_iter SQP iterations without a residual test; status / iters are those of the last QP.
 * Outputs may be NULL. */
int mpcb_solve_sqp(mpcb_handle *h, const double *x0, const double *yref, int yref_mode, const double *p, int p_mode,
                   int max_iter, const double *tol, double *u0, double *X, double *U, int32_t *status, int32_t *iters,
                   int32_t *sqp_iters, double *nlp_res, int B, void *stream);

/* Shift the stored iterate one stage forward (X_k <- X_{k+1}, U_k <- U_{k+1}, last stage repeated):
 * the standard warm start between control steps.  The reference's scripts never shift
 * (simulation_blaster.py:56-105), so this is opt-in. */
int mpcb_shift(mpcb_handle *h, int B, void *stream);

/* Plant step x+ = RK4(x,u,p) over dt for B instances (AcadosSimSolver of blastermodel.py:290).
 * p_mode: MPCB_SHARED or MPCB_PER_INSTANCE. */
int mpcb_plant_step(mpcb_handle *h, const double *x, const double *u, const double *p, int p_mode, double *xnext,
                    int B, void *stream);

/* `steps` closed-loop control steps entirely on the device (simulation_blaster.py:56-105):
 * solve from x, apply u0 to the plant, repeat.  x[B,nx] is updated in place.
 * Optional outputs: u_last[B,nu], n_fail[B] (number of steps whose solve status != 0),
 * iters_sum[B] (IPM iterations over all steps). */
int mpcb_closed_loop(mpcb_handle *h, double *x, const double *yref, int yref_mode, const double *p, int p_mode,
                     int steps, double *u_last, int32_t *n_fail, int32_t *iters_sum, int B, void *stream);

/* cost[B] = sum_k dt/2 |y_k - yref_k|^2_W + 1/2 |x_N - yref_N|^2_We at the stored iterate (get_cost()). */
int mpcb_cost(mpcb_handle *h, const double *yref, int yref_mode, double *cost, int B, void *stream);

/* Read / write the stored iterate (device pointers; either may be NULL). */
int mpcb_get_iterate(mpcb_handle *h, double *X, double *U, int B, void *stream);
int mpcb_set_iterate(mpcb_handle *h, const double *X, const double *U, int B, void *stream);

/* Test hook: run only the rollout + sensitivity kernel for instances [0,B) and return
 * BAt[B,N,nz,nx] = [B_k'; A_k'] and b[B,N,nx] (device pointers). */
int mpcb_debug_linearize(mpcb_handle *h, const double *p, int p_mode, double *BAt, double *b, int B, void *stream);

/* Test hook: the interior-point iterate the last mpcb_solve ended with and the QP it solved, copied out of the solver
 * workspace (valid right after an mpcb_solve of B <= ws_batch instances; nothing is recomputed).  Device pointers, any
 * may be NULL: z[B,N+1,nz] (z_k = [du_k; dx_k], dx_0 pinned), pi[B,N+1,nx] (pi_0 unused), slacks tl/tu and multipliers
 * ll/lu [B,N+1,nz] (zero where a component has no bound), the bounds on the increments lb/ub [B,N+1,nz] (-inf/+inf where
 * absent), the cost gradient g[B,N+1,nz], and the linearisation BAt[B,N,nz,nx] = [B_k'; A_k'], b[B,N,nx] (both or
 * neither).  tests/ evaluate the explicit KKT residuals of the GPU's solution from these. */
int mpcb_debug_qp(mpcb_handle *h, double *z, double *pi, double *tl, double *tu, double *ll, double *lu, double *lb, double *ub,
                  double *g, double *BAt, double *b, int B, void *stream);

/* Profiling aid for bench.py: when enabled, CUDA events are recorded on the launching stream
 * around the rollout kernel and the QP kernel of each solve (first workspace chunk);
 * mpcb_last_kernel_ms waits for the last solve and returns the two durations. */
int mpcb_profile(mpcb_handle *h, int enable);
int mpcb_last_kernel_ms(mpcb_handle *h, float *ms_linearize, float *ms_qp);

/* Measured FP64 FMA-pipe peak of a device (dependent-chain DFMA micro-kernel), in TFLOP/s:
 * the roofline denominator for this path (MEASURED_PEAKS.json carries only HBM and bf16). */
int mpcb_fp64_peak(int device, double *tflops);

/* Number of kernels this library has launched since it was loaded (for bench accounting). */
int64_t mpcb_kernel_launches(void);

/* Command mapping after the solve (reference mavros_blaster_sim.py:27-30,91-100):
 * attitude quaternion [w,x,y,z] of R = Rz(psi)Ry(theta)Rx(phi) from x[3:6] and the
 * normalised collective thrust from the cubic map of thrusterCumul(). quat[B,4], thrust[B]. */
int mpcb_command_map(mpcb_handle *h, const double *x, const double *u0, double *quat, double *thrust, int B, void *stream);

/* Jet point of contact (POC) and its Jacobians for B vehicle poses: the parameter generator that
 * feeds p[0:24] of the OCP.  Replaces the reference's Jacobian_POC_Solver.solveJacobians()
 * (src/scripts/Jacobian_POC_Solver.py:234-300, with setInitConditions :153-175, htm.py:7-36, the
 * ERK jet integrator :59-99 and the Newton time-of-flight search :115-152), which the scripts run
 * once on the CPU before the control loop (simulation_blaster.py:37-39).
 *   poses: euler[B,3] (phi,theta,psi), motor[B,2] (alpha1,alpha2), position[B,3] -- or x17[B,17]
 *          state vectors (then the three arrays may be NULL);
 *   stream_velocity, drag: constructor arguments of the reference class (150, 1);
 *   mode MPCB_POC_REFERENCE: the reference's algorithm step for step (RK4 x10, Newton with a
 *          forward-difference slope to |z| <= 1e-3, forward differences eps = 1e-6);
 *        MPCB_POC_ANALYTIC: closed-form flight, exact root, implicit-function Jacobians;
 *   outputs (device pointers, any may be NULL): poc[B,3], J_mot[B,3,2], J_eul[B,3,3], J_pos[B,3,3]
 *          row-major, p25[B,25] packed as simulation_blaster.py:67 does (column-major blocks, then
 *          T_blast), t_flight[B], status[B] (0 ok, 1 NaN, 2 iteration cap).
 * No handle: the call is stateless. */
#define MPCB_POC_REFERENCE 0
#define MPCB_POC_ANALYTIC 1
int mpcb_poc_jacobians(const double *euler, const double *motor, const double *position, const double *x17, int B,
                       double stream_velocity, double drag, int mode, double T_blast, double *poc, double *J_mot, double *J_eul,
                       double *J_pos, double *p25, double *t_flight, int32_t *status, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MPCB_H */
