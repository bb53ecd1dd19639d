/* Body of the C oracle, included once per model variant with
 *   NX, NU   compile-time state/input dimension
 *   SFX(x)   name-mangling macro
 *   MODEL_F, MODEL_JAC   the dynamics and its Jacobians (arrays padded to 17 states / 6 inputs)
 * TEST INFRASTRUCTURE ONLY -- see mpc_oracle.c for the header comment.
 *
 * Stage variable ordering everywhere: z_k = [du_k (NU); dx_k (NX)], NZ = NU+NX.
 * BAt_k = [B_k' ; A_k'] is NZ x NX row-major, so dx_{k+1} = BAt_k' z_k + b_k.
 */
#define NZ (NX + NU)

typedef struct {
    /* per-stage QP data (k = 0..N-1 unless noted) */
    double *BAt;  /* N * NZ*NX */
    double *b;    /* N * NX */
    double *g;    /* (N+1) * NZ  (u part unused at N) */
    double *lb, *ub; /* (N+1) * NZ, +-HUGE_VAL where absent */
    /* primal/dual iterate */
    double *z;    /* (N+1) * NZ */
    double *pi;   /* (N+1) * NX, pi[0] unused */
    double *tl, *tu, *ll, *lu; /* (N+1) * NZ */
    /* step */
    double *dz, *dpi, *dtl, *dtu, *dll, *dlu;
    /* residuals */
    double *rg, *rb, *rdl, *rdu, *q, *rml, *rmu;
    /* factorisation */
    double *L;    /* (N+1) * NZ*NZ lower, row-major */
    double *pv;   /* (N+1) * NX */
    double *lvec; /* (N+1) * NZ : Luu^{-1} l_u (first NU) */
    double *Hd;   /* (N+1) * NZ diag Hessian + barrier */
    double *H0;   /* (N+1) * NZ diag Hessian */
    int pform;    /* experiment ric_alg = 3: the xx block of L holds the full symmetric P_k, not its factor */
    int split_off; /* experiment ric_alg = 3: the list of carried columns has overflowed in this solve -> LQ from now on */
} SFX(ws_t);

static size_t SFX(ws_doubles)(int N)
{
    size_t s = (size_t)(N + 1);
    return (size_t)N * NZ * NX + (size_t)N * NX + s * NZ * 3 + s * NZ + s * NX + s * NZ * 4 + s * NZ + s * NX
           + s * NZ * 4 + s * NZ * 2 + s * NX + s * NZ * 4 + s * NZ * NZ + s * NX + s * NZ + s * NZ * 2 + 64;
}

static void SFX(ws_bind)(SFX(ws_t) * w, double *m, int N)
{
    size_t s = (size_t)(N + 1);
#define TAKE(ptr, n) do { (ptr) = m; m += (n); } while (0)
    TAKE(w->BAt, (size_t)N * NZ * NX); TAKE(w->b, (size_t)N * NX);
    TAKE(w->g, s * NZ); TAKE(w->lb, s * NZ); TAKE(w->ub, s * NZ);
    TAKE(w->z, s * NZ); TAKE(w->pi, s * NX);
    TAKE(w->tl, s * NZ); TAKE(w->tu, s * NZ); TAKE(w->ll, s * NZ); TAKE(w->lu, s * NZ);
    TAKE(w->dz, s * NZ); TAKE(w->dpi, s * NX);
    TAKE(w->dtl, s * NZ); TAKE(w->dtu, s * NZ); TAKE(w->dll, s * NZ); TAKE(w->dlu, s * NZ);
    TAKE(w->rg, s * NZ); TAKE(w->rb, s * NX); TAKE(w->rdl, s * NZ); TAKE(w->rdu, s * NZ);
    TAKE(w->q, s * NZ); TAKE(w->rml, s * NZ); TAKE(w->rmu, s * NZ);
    TAKE(w->L, s * NZ * NZ); TAKE(w->pv, s * NX); TAKE(w->lvec, s * NZ); TAKE(w->Hd, s * NZ); TAKE(w->H0, s * NZ);
#undef TAKE
}

/* ---- A1: xdot = f(x,u,p)  (blastermodel.py:93-167,191-201); QUAD12 = first 12 rows,
 *      gimbal frozen at 0, inputs 4,5 absent. -------------------------------------- */
static void SFX(f)(const orc_problem *P, const double *x, const double *u, const double *p, double *xd)
{
    double x17[17] = {0}, u6[6] = {0}, out[17];
    for (int i = 0; i < NX; i++) x17[i] = x[i];
    for (int i = 0; i < NU; i++) u6[i] = u[i];
    MODEL_F(P, x17, u6, p, out);
    for (int i = 0; i < NX; i++) xd[i] = out[i];
}

/* ---- A3: ERK4 + forward sensitivities, one step of dt.  BAt out = [B';A'] (NZ x NX),
 *      xn = phi(x,u,p).  [upstream D6] ------------------------------------------------ */
static void SFX(rk4_sens)(const orc_problem *P, const double *x, const double *u, const double *p,
                          double *xn, double *BAt)
{
    /* S is kept transposed: St[c][i] = dS_i/d(col c), columns ordered [u (NU); x (NX)] like z */
    double St[NZ][NX], Ss[NZ][NX], K[NZ][NX], Sacc[NZ][NX];
    double xs[NX], k[NX], xacc[NX];
    double x17[17] = {0}, u6[6] = {0}, f17[17], fx[17][17], fu[17][6];
    const double h = P->dt;
    const double ca[4] = {0.0, 0.5, 0.5, 1.0};
    const double cb[4] = {1.0 / 6, 2.0 / 6, 2.0 / 6, 1.0 / 6};
    for (int c = 0; c < NZ; c++)
        for (int i = 0; i < NX; i++) { St[c][i] = (c >= NU && c - NU == i) ? 1.0 : 0.0; Sacc[c][i] = St[c][i]; K[c][i] = 0.0; }
    for (int i = 0; i < NX; i++) { xacc[i] = x[i]; k[i] = 0.0; }
    for (int i = 0; i < NU; i++) u6[i] = u[i];
    for (int s = 0; s < 4; s++) {
        for (int i = 0; i < NX; i++) xs[i] = x[i] + ca[s] * h * k[i];
        for (int c = 0; c < NZ; c++)
            for (int i = 0; i < NX; i++) Ss[c][i] = St[c][i] + ca[s] * h * K[c][i];
        for (int i = 0; i < NX; i++) x17[i] = xs[i];
        MODEL_F(P, x17, u6, p, f17);
        MODEL_JAC(P, x17, u6, p, fx, fu);
        for (int i = 0; i < NX; i++) k[i] = f17[i];
        for (int c = 0; c < NZ; c++)
            for (int i = 0; i < NX; i++) {
                double acc = (c < NU) ? fu[i][c] : 0.0;
                for (int j = 0; j < NX; j++) acc += fx[i][j] * Ss[c][j];
                K[c][i] = acc;
            }
        for (int i = 0; i < NX; i++) xacc[i] += h * cb[s] * k[i];
        for (int c = 0; c < NZ; c++)
            for (int i = 0; i < NX; i++) Sacc[c][i] += h * cb[s] * K[c][i];
    }
    for (int i = 0; i < NX; i++) xn[i] = xacc[i];
    for (int c = 0; c < NZ; c++)
        for (int i = 0; i < NX; i++) BAt[c * NX + i] = Sacc[c][i];
}

static void SFX(plant_step)(const orc_problem *P, const double *x, const double *u, const double *p, double *xn)
{
    double k1[NX], k2[NX], k3[NX], k4[NX], xs[NX];
    const double h = P->dt;
    SFX(f)(P, x, u, p, k1);
    for (int i = 0; i < NX; i++) xs[i] = x[i] + 0.5 * h * k1[i];
    SFX(f)(P, xs, u, p, k2);
    for (int i = 0; i < NX; i++) xs[i] = x[i] + 0.5 * h * k2[i];
    SFX(f)(P, xs, u, p, k3);
    for (int i = 0; i < NX; i++) xs[i] = x[i] + h * k3[i];
    SFX(f)(P, xs, u, p, k4);
    for (int i = 0; i < NX; i++) xn[i] = x[i] + h / 6.0 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

/* ---- A6: backward Riccati factorisation in square-root form (HPIPM's default
 *      algorithm class):  M_k = diag(Hd_k) + BAt_k P_{k+1} BAt_k',  P_{k+1} = Lxx Lxx',
 *      L_k = chol(M_k) = [Luu 0; Lxu Lxx].  Returns 0, or 1 when a pivot is not positive. */
/* Experiment ric_alg = 3 (tools/split_factor_viability.py; never the checker): the classical Riccati recursion on the
 * matrix P_k itself (normal-equations form, input pivots only), made safe for numerically pinned states by keeping
 * their contributions OUT of the matrix: P_k = P'_k + V_k V_k' with P'_k of ordinary size and V_k a short list of huge
 * columns (sqrt(Hd_j) e_j for every state whose barrier diagonal exceeds tau_d, plus what earlier pins leave behind
 * after the inputs have been eliminated).  Stage k: Cholesky of the input block of D + [B A]' P' [B A], then the
 * columns [B A]' V are brought in by a Householder LQ update of the NU factor columns (the stable rank-h update);
 * what the update leaves in the state rows is V_k.  Columns whose squared norm has dropped below tau are folded
 * into P'.  The xx block of L then holds P_k in full (for the P r + p products of the solves). */
static int SFX(ric_factor_split)(const orc_problem *P, SFX(ws_t) * w, int N)
{
    const double tau = orc_huge_tau(), tau_d = orc_huge_tau_d();
    double Pp[NX][NX], V[NX + NX][NX];
    int h = 0;
    double *LN = w->L + (size_t)N * NZ * NZ;
    for (int i = 0; i < NZ * NZ; i++) LN[i] = 0.0;
    for (int i = 0; i < NX; i++)
        for (int c = 0; c < NX; c++) Pp[i][c] = 0.0;
    for (int i = 0; i < NX; i++) {
        double d = w->Hd[(size_t)N * NZ + NU + i];
        if (!(d > 0.0)) return 1;
        Pp[i][i] = d;
        LN[(NU + i) * NZ + NU + i] = d;
    }
    for (int k = N - 1; k >= 0; k--) {
        const double *BAt = w->BAt + (size_t)k * NZ * NX;
        double *L = w->L + (size_t)k * NZ * NZ;
        double M[NZ][NZ], T[NZ][NX], Wh[NZ][NX + NX];
        for (int i = 0; i < NZ; i++)
            for (int c = 0; c < NX; c++) {
                double acc = 0.0;
                for (int j = 0; j < NX; j++) acc += BAt[i * NX + j] * Pp[j][c];
                T[i][c] = acc;
            }
        for (int i = 0; i < NZ; i++)
            for (int j = 0; j < NZ; j++) {
                double acc = 0.0;
                for (int c = 0; c < NX; c++) acc += T[i][c] * BAt[j * NX + c];
                M[i][j] = acc;
            }
        int pin[NX], np = 0;
        for (int i = 0; i < NZ; i++) {
            const double d = w->Hd[(size_t)k * NZ + i];
            if (!(d > 0.0)) return 1;
            if (i >= NU && d > tau_d) pin[np++] = i - NU;
            else M[i][i] += d;
        }
        for (int i = 0; i < NZ; i++)
            for (int q = 0; q < h; q++) {
                double acc = 0.0;
                for (int c = 0; c < NX; c++) acc += BAt[i * NX + c] * V[q][c];
                Wh[i][q] = acc;
            }
        for (int i = 0; i < NZ * NZ; i++) L[i] = 0.0;
        for (int j = 0; j < NU; j++) {
            double d = M[j][j];
            if (!(d > 0.0)) return 1;
            d = sqrt(d);
            L[j * NZ + j] = d;
            for (int i = j + 1; i < NZ; i++) L[i * NZ + j] = M[i][j] / d;
            for (int i = j + 1; i < NZ; i++)
                for (int c = j + 1; c < NZ; c++) M[i][c] -= L[i * NZ + j] * L[c * NZ + j];
        }
        if (h)
            for (int j = 0; j < NU; j++) {
                const double d = L[j * NZ + j];
                double s2 = d * d;
                for (int q = 0; q < h; q++) s2 += Wh[j][q] * Wh[j][q];
                const double sig = sqrt(s2), v0 = d + sig, beta = 1.0 / (sig * v0);
                L[j * NZ + j] = sig;
                for (int i = j + 1; i < NZ; i++) {
                    double dot = v0 * L[i * NZ + j];
                    for (int q = 0; q < h; q++) dot += Wh[j][q] * Wh[i][q];
                    const double f = beta * dot;
                    L[i * NZ + j] = f * v0 - L[i * NZ + j];
                    for (int q = 0; q < h; q++) Wh[i][q] -= f * Wh[j][q];
                }
            }
        /* P'_k = Schur complement of the ordinary part; the list: what the update left in the state rows, and the new pins */
        for (int i = 0; i < NX; i++)
            for (int c = 0; c < NX; c++) Pp[i][c] = M[NU + i][NU + c];
        int hn = 0;
        double Vn[NX + NX][NX];
        for (int q = 0; q < h; q++) {
            double n2 = 0.0;
            for (int i = 0; i < NX; i++) n2 += Wh[NU + i][q] * Wh[NU + i][q];
            if (n2 > tau && hn + np >= orc_huge_hmax()) w->split_off = 1; /* overflow: this column is folded, later iterations use the LQ */
            if (n2 > tau && hn + np < orc_huge_hmax()) {
                for (int i = 0; i < NX; i++) Vn[hn][i] = Wh[NU + i][q];
                hn++;
            } else {
                for (int i = 0; i < NX; i++)
                    for (int c = 0; c < NX; c++) Pp[i][c] += Wh[NU + i][q] * Wh[NU + c][q];
            }
        }
        for (int q = 0; q < np; q++) {
            if (hn >= orc_huge_hmax()) { w->split_off = 1; Pp[pin[q]][pin[q]] += w->Hd[(size_t)k * NZ + NU + pin[q]]; continue; }
            for (int i = 0; i < NX; i++) Vn[hn][i] = 0.0;
            Vn[hn][pin[q]] = sqrt(w->Hd[(size_t)k * NZ + NU + pin[q]]);
            hn++;
        }
        if (k > 0) orc_huge_count(hn);
        h = hn;
        for (int q = 0; q < h; q++)
            for (int i = 0; i < NX; i++) V[q][i] = Vn[q][i];
        for (int i = 0; i < NX; i++)
            for (int c = 0; c < NX; c++) {
                double acc = Pp[i][c];
                for (int q = 0; q < h; q++) acc += V[q][i] * V[q][c];
                L[(NU + i) * NZ + NU + c] = acc;
            }
    }
    return 0;
}

static int SFX(ric_factor)(const orc_problem *P, SFX(ws_t) * w, int N, double mu)
{
    w->pform = 0;
    if (P->ric_alg == 3 && !(mu > P->ipm_mu0) && mu > orc_split_mu() && !w->split_off) { w->pform = 1; orc_split_iter(1); return SFX(ric_factor_split)(P, w, N); }
    if (P->ric_alg == 3) orc_split_iter(0);
    const int f32 = P->mixed_mu > 0.0 && mu > P->mixed_mu && mu <= P->ipm_mu0;
    double *LN = w->L + (size_t)N * NZ * NZ;
    for (int i = 0; i < NZ * NZ; i++) LN[i] = 0.0;
    for (int i = NU; i < NZ; i++) {
        double d = w->Hd[(size_t)N * NZ + i];
        if (!(d > 0.0)) return 1;
        LN[i * NZ + i] = sqrt(d);
    }
    for (int k = N - 1; k >= 0; k--) {
        const double *BAt = w->BAt + (size_t)k * NZ * NX;
        const double *Ln = w->L + (size_t)(k + 1) * NZ * NZ; /* its xx block is chol(P_{k+1}) */
        double *L = w->L + (size_t)k * NZ * NZ;
        double W[NZ][NX];
        for (int i = 0; i < NZ; i++)
            for (int c = 0; c < NX; c++) {
                double acc = 0.0;
                for (int j = c; j < NX; j++) acc += BAt[i * NX + j] * Ln[(NU + j) * NZ + NU + c];
                W[i][c] = acc;
            }
        if (f32 && P->ric_alg == 1 && getenv("ORC_MIXED_LQ")) {
            /* mixed-precision experiment, second form: the Householder LQ itself in FP32 (no pivot can go negative) */
            float Wf[NZ][NX];
            for (int i = 0; i < NZ; i++)
                for (int c = 0; c < NX; c++) {
                    float acc = 0.0f;
                    for (int j = c; j < NX; j++) acc += (float)BAt[i * NX + j] * (float)Ln[(NU + j) * NZ + NU + c];
                    Wf[i][c] = acc;
                }
            for (int i = 0; i < NZ * NZ; i++) L[i] = 0.0;
            for (int j = 0; j < NZ; j++) {
                const float hd = (float)w->Hd[(size_t)k * NZ + j];
                if (!(hd > 0.0f)) return 1;
                const float d = sqrtf(hd);
                float s2 = hd;
                for (int c = 0; c < NX; c++) s2 += Wf[j][c] * Wf[j][c];
                const float sig = sqrtf(s2);
                const float v0 = d + sig;
                const float beta = 1.0f / (sig * v0);
                L[j * NZ + j] = (double)sig;
                for (int i = j + 1; i < NZ; i++) {
                    float dot = 0.0f;
                    for (int c = 0; c < NX; c++) dot += Wf[j][c] * Wf[i][c];
                    const float f = beta * dot;
                    L[i * NZ + j] = (double)(f * v0);
                    for (int c = 0; c < NX; c++) Wf[i][c] -= f * Wf[j][c];
                }
            }
        } else if (f32) {
            /* mixed-precision experiment: the same normal-equations factorisation with every operand and operation in FP32 */
            float Wf[NZ][NX], Lf[NZ][NZ];
            for (int i = 0; i < NZ; i++)
                for (int c = 0; c < NX; c++) {
                    float acc = 0.0f;
                    for (int j = c; j < NX; j++) acc += (float)BAt[i * NX + j] * (float)Ln[(NU + j) * NZ + NU + c];
                    Wf[i][c] = acc;
                }
            for (int i = 0; i < NZ; i++)
                for (int j = 0; j <= i; j++) {
                    float acc = (i == j) ? (float)w->Hd[(size_t)k * NZ + i] : 0.0f;
                    for (int c = 0; c < NX; c++) acc += Wf[i][c] * Wf[j][c];
                    Lf[i][j] = acc;
                }
            for (int j = 0; j < NZ; j++) {
                float d = Lf[j][j];
                for (int c = 0; c < j; c++) d -= Lf[j][c] * Lf[j][c];
                if (!(d > 0.0f)) return 1;
                d = sqrtf(d);
                Lf[j][j] = d;
                const float inv = 1.0f / d;
                for (int i = j + 1; i < NZ; i++) {
                    float sacc = Lf[i][j];
                    for (int c = 0; c < j; c++) sacc -= Lf[i][c] * Lf[j][c];
                    Lf[i][j] = sacc * inv;
                }
            }
            for (int i = 0; i < NZ; i++)
                for (int j = 0; j < NZ; j++) L[i * NZ + j] = (j <= i) ? (double)Lf[i][j] : 0.0;
        } else if (P->ric_alg == 0) {
            /* classical normal-equations form: M = Hd + W W', Cholesky (kept for experiments) */
            for (int i = 0; i < NZ; i++)
                for (int j = 0; j <= i; j++) {
                    double acc = (i == j) ? w->Hd[(size_t)k * NZ + i] : 0.0;
                    for (int c = 0; c < NX; c++) acc += W[i][c] * W[j][c];
                    L[i * NZ + j] = acc;
                }
            for (int j = 0; j < NZ; j++) {
                double d = L[j * NZ + j];
                for (int c = 0; c < j; c++) d -= L[j * NZ + c] * L[j * NZ + c];
                if (!(d > 0.0)) return 1;
                d = sqrt(d);
                L[j * NZ + j] = d;
                double inv = 1.0 / d;
                for (int i = j + 1; i < NZ; i++) {
                    double s = L[i * NZ + j];
                    for (int c = 0; c < j; c++) s -= L[i * NZ + c] * L[j * NZ + c];
                    L[i * NZ + j] = s * inv;
                }
                for (int c = j + 1; c < NZ; c++) L[j * NZ + c] = 0.0;
            }
        } else {
            /* array (LQ) form: [diag(sqrt(Hd)) | W] = L Q by Householder reflections from the
             * right; never forms M, so the error is O(eps*sqrt(barrier)) instead of
             * O(eps*barrier) -- what HPIPM's lq_fact option is for. */
            for (int i = 0; i < NZ * NZ; i++) L[i] = 0.0;
            for (int j = 0; j < NZ; j++) {
                const double hd = w->Hd[(size_t)k * NZ + j];
                if (!(hd > 0.0)) return 1;
                const double d = sqrt(hd);
                double s2 = hd;
                for (int c = 0; c < NX; c++) s2 += W[j][c] * W[j][c];
                const double sig = sqrt(s2);
                const double v0 = d + sig;
                const double beta = 1.0 / (sig * v0);
                L[j * NZ + j] = sig;
                for (int i = j + 1; i < NZ; i++) {
                    double dot = 0.0;
                    for (int c = 0; c < NX; c++) dot += W[j][c] * W[i][c];
                    const double f = beta * dot;
                    L[i * NZ + j] = f * v0;
                    for (int c = 0; c < NX; c++) W[i][c] -= f * W[j][c];
                }
            }
        }
    }
    return 0;
}

/* Solve the equality-constrained step QP for gradient q and dynamics residual rb
 * with the current factorisation:
 *   min sum 1/2 dz'(M)dz + q'dz   s.t. dx_{k+1} = BAt_k' dz_k + rb_k,  dx_0 = 0.
 * Outputs dz (all stages) and dpi (k = 1..N). */
static void SFX(ric_solve)(SFX(ws_t) * w, int N)
{
    /* backward: p_N = q_N(x) */
    for (int i = 0; i < NX; i++) w->pv[(size_t)N * NX + i] = w->q[(size_t)N * NZ + NU + i];
    for (int k = N - 1; k >= 0; k--) {
        const double *BAt = w->BAt + (size_t)k * NZ * NX;
        const double *Ln = w->L + (size_t)(k + 1) * NZ * NZ;
        const double *L = w->L + (size_t)k * NZ * NZ;
        const double *r = w->rb + (size_t)k * NX;
        const double *pn = w->pv + (size_t)(k + 1) * NX;
        double t1[NX], t2[NX], l[NZ];
        /* t2 = P r + p_{k+1},  P = Lxx Lxx' */
        for (int c = 0; c < NX; c++) {
            double acc = 0.0;
            for (int j = c; j < NX; j++) acc += Ln[(NU + j) * NZ + NU + c] * r[j];
            t1[c] = acc;
        }
        for (int i = 0; i < NX; i++) {
            double acc = pn[i];
            for (int c = 0; c <= i; c++) acc += Ln[(NU + i) * NZ + NU + c] * t1[c];
            t2[i] = acc;
        }
        if (w->pform)
            for (int i = 0; i < NX; i++) {
                double acc = pn[i];
                for (int c = 0; c < NX; c++) acc += Ln[(NU + i) * NZ + NU + c] * r[c];
                t2[i] = acc;
            }
        for (int i = 0; i < NZ; i++) {
            double acc = w->q[(size_t)k * NZ + i];
            for (int j = 0; j < NX; j++) acc += BAt[i * NX + j] * t2[j];
            l[i] = acc;
        }
        /* lvec_u = Luu^{-1} l_u ; p_k = l_x - Lxu lvec_u */
        double *lv = w->lvec + (size_t)k * NZ;
        for (int i = 0; i < NU; i++) {
            double s = l[i];
            for (int c = 0; c < i; c++) s -= L[i * NZ + c] * lv[c];
            lv[i] = s / L[i * NZ + i];
        }
        for (int i = 0; i < NX; i++) {
            double s = l[NU + i];
            for (int c = 0; c < NU; c++) s -= L[(NU + i) * NZ + c] * lv[c];
            w->pv[(size_t)k * NX + i] = s;
        }
    }
    /* forward */
    double dx[NX];
    for (int i = 0; i < NX; i++) dx[i] = 0.0;
    for (int k = 0; k < N; k++) {
        const double *BAt = w->BAt + (size_t)k * NZ * NX;
        const double *L = w->L + (size_t)k * NZ * NZ;
        const double *lv = w->lvec + (size_t)k * NZ;
        double *dz = w->dz + (size_t)k * NZ;
        double y[NU];
        for (int c = 0; c < NU; c++) {
            double s = lv[c];
            for (int i = 0; i < NX; i++) s += L[(NU + i) * NZ + c] * dx[i];
            y[c] = -s;
        }
        for (int i = NU - 1; i >= 0; i--) {
            double s = y[i];
            for (int c = i + 1; c < NU; c++) s -= L[c * NZ + i] * dz[c];
            dz[i] = s / L[i * NZ + i];
        }
        for (int i = 0; i < NX; i++) dz[NU + i] = dx[i];
        double xn[NX];
        for (int j = 0; j < NX; j++) xn[j] = w->rb[(size_t)k * NX + j];
        for (int i = 0; i < NZ; i++)
            for (int j = 0; j < NX; j++) xn[j] += BAt[i * NX + j] * dz[i];
        /* dpi_{k+1} = P_{k+1} dx_{k+1} + p_{k+1} */
        const double *Ln = w->L + (size_t)(k + 1) * NZ * NZ;
        double t1[NX];
        for (int c = 0; c < NX; c++) {
            double acc = 0.0;
            for (int j = c; j < NX; j++) acc += Ln[(NU + j) * NZ + NU + c] * xn[j];
            t1[c] = acc;
        }
        for (int i = 0; i < NX; i++) {
            double acc = w->pv[(size_t)(k + 1) * NX + i];
            for (int c = 0; c <= i; c++) acc += Ln[(NU + i) * NZ + NU + c] * t1[c];
            if (w->pform) {
                acc = w->pv[(size_t)(k + 1) * NX + i];
                for (int c = 0; c < NX; c++) acc += Ln[(NU + i) * NZ + NU + c] * xn[c];
            }
            w->dpi[(size_t)(k + 1) * NX + i] = acc;
        }
        for (int i = 0; i < NX; i++) dx[i] = xn[i];
    }
    double *dzN = w->dz + (size_t)N * NZ;
    for (int i = 0; i < NU; i++) dzN[i] = 0.0;
    for (int i = 0; i < NX; i++) dzN[NU + i] = dx[i];
}

/* ---- A6: Mehrotra predictor-corrector IPM around the Riccati solver.
 * Same algorithm class as HPIPM ([upstream D8]) but converged to the FP64 floor
 * so that the unique QP optimum is reproduced (see DESIGN.md "tolerances").
 * dx0 = z[0][NU..] is fixed by the caller.  Returns status, *iters. */
#define SKIP(k, j) (((k) == 0 && (j) >= NU) || ((k) == N && (j) < NU))

static void SFX(ipm_step_from)(SFX(ws_t) * w, int N)
{
    /* q = rg + (rml + ll*rdl)/tl - (rmu + lu*rdu)/tu ; solve ; recover dt, dlam */
    const size_t n = (size_t)(N + 1) * NZ;
    for (size_t i = 0; i < n; i++)
        w->q[i] = w->rg[i] + (w->rml[i] + w->ll[i] * w->rdl[i]) / w->tl[i] - (w->rmu[i] + w->lu[i] * w->rdu[i]) / w->tu[i];
    SFX(ric_solve)(w, N);
    for (size_t i = 0; i < n; i++) {
        w->dtl[i] = w->dz[i] + w->rdl[i];
        w->dtu[i] = -w->dz[i] + w->rdu[i];
        w->dll[i] = -(w->rml[i] + w->ll[i] * w->dtl[i]) / w->tl[i];
        w->dlu[i] = -(w->rmu[i] + w->lu[i] * w->dtu[i]) / w->tu[i];
    }
}

/* Iterative refinement of the Newton step just computed by ipm_step_from (HPIPM: itref_corr_max): the residual of the
 * stationarity rows of the reduced system, rho_k = Hd_k dz_k + q_k + BAt_k dpi_{k+1} - [dpi_k]_x, is fed back through the
 * same factorisation (dynamics rows hold to rounding by construction of the forward sweep), then dt / dlam are redone.
 * Returns the inf-norm of rho before the correction. */
static double SFX(ipm_refine)(SFX(ws_t) * w, int N)
{
    if (N < 1) return 0.0;
    const size_t n = (size_t)(N + 1) * NZ;
    double *dz0 = (double *)malloc(sizeof(double) * (n + (size_t)(N + 1) * NX + n + (size_t)N * NX));
    double *dpi0 = dz0 + n, *q0 = dpi0 + (size_t)(N + 1) * NX, *rb0 = q0 + n;
    double nrm = 0.0;
    for (size_t i = 0; i < n; i++) { dz0[i] = w->dz[i]; q0[i] = w->q[i]; }
    for (size_t i = 0; i < (size_t)(N + 1) * NX; i++) dpi0[i] = w->dpi[i];
    for (size_t i = 0; i < (size_t)N * NX; i++) rb0[i] = w->rb[i];
    for (int k = 0; k <= N; k++) {
        const double *BAt = w->BAt + (size_t)k * NZ * NX;
        for (int j = 0; j < NZ; j++) {
            const size_t i = (size_t)k * NZ + j;
            if (SKIP(k, j)) { w->q[i] = 0.0; continue; }
            double r = w->Hd[i] * dz0[i] + q0[i];
            if (k < N) for (int c = 0; c < NX; c++) r += BAt[j * NX + c] * dpi0[(size_t)(k + 1) * NX + c];
            if (j >= NU) r -= dpi0[(size_t)k * NX + j - NU];
            w->q[i] = r;
            nrm = fmax(nrm, fabs(r));
        }
    }
    for (size_t i = 0; i < (size_t)N * NX; i++) w->rb[i] = 0.0;
    SFX(ric_solve)(w, N);
    for (size_t i = 0; i < n; i++) { w->dz[i] += dz0[i]; w->q[i] = q0[i]; }
    for (size_t i = 0; i < (size_t)(N + 1) * NX; i++) w->dpi[i] += dpi0[i];
    for (size_t i = 0; i < (size_t)N * NX; i++) w->rb[i] = rb0[i];
    for (size_t i = 0; i < n; i++) {
        w->dtl[i] = w->dz[i] + w->rdl[i];
        w->dtu[i] = -w->dz[i] + w->rdu[i];
        w->dll[i] = -(w->rml[i] + w->ll[i] * w->dtl[i]) / w->tl[i];
        w->dlu[i] = -(w->rmu[i] + w->lu[i] * w->dtu[i]) / w->tu[i];
    }
    free(dz0);
    return nrm;
}

static double SFX(max_step)(const SFX(ws_t) * w, int N)
{
    const size_t n = (size_t)(N + 1) * NZ;
    double a = HUGE_VAL;
    for (size_t i = 0; i < n; i++) {
        if (w->lb[i] > -HUGE_VAL) {
            if (w->dtl[i] < 0) a = fmin(a, -w->tl[i] / w->dtl[i]);
            if (w->dll[i] < 0) a = fmin(a, -w->ll[i] / w->dll[i]);
        }
        if (w->ub[i] < HUGE_VAL) {
            if (w->dtu[i] < 0) a = fmin(a, -w->tu[i] / w->dtu[i]);
            if (w->dlu[i] < 0) a = fmin(a, -w->lu[i] / w->dlu[i]);
        }
    }
    return a;
}

static int SFX(ipm)(const orc_problem *P, SFX(ws_t) * w, int N, int *iters_out)
{
    const size_t n = (size_t)(N + 1) * NZ;
    const double thr0 = P->ipm_thr0, mu0 = P->ipm_mu0;
    int nb = 0;
    /* cold start [upstream D8]: z = 0 (except the pinned dx_0), pi = 0 */
    for (size_t i = 0; i < n; i++) {
        int k = (int)(i / NZ), j = (int)(i % NZ);
        if (!(k == 0 && j >= NU)) w->z[i] = 0.0;
        if (SKIP(k, j)) { w->lb[i] = -HUGE_VAL; w->ub[i] = HUGE_VAL; }
        int hl = w->lb[i] > -HUGE_VAL, hu = w->ub[i] < HUGE_VAL;
        nb += hl + hu;
        /* slack floor: absolute (thr0 > 0) or a fraction -thr0 of the box width (thr0 < 0) */
        const double flo = (thr0 >= 0.0) ? thr0 : -thr0 * (w->ub[i] - w->lb[i]);
        w->tl[i] = hl ? fmax(w->z[i] - w->lb[i], flo) : 1.0;
        w->tu[i] = hu ? fmax(w->ub[i] - w->z[i], flo) : 1.0;
        w->ll[i] = hl ? mu0 / w->tl[i] : 0.0;
        w->lu[i] = hu ? mu0 / w->tu[i] : 0.0;
        w->rdl[i] = w->rdu[i] = w->rml[i] = w->rmu[i] = 0.0;
    }
    for (size_t i = 0; i < (size_t)(N + 1) * NX; i++) w->pi[i] = 0.0;
    int status = 2, it;
    w->split_off = 0;
    double rg_est = 0.0, rb_est = 0.0, rd_est = 0.0;
    for (it = 0; it < P->ipm_max_iter; it++) {
        /* residuals */
        double res_g = 0, res_b = 0, res_d = 0, comp = 0, mu = 0;
        for (int k = 0; k <= N; k++) {
            const double *BAt = w->BAt + (size_t)k * NZ * NX;
            for (int j = 0; j < NZ; j++) {
                size_t i = (size_t)k * NZ + j;
                if (SKIP(k, j)) { w->rg[i] = 0; continue; }
                if (it == 0 || P->rg_mode >= 1 || P->strict) {
                    double r = w->H0[i] * w->z[i] + w->g[i] - w->ll[i] + w->lu[i];
                    if (k < N) for (int c = 0; c < NX; c++) r += BAt[j * NX + c] * w->pi[(size_t)(k + 1) * NX + c];
                    if (j >= NU) r -= w->pi[(size_t)k * NX + j - NU];
                    w->rg[i] = r;
                }
                res_g = fmax(res_g, fabs(w->rg[i]));
                if (w->lb[i] > -HUGE_VAL) {
                    w->rdl[i] = w->z[i] - w->lb[i] - w->tl[i];
                    res_d = fmax(res_d, fabs(w->rdl[i]));
                    mu += w->ll[i] * w->tl[i];
                    comp = fmax(comp, w->ll[i] * w->tl[i]);
                }
                if (w->ub[i] < HUGE_VAL) {
                    w->rdu[i] = w->ub[i] - w->z[i] - w->tu[i];
                    res_d = fmax(res_d, fabs(w->rdu[i]));
                    mu += w->lu[i] * w->tu[i];
                    comp = fmax(comp, w->lu[i] * w->tu[i]);
                }
            }
            if (k < N)
                for (int c = 0; c < NX; c++) {
                    double r = w->b[(size_t)k * NX + c] - w->z[(size_t)(k + 1) * NZ + NU + c];
                    for (int j = 0; j < NZ; j++) r += BAt[j * NX + c] * w->z[(size_t)k * NZ + j];
                    w->rb[(size_t)k * NX + c] = r;
                    res_b = fmax(res_b, fabs(r));
                }
        }
        if (nb) mu /= nb;
        if (it == 0) { rg_est = res_g; rb_est = res_b; rd_est = res_d; }
        /* the three linear residuals are affine in the iterate and everything takes the same
         * step, so each shrinks by exactly (1-alpha): the stopping test uses the stationarity norm
         * extrapolated from the start, and for the dynamics / bound-slack residuals the values
         * MEASURED on the previous iterate times the (1-alpha) of the step since */
        const int extrap = P->rg_mode == 2 && !P->strict;
        const double xb = res_b, xd = res_d;
        if (extrap) { res_g = rg_est; res_b = rb_est; res_d = rd_est; }
        if (orc_debug()) fprintf(stderr, "it %d res_g %.3e res_b %.3e res_d %.3e comp %.3e mu %.3e\n", it, res_g, res_b, res_d, comp, mu);
        if (!(res_g == res_g) || !(res_b == res_b) || !(mu == mu)) { status = 1; break; }
        /* diverging multipliers = infeasible QP: stop early, same status as the min-step exit */
        if (!P->strict && mu > 1e2 * mu0) { status = 3; break; }
        if (res_g <= P->tol_stat && res_b <= P->tol_eq && res_d <= P->tol_ineq && comp <= P->tol_comp) { status = 0; break; }
        /* factorise with barrier diagonal */
        for (size_t i = 0; i < n; i++) w->Hd[i] = w->H0[i] + w->ll[i] / w->tl[i] + w->lu[i] / w->tu[i];
        if (orc_debug()) {
            double gmax = 0, gxmax = 0;
            for (size_t i = 0; i < n; i++) { double gm = w->Hd[i] - w->H0[i]; if (gm > gmax) gmax = gm; if ((int)(i % NZ) >= NU && gm > gxmax) gxmax = gm; }
            fprintf(stderr, "   barrier max %.3e  (states %.3e)\n", gmax, gxmax);
        }
        if (SFX(ric_factor)(P, w, N, mu)) { status = 4; break; }
        double alpha = 1.0;
        if (nb) {
            /* predictor */
            for (size_t i = 0; i < n; i++) { w->rml[i] = w->ll[i] * w->tl[i]; w->rmu[i] = w->lu[i] * w->tu[i]; }
            SFX(ipm_step_from)(w, N);
            double a_aff = fmin(1.0, SFX(max_step)(w, N));
            double mu_aff = 0;
            for (size_t i = 0; i < n; i++) {
                if (w->lb[i] > -HUGE_VAL) mu_aff += (w->ll[i] + a_aff * w->dll[i]) * (w->tl[i] + a_aff * w->dtl[i]);
                if (w->ub[i] < HUGE_VAL) mu_aff += (w->lu[i] + a_aff * w->dlu[i]) * (w->tu[i] + a_aff * w->dtu[i]);
            }
            mu_aff /= nb;
            double sigma = mu_aff / mu; sigma = sigma * sigma * sigma;
            /* corrector + centering */
            for (size_t i = 0; i < n; i++) {
                w->rml[i] = (w->lb[i] > -HUGE_VAL) ? w->ll[i] * w->tl[i] + w->dll[i] * w->dtl[i] - sigma * mu : 0.0;
                w->rmu[i] = (w->ub[i] < HUGE_VAL) ? w->lu[i] * w->tu[i] + w->dlu[i] * w->dtu[i] - sigma * mu : 0.0;
            }
            SFX(ipm_step_from)(w, N);
            for (int rr = 0; rr < P->itref; rr++) {
                const double e = SFX(ipm_refine)(w, N);
                if (orc_debug()) fprintf(stderr, "   itref %d: |rho| %.3e\n", rr, e);
            }
            alpha = fmin(1.0, fmax(0.995, 1.0 - mu_aff) * SFX(max_step)(w, N));
        } else {
            for (size_t i = 0; i < n; i++) { w->rml[i] = 0; w->rmu[i] = 0; }
            SFX(ipm_step_from)(w, N);
        }
        for (size_t i = 0; i < n; i++) {
            int k = (int)(i / NZ), j = (int)(i % NZ);
            if (k == 0 && j >= NU) continue;
            w->z[i] += alpha * w->dz[i];
            if (w->lb[i] > -HUGE_VAL) { w->tl[i] += alpha * w->dtl[i]; w->ll[i] += alpha * w->dll[i]; }
            if (w->ub[i] < HUGE_VAL) { w->tu[i] += alpha * w->dtu[i]; w->lu[i] += alpha * w->dlu[i]; }
        }
        for (size_t i = NX; i < (size_t)(N + 1) * NX; i++) w->pi[i] += alpha * w->dpi[i];
        /* the stationarity residual is affine in (z, pi, lam) and all take the same step, so
         * r_g <- (1-alpha) r_g exactly; evaluating it explicitly needs the multipliers of
         * numerically pinned states, whose absolute accuracy degrades like eps*lam/t. */
        if (P->rg_mode == 0 && !P->strict) for (size_t i = 0; i < n; i++) w->rg[i] *= (1.0 - alpha);
        rg_est *= (1.0 - alpha); rb_est = xb * (1.0 - alpha); rd_est = xd * (1.0 - alpha);
        if (!(alpha >= P->alpha_min)) { status = (alpha == alpha) ? 3 : 1; it++; break; } /* [upstream D9] 3 = min step */
    }
    *iters_out = it;
    return status;
}

/* ---- A3-A8: one SQP_RTI iteration on one instance (simulation_blaster.py:60-89).
 * X[(N+1)*NX], U[N*NU] = persistent iterate, updated in place with the full step
 * (FIXED_STEP, step length 1.0).  yref: (N+1)*NY if yref_per_stage else NY;
 * p: N*25 if p_per_stage else 25. */
static int SFX(rti_solve)(const orc_problem *P, double *X, double *U, const double *x0, const double *yref,
                          int yref_per_stage, const double *p, int p_per_stage, double *wsmem, int *iters)
{
    const int N = P->N, NY = NZ;
    SFX(ws_t) w;
    SFX(ws_bind)(&w, wsmem, N);
    for (int k = 0; k < N; k++) {
        const double *pk = p + (p_per_stage ? (size_t)k * 25 : 0);
        double xn[NX];
        SFX(rk4_sens)(P, X + (size_t)k * NX, U + (size_t)k * NU, pk, xn, w.BAt + (size_t)k * NZ * NX);
        for (int i = 0; i < NX; i++) w.b[(size_t)k * NX + i] = xn[i] - X[(size_t)(k + 1) * NX + i];
    }
    for (int k = 0; k <= N; k++) {
        const double *yr = yref + (yref_per_stage ? (size_t)k * NY : 0);
        double *g = w.g + (size_t)k * NZ, *H = w.H0 + (size_t)k * NZ, *lb = w.lb + (size_t)k * NZ, *ub = w.ub + (size_t)k * NZ;
        for (int j = 0; j < NU; j++) {
            if (k < N) {
                /* [upstream D1] stage cost x dt */
                H[j] = P->dt * P->R[j];
                g[j] = P->dt * P->R[j] * (U[(size_t)k * NU + j] - yr[NX + j]);
                lb[j] = P->lbu[j] - U[(size_t)k * NU + j];
                ub[j] = P->ubu[j] - U[(size_t)k * NU + j];
            } else { H[j] = 1.0; g[j] = 0; lb[j] = -HUGE_VAL; ub[j] = HUGE_VAL; }
        }
        for (int i = 0; i < NX; i++) {
            double xi = X[(size_t)k * NX + i];
            double wgt = (k < N) ? P->dt * P->Q[i] : P->Qt[i];
            H[NU + i] = wgt;
            g[NU + i] = wgt * (xi - yr[i]);
            /* [upstream D2] state bounds on stages 1..N-1 only */
            if (k >= 1 && k < N) { lb[NU + i] = P->lbx[i] - xi; ub[NU + i] = P->ubx[i] - xi; }
            else { lb[NU + i] = -HUGE_VAL; ub[NU + i] = HUGE_VAL; }
        }
    }
    /* [upstream D3] x0 pinned: dx_0 = x0 - X_0 */
    for (int i = 0; i < NX; i++) w.z[NU + i] = x0[i] - X[i];
    int status = SFX(ipm)(P, &w, N, iters);
    /* a failed QP leaves the iterate untouched (same rule as the product); with reference
     * semantics the last iterate is applied when the iteration cap was hit, as acados does */
    const int take = status == 0 || (P->strict && status == 2);
    for (int k = 0; k <= (take ? N : -1); k++) {
        if (k < N) for (int j = 0; j < NU; j++) U[(size_t)k * NU + j] += w.z[(size_t)k * NZ + j];
        for (int i = 0; i < NX; i++) X[(size_t)k * NX + i] += w.z[(size_t)k * NZ + NU + i];
    }
    return status;
}
/* ---- multi-iteration SQP to convergence on one instance (acados' SQP loop with the options the
 * reference's dump carries: nlp_solver_max_iter, nlp_solver_tol_stat/eq/ineq/comp).  Each iteration: linearise, evaluate
 * the NLP residuals with the multipliers of the previous QP (zero at the start [upstream D4]), stop if all four are within
 * tolerance, else solve the QP and take the full step.  One more evaluation after the max_iter-th QP reports the final
 * iterate's residuals (and counts as convergence if they pass).  res[4] = {stat, eq, ineq, comp}.
 * Returns 0 converged, 2 max_iter, 4 QP failure, 1 NaN; *sqp_iters = QPs solved, *qp_iters = IPM iterations in total. */
static int SFX(sqp_solve)(const orc_problem *P, double *X, double *U, const double *x0, const double *yref, int yref_per_stage,
                          const double *p, int p_per_stage, int max_iter, const double *tol, double *wsmem, int *sqp_iters,
                          int *qp_iters, double *res)
{
    const int N = P->N, NY = NZ;
    SFX(ws_t) w;
    SFX(ws_bind)(&w, wsmem, N);
    const size_t n = (size_t)(N + 1) * NZ;
    for (size_t i = 0; i < n; i++) { w.ll[i] = 0.0; w.lu[i] = 0.0; }
    for (size_t i = 0; i < (size_t)(N + 1) * NX; i++) w.pi[i] = 0.0;
    *sqp_iters = 0; *qp_iters = 0;
    for (int it = 0; it <= max_iter; it++) {
        /* linearisation at the current iterate */
        for (int k = 0; k < N; k++) {
            const double *pk = p + (p_per_stage ? (size_t)k * 25 : 0);
            double xn[NX];
            SFX(rk4_sens)(P, X + (size_t)k * NX, U + (size_t)k * NU, pk, xn, w.BAt + (size_t)k * NZ * NX);
            for (int i = 0; i < NX; i++) w.b[(size_t)k * NX + i] = xn[i] - X[(size_t)(k + 1) * NX + i];
        }
        double rs = 0, re = 0, ri = 0, rc = 0;
        for (int k = 0; k <= N; k++) {
            const double *yr = yref + (yref_per_stage ? (size_t)k * NY : 0);
            const double *BAt = w.BAt + (size_t)k * NZ * NX;
            for (int j = 0; j < NZ; j++) {
                if (SKIP(k, j)) continue;
                const size_t i = (size_t)k * NZ + j;
                const int isu = j < NU;
                const double y = isu ? U[(size_t)k * NU + j] : X[(size_t)k * NX + j - NU];
                const double yv = isu ? yr[NX + j] : yr[j - NU];
                const double wgt = isu ? P->dt * P->R[j] : (k < N ? P->dt * P->Q[j - NU] : P->Qt[j - NU]);
                const int hasb = isu ? 1 : (k >= 1 && k < N);
                double r = wgt * (y - yv);
                if (k < N) for (int c = 0; c < NX; c++) r += BAt[j * NX + c] * w.pi[(size_t)(k + 1) * NX + c];
                if (!isu) r -= w.pi[(size_t)k * NX + j - NU];
                if (hasb) {
                    const double lb = isu ? P->lbu[j] : P->lbx[j - NU], ub = isu ? P->ubu[j] : P->ubx[j - NU];
                    r += w.lu[i] - w.ll[i];
                    ri = fmax(ri, fmax(lb - y, y - ub));
                    rc = fmax(rc, fmax(fabs(w.ll[i] * (y - lb)), fabs(w.lu[i] * (ub - y))));
                }
                rs = fmax(rs, fabs(r));
            }
            if (k < N) for (int c = 0; c < NX; c++) re = fmax(re, fabs(w.b[(size_t)k * NX + c]));
            if (k == 0) for (int c = 0; c < NX; c++) re = fmax(re, fabs(x0[c] - X[c]));
        }
        if (res) { res[0] = rs; res[1] = re; res[2] = ri; res[3] = rc; }
        if (!(rs == rs) || !(re == re)) return 1;
        if (rs <= tol[0] && re <= tol[1] && ri <= tol[2] && rc <= tol[3]) return 0;
        if (it == max_iter) break;
        /* QP of this iterate (same assembly as rti_solve) */
        int qit = 0;
        const int st = SFX(rti_solve)(P, X, U, x0, yref, yref_per_stage, p, p_per_stage, wsmem, &qit);
        *sqp_iters += 1; *qp_iters += qit;
        if (!(st == 0 || (P->strict && st == 2))) return st == 1 ? 1 : 4;
        /* multipliers of components without bounds are not part of the QP: zero them for the residual */
        for (size_t i = 0; i < n; i++) {
            int k = (int)(i / NZ), j = (int)(i % NZ);
            const int hasb = !SKIP(k, j) && (j < NU || (k >= 1 && k < N));
            if (!hasb) { w.ll[i] = 0.0; w.lu[i] = 0.0; }
        }
    }
    return 2;
}
#undef SKIP
#undef NZ
