// Throughput variant of the QP solve: FOUR MPC INSTANCES PER WARP, eight lanes each.
//
// Same algorithm, same workspace records and the same results as mpcb_qp.cuh (one instance per
// warp); what changes is the mapping.  There lane j owns row j of the stage matrices, so 9 of 32
// lanes idle, half of the remaining ones idle in the triangular Householder loop, and the ~60
// scalar instructions of a pivot (shuffles, rsqrt, rcp, selects) and the shared-memory broadcast
// of the pivot row are paid per instance.  ncu on the one-instance kernel shows the shared-memory
// data pipe, the FP64 pipe and the issue slots all moderately loaded and none saturated: the
// cost is instructions per instance.  Here sub-lane s of an 8-lane group owns rows s, s+8, s+16
// (cyclic, so the triangular loop keeps every lane busy until the last rows), one warp
// instruction serves four instances, and a pivot costs one broadcast and one scalar chain per
// FOUR instances.  Finished row slots are skipped with compile-time trip counts.
//
// Used by the host scheduler for chunks of many waves; small batches keep the latency kernel.
#pragma once
#include "mpcb_qp.cuh"

namespace mpcb {

constexpr int kLPI = 8;  // lanes per instance
constexpr int kGPW = 4;  // instances (groups) per warp

template <int NX, int NU>
struct Qp8Group {
    using L = Layout<NX, NU>;
    alignas(16) double rec[L::O_Z];  // head of a stage record: [BAt | Lu | invd | lvec | rb]
    double lxx[L::LXX];              // factor of P_{k+1} / P_k, row-major, zero upper triangle
    // Arrays that are never live at the same time share storage (BLASTER17: 36.0 -> 31.9 KB per warp, 7 instead of 6
    // warps per SM; QUAD12: 19.1 -> 16.8 KB, 12 instead of 11; every hand-over has a warp_sync between the last read
    // of one and the first write of the other):
    union {
        double vrow[2][L::NXP];      // pivot row broadcast (double buffered): live inside the Householder loop only
        struct {
            double sT1[L::NXP];      // temporary L' r of t2 = P r + p: dead before the loop
            double cPv[L::NXP];      // carried p_{k+1}: consumed by t2 before the loop, rewritten after it
        };
    };
    union {
        double vz[L::NZP];           // a stage vector every lane of the group reads (z_k or dz_k): dead once r_k and the carry are formed
        double hd[L::NZP];           // Hd_k of every row, written after the W product for the pivot loop
    };
    union {
        double cPi[L::NXP];          // backward sweep S1: pi_{k+1}
        double cDx[L::NXP];          // forward sweeps: dx_k (zeroed at their start)
    };
    union {
        double sRb[L::NXP];          // r_k (S1) / dx_{k+1} (S4): last read when sT1 is formed ...
        double sT2[L::NXP];          // ... before t2 = P r + p is written (a sync lies between)
    };
    double cZx[L::NXP];              // carried dx-part of z_{k+1}
    // sqrt(Hd_k) of every row (pivot loop) lives in the [lvec | rb] slice of `rec`, which the backward sweep S1 does not use
    static_assert(L::O_Z - L::O_LVEC >= L::NZP, "rec[lvec|rb] must hold one entry per row");
    MPCB_HD double *ds() { return rec + L::O_LVEC; }
    // Pad so that consecutive groups are 32 bytes (mod 128) apart: the four groups of a warp then hit
    // distinct banks when each broadcasts one word to its lanes, and a 64-byte run per group splits into
    // the minimal two wavefronts.  (QUAD12's unpadded group is a multiple of 128 bytes: every broadcast
    // was a 4-way bank conflict, 43 % of all shared-memory wavefronts in ncu.)
    static constexpr int kBody = L::O_Z + L::LXX + 2 * L::NXP + L::NZP + 3 * L::NXP;
    double pad[(4 - kBody % 16 + 16) % 16 == 0 ? 16 : (4 - kBody % 16 + 16) % 16];
};
static_assert(sizeof(Qp8Group<17, 6>) % 128 == 32 && sizeof(Qp8Group<12, 4>) % 128 == 32 && sizeof(Qp8Group<13, 4>) % 128 == 32,
              "group stride must be 32 mod 128 bytes");
template <int NX, int NU>
struct Qp8Smem {
    Qp8Group<NX, NU> g[kGPW];
    unsigned long long mbar[kGPW];  // one mbarrier per group for its bulk copies
};

MPCB_DEV double grp_max(double v)
{
    v = fmax(v, warp_shfl_xor(v, 4)); v = fmax(v, warp_shfl_xor(v, 2)); v = fmax(v, warp_shfl_xor(v, 1));
    return v;
}
MPCB_DEV double grp_sum(double v)
{
    v += warp_shfl_xor(v, 4); v += warp_shfl_xor(v, 2); v += warp_shfl_xor(v, 1);
    return v;
}
MPCB_DEV double grp_bcast(double v, int sub) { return warp_shfl(v, (lane_id() & ~(kLPI - 1)) | sub); }

// ---- per-group staging of record runs into shared memory: TMA bulk copies (cp.async.bulk) issued
// by sub-lane 0 of each group on the group's own mbarrier; the other lanes of the group wait on its
// phase.  One instruction per run instead of a load/store loop, no registers, and the copy of the
// next stage's matrices overlaps the Householder loop of the current one.
struct GrpPipe {
    unsigned mbar;   // shared address of the group's mbarrier
    unsigned phase;  // parity the next completed fetch will have
};
#ifndef MPCB_HOST_EMU
MPCB_DEV void g8_init(GrpPipe &p, unsigned long long *mbar, bool leader)
{
    p.mbar = (unsigned)__cvta_generic_to_shared(mbar);
    p.phase = 0u;
    if (leader) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(p.mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
}
MPCB_DEV void g8_expect(const GrpPipe &p, int ndoubles, bool issue)
{
    if (issue) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(p.mbar), "r"(8 * ndoubles) : "memory");
}
MPCB_DEV void g8_copy(const GrpPipe &p, double *smem_dst, const double *gmem_src, int ndoubles, bool issue)
{
    if (issue)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                         (unsigned)__cvta_generic_to_shared(smem_dst)),
                     "l"(gmem_src), "r"(8 * ndoubles), "r"(p.mbar)
                     : "memory");
}
MPCB_DEV void g8_wait(GrpPipe &p, bool fetched)
{
    if (fetched) {
        unsigned done = 0;
        for (int spin = 0; !done; spin++) {
            asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n selp.u32 %0, 1, 0, q;\n}"
                         : "=r"(done)
                         : "r"(p.mbar), "r"(p.phase)
                         : "memory");
            if (spin > (1 << 22)) __trap();  // a lost transaction must fail loudly, never hang the GPU
        }
        p.phase ^= 1u;
    }
    __syncwarp();
}
#else
MPCB_DEV void g8_init(GrpPipe &, unsigned long long *, bool) {}
MPCB_DEV void g8_expect(const GrpPipe &, int, bool) {}
MPCB_DEV void g8_copy(const GrpPipe &, double *smem_dst, const double *gmem_src, int n, bool issue)
{
    if (issue) for (int i = 0; i < n; i++) smem_dst[i] = gmem_src[i];
}
MPCB_DEV void g8_wait(GrpPipe &, bool) { warp_sync(); }
#endif

// y_t += sum_j M[j * ld + i_t] * x[j], j < n: columns i_t of a row-major matrix in shared memory
// against a shared vector, for the lane's NT entries; a rolled loop (small code, 2 accumulators).
template <int NT_>
MPCB_DEV void g8_matvec_cols(const double *M, int ld, const double *x, int n, const int (&idx)[NT_], double (&acc)[NT_])
{
    double a1[NT_];
    MPCB_UNROLL
    for (int t = 0; t < NT_; t++) a1[t] = 0.0;
    int j = 0;
    MPCB_NOUNROLL
    for (; j + 1 < n; j += 2) {
        const double x0 = x[j], x1 = x[j + 1];
        MPCB_UNROLL
        for (int t = 0; t < NT_; t++) {
            acc[t] += M[j * ld + idx[t]] * x0;
            a1[t] += M[(j + 1) * ld + idx[t]] * x1;
        }
    }
    if (j < n) {
        const double x0 = x[j];
        MPCB_UNROLL
        for (int t = 0; t < NT_; t++) acc[t] += M[j * ld + idx[t]] * x0;
    }
    MPCB_UNROLL
    for (int t = 0; t < NT_; t++) acc[t] += a1[t];
}
// y_t += sum_c M[idx_t * ld + c] * x[c], c < n: rows idx_t against a shared vector
template <int NT_>
MPCB_DEV void g8_matvec_rows(const double *M, int ld, const double *x, int n, const int (&idx)[NT_], double (&acc)[NT_])
{
    double a1[NT_];
    MPCB_UNROLL
    for (int t = 0; t < NT_; t++) a1[t] = 0.0;
    int c = 0;
    MPCB_NOUNROLL
    for (; c + 1 < n; c += 2) {
        const double x0 = x[c], x1 = x[c + 1];
        MPCB_UNROLL
        for (int t = 0; t < NT_; t++) {
            acc[t] += M[idx[t] * ld + c] * x0;
            a1[t] += M[idx[t] * ld + c + 1] * x1;
        }
    }
    if (c < n) {
        const double x0 = x[c];
        MPCB_UNROLL
        for (int t = 0; t < NT_; t++) acc[t] += M[idx[t] * ld + c] * x0;
    }
    MPCB_UNROLL
    for (int t = 0; t < NT_; t++) acc[t] += a1[t];
}

// Forward substitution of the lane's row slots with the NU factor columns of the stage (record image in shared memory):
// l_u goes through the group's scratch once and EVERY lane runs the NU-step recurrence redundantly in registers -- the same
// operations in the same order as a chain of NU group broadcasts (bit-identical), without the NU dependent shuffles
// (WARPSYNC + 2 SHFL each, and a reconvergence before every one of them).
template <int NX, int NU, int NT_>
MPCB_DEV void g8_fwd_subst(Qp8Group<NX, NU> &sm, double (&l)[NT_], int s)
{
    using L = Layout<NX, NU>;
    static_assert(NU <= kLPI, "the input rows must sit in the first row slot");
    if (s < NU) sm.sT1[s] = l[0];
    warp_sync();
    double lv[NU];
    MPCB_UNROLL
    for (int c = 0; c < NU; c++) {
        double a = sm.sT1[c];
        MPCB_UNROLL
        for (int cc = 0; cc < c; cc++) a -= sm.rec[L::O_LU + cc * L::NZP + c] * lv[cc];
        lv[c] = a * sm.rec[L::O_INVD + c];
    }
    MPCB_UNROLL
    for (int t = 0; t < NT_; t++) {
        const int row = s + kLPI * t;
        if (row >= NU && row < L::NZ) {
            MPCB_UNROLL
            for (int c = 0; c < NU; c++) l[t] -= sm.rec[L::O_LU + c * L::NZP + row] * lv[c];
        }
    }
    MPCB_UNROLL
    for (int c = 0; c < NU; c++)
        if (s == c) l[0] = lv[c];
}

// Forward sweep (affine: FINAL = false, full step: FINAL = true), see forward_sweep in mpcb_qp.cuh.
template <int NX, int NU, bool FINAL>
MPCB_DEV void qp8_forward(const Params &P, Qp8Group<NX, NU> &sm, GrpPipe &pipe, double *__restrict__ ws, bool run, double sigmu,
                          double &imax_out, double &acc1_out, double &acc2_out)
{
    using L = Layout<NX, NU>;
    constexpr int NZ = L::NZ, NT = (NZ + kLPI - 1) / kLPI, NXT = (NX + kLPI - 1) / kLPI;
    constexpr int O_OUT = FINAL ? L::O_DZ : L::O_DZA;
    const int s = lane_id() & (kLPI - 1);
    const int N = P.N;
    const bool issue = run && s == 0;
    int xi[NXT];
    MPCB_UNROLL
    for (int t = 0; t < NXT; t++) xi[t] = (s + kLPI * t < NX) ? s + kLPI * t : 0;
    double imax = 0.0, acc1 = 0.0, acc2 = 0.0;
    MPCB_UNROLL
    for (int t = 0; t < NXT; t++)
        if (s + kLPI * t < NX) sm.cDx[s + kLPI * t] = 0.0;
    for (int k = 0; k < N; k++) {
        double *wk = ws + (size_t)k * L::STAGE;
        warp_sync();  // previous stage's readers of rec / vz / lxx are done
        g8_expect(pipe, L::O_Z + (FINAL ? L::LXX : 0), issue);
        g8_copy(pipe, sm.rec, wk, L::O_Z, issue);
        if (FINAL) g8_copy(pipe, sm.lxx, wk + L::STAGE + L::O_LXX, L::LXX, issue);
        if (k + 1 < N) {  // what the next stage will fetch: HBM -> L2
            l2_prefetch(wk + L::STAGE, L::O_G, issue);
            if (FINAL) {
                l2_prefetch(wk + L::STAGE + L::O_DZA, L::NZP, issue);
                l2_prefetch(wk + 2 * L::STAGE + L::O_LXX, L::LXX + L::NXP, issue);
            }
        }
        // this stage's box data of the lane's own rows: in flight beside the bulk copy
        double bz[NT], btl[NT], btu[NT], bll[NT], blu[NT], blb[NT], bub[NT], bdza[NT];
        bool bhb[NT];
        MPCB_UNROLL
        for (int t = 0; t < NT; t++) {
            const int row = s + kLPI * t;
            bhb[t] = run && row < NZ && var_kind<NX, NU>(k, row, N).hasb;
            bz[t] = bll[t] = blu[t] = blb[t] = bub[t] = bdza[t] = 0.0;
            btl[t] = btu[t] = 1.0;
            if (bhb[t]) {
                bz[t] = wk[L::O_Z + row]; btl[t] = wk[L::O_TL + row]; btu[t] = wk[L::O_TU + row];
                bll[t] = wk[L::O_LL + row]; blu[t] = wk[L::O_LUP + row]; blb[t] = wk[L::O_LB + row]; bub[t] = wk[L::O_UB + row];
                if (FINAL) bdza[t] = wk[L::O_DZA + row];
            }
        }
        double pvn[NXT];
        MPCB_UNROLL
        for (int t = 0; t < NXT; t++) pvn[t] = (FINAL && run && s + kLPI * t < NX) ? wk[L::STAGE + L::O_PV + s + kLPI * t] : 0.0;
        g8_wait(pipe, run);
        // du = -Luu^{-T} (lvec + Lxu' dx)
        double yy = 0.0;
        if (s < NU) {
            double a0 = sm.rec[L::O_LVEC + s], a1 = 0.0;
            const double *col = sm.rec + L::O_LU + s * L::NZP + NU;
            MPCB_UNROLL
            for (int i = 0; i + 1 < NX; i += 2) { a0 += col[i] * sm.cDx[i]; a1 += col[i + 1] * sm.cDx[i + 1]; }
            if (NX & 1) a0 += col[NX - 1] * sm.cDx[NX - 1];
            yy = -(a0 + a1);
        }
        // back substitution du = Luu^{-T} yy, redundantly in every lane of the group (see g8_fwd_subst)
        if (s < NU) sm.sT1[s] = yy;
        warp_sync();
        double du = 0.0;
        {
            double duv[NU];
            MPCB_UNROLL
            for (int i = NU - 1; i >= 0; i--) {
                double a = sm.sT1[i];
                MPCB_UNROLL
                for (int c = NU - 1; c > i; c--) a -= sm.rec[L::O_LU + i * L::NZP + c] * duv[c];
                duv[i] = a * sm.rec[L::O_INVD + i];
            }
            MPCB_UNROLL
            for (int i = 0; i < NU; i++)
                if (s == i) du = duv[i];
        }
        MPCB_UNROLL
        for (int t = 0; t < NT; t++) {
            const int row = s + kLPI * t;
            if (row >= NZ) continue;
            const double dz = (row < NU) ? du : sm.cDx[row - NU];
            if (run) wk[O_OUT + row] = dz;
            sm.vz[row] = dz;
            if (bhb[t]) {
                const double z = bz[t], tl = btl[t], tu = btu[t], ll = bll[t], lu = blu[t], lb = blb[t], ub = bub[t];
                const double itl = fast_rcp(tl), itu = fast_rcp(tu);
                double rml = ll * tl, rmu = lu * tu;
                if (FINAL) {
                    const BoxStep<double> a = box_step(z, bdza[t], lb, ub, tl, tu, ll, lu, rml, rmu, itl, itu);
                    rml += a.dll * a.dtl - sigmu;
                    rmu += a.dlu * a.dtu - sigmu;
                }
                const BoxStep<double> b = box_step(z, dz, lb, ub, tl, tu, ll, lu, rml, rmu, itl, itu);
                imax = fmax(imax, fmax(fmax(inv_ratio(b.dtl, itl), inv_ratio(b.dtu, itu)),
                                       fmax(inv_ratio(b.dll, fast_rcp(ll)), inv_ratio(b.dlu, fast_rcp(lu)))));
                if (!FINAL) {
                    acc1 += ll * b.dtl + tl * b.dll + lu * b.dtu + tu * b.dlu;
                    acc2 += b.dll * b.dtl + b.dlu * b.dtu;
                    wk[L::O_C1 + row] = b.dll * b.dtl * itl - b.dlu * b.dtu * itu;
                    wk[L::O_C2 + row] = itl - itu;
                }
            }
        }
        warp_sync();
        // dx_{k+1} = rb_k + [B A] dz_k
        double dxn[NXT];
        MPCB_UNROLL
        for (int t = 0; t < NXT; t++) dxn[t] = sm.rec[L::O_RB + xi[t]];
        g8_matvec_cols<NXT>(sm.rec + L::O_BAT, L::LDB, sm.vz, NZ, xi, dxn);
        MPCB_UNROLL
        for (int t = 0; t < NXT; t++) {
            const int i = s + kLPI * t;
            if (i < NX) { sm.cDx[i] = dxn[t]; if (FINAL) sm.sRb[i] = dxn[t]; }  // cDx was last read before the sync above
        }
        if (FINAL) {
            // dpi_{k+1} = Lxx_{k+1} (Lxx_{k+1}' dx_{k+1}) + p_{k+1}
            warp_sync();
            double t1[NXT];
            MPCB_UNROLL
            for (int t = 0; t < NXT; t++) t1[t] = 0.0;
            g8_matvec_cols<NXT>(sm.lxx, NX, sm.sRb, NX, xi, t1);
            MPCB_UNROLL
            for (int t = 0; t < NXT; t++)
                if (s + kLPI * t < NX) sm.sT1[s + kLPI * t] = t1[t];
            warp_sync();
            g8_matvec_rows<NXT>(sm.lxx, NX, sm.sT1, NX, xi, pvn);
            MPCB_UNROLL
            for (int t = 0; t < NXT; t++)
                if (run && s + kLPI * t < NX) wk[L::STAGE + L::O_DPI + s + kLPI * t] = pvn[t];
        }
    }
    warp_sync();
    // terminal stage: dz_N = [0; dx_N]
    MPCB_UNROLL
    for (int t = 0; t < NT; t++) {
        const int row = s + kLPI * t;
        if (run && row < NZ) ws[(size_t)N * L::STAGE + O_OUT + row] = (row < NU) ? 0.0 : sm.cDx[row - NU];
    }
    pipe_fence();
    imax_out = grp_max(imax);
    acc1_out = grp_sum(acc1);
    acc2_out = grp_sum(acc2);
}

// What the groups of a warp draw their instances from: one chunk of a batch (instances inst0 .. inst0+B-1
// of the persistent iterate, chunk-local workspace records) and the chunk's work counter.
struct Qp8Batch {
    double *X, *U;             // persistent iterate of the whole batch
    const double *x0, *yref;
    size_t yref_stride;        // doubles between the yref of consecutive instances (0: shared)
    int yps;                   // yref is per stage
    double *ws;                // workspace of the chunk, linearised by K1
    double *u0;                // optional outputs (whole batch)
    int32_t *status, *iters;
    int inst0, B;
    unsigned *next;            // work counter of the chunk (zero at launch)
};

// QP data and cold start of one instance (see F0 in mpcb_qp.cuh), for the groups with `act` set;
// warp-uniform control flow, so the group reductions at the end are executed by every lane.
template <int NX, int NU>
MPCB_DEV void qp8_setup(const Params &P, double *__restrict__ ws, const double *__restrict__ Xi, const double *__restrict__ Ui,
                        const double *__restrict__ x0, const double *__restrict__ yref, int yps, bool act, double &est_g, double &est_b,
                        double &est_d)
{
    using L = Layout<NX, NU>;
    constexpr int NZ = L::NZ, NT = (NZ + kLPI - 1) / kLPI, NXT = (NX + kLPI - 1) / kLPI;
    const int s = lane_id() & (kLPI - 1);
    const int N = P.N;
    const double thr0 = P.ipm_thr0, mu0 = P.ipm_mu0;
    double eg = 0.0, eb = 0.0, ed = 0.0;
    for (int k = 0; k <= N; k++) {
        double *wk = ws + (size_t)k * L::STAGE;
        const double *yr = yref + (yps ? (size_t)k * NZ : 0);
        MPCB_UNROLL
        for (int t = 0; t < NT; t++) {
            const int row = s + kLPI * t;
            if (row >= NZ || !act) continue;
            const VarKind vk = var_kind<NX, NU>(k, row, N);
            const double H0 = hess_diag<NX, NU, double>(P, k, row);
            double y = 0.0, g = 0.0, lb = 0.0, ub = 0.0, z = 0.0;
            if (row < NU) {
                if (k < N) {
                    y = Ui[(size_t)k * NU + row];
                    g = H0 * (y - yr[NX + row]);
                    lb = P.lbu[row] - y; ub = P.ubu[row] - y;
                }
            } else {
                const int i = row - NU;
                y = Xi[(size_t)k * NX + i];
                g = H0 * (y - yr[i]);
                if (vk.hasb) { lb = P.lbx[i] - y; ub = P.ubx[i] - y; }
                if (k == 0) z = x0[i] - y;
            }
            double tl = 1.0, tu = 1.0, ll = 0.0, lu = 0.0;
            if (vk.hasb) {
                const double flo = (thr0 >= 0.0) ? thr0 : -thr0 * (ub - lb);
                tl = fmax(z - lb, flo);
                tu = fmax(ub - z, flo);
                ll = mu0 / tl;
                lu = mu0 / tu;
                ed = fmax(ed, fmax(fabs(z - lb - tl), fabs(ub - z - tu)));
            }
            if (vk.var) eg = fmax(eg, fabs(H0 * z + g - ll + lu));
            wk[L::O_Z + row] = z; wk[L::O_TL + row] = tl; wk[L::O_TU + row] = tu;
            wk[L::O_LL + row] = ll; wk[L::O_LUP + row] = lu;
            wk[L::O_LB + row] = lb; wk[L::O_UB + row] = ub; wk[L::O_G + row] = g;
        }
        MPCB_UNROLL
        for (int t = 0; t < NXT; t++) {
            const int i = s + kLPI * t;
            if (i >= NX || !act) continue;
            wk[L::O_PI + i] = 0.0;
            if (k < N) {
                double rb = wk[L::O_B + i];
                if (k == 0)
                    for (int c = 0; c < NX; c++) rb += wk[L::O_BAT + (NU + c) * L::LDB + i] * (x0[c] - Xi[c]);
                eb = fmax(eb, fabs(rb));
            }
        }
    }
    eg = grp_max(eg); eb = grp_max(eb); ed = grp_max(ed);
    if (act) { est_g = eg; est_b = eb; est_d = ed; }
}

// The QP solves of a chunk, by one warp: every 8-lane group draws instances from the chunk's work
// counter, solves them one after the other, and is refilled at an IPM-iteration boundary as soon as
// its instance ends ("continuous batching").  The four groups of a warp execute the same instruction
// stream, so without the refill a warp would run for the slowest of its four instances (mean of the
// maximum of four iteration counts: 12.3 against a mean of 10.8 on the QUAD12 bench batch, +13.5 %).
// Results do not depend on which group solves an instance.
template <int NX, int NU>
MPCB_DEV void qp8_solve_queue(const Params &P, Qp8Smem<NX, NU> &smw, const Qp8Batch &job)
{
    using L = Layout<NX, NU>;
    constexpr int NZ = L::NZ, NT = (NZ + kLPI - 1) / kLPI, NXT = (NX + kLPI - 1) / kLPI;
    static_assert(NU <= kLPI, "the input block must fit one row slot");
    const int lane = lane_id();
    const int s = lane & (kLPI - 1);
    Qp8Group<NX, NU> &sm = smw.g[lane >> 3];
    const int N = P.N;
    const double mu0 = P.ipm_mu0;
    const double nb = (double)(2 * NU * N + 2 * NX * (N - 1));

    for (int idx = s; idx < L::LXX; idx += kLPI) sm.lxx[idx] = 0.0;
    GrpPipe pipe;
    g8_init(pipe, &smw.mbar[lane >> 3], s == 0);
    int xi[NXT];
    MPCB_UNROLL
    for (int t = 0; t < NXT; t++) xi[t] = (s + kLPI * t < NX) ? s + kLPI * t : 0;

    // the group's instance in flight (pointers stay valid when there is none: nothing is accessed through them then)
    double *__restrict__ ws = job.ws;
    double *__restrict__ Xi = job.X + (size_t)job.inst0 * (N + 1) * NX;
    double *__restrict__ Ui = job.U + (size_t)job.inst0 * N * NU;
    const double *__restrict__ x0 = job.x0 + (size_t)job.inst0 * NX;
    const double *__restrict__ yref = job.yref;
    int inst = job.inst0;
    bool has = false, done = false, drained = false;
    double est_g = 0.0, est_b = 0.0, est_d = 0.0;
    double comp = mu0, mu = mu0;
    int status = ST_MAXITER, git = 0;  // git: IPM iterations of the instance in flight

    for (;;) {
        // ---- the stopping tests of mpcb_qp.cuh, in the same order (an instance that uses up its iterations is not tested again)
        auto test = [&]() {
            if (has && !done) {
                if (git >= P.ipm_max_iter) { status = ST_MAXITER; done = true; }
                else if (!(est_g == est_g) || !(est_b == est_b) || !(mu == mu)) { status = ST_NAN; done = true; }
                else if (mu > kMuDiverge * mu0) { status = ST_MINSTEP; done = true; }
                else if (est_g <= P.tol_stat && est_b <= P.tol_eq && est_d <= P.tol_ineq && comp <= P.tol_comp) { status = ST_OK; done = true; }
            }
        };
        test();
        // ---- retire: RTI update X += dx, U += du (full step); a failed QP leaves the iterate untouched
        if (has && done) {
            if (status == ST_OK) {
                for (int k = 0; k <= N; k++) {
                    const double *wk = ws + (size_t)k * L::STAGE;
                    MPCB_UNROLL
                    for (int t = 0; t < NT; t++) {
                        const int row = s + kLPI * t;
                        if (row >= NZ) continue;
                        if (row < NU) { if (k < N) Ui[(size_t)k * NU + row] += wk[L::O_Z + row]; }
                        else Xi[(size_t)k * NX + row - NU] += wk[L::O_Z + row];
                    }
                }
            }
            if (s == 0) {
                if (job.status) job.status[inst] = status;
                if (job.iters) job.iters[inst] = git;
            }
            if (job.u0 && s < NU) job.u0[(size_t)inst * NU + s] = Ui[s];  // Ui[s] was updated by this very lane
            has = false;
        }
        // ---- refill from the chunk's work counter
        const bool want = !has && !drained;
        if (warp_or(want ? 1 : 0)) {
            int idx = 0;
            if (want && s == 0) idx = (int)queue_take(job.next);
            idx = warp_shfl(idx, lane & ~(kLPI - 1));
            const bool fresh = want && idx < job.B;
            if (want && !fresh) drained = true;
            if (fresh) {
                inst = job.inst0 + idx;
                ws = job.ws + (size_t)idx * L::instance_stride(N);
                Xi = job.X + (size_t)inst * (N + 1) * NX;
                Ui = job.U + (size_t)inst * N * NU;
                x0 = job.x0 + (size_t)inst * NX;
                yref = job.yref + (size_t)inst * job.yref_stride;
            }
            qp8_setup<NX, NU>(P, ws, Xi, Ui, x0, yref, job.yps, fresh, est_g, est_b, est_d);
            if (fresh) { comp = mu0; mu = mu0; status = ST_MAXITER; git = 0; has = true; done = false; }
            pipe_fence();  // the QP data written above is fetched by bulk copies
            warp_sync();
            test();        // an instance that needs no iteration at all ends here (retired at the next trip)
        }
        if (!warp_or(has ? 1 : 0)) break;
        const bool run = has && !done;
        if (!warp_or(run ? 1 : 0)) continue;
        const int yps = job.yps;
        (void)yps;

        // ================= S1: backward sweep -- residuals, factorisation, affine right-hand side
        double last_sig = 1.0;
        double xb = 0.0, xd = 0.0;  // explicit dynamics / bound-slack residual norms of this iterate (see mpcb_qp.cuh)
        const bool issue = run && s == 0;
        {
            // terminal stage N: L_N = sqrt(Q_t), p_N = q_N
            double *wN = ws + (size_t)N * L::STAGE;
            warp_sync();
            g8_expect(pipe, L::BAT, issue);
            g8_copy(pipe, sm.rec + L::O_BAT, ws + (size_t)(N - 1) * L::STAGE + L::O_BAT, L::BAT, issue);
            for (int idx = s; idx < L::LXX; idx += kLPI) sm.lxx[idx] = 0.0;
            warp_sync();
            MPCB_UNROLL
            for (int t = 0; t < NXT; t++) {
                const int i = s + kLPI * t;
                if (i >= NX) continue;
                const double H0 = P.Qt[i];
                const double zN = run ? wN[L::O_Z + NU + i] : 0.0, piN = run ? wN[L::O_PI + i] : 0.0;
                const double q = H0 * zN + (run ? wN[L::O_G + NU + i] : 0.0) - piN;
                sm.lxx[i * NX + i] = sqrt(H0);
                sm.cPv[i] = q; sm.cPi[i] = piN; sm.cZx[i] = zN;
                if (run) wN[L::O_PV + i] = q;
            }
            warp_sync();
            if (run)
                for (int idx = s; idx < L::LXX; idx += kLPI) wN[L::O_LXX + idx] = sm.lxx[idx];
        }
        for (int k = N - 1; k >= 0; k--) {
            double *wk = ws + (size_t)k * L::STAGE;
            double zr[NT], Hd[NT], q[NT];
            if (k > 0) l2_prefetch(wk - L::STAGE + L::O_Z, L::O_C1 - L::O_Z, issue);  // next stage's vectors: HBM -> L2
            {
                // the lane's own rows of the stage vectors: all loads in flight beside the bulk copy of [B A]'
                double vtl[NT], vtu[NT], vll[NT], vlu[NT], vlb[NT], vub[NT], vg[NT], vpi[NT];
                MPCB_UNROLL
                for (int t = 0; t < NT; t++) {
                    const int row = s + kLPI * t;
                    const VarKind vk = var_kind<NX, NU>(k, row, N);
                    const bool hb = run && row < NZ && vk.hasb, vv = run && row < NZ && vk.var;
                    zr[t] = (run && row < NZ) ? wk[L::O_Z + row] : 0.0;
                    vtl[t] = vtu[t] = 1.0;
                    vll[t] = vlu[t] = vlb[t] = vub[t] = vg[t] = vpi[t] = 0.0;
                    if (hb) {
                        vtl[t] = wk[L::O_TL + row]; vtu[t] = wk[L::O_TU + row]; vll[t] = wk[L::O_LL + row]; vlu[t] = wk[L::O_LUP + row];
                        vlb[t] = wk[L::O_LB + row]; vub[t] = wk[L::O_UB + row];
                    }
                    if (vv) {
                        vg[t] = wk[L::O_G + row];
                        if (row >= NU) vpi[t] = wk[L::O_PI + row - NU];
                    }
                }
                // everything of q and Hd that does not need the matrices
                MPCB_UNROLL
                for (int t = 0; t < NT; t++) {
                    const int row = s + kLPI * t;
                    const VarKind vk = var_kind<NX, NU>(k, row, N);
                    const bool hb = run && row < NZ && vk.hasb, vv = run && row < NZ && vk.var;
                    Hd[t] = hess_diag<NX, NU, double>(P, 0, row < NZ ? row : 0);
                    q[t] = 0.0;
                    if (vv) q[t] = Hd[t] * zr[t] + vg[t] - vll[t] + vlu[t] - vpi[t];
                    if (hb) {
                        const double itl = fast_rcp(vtl[t]), itu = fast_rcp(vtu[t]);
                        const double rdl = zr[t] - vlb[t] - vtl[t], rdu = vub[t] - zr[t] - vtu[t];
                        xd = fmax(xd, fmax(fabs(rdl), fabs(rdu)));
                        Hd[t] += vll[t] * itl + vlu[t] * itu;
                        q[t] += (vll[t] + vll[t] * rdl * itl) - (vlu[t] + vlu[t] * rdu * itu);
                    }
                }
            }
            double vb[NXT], vpik[NXT];
            MPCB_UNROLL
            for (int t = 0; t < NXT; t++) {
                const int i = s + kLPI * t;
                vb[t] = (run && i < NX) ? wk[L::O_B + i] : 0.0;
                vpik[t] = (run && i < NX) ? wk[L::O_PI + i] : 0.0;
            }
            MPCB_UNROLL
            for (int t = 0; t < NT; t++)
                if (s + kLPI * t < NZ) sm.vz[s + kLPI * t] = zr[t];  // readers of the previous stage's vz passed that stage's syncs
            g8_wait(pipe, run);  // [B A]' of this stage has landed (fetched during the previous stage's Householder loop)
            int ri[NT];
            bool rvar[NT];
            MPCB_UNROLL
            for (int t = 0; t < NT; t++) {
                ri[t] = (s + kLPI * t < NZ) ? s + kLPI * t : 0;
                rvar[t] = run && s + kLPI * t < NZ && var_kind<NX, NU>(k, s + kLPI * t, N).var;
            }
            // q += [B A]' pi_{k+1}
            {
                double a[NT];
                MPCB_UNROLL
                for (int t = 0; t < NT; t++) a[t] = 0.0;
                g8_matvec_rows<NT>(sm.rec + L::O_BAT, L::LDB, sm.cPi, NX, ri, a);
                MPCB_UNROLL
                for (int t = 0; t < NT; t++)
                    if (rvar[t]) q[t] += a[t];
            }
            // r_k = b_k + [B A] z_k - dx-part of z_{k+1}
            {
                double a[NXT];
                MPCB_UNROLL
                for (int t = 0; t < NXT; t++) a[t] = vb[t] - sm.cZx[xi[t]];
                g8_matvec_cols<NXT>(sm.rec + L::O_BAT, L::LDB, sm.vz, NZ, xi, a);
                MPCB_UNROLL
                for (int t = 0; t < NXT; t++) {
                    const int i = s + kLPI * t;
                    if (i < NX) {
                        sm.sRb[i] = a[t];
                        if (run) { wk[L::O_RB + i] = a[t]; xb = fmax(xb, fabs(a[t])); }
                    }
                }
            }
            warp_sync();
            // t2 = P_{k+1} r_k + p_{k+1}
            {
                double a[NXT];
                MPCB_UNROLL
                for (int t = 0; t < NXT; t++) a[t] = 0.0;
                g8_matvec_cols<NXT>(sm.lxx, NX, sm.sRb, NX, xi, a);
                MPCB_UNROLL
                for (int t = 0; t < NXT; t++)
                    if (s + kLPI * t < NX) sm.sT1[s + kLPI * t] = a[t];
                warp_sync();
                MPCB_UNROLL
                for (int t = 0; t < NXT; t++) a[t] = sm.cPv[xi[t]];
                g8_matvec_rows<NXT>(sm.lxx, NX, sm.sT1, NX, xi, a);
                MPCB_UNROLL
                for (int t = 0; t < NXT; t++)
                    if (s + kLPI * t < NX) sm.sT2[s + kLPI * t] = a[t];
            }
            // carry this stage's pi and dx-part of z to stage k-1 (cPi / cZx were consumed before the two syncs above)
            MPCB_UNROLL
            for (int t = 0; t < NXT; t++) {
                const int i = s + kLPI * t;
                if (i < NX) { sm.cPi[i] = vpik[t]; sm.cZx[i] = sm.vz[NU + i]; }
            }
            warp_sync();
            // affine backward vectors before the substitution: l = q + [B A]' t2
            double l[NT];
            {
                MPCB_UNROLL
                for (int t = 0; t < NT; t++) l[t] = q[t];
                g8_matvec_rows<NT>(sm.rec + L::O_BAT, L::LDB, sm.sT2, NX, ri, l);
                MPCB_UNROLL
                for (int t = 0; t < NT; t++)
                    if (s + kLPI * t >= NZ) l[t] = 0.0;
            }
            // W = [B A]' Lxx_{k+1}: rows of this lane, accumulated over the rows j of L (zero upper triangle)
            double w[NT][NX];
            MPCB_UNROLL
            for (int t = 0; t < NT; t++)
                MPCB_UNROLL
                for (int c = 0; c < NX; c++) w[t][c] = 0.0;
            MPCB_NOUNROLL
            for (int j = 0; j < NX; j++) {
                double lr[NX], bj[NT];
                MPCB_UNROLL
                for (int c = 0; c < NX; c++) lr[c] = sm.lxx[j * NX + c];
                MPCB_UNROLL
                for (int t = 0; t < NT; t++) bj[t] = (s + kLPI * t < NZ) ? sm.rec[L::O_BAT + ri[t] * L::LDB + j] : 0.0;
                MPCB_UNROLL
                for (int t = 0; t < NT; t++)
                    MPCB_UNROLL
                    for (int c = 0; c < NX; c++) w[t][c] += bj[t] * lr[c];
            }
            // [B A]' of this stage is dead from here on: fetch the next stage's beside the Householder loop
            warp_sync();
            if (k > 0) {
                g8_expect(pipe, L::BAT, issue);
                g8_copy(pipe, sm.rec + L::O_BAT, wk - L::STAGE + L::O_BAT, L::BAT, issue);
            }
            // ---- Householder LQ of [diag(sqrt(Hd)) | W]; pivot j lives in slot j / 8 of sub-lane j % 8.
            // No shuffles in the loop: Hd and its square root of every row are put in shared memory once
            // per stage, and every lane computes the pivot's |w_j|^2 itself from the broadcast row.
            double dsq[NT];
            MPCB_UNROLL
            for (int t = 0; t < NT; t++) {
                dsq[t] = sqrt(Hd[t]);
            }
            MPCB_UNROLL
            for (int t = 0; t < NT; t++)
                if (s + kLPI * t < NZ) { sm.hd[s + kLPI * t] = Hd[t]; sm.ds()[s + kLPI * t] = dsq[t]; }
            const int jend = (k == 0) ? NU : NZ;
            double sig = 1.0;
            MPCB_UNROLL
            for (int tj = 0; tj < NT; tj++) {
                const int jn = (jend - kLPI * tj < kLPI) ? jend - kLPI * tj : kLPI;  // pivots of this slot (uniform)
                MPCB_NOUNROLL
                for (int jj = 0; jj < jn; jj++) {
                    const int j = kLPI * tj + jj;
                    double *vr = sm.vrow[j & 1];
                    if (s == jj) {
                        MPCB_UNROLL
                        for (int c = 0; c < NX; c++) vr[c] = w[tj][c];
                    }
                    warp_sync();
                    const double hdj = sm.hd[j], dsj = sm.ds()[j];
                    double v[NX];
                    MPCB_UNROLL
                    for (int c = 0; c < NX; c++) v[c] = vr[c];
                    double djj;
                    {
                        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
                        MPCB_UNROLL
                        for (int c = 0; c + 3 < NX; c += 4) { d0 += v[c] * v[c]; d1 += v[c + 1] * v[c + 1]; d2 += v[c + 2] * v[c + 2]; d3 += v[c + 3] * v[c + 3]; }
                        MPCB_UNROLL
                        for (int c = NX & ~3; c < NX; c++) d0 += v[c] * v[c];
                        djj = (d0 + d1) + (d2 + d3);
                    }
                    const double s2v = hdj + djj;
                    const double rs = fast_rsqrt(s2v);
                    const double idjj = fast_rcp(djj);
                    sig = s2v * rs;
                    const double kap = (djj > 0.0) ? (sig - dsj) * idjj : 0.0;
                    MPCB_UNROLL
                    for (int t = tj; t < NT; t++) {
                        const int row = s + kLPI * t;
                        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
                        MPCB_UNROLL
                        for (int c = 0; c + 3 < NX; c += 4) {
                            d0 += v[c] * w[t][c]; d1 += v[c + 1] * w[t][c + 1]; d2 += v[c + 2] * w[t][c + 2]; d3 += v[c + 3] * w[t][c + 3];
                        }
                        MPCB_UNROLL
                        for (int c = NX & ~3; c < NX; c++) d0 += v[c] * w[t][c];
                        const double dot = (d0 + d1) + (d2 + d3);
                        const double lij = (row > j && row < NZ) ? dot * rs : 0.0;
                        const double f = lij * kap;
                        MPCB_UNROLL
                        for (int c = 0; c < NX; c++) w[t][c] -= f * v[c];
                        const double val = (row == j) ? sig : lij;
                        if (row >= j && row < NZ) {
                            if (j < NU) sm.rec[L::O_LU + j * L::NZP + row] = val;
                            else sm.lxx[(row - NU) * NX + (j - NU)] = val;
                        }
                    }
                    if (s == jj && j < NU) sm.rec[L::O_INVD + j] = rs;
                }
            }
            last_sig = sig;
            warp_sync();
            // forward substitution with the u-columns of L (the pivot-row buffer is dead: its first half is the scratch)
            g8_fwd_subst<NX, NU, NT>(sm, l, s);
            MPCB_UNROLL
            for (int t = 0; t < NT; t++) {
                const int row = s + kLPI * t;
                if (row >= NZ) continue;
                if (row < NU) { if (run) wk[L::O_LVEC + row] = l[t]; }
                else { if (run) wk[L::O_PV + row - NU] = l[t]; sm.cPv[row - NU] = l[t]; }
            }
            if (run) {
                for (int idx = s; idx < NU * L::NZP; idx += kLPI) wk[L::O_LU + idx] = sm.rec[L::O_LU + idx];
                if (s < NU) wk[L::O_INVD + s] = sm.rec[L::O_INVD + s];
                if (k > 0)
                    for (int idx = s; idx < L::LXX; idx += kLPI) wk[L::O_LXX + idx] = sm.lxx[idx];
            }
        }
        pipe_fence();  // L, lvec, r_b, p written by this sweep (and the shared image of Lu) are fetched / overwritten by bulk copies below
        warp_sync();
        // a breakdown (NaN) anywhere in the recursion propagates into the last pivot of stage 0
        const bool qpfail = grp_max((last_sig == last_sig && fabs(last_sig) < HUGE_VAL) ? 0.0 : 1.0) > 0.0;
        if (run && qpfail) { status = ST_QPFAIL; done = true; }
        const bool run2 = has && !done;

        // ================= S2: forward sweep, affine step
        double a_aff, mu_aff, sigmu;
        {
            double imax, s1, s2;
            qp8_forward<NX, NU, false>(P, sm, pipe, ws, run2, 0.0, imax, s1, s2);
            a_aff = (imax > 1.0) ? 1.0 / imax : 1.0;
            mu_aff = (mu * nb + a_aff * s1 + a_aff * a_aff * s2) / nb;
            double sigma = mu_aff / mu;
            sigma = sigma * sigma * sigma;
            sigmu = sigma * mu;
        }
        // ================= S3: backward sweep for the corrector increment (delta form)
        {
            warp_sync();
            MPCB_UNROLL
            for (int t = 0; t < NXT; t++)
                if (s + kLPI * t < NX) sm.cPv[s + kLPI * t] = 0.0;
            for (int k = N - 1; k >= 0; k--) {
                double *wk = ws + (size_t)k * L::STAGE;
                double l[NT], pv_old[NT], cc[NT];
                MPCB_UNROLL
                for (int t = 0; t < NT; t++) {
                    const int row = s + kLPI * t;
                    const bool hb = run2 && row < NZ && var_kind<NX, NU>(k, row, N).hasb;
                    cc[t] = hb ? wk[L::O_C1 + row] - sigmu * wk[L::O_C2 + row] : 0.0;
                    pv_old[t] = (run2 && row >= NU && row < NZ) ? wk[L::O_PV + row - NU] : 0.0;
                }
                warp_sync();
                g8_expect(pipe, L::O_RB, run2 && s == 0);
                g8_copy(pipe, sm.rec, wk, L::O_RB, run2 && s == 0);
                if (k > 0) {  // what the next stage will fetch: HBM -> L2
                    l2_prefetch(wk - L::STAGE, L::O_RB, run2 && s == 0);
                    l2_prefetch(wk - L::STAGE + L::O_C1, 2 * L::NZP, run2 && s == 0);
                    l2_prefetch(wk - L::STAGE + L::O_PV, L::NXP, run2 && s == 0);
                }
                g8_wait(pipe, run2);
                MPCB_UNROLL
                for (int t = 0; t < NT; t++) {
                    const int row = s + kLPI * t;
                    const int rr = row < NZ ? row : 0;
                    double l0 = cc[t], l1 = 0.0;
                    MPCB_UNROLL
                    for (int c = 0; c + 1 < NX; c += 2) {
                        l0 += sm.rec[L::O_BAT + rr * L::LDB + c] * sm.cPv[c];
                        l1 += sm.rec[L::O_BAT + rr * L::LDB + c + 1] * sm.cPv[c + 1];
                    }
                    if (NX & 1) l0 += sm.rec[L::O_BAT + rr * L::LDB + NX - 1] * sm.cPv[NX - 1];
                    l[t] = (row < NZ) ? l0 + l1 : 0.0;
                }
                g8_fwd_subst<NX, NU, NT>(sm, l, s);  // its warp_sync also orders every lane's reads of cPv above before the writes below
                MPCB_UNROLL
                for (int t = 0; t < NT; t++) {
                    const int row = s + kLPI * t;
                    if (row >= NZ) continue;
                    if (row < NU) { if (run2) wk[L::O_LVEC + row] = sm.rec[L::O_LVEC + row] + l[t]; }
                    else { if (run2) wk[L::O_PV + row - NU] = pv_old[t] + l[t]; sm.cPv[row - NU] = l[t]; }
                }
            }
            pipe_fence();  // lvec written above is fetched by the bulk copies of the next sweep
            warp_sync();
        }
        // ================= S4: forward sweep, full predictor-corrector step
        double alpha;
        {
            double imax, d1, d2;
            qp8_forward<NX, NU, true>(P, sm, pipe, ws, run2, sigmu, imax, d1, d2);
            const double tau = fmax(0.995, 1.0 - mu_aff);
            alpha = (imax > tau) ? tau / imax : 1.0;
        }
        // ================= F4b: take the step (KU stages per trip: their loads are all in flight together)
        {
            constexpr int KU = 1;  // stages per trip (more in flight together was measured slower: register pressure)
            double cmax = 0.0, msum = 0.0;
            for (int k0 = 0; k0 <= N; k0 += KU) {
                double z[KU][NT], dz[KU][NT], tl[KU][NT], tu[KU][NT], ll[KU][NT], lu[KU][NT], lb[KU][NT], ub[KU][NT], dza[KU][NT];
                double pi[KU][NXT], dpi[KU][NXT];
                bool hb[KU][NT], vr_[KU][NT];
                MPCB_UNROLL
                for (int u = 0; u < KU; u++) {
                    const int k = k0 + u;
                    const double *wk = ws + (size_t)k * L::STAGE;
                    MPCB_UNROLL
                    for (int t = 0; t < NT; t++) {
                        const int row = s + kLPI * t;
                        const VarKind vk = var_kind<NX, NU>(k, row, N);
                        hb[u][t] = run2 && k <= N && row < NZ && vk.hasb;
                        vr_[u][t] = run2 && k <= N && row < NZ && vk.var;
                        z[u][t] = dz[u][t] = ll[u][t] = lu[u][t] = lb[u][t] = ub[u][t] = dza[u][t] = 0.0;
                        tl[u][t] = tu[u][t] = 1.0;
                        if (vr_[u][t]) { z[u][t] = wk[L::O_Z + row]; dz[u][t] = wk[L::O_DZ + row]; }
                        if (hb[u][t]) {
                            tl[u][t] = wk[L::O_TL + row]; tu[u][t] = wk[L::O_TU + row]; ll[u][t] = wk[L::O_LL + row]; lu[u][t] = wk[L::O_LUP + row];
                            lb[u][t] = wk[L::O_LB + row]; ub[u][t] = wk[L::O_UB + row]; dza[u][t] = wk[L::O_DZA + row];
                        }
                    }
                    MPCB_UNROLL
                    for (int t = 0; t < NXT; t++) {
                        const int i = s + kLPI * t;
                        const bool on = run2 && k >= 1 && k <= N && i < NX;
                        pi[u][t] = on ? wk[L::O_PI + i] : 0.0;
                        dpi[u][t] = on ? wk[L::O_DPI + i] : 0.0;
                    }
                }
                MPCB_UNROLL
                for (int u = 0; u < KU; u++) {
                    const int k = k0 + u;
                    double *wk = ws + (size_t)k * L::STAGE;
                    MPCB_UNROLL
                    for (int t = 0; t < NT; t++) {
                        const int row = s + kLPI * t;
                        if (hb[u][t]) {
                            double tl_ = tl[u][t], tu_ = tu[u][t], ll_ = ll[u][t], lu_ = lu[u][t];
                            const double itl = fast_rcp(tl_), itu = fast_rcp(tu_);
                            const BoxStep<double> a = box_step(z[u][t], dza[u][t], lb[u][t], ub[u][t], tl_, tu_, ll_, lu_, ll_ * tl_, lu_ * tu_, itl, itu);
                            const BoxStep<double> b = box_step(z[u][t], dz[u][t], lb[u][t], ub[u][t], tl_, tu_, ll_, lu_,
                                                               ll_ * tl_ + a.dll * a.dtl - sigmu, lu_ * tu_ + a.dlu * a.dtu - sigmu, itl, itu);
                            tl_ += alpha * b.dtl; tu_ += alpha * b.dtu; ll_ += alpha * b.dll; lu_ += alpha * b.dlu;
                            wk[L::O_TL + row] = tl_; wk[L::O_TU + row] = tu_; wk[L::O_LL + row] = ll_; wk[L::O_LUP + row] = lu_;
                            cmax = fmax(cmax, fmax(ll_ * tl_, lu_ * tu_));
                            msum += ll_ * tl_ + lu_ * tu_;
                        }
                        if (vr_[u][t]) wk[L::O_Z + row] = z[u][t] + alpha * dz[u][t];
                    }
                    MPCB_UNROLL
                    for (int t = 0; t < NXT; t++) {
                        const int i = s + kLPI * t;
                        if (run2 && k >= 1 && k <= N && i < NX) wk[L::O_PI + i] = pi[u][t] + alpha * dpi[u][t];
                    }
                }
            }
            const double c_ = grp_max(cmax), m_ = grp_sum(msum) / nb;
            if (run2) { comp = c_; mu = m_; }
        }
        xb = grp_max(xb);
        xd = grp_max(xd);
        if (run2) {
            est_g *= (1.0 - alpha);
            est_b = xb * (1.0 - alpha);  // measured on this iterate by S1, then the exact-arithmetic decay of one step
            est_d = xd * (1.0 - alpha);
            git++;
            if (!(alpha >= P.alpha_min)) {
                status = (alpha == alpha) ? ST_MINSTEP : ST_NAN;
                done = true;
            }
        }
        warp_sync();
    }
}

}  // namespace mpcb
