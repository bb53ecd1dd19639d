// Host build of the warp-synchronous kernel bodies (TEST INFRASTRUCTURE ONLY).
// Compiles mpc_blaster_b200/csrc/mpcb_linearize.cuh and mpcb_qp.cuh with g++ under
// -DMPCB_HOST_EMU, where a warp is 32 fibers (warp_emu.h), and exposes a tiny C API so
// that tests/test_kernel_emulation.py can check the kernel logic against the oracle
// without a GPU.  Never part of the product library.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mpcb_common.cuh"
#include "mpcb_linearize.cuh"
#include "mpcb_qp.cuh"
#include "mpcb_qp8.cuh"
#include "mpcb_poc.cuh"

using namespace mpcb;

template <int NX, int NU, int NSLOT, bool STRICT = false>
static int rti_one(const Params &P, double *X, double *U, const double *x0, const double *yref, int yps, const double *p,
                   int p_per_stage, int *iters, double *BAt_out, double *b_out)
{
    using L = Layout<NX, NU>;
    const int N = P.N;
    std::vector<double> ws(L::instance_stride(N), 0.0);
    for (int k = 0; k < N; k++) {
        const double *pk = p + (p_per_stage ? (size_t)k * kNP : 0);
        emu::run_warp([&]() {
            linearize_warp<NX, NU, double>(P, X + (size_t)k * NX, U + (size_t)k * NU, X + (size_t)(k + 1) * NX, pk,
                                           ws.data() + (size_t)k * L::STAGE);
        });
        if (BAt_out)
            for (int r = 0; r < L::NZ; r++)
                memcpy(BAt_out + ((size_t)k * L::NZ + r) * NX, ws.data() + (size_t)k * L::STAGE + L::O_BAT + r * L::LDB, sizeof(double) * NX);
        if (b_out) memcpy(b_out + (size_t)k * NX, ws.data() + (size_t)k * L::STAGE + L::O_B, sizeof(double) * NX);
    }
    int status = -1, it = 0;
    QpSmem<NX, NU, double, NSLOT> sm;
    memset(&sm, 0, sizeof(sm));
    emu::run_warp([&]() {
        int my_it = 0;
        int st = qp_solve_warp<NX, NU, double, NSLOT, STRICT>(P, sm, ws.data(), X, U, x0, yref, yps, &my_it);
        if (emu::lane() == 0) { status = st; it = my_it; }
    });
    *iters = it;
    return status;
}

extern "C" {

size_t emu_params_size() { return sizeof(Params); }

int emu_rti_solve(const Params *P, double *X, double *U, const double *x0, const double *yref, int yps, const double *p,
                  int p_per_stage, int *iters, double *BAt_out, double *b_out)
{
    // MPCB_EMU_NSLOT=1 exercises the single-buffer (throughput) variant
    const char *ns = getenv("MPCB_EMU_NSLOT");
    const bool one = ns && ns[0] == '1';
    if (P->strict) {  // reference semantics: always the latency variant, as the host scheduler does
        if (P->variant == 17) return rti_one<17, 6, 2, true>(*P, X, U, x0, yref, yps, p, p_per_stage, iters, BAt_out, b_out);
        if (P->variant == 13) return rti_one<13, 4, 2, true>(*P, X, U, x0, yref, yps, p, p_per_stage, iters, BAt_out, b_out);
        return rti_one<12, 4, 2, true>(*P, X, U, x0, yref, yps, p, p_per_stage, iters, BAt_out, b_out);
    }
    if (P->variant == 17)
        return one ? rti_one<17, 6, 1>(*P, X, U, x0, yref, yps, p, p_per_stage, iters, BAt_out, b_out)
                   : rti_one<17, 6, 2>(*P, X, U, x0, yref, yps, p, p_per_stage, iters, BAt_out, b_out);
    if (P->variant == 13)
        return one ? rti_one<13, 4, 1>(*P, X, U, x0, yref, yps, p, p_per_stage, iters, BAt_out, b_out)
                   : rti_one<13, 4, 2>(*P, X, U, x0, yref, yps, p, p_per_stage, iters, BAt_out, b_out);
    return one ? rti_one<12, 4, 1>(*P, X, U, x0, yref, yps, p, p_per_stage, iters, BAt_out, b_out)
               : rti_one<12, 4, 2>(*P, X, U, x0, yref, yps, p, p_per_stage, iters, BAt_out, b_out);
}

void emu_plant_step(const Params *P, const double *x, const double *u, const double *p, double *xn)
{
    if (P->variant == 17) plant_step_thread<17, 6, double>(*P, x, u, p, xn);
    else if (P->variant == 13) plant_step_thread<13, 4, double>(*P, x, u, p, xn);
    else plant_step_thread<12, 4, double>(*P, x, u, p, xn);
}
}

// rotation / quaternion helpers of mpcb_model.cuh (restating reference utils/MathUtils.py)
extern "C" {
void emu_quat_mul(const double *a, const double *b, double *c) { quat_mul<double>(a, b, c); }
void emu_quat_inv(const double *q, double *r) { quat_inv_unit<double>(q, r); }
void emu_quat_to_rot(const double *q, double *R) { quat_to_rot<double>(q, R); }
void emu_euler_to_quat(double phi, double th, double psi, double *q) { euler_to_quat<double>(phi, th, psi, q); }
}

// mpcb_poc.cuh is plain per-thread code: run it directly.  out = [poc(3), J(3x8 row-major), t_flight, status]
extern "C" void emu_poc(const double *e, const double *m, const double *pos, double V, double drag, int mode, double *out, double *p25)
{
    PocOut o;
    if (mode == POC_MODE_ANALYTIC) poc_analytic(e, m, pos, V, drag, o);
    else poc_reference(e, m, pos, V, drag, o);
    for (int i = 0; i < 3; i++) out[i] = o.poc[i];
    for (int i = 0; i < 3; i++)
        for (int c = 0; c < 8; c++) out[3 + i * 8 + c] = o.J[i][c];
    out[27] = o.t_flight;
    out[28] = o.status;
    if (p25) poc_pack_params(o, 21.582, p25);
}

// Four instances per warp (mpcb_qp8.cuh).  X[nb][(N+1)*NX], U[nb][N*NU], x0[nb][NX], yref[nb][NY] (shared over stages), p[25];
// fewer than 4 instances: the other groups stay idle; more than 4: groups are refilled as their instances end.
template <int NX, int NU>
static void rti_four(const Params &P, int nb, double *X, double *U, const double *x0, const double *yref, const double *p, int *status,
                     int *iters)
{
    using L = Layout<NX, NU>;
    const int N = P.N;
    const size_t stride = L::instance_stride(N);
    std::vector<double> ws(stride * nb, 0.0);
    for (int b = 0; b < nb; b++)
        for (int k = 0; k < N; k++) {
            double *Xb = X + (size_t)b * (N + 1) * NX, *Ub = U + (size_t)b * N * NU;
            emu::run_warp([&]() {
                linearize_warp<NX, NU, double>(P, Xb + (size_t)k * NX, Ub + (size_t)k * NU, Xb + (size_t)(k + 1) * NX, p,
                                               ws.data() + b * stride + (size_t)k * L::STAGE);
            });
        }
    static Qp8Smem<NX, NU> sm;
    memset(&sm, 0, sizeof(sm));
    // the warp's four groups draw the nb instances from a work counter (nb > 4: groups are refilled)
    unsigned next = 0;
    Qp8Batch job;
    job.X = X; job.U = U; job.x0 = x0; job.yref = yref; job.yref_stride = NX + NU; job.yps = 0;
    job.ws = ws.data(); job.u0 = nullptr; job.status = status; job.iters = iters; job.inst0 = 0; job.B = nb; job.next = &next;
    emu::run_warp([&]() { qp8_solve_queue<NX, NU>(P, sm, job); });
}
extern "C" void emu_rti_solve4(const Params *P, int nb, double *X, double *U, const double *x0, const double *yref, const double *p,
                               int *status, int *iters)
{
    if (P->variant == 17) rti_four<17, 6>(*P, nb, X, U, x0, yref, p, status, iters);
    else if (P->variant == 13) rti_four<13, 4>(*P, nb, X, U, x0, yref, p, status, iters);
    else rti_four<12, 4>(*P, nb, X, U, x0, yref, p, status, iters);
}
