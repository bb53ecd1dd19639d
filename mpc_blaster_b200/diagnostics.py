"""Diagnostics of a finished solve, computed on the device from what ``mpcb_debug_qp`` exports.

``explicit_kkt_residuals`` evaluates the four KKT residuals of the interior-point iterate the last ``solve`` ended
with from the QP data itself (nothing from the solver's own bookkeeping): bench.py reports their maxima beside the
throughput, and tests/ compare them with the NumPy evaluation of the same export.  torch is used for the array
arithmetic only; the solve is libmpcb.so's.
"""
from __future__ import annotations

import torch


def explicit_kkt_residuals(mpc, B: int | None = None) -> dict:
    """Per-instance inf-norms {stat, eq, ineq, comp} [B] of the last solve of ``mpc`` (a BlasterMPC), instances [0,B).

    stat: H z + g - lam_l + lam_u + [B A]' pi_{k+1} - pi_k over the optimisation variables (dx_0 is pinned);
    eq: b_k + [B A] z_k - dx_{k+1};  ineq: z - lb - t_l and ub - z - t_u;  comp: lam * t."""
    B = mpc.batch if B is None else B
    d = mpc.debug_qp(B)
    N, nx, nu = mpc.N, mpc.nx, mpc.nu
    nz = nx + nu
    c = mpc.cfg
    dev = d["z"].device
    R = torch.tensor([c.R[i] for i in range(nu)], dtype=torch.float64, device=dev) * c.dt
    Q = torch.tensor([c.Q[i] for i in range(nx)], dtype=torch.float64, device=dev) * c.dt
    Qt = torch.tensor([c.Qt[i] for i in range(nx)], dtype=torch.float64, device=dev)
    H = torch.cat([torch.cat([R, Q]).repeat(N, 1), torch.cat([torch.ones(nu, dtype=torch.float64, device=dev), Qt])[None]], 0)
    k = torch.arange(N + 1, device=dev)[:, None]
    j = torch.arange(nz, device=dev)[None, :]
    var = torch.where(j < nu, k < N, k >= 1)
    hasb = torch.where(j < nu, k < N, (k >= 1) & (k < N))
    z, pi = d["z"], d["pi"]
    r = H * z + d["g"] - d["ll"] + d["lu"]
    r[:, :N] += torch.einsum("bkjc,bkc->bkj", d["BAt"], pi[:, 1:])
    r[:, :, nu:] -= pi
    r = torch.where(var, r, torch.zeros_like(r))
    eq = d["b"] + torch.einsum("bkjc,bkj->bkc", d["BAt"], z[:, :N]) - z[:, 1:, nu:]
    zero = torch.zeros_like(z)
    lb, ub = torch.where(hasb, d["lb"], zero), torch.where(hasb, d["ub"], zero)
    rd = torch.where(hasb, torch.maximum((z - lb - d["tl"]).abs(), (ub - z - d["tu"]).abs()), zero)
    comp = torch.where(hasb, torch.maximum(d["ll"] * d["tl"], d["lu"] * d["tu"]), zero)
    flat = lambda a: a.reshape(B, -1).max(dim=1).values
    return {"stat": flat(r.abs()), "eq": flat(eq.abs()), "ineq": flat(rd), "comp": flat(comp)}
