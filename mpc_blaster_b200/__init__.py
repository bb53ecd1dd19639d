"""B200-native batched MPC solver for the BLASTER quadrotor controller (sml93/mpc_blaster).

Hot path: rollout + sensitivities -> Gauss-Newton QP -> Riccati interior point -> RTI step,
hand-written CUDA for sm_100a behind the C ABI in include/mpcb.h.  No CPU fallback.
"""
from . import scenarios  # noqa: F401

__all__ = ["BlasterMPC", "blasterModel", "JacobianPOCSolver", "scenarios"]


def __getattr__(name):
    # torch / the CUDA library are only touched when the solver classes are asked for
    if name == "BlasterMPC":
        from .solver import BlasterMPC
        return BlasterMPC
    if name == "JacobianPOCSolver":
        from .poc import JacobianPOCSolver
        return JacobianPOCSolver
    if name == "blasterModel":
        from .acados_shim import blasterModel
        return blasterModel
    raise AttributeError(name)
