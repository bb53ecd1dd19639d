"""Generate golden fixtures by IMPORTING THE REFERENCE'S OWN PYTHON (run in the build
container, where /root/reference exists; the fixtures it writes are committed).

The reference's model file needs casadi and acados_template, neither of which is
installable here.  This script installs two *stub* modules in sys.modules -- a
sympy-backed ``casadi`` that implements exactly the SX operations blastermodel.py and
utils/MathUtils.py use, and an ``acados_template`` whose classes are attribute bags -- and
then imports /root/reference/src/scripts/blastermodel.py unchanged, runs
``generateModel()`` / ``generateController()`` with the constants of
simulation_blaster.py:12-30, and records:

  dynamics_golden.npz   f(x,u,p), df/dx, df/du of the reference's CasADi expression
                        (sympy.diff of ``model.f_expl_expr``) at seeded random points
  mathutils_golden.npz  quatMultiplication / unitQuatInversion / quat2Rot values
  ocp_golden.json       what generateController() put into the AcadosOcp (W, W_e, Vx, Vu,
                        bounds, idx, options), plus the same fields extracted from the
                        committed acados dump src/scripts/acados_ocp_blasterModel.json

    python tests/golden/make_golden.py
"""
import importlib
import json
import os
import sys
import types

import numpy as np
import sympy as sp

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------ sympy-backed casadi stub
class SX:
    __array_ufunc__ = None  # make numpy defer to our reflected operators
    _uid = 0

    def __init__(self, *a):
        if len(a) == 2 and all(isinstance(v, int) for v in a):
            self.m = sp.zeros(a[0], a[1])
        elif len(a) == 1 and isinstance(a[0], SX):
            self.m = a[0].m.copy()
        elif len(a) == 1 and isinstance(a[0], sp.MatrixBase):
            self.m = sp.Matrix(a[0])
        elif len(a) == 1 and isinstance(a[0], (list, tuple, np.ndarray)):
            arr = np.asarray(a[0], dtype=object)
            if arr.ndim == 1:
                arr = arr.reshape(-1, 1)
            self.m = sp.Matrix(arr.shape[0], arr.shape[1], lambda i, j: sp.nsimplify(arr[i, j]) if False else sp.Float(arr[i, j]) if isinstance(arr[i, j], float) else sp.sympify(arr[i, j]))
        elif len(a) == 1:
            self.m = sp.Matrix([[sp.sympify(a[0])]])
        else:
            raise TypeError(a)

    @staticmethod
    def sym(name, r=1, c=1):
        SX._uid += 1
        u = SX._uid
        return SX(sp.Matrix(r, c, lambda i, j: sp.Symbol(f"{name}__{u}_{i}_{j}", real=True)))

    @staticmethod
    def eye(n):
        return SX(sp.eye(n))

    @staticmethod
    def zeros(r, c=1):
        return SX(sp.zeros(r, c))

    def rows(self):
        return self.m.rows

    def columns(self):
        return self.m.cols

    def size(self):
        return (self.m.rows, self.m.cols)

    @property
    def shape(self):
        return self.m.shape

    def _idx(self, k):
        if isinstance(k, tuple):
            return k
        return (k % self.m.rows, k // self.m.rows)  # column-major linear index

    def __getitem__(self, k):
        return SX(self.m[self._idx(k)])

    def __setitem__(self, k, v):
        self.m[self._idx(k)] = _scalar(v)

    # -- arithmetic (elementwise with 1x1 broadcasting, like casadi)
    def _bin(self, o, f):
        o = _wrap(o)
        a, b = self.m, o.m
        if a.shape == b.shape:
            return SX(sp.Matrix(a.rows, a.cols, lambda i, j: f(a[i, j], b[i, j])))
        if a.shape == (1, 1):
            return SX(sp.Matrix(b.rows, b.cols, lambda i, j: f(a[0, 0], b[i, j])))
        if b.shape == (1, 1):
            return SX(sp.Matrix(a.rows, a.cols, lambda i, j: f(a[i, j], b[0, 0])))
        raise ValueError(f"shape mismatch {a.shape} {b.shape}")

    def __add__(self, o): return self._bin(o, lambda x, y: x + y)
    def __radd__(self, o): return _wrap(o)._bin(self, lambda x, y: x + y)
    def __sub__(self, o): return self._bin(o, lambda x, y: x - y)
    def __rsub__(self, o): return _wrap(o)._bin(self, lambda x, y: x - y)
    def __mul__(self, o): return self._bin(o, lambda x, y: x * y)
    def __rmul__(self, o): return _wrap(o)._bin(self, lambda x, y: x * y)
    def __truediv__(self, o): return self._bin(o, lambda x, y: x / y)
    def __rtruediv__(self, o): return _wrap(o)._bin(self, lambda x, y: x / y)
    def __pow__(self, e): return SX(self.m.applyfunc(lambda x: x ** e))
    def __neg__(self): return SX(-self.m)
    def __matmul__(self, o): return SX(self.m * _wrap(o).m)
    def __rmatmul__(self, o): return SX(_wrap(o).m * self.m)


def _wrap(o):
    if isinstance(o, SX):
        return o
    if isinstance(o, np.ndarray):
        arr = o if o.ndim == 2 else o.reshape(-1, 1)
        return SX(sp.Matrix(arr.shape[0], arr.shape[1], lambda i, j: sp.Float(float(arr[i, j]), 17)))
    if isinstance(o, float):
        return SX(sp.Matrix([[sp.Float(o, 17)]]))
    return SX(o)


def _scalar(v):
    if isinstance(v, SX):
        assert v.m.shape == (1, 1)
        return v.m[0, 0]
    return sp.Float(v, 17) if isinstance(v, float) else sp.sympify(v)


def vertcat(*args):
    rows = []
    for a in args:
        a = _wrap(a)
        assert a.m.cols == 1
        rows += list(a.m)
    return SX(sp.Matrix(rows))


def cos(a): return SX(_wrap(a).m.applyfunc(sp.cos))
def sin(a): return SX(_wrap(a).m.applyfunc(sp.sin))
def inv(a): return SX(_wrap(a).m.inv())
def cross(a, b): return SX(_wrap(a).m.cross(_wrap(b).m))


def reshape(a, r, c):
    flat = [a.m[i, j] for j in range(a.m.cols) for i in range(a.m.rows)]  # column-major, like casadi
    return SX(sp.Matrix(r, c, lambda i, j: flat[j * r + i]))


class Bag:
    """attribute bag standing in for AcadosModel / AcadosOcp and its members"""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def install_stubs():
    cas = types.ModuleType("casadi")
    for k, v in dict(SX=SX, vertcat=vertcat, cos=cos, sin=sin, inv=inv, cross=cross, reshape=reshape).items():
        setattr(cas, k, v)
    cas.__all__ = ["SX", "vertcat", "cos", "sin", "inv", "cross", "reshape"]
    sys.modules["casadi"] = cas
    at = types.ModuleType("acados_template")

    class AcadosModel(Bag):
        pass

    class AcadosSim(Bag):
        pass

    class AcadosOcp(Bag):
        def __init__(self):
            super().__init__(model=None, dims=Bag(), cost=Bag(), constraints=Bag(), solver_options=Bag(), parameter_values=None)

    class AcadosOcpSolver:
        last = None

        def __init__(self, ocp, json_file=None):
            self.ocp, self.json_file = ocp, json_file
            AcadosOcpSolver.last = self

    class AcadosSimSolver:
        def __init__(self, ocp, json_file=None):
            self.ocp = ocp

    for k, v in dict(AcadosModel=AcadosModel, AcadosSim=AcadosSim, AcadosOcp=AcadosOcp, AcadosOcpSolver=AcadosOcpSolver,
                     AcadosSimSolver=AcadosSimSolver).items():
        setattr(at, k, v)
    sys.modules["acados_template"] = at
    return at


def main():
    at = install_stubs()
    sys.path.insert(0, REF)                                  # for `from utils import MathUtils`
    sys.path.insert(0, os.path.join(REF, "src", "scripts"))
    blastermodel = importlib.import_module("blastermodel")   # the reference's file, unmodified
    MathUtils = importlib.import_module("utils.MathUtils")

    # constants of reference simulation_blaster.py:12-30
    mass = 9.0
    J = np.eye(3); J[0, 0] = 0.50781; J[1, 1] = 0.47314; J[2, 2] = 0.72975
    l_x, l_y, N, Tf, yaw = 0.3434, 0.3475, 60, 2.0, 0.03
    Q = np.zeros((17, 17))
    np.fill_diagonal(Q, [1e3, 1e3, 1e3, 1e3, 1e3, 1e3, 0.5e1, 0.5e1, 0.5e1, 1e1, 1e1, 1e1, 1e-2, 1e-2, 1e3, 1e3, 1e3])
    Q_t = 10 * Q
    R = np.zeros((6, 6))
    np.fill_diagonal(R, [5e-2, 5e-2, 5e-2, 5e-2, 1e-5, 1e-5])
    sB = np.array([[-1.5, -1.5, 0, -0.174532925, -0.174532925, -0.349066, -1.0, -1.0, -1.0, -0.0872665, -0.0872665, -0.0872665, -0.174532925, -0.523599, -1.5, -1.5, -2.5],
                   [1.5, 1.5, 5.0, 0.174532925, 0.174532925, 0.349066, 1.0, 1.0, 1.0, 0.0872665, 0.0872665, 0.0872665, 1.22173, 0.523599, 1.5, 1.5, 2.5]])
    cB = np.array([[0, 0, 0, 0, -0.0872665, -0.0872665], [65, 65, 65, 65, 0.0872665, 0.0872665]])
    b = blastermodel.blasterModel(mass, J, l_x, l_y, N, Tf, yaw, Q, R, Q_t, 2.2 * 9.81, sB, cB)
    assert b.generateModel() == 0
    b.generateController()
    ocp = at.AcadosOcpSolver.last.ocp
    model = ocp.model
    xs, us, ps = list(model.x.m), list(model.u.m), list(model.p.m)
    assert (len(xs), len(us), len(ps)) == (17, 6, 25)
    f = model.f_expl_expr.m
    fx = f.jacobian(xs)
    fu = f.jacobian(us)
    allsyms = xs + us + ps
    F = sp.lambdify(allsyms, f, "numpy", cse=True)
    FX = sp.lambdify(allsyms, fx, "numpy", cse=True)
    FU = sp.lambdify(allsyms, fu, "numpy", cse=True)

    rng = np.random.default_rng(20261018)
    M = 24
    X = np.zeros((M, 17)); U = np.zeros((M, 6)); Pm = np.zeros((M, 25))
    X[:, 0:3] = rng.uniform(-1.5, 1.5, (M, 3)); X[:, 3:6] = rng.uniform(-0.35, 0.35, (M, 3))
    X[:, 6:9] = rng.uniform(-1, 1, (M, 3)); X[:, 9:12] = rng.uniform(-0.3, 0.3, (M, 3))
    X[:, 12] = rng.uniform(-0.2, 1.2, M); X[:, 13] = rng.uniform(-0.5, 0.5, M); X[:, 14:17] = rng.uniform(-1, 1, (M, 3))
    U[:, 0:4] = rng.uniform(0, 65, (M, 4)); U[:, 4:6] = rng.uniform(-0.09, 0.09, (M, 2))
    Pm[:, :24] = rng.standard_normal((M, 24)); Pm[:, 24] = rng.uniform(10, 30, M)
    X[0] = 0; U[0] = 0; Pm[0] = 0; Pm[0, 24] = 2.2 * 9.81   # the reference's own start point, default params
    Fv = np.stack([np.asarray(F(*X[i], *U[i], *Pm[i]), dtype=np.float64).reshape(17) for i in range(M)])
    FXv = np.stack([np.asarray(FX(*X[i], *U[i], *Pm[i]), dtype=np.float64).reshape(17, 17) for i in range(M)])
    FUv = np.stack([np.asarray(FU(*X[i], *U[i], *Pm[i]), dtype=np.float64).reshape(17, 6) for i in range(M)])
    np.savez(os.path.join(HERE, "dynamics_golden.npz"), x=X, u=U, p=Pm, f=Fv, fx=FXv, fu=FUv,
             fx_pattern=np.array([[0 if fx[i, j] == 0 else 1 for j in range(17)] for i in range(17)], dtype=np.int8),
             fu_pattern=np.array([[0 if fu[i, j] == 0 else 1 for j in range(6)] for i in range(17)], dtype=np.int8))

    # MathUtils (dead code in the reference, but named by north_star)
    q1 = rng.standard_normal((8, 4)); q1 /= np.linalg.norm(q1, axis=1, keepdims=True)
    q2 = rng.standard_normal((8, 4)); q2 /= np.linalg.norm(q2, axis=1, keepdims=True)
    num = lambda s: np.array(s.m.evalf(17), dtype=np.float64)
    prod = np.stack([num(MathUtils.quatMultiplication(SX(list(a)), SX(list(c)))).reshape(4) for a, c in zip(q1, q2)])
    invq = np.stack([num(MathUtils.unitQuatInversion(SX(list(a)))).reshape(4) for a in q1])
    rot = np.stack([num(MathUtils.quat2Rot(SX(list(a)))).reshape(3, 3) for a in q1])
    np.savez(os.path.join(HERE, "mathutils_golden.npz"), q1=q1, q2=q2, prod=prod, inv=invq, rot=rot)

    # OCP data as generateController() set it, next to the committed acados dump
    so = ocp.solver_options
    got = dict(N=int(ocp.dims.N), W=np.asarray(ocp.cost.W).tolist(), W_e=np.asarray(ocp.cost.W_e).tolist(),
               Vx=np.asarray(ocp.cost.Vx).tolist(), Vu=np.asarray(ocp.cost.Vu).tolist(), Vx_e=np.asarray(ocp.cost.Vx_e).tolist(),
               cost_type=ocp.cost.cost_type, cost_type_e=ocp.cost.cost_type_e,
               idxbu=np.asarray(ocp.constraints.idxbu).tolist(), idxbx=np.asarray(ocp.constraints.idxbx).tolist(),
               lbu=np.asarray(ocp.constraints.lbu).tolist(), ubu=np.asarray(ocp.constraints.ubu).tolist(),
               lbx=np.asarray(ocp.constraints.lbx).tolist(), ubx=np.asarray(ocp.constraints.ubx).tolist(),
               parameter_values=np.asarray(ocp.parameter_values).tolist(),
               solver_options={k: getattr(so, k) for k in ("levenberg_marquardt", "qp_solver", "hessian_approx", "integrator_type",
                                                           "nlp_solver_type", "qp_solver_iter_max", "qp_solver_cond_N", "tf")})
    dump = json.load(open(os.path.join(REF, "src", "scripts", "acados_ocp_blasterModel.json")))
    keep_opts = ("globalization", "nlp_solver_step_length", "qp_solver_warm_start", "sim_method_num_stages", "sim_method_num_steps",
                 "qp_solver", "qp_solver_cond_N", "qp_solver_iter_max", "nlp_solver_type", "integrator_type", "hessian_approx",
                 "levenberg_marquardt", "tf", "qp_solver_tol_stat", "qp_solver_tol_eq", "qp_solver_tol_ineq", "qp_solver_tol_comp")
    ext = dict(dims={k: dump["dims"][k] for k in ("N", "nx", "nu", "np", "ny", "ny_e", "nbx", "nbu", "nbx_0", "nbx_e", "nbxe_0")},
               W_diag=np.diag(np.array(dump["cost"]["W"])).tolist(), W_offdiag_nnz=int(np.count_nonzero(np.array(dump["cost"]["W"]) - np.diag(np.diag(np.array(dump["cost"]["W"]))))),
               W_e_diag=np.diag(np.array(dump["cost"]["W_e"])).tolist(),
               lbx=dump["constraints"]["lbx"], ubx=dump["constraints"]["ubx"], lbu=dump["constraints"]["lbu"], ubu=dump["constraints"]["ubu"],
               idxbx=dump["constraints"]["idxbx"], idxbu=dump["constraints"]["idxbu"], idxbxe_0=dump["constraints"]["idxbxe_0"],
               parameter_values=dump["parameter_values"],
               time_steps_unique=sorted(set(np.round(dump["solver_options"]["time_steps"], 15).tolist())),
               solver_options={k: dump["solver_options"].get(k) for k in keep_opts})
    json.dump(dict(generateController=got, acados_dump_extract=ext), open(os.path.join(HERE, "ocp_golden.json"), "w"), indent=1)
    # a subset of the dump in its own (acados) format, for the JSON-loader tests
    sub = dict(dims={k: dump["dims"][k] for k in ("N", "nx", "nu", "np", "ny", "ny_e", "nbx", "nbu")},
               cost={k: dump["cost"][k] for k in ("W", "W_e", "cost_type", "cost_type_e")},
               constraints={k: dump["constraints"][k] for k in ("lbx", "ubx", "lbu", "ubu", "idxbx", "idxbu")},
               parameter_values=dump["parameter_values"],
               solver_options={k: dump["solver_options"][k] for k in ("tf", "nlp_solver_type", "integrator_type", "hessian_approx",
                                                                      "qp_solver", "qp_solver_cond_N", "qp_solver_iter_max",
                                                                      "globalization", "nlp_solver_step_length", "qp_solver_warm_start")})
    json.dump(sub, open(os.path.join(HERE, "acados_ocp_subset.json"), "w"))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
