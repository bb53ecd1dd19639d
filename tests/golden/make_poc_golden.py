"""Golden fixtures for the jet point-of-contact (POC) Jacobian generator, produced by RUNNING
THE REFERENCE'S OWN PYTHON (run in the build container, where /root/reference exists).

/root/reference/src/scripts/Jacobian_POC_Solver.py and htm.py are imported unchanged.  What they
need and this image lacks is replaced by stubs: the sympy-backed ``casadi`` of make_golden.py,
an empty ``matplotlib.pyplot``, and an ``AcadosSimSolver`` that integrates the model the
reference hands it (``sim.model.f_expl_expr``) with the scheme the reference asks for --
explicit Runge-Kutta, 4 stages, ``num_steps`` = 10 uniform steps over T
(Jacobian_POC_Solver.py:92-96).  The root finding, the finite differences and the homogeneous
transforms are the reference's code, quirks included:

  * the reference's call pattern (``initialise()``, :53-57) passes Python *lists*.  ``position +=
    eps`` (:284) on a list and an ndarray is resolved by numpy's ``__radd__`` (a new ndarray, not an
    in-place extend), so each coordinate is perturbed on its own and ``J_pos`` is a proper forward
    difference -- these are the golden values (``J_pos``);
  * called with *ndarrays* instead, the same statement perturbs the caller's array in place and
    ``position = self._positions`` (:296) re-binds the same object, so the perturbations accumulate:
    column i of ``J_pos`` is then the difference for a shift of all coordinates 0..i.  Recorded as
    ``J_pos_ndarray_call`` for the first pose only, to document the behaviour; not a parity target.

    python tests/golden/make_poc_golden.py   ->  tests/golden/poc_golden.npz
"""
import importlib
import os
import sys
import types

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (the casadi / acados_template stubs)

REF = "/root/reference"


def install():
    at = mg.install_stubs()
    # casadi's mtimes accepts a scalar factor (the reference writes `- self._M_c @ self._v` with M_c = 1)
    def rmatmul(self, o):
        o = mg._wrap(o)
        return self._bin(o, lambda x, y: y * x) if o.m.shape == (1, 1) else mg.SX(o.m * self.m)
    mg.SX.__rmatmul__ = rmatmul
    _init = mg.SX.__init__

    def init(self, *a):
        if len(a) == 1 and isinstance(a[0], (list, tuple)) and len(a[0]) == 0:
            self.m = sp.zeros(0, 1)
        else:
            _init(self, *a)
    mg.SX.__init__ = init

    class AcadosSimSolver:
        """ERK integrator of sim.model.f_expl_expr: num_stages = 4 (classic RK4), num_steps uniform steps over T."""

        def __init__(self, sim, json_file=None):
            o = sim.solver_options
            assert o.integrator_type == "ERK" and o.num_stages == 4
            self.steps, self.T = int(o.num_steps), float(o.T)
            xs = list(sim.model.x.m)
            self.f = sp.lambdify(xs, sim.model.f_expl_expr.m, "numpy")
            self.x = np.zeros(len(xs))      # sim_in.x: stays as set
            self.xn = np.zeros(len(xs))     # sim_out.xn: what get('x') returns

        def set(self, field, v):
            if field == "T":
                self.T = float(v)
            elif field == "x":
                self.x = np.array(v, dtype=np.float64).reshape(-1)
            else:
                raise KeyError(field)

        def solve(self):
            h = self.T / self.steps
            f = lambda x: np.asarray(self.f(*x), dtype=np.float64).reshape(-1)
            x = self.x
            for _ in range(self.steps):
                k1 = f(x); k2 = f(x + 0.5 * h * k1); k3 = f(x + 0.5 * h * k2); k4 = f(x + h * k3)
                x = x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
            self.xn = x
            return 0

        def get(self, field):
            assert field == "x"
            return self.xn.copy()

    at.AcadosSimSolver = AcadosSimSolver

    class AcadosSim(mg.Bag):
        def __init__(self):
            super().__init__(model=None, solver_options=mg.Bag())

    at.AcadosSim = AcadosSim
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def main():
    install()
    sys.path.insert(0, os.path.join(REF, "src", "scripts"))
    htm = importlib.import_module("htm")
    mod = importlib.import_module("Jacobian_POC_Solver")   # the reference's file, unmodified

    # 1. the committed entry point (simulation_blaster.py:37-39): hover pose, list arguments
    s = mod.Jacobian_POC_Solver(150, 1, 0.000015)
    s.initialise()
    init_J = [j.copy() for j in s.getJacobians()]

    # 2. homogeneous transforms (htm.py:7-36)
    rng = np.random.default_rng(20261019)
    M = 12
    eul = np.concatenate([rng.uniform(-0.17, 0.17, (M, 2)), rng.uniform(-0.35, 0.35, (M, 1))], axis=1)
    mot = np.stack([rng.uniform(-0.17, 1.2, M), rng.uniform(-0.5, 0.5, M)], axis=1)
    pos = np.concatenate([rng.uniform(-1.5, 1.5, (M, 2)), rng.uniform(1.0, 5.0, (M, 1))], axis=1)
    # the reference's own example pose (Jacobian_POC_Solver.py:308) and the hover pose of initialise()
    eul[0], mot[0], pos[0] = [0, -0.05, 0], [0.2117, 0], [0.6, 0, 3.5]
    eul[1], mot[1], pos[1] = [0, 0, 0], [0, 0], [0, 0, 4]
    T_bs = np.stack([htm.compute_T_b_s2(*mot[i]) for i in range(M)])
    T_wb = np.stack([htm.compute_T_w_b(*eul[i], pos[i]) for i in range(M)])

    # 3. solveJacobians, called the way the reference calls it (lists)
    poc, Jm, Je, Jp, x_init, Tf = [], [], [], [], [], []
    for i in range(M):
        s = mod.Jacobian_POC_Solver(150, 1, 0.000015)
        s._createIntegrator()
        s.solveJacobians(list(eul[i]), list(mot[i]), list(pos[i]))
        jm, je, jp = s.getJacobians()
        poc.append(s._POC.copy()); Jm.append(jm.copy()); Je.append(je.copy()); Jp.append(jp.copy())
        x_init.append(s.setInitConditions_Plus(eul[i], mot[i], pos[i]))
        Tf.append(s._solveRootFindingProblem(0.1, s._function, x_init[-1]))
    s = mod.Jacobian_POC_Solver(150, 1, 0.000015)
    s._createIntegrator()
    s.solveJacobians(eul[0].copy(), mot[0].copy(), pos[0].copy())
    Jp_nd = s.getJacobians()[2].copy()
    np.savez(os.path.join(HERE, "poc_golden.npz"), euler=eul, motor=mot, position=pos, T_b_s2=T_bs, T_w_b=T_wb,
             x_init=np.array(x_init), t_flight=np.array(Tf), poc=np.array(poc), J_mot=np.array(Jm), J_eul=np.array(Je),
             J_pos=np.array(Jp), J_pos_ndarray_call=Jp_nd, init_J_mot=init_J[0], init_J_eul=init_J[1], init_J_pos=init_J[2],
             stream_velocity=150.0, M_c=1.0)
    print("example pose: POC", poc[0], "t_flight", Tf[0])
    print("J_mot\n", Jm[0], "\nJ_eul\n", Je[0], "\nJ_pos\n", Jp[0], "\nJ_pos when called with ndarrays (accumulated)\n", Jp_nd)


if __name__ == "__main__":
    main()
