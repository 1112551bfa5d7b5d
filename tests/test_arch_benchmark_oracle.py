"""
CPU: BASELINE config 2 on the ORACLE -- the analytical arch benchmark of the reference
(tests/test_arch_benchmark.py: n_v in {5, 50, 100, 500}, ONE grouped force density, PredictionError on
NetworkLoadPathGoal, L-BFGS-B with bounds (-1000, -1e-3), maxiter 100, tol 1e-9).  The optimised rise must meet
sqrt(3) d / 4 within 0.1 % and the load path rho d^2 / sqrt(3) within 1.1 % (n_v = 100) / 0.3 % (n_v = 500), and the
load-path error must fall monotonically with n_v (:135-166).  This pins the oracle (and the synthetic generator)
to the reference's closed-form KAT; tests/test_gpu_arch_benchmark.py runs the same on the CUDA path.
"""
from math import sqrt

import numpy as np
import pytest
from scipy.optimize import minimize

import oracle
from oracle import goals as og
from jax_fdm_b200 import datasets

SPAN, RHO = 10.0, 1.0
RISE = sqrt(3.0) * SPAN / 4.0
LOADPATH = RHO * SPAN ** 2 / sqrt(3.0)
TERMS = [("PredictionError", [("NetworkLoadPathGoal", 0, 0.0, 1.0)], 1.0)]


def oracle_arch_optimum(nv):
    prob = datasets.arch_benchmark_problem(nv)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)

    def f_and_g(theta):
        q = np.full(prob.q.size, theta[0])

        def cot(st, xf):
            # d loadpath / d state: sum_e |l_e f_e|  (goals/network/loadpath.py)
            s = np.sign(st["lengths"] * st["forces"])
            return {"lengths": s * st["forces"], "forces": s * st["lengths"]}

        out = oracle.forward_adjoint(q, prob.xyz_fixed, prob.loads, so, cot)
        return float(og.loss(TERMS, out["state"])), np.array([out["q_bar"].sum()])

    res = minimize(f_and_g, np.array([prob.q[0]]), jac=True, method="L-BFGS-B", bounds=[(-1000.0, -1e-3)], tol=1e-9,
                   options={"maxiter": 100})
    st, _ = oracle.forward(np.full(prob.q.size, res.x[0]), prob.xyz_fixed, prob.loads, so)
    return float(np.abs(st["xyz"][:, 1]).max()), float(res.fun), float(res.x[0])


@pytest.mark.parametrize("nv,rise_tol,lp_tol", [(100, 1e-3, 1.1e-2), (500, 1e-3, 3e-3)])
def test_oracle_arch_matches_the_closed_form(nv, rise_tol, lp_tol):
    rise, lp, _ = oracle_arch_optimum(nv)
    assert abs(rise - RISE) / RISE < rise_tol
    assert abs(lp - LOADPATH) / LOADPATH < lp_tol


def test_oracle_arch_error_decreases_with_discretisation():
    errs = [abs(oracle_arch_optimum(nv)[1] - LOADPATH) / LOADPATH for nv in (5, 50, 100, 500)]
    assert all(b < a for a, b in zip(errs, errs[1:])), errs
