"""
The seam between the device path and `scipy.optimize` (mirror of the reference's
optimization/optimizers/optimizer.py:283-415, 453-515: `problem()` builds `loss_and_grad(x)` over the flat
parameter vector, `solve()` hands it to `scipy.optimize.minimize(jac=True)`), for the common case where the
parameters are the force densities: every iteration is one forward + adjoint pass on the GPU (model, fused
goal / loss kernel, state VJP, adjoint solve with the stored factor), and only the scalar loss and the
gradient `q_bar` cross to the host.
"""
import numpy as np
import torch

from .equilibrium import EquilibriumParametersState, LoadState

__all__ = ["OptProblem", "Optimizer", "LBFGSB", "SLSQP", "TrustRegionConstrained", "NewtonCG", "optimize_q"]

F64 = torch.float64


class OptProblem:
    """optimizer.py:47-85: what scipy.optimize.minimize needs."""

    def __init__(self, fun, x0, bounds, method, tol, maxiter, callback=None):
        self.fun, self.x0, self.bounds, self.method = fun, x0, bounds, method
        self.tol, self.maxiter, self.callback = tol, maxiter, callback

    constraints = None      # scipy NonlinearConstraint list (constrained.py:38-102)
    hess = None             # callable Hessian (second_order.py:20-39)

    def to_kwargs(self):
        kw = dict(fun=self.fun, jac=True, x0=self.x0, method=self.method, bounds=self.bounds, tol=self.tol,
                  callback=self.callback, options={"maxiter": self.maxiter})
        if self.constraints:
            kw["constraints"] = self.constraints
        if self.hess is not None:
            kw["hess"] = self.hess
        if self.method == "Newton-CG":
            kw.pop("bounds")
        return kw


class Optimizer:
    """optimizer.py:88-515 for parameters = q (lower/upper bounds per edge or scalars)."""
    name = None

    def __init__(self, disp=False):
        self.disp = disp

    def problem(self, model, structure, loss, q0, xyz_fixed, loads, bounds=(None, None), maxiter=100, tol=1e-6,
                callback=None, opt_xyz_fixed=None, bounds_xyz_fixed=(None, None), opt_loads=None,
                bounds_loads=(None, None), q_groups=None, constraints=None, hessian=False):
        """
        The flat parameter vector is [q, xyz_fixed[opt_xyz_fixed], loads[opt_loads]]: every force density, plus
        the support coordinates and nodal load components selected by the boolean masks (Ns,3) / (N,3) -- the
        parameter kinds of the reference (parameters/parameters.py: EdgeForceDensityParameter,
        NodeSupport{X,Y,Z}Parameter, NodeLoad{X,Y,Z}Parameter), as in its gridshell inverse-design examples.
        `q_groups`: a list of edge-index arrays, each ONE parameter shared by its edges (the reference's
        EdgeGroupForceDensityParameter, parameters/parameters.py; tests/test_arch_benchmark.py:107): the q part of the
        flat vector is then [one value per group, the ungrouped edges in edge order]; a group starts from the mean of
        its edges' q0 and its gradient is the sum of its edges' q_bar.
        `constraints`: a list of `jax_fdm_b200.constraints.Constraint` (SLSQP, trust-constr): values from one forward
        solve, Jacobian by the sparse path's forward mode -- one tangent per optimisation parameter, all solved with
        the primal factor in one batched launch (the reference: `jacfwd` on the dense model, constrained.py:38-102).
        `hessian`: also hand scipy a Hessian (second_order.py:20-39: jacfwd(jacrev(loss))).  Here it is built from the
        EXACT adjoint gradients of 2 n perturbed parameter vectors evaluated as ONE batch (central differences, step
        1e-6 relative): second derivatives of the goal kernels are not built, a batch of 2 n designs is one pass.
        """
        dev = torch.device("cuda", structure.device) if isinstance(structure.device, int) else torch.device(structure.device)
        xs0 = np.ascontiguousarray(xyz_fixed, dtype=np.float64)
        p0 = np.ascontiguousarray(loads, dtype=np.float64)
        q_full0 = np.ascontiguousarray(q0, dtype=np.float64).reshape(-1)
        x0 = q_full0
        expand = None                                      # (E,) position in the packed q part of every edge
        if q_groups:
            slot = np.full(q_full0.size, -1, dtype=np.int64)
            for gi, idx in enumerate(q_groups):
                idx = np.asarray(idx, dtype=np.int64).reshape(-1)
                assert idx.size and np.all(slot[idx] < 0), "q_groups must be non-empty and disjoint"
                slot[idx] = gi
            rest = np.flatnonzero(slot < 0)
            slot[rest] = len(q_groups) + np.arange(rest.size)
            x0 = np.concatenate([[q_full0[np.asarray(idx, dtype=np.int64)].mean() for idx in q_groups], q_full0[rest]])
            expand = torch.as_tensor(slot, device=dev)
        nq = x0.size
        mxs = None if opt_xyz_fixed is None else np.asarray(opt_xyz_fixed, dtype=bool).reshape(xs0.shape)
        mp = None if opt_loads is None else np.asarray(opt_loads, dtype=bool).reshape(p0.shape)
        nxs = 0 if mxs is None else int(mxs.sum())
        xs_t = torch.as_tensor(xs0, device=dev)
        p_t = torch.as_tensor(p0, device=dev)
        mxs_t = None if mxs is None else torch.as_tensor(mxs, device=dev)
        mp_t = None if mp is None else torch.as_tensor(mp, device=dev)

        def unpack(x):
            """x (numpy) -> q, xyz_fixed, loads as leaf tensors on the device."""
            xt = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=dev)
            q = xt[:nq].clone().requires_grad_(True)
            if expand is not None:                          # leaf = the packed vector; autograd sums a group's q_bar
                q = (q, q.index_select(0, expand))
            xs, p = xs_t, p_t
            if mxs_t is not None:
                xs = xs_t.clone()
                xs[mxs_t] = xt[nq:nq + nxs]
            if mp_t is not None:
                p = p_t.clone()
                p[mp_t] = xt[nq + nxs:]
            return q, xs.requires_grad_(mxs_t is not None), p.requires_grad_(mp_t is not None)

        def loss_and_grad(x):                            # optimizer.py:365-370: jit(value_and_grad(loss))
            q, xs, p = unpack(x)
            q_leaf, q = q if isinstance(q, tuple) else (q, q)
            val = loss(EquilibriumParametersState(q, xs, LoadState(p, 0.0, 0.0)), model, structure)
            val.backward()
            grads = [q_leaf.grad]
            if mxs_t is not None:
                grads.append(xs.grad[mxs_t])
            if mp_t is not None:
                grads.append(p.grad[mp_t])
            return float(val.item()), torch.cat([g.reshape(-1) for g in grads]).cpu().numpy()

        parts = [x0] + ([xs0[mxs]] if mxs is not None else []) + ([p0[mp]] if mp is not None else [])
        x0 = np.concatenate([a.reshape(-1) for a in parts])
        val, grad = loss_and_grad(x0)                    # optimizer.py:372-380: warm start, no NaN allowed
        assert np.isfinite(val) and np.all(np.isfinite(grad)), "the loss or its gradient is not finite at x0"
        bnds = []
        for (lo, hi), n in ((bounds, nq), (bounds_xyz_fixed, nxs), (bounds_loads, x0.size - nq - nxs)):
            lo = np.broadcast_to(-np.inf if lo is None else lo, (n,))
            hi = np.broadcast_to(np.inf if hi is None else hi, (n,))
            bnds += list(zip(lo, hi))
        prob = OptProblem(loss_and_grad, x0, bnds, self.name, tol, maxiter, callback)
        prob.unpack = lambda x: tuple((t[1] if isinstance(t, tuple) else t).detach().cpu().numpy() for t in unpack(x))

        # ---- tangents of (q, xyz_fixed, loads) for the unit vectors of the flat parameter vector (static)
        def tangent_basis():
            n = x0.size
            E_ = q_full0.size
            qd = np.zeros((n, E_))
            if expand is None:
                qd[np.arange(nq), np.arange(nq)] = 1.0
            else:
                slot_np = expand.cpu().numpy()
                qd[slot_np, np.arange(E_)] = 1.0
            xsd = np.zeros((n,) + xs0.shape)
            pd = np.zeros((n,) + p0.shape)
            if mxs is not None:
                idx = np.argwhere(mxs)
                xsd[nq + np.arange(nxs), idx[:, 0], idx[:, 1]] = 1.0
            if mp is not None:
                idx = np.argwhere(mp)
                pd[nq + nxs + np.arange(idx.shape[0]), idx[:, 0], idx[:, 1]] = 1.0
            T_ = lambda a: torch.as_tensor(a, device=dev)
            return T_(qd), (T_(xsd) if mxs is not None else None), (T_(pd) if mp is not None else None)

        if constraints:
            from scipy.optimize import NonlinearConstraint
            basis = tangent_basis()
            groups = {}
            for c in constraints:                          # one NonlinearConstraint per state field
                groups.setdefault(c.field, []).append(c)
            prob.constraints = []
            for field, cs in groups.items():
                rows = torch.as_tensor(np.concatenate([c.rows(structure) for c in cs]), device=dev)
                lb = np.concatenate([np.full(c.rows(structure).size, c.bound_low, dtype=np.float64) for c in cs])
                ub = np.concatenate([np.full(c.rows(structure).size, c.bound_up, dtype=np.float64) for c in cs])

                def values(x, field=field, rows=rows):
                    q, xs, p = unpack(x)
                    q = q[1] if isinstance(q, tuple) else q
                    with torch.no_grad():
                        st = model(EquilibriumParametersState(q.detach(), xs.detach(), LoadState(p.detach(), 0.0, 0.0)), structure)
                    return getattr(st, field).reshape(-1).index_select(0, rows).cpu().numpy()

                def jacobian(x, field=field, rows=rows):   # (constraints, parameters)
                    q, xs, p = unpack(x)
                    q = q[1] if isinstance(q, tuple) else q
                    with torch.no_grad():
                        _, sd = model.jvp(EquilibriumParametersState(q.detach(), xs.detach(), LoadState(p.detach(), 0.0, 0.0)),
                                          structure, q_dot=basis[0], xyz_fixed_dot=basis[1], loads_dot=basis[2])
                    J = getattr(sd, field).reshape(x0.size, -1).index_select(1, rows)
                    return J.t().contiguous().cpu().numpy()

                prob.constraints.append(NonlinearConstraint(fun=values, jac=jacobian, lb=lb, ub=ub))
            prob.constraint_functions = [(c.fun, c.jac) for c in prob.constraints]

        if hessian:
            def hess(x, h=1e-6):
                """Central differences of the exact gradient, all 2 n perturbed parameter vectors as one batch."""
                n = x.size
                step = h * np.maximum(1.0, np.abs(x))
                X = np.repeat(x[None, :], 2 * n, axis=0)
                X[np.arange(n), np.arange(n)] += step
                X[n + np.arange(n), np.arange(n)] -= step
                G = batched_grad(X)
                H = (G[:n] - G[n:]) / (2.0 * step[:, None])
                return 0.5 * (H + H.T)

            def batched_grad(X):
                B_ = X.shape[0]
                Xt = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64), device=dev)
                qp = Xt[:, :nq].clone().requires_grad_(True)
                q = qp if expand is None else qp.index_select(1, expand)
                xs, p = xs_t, p_t
                if mxs_t is not None:
                    xs = xs_t.expand(B_, *xs_t.shape).clone()
                    xs[:, mxs_t] = Xt[:, nq:nq + nxs]
                    xs.requires_grad_(True)
                if mp_t is not None:
                    p = p_t.expand(B_, *p_t.shape).clone()
                    p[:, mp_t] = Xt[:, nq + nxs:]
                    p.requires_grad_(True)
                val = loss(EquilibriumParametersState(q, xs, LoadState(p, 0.0, 0.0)), model, structure)
                val.sum().backward()
                grads = [qp.grad]
                if mxs_t is not None:
                    grads.append(xs.grad[:, mxs_t])
                if mp_t is not None:
                    grads.append(p.grad[:, mp_t])
                return torch.cat([g.reshape(B_, -1) for g in grads], dim=1).cpu().numpy()

            prob.hess = hess
        return prob

    def solve(self, opt_problem):
        from scipy.optimize import minimize
        res = minimize(**opt_problem.to_kwargs())
        self.result = res
        return res.x


class LBFGSB(Optimizer):
    """optimizers/gradient_based.py: L-BFGS-B."""
    name = "L-BFGS-B"


class SLSQP(Optimizer):
    """optimizers/gradient_based.py: SLSQP."""
    name = "SLSQP"


class TrustRegionConstrained(Optimizer):
    """optimizers/second_order.py: trust-constr (constraints; Hessian with problem(hessian=True))."""
    name = "trust-constr"


class NewtonCG(Optimizer):
    """optimizers/second_order.py: Newton-CG (unbounded; needs problem(hessian=True))."""
    name = "Newton-CG"


def optimize_q(model, structure, loss, q0, xyz_fixed, loads, optimizer=None, **kw):
    """constrained_fdm (fdm.py:170-330) on arrays, parameters = q: returns (q_opt, scipy result)."""
    opt = optimizer or LBFGSB()
    x = opt.solve(opt.problem(model, structure, loss, q0, xyz_fixed, loads, **kw))
    return x, opt.result
