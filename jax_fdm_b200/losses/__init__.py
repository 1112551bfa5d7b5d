"""
Losses (mirror of `jax_fdm.losses`): error terms over goals, regularizers, and `Loss`.

`Loss.__call__(params, model, structure)` follows losses/loss.py:60-90: the model's equilibrium state,
plus every error term, plus every regularizer.  All error terms of a loss are compiled once per
structure into ONE goal program and evaluated by ONE kernel launch (`fdm_goals_loss`) that also writes
the cotangents of the state, so the reverse pass continues straight into `fdm_state_vjp`.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib as L
from ..goals import Goal

__all__ = ["Error", "SquaredError", "MeanSquaredError", "RootMeanSquaredError", "AbsoluteError", "MeanAbsoluteError",
           "PredictionError", "MeanPredictionError", "LogMaxError", "Regularizer", "L2Regularizer", "Loss"]

F64 = torch.float64


def _over_features(e, gs):
    """Sum over the feature axis of a vector goal (xyz, residual vector); scalar goals have none."""
    return e.sum(dim=-1) if gs.goal.dim() else e


class Error:
    """losses/errors.py:34-230: alpha * errors(sum over the goals of error(goal state))."""
    err_kind = -1
    mean = False
    root = False

    def __init__(self, goals, alpha=1.0, name=None):
        self.goals = list(goals)
        for g in self.goals:
            if not isinstance(g, Goal) or g.kind < 0:
                raise TypeError(f"{type(g).__name__} is not a goal the fused loss kernel evaluates")
        self.alpha = float(alpha)
        self.name = name or type(self).__name__

    def number_of_goals(self):
        return len(self.goals)

    @staticmethod
    def error(gstate):
        raise NotImplementedError

    def errors(self, total):
        if self.mean:
            total = total / self.number_of_goals()
        return total.sqrt() if self.root else total

    def evaluate_state(self, eq_state, structure):
        """errors.py:165-200: the same value goal by goal in torch (tests, small problems; not the hot path)."""
        total = sum(self.error(g(eq_state, structure)) for g in self.goals)
        return self.errors(total) * self.alpha


class SquaredError(Error):
    """errors.py:232-258."""
    err_kind = 0

    @staticmethod
    def error(gs):
        return _over_features(gs.weight * (gs.prediction - gs.goal) ** 2, gs)


class MeanSquaredError(SquaredError):
    """errors.py:261-281."""
    mean = True


class RootMeanSquaredError(MeanSquaredError):
    """errors.py:284-301."""
    root = True


class PredictionError(Error):
    """errors.py:304-330."""
    err_kind = 2

    @staticmethod
    def error(gs):
        return _over_features(gs.weight * gs.prediction, gs)


class MeanPredictionError(PredictionError):
    """errors.py:333-352."""
    mean = True


class AbsoluteError(Error):
    """errors.py:355-380."""
    err_kind = 1

    @staticmethod
    def error(gs):
        return _over_features(gs.weight * (gs.prediction - gs.goal).abs(), gs)


class MeanAbsoluteError(AbsoluteError):
    """errors.py:383-400."""
    mean = True


class LogMaxError(Error):
    """errors.py:403-431."""
    err_kind = 3

    @staticmethod
    def error(gs):
        return _over_features(gs.weight * torch.log1p(torch.clamp(gs.prediction - gs.goal, min=0.0)), gs)


class Regularizer:
    """losses/regularizers.py:14-40."""

    def __init__(self, alpha, name=None):
        self.alpha = float(alpha)
        self.name = name or type(self).__name__

    def __call__(self, params):
        raise NotImplementedError("Regularizer subclasses must implement __call__")


class L2Regularizer(Regularizer):
    """regularizers.py:43-70: alpha * sum(q^2)."""

    def __call__(self, params):
        return self.alpha * (params.q ** 2).sum(dim=-1)


class _Program:
    """A Loss's error terms compiled against one structure: device tables + the fdm_goal_program struct."""

    def __init__(self, errors, structure, device):
        kind, index, err, weight, target, gptr, gfinal, ginner, gouter = [], [], [], [], [], [0], [], [], []
        agg_ptr, agg_index, n_agg = [0], [], 0
        for term in errors:
            for g in term.goals:
                if getattr(g, "is_aggregate", False):         # index = running number of the aggregate, set = its elements
                    members = [int(i) for i in g.indices(structure)]
                    limit = structure.num_nodes if g.kind == 19 else structure.num_edges
                    if not members or min(members) < 0 or max(members) >= limit:
                        raise IndexError(f"{type(g).__name__}: element set outside [0, {limit})")
                    i = n_agg
                    n_agg += 1
                    agg_index += members
                else:
                    i = int(g.index(structure))
                    limit = structure.num_edges if g.kind in (3, 4, 5, 15, 16) else structure.num_nodes
                    if not 0 <= i < limit:                      # the kernel trusts the table
                        raise IndexError(f"{type(g).__name__}: element index {i} outside [0, {limit})")
                agg_ptr.append(len(agg_index))
                kind.append(g.kind)
                index.append(i)
                err.append(term.err_kind)
                weight.append(g.weight)
                t = np.zeros(6)
                t[:g.target.size] = g.target.ravel()
                target.append(t)
            gptr.append(len(kind))
            gfinal.append(1 if term.root else 0)
            ginner.append(1.0 / term.number_of_goals() if term.mean else 1.0)
            gouter.append(term.alpha)
        self.num_terms, self.num_groups = len(kind), len(gfinal)
        # which of (xyz, lengths, forces, residuals) the terms differentiate into: the reverse pass neither zeroes nor
        # writes the other cotangent arrays (include/fdm_b200.h: fdm_goal_kind_arrays)
        mask = 0
        for k in set(kind):
            mask |= int(L.lib().fdm_goal_kind_arrays(int(k)))
        self.touches = tuple(bool(mask & bit) for bit in (1, 2, 4, 8))

        host = [(kind, np.int32), (index, np.int32), (err, np.int32), (weight, np.float64),
                (np.asarray(target).reshape(-1, 6), np.float64), (gptr, np.int32), (gfinal, np.int32),
                (ginner, np.float64), (gouter, np.float64), (agg_ptr, np.int32),
                (agg_index if agg_index else [0], np.int32)]
        # the same tables on the host (fdm_fa_set_goal_program copies them to the device itself)
        self.host_arrays = [np.ascontiguousarray(a, dtype=dt) for a, dt in host]
        self.host_struct = L.GoalProgram(self.num_terms, self.num_groups, int(6 in kind), n_agg,
                                         *[C.c_void_p(a.ctypes.data) for a in self.host_arrays])
        self.arrays, self.struct = None, None
        if device is not None:
            self.arrays = [torch.as_tensor(a, device=device) for a in self.host_arrays]
            ptr = [C.c_void_p(a.data_ptr()) for a in self.arrays]
            self.struct = L.GoalProgram(self.num_terms, self.num_groups, int(6 in kind), n_agg, *ptr)


def _goals_call(structure, program, xyz, lengths, forces, residuals, loss, loss_cot, bars):
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    st = C.c_void_p(torch.cuda.current_stream(xyz.device).cuda_stream)
    B = xyz.shape[0] if xyz.dim() == 3 else 1
    L.check(L.lib().fdm_goals_loss(structure.plan, C.byref(program.struct), P(xyz), P(lengths), P(forces), P(residuals),
                                   P(loss), None, P(loss_cot), *[P(t) for t in bars], B, st))


class _GoalsLoss(torch.autograd.Function):
    """Forward: one launch, the loss per design.  Backward: one launch of the same kernel in its reverse mode --
    d loss / d (xyz, lengths, forces, residuals) already scaled by the upstream cotangent (no elementwise passes)."""

    @staticmethod
    def forward(ctx, xyz, lengths, forces, residuals, structure, program):
        batched = xyz.dim() == 3
        B = xyz.shape[0] if batched else 1
        xyz, lengths, forces, residuals = (t.contiguous() for t in (xyz, lengths, forces, residuals))
        loss = torch.empty((B,), dtype=F64, device=xyz.device)
        _goals_call(structure, program, xyz, lengths, forces, residuals, loss, None, [None] * 4)
        ctx.save_for_backward(xyz, lengths, forces, residuals)
        ctx.structure, ctx.program, ctx.batched = structure, program, batched
        return loss if batched else loss[0]

    @staticmethod
    def backward(ctx, g):
        xyz, lengths, forces, residuals = ctx.saved_tensors
        touched = ctx.program.touches
        bars = [torch.empty_like(t) if on else None for t, on in zip((xyz, lengths, forces, residuals), touched)]
        if any(touched):
            _goals_call(ctx.structure, ctx.program, xyz, lengths, forces, residuals, None, g.reshape(-1).contiguous(), bars)
        return (*bars, None, None)


class _FusedValueAndGrad(torch.autograd.Function):
    """
    `Loss(params, model, structure)` for the linear model (tmax = 1, nodal loads) as ONE autograd node: forward =
    solve from q -> positions -> state -> goal kernel; backward = goal kernel (reverse mode, scaled by the upstream
    cotangent) -> state VJP -> split of xyz_bar -> adjoint solve with the stored factor -> gradient kernel
    accumulating on the direct terms.  The same launches as the library's own pipeline (fdm_fa_*), with no autograd
    bookkeeping or elementwise passes between them.
    """

    @staticmethod
    def forward(ctx, q, xyz_fixed, loads, structure, program, solve_opts):
        from ..equilibrium import ops
        s = structure
        ws = ops.Workspace()
        x = K = None
        opts = {k: v for k, v in solve_opts.items() if k != "fused_assembly"}
        if solve_opts.get("fused_assembly", True) and opts.get("solver", "auto") in ("auto", "cholesky"):
            x = ops.solve_from_q(s, q, xyz_fixed, loads, workspace=ws, nan_on_failure=True)
            if x is not None:
                opts["solver"] = "cholesky"
        if x is None:
            K, rhs = ops.assemble(s, q, xyz_fixed, loads)
            x = ops.solve(s, K, rhs, workspace=ws, nan_on_failure=True, **opts)
        xyz = ops.positions(s, x, xyz_fixed)
        # a loss of node goals only reads the positions: no state kernel (and no state VJP on the way back)
        ctx.needs_state = any(program.touches[1:])
        if ctx.needs_state:
            _, lengths, forces, residuals = ops.state(s, q, xyz, loads)
        else:
            lengths = forces = residuals = None
        batched = xyz.dim() == 3
        loss = torch.empty((xyz.shape[0] if batched else 1,), dtype=F64, device=xyz.device)
        _goals_call(s, program, xyz, lengths, forces, residuals, loss, None, [None] * 4)
        e = x.new_empty(0)
        ctx.save_for_backward(q, xyz_fixed, loads, xyz, *(t if t is not None else e for t in (lengths, forces, residuals)),
                              K if K is not None else e)
        ctx.structure, ctx.program, ctx.opts, ctx.ws, ctx.batched = s, program, opts, ws, batched
        return loss if batched else loss[0]

    @staticmethod
    def backward(ctx, g):
        from ..equilibrium import ops
        from ..distributed import _batch_sum
        q, xyz_fixed, loads, xyz, lengths, forces, residuals, K = ctx.saved_tensors
        s = ctx.structure
        touched = ctx.program.touches                       # untouched state arrays: no cotangent array at all
        if not ctx.needs_state:
            lengths = forces = residuals = None
        bars = [torch.empty_like(xyz) if touched[0] else None] + [
            torch.empty_like(t) if on else None for t, on in zip((lengths, forces, residuals), touched[1:])]
        if any(touched):
            _goals_call(s, ctx.program, xyz, lengths, forces, residuals, None, g.reshape(-1).contiguous(), bars)
        if ctx.needs_state or bars[0] is None:
            q_bar, xyz_bar, loads_bar = ops.state_vjp(s, q, xyz, xyz_cot=bars[0], residuals_cot=bars[3],
                                                      lengths_cot=bars[1], forces_cot=bars[2])
        else:                                               # only xyz was read: its cotangent IS xyz_bar, no direct terms
            xyz_bar = bars[0]
            q_bar = torch.zeros(xyz.shape[:-2] + (s.num_edges,), dtype=F64, device=xyz.device)
            loads_bar = torch.zeros_like(xyz)
        gx, xs_bar = ops.positions_vjp(s, xyz_bar)
        lam = ops.solve(s, K if K.numel() else None, gx, workspace=ctx.ws, reuse_factor=True, **ctx.opts)
        ops.adjoint_grads(s, q, xyz, lam, q_bar=q_bar, loads_bar=loads_bar, xyz_fixed_bar=xs_bar)

        def to_shape(grad, like):                          # a parameter shared by the batch takes the sum over designs
            if grad.dim() == like.dim():
                return grad
            out = torch.empty(like.shape, dtype=F64, device=grad.device)
            return _batch_sum(grad.reshape(grad.shape[0], -1), out.view(-1)).view(like.shape)
        return (to_shape(q_bar, q) if ctx.needs_input_grad[0] else None,
                to_shape(xs_bar, xyz_fixed) if ctx.needs_input_grad[1] else None,
                to_shape(loads_bar, loads) if ctx.needs_input_grad[2] else None, None, None, None)


class Loss:
    """losses/loss.py:22-171."""

    def __init__(self, *args, name=None):
        self.terms_error = [t for t in args if isinstance(t, Error)]
        self.terms_regularization = [t for t in args if isinstance(t, Regularizer)]
        self.name = name or type(self).__name__
        self._programs = {}

    @property
    def terms(self):
        return self.terms_error + self.terms_regularization

    def number_of_goals(self):
        return sum(t.number_of_goals() for t in self.terms_error)

    def number_of_regularizers(self):
        return len(self.terms_regularization)

    def program(self, structure, device):
        key = (id(structure), str(device))
        if key not in self._programs:
            self._programs[key] = _Program(self.terms_error, structure, device)
        return self._programs[key]

    def errors_from_state(self, eq_state, structure):
        """Sum of the error terms on an equilibrium state: one fused kernel (fdm_goals_loss)."""
        prog = self.program(structure, eq_state.xyz.device)
        return _GoalsLoss.apply(eq_state.xyz, eq_state.lengths, eq_state.forces, eq_state.residuals, structure, prog)

    def __call__(self, params, model, structure):
        load_state = params[2]
        nodes_load = load_state.nodes if hasattr(load_state, "nodes") else load_state
        shape_dependent = any(isinstance(t, torch.Tensor) and t.numel() > 1
                              for t in (getattr(load_state, "edges", 0.0), getattr(load_state, "faces", 0.0)))
        if getattr(model, "tmax", 1) == 1 and not shape_dependent and params[0].is_cuda and self.terms_error:
            # the linear model: forward + loss (+ its whole reverse pass) as one node of this library's kernels
            prog = self.program(structure, params[0].device)
            loss = _FusedValueAndGrad.apply(params[0], params[1], nodes_load, structure, prog, model.solve_opts)
        else:
            eq_state = model(params, structure)
            loss = self.errors_from_state(eq_state, structure)
        for reg in self.terms_regularization:
            loss = loss + reg(params)
        return loss
