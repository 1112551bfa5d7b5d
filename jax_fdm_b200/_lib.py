"""
ctypes binding of the C ABI in include/fdm_b200.h (libfdm_b200.so, built in-tree by
`__graft_entry__.build()` / `python -m jax_fdm_b200.build`).

There is no fallback: if the shared library is missing, import of any compute entry
point raises.  The CPU oracle under oracle/ is never imported from here.
"""

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# FDM_B200_LIB: alternative build of the same library (e.g. the -DFDM_PCG2_PROFILE diagnostics build)
LIB_PATH = os.environ.get("FDM_B200_LIB") or os.path.join(HERE, "lib", "libfdm_b200.so")

# ---- constants mirrored from include/fdm_b200.h ----
FDM_OK = 0
ERRORS = {-1: "INVALID", -2: "CUDA", -3: "MULTI_EDGE", -4: "NO_SUPPORT", -5: "UNSUPPORTED", -6: "WORKSPACE"}
STATUS_OK, STATUS_NOT_CONVERGED, STATUS_INDEFINITE, STATUS_BREAKDOWN = 0, 1, 2, 3
INFO_STATUS, INFO_ITERS, INFO_RELRES, INFO_SIGN, INFO_STRIDE = 0, 1, 2, 3, 8
SIZE_NODES, SIZE_EDGES, SIZE_FREE, SIZE_FIXED, SIZE_NNZ, SIZE_DIAGS_NNZ, SIZE_HBW, SIZE_PARTS = range(8)
(ARR_INDICES_FREE, ARR_INDICES_FIXED, ARR_INDICES_FREEFIXED, ARR_INDEX_DATA, ARR_INDEX_INDICES,
 ARR_INDEX_INDPTR, ARR_DIAG_INDICES, ARR_DIAGS_INDPTR, ARR_DIAGS_INDICES, ARR_BAND_PERM,
 ARR_PART_OF_ROW) = range(11)
SOLVER_AUTO, SOLVER_PCG, SOLVER_PCG_PERSISTENT, SOLVER_CHOLESKY_BATCHED = 0, 1, 2, 3
SOLVERS = {"auto": 0, "pcg": 1, "pcg_persistent": 2, "cholesky": 3, "ldlt": 4}

EXPORTS = [
    "fdm_plan_create", "fdm_plan_create_host", "fdm_plan_destroy", "fdm_plan_size",
    "fdm_plan_get_array", "fdm_assemble", "fdm_workspace_bytes", "fdm_solve", "fdm_solve_from_q", "fdm_spmm",
    "fdm_positions", "fdm_state", "fdm_state_vjp", "fdm_state_jvp", "fdm_adjoint_grads", "fdm_edge_loads",
    "fdm_edge_loads_vjp", "fdm_face_loads", "fdm_face_loads_vjp", "fdm_goals_loss", "fdm_goal_kind_arrays", "fdm_loss_sqdist",
    "fdm_fa_create", "fdm_fa_destroy", "fdm_fa_set_constants", "fdm_fa_set_goal_program", "fdm_fa_set_q",
    "fdm_fa_step_device", "fdm_fa_enqueue_host", "fdm_fa_enqueue_copies_only", "fdm_fa_wait", "fdm_fa_run_host", "fdm_fa_device_array", "fdm_fa_info",
    "fdm_fp_create", "fdm_fp_destroy", "fdm_fp_forward", "fdm_fp_adjoint", "fdm_fp_info",
    "fdm_host_alloc", "fdm_host_free", "fdm_batch_sum", "fdm_lincomb", "fdm_kbar_pattern", "fdm_stiffness_vjp",
    "fdm_last_error", "fdm_version",
]


class SolveOpts(C.Structure):
    _fields_ = [("solver", C.c_int32), ("max_iters", C.c_int32), ("tol", C.c_double),
                ("check_every", C.c_int32), ("reuse_factor", C.c_int32), ("stages", C.c_int32),
                ("nan_on_failure", C.c_int32), ("concurrent_designs", C.c_int32)]


class FpOpts(C.Structure):
    """fdm_fp_opts (include/fdm_b200.h)."""
    _fields_ = [("tmax", C.c_int32), ("is_load_local", C.c_int32), ("adjoint_max_iters", C.c_int32), ("eta", C.c_double),
                ("adjoint_tol", C.c_double)]


class FaOpts(C.Structure):
    """fdm_fa_opts (include/fdm_b200.h)."""
    _fields_ = [("chunks", C.c_int32), ("with_state", C.c_int32), ("solver", C.c_int32), ("use_graph", C.c_int32),
                ("max_iters", C.c_int32), ("tol", C.c_double)]


(FA_Q, FA_X_FREE, FA_XYZ, FA_LOSS, FA_Q_BAR, FA_LOADS_BAR, FA_XYZ_FIXED_BAR, FA_INFO_FORWARD, FA_INFO_ADJOINT,
 FA_LENGTHS, FA_FORCES, FA_RESIDUALS, FA_VECTORS, FA_XYZ_FIXED, FA_LOADS) = range(15)
FA_INFO_CHUNKS, FA_INFO_LAUNCHES, FA_INFO_FUSED_ASSEMBLY, FA_INFO_GRAPH, FA_INFO_BATCH = range(5)


class GoalProgram(C.Structure):
    """fdm_goal_program (include/fdm_b200.h): device pointers of a compiled Loss."""
    _fields_ = [("num_terms", C.c_int32), ("num_groups", C.c_int32), ("has_network_loadpath", C.c_int32),
                ("num_aggregates", C.c_int32), ("kind", C.c_void_p), ("index", C.c_void_p), ("error", C.c_void_p), ("weight", C.c_void_p),
                ("target", C.c_void_p), ("group_ptr", C.c_void_p), ("group_final", C.c_void_p),
                ("group_inner", C.c_void_p), ("group_outer", C.c_void_p), ("agg_ptr", C.c_void_p),
                ("agg_index", C.c_void_p)]


class FdmError(RuntimeError):
    def __init__(self, code, text):
        self.code = code
        super().__init__(f"fdm_b200 error {code} ({ERRORS.get(code, '?')}): {text}")


_lib = None


def lib():
    """Load (once) and return the shared library; raise loudly when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_size_t
    L.fdm_plan_create.argtypes = [vp, i64, i64, vp, C.c_int, C.POINTER(vp)]
    L.fdm_plan_create.restype = C.c_int
    L.fdm_plan_create_host.argtypes = [vp, i64, i64, vp, C.POINTER(vp)]
    L.fdm_plan_create_host.restype = C.c_int
    L.fdm_plan_destroy.argtypes = [vp]
    L.fdm_plan_destroy.restype = None
    L.fdm_plan_size.argtypes = [vp, C.c_int]
    L.fdm_plan_size.restype = i64
    L.fdm_plan_get_array.argtypes = [vp, C.c_int, vp, i64]
    L.fdm_plan_get_array.restype = i64
    L.fdm_assemble.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, vp]
    L.fdm_assemble.restype = C.c_int
    L.fdm_workspace_bytes.argtypes = [vp, i64, i32]
    L.fdm_workspace_bytes.restype = sz
    L.fdm_solve.argtypes = [vp, vp, vp, vp, vp, vp, sz, i64, C.POINTER(SolveOpts), vp]
    L.fdm_solve.restype = C.c_int
    L.fdm_solve_from_q.argtypes = [vp, vp, vp, vp, vp, vp, vp, sz, i64, i64, i64, i64, C.POINTER(SolveOpts), vp]
    L.fdm_solve_from_q.restype = C.c_int
    L.fdm_spmm.argtypes = [vp, vp, vp, vp, i64, vp]
    L.fdm_spmm.restype = C.c_int
    L.fdm_positions.argtypes = [vp, vp, vp, vp, i64, i64, C.c_int, vp]
    L.fdm_positions.restype = C.c_int
    L.fdm_state.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, vp]
    L.fdm_state.restype = C.c_int
    L.fdm_state_vjp.argtypes = [vp] + [vp] * 8 + [vp] * 3 + [i64, i64, vp]
    L.fdm_state_vjp.restype = C.c_int
    L.fdm_state_jvp.argtypes = [vp] * 10 + [i64] * 5 + [vp]
    L.fdm_state_jvp.restype = C.c_int
    L.fdm_adjoint_grads.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, C.c_int, vp]
    L.fdm_adjoint_grads.restype = C.c_int
    L.fdm_edge_loads.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, C.c_int, vp]
    L.fdm_edge_loads.restype = C.c_int
    L.fdm_edge_loads_vjp.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, C.c_int, vp]
    L.fdm_edge_loads_vjp.restype = C.c_int
    L.fdm_face_loads.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, C.c_int, vp]
    L.fdm_face_loads.restype = C.c_int
    L.fdm_face_loads_vjp.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, C.c_int, vp]
    L.fdm_face_loads_vjp.restype = C.c_int
    L.fdm_goals_loss.argtypes = [vp, C.POINTER(GoalProgram), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
    L.fdm_goals_loss.restype = C.c_int
    L.fdm_goal_kind_arrays.argtypes = [C.c_int]
    L.fdm_goal_kind_arrays.restype = C.c_int
    L.fdm_loss_sqdist.argtypes = [vp, vp, vp, vp, i64, i64, i64, vp]
    L.fdm_loss_sqdist.restype = C.c_int
    L.fdm_fa_create.argtypes = [vp, i64, C.POINTER(FaOpts), C.POINTER(vp)]
    L.fdm_fa_create.restype = C.c_int
    L.fdm_fa_destroy.argtypes = [vp]
    L.fdm_fa_destroy.restype = None
    L.fdm_fa_set_constants.argtypes = [vp, vp, vp, vp]
    L.fdm_fa_set_constants.restype = C.c_int
    L.fdm_fa_set_goal_program.argtypes = [vp, C.POINTER(GoalProgram)]
    L.fdm_fa_set_goal_program.restype = C.c_int
    L.fdm_fa_set_q.argtypes = [vp, vp]
    L.fdm_fa_set_q.restype = C.c_int
    L.fdm_fa_step_device.argtypes = [vp, vp]
    L.fdm_fa_step_device.restype = C.c_int
    L.fdm_fa_enqueue_host.argtypes = [vp, vp, vp, vp]
    L.fdm_fa_enqueue_host.restype = C.c_int
    L.fdm_fa_enqueue_copies_only.argtypes = [vp, vp, vp, vp]
    L.fdm_fa_enqueue_copies_only.restype = C.c_int
    L.fdm_fa_wait.argtypes = [vp]
    L.fdm_fa_wait.restype = C.c_int
    L.fdm_fa_run_host.argtypes = [vp, vp, vp, vp]
    L.fdm_fa_run_host.restype = C.c_int
    L.fdm_fa_device_array.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(i64)]
    L.fdm_fa_device_array.restype = C.c_int
    L.fdm_fa_info.argtypes = [vp, C.c_int]
    L.fdm_fa_info.restype = i64
    L.fdm_kbar_pattern.argtypes = [vp, vp, vp, vp, C.c_double, i64, vp]
    L.fdm_kbar_pattern.restype = C.c_int
    L.fdm_stiffness_vjp.argtypes = [vp, vp, vp, i64, vp]
    L.fdm_stiffness_vjp.restype = C.c_int
    L.fdm_lincomb.argtypes = [vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, i64, vp]
    L.fdm_lincomb.restype = C.c_int
    L.fdm_batch_sum.argtypes = [vp, vp, i64, i64, vp]
    L.fdm_batch_sum.restype = C.c_int
    L.fdm_fp_create.argtypes = [vp, C.POINTER(FpOpts), C.c_int, vp, i64, i64, vp, vp, vp, vp, C.POINTER(vp)]
    L.fdm_fp_create.restype = C.c_int
    L.fdm_fp_destroy.argtypes = [vp]
    L.fdm_fp_destroy.restype = None
    L.fdm_fp_forward.argtypes = [vp] * 9 + [vp]
    L.fdm_fp_forward.restype = C.c_int
    L.fdm_fp_adjoint.argtypes = [vp, vp, vp, vp, vp]
    L.fdm_fp_adjoint.restype = C.c_int
    L.fdm_fp_info.argtypes = [vp, vp, vp]
    L.fdm_fp_info.restype = C.c_int
    L.fdm_host_alloc.argtypes = [sz]
    L.fdm_host_alloc.restype = vp
    L.fdm_host_free.argtypes = [vp]
    L.fdm_host_free.restype = None
    L.fdm_last_error.argtypes = []
    L.fdm_last_error.restype = C.c_char_p
    L.fdm_version.argtypes = []
    L.fdm_version.restype = C.c_char_p
    _lib = L
    return L


def check(rc):
    if rc != FDM_OK:
        raise FdmError(rc, lib().fdm_last_error().decode())
