/*
 * fdm_b200.h -- C ABI of the B200-native force-density equilibrium hot path.
 *
 * Drop-in boundary for the equilibrium path of arpastrana/jax_fdm.  Every entry point
 * names the reference interface it replaces (file:line under src/jax_fdm/equilibrium/).
 * The reference has no FFI of its own (it is pure Python/JAX); these are the functions a
 * `jax.ffi` / ctypes binding on the reference side would bind.
 *
 * Conventions
 *   - plain C: pointers, sizes, int return codes (0 = OK, negative = error; text via
 *     fdm_last_error(), thread-local).  No torch / JAX types in any signature.
 *   - all `double*` / `int*` arguments named d_* are DEVICE pointers owned by the caller;
 *     h_* are HOST pointers.  Arrays are row-major, FP64 values (the reference fixes
 *     float64 at import, src/jax_fdm/__init__.py:62-72).
 *   - every compute call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     allocates nothing, never synchronises, and is CUDA-graph capturable -- except where a
 *     function says otherwise.  Scratch memory is caller-provided (fdm_workspace_bytes).
 *   - batch: B independent designs sharing one topology (the `vmap(model, in_axes=(0, None))`
 *     of visualization/plotters/loss_plotter.py:98-103).  Per-array batch strides are in
 *     ELEMENTS; stride 0 means "shared by all designs".
 */
#ifndef FDM_B200_H
#define FDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fdm_plan fdm_plan;

/* ---- error codes ------------------------------------------------------------------ */
enum {
  FDM_OK = 0,
  FDM_ERR_INVALID = -1,        /* bad argument / size                                   */
  FDM_ERR_CUDA = -2,           /* CUDA runtime error (text in fdm_last_error)           */
  FDM_ERR_MULTI_EDGE = -3,     /* parallel edges or self loops: reference index_array is ill-defined
                                  (structures.py:248-288 sums e+1 over duplicates)      */
  FDM_ERR_NO_SUPPORT = -4,     /* no fixed node / no edge (fdm.py:491-522 asserts)      */
  FDM_ERR_UNSUPPORTED = -5,    /* configuration outside this build (e.g. band too wide) */
  FDM_ERR_WORKSPACE = -6       /* workspace too small                                   */
};

/* device-side solve status, written to info[FDM_INFO_STATUS] */
enum {
  FDM_STATUS_OK = 0,
  FDM_STATUS_NOT_CONVERGED = 1,
  FDM_STATUS_INDEFINITE = 2,   /* mixed-sign diagonal / pivot: K neither SPD nor -SPD.  The batched band Cholesky
                                  solves such designs itself when hbw <= 30 and FDM_SOLVER_LDLT_BAND takes over
                                  behind the other solvers when hbw <= 255 (FDM_INFO_SIGN 0); the status only
                                  remains for wider bands or when the fallback's workspace slots run out */
  FDM_STATUS_BREAKDOWN = 3     /* non-finite scalar met in the iteration / vanishing pivot */
};

/* layout of the per-solve info record (doubles), one record per design of the batch */
enum {
  FDM_INFO_STATUS = 0,
  FDM_INFO_ITERS = 1,
  FDM_INFO_RELRES = 2,         /* max over rhs of ||r|| / ||b|| (recurrence residual)   */
  FDM_INFO_SIGN = 3,           /* +1: K SPD, -1: -K SPD, 0: indefinite K factored as L D L^T (signed pivots) */
  FDM_INFO_STRIDE = 8
};

/* ---- plan: topology -> static device arrays ----------------------------------------
 * Replaces Graph.__init__/_edges_indexed (structures/graphs.py:44-97, done by the caller:
 * `h_edges` are already node INDICES), connectivity_matrix (graphs.py:176-211),
 * EquilibriumStructure.__init__ + _indices_* (structures/structures.py:53-69, 152-188) and
 * EquilibriumStructureSparse.__init__ (structures.py:213-350: index_array, diag_indices, diags).
 * Host work + uploads; synchronous; the only call that allocates device memory.
 *   h_edges    (E,2) int64  tail, head node index
 *   h_supports (N)   int64  0 = free, nonzero = fixed
 */
int fdm_plan_create(const int64_t* h_edges, int64_t num_edges, int64_t num_nodes,
                    const int64_t* h_supports, int device, fdm_plan** out);
void fdm_plan_destroy(fdm_plan* plan);

/* host-only variant (no CUDA calls): builds every topology array but uploads nothing.
 * Used by CPU tests of the index maps; compute calls on such a plan return FDM_ERR_INVALID. */
int fdm_plan_create_host(const int64_t* h_edges, int64_t num_edges, int64_t num_nodes,
                         const int64_t* h_supports, fdm_plan** out);

enum {
  FDM_SIZE_NODES = 0, FDM_SIZE_EDGES = 1, FDM_SIZE_FREE = 2, FDM_SIZE_FIXED = 3,
  FDM_SIZE_NNZ = 4,            /* stored entries of K = Nf + 2 * (#free-free edges)     */
  FDM_SIZE_DIAGS_NNZ = 5,      /* stored entries of `diags`                             */
  FDM_SIZE_HALF_BANDWIDTH = 6, /* of K after the plan's band ordering (batched Cholesky) */
  FDM_SIZE_NUM_PARTS = 7       /* aggregates of the two-level preconditioner            */
};
int64_t fdm_plan_size(const fdm_plan* plan, int which);

/* Topology arrays as the reference defines them; copied to a HOST
 * buffer of `capacity` bytes.  Returns bytes written or a negative error.  dtypes follow the
 * reference (int64 index maps; index_array.indices/indptr int32 as scipy yields them). */
enum {
  FDM_ARR_INDICES_FREE = 0,      /* (Nf)   int64  structures.py:152-161                  */
  FDM_ARR_INDICES_FIXED = 1,     /* (Ns)   int64  structures.py:163-172                  */
  FDM_ARR_INDICES_FREEFIXED = 2, /* (N)    int64  structures.py:174-188                  */
  FDM_ARR_INDEX_DATA = 3,        /* (nnz)  int64  index_array.data   structures.py:248-288 */
  FDM_ARR_INDEX_INDICES = 4,     /* (nnz)  int32  index_array.indices                    */
  FDM_ARR_INDEX_INDPTR = 5,      /* (Nf+1) int32  index_array.indptr                     */
  FDM_ARR_DIAG_INDICES = 6,      /* (Nf)   int64  structures.py:290-314                  */
  FDM_ARR_DIAGS_INDPTR = 7,      /* (Nf+1) int32  diags (BCSR Nf x E) structures.py:316-350 */
  FDM_ARR_DIAGS_INDICES = 8,     /* (dnnz) int32                                         */
  FDM_ARR_BAND_PERM = 9,         /* (Nf)   int32  free index -> band-ordered index       */
  FDM_ARR_PART_OF_ROW = 10,      /* (Nf)   int32  free index -> aggregate id             */
  /* tables of the two-sided factorisation (diagnostics / tests), half h = 0, 1 at +h: row -> free index (-1 =
   * padding or the other half's row), first entry group of every 32-row block, and the (slot, K.data position)
   * and (slot, edge id) entry lists; empty when the plan has no two-sided tables */
  FDM_ARR_TW_IPERM = 11, FDM_ARR_TW_STAGE_PTR = 13, FDM_ARR_TW_STAGE_KENT = 15, FDM_ARR_TW_STAGE_EENT = 17
};
int64_t fdm_plan_get_array(const fdm_plan* plan, int which, void* h_out, int64_t capacity);

/* ---- K1: assembly --------------------------------------------------------------------
 * Replaces EquilibriumModelSparse.stiffness_matrix (models.py:1037-1079):
 *     K.data[k] = -q[index_array.data[k]-1],  K.data[diag_indices] = diags @ q
 * and load_matrix / residual_fixed_matrix (models.py:828-856, 949-976):
 *     rhs = loads[indices_free] - C_f^T (q * (C_s xyz_fixed)).
 *   d_q (B,E)  d_xyz_fixed (B|1,Ns,3)  d_loads (B|1,N,3)  ->  d_Kdata (B,nnz)  d_rhs (B,Nf,3)
 * d_Kdata or d_rhs may be NULL to skip that half.
 * A batch (>= 8 designs) on a small topology takes streaming kernels that read the index streams once per launch
 * (bit-identical output); every call is enqueued on `stream` only. */
int fdm_assemble(const fdm_plan* plan, const double* d_q, const double* d_xyz_fixed,
                 const double* d_loads, double* d_Kdata, double* d_rhs, int64_t batch,
                 int64_t q_stride, int64_t xyz_fixed_stride, int64_t loads_stride, void* stream);

/* ---- K2: the 3-rhs solve ---------------------------------------------------------------
 * Replaces sparse_solve / spsolve_cpu / spsolve_gpu_ravel (sparse.py:207-230, 130-159, 54-84):
 * x = K^{-1} rhs with K given by its CSC data array on the plan's pattern (K symmetric).
 * Also serves the adjoint solve lam = K^{-1} g of sparse_solve_bwd (sparse.py:296).
 */
enum {
  FDM_SOLVER_AUTO = 0,
  FDM_SOLVER_PCG = 1,            /* Jacobi-PCG, one kernel per phase (any size)          */
  FDM_SOLVER_PCG_PERSISTENT = 2, /* two-level PCG, one cooperative kernel, K and vectors
                                    resident in shared memory (meshes that fit)          */
  FDM_SOLVER_CHOLESKY_BATCHED = 3, /* one CTA per design, banded LL^T in shared memory    */
  FDM_SOLVER_LDLT_BAND = 4       /* band L D L^T without pivoting in global memory, one CTA per design, half
                                    bandwidth <= 255: what FDM_SOLVER_AUTO falls back to, design by design and
                                    without a host round trip, when the solver above reports
                                    FDM_STATUS_INDEFINITE (mixed-sign q: the reference's LU, sparse.py:147-149,
                                    solves such systems); slow (two CTA barriers per column), kept for coverage */
};

typedef struct fdm_solve_opts {
  int32_t solver;       /* FDM_SOLVER_*                                                   */
  int32_t max_iters;    /* PCG; <=0 -> default 20000                                     */
  double tol;           /* PCG: stop when ||r|| <= tol * ||b|| for every rhs; <=0 -> 1e-12 */
  int32_t check_every;  /* multi-kernel PCG: host polls convergence every this many iterations
                           (synchronises the stream; not graph-capturable); <=0 -> 50    */
  int32_t reuse_factor; /* Cholesky: nonzero -> d_ws already holds the factor of this K (adjoint
                           solve after the forward solve, tmax>1 iterations): skip refactoring.
                           2 (warp-per-design band Cholesky): ONE factor, in the first design's slot of a workspace
                           sized for `batch` designs, serves every right-hand-side set of the batch (tangent /
                           multi-rhs solves: same K, many right-hand sides) */
  int32_t stages;       /* batched Cholesky, measurement only: 0 = the whole solve; otherwise a mask of the
                           launches to issue (1 factorisation [+ its signed second pass], 2 diagonal-block inversion
                           [separate launch at half bandwidth 31 only], 4 substitution),
                           so that a benchmark can time each kernel of one solve on its own */
  int32_t nan_on_failure; /* nonzero: a design whose status is not FDM_STATUS_OK gets x = NaN (what the reference's
                           solvers hand to the optimiser's NaN check, optimizer.py:372-380); needs d_info */
  int32_t concurrent_designs; /* batched Cholesky: designs in flight on the device including other streams' calls
                           (a sliced batch: the whole batch); 0 = this call's batch.  Small batches take the
                           two-sided factorisation (a warp pair per design: half the dependency chain), which only
                           pays while SMs are short of warps */
} fdm_solve_opts;

/* Bytes of scratch fdm_solve / fdm_solve_from_q want for `batch` designs with this solver choice.  It includes room
 * for the band L D L^T fallback behind solvers that need a definite K (up to 256 MB: min(batch, 256 MB / band size)
 * designs can fall back in one call; nothing when one band alone is larger); a smaller workspace is accepted as long
 * as it covers the solver itself -- the fallback is then skipped and such designs keep FDM_STATUS_INDEFINITE. */
size_t fdm_workspace_bytes(const fdm_plan* plan, int64_t batch, int32_t solver);

int fdm_solve(const fdm_plan* plan, const double* d_Kdata, const double* d_rhs, double* d_x,
              double* d_info /* (B, FDM_INFO_STRIDE) or NULL */, void* d_ws, size_t ws_bytes,
              int64_t batch, const fdm_solve_opts* opts, void* stream);

/* K1 + K2 in one call: x = K(q)^{-1} (loads_f - C_f^T Q C_s xyz_fixed), i.e. nodes_free_positions
 * (models.py:222-256), without materialising K.data or the right-hand side.  Available where the
 * batched band Cholesky assembles its rows from q in place (half bandwidth <= 30); otherwise returns
 * FDM_ERR_UNSUPPORTED and the caller uses fdm_assemble + fdm_solve.  The factor left in d_ws serves
 * later fdm_solve calls with opts.reuse_factor (the adjoint solve). */
int fdm_solve_from_q(const fdm_plan* plan, const double* d_q, const double* d_xyz_fixed,
                     const double* d_loads, double* d_x, double* d_info, void* d_ws, size_t ws_bytes,
                     int64_t batch, int64_t q_stride, int64_t xyz_fixed_stride, int64_t loads_stride,
                     const fdm_solve_opts* opts, void* stream);

/* ---- SpMM: Y = K X, X (Nf,3) -----------------------------------------------------------
 * Replaces `K @ xyz_free` (models.py:1012) and `A @ xk` (sparse.py:301). */
int fdm_spmm(const fdm_plan* plan, const double* d_Kdata, const double* d_X, double* d_Y,
             int64_t batch, void* stream);

/* ---- positions ---------------------------------------------------------------------------
 * Replaces nodes_positions (models.py:195-220): xyz = concat(x_free, xyz_fixed)[indices_freefixed]
 * and, with transpose != 0, its VJP: x_free_bar = xyz_bar[indices_free], xyz_fixed_bar = xyz_bar[indices_fixed].
 */
int fdm_positions(const fdm_plan* plan, double* d_x_free, double* d_xyz_fixed, double* d_xyz,
                  int64_t batch, int64_t xyz_fixed_stride, int transpose, void* stream);

/* ---- K4: equilibrium state and its VJP ---------------------------------------------------
 * Replaces equilibrium_state (models.py:753-794): vectors = C xyz, lengths = |vectors|,
 * residuals = loads - C^T (q * vectors), forces = q * lengths.  Any output may be NULL.
 */
int fdm_state(const fdm_plan* plan, const double* d_q, const double* d_xyz, const double* d_loads,
              double* d_vectors /*(B,E,3)*/, double* d_lengths /*(B,E)*/, double* d_forces /*(B,E)*/,
              double* d_residuals /*(B,N,3)*/, int64_t batch, int64_t q_stride,
              int64_t loads_stride, void* stream);

/* VJP of fdm_state (JAX autodiff of models.py:780-794).  Input cotangents may be NULL (= 0).
 * Outputs are OVERWRITTEN: q_bar (B,E), xyz_bar (B,N,3), loads_bar (B,N,3) [loads_bar =
 * loads_cot + residuals_cot].  A zero-length edge with nonzero length/force cotangent gives
 * NaN, like jnp.linalg.norm's gradient. */
int fdm_state_vjp(const fdm_plan* plan, const double* d_q, const double* d_xyz,
                  const double* d_xyz_cot, const double* d_residuals_cot,
                  const double* d_lengths_cot, const double* d_forces_cot,
                  const double* d_loads_cot, const double* d_vectors_cot,
                  double* d_q_bar, double* d_xyz_bar, double* d_loads_bar,
                  int64_t batch, int64_t q_stride, void* stream);

/* JVP of fdm_state: tangents of vectors / lengths / forces / residuals for tangents (q_dot, xyz_dot, loads_dot) at
 * the primal (q, xyz).  `batch` tangents; a primal stride of 0 shares one primal among them (jacfwd-style use: the
 * reference forces the dense model for this, fdm.py:266-271).  q_dot, loads_dot may be NULL (= 0); any output may
 * be NULL.  Strides in elements. */
int fdm_state_jvp(const fdm_plan* plan, const double* d_q, const double* d_xyz, const double* d_q_dot,
                  const double* d_xyz_dot, const double* d_loads_dot, double* d_vectors_dot, double* d_lengths_dot,
                  double* d_forces_dot, double* d_residuals_dot, int64_t batch, int64_t q_stride, int64_t xyz_stride,
                  int64_t q_dot_stride, int64_t loads_dot_stride, void* stream);

/* ---- K3: gradients of the implicit solve ---------------------------------------------------
 * Replaces the tail of sparse_solve_bwd (sparse.py:299-308) plus the autodiff transposes of
 * stiffness_matrix (models.py:1069-1077) and load_matrix (models.py:856, 973-976), given the
 * adjoint solution lam = K^{-1} g:
 *     q_bar_e        (+)= -(lam_head - lam_tail) . (xyz_head - xyz_tail)     (lam = 0 on supports)
 *     loads_bar[free] (+)= lam
 *     xyz_fixed_bar_s (+)= sum_{e=(s,j), j free} q_e lam_j
 * accumulate != 0 adds to what the output buffers hold (direct terms from fdm_state_vjp /
 * fdm_positions(transpose)); accumulate == 0 overwrites (loads_bar rows of supports <- 0).
 */
int fdm_adjoint_grads(const fdm_plan* plan, const double* d_q, const double* d_xyz,
                      const double* d_lam, double* d_q_bar, double* d_loads_bar,
                      double* d_xyz_fixed_bar, int64_t batch, int64_t q_stride, int accumulate,
                      void* stream);

/* ---- shape-dependent loads (tmax > 1) -------------------------------------------------------
 * Replaces nodes_load with non-scalar edge loads (models.py:290-344, loads.py:296-331, 340-465): line
 * loads scaled by the edge length, half to each end node; is_local != 0: the load is given in the edge's
 * local frame (geometry.py:581-618 line_lcs) and follows the deformed geometry:
 *     loads_out_n = loads_nodes_n + 1/2 sum_{e incident to n} frame_e(edges_load_e) |x_head - x_tail|
 *   d_xyz (B,N,3)  d_loads_nodes (B|1,N,3)  d_edges_load (B|1,E,3)  ->  d_loads_out (B,N,3)
 */
int fdm_edge_loads(const fdm_plan* plan, const double* d_xyz, const double* d_loads_nodes,
                   const double* d_edges_load, double* d_loads_out, int64_t batch,
                   int64_t loads_stride, int64_t edges_load_stride, int is_local, void* stream);

/* VJP of fdm_edge_loads at the cotangent d_cot (B,N,3) of loads_out.  Outputs are OVERWRITTEN and
 * any of them may be NULL:  xyz_bar (B,N,3) [through the edge lengths; a zero-length edge gives
 * NaN like jnp.linalg.norm], edges_load_bar (B,E,3);  loads_nodes_bar is d_cot itself. */
int fdm_edge_loads_vjp(const fdm_plan* plan, const double* d_xyz, const double* d_edges_load,
                       const double* d_cot, double* d_xyz_bar, double* d_edges_load_bar,
                       int64_t batch, int64_t edges_load_stride, int is_local, void* stream);

/* Tributary FACE (area) loads (models.py:333-344, loads.py:35-71, 80-185, 187-288): every edge takes,
 * from each of its (first two) incident faces, the face load times the area of the triangle (edge, face
 * centroid); the edge total goes half to each end node.  is_local != 0: the face load is given in the
 * face's local frame (geometry.py:621-657 polygon_lcs; identity for a degenerate normal) and follows the
 * deformed geometry.  The mesh arrays are device int32 operands built by the host mirror of
 * EquilibriumMeshStructure (structures/meshes.py:108-204):
 *   d_faces (F,D) vertex indices padded -1 (D <= 8);  d_edges_faces (E,2) padded -1;
 *   d_faces_edges (F,D): edge on side i of the face, -1 when the face is not among that edge's two;
 *   d_node_faces_ptr (N+1), d_node_faces (sum deg): faces around each node.
 * Caller-provided scratch: d_centroids (B,F,3), d_gl (B,F,3) (face loads in global coordinates).
 * loads_out_n = loads_in_n + tributary face loads. */
int fdm_face_loads(const fdm_plan* plan, const int32_t* d_faces, int64_t num_faces, int64_t max_deg,
                   const int32_t* d_edges_faces, const double* d_xyz, const double* d_loads_in,
                   const double* d_faces_load, double* d_centroids, double* d_gl, double* d_loads_out,
                   int64_t batch, int64_t loads_stride, int64_t faces_load_stride, int is_local, void* stream);

/* VJP of fdm_face_loads at the cotangent d_cot (B,N,3).  Outputs OVERWRITTEN, any may be NULL:
 * xyz_bar (B,N,3) [through areas, centroids and, for local loads, the face frames], faces_load_bar (B,F,3).
 * d_centroids / d_gl as filled by the forward call; scratch: d_cen_bar (B,F,3), d_vert_bar (B,F,D,3)
 * (local loads only). */
int fdm_face_loads_vjp(const fdm_plan* plan, const int32_t* d_faces, int64_t num_faces, int64_t max_deg,
                       const int32_t* d_edges_faces, const int32_t* d_faces_edges,
                       const int32_t* d_node_faces_ptr, const int32_t* d_node_faces, const double* d_xyz,
                       const double* d_faces_load, const double* d_centroids, const double* d_gl,
                       const double* d_cot, double* d_cen_bar, double* d_vert_bar, double* d_xyz_bar,
                       double* d_faces_load_bar, int64_t batch, int64_t faces_load_stride, int is_local,
                       void* stream);

/* ---- goals and losses fused after the state --------------------------------
 * Replaces Loss.__call__ (losses/loss.py:60-90) / Error.__call__ (losses/errors.py:100-135) for goals
 * whose prediction reads one element of the equilibrium state, and the reverse pass JAX derives:
 *   kind 0 NodePointGoal (goals/node/point.py)            xyz[i]            vs target[3]
 *        1 NodeResidualForceGoal (goals/node/residual.py) |residuals[i]|
 *        2 NodeResidualVectorGoal                         residuals[i]      vs target[3]
 *        3 EdgeLengthGoal (goals/edge/length.py)          lengths[e]
 *        4 EdgeForceGoal (goals/edge/force.py)            forces[e]
 *        5 EdgeLoadPathGoal (goals/edge/loadpath.py)      |forces[e]| lengths[e]
 *        6 NetworkLoadPathGoal (goals/network/loadpath.py) sum_e |forces[e] lengths[e]|   (index unused)
 *        7, 8, 9 Node{X,Y,Z}CoordinateGoal (goals/node/coordinates.py)  xyz[i, 0 | 1 | 2]
 *        10 NodeLineGoal, 11 NodePlaneGoal, 12 NodeSegmentGoal (goals/node/line.py, plane.py, segment.py):
 *           xyz[i] vs its projection on the line / plane / segment given by target[0:6] (two points; origin and
 *           normal); the projection is differentiated through, as JAX does
 *        13 NodeResidualDirectionGoal, 14 NodeResidualPlaneGoal (goals/node/residual.py:86-219): the unit
 *           residual against the unit target / against its projection on the plane with normal target[0:3]
 *        15 EdgeDirectionGoal (goals/edge/direction.py): unit edge vector xyz[head]-xyz[tail] vs unit target
 *        16 EdgeAngleGoal (goals/edge/angle.py): angle between the edge vector and target[1:4] vs target[0]
 *        17 EdgesLengthEqualGoal (goals/edge/length.py:49-75): var(lengths[S]) / mean(lengths[S]) over the edge
 *           set S = agg_index[agg_ptr[t] .. agg_ptr[t+1]); 18 EdgesForceEqualGoal (goals/edge/force.py:49-75):
 *           var(forces[S]) / mean(|forces[S]|); 19 NodesColinearGoal (goals/node/colinear.py:19-60): colinearity
 *           energy of the ordered node sequence S; index[t] = running number of the aggregate goal (< 16)
 *   error 0 SquaredError w (p-t)^2, 1 AbsoluteError w |p-t|, 2 PredictionError w p,
 *         3 LogMaxError w log1p(max(p-t, 0))                                   (errors.py:232-431)
 * Terms are grouped by Error instance (terms of group g = [group_ptr[g], group_ptr[g+1])):
 *   loss = sum_g group_outer[g] * f_g(group_inner[g] * S_g),  S_g = sum of the group's term errors,
 *   f = identity (group_final 0) or sqrt (1): alpha, the 1/number_of_goals of the Mean* errors and the root
 *   of RootMeanSquaredError (errors.py:262-300).
 * All arrays of the program are DEVICE pointers shared by the batch. */
typedef struct fdm_goal_program {
  int32_t num_terms, num_groups, has_network_loadpath, num_aggregates;
  const int32_t* kind;        /* (T) */
  const int32_t* index;       /* (T) node or edge index (structure order) */
  const int32_t* error;       /* (T) */
  const double* weight;       /* (T) */
  const double* target;       /* (T,6) scalar targets in column 0, points in columns 0..2 */
  const int32_t* group_ptr;   /* (G+1) */
  const int32_t* group_final; /* (G) */
  const double* group_inner;  /* (G) */
  const double* group_outer;  /* (G) */
  const int32_t* agg_ptr;     /* (T+1) element sets of the aggregate goals (empty ranges for the others), or NULL */
  const int32_t* agg_index;   /* (sum of set sizes) */
} fdm_goal_program;

/* loss (B) [NULL in a reverse-only call]; group_sums (B,G) or NULL.  With cotangent arrays the kernel also writes loss_cot_b * d loss_b / d state: xyz_bar (B,N,3), lengths_bar (B,E), forces_bar
 * (B,E), residuals_bar (B,N,3) -- the cotangents fdm_state_vjp takes; d_loss_cot (B) is the upstream cotangent of
 * the loss (NULL = 1).
 * d_lengths / d_forces / d_residuals may be NULL when no goal of the program reads them (fdm_goal_kind_arrays), and a
 * cotangent array that no term of the program differentiates into (fdm_goal_kind_arrays) may be NULL --
 * it is then neither zeroed nor written, and fdm_state_vjp takes a NULL cotangent as zero; an array a term does read
 * must be given. */
int fdm_goals_loss(const fdm_plan* plan, const fdm_goal_program* program, const double* d_xyz,
                   const double* d_lengths, const double* d_forces, const double* d_residuals, double* d_loss,
                   double* d_group_sums, const double* d_loss_cot, double* d_xyz_bar, double* d_lengths_bar,
                   double* d_forces_bar, double* d_residuals_bar, int64_t batch, void* stream);
/* bit mask of the state arrays a goal kind differentiates into: 1 xyz, 2 lengths, 4 forces, 8 residuals */
int fdm_goal_kind_arrays(int kind);

/* ---- helper used by the benchmark's loss (above the path; kept here so the timed region
 * launches only this library's kernels): loss_b = 0.5 * sum (x - x0)^2, g = x - x0. */
int fdm_loss_sqdist(const double* d_x, const double* d_x0, double* d_g, double* d_loss,
                    int64_t n_per_design, int64_t batch, int64_t x0_stride, void* stream);

/* ---- the method-by-method route's transposes ------------------------------------------------------
 * Kbar.data[k] = sign * sum_xyz a[row_k] . b[col_k] on the plan's pattern: the cotangent of K that
 * sparse_solve_bwd hands back (sparse.py:299-306, a = lam, b = x, sign = -1) and the VJP of `K @ x`
 * (a = cotangent, b = x, sign = +1).   a, b (B,Nf,3) -> Kbar (B,nnz) */
int fdm_kbar_pattern(const fdm_plan* plan, const double* d_a, const double* d_b, double* d_Kbar_data, double sign,
                     int64_t batch, void* stream);
/* transpose of stiffness_matrix (models.py:1069-1077): q_bar (B,E) from Kbar.data (B,nnz); overwritten */
int fdm_stiffness_vjp(const fdm_plan* plan, const double* d_Kbar_data, double* d_q_bar, int64_t batch, void* stream);

/* out = a x + b y + c z over n doubles (y, z may be NULL; out may alias an input): the vector combinations the
 * host-side rules need between kernels (tangent right-hand sides b_dot - K_dot x, fixed-point updates) */
int fdm_lincomb(double* d_out, const double* d_x, const double* d_y, const double* d_z, double a, double b, double c,
                int64_t n, void* stream);

/* out[i] = sum_b in[b, i], fixed summation order: the local half of the two exchanges a sharded batch has (the
 * scalar loss, n = 1, and the gradient of a parameter shared by all designs), before the NCCL all-reduce. */
int fdm_batch_sum(const double* d_in /*(B,n)*/, double* d_out /*(n)*/, int64_t batch, int64_t n, void* stream);

/* ---- value and gradient of a loss over a batch, host in / host out ------------------------------
 * Replaces `jit(value_and_grad(loss_fn))(params)` (optimization/optimizers/optimizer.py:365-370) and the
 * vmap over designs of visualization/plotters/loss_plotter.py:98-103 for parameters = force densities:
 * one object that owns the device buffers, streams, events and CUDA graphs of the whole forward + adjoint
 * pass, so a step costs the host a handful of driver calls.  The batch is cut into `chunks` contiguous
 * slices; slice c copies its q in, runs (fdm_solve_from_q | fdm_assemble + fdm_solve) -> fdm_positions ->
 * fdm_state -> loss and cotangents -> fdm_solve with the stored factor -> fdm_adjoint_grads, and copies its
 * loss and q_bar out, all on its own stream, so copies and kernels of different slices overlap.
 * Loss: L_b = 1/2 |x_free - x0|^2 (default, fdm_loss_sqdist) or a goal program (fdm_fa_set_goal_program:
 * fdm_goals_loss + fdm_state_vjp + the transposes).  A design whose solve fails yields NaN loss / gradient. */
typedef struct fdm_fa fdm_fa;
typedef struct fdm_fa_opts {
  int32_t chunks;      /* slices of the batch (<= 0 -> 1)                                              */
  int32_t with_state;  /* also write vectors / lengths / forces / residuals (always on with a goal program) */
  int32_t solver;      /* FDM_SOLVER_*                                                                   */
  int32_t use_graph;   /* replay each slice (and the device-resident step) as a CUDA graph when capturable */
  int32_t max_iters;   /* PCG                                                                            */
  double tol;          /* PCG; <= 0 -> 1e-12                                                             */
} fdm_fa_opts;
int fdm_fa_create(const fdm_plan* plan, int64_t batch, const fdm_fa_opts* opts, fdm_fa** out);  /* allocates; synchronous */
void fdm_fa_destroy(fdm_fa* fa);
/* shared by the batch: xyz_fixed (Ns,3), loads (N,3), x0 (Nf,3) [may be NULL with a goal program]; synchronous */
int fdm_fa_set_constants(fdm_fa* fa, const double* h_xyz_fixed, const double* h_loads, const double* h_x0);
/* `h_program`: a fdm_goal_program whose arrays are HOST pointers; copied to the device; synchronous */
int fdm_fa_set_goal_program(fdm_fa* fa, const fdm_goal_program* h_program);
int fdm_fa_set_q(fdm_fa* fa, const double* h_q /*(B,E)*/);                       /* synchronous upload */
/* one forward + adjoint pass on the q resident on the device; asynchronous, ordered on `stream` (fork / join) */
int fdm_fa_step_device(fdm_fa* fa, void* stream);
/* end to end: h_q (B,E) in, h_loss (B) and h_q_bar (B,E) [may be NULL] out.  Pinned host memory
 * (fdm_host_alloc) makes the copies asynchronous.  enqueue + wait = run. */
int fdm_fa_enqueue_host(fdm_fa* fa, const double* h_q, double* h_loss, double* h_q_bar);
int fdm_fa_wait(fdm_fa* fa);
/* measurement only: the copies of fdm_fa_enqueue_host without the kernels (the link's ceiling for this call) */
int fdm_fa_enqueue_copies_only(fdm_fa* fa, const double* h_q, double* h_loss, double* h_q_bar);
int fdm_fa_run_host(fdm_fa* fa, const double* h_q, double* h_loss, double* h_q_bar);
enum {
  FDM_FA_Q = 0, FDM_FA_X_FREE = 1, FDM_FA_XYZ = 2, FDM_FA_LOSS = 3, FDM_FA_Q_BAR = 4, FDM_FA_LOADS_BAR = 5,
  FDM_FA_XYZ_FIXED_BAR = 6, FDM_FA_INFO_FORWARD = 7, FDM_FA_INFO_ADJOINT = 8, FDM_FA_LENGTHS = 9, FDM_FA_FORCES = 10,
  FDM_FA_RESIDUALS = 11, FDM_FA_VECTORS = 12, FDM_FA_XYZ_FIXED = 13, FDM_FA_LOADS = 14
};
/* device pointer of one of the object's arrays (valid until fdm_fa_destroy) and its doubles per design */
int fdm_fa_device_array(fdm_fa* fa, int which, void** d_out, int64_t* per_design);
enum { FDM_FA_INFO_CHUNKS = 0, FDM_FA_INFO_LAUNCHES = 1, FDM_FA_INFO_FUSED_ASSEMBLY = 2, FDM_FA_INFO_GRAPH = 3,
       FDM_FA_INFO_BATCH = 4 };
int64_t fdm_fa_info(const fdm_fa* fa, int which);   /* LAUNCHES: kernels per step (0 before the first step) */
/* ---- tmax > 1: the fixed point and its implicit adjoint as device-resident loops -------------------
 * Replaces equilibrium_iterative_xyz + solver_forward (models.py:512-618, solvers/fixed_point.py:139-198) and the
 * adjoint system of fixed_point_bwd_adjoint (fixed_point.py:501-604) for ONE design on the band Cholesky (half
 * bandwidth <= 31; else FDM_ERR_UNSUPPORTED and the caller iterates from the host):
 *     forward   x <- K^{-1} P(x),  P(x) = nodes_load(xyz(x))[free] - R_fixed,  until mean_n |x_prev - x|_2 <= eta
 *     adjoint   w <- x_bar + J_P^T K^{-1} w   until max |w_new - w| <= adjoint_tol * max |x_bar|;   lam = K^{-1} w
 * Each loop is one CUDA graph with a conditional WHILE node whose condition a kernel sets on the device: no host
 * synchronisation per iteration.  K is factored once per forward call and the factor serves every iteration of
 * both loops.  The object keeps the converged geometry between fdm_fp_forward and fdm_fp_adjoint.
 * Mesh arrays (face loads): device int32, as fdm_face_loads / fdm_face_loads_vjp take them; NULL without faces. */
typedef struct fdm_fp fdm_fp;
typedef struct fdm_fp_opts {
  int32_t tmax;               /* <= 0 -> 100 (models.py:77)                               */
  int32_t is_load_local;      /* edge / face loads given in the element frames            */
  int32_t adjoint_max_iters;  /* <= 0 -> 200                                              */
  double eta;                 /* <= 0 -> 1e-6                                             */
  double adjoint_tol;         /* <= 0 -> 1e-12                                            */
} fdm_fp_opts;
int fdm_fp_create(const fdm_plan* plan, const fdm_fp_opts* opts, int has_edges_load, const int32_t* d_faces,
                  int64_t num_faces, int64_t max_deg, const int32_t* d_edges_faces, const int32_t* d_faces_edges,
                  const int32_t* d_node_faces_ptr, const int32_t* d_node_faces, fdm_fp** out);   /* allocates; synchronous */
void fdm_fp_destroy(fdm_fp* fp);
/* d_x_init (Nf,3) or NULL (= the linear step with the nodal loads); d_x_free (Nf,3) out; d_loads_out (N,3) or NULL:
 * the nodal loads at the converged geometry.  Asynchronous on `stream`. */
int fdm_fp_forward(fdm_fp* fp, const double* d_q, const double* d_xyz_fixed, const double* d_loads_nodes,
                   const double* d_edges_load, const double* d_faces_load, const double* d_x_init, double* d_x_free,
                   double* d_loads_out, void* stream);
/* after fdm_fp_forward: d_x_bar (Nf,3) -> d_w (Nf,3, may be NULL), d_lam (Nf,3).  Asynchronous on `stream`. */
int fdm_fp_adjoint(fdm_fp* fp, const double* d_x_bar, double* d_w, double* d_lam, void* stream);
/* h_out5: forward steps (solver_forward's count), last forward distance, adjoint iterations, last adjoint change,
 * status of the last substitution.  Synchronises `stream`. */
int fdm_fp_info(fdm_fp* fp, double* h_out5, void* stream);

/* page-locked host memory for the end-to-end calls */
void* fdm_host_alloc(size_t bytes);
void fdm_host_free(void* p);

const char* fdm_last_error(void);
const char* fdm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FDM_B200_H */
