// fdm_fp: the tmax > 1 equilibrium (shape-dependent loads) and its implicit adjoint as DEVICE-RESIDENT loops.
//
// Replaces equilibrium_iterative_xyz + solver_forward (models.py:512-618, solvers/fixed_point.py:139-198) and the
// adjoint system of fixed_point_bwd_adjoint (fixed_point.py:501-604) for one design on the band Cholesky:
//     forward   x <- K^{-1} P(x),  P(x) = nodes_load(xyz(x))[free] - R_fixed,   until mean_n |x_prev - x|_2 <= eta
//     adjoint   w <- x_bar + J_P^T K^{-1} w                                     until max |w_new - w| <= tol max |x_bar|
// Each loop is ONE CUDA graph with a conditional WHILE node (cudaGraphCondTypeWhile): the body -- positions, tributary
// edge / face loads (or their transposes), the right-hand side, the substitution with the factor left by the first
// solve, the distance -- is captured once; a one-CTA kernel reduces the distance and sets the loop condition on the
// device (cudaGraphSetConditional).  No host synchronisation or host read per iteration; the step count and the last
// distance stay in device memory until somebody asks.  K is factored once per forward call; every iteration of both
// loops reuses that factor (the reference re-factorises per iteration, models.py:591).
#include <cuda_runtime.h>

#include <cstring>
#include <new>
#include <vector>

#include "launch.h"

extern "C" size_t fdm_workspace_bytes(const fdm_plan* p, int64_t batch, int32_t solver);

namespace {

// mean over nodes of |a_n - b_n|_2 (fixed_point.py:176-183) by one CTA; a <- b... (see callers)
// mode 0 (forward):  d = mean_n |x_new_n - x_n|_2; x <- x_new; evals += 1; continue while (evals - 1 < tmax && d > eta)
// mode 1 (adjoint):  d = max |w_new - w|;          w <- w_new; its   += 1; continue while (its < max_iters && d > tol * scale)
__global__ void __launch_bounds__(1024)
fp_check_kernel(int mode, const double* __restrict__ fresh, double* __restrict__ cur, int64_t nrows, double thresh,
                const double* __restrict__ scale, int limit, double* __restrict__ state /* [0] count, [1] distance */,
                cudaGraphConditionalHandle handle) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t r = threadIdx.x; r < nrows; r += blockDim.x) {
    const double dx = fresh[3 * r] - cur[3 * r], dy = fresh[3 * r + 1] - cur[3 * r + 1], dz = fresh[3 * r + 2] - cur[3 * r + 2];
    if (mode == 0) acc += sqrt(dx * dx + dy * dy + dz * dz);
    else acc = fmax(acc, fmax(fabs(dx), fmax(fabs(dy), fabs(dz))));
    cur[3 * r] = fresh[3 * r]; cur[3 * r + 1] = fresh[3 * r + 1]; cur[3 * r + 2] = fresh[3 * r + 2];
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, acc, o);
    acc = mode == 0 ? acc + other : fmax(acc, other);
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, acc, o);
      acc = mode == 0 ? acc + other : fmax(acc, other);
    }
    if (threadIdx.x == 0) {
      const double d = mode == 0 ? acc / (double)nrows : acc;
      const double count = state[0] + 1.0;
      state[0] = count;
      state[1] = d;
      bool go;
      if (mode == 0) go = (count - 1.0 < (double)limit) && (d > thresh);           // solver_forward's while condition
      else go = (count < (double)limit) && !(d <= thresh * fmax(scale[0], 1e-300));
      cudaGraphSetConditional(handle, go ? 1u : 0u);
    }
  }
}

__global__ void __launch_bounds__(1024) fp_absmax_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc = fmax(acc, fabs(v[i]));
  for (int o = 16; o > 0; o >>= 1) acc = fmax(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc = fmax(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if (threadIdx.x == 0) out[0] = acc;
  }
}

}  // namespace

struct fdm_fp {
  const fdm_plan* p = nullptr;
  int device = 0;          // kept here: the plan may be gone by the time the object is destroyed
  fdm_fp_opts o{};
  // mesh arrays (device, caller-owned) for face loads
  const int32_t *faces = nullptr, *edges_faces = nullptr, *faces_edges = nullptr, *nf_ptr = nullptr, *nf = nullptr;
  int64_t F = 0, D = 0;
  bool has_edges = false, has_faces = false;
  // own buffers
  std::vector<void*> allocs;
  double *q = nullptr, *xs = nullptr, *loads_nodes = nullptr, *edges_load = nullptr, *faces_load = nullptr;
  double *K = nullptr, *rhs = nullptr, *x = nullptr, *x_new = nullptr, *xyz = nullptr, *nl_a = nullptr, *nl_b = nullptr;
  double *cen = nullptr, *gl = nullptr, *info = nullptr, *zeros_ns = nullptr, *zero_edges = nullptr;
  double *vec = nullptr, *w = nullptr, *w_new = nullptr, *lam = nullptr, *cot = nullptr, *xb_a = nullptr, *xb_b = nullptr;
  double *cen_bar = nullptr, *vert_bar = nullptr, *xfree_bar = nullptr, *xs_tmp = nullptr;
  double *fstate = nullptr, *astate = nullptr, *scale = nullptr;     // device: [count, distance]
  void* ws = nullptr;
  size_t ws_bytes = 0;
  cudaStream_t cap = nullptr;
  cudaGraphExec_t fwd_exec = nullptr, adj_exec = nullptr;
  cudaGraph_t fwd_graph = nullptr, adj_graph = nullptr;
};

namespace {

int fp_alloc(fdm_fp* f, double** out, size_t count) {
  void* m = nullptr;
  const size_t bytes = std::max<size_t>(count, 2) * sizeof(double);
  FDM_CUDA(cudaMalloc(&m, bytes));
  FDM_CUDA(cudaMemset(m, 0, bytes));
  f->allocs.push_back(m);
  *out = (double*)m;
  return FDM_OK;
}

fdm_solve_opts solve_opts(const fdm_fp* f, int reuse) {
  fdm_solve_opts so;
  std::memset(&so, 0, sizeof(so));
  so.solver = FDM_SOLVER_CHOLESKY_BATCHED;
  so.reuse_factor = reuse;
  so.nan_on_failure = 1;
  return so;
}

// nodes_load(xyz(x)) -> rhs = loads[free] - R_fixed (models.py:290-344, 858-896), on stream st.  Returns the buffer
// that holds the nodal loads.
int launch_P(fdm_fp* f, const double* x, cudaStream_t st) {
  const fdm_plan* p = f->p;
  int rc = fdm_positions(p, const_cast<double*>(x), f->xs, f->xyz, 1, 0, 0, st);
  if (rc != FDM_OK) return rc;
  const double* loads = f->loads_nodes;
  if (f->has_edges) {
    rc = fdm_edge_loads(p, f->xyz, loads, f->edges_load, f->nl_a, 1, 0, 0, f->o.is_load_local, st);
    if (rc != FDM_OK) return rc;
    loads = f->nl_a;
  }
  if (f->has_faces) {
    rc = fdm_face_loads(p, f->faces, f->F, f->D, f->edges_faces, f->xyz, loads, f->faces_load, f->cen, f->gl, f->nl_b, 1, 0,
                        0, f->o.is_load_local, st);
    if (rc != FDM_OK) return rc;
    loads = f->nl_b;
  }
  return fdm_assemble(p, f->q, f->xs, loads, nullptr, f->rhs, 1, p->E, 0, 0, st);
}

// the WHILE graph: body captured on f->cap into the conditional node's body graph
template <class Body>
int build_while(fdm_fp* f, cudaGraph_t* graph_out, cudaGraphExec_t* exec_out, Body body) {
  cudaGraph_t g = nullptr;
  FDM_CUDA(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle handle;
  FDM_CUDA(cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = handle;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t node;
  FDM_CUDA(cudaGraphAddNode(&node, g, nullptr, 0, &np));
  cudaGraph_t bodyg = np.conditional.phGraph_out[0];
  FDM_CUDA(cudaStreamBeginCaptureToGraph(f->cap, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  int rc = body(f->cap, handle);
  cudaGraph_t tmp = nullptr;
  cudaError_t e = cudaStreamEndCapture(f->cap, &tmp);
  if (rc != FDM_OK) { cudaGraphDestroy(g); return rc; }
  if (e != cudaSuccess) {
    fdm_set_error(std::string("fdm_fp: capturing the loop body: ") + cudaGetErrorString(e));
    cudaGraphDestroy(g);
    return FDM_ERR_CUDA;
  }
  e = cudaGraphInstantiate(exec_out, g, 0);
  if (e != cudaSuccess) {
    fdm_set_error(std::string("fdm_fp: instantiating the loop graph: ") + cudaGetErrorString(e));
    cudaGraphDestroy(g);
    return FDM_ERR_CUDA;
  }
  *graph_out = g;
  return FDM_OK;
}

void fp_destroy(fdm_fp* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  if (f->cap) cudaStreamSynchronize(f->cap);
  if (f->fwd_exec) cudaGraphExecDestroy(f->fwd_exec);
  if (f->adj_exec) cudaGraphExecDestroy(f->adj_exec);
  if (f->fwd_graph) cudaGraphDestroy(f->fwd_graph);
  if (f->adj_graph) cudaGraphDestroy(f->adj_graph);
  if (f->cap) cudaStreamDestroy(f->cap);
  if (f->ws) cudaFree(f->ws);
  for (void* m : f->allocs) cudaFree(m);
  delete f;
}

}  // namespace

extern "C" {

int fdm_fp_create(const fdm_plan* p, const fdm_fp_opts* opts, int has_edges_load, const int32_t* d_faces, int64_t num_faces,
                  int64_t max_deg, const int32_t* d_edges_faces, const int32_t* d_faces_edges, const int32_t* d_node_faces_ptr,
                  const int32_t* d_node_faces, fdm_fp** out) {
  if (!out) return FDM_ERR_INVALID;
  *out = nullptr;
  if (!p || !p->on_device || !opts) {
    fdm_set_error("fdm_fp_create: plan is null or host-only, or no options");
    return FDM_ERR_INVALID;
  }
  if (!chol_warp_supported(p)) {
    fdm_set_error("fdm_fp_create: the device-resident loops run on the warp-per-design band Cholesky (half bandwidth <= 31); "
                  "wider bands iterate from the host");
    return FDM_ERR_UNSUPPORTED;
  }
  if (d_faces && (max_deg < 3 || max_deg > 8 || num_faces <= 0 || !d_edges_faces || !d_faces_edges || !d_node_faces_ptr || !d_node_faces)) {
    fdm_set_error("fdm_fp_create: incomplete mesh arrays");
    return FDM_ERR_INVALID;
  }
  FDM_CUDA(cudaSetDevice(p->device));
  fdm_fp* f = new (std::nothrow) fdm_fp();
  if (!f) return FDM_ERR_INVALID;
  f->p = p;
  f->device = p->device;
  f->o = *opts;
  if (f->o.tmax <= 0) f->o.tmax = 100;
  if (f->o.eta <= 0) f->o.eta = 1e-6;
  if (f->o.adjoint_max_iters <= 0) f->o.adjoint_max_iters = 200;
  if (f->o.adjoint_tol <= 0) f->o.adjoint_tol = 1e-12;
  f->has_edges = has_edges_load != 0;
  f->has_faces = d_faces != nullptr;
  f->faces = d_faces; f->F = num_faces; f->D = max_deg; f->edges_faces = d_edges_faces; f->faces_edges = d_faces_edges;
  f->nf_ptr = d_node_faces_ptr; f->nf = d_node_faces;
  const size_t E = (size_t)p->E, N = (size_t)p->N, Nf = (size_t)p->Nf, Ns = (size_t)p->Ns, F = (size_t)std::max<int64_t>(num_faces, 1);
  int rc = FDM_OK;
  auto A = [&](double** ptr, size_t n) { if (rc == FDM_OK) rc = fp_alloc(f, ptr, n); };
  A(&f->q, E); A(&f->xs, Ns * 3); A(&f->loads_nodes, N * 3); A(&f->edges_load, E * 3); A(&f->faces_load, F * 3);
  A(&f->K, (size_t)p->nnz); A(&f->rhs, Nf * 3); A(&f->x, Nf * 3); A(&f->x_new, Nf * 3); A(&f->xyz, N * 3);
  A(&f->nl_a, N * 3); A(&f->nl_b, N * 3); A(&f->cen, F * 3); A(&f->gl, F * 3); A(&f->info, FDM_INFO_STRIDE);
  A(&f->zeros_ns, Ns * 3); A(&f->vec, Nf * 3); A(&f->w, Nf * 3); A(&f->w_new, Nf * 3); A(&f->lam, Nf * 3);
  A(&f->cot, N * 3); A(&f->xb_a, N * 3); A(&f->xb_b, N * 3); A(&f->cen_bar, F * 3); A(&f->vert_bar, F * 8 * 3);
  A(&f->xfree_bar, Nf * 3); A(&f->xs_tmp, Ns * 3); A(&f->fstate, 2); A(&f->astate, 2); A(&f->scale, 2);
  if (rc == FDM_OK) {
    f->ws_bytes = std::max<size_t>(fdm_workspace_bytes(p, 1, FDM_SOLVER_CHOLESKY_BATCHED), 16);
    if (cudaMalloc(&f->ws, f->ws_bytes) != cudaSuccess) rc = FDM_ERR_CUDA;
  }
  if (rc == FDM_OK && cudaStreamCreateWithFlags(&f->cap, cudaStreamNonBlocking) != cudaSuccess) rc = FDM_ERR_CUDA;
  // warm the kernels once outside capture (function attributes are set on first use), then capture both loops
  fdm_solve_opts so0 = solve_opts(f, 0), so1 = solve_opts(f, 1);
  if (rc == FDM_OK) {
    // a harmless solve: K = I-like from q = 1 is not needed; attributes are set by launching on the zeroed buffers
    rc = launch_P(f, f->x, f->cap);
    if (rc == FDM_OK) rc = fdm_solve(p, f->K, f->rhs, f->x_new, f->info, f->ws, f->ws_bytes, 1, &so0, f->cap);
    if (rc == FDM_OK) rc = fdm_solve(p, f->K, f->rhs, f->x_new, f->info, f->ws, f->ws_bytes, 1, &so1, f->cap);
    if (rc == FDM_OK && cudaStreamSynchronize(f->cap) != cudaSuccess) rc = FDM_ERR_CUDA;
  }
  if (rc == FDM_OK) {
    rc = build_while(f, &f->fwd_graph, &f->fwd_exec, [&](cudaStream_t st, cudaGraphConditionalHandle h) {
      int r = launch_P(f, f->x, st);
      if (r != FDM_OK) return r;
      r = fdm_solve(p, f->K, f->rhs, f->x_new, f->info, f->ws, f->ws_bytes, 1, &so1, st);
      if (r != FDM_OK) return r;
      fp_check_kernel<<<1, 1024, 0, st>>>(0, f->x_new, f->x, p->Nf, f->o.eta, nullptr, f->o.tmax, f->fstate, h);
      return fdm_check_launch("fp_check");
    });
  }
  if (rc == FDM_OK) {
    rc = build_while(f, &f->adj_graph, &f->adj_exec, [&](cudaStream_t st, cudaGraphConditionalHandle h) {
      // w_new = vec + J_P^T lam  (lam = K^{-1} w already in f->lam), then lam = K^{-1} w_new
      int r = fdm_positions(p, f->lam, f->zeros_ns, f->cot, 1, 0, 0, st);          // lam on the free rows, 0 on supports
      if (r != FDM_OK) return r;
      const double* xb = nullptr;
      if (f->has_edges) {
        r = fdm_edge_loads_vjp(p, f->xyz, f->edges_load, f->cot, f->xb_a, nullptr, 1, 0, f->o.is_load_local, st);
        if (r != FDM_OK) return r;
        xb = f->xb_a;
      }
      if (f->has_faces) {
        r = fdm_face_loads_vjp(p, f->faces, f->F, f->D, f->edges_faces, f->faces_edges, f->nf_ptr, f->nf, f->xyz, f->faces_load,
                               f->cen, f->gl, f->cot, f->cen_bar, f->vert_bar, f->xb_b, nullptr, 1, 0, f->o.is_load_local, st);
        if (r != FDM_OK) return r;
        if (xb) {
          r = fdm_lincomb(f->xb_b, f->xb_a, f->xb_b, nullptr, 1.0, 1.0, 0.0, p->N * 3, st);
          if (r != FDM_OK) return r;
        }
        xb = f->xb_b;
      }
      if (xb) {
        r = fdm_positions(p, f->xfree_bar, f->xs_tmp, const_cast<double*>(xb), 1, p->Ns * 3, 1, st);   // free rows of J_P^T lam
        if (r != FDM_OK) return r;
        r = fdm_lincomb(f->w_new, f->vec, f->xfree_bar, nullptr, 1.0, 1.0, 0.0, p->Nf * 3, st);
      } else {
        r = fdm_lincomb(f->w_new, f->vec, nullptr, nullptr, 1.0, 0.0, 0.0, p->Nf * 3, st);
      }
      if (r != FDM_OK) return r;
      r = fdm_solve(p, f->K, f->w_new, f->lam, f->info, f->ws, f->ws_bytes, 1, &so1, st);
      if (r != FDM_OK) return r;
      fp_check_kernel<<<1, 1024, 0, st>>>(1, f->w_new, f->w, p->Nf, f->o.adjoint_tol, f->scale, f->o.adjoint_max_iters, f->astate, h);
      return fdm_check_launch("fp_check");
    });
  }
  if (rc != FDM_OK) {
    if (rc == FDM_ERR_CUDA && std::string(fdm_last_error()).empty())
      fdm_set_error(std::string("fdm_fp_create: ") + cudaGetErrorString(cudaGetLastError()));
    fp_destroy(f);
    return rc;
  }
  *out = f;
  return FDM_OK;
}

void fdm_fp_destroy(fdm_fp* f) { fp_destroy(f); }

int fdm_fp_forward(fdm_fp* f, const double* d_q, const double* d_xyz_fixed, const double* d_loads_nodes,
                   const double* d_edges_load, const double* d_faces_load, const double* d_x_init, double* d_x_free,
                   double* d_loads_out, void* stream) {
  if (!f || !d_q || !d_xyz_fixed || !d_loads_nodes || !d_x_free || (f->has_edges && !d_edges_load) ||
      (f->has_faces && !d_faces_load)) {
    fdm_set_error("fdm_fp_forward: bad arguments");
    return FDM_ERR_INVALID;
  }
  const fdm_plan* p = f->p;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t E = (size_t)p->E, N = (size_t)p->N, Nf = (size_t)p->Nf, Ns = (size_t)p->Ns;
  auto copy = [&](double* dst, const double* src, size_t n) {
    return cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyDeviceToDevice, st) == cudaSuccess ? FDM_OK : FDM_ERR_CUDA;
  };
  int rc = copy(f->q, d_q, E);
  if (rc == FDM_OK) rc = copy(f->xs, d_xyz_fixed, Ns * 3);
  if (rc == FDM_OK) rc = copy(f->loads_nodes, d_loads_nodes, N * 3);
  if (rc == FDM_OK && f->has_edges) rc = copy(f->edges_load, d_edges_load, E * 3);
  if (rc == FDM_OK && f->has_faces) rc = copy(f->faces_load, d_faces_load, (size_t)f->F * 3);
  if (rc != FDM_OK) { fdm_set_error("fdm_fp_forward: input copy failed"); return rc; }
  // K once; the first solve factors it.  x_init given: start the iteration there (models.py:556-558), else from
  // the linear step with the nodal loads alone (models.py:449-452).
  rc = fdm_assemble(p, f->q, f->xs, f->loads_nodes, f->K, f->rhs, 1, p->E, 0, 0, st);
  if (rc != FDM_OK) return rc;
  fdm_solve_opts so0 = solve_opts(f, 0);
  rc = fdm_solve(p, f->K, f->rhs, f->x, f->info, f->ws, f->ws_bytes, 1, &so0, st);
  if (rc != FDM_OK) return rc;
  if (d_x_init) {
    rc = copy(f->x, d_x_init, Nf * 3);
    if (rc != FDM_OK) return rc;
  }
  FDM_CUDA(cudaMemsetAsync(f->fstate, 0, 16, st));
  FDM_CUDA(cudaGraphLaunch(f->fwd_exec, st));
  rc = copy(d_x_free, f->x, Nf * 3);
  if (rc != FDM_OK) return rc;
  // the converged geometry and its loads stay in the object for the adjoint; the nodal loads at x* on request
  rc = launch_P(f, f->x, st);
  if (rc != FDM_OK) return rc;
  if (d_loads_out) {
    const double* loads = f->has_faces ? f->nl_b : f->has_edges ? f->nl_a : f->loads_nodes;
    rc = copy(d_loads_out, loads, N * 3);
  }
  return rc;
}

int fdm_fp_adjoint(fdm_fp* f, const double* d_x_bar, double* d_w, double* d_lam, void* stream) {
  if (!f || !d_x_bar || !d_lam) {
    fdm_set_error("fdm_fp_adjoint: bad arguments");
    return FDM_ERR_INVALID;
  }
  const fdm_plan* p = f->p;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t Nf = (size_t)p->Nf;
  FDM_CUDA(cudaMemcpyAsync(f->vec, d_x_bar, Nf * 24, cudaMemcpyDeviceToDevice, st));
  FDM_CUDA(cudaMemcpyAsync(f->w, d_x_bar, Nf * 24, cudaMemcpyDeviceToDevice, st));
  fp_absmax_kernel<<<1, 1024, 0, st>>>(f->vec, (int64_t)Nf * 3, f->scale);
  fdm_solve_opts so1 = solve_opts(f, 1);
  int rc = fdm_solve(p, f->K, f->w, f->lam, f->info, f->ws, f->ws_bytes, 1, &so1, st);   // lam = K^{-1} x_bar
  if (rc != FDM_OK) return rc;
  FDM_CUDA(cudaMemsetAsync(f->astate, 0, 16, st));
  FDM_CUDA(cudaGraphLaunch(f->adj_exec, st));
  if (d_w) FDM_CUDA(cudaMemcpyAsync(d_w, f->w, Nf * 24, cudaMemcpyDeviceToDevice, st));
  FDM_CUDA(cudaMemcpyAsync(d_lam, f->lam, Nf * 24, cudaMemcpyDeviceToDevice, st));
  return fdm_check_launch("fdm_fp_adjoint");
}

/* which: 0 forward steps (solver_forward's count), 1 last forward distance, 2 adjoint iterations, 3 last adjoint
 * change, 4 solve status of the last substitution; one small device-to-host copy, synchronises `stream` */
int fdm_fp_info(fdm_fp* f, double* h_out5, void* stream) {
  if (!f || !h_out5) return FDM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  double a[2], b[2], inf[FDM_INFO_STRIDE];
  FDM_CUDA(cudaMemcpyAsync(a, f->fstate, 16, cudaMemcpyDeviceToHost, st));
  FDM_CUDA(cudaMemcpyAsync(b, f->astate, 16, cudaMemcpyDeviceToHost, st));
  FDM_CUDA(cudaMemcpyAsync(inf, f->info, sizeof(inf), cudaMemcpyDeviceToHost, st));
  FDM_CUDA(cudaStreamSynchronize(st));
  h_out5[0] = a[0] > 0 ? a[0] - 1.0 : 0.0;
  h_out5[1] = a[1];
  h_out5[2] = b[0];
  h_out5[3] = b[1];
  h_out5[4] = inf[FDM_INFO_STATUS];
  return FDM_OK;
}

}  // extern "C"
