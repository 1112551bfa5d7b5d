// C ABI (include/fdm_b200.h): plan lifetime, array getters, argument checks, dispatch.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <new>

#include "launch.h"

namespace {

template <class T>
int upload(const std::vector<T>& h, T** d, size_t slack_bytes = 16) {
  *d = nullptr;
  const size_t bytes = h.size() * sizeof(T);
  FDM_CUDA(cudaMalloc((void**)d, bytes + slack_bytes));
  if (bytes) FDM_CUDA(cudaMemcpy(*d, h.data(), bytes, cudaMemcpyHostToDevice));
  FDM_CUDA(cudaMemset((char*)*d + bytes, 0, slack_bytes));
  return FDM_OK;
}

void free_device(fdm_plan* p) {
  DeviceArrays& d = p->d;
  void* ptrs[] = {d.node_slot, d.edge_nodes, d.ninc_ptr, d.ninc_edge, d.ninc_other, d.free_node,
                  d.fixed_node, d.rowptr, d.colidx, d.entry_edge, d.diag_pos, d.band_perm,
                  d.band_iperm, d.band_kidx, d.band_rowkidx, d.diags_ptr, d.diags_idx, d.sup_mask,
                  d.band_stage_ptr, d.band_stage_kent, d.band_stage_eent, d.band_row_diag, d.band_row_sup, d.band_row_node,
                  d.asm_code, d.rsup_ptr, d.rsup_rows, d.rsup_edge, d.rsup_slot, d.rhs_src};
  for (void* q : ptrs)
    if (q) cudaFree(q);
  for (int h = 0; h < 2; ++h) {
    void* tw[] = {d.tw_iperm[h], d.tw_stage_ptr[h], d.tw_stage_kent[h], d.tw_stage_eent[h], d.tw_row_diag[h],
                  d.tw_row_sup[h], d.tw_row_node[h]};
    for (void* q : tw)
      if (q) cudaFree(q);
  }
  d = DeviceArrays();
}

int build_host(fdm_plan* p, const int64_t* edges, int64_t E, int64_t N, const int64_t* supports,
               int num_parts) {
  int rc = fdm_build_topology(*p, edges, E, N, supports);
  if (rc != FDM_OK) return rc;
  fdm_build_band_order(*p);
  fdm_build_twisted(*p);
  fdm_check_assemble_batch(*p);
  fdm_build_partition(*p, num_parts);
  return FDM_OK;
}

template <class T>
int64_t copy_out(const std::vector<T>& v, void* out, int64_t cap) {
  const int64_t bytes = (int64_t)(v.size() * sizeof(T));
  if (!out || cap < bytes) {
    fdm_set_error("fdm_plan_get_array: buffer too small (need " + std::to_string(bytes) + " bytes)");
    return FDM_ERR_INVALID;
  }
  if (bytes) std::memcpy(out, v.data(), (size_t)bytes);
  return bytes;
}

inline bool ready(const fdm_plan* p, const char* who) {
  if (!p || !p->on_device) {
    fdm_set_error(std::string(who) + ": plan is null or host-only");
    return false;
  }
  return true;
}

}  // namespace

extern "C" {

const char* fdm_version(void) { return "fdm_b200 0.1 (sm_100a)"; }

int fdm_plan_create_host(const int64_t* h_edges, int64_t num_edges, int64_t num_nodes,
                         const int64_t* h_supports, fdm_plan** out) {
  if (!out) return FDM_ERR_INVALID;
  *out = nullptr;
  fdm_plan* p = new (std::nothrow) fdm_plan();
  if (!p) return FDM_ERR_INVALID;
  int rc = build_host(p, h_edges, num_edges, num_nodes, h_supports, 148);
  if (rc != FDM_OK) {
    delete p;
    return rc;
  }
  *out = p;
  return FDM_OK;
}

int fdm_plan_create(const int64_t* h_edges, int64_t num_edges, int64_t num_nodes,
                    const int64_t* h_supports, int device, fdm_plan** out) {
  if (!out) return FDM_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  FDM_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) {
    fdm_set_error("fdm_plan_create: no such CUDA device " + std::to_string(device));
    return FDM_ERR_CUDA;
  }
  FDM_CUDA(cudaSetDevice(device));
  int sms = 0;
  FDM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  fdm_plan* p = new (std::nothrow) fdm_plan();
  if (!p) return FDM_ERR_INVALID;
  p->device = device;
  p->sm_count = sms;
  int parts = sms;
  if (const char* e = std::getenv("FDM_PCG2_PARTS")) parts = std::max(2, std::min(sms, std::atoi(e)));
  int rc = build_host(p, h_edges, num_edges, num_nodes, h_supports, parts);
  if (rc == FDM_OK) {
    DeviceArrays& d = p->d;
    int (*steps[])(fdm_plan*) = {
        [](fdm_plan* q) { return upload(q->node_slot, &q->d.node_slot); },
        [](fdm_plan* q) { return upload(q->edge_nodes, &q->d.edge_nodes); },
        [](fdm_plan* q) { return upload(q->ninc_ptr, &q->d.ninc_ptr); },
        [](fdm_plan* q) { return upload(q->ninc_edge, &q->d.ninc_edge); },
        [](fdm_plan* q) { return upload(q->ninc_other, &q->d.ninc_other); },
        [](fdm_plan* q) { return upload(q->free_node, &q->d.free_node); },
        [](fdm_plan* q) { return upload(q->fixed_node, &q->d.fixed_node); },
        [](fdm_plan* q) { return upload(q->index_indptr, &q->d.rowptr); },
        [](fdm_plan* q) { return upload(q->index_indices, &q->d.colidx); },
        [](fdm_plan* q) { return upload(q->entry_edge, &q->d.entry_edge); },
        [](fdm_plan* q) { return upload(q->diag_pos, &q->d.diag_pos); },
        [](fdm_plan* q) { return upload(q->band_perm, &q->d.band_perm); },
        [](fdm_plan* q) { return upload(q->band_iperm, &q->d.band_iperm); },
        [](fdm_plan* q) { return upload(q->band_kidx, &q->d.band_kidx); },
        [](fdm_plan* q) { return upload(q->band_rowkidx, &q->d.band_rowkidx); },
        [](fdm_plan* q) { return upload(q->band_stage_ptr, &q->d.band_stage_ptr); },
        [](fdm_plan* q) { return upload(q->band_stage_kent, &q->d.band_stage_kent); },
        [](fdm_plan* q) { return upload(q->band_stage_eent, &q->d.band_stage_eent); },
        [](fdm_plan* q) { return upload(q->band_row_diag, &q->d.band_row_diag); },
        [](fdm_plan* q) { return upload(q->band_row_sup, &q->d.band_row_sup); },
        [](fdm_plan* q) { return upload(q->band_row_node, &q->d.band_row_node); },
        [](fdm_plan* q) { return upload(q->diags_indptr, &q->d.diags_ptr); },
        [](fdm_plan* q) { return upload(q->diags_indices, &q->d.diags_idx); },
        [](fdm_plan* q) { return upload(q->sup_mask, &q->d.sup_mask); },
        [](fdm_plan* q) { return upload(q->asm_code, &q->d.asm_code); },
        [](fdm_plan* q) { return upload(q->rsup_ptr, &q->d.rsup_ptr); },
        [](fdm_plan* q) { return upload(q->rsup_rows, &q->d.rsup_rows); },
        [](fdm_plan* q) { return upload(q->rhs_src, &q->d.rhs_src); },
        [](fdm_plan* q) { return upload(q->rsup_edge, &q->d.rsup_edge); },
        [](fdm_plan* q) { return upload(q->rsup_slot, &q->d.rsup_slot); },
    };
    for (auto f : steps) {
      rc = f(p);
      if (rc != FDM_OK) break;
    }
    for (int h = 0; h < 2 && rc == FDM_OK && p->tw[0].nblk > 0; ++h) {
      const BandHalf& H = p->tw[h];
      if (rc == FDM_OK) rc = upload(H.iperm, &d.tw_iperm[h]);
      if (rc == FDM_OK) rc = upload(H.stage_ptr, &d.tw_stage_ptr[h]);
      if (rc == FDM_OK) rc = upload(H.stage_kent, &d.tw_stage_kent[h]);
      if (rc == FDM_OK) rc = upload(H.stage_eent, &d.tw_stage_eent[h]);
      if (rc == FDM_OK) rc = upload(H.row_diag, &d.tw_row_diag[h]);
      if (rc == FDM_OK) rc = upload(H.row_sup, &d.tw_row_sup[h]);
      if (rc == FDM_OK) rc = upload(H.row_node, &d.tw_row_node[h]);
    }
  }
  if (rc != FDM_OK) {
    free_device(p);
    delete p;
    return rc;
  }
  p->on_device = true;
  // the persistent PCG's per-aggregate tables are built and uploaded HERE (plan_create is the only call that
  // allocates or synchronises); pcg2_supported() below is then a pure query, safe from any thread
  pcg2_prepare(p);
  *out = p;
  return FDM_OK;
}

void fdm_plan_destroy(fdm_plan* plan) {
  if (!plan) return;
  if (plan->on_device) {
    cudaSetDevice(plan->device);
    pcg2_destroy(plan);
    free_device(plan);
  }
  delete plan;
}

int64_t fdm_plan_size(const fdm_plan* p, int which) {
  if (!p) return FDM_ERR_INVALID;
  switch (which) {
    case FDM_SIZE_NODES: return p->N;
    case FDM_SIZE_EDGES: return p->E;
    case FDM_SIZE_FREE: return p->Nf;
    case FDM_SIZE_FIXED: return p->Ns;
    case FDM_SIZE_NNZ: return p->nnz;
    case FDM_SIZE_DIAGS_NNZ: return p->dnnz;
    case FDM_SIZE_HALF_BANDWIDTH: return p->hbw;
    case FDM_SIZE_NUM_PARTS: return p->num_parts;
  }
  return FDM_ERR_INVALID;
}

int64_t fdm_plan_get_array(const fdm_plan* p, int which, void* h_out, int64_t capacity) {
  if (!p) return FDM_ERR_INVALID;
  switch (which) {
    case FDM_ARR_INDICES_FREE: return copy_out(p->indices_free, h_out, capacity);
    case FDM_ARR_INDICES_FIXED: return copy_out(p->indices_fixed, h_out, capacity);
    case FDM_ARR_INDICES_FREEFIXED: return copy_out(p->indices_freefixed, h_out, capacity);
    case FDM_ARR_INDEX_DATA: return copy_out(p->index_data, h_out, capacity);
    case FDM_ARR_INDEX_INDICES: return copy_out(p->index_indices, h_out, capacity);
    case FDM_ARR_INDEX_INDPTR: return copy_out(p->index_indptr, h_out, capacity);
    case FDM_ARR_DIAG_INDICES: return copy_out(p->diag_indices, h_out, capacity);
    case FDM_ARR_DIAGS_INDPTR: return copy_out(p->diags_indptr, h_out, capacity);
    case FDM_ARR_DIAGS_INDICES: return copy_out(p->diags_indices, h_out, capacity);
    case FDM_ARR_BAND_PERM: return copy_out(p->band_perm, h_out, capacity);
    case FDM_ARR_PART_OF_ROW: return copy_out(p->part_of_row, h_out, capacity);
    case FDM_ARR_TW_IPERM: case FDM_ARR_TW_IPERM + 1: return copy_out(p->tw[which - FDM_ARR_TW_IPERM].iperm, h_out, capacity);
    case FDM_ARR_TW_STAGE_PTR: case FDM_ARR_TW_STAGE_PTR + 1:
      return copy_out(p->tw[which - FDM_ARR_TW_STAGE_PTR].stage_ptr, h_out, capacity);
    case FDM_ARR_TW_STAGE_KENT: case FDM_ARR_TW_STAGE_KENT + 1:
      return copy_out(p->tw[which - FDM_ARR_TW_STAGE_KENT].stage_kent, h_out, capacity);
    case FDM_ARR_TW_STAGE_EENT: case FDM_ARR_TW_STAGE_EENT + 1:
      return copy_out(p->tw[which - FDM_ARR_TW_STAGE_EENT].stage_eent, h_out, capacity);
  }
  fdm_set_error("fdm_plan_get_array: unknown array id");
  return FDM_ERR_INVALID;
}

int fdm_assemble(const fdm_plan* p, const double* q, const double* xs, const double* loads,
                 double* Kdata, double* rhs, int64_t batch, int64_t qs, int64_t xss, int64_t ls,
                 void* stream) {
  if (!ready(p, "fdm_assemble")) return FDM_ERR_INVALID;
  if (!q || batch <= 0 || (rhs && (!xs || !loads)) || (!Kdata && !rhs)) {
    fdm_set_error("fdm_assemble: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_assemble(p, q, xs, loads, Kdata, rhs, batch, qs, xss, ls, (cudaStream_t)stream);
}

static int pick_solver(const fdm_plan* p, int64_t batch, int32_t want) {
  if (want != FDM_SOLVER_AUTO) return want;
  // narrow bands (one warp per design): direct, deterministic, and the only solver here that also handles
  // mixed-sign q -- for one design as well as for a batch
  if (chol_warp_supported(p)) return FDM_SOLVER_CHOLESKY_BATCHED;
  if (chol_supported(p) && (batch > 1 || !pcg2_supported(p))) return FDM_SOLVER_CHOLESKY_BATCHED;
  if (pcg2_supported(p)) return FDM_SOLVER_PCG_PERSISTENT;
  return FDM_SOLVER_PCG;
}

static size_t primary_workspace_bytes(const fdm_plan* p, int64_t batch, int32_t picked) {
  switch (picked) {
    case FDM_SOLVER_PCG: return pcg_workspace_bytes(p);
    case FDM_SOLVER_PCG_PERSISTENT: return pcg2_workspace_bytes(p);
    case FDM_SOLVER_CHOLESKY_BATCHED: return chol_workspace_bytes(p, batch);
  }
  return 0;
}

// Does a design this solver reports as indefinite go on to the band L D L^T?  The warp-per-design Cholesky handles
// mixed-sign q itself up to half bandwidth 30; everything else with a band table does.
static bool ldlt_fallback(const fdm_plan* p, int32_t picked) {
  if (!ldlt_supported(p) || picked == FDM_SOLVER_LDLT_BAND || ldlt_fallback_slots(p, 1) < 1) return false;
  if (picked == FDM_SOLVER_CHOLESKY_BATCHED && chol_warp_supported(p) && p->hbw <= 30) return false;
  static const bool off = std::getenv("FDM_LDLT_FALLBACK") && std::atoi(std::getenv("FDM_LDLT_FALLBACK")) == 0;
  return !off;
}

static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

size_t fdm_workspace_bytes(const fdm_plan* p, int64_t batch, int32_t solver) {
  if (!p || batch <= 0) return 0;
  const int32_t picked = pick_solver(p, batch, solver);
  if (picked == FDM_SOLVER_LDLT_BAND) return ldlt_workspace_bytes(p, batch, batch);
  size_t bytes = primary_workspace_bytes(p, batch, picked);
  if (ldlt_fallback(p, picked)) bytes = align256(bytes) + ldlt_workspace_bytes(p, batch, ldlt_fallback_slots(p, batch));
  return bytes;
}

int fdm_solve(const fdm_plan* p, const double* Kdata, const double* rhs, double* x, double* info,
              void* ws, size_t ws_bytes, int64_t batch, const fdm_solve_opts* opts, void* stream) {
  if (!ready(p, "fdm_solve")) return FDM_ERR_INVALID;
  if (!Kdata || !rhs || !x || batch <= 0 || (!ws && ws_bytes)) {
    fdm_set_error("fdm_solve: bad arguments");
    return FDM_ERR_INVALID;
  }
  fdm_solve_opts o;
  std::memset(&o, 0, sizeof(o));
  if (opts) o = *opts;
  cudaStream_t st = (cudaStream_t)stream;
  if (o.nan_on_failure && !info) {
    fdm_set_error("fdm_solve: nan_on_failure needs d_info");
    return FDM_ERR_INVALID;
  }
  int rc;
  const int32_t picked = pick_solver(p, batch, o.solver);
  bool nan_done = false;
  switch (picked) {
    case FDM_SOLVER_PCG: rc = pcg_solve(p, Kdata, rhs, x, info, ws, ws_bytes, batch, o, st); break;
    case FDM_SOLVER_PCG_PERSISTENT: rc = pcg2_solve(p, Kdata, rhs, x, info, ws, ws_bytes, batch, o, st); break;
    case FDM_SOLVER_CHOLESKY_BATCHED:
      rc = chol_solve(p, Kdata, rhs, x, info, ws, ws_bytes, batch, o, st);
      nan_done = chol_warp_supported(p);                 // its substitution kernel writes the NaNs itself
      break;
    case FDM_SOLVER_LDLT_BAND:
      if (o.reuse_factor == 2) {
        fdm_set_error("fdm_solve: the band L D L^T has no shared-factor mode");
        return FDM_ERR_UNSUPPORTED;
      }
      rc = ldlt_solve(p, Kdata, rhs, x, info, ws, ws_bytes, batch, batch, 0, o.reuse_factor != 0, st);
      break;
    default:
      fdm_set_error("fdm_solve: unknown solver id");
      return FDM_ERR_INVALID;
  }
  // mixed-sign q behind a solver that needs a definite K: the designs it flagged go through the band L D L^T (the
  // kernel returns at once for all others), when the caller's workspace has the room fdm_workspace_bytes asked for
  if (rc == FDM_OK && info && ldlt_fallback(p, picked) && o.reuse_factor != 2) {
    const size_t off = align256(primary_workspace_bytes(p, batch, picked));
    const int64_t slots = ldlt_fallback_slots(p, batch);
    if (ws_bytes >= off + ldlt_workspace_bytes(p, batch, slots)) {
      rc = ldlt_solve(p, Kdata, rhs, x, info, static_cast<char*>(ws) + off, ws_bytes - off, batch, slots,
                      picked == FDM_SOLVER_CHOLESKY_BATCHED ? 1 : 2, o.reuse_factor != 0, st);
      nan_done = false;                                  // a design the fallback solved must lose its NaNs
      if (rc == FDM_OK && o.nan_on_failure && chol_warp_supported(p) && picked == FDM_SOLVER_CHOLESKY_BATCHED) {
        // (the warp kernels wrote NaN for the designs they gave up on before the fallback ran: it has overwritten x for
        //  the ones it solved, the others keep their NaN)
        nan_done = true;
      }
    }
  }
  if (rc == FDM_OK && o.nan_on_failure && !nan_done) rc = launch_nan_failed(info, x, p->Nf * 3, batch, st);
  return rc;
}

int fdm_solve_from_q(const fdm_plan* p, const double* q, const double* xs, const double* loads, double* x,
                     double* info, void* ws, size_t ws_bytes, int64_t batch, int64_t qs, int64_t xss, int64_t ls,
                     const fdm_solve_opts* opts, void* stream) {
  if (!ready(p, "fdm_solve_from_q")) return FDM_ERR_INVALID;
  if (!q || !xs || !loads || !x || batch <= 0 || (!ws && ws_bytes)) {
    fdm_set_error("fdm_solve_from_q: bad arguments");
    return FDM_ERR_INVALID;
  }
  fdm_solve_opts o;
  std::memset(&o, 0, sizeof(o));
  if (opts) o = *opts;
  if (o.reuse_factor) {
    fdm_set_error("fdm_solve_from_q: reuse_factor needs a right-hand side, call fdm_solve");
    return FDM_ERR_INVALID;
  }
  if (pick_solver(p, batch, o.solver) != FDM_SOLVER_CHOLESKY_BATCHED || !chol_warp_from_q_supported(p)) {
    fdm_set_error("fdm_solve_from_q: only the warp-per-design band Cholesky (hbw <= 30) assembles in place; "
                  "call fdm_assemble + fdm_solve");
    return FDM_ERR_UNSUPPORTED;
  }
  if (o.nan_on_failure && !info) {
    fdm_set_error("fdm_solve_from_q: nan_on_failure needs d_info");
    return FDM_ERR_INVALID;
  }
  return chol_warp_solve_from_q(p, q, xs, loads, x, info, ws, ws_bytes, batch, qs, xss, ls, o, (cudaStream_t)stream);
}

int fdm_spmm(const fdm_plan* p, const double* Kdata, const double* X, double* Y, int64_t batch,
             void* stream) {
  if (!ready(p, "fdm_spmm")) return FDM_ERR_INVALID;
  if (!Kdata || !X || !Y || batch <= 0) {
    fdm_set_error("fdm_spmm: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_spmm(p, Kdata, X, Y, batch, (cudaStream_t)stream);
}

int fdm_positions(const fdm_plan* p, double* xf, double* xs, double* xyz, int64_t batch, int64_t xss,
                  int transpose, void* stream) {
  if (!ready(p, "fdm_positions")) return FDM_ERR_INVALID;
  if (!xyz || batch <= 0 || (!transpose && (!xf || !xs))) {
    fdm_set_error("fdm_positions: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_positions(p, xf, xs, xyz, batch, xss, transpose, (cudaStream_t)stream);
}

int fdm_state(const fdm_plan* p, const double* q, const double* xyz, const double* loads,
              double* vectors, double* lengths, double* forces, double* residuals, int64_t batch,
              int64_t qs, int64_t ls, void* stream) {
  if (!ready(p, "fdm_state")) return FDM_ERR_INVALID;
  if (!q || !xyz || batch <= 0 || (residuals && !loads)) {
    fdm_set_error("fdm_state: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_state(p, q, xyz, loads, vectors, lengths, forces, residuals, batch, qs, ls,
                      (cudaStream_t)stream);
}

int fdm_state_vjp(const fdm_plan* p, const double* q, const double* xyz, const double* xyz_cot,
                  const double* rbar, const double* lbar, const double* fbar, const double* loads_cot,
                  const double* vbar, double* q_bar, double* xyz_bar, double* loads_bar,
                  int64_t batch, int64_t qs, void* stream) {
  if (!ready(p, "fdm_state_vjp")) return FDM_ERR_INVALID;
  if (!q || !xyz || batch <= 0) {
    fdm_set_error("fdm_state_vjp: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_state_vjp(p, q, xyz, xyz_cot, rbar, lbar, fbar, loads_cot, vbar, q_bar, xyz_bar,
                          loads_bar, batch, qs, (cudaStream_t)stream);
}

int fdm_state_jvp(const fdm_plan* p, const double* q, const double* xyz, const double* q_dot, const double* xyz_dot,
                  const double* loads_dot, double* vectors_dot, double* lengths_dot, double* forces_dot,
                  double* residuals_dot, int64_t batch, int64_t q_stride, int64_t xyz_stride, int64_t q_dot_stride,
                  int64_t loads_dot_stride, void* stream) {
  if (!ready(p, "fdm_state_jvp")) return FDM_ERR_INVALID;
  if (!q || !xyz || !xyz_dot || batch <= 0) {
    fdm_set_error("fdm_state_jvp: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_state_jvp(p, q, xyz, q_dot, xyz_dot, loads_dot, vectors_dot, lengths_dot, forces_dot, residuals_dot, batch,
                          q_stride, xyz_stride, q_dot_stride, loads_dot_stride, (cudaStream_t)stream);
}

int fdm_adjoint_grads(const fdm_plan* p, const double* q, const double* xyz, const double* lam,
                      double* q_bar, double* loads_bar, double* xs_bar, int64_t batch, int64_t qs,
                      int accumulate, void* stream) {
  if (!ready(p, "fdm_adjoint_grads")) return FDM_ERR_INVALID;
  if (!q || !xyz || !lam || batch <= 0) {
    fdm_set_error("fdm_adjoint_grads: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_adjoint_grads(p, q, xyz, lam, q_bar, loads_bar, xs_bar, batch, qs, accumulate,
                              (cudaStream_t)stream);
}

int fdm_edge_loads(const fdm_plan* p, const double* xyz, const double* loads_nodes, const double* edges_load,
                   double* loads_out, int64_t batch, int64_t ls, int64_t es, int is_local, void* stream) {
  if (!ready(p, "fdm_edge_loads")) return FDM_ERR_INVALID;
  if (!xyz || !loads_nodes || !edges_load || !loads_out || batch <= 0) {
    fdm_set_error("fdm_edge_loads: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_edge_loads(p, xyz, loads_nodes, edges_load, loads_out, batch, ls, es, is_local, (cudaStream_t)stream);
}

int fdm_edge_loads_vjp(const fdm_plan* p, const double* xyz, const double* edges_load, const double* cot,
                       double* xyz_bar, double* edges_load_bar, int64_t batch, int64_t es, int is_local, void* stream) {
  if (!ready(p, "fdm_edge_loads_vjp")) return FDM_ERR_INVALID;
  if (!xyz || !edges_load || !cot || batch <= 0) {
    fdm_set_error("fdm_edge_loads_vjp: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_edge_loads_vjp(p, xyz, edges_load, cot, xyz_bar, edges_load_bar, batch, es, is_local, (cudaStream_t)stream);
}

int fdm_face_loads(const fdm_plan* p, const int32_t* faces, int64_t F, int64_t D, const int32_t* edges_faces,
                   const double* xyz, const double* loads_in, const double* faces_load, double* cen, double* gl,
                   double* out, int64_t batch, int64_t ls, int64_t fs, int is_local, void* stream) {
  if (!ready(p, "fdm_face_loads")) return FDM_ERR_INVALID;
  if (!faces || !edges_faces || !xyz || !loads_in || !faces_load || !cen || !gl || !out || F <= 0 || D < 3 || batch <= 0) {
    fdm_set_error("fdm_face_loads: bad arguments");
    return FDM_ERR_INVALID;
  }
  if (D > 8) {
    fdm_set_error("fdm_face_loads: faces with more than 8 vertices are not supported");
    return FDM_ERR_UNSUPPORTED;
  }
  return launch_face_loads(p, faces, F, D, edges_faces, xyz, loads_in, faces_load, cen, gl, out, batch, ls, fs, is_local,
                           (cudaStream_t)stream);
}

int fdm_face_loads_vjp(const fdm_plan* p, const int32_t* faces, int64_t F, int64_t D, const int32_t* edges_faces,
                       const int32_t* faces_edges, const int32_t* nf_ptr, const int32_t* nf, const double* xyz,
                       const double* faces_load, const double* cen, const double* gl, const double* cot,
                       double* cen_bar, double* vert_bar, double* xyz_bar, double* fl_bar, int64_t batch, int64_t fs,
                       int is_local, void* stream) {
  if (!ready(p, "fdm_face_loads_vjp")) return FDM_ERR_INVALID;
  if (!faces || !edges_faces || !faces_edges || !nf_ptr || !nf || !xyz || !faces_load || !cen || !gl || !cot ||
      !cen_bar || (is_local && !vert_bar) || F <= 0 || D < 3 || D > 8 || batch <= 0) {
    fdm_set_error("fdm_face_loads_vjp: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_face_loads_vjp(p, faces, F, D, edges_faces, faces_edges, nf_ptr, nf, xyz, faces_load, cen, gl, cot,
                               cen_bar, vert_bar, xyz_bar, fl_bar, batch, fs, is_local, (cudaStream_t)stream);
}

int fdm_goals_loss(const fdm_plan* p, const fdm_goal_program* g, const double* xyz, const double* lengths,
                   const double* forces, const double* residuals, double* loss, double* group_sums,
                   const double* loss_cot, double* xyz_bar, double* lengths_bar, double* forces_bar, double* residuals_bar,
                   int64_t batch, void* stream) {
  if (!ready(p, "fdm_goals_loss")) return FDM_ERR_INVALID;
  const bool any_bar = xyz_bar || lengths_bar || forces_bar || residuals_bar;
  if (!g || !xyz || (!loss && !any_bar) || batch <= 0 || g->num_terms < 0 || g->num_groups < 0 ||
      (g->num_terms > 0 && (!g->kind || !g->index || !g->error || !g->weight || !g->target)) ||
      (g->num_groups > 0 && (!g->group_ptr || !g->group_final || !g->group_inner || !g->group_outer))) {
    fdm_set_error("fdm_goals_loss: bad arguments");
    return FDM_ERR_INVALID;
  }
  if (g->num_aggregates > 16 || (g->num_aggregates > 0 && (!g->agg_ptr || !g->agg_index))) {
    fdm_set_error("fdm_goals_loss: at most 16 aggregate goals per loss, with their element sets");
    return FDM_ERR_UNSUPPORTED;
  }
  if (g->num_groups > 64) {
    fdm_set_error("fdm_goals_loss: more than 64 error terms in one loss");
    return FDM_ERR_UNSUPPORTED;
  }
  return launch_goals_loss(p, g->num_terms, g->kind, g->index, g->error, g->weight, g->target, g->num_groups,
                           g->group_ptr, g->group_final, g->group_inner, g->group_outer, g->has_network_loadpath,
                           g->agg_ptr, g->agg_index, g->num_aggregates, xyz,
                           lengths, forces, residuals, loss, group_sums, loss_cot, xyz_bar, lengths_bar, forces_bar, residuals_bar,
                           batch, (cudaStream_t)stream);
}

int fdm_loss_sqdist(const double* x, const double* x0, double* g, double* loss, int64_t n,
                    int64_t batch, int64_t x0_stride, void* stream) {
  if (!x || !x0 || n <= 0 || batch <= 0) {
    fdm_set_error("fdm_loss_sqdist: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_loss_sqdist(x, x0, g, loss, n, batch, x0_stride, (cudaStream_t)stream);
}

int fdm_kbar_pattern(const fdm_plan* p, const double* a, const double* b, double* Kbar, double sign, int64_t batch,
                     void* stream) {
  if (!ready(p, "fdm_kbar_pattern")) return FDM_ERR_INVALID;
  if (!a || !b || !Kbar || batch <= 0) {
    fdm_set_error("fdm_kbar_pattern: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_kbar_pattern(p, a, b, Kbar, sign, batch, (cudaStream_t)stream);
}

int fdm_stiffness_vjp(const fdm_plan* p, const double* Kbar, double* q_bar, int64_t batch, void* stream) {
  if (!ready(p, "fdm_stiffness_vjp")) return FDM_ERR_INVALID;
  if (!Kbar || !q_bar || batch <= 0) {
    fdm_set_error("fdm_stiffness_vjp: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_stiffness_vjp(p, Kbar, q_bar, batch, (cudaStream_t)stream);
}

int fdm_lincomb(double* out, const double* x, const double* y, const double* z, double a, double b, double c, int64_t n,
                void* stream) {
  if (!out || !x || n <= 0) {
    fdm_set_error("fdm_lincomb: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_lincomb(out, x, y, z, a, b, c, n, (cudaStream_t)stream);
}

int fdm_batch_sum(const double* d_in, double* d_out, int64_t batch, int64_t n, void* stream) {
  if (!d_in || !d_out || batch <= 0 || n <= 0) {
    fdm_set_error("fdm_batch_sum: bad arguments");
    return FDM_ERR_INVALID;
  }
  return launch_batch_sum(d_in, d_out, batch, n, (cudaStream_t)stream);
}

}  // extern "C"
