// Launcher declarations shared between the kernel translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>

#include "plan.h"

// gridDim.y carries the design index; chunk batches larger than its limit.
#define FDM_BATCH_LOOP(batch, b0, nb)                                                          \
  for (int64_t b0 = 0, nb = 0;                                                                 \
       b0 < (batch) && ((nb = std::min<int64_t>((int64_t)(batch) - b0, 65535)), true); b0 += nb)

inline int fdm_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fdm_set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return FDM_ERR_CUDA;
  }
  return FDM_OK;
}

#define FDM_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      fdm_set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                      \
      return FDM_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

// In-kernel bounds checks of the -DFDM_DEBUG build (jax_fdm_b200.build.build(defines=("FDM_DEBUG",))): the pool
// has no compute-sanitizer, so the index tables the kernels trust are checked where they are consumed; a failed
// check prints file:line and the values, then traps (the launch fails with an error the C ABI reports).
#ifdef FDM_DEBUG
#include <cstdio>
#define FDM_ASSERT(cond, fmt, ...)                                                                         \
  do {                                                                                                     \
    if (!(cond)) {                                                                                         \
      printf("FDM_ASSERT %s:%d (block %d thread %d): " #cond " -- " fmt "\n", __FILE__, __LINE__,          \
             (int)blockIdx.x, (int)threadIdx.x, ##__VA_ARGS__);                                            \
      __trap();                                                                                            \
    }                                                                                                      \
  } while (0)
#else
#define FDM_ASSERT(cond, fmt, ...) ((void)0)
#endif

// kernels_basic.cu
int launch_assemble(const fdm_plan* p, const double* q, const double* xs, const double* loads,
                    double* Kdata, double* rhs, int64_t batch, int64_t qs, int64_t xss, int64_t ls,
                    cudaStream_t st);
int launch_positions(const fdm_plan* p, double* xf, double* xs, double* xyz, int64_t batch,
                     int64_t xss, int transpose, cudaStream_t st);
int launch_state(const fdm_plan* p, const double* q, const double* xyz, const double* loads,
                 double* vectors, double* lengths, double* forces, double* residuals,
                 int64_t batch, int64_t qs, int64_t ls, cudaStream_t st);
int launch_state_jvp(const fdm_plan* p, const double* q, const double* xyz, const double* q_dot, const double* xyz_dot,
                     const double* loads_dot, double* vd, double* ld, double* fd, double* rd, int64_t batch, int64_t qs,
                     int64_t xs, int64_t qds, int64_t lds, cudaStream_t st);
int launch_state_vjp(const fdm_plan* p, const double* q, const double* xyz, const double* xyz_cot,
                     const double* rbar, const double* lbar, const double* fbar,
                     const double* loads_cot, const double* vbar, double* q_bar, double* xyz_bar,
                     double* loads_bar, int64_t batch, int64_t qs, cudaStream_t st);
int launch_adjoint_grads(const fdm_plan* p, const double* q, const double* xyz, const double* lam,
                         double* q_bar, double* loads_bar, double* xs_bar, int64_t batch, int64_t qs,
                         int accumulate, cudaStream_t st);
int launch_edge_loads(const fdm_plan* p, const double* xyz, const double* loads_nodes, const double* edges_load,
                      double* loads_out, int64_t batch, int64_t ls, int64_t es, int local, cudaStream_t st);
int launch_edge_loads_vjp(const fdm_plan* p, const double* xyz, const double* edges_load, const double* cot,
                          double* xyz_bar, double* edges_load_bar, int64_t batch, int64_t es, int local, cudaStream_t st);
int launch_face_loads(const fdm_plan* p, const int32_t* faces, int64_t F, int64_t D, const int32_t* edges_faces,
                      const double* xyz, const double* loads_in, const double* faces_load, double* cen, double* gl,
                      double* out, int64_t batch, int64_t ls, int64_t fs, int local, cudaStream_t st);
int launch_face_loads_vjp(const fdm_plan* p, const int32_t* faces, int64_t F, int64_t D, const int32_t* edges_faces,
                          const int32_t* faces_edges, const int32_t* nf_ptr, const int32_t* nf, const double* xyz,
                          const double* faces_load, const double* cen, const double* gl, const double* cot,
                          double* cen_bar, double* vert_bar, double* xyz_bar, double* fl_bar, int64_t batch, int64_t fs,
                          int local, cudaStream_t st);
int launch_goals_loss(const fdm_plan* p, int n_terms, const int32_t* kind, const int32_t* index, const int32_t* err,
                      const double* weight, const double* target, int n_groups, const int32_t* group_ptr,
                      const int32_t* gfinal, const double* ginner, const double* gouter, int need_loadpath,
                      const int32_t* agg_ptr, const int32_t* agg_index, int num_agg,
                      const double* xyz, const double* lengths, const double* forces, const double* residuals,
                      double* loss, double* gsum, const double* loss_cot, double* xyz_bar, double* lengths_bar,
                      double* forces_bar, double* residuals_bar, int64_t batch, cudaStream_t st);
int launch_kbar_pattern(const fdm_plan* p, const double* a, const double* bvec, double* out, double sign, int64_t batch,
                        cudaStream_t st);
int launch_stiffness_vjp(const fdm_plan* p, const double* Kbar, double* q_bar, int64_t batch, cudaStream_t st);
int launch_lincomb(double* out, const double* x, const double* y, const double* z, double a, double b, double c, int64_t n,
                   cudaStream_t st);
int launch_batch_sum(const double* in, double* out, int64_t batch, int64_t n, cudaStream_t st);
// x[b] = NaN for every design whose info status is not OK (fdm_solve_opts.nan_on_failure)
int launch_nan_failed(const double* info, double* x, int64_t n_per_design, int64_t batch, cudaStream_t st);
int launch_loss_sqdist(const double* x, const double* x0, double* g, double* loss, int64_t n,
                       int64_t batch, int64_t x0_stride, cudaStream_t st);

// spmm.cu
int launch_spmm(const fdm_plan* p, const double* Kdata, const double* X, double* Y, int64_t batch,
                cudaStream_t st);

// solver_pcg.cu  (multi-kernel Jacobi-PCG, any size)
size_t pcg_workspace_bytes(const fdm_plan* p);
int pcg_solve(const fdm_plan* p, const double* Kdata, const double* rhs, double* x, double* info,
              void* ws, size_t ws_bytes, int64_t batch, const fdm_solve_opts& o, cudaStream_t st);

// solver_chol.cu  (batched band Cholesky)
size_t chol_workspace_bytes(const fdm_plan* p, int64_t batch);
bool chol_supported(const fdm_plan* p);
int chol_solve(const fdm_plan* p, const double* Kdata, const double* rhs, double* x, double* info,
               void* ws, size_t ws_bytes, int64_t batch, const fdm_solve_opts& o, cudaStream_t st);

// solver_chol_warp.cu  (batched band Cholesky, one warp per design, hbw <= 31)
bool chol_warp_supported(const fdm_plan* p);
size_t chol_warp_workspace_bytes(const fdm_plan* p, int64_t batch);
int chol_warp_solve(const fdm_plan* p, const double* Kdata, const double* rhs, double* x, double* info,
                    void* ws, size_t ws_bytes, int64_t batch, const fdm_solve_opts& o, cudaStream_t st);

bool chol_warp_from_q_supported(const fdm_plan* p);
bool chol_warp_is_factor_kernel(const void* func);   // a graph kernel node's function: one of the factor kernels?
int chol_warp_solve_from_q(const fdm_plan* p, const double* q, const double* xs, const double* loads, double* x,
                           double* info, void* ws, size_t ws_bytes, int64_t batch, int64_t qs, int64_t xss,
                           int64_t ls, const fdm_solve_opts& o, cudaStream_t st);

// solver_ldlt.cu  (band L D L^T in global memory, one CTA per design: the fallback for indefinite K)
bool ldlt_supported(const fdm_plan* p);
int64_t ldlt_fallback_slots(const fdm_plan* p, int64_t batch);
size_t ldlt_workspace_bytes(const fdm_plan* p, int64_t batch, int64_t slots);
int ldlt_solve(const fdm_plan* p, const double* Kdata, const double* rhs, double* x, double* info, void* ws,
               size_t ws_bytes, int64_t batch, int64_t slots, int only_failed, bool reuse, cudaStream_t st);

// solver_pcg2.cu  (persistent two-level PCG)
size_t pcg2_workspace_bytes(const fdm_plan* p);
bool pcg2_supported(const fdm_plan* p);
int pcg2_solve(const fdm_plan* p, const double* Kdata, const double* rhs, double* x, double* info,
               void* ws, size_t ws_bytes, int64_t batch, const fdm_solve_opts& o, cudaStream_t st);
void pcg2_destroy(fdm_plan* p);
void pcg2_prepare(fdm_plan* p);
