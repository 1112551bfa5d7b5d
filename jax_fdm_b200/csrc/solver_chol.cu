// Batched band Cholesky: one CTA per design, factor held in shared memory.
// Replaces the vmap(EquilibriumModel) dense solve (models.py:88, loss_plotter.py:98-103) for
// batches of small designs that share one topology.
//
// Band layout (column-oriented, band order = plan.band_perm): a[j*W + d] = A(j+d, j), W = hbw+1.
// sigma = sign of the first diagonal entry; the factor is of sigma*K (all q<0 => K = -SPD).
// The factor is also written to the workspace so the adjoint solve (sparse.py:296) and tmax>1
// iterations can reuse it (opts.reuse_factor) instead of refactoring as the reference does.
#include <cuda_runtime.h>

#include "launch.h"

namespace {

constexpr int kThreads = 256;

struct CholArgs {
  int n, W;                       // order, hbw+1
  const int32_t* kidx;            // (n, W) band slot -> position in K.data, -1 zero
  const int32_t* iperm;           // band position -> free index
  const double* Kdata;            // (B, nnz)
  const double* rhs;              // (B, n, 3) free order
  double* x;                      // (B, n, 3) free order
  double* info;                   // (B, stride) or null
  double* factor;                 // (B, n*W + 8) workspace
  int64_t nnz;
  int reuse;
};

__global__ void __launch_bounds__(kThreads) chol_band_kernel(CholArgs A) {
  extern __shared__ __align__(16) double smem[];
  const int n = A.n, W = A.W, bw = W - 1;
  double* a = smem;                      // n*W
  double* y = a + (size_t)n * W;         // n*3 (rhs -> solution), stored [c*n + j]
  __shared__ double s_inv;
  __shared__ int s_bad;
  const int64_t b = blockIdx.x;
  const double* Kd = A.Kdata + b * A.nnz;
  double* fac = A.factor + b * ((int64_t)n * W + 8);
  const int tid = threadIdx.x;

  double sigma;
  if (A.reuse) {
    for (int i = tid; i < n * W; i += kThreads) a[i] = fac[i];
    sigma = fac[(int64_t)n * W];
  } else {
    sigma = __ldg(Kd + __ldg(A.kidx)) < 0.0 ? -1.0 : 1.0;
    for (int i = tid; i < n * W; i += kThreads) {
      const int k = __ldg(A.kidx + i);
      a[i] = k >= 0 ? sigma * __ldg(Kd + k) : 0.0;
    }
  }
  for (int i = tid; i < n * 3; i += kThreads) {
    const int j = i / 3, c = i - 3 * j;
    y[c * n + j] = sigma * A.rhs[(b * n + __ldg(A.iperm + j)) * 3 + c];
  }
  // the pivot flag lives beside the factor (slot 1 after sigma): a factor that failed keeps reporting so when reused
  if (tid == 0) s_bad = A.reuse ? (int)fac[(int64_t)n * W + 1] : 0;
  __syncthreads();

  if (!A.reuse) {
    // right-looking column Cholesky inside the band
    for (int j = 0; j < n; ++j) {
      const int m = min(bw, n - 1 - j);
      if (tid == 0) {
        const double piv = a[j * W];
        if (!(piv > 0.0)) s_bad = 1;
        const double s = sqrt(piv);
        a[j * W] = s;
        s_inv = 1.0 / s;
      }
      __syncthreads();
      if (tid < m) a[j * W + 1 + tid] *= s_inv;
      __syncthreads();
      // trailing update: for 1 <= c <= r <= m:  A(j+r, j+c) -= l_r l_c
      const int tri = m * (m + 1) / 2;
      for (int t = tid; t < tri; t += kThreads) {
        // unrank t -> (c, r) with c <= r, enumerating column by column
        int c = 1, rem = t;
        while (rem >= m - c + 1) { rem -= m - c + 1; ++c; }
        const int r = c + rem;
        a[(j + c) * W + (r - c)] -= a[j * W + r] * a[j * W + c];
      }
      __syncthreads();
    }
    for (int i = tid; i < n * W; i += kThreads) fac[i] = a[i];
    if (tid == 0) {
      fac[(int64_t)n * W] = sigma;
      fac[(int64_t)n * W + 1] = (double)s_bad;
    }
  }

  // solves: warp c handles rhs column c (3 warps), column-oriented forward, row-oriented backward
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < 3) {
    double* yc = y + warp * n;
    for (int j = 0; j < n; ++j) {
      const int m = min(bw, n - 1 - j);
      const double yj = yc[j] / a[j * W];
      __syncwarp();
      if (lane == 0) yc[j] = yj;
      for (int d = 1 + lane; d <= m; d += 32) yc[j + d] -= a[j * W + d] * yj;
      __syncwarp();
    }
    for (int j = n - 1; j >= 0; --j) {
      const int m = min(bw, n - 1 - j);
      double acc = 0.0;
      for (int d = 1 + lane; d <= m; d += 32) acc += a[j * W + d] * yc[j + d];
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) yc[j] = (yc[j] - acc) / a[j * W];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < n * 3; i += kThreads) {
    const int j = i / 3, c = i - 3 * j;
    A.x[(b * n + __ldg(A.iperm + j)) * 3 + c] = y[c * n + j];
  }
  if (A.info && tid == 0) {
    double* inf = A.info + b * FDM_INFO_STRIDE;
    inf[FDM_INFO_STATUS] = s_bad ? (double)FDM_STATUS_INDEFINITE : (double)FDM_STATUS_OK;
    inf[FDM_INFO_ITERS] = 0.0;
    inf[FDM_INFO_RELRES] = 0.0;
    inf[FDM_INFO_SIGN] = sigma;
  }
}

size_t smem_bytes(const fdm_plan* p) {
  return ((size_t)p->Nf * (p->hbw + 1) + (size_t)p->Nf * 3) * sizeof(double);
}

}  // namespace

static bool chol_cta_supported(const fdm_plan* p) {
  return !p->band_kidx.empty() && smem_bytes(p) <= 227 * 1024 - 1024;
}

bool chol_supported(const fdm_plan* p) { return chol_warp_supported(p) || chol_cta_supported(p); }

size_t chol_workspace_bytes(const fdm_plan* p, int64_t batch) {
  if (chol_warp_supported(p)) return chol_warp_workspace_bytes(p, batch);
  if (!chol_cta_supported(p)) return 0;
  return (size_t)batch * ((size_t)p->Nf * (p->hbw + 1) + 8) * sizeof(double);
}

int chol_solve(const fdm_plan* p, const double* Kdata, const double* rhs, double* x, double* info,
               void* ws, size_t ws_bytes, int64_t batch, const fdm_solve_opts& o, cudaStream_t st) {
  // narrow bands (hbw <= 31): one warp per design, any n.  Wider bands that still fit one
  // CTA's shared memory: one CTA per design (below).
  if (chol_warp_supported(p)) return chol_warp_solve(p, Kdata, rhs, x, info, ws, ws_bytes, batch, o, st);
  if (!chol_cta_supported(p)) {
    fdm_set_error("chol_solve: band (n=" + std::to_string(p->Nf) + ", hbw=" + std::to_string(p->hbw) +
                  ") does not fit in shared memory");
    return FDM_ERR_UNSUPPORTED;
  }
  if (ws_bytes < chol_workspace_bytes(p, batch)) {
    fdm_set_error("chol_solve: workspace too small");
    return FDM_ERR_WORKSPACE;
  }
  const size_t smem = smem_bytes(p);
  FDM_CUDA(cudaFuncSetAttribute(chol_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CholArgs A;
  A.n = (int)p->Nf;
  A.W = p->hbw + 1;
  A.kidx = p->d.band_kidx;
  A.iperm = p->d.band_iperm;
  A.Kdata = Kdata;
  A.rhs = rhs;
  A.x = x;
  A.info = info;
  A.factor = (double*)ws;
  A.nnz = p->nnz;
  A.reuse = o.reuse_factor;
  chol_band_kernel<<<(unsigned)batch, kThreads, smem, st>>>(A);
  return fdm_check_launch("chol_solve");
}
