// Multi-kernel Jacobi-preconditioned CG for K X = B with 3 right-hand sides (any mesh size).
// Replaces sparse_solve / spsolve_* (sparse.py:54-159, 207-230).  The three coordinate
// systems share K, so each row's CSR entries are read once per iteration for all three.
//
// Three kernels per iteration (fixed grid, grid-stride rows):
//   spmv      Ap = K p,              per-block partial <p,Ap>
//   update    alpha = rz/<p,Ap>;  x += alpha p;  r -= alpha Ap;  per-block partial <r,D^-1 r>, <r,r>
//   direction beta = rz'/rz;      p = D^-1 r + beta p;  block 0 publishes scalars / convergence
// Scalars never leave the device: every block re-reduces the per-block partials (fixed order
// => deterministic).  The host polls one `done` word every `check_every` iterations.
// K negative definite (all q < 0) needs no special casing: CG's recurrences are invariant under
// (K, b, D) -> (-K, -b, -D).
#include <cuda_runtime.h>

#include "launch.h"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 1024;

struct PcgScalars {
  double rz[2][3];
  double bnorm2[3];
  double rr[3];
  int done;
  int iters;
  int status;
  int pad;
};

struct PcgWs {
  double *r, *p, *Ap, *dinv;     // (Nf,3) x3, (Nf)
  double *part_a, *part_b, *part_c;  // (kMaxBlocks,3) each
  PcgScalars* sc;
};

__device__ __forceinline__ void block_sum3(double& a, double& b, double& c) {
  __shared__ double red[3][kThreads / 32];
  __shared__ double out[3];
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // protect `red`/`out` against the previous call
  if (l == 0) { red[0][w] = a; red[1][w] = b; red[2][w] = c; }
  __syncthreads();
  if (w == 0) {
    a = l < kThreads / 32 ? red[0][l] : 0.0;
    b = l < kThreads / 32 ? red[1][l] : 0.0;
    c = l < kThreads / 32 ? red[2][l] : 0.0;
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (l == 0) { out[0] = a; out[1] = b; out[2] = c; }
  }
  __syncthreads();
  a = out[0]; b = out[1]; c = out[2];
}

// sum of `nblk` per-block partial triples, same value in every thread of every block
__device__ __forceinline__ void reduce_partials(const double* __restrict__ part, int nblk,
                                                double& a, double& b, double& c) {
  a = b = c = 0.0;
  for (int i = threadIdx.x; i < nblk; i += kThreads) {
    a += part[3 * i]; b += part[3 * i + 1]; c += part[3 * i + 2];
  }
  block_sum3(a, b, c);
}

__global__ void __launch_bounds__(kThreads)
pcg_init_kernel(int Nf, const int32_t* __restrict__ diag_pos, const double* __restrict__ Kdata,
                const double* __restrict__ rhs, double* __restrict__ x, PcgWs w) {
  double s0 = 0, s1 = 0, s2 = 0, n0 = 0, n1 = 0, n2 = 0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < Nf; i += gridDim.x * kThreads) {
    const double di = 1.0 / __ldg(Kdata + __ldg(diag_pos + i));
    w.dinv[i] = di;
    const double b0 = rhs[3 * i], b1 = rhs[3 * i + 1], b2 = rhs[3 * i + 2];
    x[3 * i] = 0.0; x[3 * i + 1] = 0.0; x[3 * i + 2] = 0.0;
    w.r[3 * i] = b0; w.r[3 * i + 1] = b1; w.r[3 * i + 2] = b2;
    w.p[3 * i] = di * b0; w.p[3 * i + 1] = di * b1; w.p[3 * i + 2] = di * b2;
    s0 += b0 * di * b0; s1 += b1 * di * b1; s2 += b2 * di * b2;
    n0 += b0 * b0; n1 += b1 * b1; n2 += b2 * b2;
  }
  block_sum3(s0, s1, s2);
  block_sum3(n0, n1, n2);
  if (threadIdx.x == 0) {
    w.part_a[3 * blockIdx.x] = s0; w.part_a[3 * blockIdx.x + 1] = s1; w.part_a[3 * blockIdx.x + 2] = s2;
    w.part_b[3 * blockIdx.x] = n0; w.part_b[3 * blockIdx.x + 1] = n1; w.part_b[3 * blockIdx.x + 2] = n2;
  }
}

__global__ void __launch_bounds__(kThreads) pcg_init_scalars_kernel(int nblk, PcgWs w) {
  double a, b, c, d, e, f;
  reduce_partials(w.part_a, nblk, a, b, c);
  reduce_partials(w.part_b, nblk, d, e, f);
  if (threadIdx.x == 0) {
    PcgScalars* s = w.sc;
    s->rz[0][0] = a; s->rz[0][1] = b; s->rz[0][2] = c;
    s->bnorm2[0] = d; s->bnorm2[1] = e; s->bnorm2[2] = f;
    s->rr[0] = d; s->rr[1] = e; s->rr[2] = f;
    s->iters = 0;
    s->status = FDM_STATUS_OK;
    s->done = (d == 0.0 && e == 0.0 && f == 0.0) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(kThreads)
pcg_spmv_kernel(int Nf, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const double* __restrict__ Kdata, PcgWs w) {
  if (w.sc->done) return;
  double s0 = 0, s1 = 0, s2 = 0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < Nf; i += gridDim.x * kThreads) {
    const int k0 = __ldg(rowptr + i), k1 = __ldg(rowptr + i + 1);
    double a0 = 0, a1 = 0, a2 = 0;
    for (int k = k0; k < k1; ++k) {
      const double v = __ldg(Kdata + k);
      const double* pc = w.p + 3 * (int64_t)__ldg(colidx + k);
      a0 += v * pc[0]; a1 += v * pc[1]; a2 += v * pc[2];
    }
    w.Ap[3 * i] = a0; w.Ap[3 * i + 1] = a1; w.Ap[3 * i + 2] = a2;
    s0 += w.p[3 * i] * a0; s1 += w.p[3 * i + 1] * a1; s2 += w.p[3 * i + 2] * a2;
  }
  block_sum3(s0, s1, s2);
  if (threadIdx.x == 0) {
    w.part_c[3 * blockIdx.x] = s0; w.part_c[3 * blockIdx.x + 1] = s1; w.part_c[3 * blockIdx.x + 2] = s2;
  }
}

__global__ void __launch_bounds__(kThreads)
pcg_update_kernel(int Nf, int iter, double* __restrict__ x, PcgWs w) {
  if (w.sc->done) return;
  double pap0, pap1, pap2;
  reduce_partials(w.part_c, gridDim.x, pap0, pap1, pap2);
  const double* rz = w.sc->rz[iter & 1];
  const double al0 = pap0 != 0.0 ? rz[0] / pap0 : 0.0;
  const double al1 = pap1 != 0.0 ? rz[1] / pap1 : 0.0;
  const double al2 = pap2 != 0.0 ? rz[2] / pap2 : 0.0;
  double s0 = 0, s1 = 0, s2 = 0, n0 = 0, n1 = 0, n2 = 0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < Nf; i += gridDim.x * kThreads) {
    const double di = w.dinv[i];
    x[3 * i] += al0 * w.p[3 * i]; x[3 * i + 1] += al1 * w.p[3 * i + 1]; x[3 * i + 2] += al2 * w.p[3 * i + 2];
    const double r0 = w.r[3 * i] - al0 * w.Ap[3 * i];
    const double r1 = w.r[3 * i + 1] - al1 * w.Ap[3 * i + 1];
    const double r2 = w.r[3 * i + 2] - al2 * w.Ap[3 * i + 2];
    w.r[3 * i] = r0; w.r[3 * i + 1] = r1; w.r[3 * i + 2] = r2;
    s0 += r0 * di * r0; s1 += r1 * di * r1; s2 += r2 * di * r2;
    n0 += r0 * r0; n1 += r1 * r1; n2 += r2 * r2;
  }
  block_sum3(s0, s1, s2);
  block_sum3(n0, n1, n2);
  if (threadIdx.x == 0) {
    w.part_a[3 * blockIdx.x] = s0; w.part_a[3 * blockIdx.x + 1] = s1; w.part_a[3 * blockIdx.x + 2] = s2;
    w.part_b[3 * blockIdx.x] = n0; w.part_b[3 * blockIdx.x + 1] = n1; w.part_b[3 * blockIdx.x + 2] = n2;
  }
}

__global__ void __launch_bounds__(kThreads)
pcg_direction_kernel(int Nf, int iter, double tol2, PcgWs w) {
  if (w.sc->done) return;
  double z0, z1, z2, n0, n1, n2;
  reduce_partials(w.part_a, gridDim.x, z0, z1, z2);
  reduce_partials(w.part_b, gridDim.x, n0, n1, n2);
  const double* rz = w.sc->rz[iter & 1];
  const double be0 = rz[0] != 0.0 ? z0 / rz[0] : 0.0;
  const double be1 = rz[1] != 0.0 ? z1 / rz[1] : 0.0;
  const double be2 = rz[2] != 0.0 ? z2 / rz[2] : 0.0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < Nf; i += gridDim.x * kThreads) {
    const double di = w.dinv[i];
    w.p[3 * i] = di * w.r[3 * i] + be0 * w.p[3 * i];
    w.p[3 * i + 1] = di * w.r[3 * i + 1] + be1 * w.p[3 * i + 1];
    w.p[3 * i + 2] = di * w.r[3 * i + 2] + be2 * w.p[3 * i + 2];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    PcgScalars* s = w.sc;
    double* nz = s->rz[(iter + 1) & 1];
    nz[0] = z0; nz[1] = z1; nz[2] = z2;
    s->rr[0] = n0; s->rr[1] = n1; s->rr[2] = n2;
    s->iters = iter + 1;
    const bool finite = isfinite(z0) && isfinite(z1) && isfinite(z2);
    const bool conv = n0 <= tol2 * s->bnorm2[0] && n1 <= tol2 * s->bnorm2[1] && n2 <= tol2 * s->bnorm2[2];
    if (!finite) s->status = FDM_STATUS_BREAKDOWN;
    if (conv || !finite) s->done = 1;
  }
}

__global__ void pcg_finish_kernel(int Nf, const int32_t* __restrict__ diag_pos,
                                  const double* __restrict__ Kdata, PcgWs w, double* info) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  PcgScalars* s = w.sc;
  double rel = 0.0;
  for (int c = 0; c < 3; ++c)
    if (s->bnorm2[c] > 0.0) rel = fmax(rel, sqrt(s->rr[c] / s->bnorm2[c]));
  int status = s->status;
  if (status == FDM_STATUS_OK && !s->done) status = FDM_STATUS_NOT_CONVERGED;
  if (info) {
    info[FDM_INFO_STATUS] = (double)status;
    info[FDM_INFO_ITERS] = (double)s->iters;
    info[FDM_INFO_RELRES] = rel;
    info[FDM_INFO_SIGN] = Kdata[diag_pos[0]] < 0.0 ? -1.0 : 1.0;
  }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

PcgWs carve(const fdm_plan* p, void* ws) {
  char* c = (char*)ws;
  PcgWs w;
  const size_t v3 = align_up((size_t)p->Nf * 3 * 8, 256), v1 = align_up((size_t)p->Nf * 8, 256);
  const size_t pp = align_up((size_t)kMaxBlocks * 3 * 8, 256);
  w.r = (double*)c; c += v3;
  w.p = (double*)c; c += v3;
  w.Ap = (double*)c; c += v3;
  w.dinv = (double*)c; c += v1;
  w.part_a = (double*)c; c += pp;
  w.part_b = (double*)c; c += pp;
  w.part_c = (double*)c; c += pp;
  w.sc = (PcgScalars*)c;
  return w;
}

}  // namespace

size_t pcg_workspace_bytes(const fdm_plan* p) {
  const size_t v3 = align_up((size_t)p->Nf * 3 * 8, 256), v1 = align_up((size_t)p->Nf * 8, 256);
  const size_t pp = align_up((size_t)kMaxBlocks * 3 * 8, 256);
  return 3 * v3 + v1 + 3 * pp + align_up(sizeof(PcgScalars), 256);
}

int pcg_solve(const fdm_plan* p, const double* Kdata, const double* rhs, double* x, double* info,
              void* ws, size_t ws_bytes, int64_t batch, const fdm_solve_opts& o, cudaStream_t st) {
  if (ws_bytes < pcg_workspace_bytes(p)) {
    fdm_set_error("pcg_solve: workspace too small");
    return FDM_ERR_WORKSPACE;
  }
  const int Nf = (int)p->Nf;
  const int nblk = (int)std::min<int64_t>(kMaxBlocks, std::min<int64_t>((Nf + kThreads - 1) / kThreads,
                                                                        (int64_t)p->sm_count * 4));
  const int max_iters = o.max_iters > 0 ? o.max_iters : 20000;
  const double tol = o.tol > 0 ? o.tol : 1e-12;
  const int check = o.check_every > 0 ? o.check_every : 50;
  PcgWs w = carve(p, ws);
  for (int64_t b = 0; b < batch; ++b) {
    const double* Kb = Kdata + b * p->nnz;
    const double* rb = rhs + b * p->Nf * 3;
    double* xb = x + b * p->Nf * 3;
    pcg_init_kernel<<<nblk, kThreads, 0, st>>>(Nf, p->d.diag_pos, Kb, rb, xb, w);
    pcg_init_scalars_kernel<<<1, kThreads, 0, st>>>(nblk, w);
    int it = 0;
    while (it < max_iters) {
      const int stop = std::min(max_iters, it + check);
      for (; it < stop; ++it) {
        pcg_spmv_kernel<<<nblk, kThreads, 0, st>>>(Nf, p->d.rowptr, p->d.colidx, Kb, w);
        pcg_update_kernel<<<nblk, kThreads, 0, st>>>(Nf, it, xb, w);
        pcg_direction_kernel<<<nblk, kThreads, 0, st>>>(Nf, it, tol * tol, w);
      }
      int done = 0;
      FDM_CUDA(cudaMemcpyAsync(&done, &w.sc->done, sizeof(int), cudaMemcpyDeviceToHost, st));
      FDM_CUDA(cudaStreamSynchronize(st));
      if (done) break;
    }
    pcg_finish_kernel<<<1, 32, 0, st>>>(Nf, p->d.diag_pos, Kb, w, info ? info + b * FDM_INFO_STRIDE : nullptr);
  }
  return fdm_check_launch("pcg_solve");
}
