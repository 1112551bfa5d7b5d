// Batched band Cholesky, ONE WARP PER DESIGN (half bandwidth <= 31), many designs per SM.
// Replaces the vmap(EquilibriumModel) dense solve (models.py:88, loss_plotter.py:98-103) for
// batches of small designs that share one topology, and the adjoint solve of
// sparse_solve_bwd (sparse.py:296) with the factor reused instead of recomputed.
//
// Layout.  Band order = plan.band_perm.  Lane (R & 31) owns row R of the lower triangle while R
// is inside the 32-row window [j, j+32) of the right-looking factorisation; the row lives in 32
// REGISTERS, slot (C & 31) holding A(R, C).  Step j:
//     piv  = A(j,j)                    shuffle from lane j&31
//     l    = A(R,j) / sqrt(piv)        every lane, own register (exactly 0 outside the band)
//     A(R, j+k) -= l_R * l_{j+k}       k = 1..KMAX: l_{j+k} is a shared-memory BROADCAST (one
//                                      wavefront per pair), the accumulator never leaves registers
// so a step costs ~KMAX FP64 FMAs and KMAX/2 uniform LDS.128 per lane -- no matrix traffic through
// shared memory.  The updates with j+k > R fall into slots of columns that are already
// eliminated (dead) and are overwritten when the lane takes row R+32, whose 32 slots are read
// from a staged shared-memory block (filled once per 32 steps from K.data through the
// row-oriented index table plan.band_rowkidx).  The three right-hand sides ride along as three
// more registers per lane (forward substitution fused into the factorisation).
//
// The factor goes to the workspace column by column (fac[j][d] = L(j+d, j), d = 1..31, and
// fac[j][0] = 1/L(j,j)): one coalesced 256-byte store per step.  The backward substitution and
// the adjoint's forward substitution read it back in 32x32 blocks staged in shared memory.
// Rows are padded to a multiple of 32 with identity rows, so there is no tail code.
// sigma = sign of the first diagonal entry; the factor is of sigma*K (all q<0 => K = -SPD).
#include <cuda_runtime.h>

#include "launch.h"

namespace {

constexpr int kWarps = 8;  // designs per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kFS = 34;    // row stride (doubles) of the staged 32x32 block: 16-byte aligned rows, conflict-free columns
constexpr int kSL = 40;    // doubles per broadcast buffer: [1..32] l, [34..36] solved rhs entries
constexpr int kWarpSmem = 32 * kFS + 2 * kSL;
constexpr unsigned kFull = 0xffffffffu;

struct WArgs {
  int n, bw, nblk;
  const int32_t* rowkidx;  // ((nblk+1)*32, 32)
  const int32_t* iperm;    // (n) band position -> free index
  const double* Kdata;     // (B, nnz)
  const double* rhs;       // (B, n, 3) free order
  double* x;               // (B, n, 3) free order
  double* info;            // (B, stride) or null
  double* ws;              // (B, ws_stride)
  int64_t ws_stride;       // nblk*1024 (factor) + nblk*128 (intermediate rhs) + 8 (sigma, bad)
  int64_t nnz, batch;
  int reuse;
};

template <int KMAX>
__global__ void __launch_bounds__(kThreads, 2) chol_warp_kernel(WArgs A) {
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * kWarps + warp;
  if (b >= A.batch) return;  // warps are independent: no block-level barrier anywhere below
  double* fs = smem + warp * kWarpSmem;
  double* sl = fs + 32 * kFS;
  const int n = A.n, nblk = A.nblk;
  const double* Kd = A.Kdata + b * A.nnz;
  const double* rhs = A.rhs + b * (int64_t)n * 3;
  double* fac = A.ws + b * A.ws_stride;
  double* ysol = fac + (int64_t)nblk * 1024;
  double* meta = ysol + (int64_t)nblk * 128;
  double sigma;
  int bad = 0;

  auto load_rhs = [&](int R, double (&out)[3]) {
    if (R < n) {
      const int g = __ldg(A.iperm + R);
      out[0] = sigma * rhs[3 * g];
      out[1] = sigma * rhs[3 * g + 1];
      out[2] = sigma * rhs[3 * g + 2];
    } else {
      out[0] = out[1] = out[2] = 0.0;
    }
  };
  // fs[r][c] = fac[j0 + r][c]
  auto stage_fac = [&](int j0) {
    const double* src = fac + (int64_t)j0 * 32 + lane;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) fs[r * kFS + lane] = src[r * 32];
  };

  if (!A.reuse) {
    // ------------------------------------------------------------------ factor + forward substitution
    sigma = __ldg(Kd + __ldg(A.rowkidx)) < 0.0 ? -1.0 : 1.0;
    // fs[r][s] = sigma * A(R0 + r, C) for the column C = s (mod 32) of the 32 columns ending at row R0 + r
    auto stage_rows = [&](int R0) {
      const int32_t* idxp = A.rowkidx + (int64_t)R0 * 32;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const int idx = __ldg(idxp + r * 32 + ((r - lane) & 31));
        double v = 0.0;
        if (idx >= 0) v = sigma * __ldg(Kd + idx);
        else if (idx == -2) v = 1.0;
        fs[r * kFS + lane] = v;
      }
    };
    double a[32], y[3], yn[3];
    stage_rows(0);
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 32; s += 2) {
      const double2 v = *reinterpret_cast<const double2*>(fs + lane * kFS + s);
      a[s] = v.x;
      a[s + 1] = v.y;
    }
    load_rhs(lane, y);
    for (int jb = 0; jb < nblk; ++jb) {
      const int j0 = jb * 32;
      __syncwarp();  // every lane has consumed the previous staged block
      stage_rows(j0 + 32);
      load_rhs(j0 + 32 + lane, yn);
      __syncwarp();
#pragma unroll
      for (int p = 0; p < 32; ++p) {
        const int j = j0 + p;
        const int delta = (lane - p) & 31;
        double* buf = sl + (p & 1) * kSL;
        const double piv = __shfl_sync(kFull, a[p], p);
        bad |= !(piv > 0.0);
        const double inv = rsqrt(piv);
        const double l = a[p] * inv;
        fac[(int64_t)j * 32 + delta] = delta == 0 ? inv : l;
        buf[delta + 1] = l;
        if (delta == 0) {
          const double y0 = y[0] * inv, y1 = y[1] * inv, y2 = y[2] * inv;
          buf[34] = y0; buf[35] = y1; buf[36] = y2;
          double* yo = ysol + (int64_t)j * 4;
          yo[0] = y0; yo[1] = y1; yo[2] = y2;
        }
        __syncwarp();
#pragma unroll
        for (int k = 1; k <= KMAX; k += 2) {
          const double2 lc = *reinterpret_cast<const double2*>(buf + k + 1);
          a[(p + k) & 31] = fma(-l, lc.x, a[(p + k) & 31]);
          if (k + 1 <= KMAX) a[(p + k + 1) & 31] = fma(-l, lc.y, a[(p + k + 1) & 31]);
        }
        {
          const double2 t = *reinterpret_cast<const double2*>(buf + 34);
          const double t2 = buf[36];
          y[0] = fma(-l, t.x, y[0]);
          y[1] = fma(-l, t.y, y[1]);
          y[2] = fma(-l, t2, y[2]);
        }
        if (delta == 0) {  // row j is done: this lane takes row j + 32
#pragma unroll
          for (int s = 0; s < 32; s += 2) {
            const double2 v = *reinterpret_cast<const double2*>(fs + p * kFS + s);
            a[s] = v.x;
            a[s + 1] = v.y;
          }
          y[0] = yn[0]; y[1] = yn[1]; y[2] = yn[2];
        }
      }
    }
    if (lane == 0) {
      meta[0] = sigma;
      meta[1] = (double)bad;
    }
  } else {
    // ------------------------------------------------------------------ forward substitution with the stored factor
    sigma = meta[0];
    bad = meta[1] != 0.0;
    double acc[3], yn[3];
    load_rhs(lane, acc);
    for (int jb = 0; jb < nblk; ++jb) {
      const int j0 = jb * 32;
      __syncwarp();
      stage_fac(j0);
      load_rhs(j0 + 32 + lane, yn);
      __syncwarp();
#pragma unroll
      for (int p = 0; p < 32; ++p) {
        const int delta = (lane - p) & 31;
        const double inv = fs[p * kFS];
        const double lv = fs[p * kFS + delta];
        const double z0 = __shfl_sync(kFull, acc[0] * inv, p);
        const double z1 = __shfl_sync(kFull, acc[1] * inv, p);
        const double z2 = __shfl_sync(kFull, acc[2] * inv, p);
        if (delta == 0) {
          double* yo = ysol + (int64_t)(j0 + p) * 4;
          yo[0] = z0; yo[1] = z1; yo[2] = z2;
          acc[0] = yn[0]; acc[1] = yn[1]; acc[2] = yn[2];
        } else {
          acc[0] = fma(-lv, z0, acc[0]);
          acc[1] = fma(-lv, z1, acc[1]);
          acc[2] = fma(-lv, z2, acc[2]);
        }
      }
    }
  }

  // -------------------------------------------------------------------- backward substitution  L^T x = y
  double xprev[3] = {0.0, 0.0, 0.0};
  for (int jb = nblk - 1; jb >= 0; --jb) {
    const int j0 = jb * 32;
    __syncwarp();  // orders this warp's factor / ysol stores before the loads below, and protects fs
    stage_fac(j0);
    double acc[3];
    {
      const double* yi = ysol + (int64_t)(j0 + lane) * 4;
      acc[0] = yi[0]; acc[1] = yi[1]; acc[2] = yi[2];
    }
    __syncwarp();
    const double* row = fs + lane * kFS;  // column j0+lane of L: row[d] = L(j0+lane+d, j0+lane)
    const double myinv = row[0];
    // rows of the block above: x_i for i = j0+32 .. j0+31+KMAX, held by lane i-j0-32 in xprev
#pragma unroll
    for (int ii = KMAX; ii >= 1; --ii) {
      const double x0 = __shfl_sync(kFull, xprev[0], ii - 1);
      const double x1 = __shfl_sync(kFull, xprev[1], ii - 1);
      const double x2 = __shfl_sync(kFull, xprev[2], ii - 1);
      if (lane >= ii) {
        const double lv = row[31 + ii - lane];
        acc[0] = fma(-lv, x0, acc[0]);
        acc[1] = fma(-lv, x1, acc[1]);
        acc[2] = fma(-lv, x2, acc[2]);
      }
    }
    double xf[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 31; r >= 0; --r) {
      const double x0 = __shfl_sync(kFull, acc[0] * myinv, r);
      const double x1 = __shfl_sync(kFull, acc[1] * myinv, r);
      const double x2 = __shfl_sync(kFull, acc[2] * myinv, r);
      if (lane < r) {
        const double lv = row[r - lane];
        acc[0] = fma(-lv, x0, acc[0]);
        acc[1] = fma(-lv, x1, acc[1]);
        acc[2] = fma(-lv, x2, acc[2]);
      } else if (lane == r) {
        xf[0] = x0; xf[1] = x1; xf[2] = x2;
      }
    }
    xprev[0] = xf[0]; xprev[1] = xf[1]; xprev[2] = xf[2];
    const int R = j0 + lane;
    if (R < n) {
      double* xo = A.x + (b * (int64_t)n + __ldg(A.iperm + R)) * 3;
      xo[0] = xf[0]; xo[1] = xf[1]; xo[2] = xf[2];
    }
  }
  if (A.info && lane == 0) {
    double* inf = A.info + b * FDM_INFO_STRIDE;
    inf[FDM_INFO_STATUS] = bad ? (double)FDM_STATUS_INDEFINITE : (double)FDM_STATUS_OK;
    inf[FDM_INFO_ITERS] = 0.0;
    inf[FDM_INFO_RELRES] = 0.0;
    inf[FDM_INFO_SIGN] = sigma;
  }
}

inline int64_t ws_stride_of(const fdm_plan* p) {
  const int64_t nblk = (p->Nf + 31) / 32;
  return nblk * 1024 + nblk * 128 + 8;
}

}  // namespace

bool chol_warp_supported(const fdm_plan* p) { return !p->band_rowkidx.empty(); }

size_t chol_warp_workspace_bytes(const fdm_plan* p, int64_t batch) {
  return (size_t)batch * (size_t)ws_stride_of(p) * sizeof(double);
}

int chol_warp_solve(const fdm_plan* p, const double* Kdata, const double* rhs, double* x, double* info,
                    void* ws, size_t ws_bytes, int64_t batch, const fdm_solve_opts& o, cudaStream_t st) {
  if (!chol_warp_supported(p)) {
    fdm_set_error("chol_warp_solve: half bandwidth " + std::to_string(p->hbw) + " > 31");
    return FDM_ERR_UNSUPPORTED;
  }
  if (ws_bytes < chol_warp_workspace_bytes(p, batch)) {
    fdm_set_error("chol_warp_solve: workspace too small");
    return FDM_ERR_WORKSPACE;
  }
  WArgs A;
  A.n = (int)p->Nf;
  A.bw = p->hbw;
  A.nblk = (int)((p->Nf + 31) / 32);
  A.rowkidx = p->d.band_rowkidx;
  A.iperm = p->d.band_iperm;
  A.Kdata = Kdata;
  A.rhs = rhs;
  A.x = x;
  A.info = info;
  A.ws = (double*)ws;
  A.ws_stride = ws_stride_of(p);
  A.nnz = p->nnz;
  A.batch = batch;
  A.reuse = o.reuse_factor;
  void (*kern)(WArgs) = nullptr;
  if (p->hbw <= 7) kern = chol_warp_kernel<7>;
  else if (p->hbw <= 15) kern = chol_warp_kernel<15>;
  else if (p->hbw <= 23) kern = chol_warp_kernel<23>;
  else kern = chol_warp_kernel<31>;
  const size_t smem = (size_t)kWarps * kWarpSmem * sizeof(double);
  FDM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)((batch + kWarps - 1) / kWarps);
  kern<<<grid, kThreads, smem, st>>>(A);
  return fdm_check_launch("chol_warp_solve");
}
