// Band L D L^T without pivoting in global memory, one CTA per design: the direct solver behind mixed-sign q
// (indefinite K) wherever the other solvers stop with FDM_STATUS_INDEFINITE -- one large mesh (the PCG solvers need
// a definite K), batches of wide-band designs (the shared-memory band Cholesky), half bandwidth 31 (the rank-1 warp
// kernel).  The reference's `spsolve` is an LU and solves such systems (sparse.py:147-149); this keeps that behaviour
// on the device.  It is a fallback, built for coverage more than for speed: a blocked right-looking sweep, 32 columns
// at a time through shared memory (the 40 000 columns of a 200x200 net: see DESIGN.md §5).
//
//   storage (per design, in the caller's workspace): Ab[j*(hbw+1) + d] = A(j+d, j) in band order (plan.band_kidx
//   says where each entry sits in K.data), overwritten by L (unit diagonal implied) with d_j in slot 0; then
//   dinv[j] = 1/d_j and three work vectors.
//   factor:   for j: l = A(j+1.., j) / d_j;  A(j+c.., j+c) -= l_c d_j l   (c = 1..hbw), blocked
//   solve:    L y = P b (column sweep), y /= d, L^T z = y (row dots), x = P^T z
//   a pivot below 1e-13 of the largest |diagonal| is a breakdown (FDM_STATUS_BREAKDOWN), as in the warp kernels;
//   the residual b - K x is formed against K.data, one step of iterative refinement follows when it is not at
//   rounding level, and max ||r|| / ||b|| goes to FDM_INFO_RELRES (above 1e-6 the design is reported as a breakdown).
//
// `only_failed`: the kernel returns at once for every design whose status is not FDM_STATUS_INDEFINITE, so fdm_solve
// can launch it behind another solver without a host round trip (CUDA-graph capturable).  Designs that do need it take
// a slot of the workspace from an atomic counter (the workspace holds min(batch, cap) slots); the slot is remembered
// per design so that the adjoint solve (reuse_factor) finds the factor again.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "launch.h"
#include "plan.h"

namespace {

constexpr int kLThreads = 1024;
constexpr int kNB = 32;              // columns per block

struct LArgs {
  int n, hbw, batch, cap, only_failed, reuse;
  int64_t nnz, per;                    // doubles per slot
  const int32_t* band_kidx;            // (n, hbw+1)
  const int32_t* band_iperm;           // band position -> free index
  const int32_t* rowptr;               // CSR pattern of K (free order): the residual check
  const int32_t* colidx;
  const double* Kdata;
  const double* rhs;
  double* x;
  double* info;
  int32_t* counter;                    // slots handed out so far
  int32_t* slots;                      // (batch) slot of a design, -1 none
  double* pool;                        // cap slots of `per` doubles
};

// Blocked right-looking sweep, kNB columns per block:
//   panel   : the block's columns (rows j0 .. j0 + kNB - 1 + hbw) are copied to shared memory, factored there
//             (two CTA barriers per column, all traffic on chip) and written back;
//   update  : the hbw x hbw window behind the panel loses P D P^T in ONE read-modify-write pass over global memory per
//             block (each thread a 1 x 2 tile, the panel read from shared memory);
//   solves  : the same panels, loaded block by block, sweep a shared-memory segment of the three right-hand sides.
// P[c][r] (c < kNB, r < kNB + hbw, rows relative to j0) = L(j0 + r, j0 + c), zero outside the band.
__device__ __forceinline__ void ldlt_load_panel(double* P, const double* Ab, int j0, int nb, int n, int hbw, int RP, int tid) {
  const int w = hbw + 1;
  for (int i = tid; i < kNB * RP; i += kLThreads) P[i] = 0.0;
  __syncthreads();
  for (int i = tid; i < nb * w; i += kLThreads) {
    const int c = i / w, d = i - c * w;
    if (j0 + c + d < n) P[c * RP + c + d] = Ab[(int64_t)(j0 + c) * w + d];
  }
  __syncthreads();
}

// y (n, 3) in band order <- (L D L^T)^-1 y, block by block through shared memory
__device__ void ldlt_substitute(double* P, double* sy, const double* Ab, const double* dinv, double* y, int n, int hbw,
                                int RP, int tid) {
  for (int j0 = 0; j0 < n; j0 += kNB) {                    // L y = b
    const int nb = min(kNB, n - j0), R = min(nb + hbw, n - j0);
    ldlt_load_panel(P, Ab, j0, nb, n, hbw, RP, tid);
    for (int i = tid; i < 3 * R; i += kLThreads) sy[i] = y[3 * (int64_t)j0 + i];
    __syncthreads();
    for (int c = 0; c < nb; ++c) {
      const double* pc = P + c * RP;
      const int top = min(c + hbw, R - 1);
      for (int i = tid; i < 3 * (top - c); i += kLThreads) {
        const int r = c + 1 + i / 3, k = i % 3;
        sy[3 * r + k] -= pc[r] * sy[3 * c + k];
      }
      __syncthreads();
    }
    for (int i = tid; i < 3 * R; i += kLThreads) y[3 * (int64_t)j0 + i] = sy[i];
    __syncthreads();
  }
  for (int i = tid; i < n; i += kLThreads) {               // y /= d
    const double di = dinv[i];
    y[3 * i] *= di; y[3 * i + 1] *= di; y[3 * i + 2] *= di;
  }
  __syncthreads();
  for (int j0 = ((n - 1) / kNB) * kNB; j0 >= 0; j0 -= kNB) {   // L^T z = y
    const int nb = min(kNB, n - j0), R = min(nb + hbw, n - j0);
    ldlt_load_panel(P, Ab, j0, nb, n, hbw, RP, tid);
    for (int i = tid; i < 3 * R; i += kLThreads) sy[i] = y[3 * (int64_t)j0 + i];
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int c = nb - 1; c >= 0; --c) {
      if (warp < 3) {                                      // warp k: component k of row c
        const double* pc = P + c * RP;
        const int top = min(c + hbw, R - 1);
        double acc = 0.0;
        for (int r = c + 1 + lane; r <= top; r += 32) acc += pc[r] * sy[3 * r + warp];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) sy[3 * c + warp] -= acc;
      }
      __syncthreads();
    }
    for (int i = tid; i < 3 * nb; i += kLThreads) y[3 * (int64_t)j0 + i] = sy[i];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kLThreads) ldlt_band_kernel(LArgs A) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_slot, s_bad, s_fresh;
  __shared__ double s_dmax, s_d[kNB];                      // s_d: reciprocal pivots of the current panel
  const int b = blockIdx.x, tid = threadIdx.x, n = A.n, hbw = A.hbw, w = hbw + 1;
  const int RP = (kNB + hbw + 2) & ~1;                     // rows of a panel in shared memory (even: 16-byte pairs)
  double* P = smem;                                        // [kNB][RP]
  double* sy = P + kNB * RP;                               // [RP][3]
  double* info = A.info + (int64_t)b * FDM_INFO_STRIDE;
  // only_failed 1: the designs flagged indefinite (behind a Cholesky kernel: its other failures are breakdowns this
  // factorisation would meet too); 2: every design an iterative solver did not finish (indefinite K with a
  // same-sign diagonal shows up there as a breakdown or as no convergence)
  const double st0 = info[FDM_INFO_STATUS];
  if ((A.only_failed == 1 && st0 != (double)FDM_STATUS_INDEFINITE) || (A.only_failed == 2 && st0 == (double)FDM_STATUS_OK)) {
    if (!A.reuse && tid == 0) A.slots[b] = -1;           // no factor of this design in the pool
    return;
  }
  if (tid == 0) {
    int slot = A.reuse ? A.slots[b] : -1;
    s_fresh = 0;
    if (slot < 0) {                                       // first visit (also: an adjoint solve whose forward solve
      slot = atomicAdd(A.counter, 1);                     // did not need the fallback): take a slot and factor
      if (slot >= A.cap) slot = -1;
      A.slots[b] = slot;
      s_fresh = 1;
    }
    s_slot = slot;
    s_bad = 0;
    s_dmax = 0.0;
  }
  __syncthreads();
  if (s_slot < 0) return;              // no room: the design keeps its status
  const bool factor = s_fresh != 0;
  const long long t_start = clock64();
  long long t_loaded = t_start, t_factored = t_start;
  double* Ab = A.pool + (int64_t)s_slot * A.per;
  double* dinv = Ab + (int64_t)n * w;
  double* y = dinv + n;                // (n, 3)
  if (factor) {
    const double* Kd = A.Kdata + (int64_t)b * A.nnz;
    double dm = 0.0;
    for (int64_t i = tid; i < (int64_t)n * w; i += kLThreads) {
      const int k = __ldg(A.band_kidx + i);
      const double v = k >= 0 ? Kd[k] : 0.0;
      Ab[i] = v;
      if (i % w == 0) dm = fmax(dm, fabs(v));
    }
    // largest |diagonal entry|: the breakdown threshold (any order: fmax is exact)
    for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
    if ((tid & 31) == 0 && dm > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(&s_dmax), (unsigned long long)__double_as_longlong(dm));
    __syncthreads();
    t_loaded = clock64();
    const double thr = 1e-13 * s_dmax;
    bool bad = false;
    for (int j0 = 0; j0 < n && !bad; j0 += kNB) {
      const int nb = min(kNB, n - j0);
      ldlt_load_panel(P, Ab, j0, nb, n, hbw, RP, tid);
      // ---- the panel, in shared memory.  Column c stays UNSCALED there (a_rc = l_rc d_c): the updates use
      //      l_c' d l_r = a_c' a_r / d, so a column costs one barrier; L = a / d is formed when the panel is written back
      for (int c = 0; c < nb; ++c) {
        const double* pc = P + c * RP;
        const double d = pc[c];                            // every thread reads the pivot: the test below is uniform
        if (!(fabs(d) > thr)) { bad = true; break; }
        const double di = 1.0 / d;
        if (tid == 0) { s_d[c] = di; dinv[j0 + c] = di; }
        const int ncol = nb - 1 - c;
        for (int i = tid; i < ncol * hbw; i += kLThreads) {
          const int cc = c + 1 + i / hbw, r = c + 1 + i % hbw;
          if (r >= cc) P[cc * RP + r] -= pc[cc] * pc[r] * di;
        }
        __syncthreads();
      }
      if (bad) break;
      // ---- write the panel back: d_c in slot 0, L below
      for (int i = tid; i < nb * w; i += kLThreads) {
        const int c = i / w, dd = i - c * w;
        if (j0 + c + dd < n) Ab[(int64_t)(j0 + c) * w + dd] = dd == 0 ? P[c * RP + c] : P[c * RP + c + dd] * s_d[c];
      }
      // ---- the window behind the panel: columns J = j0 + nb + cc, rows j0 + nb + rr, cc <= rr < hbw
      if (nb == kNB) {
        const int half = (hbw + 1) >> 1;                   // pairs of rows
        for (int i = tid; i < hbw * half; i += kLThreads) {
          const int cc = i / half, rr = 2 * (i - cc * half);
          const int J = j0 + kNB + cc;
          if (J >= n || rr + 1 < cc) continue;
          double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
          for (int c = 0; c < kNB; ++c) {
            const double f = s_d[c] * P[c * RP + kNB + cc];      // s_d = 1 / d_c: a_cc a_rr / d
            const double2 pr = *reinterpret_cast<const double2*>(P + c * RP + kNB + rr);
            a0 += f * pr.x;
            a1 += f * pr.y;
          }
          double* col = Ab + (int64_t)J * w - cc;
          if (rr >= cc && j0 + kNB + rr < n) col[rr] -= a0;
          if (rr + 1 >= cc && rr + 1 < hbw && j0 + kNB + rr + 1 < n) col[rr + 1] -= a1;
        }
      }
      __syncthreads();
    }
    if (bad && tid == 0) s_bad = 1;
    __syncthreads();
    if (s_bad) {
      if (tid == 0) {
        info[FDM_INFO_STATUS] = (double)FDM_STATUS_BREAKDOWN;
        A.slots[b] = -1;
      }
      return;
    }
  }
  t_factored = clock64();
  // ---- three right-hand sides: solve, residual against K.data, one step of iterative refinement when the residual
  //      is not at rounding level (no pivoting: element growth can cost digits on an indefinite K), residual reported
  const double* Kd = A.Kdata + (int64_t)b * A.nnz;
  const double* rhs = A.rhs + (int64_t)b * n * 3;
  double* x = A.x + (int64_t)b * n * 3;
  __shared__ double s_red[6];
  auto residual_into_y = [&]() {                           // y_j = b_f - (K x)_f, f = iperm[j]; returns max ||r|| / ||b||
    if (tid < 6) s_red[tid] = 0.0;
    __syncthreads();
    double rr[3] = {0.0, 0.0, 0.0}, bb[3] = {0.0, 0.0, 0.0};
    for (int j = tid; j < n; j += kLThreads) {
      const int f = __ldg(A.band_iperm + j);
      double r0 = rhs[3 * f], r1 = rhs[3 * f + 1], r2 = rhs[3 * f + 2];
      bb[0] += r0 * r0; bb[1] += r1 * r1; bb[2] += r2 * r2;
      for (int k = __ldg(A.rowptr + f); k < __ldg(A.rowptr + f + 1); ++k) {
        const double v = Kd[k];
        const double* xc = x + 3 * (int64_t)__ldg(A.colidx + k);
        r0 -= v * xc[0]; r1 -= v * xc[1]; r2 -= v * xc[2];
      }
      y[3 * j] = r0; y[3 * j + 1] = r1; y[3 * j + 2] = r2;
      rr[0] += r0 * r0; rr[1] += r1 * r1; rr[2] += r2 * r2;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      for (int o = 16; o > 0; o >>= 1) {
        rr[k] += __shfl_xor_sync(0xffffffffu, rr[k], o);
        bb[k] += __shfl_xor_sync(0xffffffffu, bb[k], o);
      }
      if ((tid & 31) == 0) { atomicAdd(&s_red[k], rr[k]); atomicAdd(&s_red[3 + k], bb[k]); }
    }
    __syncthreads();
    double rel = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) rel = fmax(rel, s_red[3 + k] > 0.0 ? sqrt(s_red[k] / s_red[3 + k]) : 0.0);
    __syncthreads();
    return rel;
  };
  for (int i = tid; i < n; i += kLThreads) {
    const int f = __ldg(A.band_iperm + i);
    y[3 * i] = rhs[3 * f]; y[3 * i + 1] = rhs[3 * f + 1]; y[3 * i + 2] = rhs[3 * f + 2];
  }
  __syncthreads();
  ldlt_substitute(P, sy, Ab, dinv, y, n, hbw, RP, tid);
  for (int i = tid; i < n; i += kLThreads) {
    const int f = __ldg(A.band_iperm + i);
    x[3 * f] = y[3 * i]; x[3 * f + 1] = y[3 * i + 1]; x[3 * f + 2] = y[3 * i + 2];
  }
  __syncthreads();
  double rel = residual_into_y();
  int refined = 0;
  if (rel > 1e-13 && isfinite(rel)) {
    ldlt_substitute(P, sy, Ab, dinv, y, n, hbw, RP, tid);
    for (int i = tid; i < n; i += kLThreads) {
      const int f = __ldg(A.band_iperm + i);
      x[3 * f] += y[3 * i]; x[3 * f + 1] += y[3 * i + 1]; x[3 * f + 2] += y[3 * i + 2];
    }
    __syncthreads();
    rel = residual_into_y();
    refined = 1;
  }
  if (tid == 0) {
    // a solution that does not satisfy the system is not handed out as one (the sums are atomics: the residual is
    // reproducible to rounding only, the solution itself is deterministic)
    info[FDM_INFO_STATUS] = rel <= 1e-6 ? (double)FDM_STATUS_OK : (double)FDM_STATUS_BREAKDOWN;
    info[FDM_INFO_ITERS] = (double)refined;
    info[FDM_INFO_RELRES] = rel;
    info[FDM_INFO_SIGN] = 0.0;
    // slots 4..6: cycles spent loading the band, factoring, substituting (diagnostic)
    info[4] = (double)(t_loaded - t_start); info[5] = (double)(t_factored - t_loaded); info[6] = (double)(clock64() - t_factored);
  }
}

inline int64_t per_design(const fdm_plan* p) { return (int64_t)p->Nf * (p->hbw + 1) + 4 * (int64_t)p->Nf; }
inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace

bool ldlt_supported(const fdm_plan* p) { return !p->band_kidx.empty() && p->d.band_kidx && p->d.band_iperm; }

// slots of the fallback behind another solver: what fits 256 MB (none when a single band is larger than that: the
// fallback must not multiply the workspace of a solver that would have done without it)
int64_t ldlt_fallback_slots(const fdm_plan* p, int64_t batch) {
  const int64_t bytes = per_design(p) * 8;
  return std::min<int64_t>(batch, (int64_t(256) << 20) / bytes);
}

size_t ldlt_workspace_bytes(const fdm_plan* p, int64_t batch, int64_t slots) {
  if (!ldlt_supported(p)) return 0;
  return 256 + align256((size_t)batch * 4) + (size_t)slots * (size_t)per_design(p) * 8;
}

int ldlt_solve(const fdm_plan* p, const double* Kdata, const double* rhs, double* x, double* info, void* ws,
               size_t ws_bytes, int64_t batch, int64_t slots, int only_failed, bool reuse, cudaStream_t st) {
  if (!ldlt_supported(p)) {
    fdm_set_error("ldlt: the plan has no band table (half bandwidth > 255)");
    return FDM_ERR_UNSUPPORTED;
  }
  if (!info || ws_bytes < ldlt_workspace_bytes(p, batch, slots)) {
    fdm_set_error("ldlt: needs d_info and a workspace of fdm_workspace_bytes");
    return FDM_ERR_INVALID;
  }
  LArgs A;
  std::memset(&A, 0, sizeof(A));
  A.n = (int)p->Nf; A.hbw = p->hbw; A.batch = (int)batch; A.cap = (int)slots;
  A.only_failed = only_failed; A.reuse = reuse ? 1 : 0;
  A.nnz = p->nnz; A.per = per_design(p);
  A.band_kidx = p->d.band_kidx; A.band_iperm = p->d.band_iperm;
  A.rowptr = p->d.rowptr; A.colidx = p->d.colidx;
  A.Kdata = Kdata; A.rhs = rhs; A.x = x; A.info = info;
  char* w = static_cast<char*>(ws);
  A.counter = reinterpret_cast<int32_t*>(w);
  A.slots = reinterpret_cast<int32_t*>(w + 256);
  A.pool = reinterpret_cast<double*>(w + 256 + align256((size_t)batch * 4));
  if (!reuse) FDM_CUDA(cudaMemsetAsync(A.counter, 0, 256, st));
  const int RP = (kNB + p->hbw + 2) & ~1;
  const size_t smem = (size_t)(kNB * RP + 3 * RP) * 8;
  FDM_CUDA(cudaFuncSetAttribute(ldlt_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ldlt_band_kernel<<<(unsigned)batch, kLThreads, smem, st>>>(A);
  return fdm_check_launch("ldlt_band");
}
