// Persistent two-level PCG: ONE cooperative kernel per 3-rhs solve, K and every CG vector
// resident on chip for the whole solve.
// Replaces sparse_solve / spsolve_* (sparse.py:54-159, 207-230) and the adjoint solve of
// sparse_solve_bwd (sparse.py:296) for meshes whose partition fits shared memory
// (the 200x200 cable-net of BASELINE config 4 uses ~40 KB per CTA).
//
// Decomposition: the free nodes are split into P <= #SM compact aggregates (plan.cpp,
// fdm_build_partition); CTA c owns aggregate c: its rows of K (local CSR, values gathered from
// K.data once, kept in shared memory), and x, r, A p for those rows in registers.  Only the
// search direction p is shared between threads (shared memory: own rows + halo).
//
// Preconditioner: M^-1 = D^-1 + P Ac^-1 P^T  (Jacobi + piecewise-constant coarse correction,
// Ac = P^T K P is PxP, inverted once per K by a small kernel and reused by the adjoint solve).
// The coarse residual P^T r is carried by recurrence (P^T r -= alpha P^T A p), so one CG
// iteration needs exactly two grid-wide exchanges:
//   A: all-gather { <p,Ap>_c [3],  (P^T A p)_c [3] }                      -> alpha, coarse residual
//   B: all-gather { <r,z>_c [3], <r,r>_c [3] } + boundary values of z     -> beta, halo of p
// Exchanges are flag-in-data ("LL"): every double travels as two 8-byte words {32 payload bits,
// 32-bit tag}; readers poll the words themselves, so an exchange costs one L2 round trip with
// no fence, no atomic and no separate barrier.  All CTAs reduce the gathered partials in the
// same fixed order, hence take identical branches and the result is run-to-run deterministic.
// CG is invariant under (K, b, M) -> (-K, -b, -M): all q < 0 (K = -SPD) needs no special case.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <map>

#include "launch.h"

// -DFDM_PCG2_PROFILE: CTA 0 accumulates clock64() per phase and writes the totals after the info
// record (slots 4..7: coarse+z, exchange B, halo+SpMV, exchange A); off in the shipped build.
#ifdef FDM_PCG2_PROFILE
#define PROF(...) __VA_ARGS__
#else
#define PROF(...)
#endif

namespace {

constexpr int kThreads = 512;
constexpr int kMaxRowsPerThread = 4;
constexpr int kChebDegree = 3;         // tuned on the 200x200 cable-net: 307 / 201 / 170 / 155 iterations at degree 1 / 2 / 3 / 4
constexpr double kChebAlpha = 10.0;
constexpr int kVals = 8;               // doubles per CTA record in every exchange

struct Pcg2Device {
  int P = 0, nmax = 0, emax = 0, hmax = 0, xmax = 0, rows_per_thread = 1;
  size_t smem_bytes = 0;
  // per-part offsets (P+1 each)
  int32_t *row_off = nullptr, *ent_off = nullptr, *halo_off = nullptr, *exp_off = nullptr;
  // flat per-part data
  int32_t* rows = nullptr;        // (Nf) free index of each local row
  int32_t* lrowptr = nullptr;     // (Nf + P) local CSR row pointers (relative to the part's ent_off)
  uint16_t* lcol = nullptr;       // (#offdiag) local column: own [0,n_c) or halo n_c + h
  int32_t* lkidx = nullptr;       // (#offdiag) position in K.data
  int32_t* dkidx = nullptr;       // (Nf) position of the diagonal in K.data
  int32_t* exp_slot = nullptr;    // (Nf) export slot of each local row (global slot id) or -1
  int32_t* halo_src = nullptr;    // (#halo) global export slot each halo column reads
  // coarse assembly: Ac entry groups
  int32_t *cg_ptr = nullptr, *cg_ab = nullptr, *cg_kidx = nullptr;
  int n_cgroups = 0;
  int total_exports = 0;
};

struct Pcg2Args {
  Pcg2Device d;
  const double* Kdata;
  const double* rhs;
  double* x;
  double* info;
  double* ainv;          // (P,P) workspace
  uint64_t* xS;          // (P, kVals, 2) LL words
  uint64_t* xA;
  uint64_t* xB;
  uint64_t* xE;          // (total_exports, 3, 2) LL words
  double tol2;
  int max_iters;
  int cheb_k;            // degree of the Chebyshev smoother (1 = scaled Jacobi)
  double cheb_alpha;     // its interval is [2 / cheb_alpha, 2] in the spectrum of D^-1 K
};

// ------------------------------------------------------------------------------------------
// LL words
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint64_t* slot, double v, uint32_t tag) {
  const uint64_t bits = (uint64_t)__double_as_longlong(v);
  const uint64_t w0 = (bits & 0xffffffffull) | ((uint64_t)tag << 32);
  const uint64_t w1 = (bits >> 32) | ((uint64_t)tag << 32);
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ double ll_load(const uint64_t* slot, uint32_t tag) {
  uint64_t w0, w1;
  do {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
  } while ((uint32_t)(w0 >> 32) != tag || (uint32_t)(w1 >> 32) != tag);
  return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}

// ------------------------------------------------------------------------------------------
// coarse operator: Ac = P^T K P, Ainv = Ac^-1 (Gauss-Jordan in shared memory, one CTA); the other
// CTAs of this launch clear the exchange buffers (tags restart at 1 every solve).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) pcg2_setup_kernel(Pcg2Args a, size_t xwords) {
  if (blockIdx.x > 0) {
    // exchange buffers are contiguous: xS | xA | xB | xE
    const size_t per = (xwords + gridDim.x - 2) / (gridDim.x - 1);
    const size_t lo = (size_t)(blockIdx.x - 1) * per, hi = min(xwords, lo + per);
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) a.xS[i] = 0ull;
    return;
  }
  extern __shared__ __align__(16) double sm[];
  const int P = a.d.P;
  double* A = sm;  // P*P
  for (int i = threadIdx.x; i < P * P; i += blockDim.x) A[i] = 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int gidx = warp; gidx < a.d.n_cgroups; gidx += nw) {
    const int k0 = a.d.cg_ptr[gidx], k1 = a.d.cg_ptr[gidx + 1];
    double s = 0.0;
    for (int k = k0 + lane; k < k1; k += 32) s += __ldg(a.Kdata + __ldg(a.d.cg_kidx + k));
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) A[a.d.cg_ab[gidx]] = s;
  }
  __syncthreads();
  // in-place Gauss-Jordan inversion without pivoting (Ac is SPD or -SPD); warp per row.
  __shared__ double s_piv;
  for (int k = 0; k < P; ++k) {
    if (threadIdx.x == 0) s_piv = 1.0 / A[k * P + k];
    __syncthreads();
    const double ipiv = s_piv;
    for (int i = warp; i < P; i += nw) {
      if (i == k) continue;
      const double f = A[i * P + k] * ipiv;
      for (int j = lane; j < P; j += 32)
        if (j != k) A[i * P + j] -= f * A[k * P + j];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < P; j += blockDim.x) {
      if (j != k) {
        const double rkj = A[k * P + j] * ipiv;
        const double cik = -A[j * P + k] * ipiv;
        A[k * P + j] = rkj;
        A[j * P + k] = cik;
      }
    }
    if (threadIdx.x == 0) A[k * P + k] = ipiv;
    __syncthreads();
  }
  for (int i = threadIdx.x; i < P * P; i += blockDim.x) a.ainv[i] = A[i];
}

// ------------------------------------------------------------------------------------------
// the solver
// ------------------------------------------------------------------------------------------
// Block partial sums -> this CTA's record.  Every thread t < nact has stored its V partials in
// s_part[k*kThreads + t]; warp k < V sums column k in a fixed order and its lane 0 publishes it.
__device__ __forceinline__ void reduce_publish(const double* s_part, int nact, int V, uint64_t* myrec,
                                               uint32_t tag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < V) {
    double v = 0.0;
    for (int i = lane; i < nact; i += 32) v += s_part[warp * kThreads + i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) ll_store(myrec + warp * 2, v, tag);
  }
}

// One warp gathers value `col` of every CTA's record (lane l takes CTAs l, l+32, ...), optionally
// keeps the individual values in shared memory, and returns their sum (same fixed order in every
// CTA => bitwise identical everywhere).  P <= 32 * kMaxPer.
constexpr int kMaxPer = 5;
__device__ __forceinline__ double poll_column(const uint64_t* buf, int V, int col, uint32_t tag, int P,
                                              double* store, int sstride, int soff) {
  const int lane = threadIdx.x & 31;
  double v[kMaxPer];
  unsigned pending = 0;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    v[j] = 0.0;
    if (lane + 32 * j < P) pending |= 1u << j;
  }
  while (pending) {
    uint64_t w0[kMaxPer], w1[kMaxPer];
#pragma unroll
    for (int j = 0; j < kMaxPer; ++j) {     // all outstanding loads in flight together
      if (pending & (1u << j)) {
        const uint64_t* slot = buf + ((size_t)(lane + 32 * j) * V + col) * 2;
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[j]), "=l"(w1[j]) : "l"(slot));
      }
    }
#pragma unroll
    for (int j = 0; j < kMaxPer; ++j) {
      if ((pending & (1u << j)) && (uint32_t)(w0[j] >> 32) == tag && (uint32_t)(w1[j] >> 32) == tag) {
        v[j] = __longlong_as_double((long long)((w0[j] & 0xffffffffull) | (w1[j] << 32)));
        pending &= ~(1u << j);
      }
    }
  }
  if (store) {
#pragma unroll
    for (int j = 0; j < kMaxPer; ++j)
      if (lane + 32 * j < P) store[(lane + 32 * j) * sstride + soff] = v[j];
  }
  double sum = v[0];
#pragma unroll
  for (int j = 1; j < kMaxPer; ++j) sum += v[j];
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  return sum;
}

template <int R>
__global__ void __launch_bounds__(kThreads, 1) pcg2_kernel(Pcg2Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Pcg2Device& d = a.d;
  const int c = blockIdx.x, P = d.P;
  const int r0 = d.row_off[c], n = d.row_off[c + 1] - r0;
  const int e0 = d.ent_off[c], ne = d.ent_off[c + 1] - e0;
  const int h0 = d.halo_off[c], nh = d.halo_off[c + 1] - h0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nact = min(kThreads, ((min(n, kThreads) + 31) >> 5) << 5);   // threads that own rows

  // ---- shared memory carve-up (sizes from the host: nmax, emax, hmax) ----
  unsigned char* sp = smem_raw;
  auto take = [&](size_t bytes) { unsigned char* p = sp; sp += (bytes + 15) & ~size_t(15); return p; };
  double* s_val = (double*)take(sizeof(double) * d.emax);
  double* s_p = (double*)take(sizeof(double) * 3 * (d.nmax + d.hmax));   // [rhs][own | halo]
  double* s_t = (double*)take(sizeof(double) * 3 * (d.nmax + d.hmax));   // smoother iterate, same layout
  double* s_part = (double*)take(sizeof(double) * kVals * kThreads);     // per-thread partials
  double* s_rc = (double*)take(sizeof(double) * P * 3);                  // coarse residual [P][3]
  double* s_cap = (double*)take(sizeof(double) * P * 3);                 // gathered P^T A p [P][3]
  double* s_ainv = (double*)take(sizeof(double) * P);                    // row c of Ac^-1
  double* s_sc = (double*)take(sizeof(double) * 32);                     // scalars
  int32_t* s_rowptr = (int32_t*)take(sizeof(int32_t) * (d.nmax + 1));
  int32_t* s_halo_src = (int32_t*)take(sizeof(int32_t) * d.hmax);
  uint16_t* s_col = (uint16_t*)take(sizeof(uint16_t) * d.emax);
  const int pstride = d.nmax + d.hmax;

  // ---- load the static slice + gather K values ----
  for (int i = tid; i <= n; i += kThreads) s_rowptr[i] = d.lrowptr[r0 + c + i];
  for (int i = tid; i < ne; i += kThreads) {
    s_col[i] = d.lcol[e0 + i];
    s_val[i] = __ldg(a.Kdata + __ldg(d.lkidx + e0 + i));
  }
  for (int i = tid; i < nh; i += kThreads) s_halo_src[i] = d.halo_src[h0 + i];
  for (int i = tid; i < P; i += kThreads) s_ainv[i] = a.ainv[(size_t)c * P + i];
  for (int i = tid; i < 3 * pstride; i += kThreads) s_p[i] = 0.0;
  for (int i = tid; i < 3 * P; i += kThreads) s_cap[i] = 0.0;

  double x[R][3], r[R][3], w[R][3], dinv[R], diag[R];
  int expslot[R];
  double acc[kVals];
#pragma unroll
  for (int k = 0; k < kVals; ++k) acc[k] = 0.0;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int row = tid + j * kThreads;
    expslot[j] = -1;
    dinv[j] = 0.0; diag[j] = 0.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) { x[j][q] = 0.0; r[j][q] = 0.0; w[j][q] = 0.0; }
    if (row < n) {
      const int gi = d.rows[r0 + row];
      diag[j] = __ldg(a.Kdata + __ldg(d.dkidx + r0 + row));
      dinv[j] = 1.0 / diag[j];
      expslot[j] = d.exp_slot[r0 + row];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        r[j][q] = a.rhs[3 * (size_t)gi + q];
        acc[q] += r[j][q];                 // P^T r (own aggregate)
        acc[3 + q] += r[j][q] * r[j][q];   // <b,b>
      }
      acc[6] += diag[j] > 0.0 ? 1.0 : 0.0;
      acc[7] += diag[j] < 0.0 ? 1.0 : 0.0;
    }
  }
  if (tid < nact) {
#pragma unroll
    for (int k = 0; k < kVals; ++k) s_part[k * kThreads + tid] = acc[k];
  }
  __syncthreads();

  // ---- exchange S: coarse residual of r0 = b, ||b||^2, sign census ----
  reduce_publish(s_part, nact, 8, a.xS + (size_t)c * 8 * 2, 1u);
  if (warp < 8) {
    const double sum = poll_column(a.xS, 8, warp, 1u, P, warp < 3 ? s_rc : nullptr, 3, warp);
    if (lane == 0) s_sc[warp] = sum;          // [3..5] = bb, [6] = npos, [7] = nneg
  }
  __syncthreads();
  const double bb0 = s_sc[3], bb1 = s_sc[4], bb2 = s_sc[5];
  const bool indefinite = s_sc[6] > 0.0 && s_sc[7] > 0.0;
  const double sign = s_sc[7] > 0.0 && s_sc[6] == 0.0 ? -1.0 : 1.0;

  double rz_old[3] = {0.0, 0.0, 0.0}, alpha[3] = {0.0, 0.0, 0.0};
  double rr[3] = {bb0, bb1, bb2};
  int it = 0;
  PROF(long long tB = 0, tA = 0, tH = 0, tC = 0, tM = 0, t0, t1; const long long tstart = clock64();)
  int status = indefinite ? FDM_STATUS_INDEFINITE : FDM_STATUS_OK;
  const bool zero_rhs = (bb0 == 0.0 && bb1 == 0.0 && bb2 == 0.0);

  // a mixed-sign diagonal: K is indefinite, CG has nothing to offer -- report at once (the band L D L^T takes over,
  // solver_ldlt.cu); the flag comes from the all-gathered sums above, so every CTA takes the same branch
  while (!zero_rhs && !indefinite) {
    const uint32_t tag = (uint32_t)it + 1u;
    PROF(t0 = clock64();)
    // ---- coarse residual recurrence + coarse solve for the own aggregate (warps 0..2) ----
    //      rc -= alpha * (P^T A p);  yc = Ainv[c,:] . rc
    if (warp < 3) {
      double v = 0.0;
      for (int j = lane; j < P; j += 32) {
        const double nv = s_rc[j * 3 + warp] - alpha[warp] * s_cap[j * 3 + warp];
        s_rc[j * 3 + warp] = nv;
        v += s_ainv[j] * nv;
      }
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_sc[8 + warp] = v;
    }
    __syncthreads();
    const double yc[3] = {s_sc[8], s_sc[9], s_sc[10]};
    // ---- smoother: zs = p_k(D^-1 K) D^-1 r, Chebyshev polynomial of degree cheb_k on [2/alpha, 2] (the
    //      spectrum of D^-1 K lies in (0, 2) for an all-positive or all-negative q); every extra degree is
    //      one halo hand-over + one SpMV, no grid-wide exchange.  z = zs + yc.
    const double cb = 2.0, ca = 2.0 / a.cheb_alpha;
    const double theta = 0.5 * (cb + ca), delta = 0.5 * (cb - ca), sigma1 = theta / delta, itheta = 1.0 / theta;
    double zs[R][3], dv[R][3];
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
      for (int q = 0; q < 3; ++q) { dv[j][q] = (dinv[j] * itheta) * r[j][q]; zs[j][q] = dv[j][q]; }
    double rho = 1.0 / sigma1;
    for (int sstep = 1; sstep < a.cheb_k; ++sstep) {
      const uint32_t stag = (uint32_t)it * (uint32_t)a.cheb_k + (uint32_t)sstep;
      uint64_t* xEb = a.xE + (size_t)(stag & 1u) * d.total_exports * 6;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int row = tid + j * kThreads;
        if (row < n) {
#pragma unroll
          for (int q = 0; q < 3; ++q) s_t[q * pstride + row] = zs[j][q];
          if (expslot[j] >= 0) {
            uint64_t* e = xEb + (size_t)expslot[j] * 6;
            ll_store(e, zs[j][0], stag); ll_store(e + 2, zs[j][1], stag); ll_store(e + 4, zs[j][2], stag);
          }
        }
      }
      for (int i = tid; i < nh * 3; i += kThreads) {
        const int h = i / 3, q = i - 3 * h;
        s_t[q * pstride + d.nmax + h] = ll_load(xEb + (size_t)s_halo_src[h] * 6 + 2 * q, stag);
      }
      __syncthreads();
      const double rho_new = 1.0 / (2.0 * sigma1 - rho);
      const double c1 = rho_new * rho, c2 = 2.0 * rho_new / delta;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int row = tid + j * kThreads;
        if (row < n) {
          double t0 = diag[j] * s_t[row], t1 = diag[j] * s_t[pstride + row], t2 = diag[j] * s_t[2 * pstride + row];
          const int k0 = s_rowptr[row], k1 = s_rowptr[row + 1];
          for (int k = k0; k < k1; ++k) {
            const double v = s_val[k];
            int cidx = s_col[k];
            cidx = cidx < n ? cidx : d.nmax + (cidx - n);
            t0 += v * s_t[cidx]; t1 += v * s_t[pstride + cidx]; t2 += v * s_t[2 * pstride + cidx];
          }
          const double tt[3] = {t0, t1, t2};
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            dv[j][q] = c1 * dv[j][q] + c2 * dinv[j] * (r[j][q] - tt[q]);
            zs[j][q] += dv[j][q];
          }
        }
      }
      rho = rho_new;
      __syncthreads();                        // s_t is rewritten by the next degree
    }
    // ---- z = zs + yc ; partial <r,z>, <r,r> ; export boundary z ----
    const uint32_t ztag = ((uint32_t)it + 1u) * (uint32_t)a.cheb_k;
    uint64_t* xEz = a.xE + (size_t)(ztag & 1u) * d.total_exports * 6;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] = 0.0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int row = tid + j * kThreads;
      if (row < n) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const double z = zs[j][q] + yc[q];
          w[j][q] = z;
          acc[q] += r[j][q] * z;
          acc[3 + q] += r[j][q] * r[j][q];
        }
        if (expslot[j] >= 0) {
          uint64_t* e = xEz + (size_t)expslot[j] * 6;
          ll_store(e, w[j][0], ztag); ll_store(e + 2, w[j][1], ztag); ll_store(e + 4, w[j][2], ztag);
        }
      }
    }
    if (tid < nact) {
#pragma unroll
      for (int k = 0; k < 6; ++k) s_part[k * kThreads + tid] = acc[k];
    }
    __syncthreads();
    PROF(t1 = clock64(); tC += t1 - t0; t0 = t1;)
    reduce_publish(s_part, nact, 6, a.xB + (size_t)c * 6 * 2, tag);
    if (warp < 6) {
      const double sum = poll_column(a.xB, 6, warp, tag, P, nullptr, 0, 0);
      if (lane == 0) s_sc[12 + warp] = sum;   // [12..14] = <r,z>, [15..17] = <r,r>
    }
    __syncthreads();
    PROF(t1 = clock64(); tB += t1 - t0; t0 = t1;)
    double rz_new[3], beta[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      rz_new[q] = s_sc[12 + q];
      rr[q] = s_sc[15 + q];
      beta[q] = (it > 0 && rz_old[q] != 0.0) ? rz_new[q] / rz_old[q] : 0.0;
    }
    if (rr[0] <= a.tol2 * bb0 && rr[1] <= a.tol2 * bb1 && rr[2] <= a.tol2 * bb2) break;
    if (!(isfinite(rz_new[0]) && isfinite(rz_new[1]) && isfinite(rz_new[2]))) { status = FDM_STATUS_BREAKDOWN; break; }
    if (it >= a.max_iters) { status = FDM_STATUS_NOT_CONVERGED; break; }
    // ---- p = z + beta p on own rows and halo ----
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int row = tid + j * kThreads;
      if (row < n) {
#pragma unroll
        for (int q = 0; q < 3; ++q) s_p[q * pstride + row] = w[j][q] + beta[q] * s_p[q * pstride + row];
      }
    }
    for (int i = tid; i < nh * 3; i += kThreads) {
      const int h = i / 3, q = i - 3 * h;
      const double z = ll_load(xEz + (size_t)s_halo_src[h] * 6 + 2 * q, ztag);
      double* pp = s_p + q * pstride + d.nmax + h;
      *pp = z + beta[q] * (*pp);
    }
    __syncthreads();
    PROF(t1 = clock64(); tH += t1 - t0; t0 = t1;)
    // ---- w = A p (own rows), partial <p,Ap>, own coarse sum of Ap ----
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] = 0.0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int row = tid + j * kThreads;
      if (row < n) {
        const double p0 = s_p[row], p1 = s_p[pstride + row], p2 = s_p[2 * pstride + row];
        double a0 = diag[j] * p0, a1 = diag[j] * p1, a2 = diag[j] * p2;
        const int k0 = s_rowptr[row], k1 = s_rowptr[row + 1];
        for (int k = k0; k < k1; ++k) {
          const double v = s_val[k];
          int cidx = s_col[k];
          cidx = cidx < n ? cidx : d.nmax + (cidx - n);
          a0 += v * s_p[cidx]; a1 += v * s_p[pstride + cidx]; a2 += v * s_p[2 * pstride + cidx];
        }
        w[j][0] = a0; w[j][1] = a1; w[j][2] = a2;
        acc[0] += p0 * a0; acc[1] += p1 * a1; acc[2] += p2 * a2;
        acc[3] += a0; acc[4] += a1; acc[5] += a2;
      }
    }
    if (tid < nact) {
#pragma unroll
      for (int k = 0; k < 6; ++k) s_part[k * kThreads + tid] = acc[k];
    }
    __syncthreads();
    PROF(t1 = clock64(); tM += t1 - t0; t0 = t1;)
    reduce_publish(s_part, nact, 6, a.xA + (size_t)c * 6 * 2, tag);
    if (warp < 6) {
      const double sum = poll_column(a.xA, 6, warp, tag, P, warp >= 3 ? s_cap : nullptr, 3, warp - 3);
      if (lane == 0) s_sc[18 + warp] = sum;   // [18..20] = <p,Ap>
    }
    __syncthreads();
    PROF(t1 = clock64(); tA += t1 - t0; t0 = t1;)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double pap = s_sc[18 + q];
      alpha[q] = pap != 0.0 ? rz_new[q] / pap : 0.0;
      rz_old[q] = rz_new[q];
    }
    // ---- x += alpha p ; r -= alpha A p  (coarse recurrence happens at the top of the loop) ----
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int row = tid + j * kThreads;
      if (row < n) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          x[j][q] += alpha[q] * s_p[q * pstride + row];
          r[j][q] -= alpha[q] * w[j][q];
        }
      }
    }
    ++it;
  }

  // ---- write back ----
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int row = tid + j * kThreads;
    if (row < n) {
      const int gi = d.rows[r0 + row];
#pragma unroll
      for (int q = 0; q < 3; ++q) a.x[3 * (size_t)gi + q] = x[j][q];
    }
  }
  if (c == 0 && tid == 0 && a.info) {
    double rel = 0.0;
    const double bb[3] = {bb0, bb1, bb2};
    for (int q = 0; q < 3; ++q)
      if (bb[q] > 0.0) rel = fmax(rel, sqrt(rr[q] / bb[q]));
    a.info[FDM_INFO_STATUS] = (double)status;
    a.info[FDM_INFO_ITERS] = (double)it;
    a.info[FDM_INFO_RELRES] = rel;
    a.info[FDM_INFO_SIGN] = sign;
    // phase cycle totals of CTA 0 (diagnostics build only): coarse+z, exchange B, p/halo, SpMV, exchange A, total
    PROF(a.info[4] = (double)tC; a.info[5] = (double)tB; a.info[6] = (double)(tH + tM); a.info[7] = (double)tA;
         (void)tstart;)
  }
}

// ------------------------------------------------------------------------------------------
// host: per-part layout
// ------------------------------------------------------------------------------------------
struct Pcg2Host {
  Pcg2Device dev;
  std::vector<void*> allocs;
  size_t xwords = 0;
  int coarse_smem = 0;
};

template <class T>
int up(Pcg2Host* h, const std::vector<T>& v, T** out) {
  *out = nullptr;
  FDM_CUDA(cudaMalloc((void**)out, std::max<size_t>(16, v.size() * sizeof(T))));
  h->allocs.push_back(*out);
  if (!v.empty()) FDM_CUDA(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return FDM_OK;
}

size_t pcg2_smem(int P, int nmax, int emax, int hmax) {
  auto al = [](size_t b) { return (b + 15) & ~size_t(15); };
  return al(8 * (size_t)emax) + 2 * al(8 * 3 * (size_t)(nmax + hmax)) + al(8 * (size_t)kVals * kThreads) +
         al(8 * (size_t)P * 3) * 2 + al(8 * (size_t)P) + al(8 * 32) +
         al(4 * (size_t)(nmax + 1)) + al(4 * (size_t)hmax) + al(2 * (size_t)emax);
}

// Build (once per plan) the per-aggregate local CSR, halo and export maps.
int pcg2_build(fdm_plan* p) {
  if (p->d.pcg2) return FDM_OK;
  const int P = p->num_parts;
  const int32_t n = (int32_t)p->Nf;
  auto* H = new Pcg2Host();
  std::vector<int32_t> row_off(P + 1, 0), loc(n), rows(n);
  for (int32_t i = 0; i < n; ++i) row_off[p->part_of_row[i] + 1]++;
  for (int c = 0; c < P; ++c) row_off[c + 1] += row_off[c];
  {
    std::vector<int32_t> fill(row_off.begin(), row_off.end() - 1);
    for (int32_t i = 0; i < n; ++i) {
      const int c = p->part_of_row[i];
      loc[i] = fill[c] - row_off[c];
      rows[fill[c]++] = i;
    }
  }
  // exports: rows referenced from another aggregate, numbered part by part in local order
  std::vector<int32_t> exp_slot(n, -1), exp_off(P + 1, 0);
  {
    std::vector<char> is_exp(n, 0);
    for (int32_t i = 0; i < n; ++i)
      for (int32_t k = p->index_indptr[i]; k < p->index_indptr[i + 1]; ++k) {
        const int32_t j = p->index_indices[k];
        if (p->part_of_row[j] != p->part_of_row[i]) is_exp[j] = 1;
      }
    int32_t slot = 0;
    for (int c = 0; c < P; ++c) {
      exp_off[c] = slot;
      for (int32_t t = row_off[c]; t < row_off[c + 1]; ++t)
        if (is_exp[rows[t]]) exp_slot[rows[t]] = slot++;
    }
    exp_off[P] = slot;
  }
  std::vector<int32_t> ent_off(P + 1, 0), halo_off(P + 1, 0), lrowptr, lkidx, dkidx(n), halo_src, exp_slot_local(n);
  std::vector<uint16_t> lcol;
  lrowptr.reserve(n + P);
  int nmax = 0, emax = 0, hmax = 0;
  std::map<int32_t, int32_t> halo_of;  // free index -> halo position (per part)
  for (int c = 0; c < P; ++c) {
    const int32_t nc = row_off[c + 1] - row_off[c];
    halo_of.clear();
    // halo columns ordered by export slot (groups them by owner, in owner-local order)
    for (int32_t t = row_off[c]; t < row_off[c + 1]; ++t) {
      const int32_t i = rows[t];
      for (int32_t k = p->index_indptr[i]; k < p->index_indptr[i + 1]; ++k) {
        const int32_t j = p->index_indices[k];
        if (p->part_of_row[j] != c) halo_of[exp_slot[j]] = 0;
      }
    }
    int32_t hpos = 0;
    for (auto& kv : halo_of) {
      kv.second = hpos++;
      halo_src.push_back(kv.first);
    }
    if (nc + hpos >= 65535) {
      delete H;
      fdm_set_error("pcg2: aggregate too large for 16-bit local columns");
      return FDM_ERR_UNSUPPORTED;
    }
    const int32_t ebase = (int32_t)lcol.size();
    for (int32_t t = row_off[c]; t < row_off[c + 1]; ++t) {
      const int32_t i = rows[t];
      lrowptr.push_back((int32_t)lcol.size() - ebase);
      exp_slot_local[t] = exp_slot[i];
      for (int32_t k = p->index_indptr[i]; k < p->index_indptr[i + 1]; ++k) {
        const int32_t j = p->index_indices[k];
        if (j == i) { dkidx[t] = k; continue; }
        lcol.push_back((uint16_t)(p->part_of_row[j] == c ? loc[j] : nc + halo_of[exp_slot[j]]));
        lkidx.push_back(k);
      }
    }
    lrowptr.push_back((int32_t)lcol.size() - ebase);
    ent_off[c + 1] = (int32_t)lcol.size();
    halo_off[c + 1] = (int32_t)halo_src.size();
    nmax = std::max(nmax, (int)nc);
    emax = std::max(emax, (int)(ent_off[c + 1] - ent_off[c]));
    hmax = std::max(hmax, (int)hpos);
  }
  // coarse groups: (a,b) -> K entries with part(row)=a, part(col)=b
  std::map<int64_t, std::vector<int32_t>> groups;
  for (int32_t i = 0; i < n; ++i)
    for (int32_t k = p->index_indptr[i]; k < p->index_indptr[i + 1]; ++k)
      groups[(int64_t)p->part_of_row[i] * P + p->part_of_row[p->index_indices[k]]].push_back(k);
  std::vector<int32_t> cg_ptr(1, 0), cg_ab, cg_kidx;
  for (auto& kv : groups) {
    cg_ab.push_back((int32_t)kv.first);
    cg_kidx.insert(cg_kidx.end(), kv.second.begin(), kv.second.end());
    cg_ptr.push_back((int32_t)cg_kidx.size());
  }

  Pcg2Device& D = H->dev;
  D.P = P; D.nmax = nmax; D.emax = std::max(emax, 1); D.hmax = std::max(hmax, 1);
  D.rows_per_thread = (nmax + kThreads - 1) / kThreads;
  D.smem_bytes = pcg2_smem(P, D.nmax, D.emax, D.hmax);
  D.n_cgroups = (int)cg_ab.size();
  D.total_exports = exp_off[P];
  H->coarse_smem = (int)(sizeof(double) * P * P);
  int rc = FDM_OK;
  if ((rc = up(H, row_off, &D.row_off)) || (rc = up(H, ent_off, &D.ent_off)) ||
      (rc = up(H, halo_off, &D.halo_off)) || (rc = up(H, exp_off, &D.exp_off)) ||
      (rc = up(H, rows, &D.rows)) || (rc = up(H, lrowptr, &D.lrowptr)) || (rc = up(H, lcol, &D.lcol)) ||
      (rc = up(H, lkidx, &D.lkidx)) || (rc = up(H, dkidx, &D.dkidx)) ||
      (rc = up(H, exp_slot_local, &D.exp_slot)) || (rc = up(H, halo_src, &D.halo_src)) ||
      (rc = up(H, cg_ptr, &D.cg_ptr)) || (rc = up(H, cg_ab, &D.cg_ab)) || (rc = up(H, cg_kidx, &D.cg_kidx))) {
    for (void* q : H->allocs) cudaFree(q);
    delete H;
    return rc;
  }
  H->xwords = ((size_t)3 * P * kVals + (size_t)2 * D.total_exports * 3) * 2;   // exports double buffered by parity
  p->d.pcg2 = H;
  return FDM_OK;
}

bool shape_ok(const fdm_plan* p) {
  // P CTAs must be co-resident (one per SM) and the worst aggregate must fit shared memory.
  // Estimate before building: rows/part <= kThreads*kMaxRowsPerThread, dense coarse inverse fits one CTA.
  if (p->num_parts < 2 || p->num_parts > p->sm_count || p->num_parts > 32 * kMaxPer) return false;
  if ((size_t)p->num_parts * p->num_parts * 8 > 200 * 1024) return false;
  return true;
}

size_t ws_layout(const Pcg2Host* H, size_t* ainv_off, size_t* x_off) {
  const size_t ainv = ((size_t)H->dev.P * H->dev.P * 8 + 255) & ~size_t(255);
  *ainv_off = 0;
  *x_off = ainv;
  return ainv + H->xwords * 8 + 256;
}

}  // namespace

// Called once by fdm_plan_create (current device = the plan's): builds and uploads the tables when the mesh
// has the shape the persistent kernel handles.  Failure only means "not supported".
void pcg2_prepare(fdm_plan* p) {
  if (!p->on_device || p->d.pcg2 || !shape_ok(p)) return;
  for (int32_t c : p->part_of_row)
    if (c < 0 || c >= p->num_parts) return;               // every row must have an owner
  if (pcg2_build(p) != FDM_OK) {
    pcg2_destroy(p);
    cudaGetLastError();
  }
}

bool pcg2_supported(const fdm_plan* p) {
  if (!p->on_device || !p->d.pcg2) return false;
  const Pcg2Host* H = (const Pcg2Host*)p->d.pcg2;
  return H->dev.rows_per_thread <= kMaxRowsPerThread && H->dev.smem_bytes <= 227 * 1024 - 1024;
}

size_t pcg2_workspace_bytes(const fdm_plan* p) {
  if (!pcg2_supported(p)) return 0;
  size_t a, x;
  return ws_layout((const Pcg2Host*)p->d.pcg2, &a, &x);
}

void pcg2_destroy(fdm_plan* p) {
  auto* H = (Pcg2Host*)p->d.pcg2;
  if (!H) return;
  for (void* q : H->allocs) cudaFree(q);
  delete H;
  p->d.pcg2 = nullptr;
}

int pcg2_solve(const fdm_plan* p, const double* Kdata, const double* rhs, double* x, double* info,
               void* ws, size_t ws_bytes, int64_t batch, const fdm_solve_opts& o, cudaStream_t st) {
  if (!pcg2_supported(p)) {
    fdm_set_error("pcg2_solve: mesh does not fit the persistent kernel (use FDM_SOLVER_PCG)");
    return FDM_ERR_UNSUPPORTED;
  }
  const Pcg2Host* H = (const Pcg2Host*)p->d.pcg2;
  size_t ainv_off, x_off;
  if (ws_bytes < ws_layout(H, &ainv_off, &x_off)) {
    fdm_set_error("pcg2_solve: workspace too small");
    return FDM_ERR_WORKSPACE;
  }
  const Pcg2Device& D = H->dev;
  const int P = D.P;
  Pcg2Args a;
  a.d = D;
  a.ainv = (double*)((char*)ws + ainv_off);
  a.xS = (uint64_t*)((char*)ws + x_off);
  a.xA = a.xS + (size_t)P * kVals * 2;
  a.xB = a.xA + (size_t)P * kVals * 2;
  a.xE = a.xB + (size_t)P * kVals * 2;
  const double tol = o.tol > 0 ? o.tol : 1e-12;
  a.tol2 = tol * tol;
  a.max_iters = o.max_iters > 0 ? o.max_iters : 20000;
  {
    static const int env_k = std::getenv("FDM_PCG2_CHEB_K") ? std::atoi(std::getenv("FDM_PCG2_CHEB_K")) : 0;
    static const double env_a = std::getenv("FDM_PCG2_CHEB_ALPHA") ? std::atof(std::getenv("FDM_PCG2_CHEB_ALPHA")) : 0.0;
    a.cheb_k = env_k > 0 ? env_k : kChebDegree;
    a.cheb_alpha = env_a > 1.0 ? env_a : kChebAlpha;
  }

  void (*kern)(Pcg2Args) = nullptr;
  switch (D.rows_per_thread) {
    case 1: kern = pcg2_kernel<1>; break;
    case 2: kern = pcg2_kernel<2>; break;
    case 3: case 4: kern = pcg2_kernel<4>; break;
    default: fdm_set_error("pcg2_solve: too many rows per thread"); return FDM_ERR_UNSUPPORTED;
  }
  FDM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.smem_bytes));
  FDM_CUDA(cudaFuncSetAttribute(pcg2_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H->coarse_smem));
  for (int64_t b = 0; b < batch; ++b) {
    a.Kdata = Kdata + b * p->nnz;
    a.rhs = rhs + b * p->Nf * 3;
    a.x = x + b * p->Nf * 3;
    a.info = info ? info + b * FDM_INFO_STRIDE : nullptr;
    // the coarse inverse in the workspace belongs to ONE K: it is reused only for a single design
    if (!o.reuse_factor || batch > 1) {
      // coarse inverse + clear exchange tags (CTA 0: coarse; CTAs 1..: clear)
      pcg2_setup_kernel<<<1 + 16, 1024, H->coarse_smem, st>>>(a, H->xwords);
    } else {
      FDM_CUDA(cudaMemsetAsync(a.xS, 0, H->xwords * 8, st));
    }
    void* args[] = {&a};
    FDM_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(P), dim3(kThreads), args, D.smem_bytes, st));
  }
  return fdm_check_launch("pcg2_solve");
}
