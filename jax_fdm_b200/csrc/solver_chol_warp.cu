// Batched band Cholesky, ONE WARP PER DESIGN (half bandwidth <= 31), many designs per SM.
// Replaces the vmap(EquilibriumModel) dense solve (models.py:88, loss_plotter.py:98-103) for
// batches of small designs that share one topology, and the adjoint solve of
// sparse_solve_bwd (sparse.py:296) with the factor reused instead of recomputed.
//
// Kernels:
//
// chol_factor_mma_kernel<FROM_Q, PW, SIGNED>  (half bandwidth <= 30) factorisation + fused forward substitution
//   on the FP64 tensor pipe, see the comment above the kernel: the trailing window in registers as 8x8 tiles,
//   PW = 4 or 2 columns per panel, the right-hand sides as extra window rows (PW = 4), the diagonal 8x8 blocks
//   of every 32-column block inverted in place; FROM_Q assembles its rows from q (fdm_solve_from_q); SIGNED is
//   the second pass over designs with a non-positive pivot (sigma K = L D L^T).
//
// chol_factor_kernel<KMAX>  (half bandwidth 31, and the A/B reference) the same factorisation with rank-1 updates.
//   Band order = plan.band_perm.
//   Lane (R & 31) owns row R of the lower triangle while R is inside the 32-row window [j, j+32)
//   of the right-looking factorisation; the row lives in 32 REGISTERS, slot (C & 31) holding
//   A(R, C).  Step j:
//       piv  = A(j,j)                    shuffle from lane j&31
//       l    = A(R,j) / sqrt(piv)        every lane, own register (exactly 0 outside the band)
//       A(R, j+k) -= l_R * l_{j+k}       k = 1..KMAX: l_{j+k} is a shared-memory BROADCAST, the
//                                        accumulator never leaves registers
//   The updates with j+k > R fall into slots of columns that are already eliminated (dead) and
//   are overwritten when the lane takes row R+32, whose 32 slots are read from a staged
//   shared-memory block (filled once per 32 steps from K.data through plan.band_rowkidx).  The
//   three right-hand sides ride along as three more registers per lane.
//   The factor goes to the workspace column by column (fac[j][d] = L(j+d, j), d = 1..31), one
//   coalesced 256-byte store per step.  Columns are grouped in eights; the 8x8 diagonal block of
//   every group is stored INVERTED (fac[8k+b][a-b] = inv(L_kk)(a,b), hence fac[j][0] = 1/L(j,j)).
//   Inside its 32-double record, column j is rotated by 5*(j&3) slots (FROT): a 32-column block is
//   then one contiguous 8 KB that a single TMA bulk copy drops into shared memory, and the tile
//   fragment reads of the substitution kernel (4 columns x 8 rows per access) hit 16 distinct banks.
//
// chol_subst_kernel   substitutions with the stored factor, as dependency-free 8x8 tile products
//       forward   z_k = inv(L_kk) g_k ;  g_{k+i} -= L_{k+i,k} z_k           (i = 1..NBL)
//       backward  x_k = inv(L_kk)^T (z_k - sum_i L_{k+i,k}^T x_{k+i})
//   on the FP64 tensor pipe (mma.sync.m8n8k4.f64, SASS DMMA); the three right-hand sides are
//   columns 0..2 of an 8-wide tile.  The factor streams through shared memory in 32-column blocks,
//   double buffered: one 8 KB TMA bulk copy (cp.async.bulk, SASS UBLKCP) per block, completion on
//   an mbarrier, so the next block is in flight while this one is used.
//   The forward solve runs it backward-only (its forward substitution is fused in the factor
//   kernel); the adjoint solve (opts.reuse_factor) runs forward then backward.
//
// Rows are padded to a multiple of 32 with identity rows, so there is no tail code.
// sigma = sign of the first diagonal entry; the factor is of sigma*K (all q<0 => K = -SPD).
#include <cuda_runtime.h>

#include <cstdlib>

#include "launch.h"

namespace {

constexpr int kFS = 34;    // row stride (doubles) of a staged 32x32 block: 16-byte aligned rows, conflict-free columns
constexpr unsigned kFull = 0xffffffffu;
// slot of L(j+d, j) inside the 32-double record of factor column j
#define FROT(j, d) (((d) + 5 * ((j) & 3)) & 31)

// tables of one half of the two-sided factorisation (plan.h: BandHalf); the one-sided kernels use h[0] alone
struct WHalf {
  const int32_t *iperm, *stage_ptr, *stage_ent, *row_diag, *row_sup, *row_node;
  int nblk;        // blocks this half eliminates
  int rows;        // rows covered by iperm (one-sided: n; two-sided: every staged row, -1 = padding)
  int block0;      // first factor block of this half inside the design's workspace
};

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
  int forward;             // subst kernel: run the forward substitution on `rhs` first
  int nan_fail;            // subst kernel: a design whose factorisation failed gets x = NaN
  // in-place assembly (chol_factor_mma_kernel<true>): K and the load term straight from q
  const double* q;         // (B, E)
  const double* xs;        // (B|1, Ns, 3)
  const double* loads;     // (B|1, N, 3)
  int64_t q_stride, xs_stride, loads_stride;
  const int32_t *diags_ptr, *diags_idx, *free_node, *node_slot, *ninc_ptr, *ninc_edge, *ninc_other;
  const uint32_t* sup_mask;
  const int32_t *stage_ptr, *stage_ent;   // plan.band_stage_ptr and band_stage_kent (from K.data) or band_stage_eent (from q)
  const int32_t *row_diag, *row_sup, *row_node;   // plan.band_row_* (assembly in place)
  WHalf h[2];
  int shared_factor;       // substitution kernel: one factor (workspace slot 0) for every right-hand-side set of the batch
  int sub_pair;            // substitution kernel: a warp pair per design (two-sided factor): T and B sweeps in parallel
  int twisted;             // two warps per design: warp 2k factors [T | M] top-down, warp 2k+1 factors B bottom-up
  int sigma_f, sigma_k;    // free index / position in K.data of the first diagonal entry (the sign of K)
  int num_edges, num_nodes, num_fixed;   // bounds of the tables' targets (checked by the -DFDM_DEBUG build)
};

// 1/sqrt(x) for a positive normal double: hardware seed (MUFU.RSQ64H, ~20 bits) + ONE third-order step
// y (1 + e/2 + 3e^2/8), e = 1 - x y^2 (error e^3: below double rounding) -- four dependent FP64 operations
// on the pivot chain instead of the six of two Newton steps.  No special-case path: a non-positive pivot is
// reported through `bad`.
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
}

// -------------------------------------------------------------------------------------------
// After the factorisation: replace the 8x8 lower-triangular diagonal block of every group of 8
// columns by its inverse, in place (the diagonal slots already hold 1/L(j,j)).  Eight lanes per
// block: lane c solves L X(:,c) = e_c by forward substitution with the block in registers.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chol_invdiag_kernel(WArgs A) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int groups = A.nblk * 4;
  const int64_t item = t >> 3;            // (design, group)
  const int c = (int)(t & 7), lane = threadIdx.x & 31, base = lane & ~7;
  const bool live = item < A.batch * groups;
  double* fb = nullptr;
  // lane c reads the block entries of ITS column once (a contiguous run of 8 - c doubles inside the
  // rotated record: one or two 32-byte sectors); the others get them by shuffle
  double col[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) col[d] = 0.0;
  if (live) {
    const int64_t b = item / groups;
    const int grp = (int)(item - b * groups);
    fb = A.ws + b * A.ws_stride + (int64_t)grp * 8 * 32;
#pragma unroll
    for (int d = 0; d < 8; ++d)
      if (d < 8 - c) col[d] = fb[c * 32 + FROT(c, d)];      // L(c + d, c); 1/L(c,c) at d = 0
  }
  double Lr[8][8];                         // Lr[r][k] = L(r, k), k <= r
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int r = k; r < 8; ++r) Lr[r][k] = __shfl_sync(0xffffffffu, col[r - k], base + k);
  if (!live) return;
  double X[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    double sacc = 0.0;
#pragma unroll
    for (int k = 0; k < r; ++k) sacc = fma(Lr[r][k], X[k], sacc);   // X[k] = 0 for k < c
    X[r] = Lr[r][r] * ((r == c ? 1.0 : 0.0) - sacc);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if (r >= c) fb[c * 32 + FROT(c, r - c)] = X[r];
}

// ===========================================================================================
// factorisation + fused forward substitution
// ===========================================================================================
constexpr int kFacWarps = 8;  // designs per CTA
constexpr int kFacThreads = kFacWarps * 32;
constexpr int kSL = 40;       // doubles per broadcast buffer: [1..32] l, [34..36] solved rhs entries
constexpr int kFacWarpSmem = 32 * kFS + 2 * kSL;       // staged rows + 2 broadcast buffers

template <int KMAX>
__global__ void __launch_bounds__(kFacThreads, 2) chol_factor_kernel(WArgs A) {
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * kFacWarps + warp;
  if (b >= A.batch) return;  // warps are independent: no block-level barrier anywhere below
  double* fs = smem + warp * kFacWarpSmem;
  double* sl = fs + 32 * kFS;
  const int n = A.n, nblk = A.nblk;
  const double* Kd = A.Kdata + b * A.nnz;
  const double* rhs = A.rhs + b * (int64_t)n * 3;
  double* fac = A.ws + b * A.ws_stride;
  double* ysol = fac + (int64_t)nblk * 1024;
  double* meta = ysol + (int64_t)nblk * 128;
  int bad = 0;
  const double sigma = __ldg(Kd + __ldg(A.rowkidx)) < 0.0 ? -1.0 : 1.0;

  auto load_rhs = [&](int R, double (&out)[3]) {
    if (R < n) {
      const int gi = __ldg(A.iperm + R);
      out[0] = sigma * rhs[3 * gi];
      out[1] = sigma * rhs[3 * gi + 1];
      out[2] = sigma * rhs[3 * gi + 2];
    } else {
      out[0] = out[1] = out[2] = 0.0;
    }
  };
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
  // per-lane constants of the unrolled step body (everything else is an immediate)
  const double* const fsrow0 = fs;                   // staged rows, indexed by the compile-time step
  for (int jb = 0; jb < nblk; ++jb) {
    const int j0 = jb * 32;
    double* const fb = fac + (int64_t)j0 * 32;       // factor records of this block's 32 columns
    double* const yb = ysol + (int64_t)j0 * 4;
    __syncwarp();  // every lane has consumed the previous staged block
    stage_rows(j0 + 32);
    load_rhs(j0 + 32 + lane, yn);
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 32; ++p) {
      const int delta = (lane - p) & 31;
      double* buf = sl + (p & 1) * kSL;
      const double piv = __shfl_sync(kFull, a[p], p);
      bad |= !(piv > 0.0);
      const double inv = rsqrt(piv);
      const double l = a[p] * inv;
      // factor record of column j (rotated, see FROT), 1/L(j,j) in the diagonal slot; the 8x8 diagonal
      // blocks are inverted afterwards by chol_invdiag_kernel
      fb[p * 32 + ((lane + ((5 * (p & 3) - p) & 31)) & 31)] = delta == 0 ? inv : l;
      buf[delta + 1] = l;
      if (delta == 0) {
        const double y0 = y[0] * inv, y1 = y[1] * inv, y2 = y[2] * inv;
        buf[34] = y0; buf[35] = y1; buf[36] = y2;
        yb[p * 4] = y0; yb[p * 4 + 1] = y1; yb[p * 4 + 2] = y2;
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
          const double2 v = *reinterpret_cast<const double2*>(fsrow0 + p * kFS + s);
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
    meta[2] = 0.0;
    meta[3] = 0.0;
  }
}

// ===========================================================================================
// factorisation + fused forward substitution, FP64 tensor-core trailing updates (hbw <= 30)
//
// The 32x32 trailing window (rows / columns j..j+31, slot = index & 31, symmetric) lives in REGISTERS as
// tiles of 8x8 in the accumulator layout of mma.sync.m8n8k4.f64: lane (g = lane>>2, m = lane&3) holds rows
// 8tr+g, columns 8tc+2m and 8tc+2m+1 of tile (tr, tc).
//
// PW = 4 (half bandwidth <= 29), four columns per panel; only the tiles on and below the diagonal are kept and
// the three right-hand sides are a fifth tile row Y (rows 32..34):
//   1. the four panel columns (and the right-hand sides' entries in them) go through the shared-memory exchange
//      to a lanes = rows layout; rows of tile rows above the panel's come from the transposed tiles
//   2. every lane reads the 4x4 pivot block from the exchange and factors it redundantly (four rsqrt, no
//      shuffle on the chain), then computes its own row of the four columns of L            [FP64 FMA pipe]
//   3. lanes (g < 3, m) run the four pivot columns on right-hand side g: w_m is lane (g, m)'s mma A fragment
//      and the solved entry of column j+m (stored once per panel)
//   4. the factor records (rotated, see FROT)
//   5. rows / columns j+32..j+35 REFILL the four retired slots from the staged block
//   6. the panel P (incoming rows: zeros, except L(j+32, j+3) for row j+32) goes back through the exchange as
//      A/B fragments: W -= P P^T as 10 DMMAs, Y -= w P^T as 4 more (all four k slices used) [FP64 tensor pipe]
// After the 8 panels of a 32-column block the block's four 8x8 diagonal blocks are inverted in place.
//
// PW = 2 (half bandwidth 30), two columns per panel, full window, right-hand sides in per-lane registers:
//   pivots by shuffle, forward substitution by shuffle, update (k = 2 of 4 used, 16 DMMAs), then refill.
// Out-of-band entries of the window stay exactly zero (an update needs both factors non-zero).
// ===========================================================================================
constexpr int kMS = 36;                                  // staged row: 32 column slots + 3 right-hand sides + pad
// panel exchange of the four-column kernel: entry (row, column c) at h * kPH + 2 * (row ^ h) + (c & 1), h = c >> 1:
// a lane's 16-byte accesses are contiguous across the warp, and the 16-byte column writes, the 8-byte transposed
// writes and the fragment reads are all free of bank conflicts
constexpr int kPH = 72;
constexpr int kMmaWarpSmem = 32 * kMS + 160;             // staged rows + panel exchange (40 rows x up to 4 columns)

__device__ __forceinline__ void dmma_f(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// The 8x8 diagonal blocks of one 32-column block inverted in place by the warp that has just written the
// block's records (same arithmetic as chol_invdiag_kernel: lane c of every 8-lane group solves
// L X(:,c) = e_c; the entries of L come from the lane that owns their column, by shuffle).
__device__ __forceinline__ void invdiag_load(const double* fb, int lane, double (&col)[8]) {
  const int c = lane & 7;
  const double* gb = fb + (lane >> 3) * 256;
#pragma unroll
  for (int d = 0; d < 8; ++d) col[d] = d < 8 - c ? gb[c * 32 + FROT(c, d)] : 0.0;   // L(c + d, c); 1/L(c,c) at d = 0
}
__device__ __forceinline__ void invdiag_solve_store(double* fb, int lane, const double (&col)[8]) {
  const int c = lane & 7, base = lane & ~7;
  double* gb = fb + (lane >> 3) * 256;
  double X[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    double sacc = 0.0;
#pragma unroll
    for (int k = 0; k < r; ++k) sacc = fma(__shfl_sync(kFull, col[r - k], base + k), X[k], sacc);   // X[k] = 0 for k < c
    X[r] = __shfl_sync(kFull, col[0], base + r) * ((r == c ? 1.0 : 0.0) - sacc);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if (r >= c) gb[c * 32 + FROT(c, r - c)] = X[r];
}

// SIGNED: second pass over the designs the plain factorisation flagged (a pivot of sigma*K was not positive:
// mixed-sign q).  Same kernel with signed pivots, sigma*K = L D L^T, D = diag(+-1), no pivoting: column k is
// L(:,k) = A(:,k) / (d_k sqrt|piv_k|), the update W -= d_k l_k l_k^T, and the forward-substituted right-hand
// sides are stored multiplied by D so that the backward substitution runs unchanged.  The signs go to the
// fourth slot of the intermediate right-hand sides (the adjoint's forward sweep needs them), meta[2] = 1.
// A pivot below 1e-13 of the largest diagonal entry seen is reported as a breakdown (meta[1] = 2).
// measurement builds only: -DFDM_FAC_MAXNREG=112 trades the launch bound for a register cap (18 warps per SM with
// FDM_CHOL_FACWARPS=6); DESIGN.md §5 has the outcome
#ifdef FDM_FAC_MAXNREG
#define FDM_FAC_BOUNDS __maxnreg__(FDM_FAC_MAXNREG)
#else
#define FDM_FAC_BOUNDS __launch_bounds__(kFacThreads, 2)
#endif
template <bool FROM_Q, int PW, bool SIGNED, bool TW = false>   // TW: two-sided, a warp pair per design
__global__ void FDM_FAC_BOUNDS chol_factor_mma_kernel(WArgs A) {
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // designs per CTA = warps per CTA (host's choice); two-sided: a design is the warp pair (2k, 2k+1)
  const int half = TW ? (warp & 1) : 0;
  const int64_t b = TW ? (int64_t)blockIdx.x * (blockDim.x >> 6) + (warp >> 1) : (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= A.batch) return;
  const WHalf H = TW ? A.h[half]
                            : WHalf{A.iperm, A.stage_ptr, A.stage_ent, A.row_diag, A.row_sup, A.row_node, A.nblk, A.n, 0};
  double* fs = smem + warp * kMmaWarpSmem;
  double* Pn = fs + 32 * kMS;
  const int g = lane >> 2, m = lane & 3;
  const int n = A.n, nblk = H.nblk;
  const double* Kd = FROM_Q ? nullptr : A.Kdata + b * A.nnz;
  const double* rhs = FROM_Q ? nullptr : A.rhs + b * (int64_t)n * 3;
  const double* q = FROM_Q ? A.q + b * A.q_stride : nullptr;
  double* const ws0 = A.ws + b * A.ws_stride;
  double* fac = ws0 + (int64_t)H.block0 * 1024;
  double* ysol = ws0 + (int64_t)A.nblk * 1024 + (int64_t)H.block0 * 128;
  double* meta = ws0 + (int64_t)A.nblk * (1024 + 128);
  int bad = 0;
  if (SIGNED && meta[1] == 0.0) return;                    // the plain factorisation succeeded for this design
  double dmax = 0.0;                                       // SIGNED: largest |diagonal entry| staged so far

  // K(i,i) = sum of q over the incident edges of free row f, in ascending edge id (the order of the
  // reference's `diags @ q`, structures.py:316-350 -- same additions as assemble_kernel)
  auto diag_of = [&](int f) {
    double dsum = 0.0;
    const int k0 = __ldg(A.diags_ptr + f), k1 = __ldg(A.diags_ptr + f + 1);
    for (int k = k0; k < k1; ++k) dsum += __ldg(q + __ldg(A.diags_idx + k));
    return dsum;
  };
  double sigma;
  if (FROM_Q) sigma = diag_of(A.sigma_f) < 0.0 ? -1.0 : 1.0;
  else sigma = __ldg(Kd + A.sigma_k) < 0.0 ? -1.0 : 1.0;

  // fs[r][s] = sigma * A(R0 + r, C), C = the column with slot s among the 32 ending at row R0 + r;
  // fs[r][32..34] = sigma * rhs(R0 + r).  K is sparse inside the band: the plan lists the non-zero (r, s) of
  // every block (band_stage_*), a few per lane; fs is all zero between blocks (the previous block's entries
  // are taken back first), so a block costs one round of index loads and one of value loads.
  const int2* ent = reinterpret_cast<const int2*>(H.stage_ent);
  auto stage_rows = [&](int blk) {
    const int R0 = blk * 32;
    // loads are issued in the order that keeps two chains in flight side by side: entry lists -> K.data,
    // band-to-free index -> right-hand side
    const int g0 = __ldg(H.stage_ptr + blk), g1 = __ldg(H.stage_ptr + blk + 1);
    int gi = -1;
    if (R0 + lane < H.rows) gi = __ldg(H.iperm + R0 + lane);
    FDM_ASSERT(gi >= -1 && gi < n, "iperm[%d] = %d, n = %d, half %d", R0 + lane, gi, n, half);
    FDM_ASSERT(g0 <= g1 && g1 - g0 <= 64, "stage groups [%d, %d) of block %d, half %d", g0, g1, blk, half);
    int2 e[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) e[u] = g0 + u < g1 ? __ldg(ent + (g0 + u) * 32 + lane) : make_int2(0, -4);
    if (blk > 0) {                                           // take back the previous block's entries
      const int h0 = __ldg(H.stage_ptr + blk - 1);
      for (int gq = h0; gq < g0; gq += 4) {
        int2 o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) o[u] = gq + u < g0 ? __ldg(ent + (gq + u) * 32 + lane) : make_int2(0, -4);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (o[u].y != -4) fs[o[u].x + (o[u].x >> 5) * (kMS - 32)] = 0.0;
      }
      __syncwarp();
    }
    double rr0 = 0.0, rr1 = 0.0, rr2 = 0.0;
    if (!FROM_Q && gi >= 0) { rr0 = rhs[3 * gi]; rr1 = rhs[3 * gi + 1]; rr2 = rhs[3 * gi + 2]; }
    for (int gq = g0; gq < g1; gq += 4) {
      double v[4];
      if (gq > g0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) e[u] = gq + u < g1 ? __ldg(ent + (gq + u) * 32 + lane) : make_int2(0, -4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = e[u].y == -2 ? 1.0 : 0.0;
        FDM_ASSERT(e[u].y == -4 || (e[u].x >= 0 && e[u].x < 1024), "stage slot %d (block %d, half %d)", e[u].x, blk, half);
        FDM_ASSERT(e[u].y < (FROM_Q ? A.num_edges : (int)A.nnz), "stage index %d (block %d, half %d)", e[u].y, blk, half);
        if (e[u].y >= 0) v[u] = FROM_Q ? -sigma * __ldg(q + e[u].y) : sigma * __ldg(Kd + e[u].y);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (e[u].y != -4) fs[e[u].x + (e[u].x >> 5) * (kMS - 32)] = v[u];
    }
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
    if (FROM_Q) __syncwarp();                                // the list wrote a zero on the diagonals, see below
    if (!FROM_Q) {
      r0 = sigma * rr0; r1 = sigma * rr1; r2 = sigma * rr2;
    } else if (gi >= 0) {
      // assembly in place (models.py:1064-1079, 828-856, 949-976): the off-diagonals are -q[e] (above); lane r
      // finishes row R0 + r: its diagonal and its load term loads_f + sum_{e to support s} q_e x_s.  The row's
      // edge ids / support pairs / node id come as padded vectors (plan.band_row_*): one round of index loads,
      // one of value loads; the sums run in the order of the assembly kernel (same bits).
      const int f = gi;
      const int R = R0 + lane;
      const int4 da = __ldg(reinterpret_cast<const int4*>(H.row_diag) + 2 * R);
      const int4 db = __ldg(reinterpret_cast<const int4*>(H.row_diag) + 2 * R + 1);
      const int4 sa = __ldg(reinterpret_cast<const int4*>(H.row_sup) + 2 * R);
      const int4 sb = __ldg(reinterpret_cast<const int4*>(H.row_sup) + 2 * R + 1);
      const int node = __ldg(H.row_node + R);
      FDM_ASSERT(node >= 0 && node < A.num_nodes, "row_node[%d] = %d (half %d)", R, node, half);
      const double* lp = A.loads + b * A.loads_stride + 3 * (int64_t)node;
      const double p0 = __ldg(lp), p1 = __ldg(lp + 1), p2 = __ldg(lp + 2);
      double dsum;
      if (da.x == -2) {
        dsum = diag_of(f);
      } else {
        const int de[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
        double qd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) qd[j] = de[j] >= 0 ? __ldg(q + de[j]) : 0.0;
        dsum = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (de[j] >= 0) dsum += qd[j];
      }
      fs[lane * kMS + lane] = sigma * dsum;                  // slot of column R0 + lane is lane
      double3 acc = make_double3(0.0, 0.0, 0.0);
      const double* xsb = A.xs + b * A.xs_stride;
      if (sa.x == -2) {
        const int i0 = __ldg(A.ninc_ptr + node), i1 = __ldg(A.ninc_ptr + node + 1);
        for (int k = i0; k < i1; ++k) {
          const int slot = __ldg(A.node_slot + __ldg(A.ninc_other + k));
          if (slot < 0) {
            const double qe = __ldg(q + (__ldg(A.ninc_edge + k) >> 1));
            const double* xp = xsb + 3 * (int64_t)(-slot - 1);
            acc.x += qe * __ldg(xp); acc.y += qe * __ldg(xp + 1); acc.z += qe * __ldg(xp + 2);
          }
        }
      } else if (sa.x >= 0) {
        const int se[4] = {sa.x, sa.z, sb.x, sb.z}, ss[4] = {sa.y, sa.w, sb.y, sb.w};
        double qe[4], x0[4], x1[4], x2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool on = se[j] >= 0;
          const double* xp = xsb + 3 * (int64_t)(on ? ss[j] : 0);
          qe[j] = on ? __ldg(q + se[j]) : 0.0;
          x0[j] = on ? __ldg(xp) : 0.0; x1[j] = on ? __ldg(xp + 1) : 0.0; x2[j] = on ? __ldg(xp + 2) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (se[j] >= 0) { acc.x += qe[j] * x0[j]; acc.y += qe[j] * x1[j]; acc.z += qe[j] * x2[j]; }
      }
      r0 = sigma * (p0 + acc.x); r1 = sigma * (p1 + acc.y); r2 = sigma * (p2 + acc.z);
    }
    fs[lane * kMS + 32] = r0; fs[lane * kMS + 33] = r1; fs[lane * kMS + 34] = r2;
  };

#pragma unroll 8
  for (int r = 0; r < 32; ++r) fs[r * kMS + lane] = 0.0;
  __syncwarp();
  auto track_dmax = [&]() {                                // after a __syncwarp that follows stage_rows
    double dm = fabs(fs[lane * kMS + lane]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(kFull, dm, o));
    dmax = fmax(dmax, dm);
  };
  double W[4][4][2], y[3];
  stage_rows(0);
  __syncwarp();
  if (SIGNED) track_dmax();
#pragma unroll
  for (int tr = 0; tr < 4; ++tr)
#pragma unroll
    for (int tc = 0; tc < 4; ++tc)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int R = 8 * tr + g, C = 8 * tc + 2 * m + e;
        W[tr][tc][e] = C <= R ? fs[R * kMS + C] : fs[C * kMS + R];
      }
  y[0] = fs[lane * kMS + 32]; y[1] = fs[lane * kMS + 33]; y[2] = fs[lane * kMS + 34];   // two-column panels
  double Y[4][2];                                                                         // four-column panels
#pragma unroll
  for (int tc = 0; tc < 4; ++tc)
#pragma unroll
    for (int e = 0; e < 2; ++e) Y[tc][e] = g < 3 ? fs[(8 * tc + 2 * m + e) * kMS + 32 + g] : 0.0;

  for (int jb = 0; jb < nblk; ++jb) {
    const int j0 = jb * 32;
    double* const fb = fac + (int64_t)j0 * 32;
    double* const yb = ysol + (int64_t)j0 * 4;
    FDM_ASSERT(H.block0 + jb < A.nblk, "factor block %d + %d of %d (half %d)", H.block0, jb, A.nblk, half);
    if constexpr (PW == 4 && TW) {
      if (half == 0 && jb == nblk - 1) {
        // Two-sided: the window now holds the middle block M with every update from T.  The partner warp has
        // left B's updates of M in its staging area, in ITS (reversed) numbering: entry (r, c) here is
        // (31 - r, 31 - c) there; right-hand side column c is 31 - c.  Add them; M is then factored like any block.
        asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp >> 1)) : "memory");
        const double* fo = fs + kMmaWarpSmem;            // warp + 1's staging area
#pragma unroll
        for (int tr = 0; tr < 4; ++tr)
#pragma unroll
          for (int tc = 0; tc <= tr; ++tc)
#pragma unroll
            for (int e = 0; e < 2; ++e) W[tr][tc][e] += fo[(31 - (8 * tr + g)) * kMS + 31 - (8 * tc + 2 * m + e)];
        if (g < 3) {
#pragma unroll
          for (int tc = 0; tc < 4; ++tc)
#pragma unroll
            for (int e = 0; e < 2; ++e) Y[tc][e] += fo[(31 - (8 * tc + 2 * m + e)) * kMS + 32 + g];
        }
        const int obad = (int)fo[32 * kMS];              // the partner's pivot flag
        if (SIGNED) bad = obad == 2 ? 2 : bad; else bad |= obad;
      }
    }
    __syncwarp();
    stage_rows(jb + 1);
    __syncwarp();
    if (SIGNED) track_dmax();
    if constexpr (PW == 4) {
      // Four columns per panel (hbw <= 29): all four k-slices of the DMMA carry data.  Column j+3 can reach
      // row j+32, one row past the window: that entry has no fill before it (rows j+32 and j..j+2 are more
      // than hbw apart), so it is x3 = A(j+32, j+3) / L(j+3, j+3) straight from the staged block.  The window
      // is REFILLED FIRST and the update runs afterwards with the panel rows of the four incoming rows set to
      // (0, 0, 0, x3), (0..), (0..), (0..): row j+32 receives its update from column j+3 with everyone else.
#pragma unroll
      for (int p4 = 0; p4 < 8; ++p4) {
        const int s0 = 4 * p4;            // slots (= columns of this block) of the panel: s0 .. s0 + 3
        const int tc0 = p4 >> 1, mh = p4 & 1;
        const bool collane = (m >> 1) == mh;               // holds panel columns s0 + 2 (m & 1), + 1 of every tile row
        const bool rowlane = (g >> 2) == mh;               // holds panel row 8 tc0 + g of tile row tc0
        // ---- 1. panel columns -> lanes = rows
        __syncwarp();                                      // the previous panel's fragment reads of Pn are done
        // (only the tiles on and below the diagonal of the symmetric window are kept: rows of tile rows above
        //  tc0 come from the transposed tiles W[tc0][tr], held by the lanes that own the panel columns as ROWS)
        if (collane) {
#pragma unroll
          for (int tr = tc0; tr < 4; ++tr)
            *reinterpret_cast<double2*>(Pn + (m & 1) * kPH + 2 * ((8 * tr + g) ^ (m & 1))) = make_double2(W[tr][tc0][0], W[tr][tc0][1]);
        }
        if (collane && g < 3)                              // rows 32..34 of the exchange: the three right-hand sides
          *reinterpret_cast<double2*>(Pn + (m & 1) * kPH + 2 * ((32 + g) ^ (m & 1))) = make_double2(Y[tc0][0], Y[tc0][1]);
        if (rowlane) {
#pragma unroll
          for (int tr = 0; tr < tc0; ++tr) {
            const int h = (g >> 1) & 1;
            Pn[h * kPH + 2 * ((8 * tr + 2 * m) ^ h) + (g & 1)] = W[tc0][tr][0];
            Pn[h * kPH + 2 * ((8 * tr + 2 * m + 1) ^ h) + (g & 1)] = W[tc0][tr][1];
          }
        }
        __syncwarp();
        const double2 va = *reinterpret_cast<const double2*>(Pn + 2 * lane);
        const double2 vb = *reinterpret_cast<const double2*>(Pn + kPH + 2 * (lane ^ 1));
        const int d0 = (lane - s0) & 31;                 // row of this lane = j0 + s0 + d0
        // signed pivots (SIGNED): d_k = sign(piv_k), l_k = v_k / (d_k sqrt|piv_k|), update with d_k l_k l_k^T
        // The 4x4 pivot block is read by every lane (broadcast reads of rows s0..s0+3 of the exchange) and
        // factored redundantly: no shuffle on the pivot chain; each operation is the one the lane that owns the
        // row performs below, so the pivots and the block's entries of L are the same bits in every lane.
        const double thr = SIGNED ? 1e-13 * dmax : 0.0;
        const double* B0 = Pn + 2 * s0;                          // columns 0, 1 of rows s0 + k at B0[2k], B0[2k+1]
        const double* B1 = Pn + kPH;                             // columns 2, 3 of row r at B1[2 (r ^ 1)], + 1
        const double piv0 = B0[0];
        const double sg0 = SIGNED && piv0 < 0.0 ? -1.0 : 1.0;
        const double inv0 = rsqrt_nr(SIGNED ? fabs(piv0) : piv0);
        const double i0s = SIGNED ? sg0 * inv0 : inv0;
        const double L10 = B0[2] * i0s, L20 = B0[4] * i0s, L30 = B0[6] * i0s;
        const double l0 = va.x * i0s;
        const double m0s = SIGNED ? -sg0 * l0 : -l0;
        const double n10 = SIGNED ? -sg0 * L10 : -L10, n20 = SIGNED ? -sg0 * L20 : -L20, n30 = SIGNED ? -sg0 * L30 : -L30;
        double v1 = fma(m0s, L10, va.y);
        double v2 = fma(m0s, L20, vb.x);
        double v3 = fma(m0s, L30, vb.y);
        const double piv1 = fma(n10, L10, B0[3]);
        const double sg1 = SIGNED && piv1 < 0.0 ? -1.0 : 1.0;
        const double inv1 = rsqrt_nr(SIGNED ? fabs(piv1) : piv1);
        const double i1s = SIGNED ? sg1 * inv1 : inv1;
        const double L21 = fma(n20, L10, B0[5]) * i1s, L31 = fma(n30, L10, B0[7]) * i1s;
        const double l1 = v1 * i1s;
        const double m1s = SIGNED ? -sg1 * l1 : -l1;
        const double n21 = SIGNED ? -sg1 * L21 : -L21, n31 = SIGNED ? -sg1 * L31 : -L31;
        v2 = fma(m1s, L21, v2);
        v3 = fma(m1s, L31, v3);
        const double2 c2 = *reinterpret_cast<const double2*>(B1 + 2 * ((s0 + 2) ^ 1));   // A(s0+2, s0+2), A(s0+2, s0+3)
        const double2 c3 = *reinterpret_cast<const double2*>(B1 + 2 * ((s0 + 3) ^ 1));   // A(s0+3, s0+2), A(s0+3, s0+3)
        const double piv2 = fma(n21, L21, fma(n20, L20, c2.x));
        const double sg2 = SIGNED && piv2 < 0.0 ? -1.0 : 1.0;
        const double inv2 = rsqrt_nr(SIGNED ? fabs(piv2) : piv2);
        const double i2s = SIGNED ? sg2 * inv2 : inv2;
        const double L32 = fma(n31, L21, fma(n30, L20, c3.x)) * i2s;
        const double l2 = v2 * i2s;
        const double m2s = SIGNED ? -sg2 * l2 : -l2;
        const double n32 = SIGNED ? -sg2 * L32 : -L32;
        v3 = fma(m2s, L32, v3);
        const double piv3 = fma(n32, L32, fma(n31, L31, fma(n30, L30, c3.y)));
        const double sg3 = SIGNED && piv3 < 0.0 ? -1.0 : 1.0;
        const double inv3 = rsqrt_nr(SIGNED ? fabs(piv3) : piv3);
        const double l3 = v3 * (SIGNED ? sg3 * inv3 : inv3);
        if (SIGNED) {
          if (!(fabs(piv0) > thr) | !(fabs(piv1) > thr) | !(fabs(piv2) > thr) | !(fabs(piv3) > thr)) bad = 2;
        } else {
          bad |= !(piv0 > 0.0) | !(piv1 > 0.0) | !(piv2 > 0.0) | !(piv3 > 0.0);
        }
        const double x3 = fs[s0 * kMS + s0 + 3] * (SIGNED ? sg3 * inv3 : inv3);  // L(j+32, j+3), see above
        // ---- 2. right-hand sides.  They are three more ROWS of the window (tile row 4: lane (g, m) holds rhs g,
        //         columns 8 tc + 2m, + 1 of Y[tc]; g < 3): forward substitution is the same elimination.  Lane
        //         (g, m) runs the four pivot columns on rhs g (the entries of the pivot block are the shuffled
        //         l's above), keeps w_m as its mma A fragment and stores it as the solved entry of column j + m.
        double ay;
        {
          const double2 ya = *reinterpret_cast<const double2*>(Pn + 2 * (32 + g));
          const double2 yc = *reinterpret_cast<const double2*>(Pn + kPH + 2 * ((32 + g) ^ 1));
          const double w0 = ya.x * inv0;
          const double w1 = fma(-L10, w0, ya.y) * inv1;
          const double w2 = fma(-L21, w1, fma(-L20, w0, yc.x)) * inv2;
          const double w3 = fma(-L32, w2, fma(-L31, w1, fma(-L30, w0, yc.y))) * inv3;
          ay = m == 0 ? w0 : m == 1 ? w1 : m == 2 ? w2 : w3;
          const double dk = SIGNED ? (m == 0 ? sg0 : m == 1 ? sg1 : m == 2 ? sg2 : sg3) : 1.0;
          if (g < 3) yb[(s0 + m) * 4 + g] = dk * ay;      // stored as D w: the backward substitution starts from it
          if (SIGNED && g == 3) yb[(s0 + m) * 4 + 3] = dk;
          if (g >= 3) ay = 0.0;
        }
        // ---- 3. factor records (the diagonal slot holds 1/L(j,j), see chol_invdiag_kernel); rows above the
        //         diagonal of the 4x4 pivot block are zero, row j+32 of column j+3 is x3
        {
          double r1 = l1, r2 = l2, r3 = l3;
          if (lane == s0) { r1 = 0.0; r2 = 0.0; r3 = x3; }
          if (lane == s0 + 1) { r1 = inv1; r2 = 0.0; r3 = 0.0; }
          if (lane == s0 + 2) { r2 = inv2; r3 = 0.0; }
          if (lane == s0 + 3) r3 = inv3;
          fb[s0 * 32 + FROT(s0, d0)] = lane == s0 ? inv0 : l0;
          fb[(s0 + 1) * 32 + FROT(s0 + 1, (d0 - 1) & 31)] = r1;
          fb[(s0 + 2) * 32 + FROT(s0 + 2, (d0 - 2) & 31)] = r2;
          fb[(s0 + 3) * 32 + FROT(s0 + 3, (d0 - 3) & 31)] = r3;
        }
        // ---- 4. rows / columns j+32 .. j+35 enter at slots s0 .. s0+3 (fs holds lower triangles: entry
        //         (R, C) of two incoming rows is staged with the later one as the row)
        if (rowlane) {
          const int R = 8 * tc0 + g;
          const double* fr = fs + R * kMS;
#pragma unroll
          for (int tc = 0; tc <= tc0; ++tc) {
            const double2 w = *reinterpret_cast<const double2*>(fr + 8 * tc + 2 * m);
            W[tc0][tc][0] = w.x;
            W[tc0][tc][1] = w.y;
          }
          if (collane) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int C = 8 * tc0 + 2 * m + e;
              if (C > R) W[tc0][tc0][e] = fs[C * kMS + R];
            }
          }
        }
        if (collane) {
#pragma unroll
          for (int tr = tc0; tr < 4; ++tr) {
            if (!(tr == tc0 && rowlane)) {
              W[tr][tc0][0] = fs[(8 * tc0 + 2 * m) * kMS + 8 * tr + g];
              W[tr][tc0][1] = fs[(8 * tc0 + 2 * m + 1) * kMS + 8 * tr + g];
            }
          }
        }
        if (collane && g < 3) {                            // right-hand sides of the incoming rows (the update below
          Y[tc0][0] = fs[(8 * tc0 + 2 * m) * kMS + 32 + g];      // carries x3 * w_3 into row j+32's)
          Y[tc0][1] = fs[(8 * tc0 + 2 * m + 1) * kMS + 32 + g];
        }
        // ---- 5. W -= P P^T on the tensor pipe, P = panel with the incoming rows' entries
        const bool inc = d0 < 4;
        *reinterpret_cast<double2*>(Pn + 2 * lane) = inc ? make_double2(0.0, 0.0) : make_double2(l0, l1);
        *reinterpret_cast<double2*>(Pn + kPH + 2 * (lane ^ 1)) = inc ? make_double2(0.0, d0 == 0 ? x3 : 0.0) : make_double2(l2, l3);
        __syncwarp();
        double af[4], nf[4];
        const double dm = SIGNED ? (m == 0 ? sg0 : m == 1 ? sg1 : m == 2 ? sg2 : sg3) : 1.0;   // this lane's k slice
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          af[t] = Pn[(m >> 1) * kPH + 2 * ((8 * t + g) ^ (m >> 1)) + (m & 1)];
          nf[t] = SIGNED ? -dm * af[t] : -af[t];
        }
#pragma unroll
        for (int tr = 0; tr < 4; ++tr)
#pragma unroll
          for (int tc = 0; tc <= tr; ++tc) dmma_f(W[tr][tc][0], W[tr][tc][1], nf[tr], af[tc]);
        const double nay = -ay;
#pragma unroll
        for (int tc = 0; tc < 4; ++tc) dmma_f(Y[tc][0], Y[tc][1], nay, af[tc]);
      }
    } else {
#pragma unroll
    for (int p2 = 0; p2 < 16; ++p2) {
      const int s0 = 2 * p2;            // slots (= columns of this block) of the panel: s0, s0 + 1
      const int tc0 = p2 >> 2, m0 = p2 & 3;
      // ---- 1. panel columns -> lanes = rows
      __syncwarp();                                      // the previous panel's fragment reads of Pn are done
      if (m == m0) {
#pragma unroll
        for (int tr = 0; tr < 4; ++tr)
          *reinterpret_cast<double2*>(Pn + (8 * tr + g) * 2) = make_double2(W[tr][tc0][0], W[tr][tc0][1]);
      }
      __syncwarp();
      const double2 v = *reinterpret_cast<const double2*>(Pn + 2 * lane);
      const int d0 = (lane - s0) & 31;                 // row of this lane = j0 + s0 + d0
      const double thr = SIGNED ? 1e-13 * dmax : 0.0;
      const double piv0 = __shfl_sync(kFull, v.x, s0);
      const double sg0 = SIGNED && piv0 < 0.0 ? -1.0 : 1.0;
      const double inv0 = rsqrt_nr(SIGNED ? fabs(piv0) : piv0);
      const double l0 = v.x * (SIGNED ? sg0 * inv0 : inv0);
      const double c10 = __shfl_sync(kFull, l0, s0 + 1);
      const double t1 = fma(SIGNED ? -sg0 * l0 : -l0, c10, v.y);
      const double piv1 = __shfl_sync(kFull, t1, s0 + 1);
      const double sg1 = SIGNED && piv1 < 0.0 ? -1.0 : 1.0;
      const double inv1 = rsqrt_nr(SIGNED ? fabs(piv1) : piv1);
      const double l1 = t1 * (SIGNED ? sg1 * inv1 : inv1);
      if (SIGNED) {
        if (!(fabs(piv0) > thr) | !(fabs(piv1) > thr)) bad = 2;
      } else {
        bad |= !(piv0 > 0.0) | !(piv1 > 0.0);
      }
      // ---- 2. right-hand sides (SIGNED: stored as D w, the sign in the spare fourth slot)
      const double ya0 = __shfl_sync(kFull, y[0] * inv0, s0);
      const double ya1 = __shfl_sync(kFull, y[1] * inv0, s0);
      const double ya2 = __shfl_sync(kFull, y[2] * inv0, s0);
      y[0] = fma(-l0, ya0, y[0]); y[1] = fma(-l0, ya1, y[1]); y[2] = fma(-l0, ya2, y[2]);
      const double yc0 = __shfl_sync(kFull, y[0] * inv1, s0 + 1);
      const double yc1 = __shfl_sync(kFull, y[1] * inv1, s0 + 1);
      const double yc2 = __shfl_sync(kFull, y[2] * inv1, s0 + 1);
      y[0] = fma(-l1, yc0, y[0]); y[1] = fma(-l1, yc1, y[1]); y[2] = fma(-l1, yc2, y[2]);
      if (lane == s0) {
        yb[s0 * 4] = sg0 * ya0; yb[s0 * 4 + 1] = sg0 * ya1; yb[s0 * 4 + 2] = sg0 * ya2;
        if (SIGNED) yb[s0 * 4 + 3] = sg0;
      }
      if (lane == s0 + 1) {
        yb[s0 * 4 + 4] = sg1 * yc0; yb[s0 * 4 + 5] = sg1 * yc1; yb[s0 * 4 + 6] = sg1 * yc2;
        if (SIGNED) yb[s0 * 4 + 7] = sg1;
      }
      // ---- 3. factor records and the diagonal block of the column group
      // (the diagonal slot holds 1/L(j,j); chol_invdiag_kernel turns the 8x8 diagonal blocks into
      //  their inverses afterwards)
      fb[s0 * 32 + FROT(s0, d0)] = d0 == 0 ? inv0 : l0;
      fb[(s0 + 1) * 32 + FROT(s0 + 1, (d0 - 1) & 31)] = d0 == 0 ? 0.0 : (d0 == 1 ? inv1 : l1);
      // ---- 4. W -= L_panel L_panel^T on the tensor pipe (k = 2 of 4 used)
      *reinterpret_cast<double2*>(Pn + 2 * lane) = make_double2(l0, l1);
      __syncwarp();
      double af[4], nf[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        af[t] = m < 2 ? Pn[(8 * t + g) * 2 + m] : 0.0;
        nf[t] = SIGNED ? -(m == 0 ? sg0 : sg1) * af[t] : -af[t];
      }
#pragma unroll
      for (int tr = 0; tr < 4; ++tr)
#pragma unroll
        for (int tc = 0; tc < 4; ++tc) dmma_f(W[tr][tc][0], W[tr][tc][1], nf[tr], af[tc]);
      // ---- 5. rows / columns j+32, j+33 enter at slots s0, s0+1
      if ((g >> 1) == m0) {                              // tile row tc0, rows s0 + (g & 1)
        const double* fr = fs + (s0 + (g & 1)) * kMS;
#pragma unroll
        for (int tc = 0; tc < 4; ++tc) {
          const double2 w = *reinterpret_cast<const double2*>(fr + 8 * tc + 2 * m);
          W[tc0][tc][0] = w.x;
          W[tc0][tc][1] = w.y;
        }
        if (m == m0 && (g & 1) == 0) W[tc0][tc0][1] = fs[(s0 + 1) * kMS + s0];   // (s0, s0+1) mirrors (s0+1, s0)
      }
      if (m == m0) {                                     // tile column tc0, columns s0, s0 + 1
#pragma unroll
        for (int tr = 0; tr < 4; ++tr) {
          if (!(tr == tc0 && (g >> 1) == m0)) {
            W[tr][tc0][0] = fs[s0 * kMS + 8 * tr + g];
            W[tr][tc0][1] = fs[(s0 + 1) * kMS + 8 * tr + g];
          }
        }
      }
      if ((lane >> 1) == p2) { y[0] = fs[lane * kMS + 32]; y[1] = fs[lane * kMS + 33]; y[2] = fs[lane * kMS + 34]; }
    }
    }
    {                                                    // the block's diagonal 8x8 blocks, inverted in place
      __syncwarp();                                      // (its records are written)
      double dcol[8];
      invdiag_load(fb, lane, dcol);
      invdiag_solve_store(fb, lane, dcol);
    }
  }
  if constexpr (PW == 4 && TW) {
    if (half == 1) {
      // bottom-up half: leave the updated window (M in reversed numbering, as a full symmetric matrix), its
      // right-hand-side rows and the pivot flag in the staging area for the partner, and retire
      __syncwarp();
#pragma unroll
      for (int tr = 0; tr < 4; ++tr)
#pragma unroll
        for (int tc = 0; tc <= tr; ++tc)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int R = 8 * tr + g, C = 8 * tc + 2 * m + e;
            fs[R * kMS + C] = W[tr][tc][e];
            if (tc < tr) fs[C * kMS + R] = W[tr][tc][e];
          }
      if (g < 3) {
#pragma unroll
        for (int tc = 0; tc < 4; ++tc)
#pragma unroll
          for (int e = 0; e < 2; ++e) fs[(8 * tc + 2 * m + e) * kMS + 32 + g] = Y[tc][e];
      }
      if (lane == 0) fs[32 * kMS] = (double)bad;
      __threadfence_block();
      asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp >> 1)) : "memory");
      return;
    }
  }
  if (lane == 0) {
    meta[0] = sigma;
    meta[1] = (double)bad;
    meta[2] = SIGNED ? 1.0 : 0.0;
    meta[3] = TW ? 1.0 : 0.0;                              // layout of the stored factor, read by chol_subst_kernel
  }
}

// ===========================================================================================
// substitutions with the stored factor
// ===========================================================================================
constexpr int kSubWarps = 11;  // designs per CTA; one CTA per SM (214 KB of shared memory)
constexpr int kSubThreads = kSubWarps * 32;
constexpr int kSS = 32;        // staged factor blocks are verbatim copies of global memory
constexpr int kSubWarpSmem = 2 * 32 * kSS + 6 * 64 + 2;  // 2 staged blocks, 2 transposition tiles, ring of 4 solution tiles, 2 mbarriers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(phase)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// D(8x8) += A(8x4) * B(4x8) in FP64 on the tensor pipe.  Lane (g = lane>>2, m = lane&3) holds
// A[g][m], B[m][g] and D[g][2m], D[g][2m+1].
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// TWC: the instance that also understands two-sided factors (small batches); the one-sided instance keeps the
// lean visit schedule the bandwidth-bound full batches run
template <int NBL, bool TWC = false>  // column group k reaches row groups k+1 .. k+NBL
__global__ void __launch_bounds__(kSubThreads, 1) chol_subst_kernel(WArgs A) {
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // designs per CTA = warps per CTA (host's choice); A.sub_pair: a design is the warp pair (2k, 2k+1)
  const bool pair = TWC && A.sub_pair;
  const int64_t b = pair ? (int64_t)blockIdx.x * (blockDim.x >> 6) + (warp >> 1)
                         : (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= A.batch || (pair && warp >= (int)((blockDim.x >> 6) << 1))) return;
  double* fsb = smem + warp * kSubWarpSmem;   // two staged blocks
  double* zt1 = fsb + 2 * 32 * kSS;
  double* zt2 = zt1 + 64;
  double* xt = zt2 + 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xt + 4 * 64);
  double* xt_other = smem + (warp ^ 1) * kSubWarpSmem + 2 * 32 * kSS + 128;   // the partner's ring (A.sub_pair)
  const int g = lane >> 2, m = lane & 3;
  const int n = A.n, nblk = A.nblk;
  const double* rhs = A.rhs + b * (int64_t)n * 3;
  // A.shared_factor (opts.reuse_factor == 2): every right-hand-side set of the batch is solved with the ONE factor
  // in slot 0 of the workspace (a tangent / multi-rhs solve: same K, many right-hand sides); each set keeps its own
  // intermediate vector in its own slot
  double* fac = A.ws + (A.shared_factor ? 0 : b * A.ws_stride);
  double* ysol = A.ws + b * A.ws_stride + (int64_t)nblk * 1024;
  const double* dsign = fac + (int64_t)nblk * 1024;        // D of L D L^T lives with the factor (slot 3 of ITS records)
  const double* meta = fac + (int64_t)nblk * (1024 + 128);
  const double sigma = meta[0];
  // a two-sided factor met by the one-sided instance (the two calls of a solve disagreed on the batch in flight) is
  // reported as a breakdown rather than read wrongly
  const int bad = !TWC && meta[3] != 0.0 ? 2 : (int)meta[1];
  const bool sgn = meta[2] != 0.0;                        // sigma*K = L D L^T: D in slot 3 of the intermediate rhs
  // Two-sided factor (meta[3], written by the factor kernel): half 0 = [T | M] in blocks 0 .. nbT, half 1 = B
  // bottom-up in the blocks after them.  One warp (role 0): forward T, forward B, forward M; backward M, T, B; the
  // running right-hand side of M after T is set aside while B runs and B's part (reversed numbering) is added to
  // it; x_M is handed to B's backward sweep the same way.  A warp pair (roles 1 = top, 2 = bottom) runs the T and B
  // sweeps side by side and meets on M through the partner's shared-memory ring and two named barriers.
  const bool tw = TWC && NBL >= 3 && meta[3] != 0.0;
  int role = pair ? 1 + (warp & 1) : 0;
  if (role && !tw) {                                       // paired launch on a one-sided factor: the first warp does it all
    if (role == 2) return;
    role = 0;
  }
  const int nbT = tw ? A.h[0].nblk - 1 : nblk, nbB = tw ? A.h[1].nblk : 0;
  const int bar_a = 1 + 2 * (warp >> 1), bar_b = bar_a + 1;   // named barriers of the pair

  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = lane; i < 4 * 64; i += 32) xt[i] = 0.0;
  __syncwarp();

  // visit v = 0 .. nvis-1: forward pass (if A.forward), then backward pass; (half, block in the half) of a visit
  const int nmine = role == 0 ? nblk : role == 1 ? nbT + 1 : nbB;   // blocks this warp sweeps
  const int nfwd = A.forward ? nmine : 0;
  const int nvis = nfwd + nmine;
  auto half_of = [&](int v) {
    if (role) return role - 1;
    if (!tw) return 0;
    const int u = v < nfwd ? v : v - nfwd;
    return v < nfwd ? (u >= nbT && u < nbT + nbB ? 1 : 0) : (u > nbT ? 1 : 0);
  };
  auto blk_of = [&](int v) {                               // block index inside its half
    if (role || !tw) return v < nfwd ? v : nmine - 1 - (v - nfwd);
    const int u = v < nfwd ? v : v - nfwd;
    if (v < nfwd) return u < nbT ? u : (u < nbT + nbB ? u - nbT : nbT);
    return u == 0 ? nbT : (u <= nbT ? nbT - u : nbB - (u - nbT));
  };
  auto gblock_of = [&](int v) { return (tw ? A.h[half_of(v)].block0 : 0) + blk_of(v); };   // factor block in the workspace
  auto iperm_of = [&](int v) { return tw ? A.h[half_of(v)].iperm : A.iperm; };
  auto rows_in = [&](int v) { return tw ? A.h[half_of(v)].rows : n; };
  auto issue = [&](int v) {  // one 8 KB bulk copy: the 32 factor columns of the visit's block into buffer v&1
    FDM_ASSERT(gblock_of(v) >= 0 && gblock_of(v) < nblk, "visit %d -> factor block %d of %d (role %d)", v, gblock_of(v), nblk, role);
    if (lane == 0) {
      uint64_t* bar = bars + (v & 1);
      mbar_expect_tx(bar, 32 * 256);
      tma_load_1d(fsb + (v & 1) * 32 * kSS, fac + (int64_t)gblock_of(v) * 1024, 32 * 256, bar);
    }
  };
  uint32_t phases = 0u;   // bit i: parity the next wait on barrier i expects
  issue(0);

  // forward-pass state: T[i] = rows of group k+i of the running right-hand side, 8x8 tile (columns 0..2
  // used) in the accumulator layout of the FP64 mma: lane (g, m) holds row g, columns 2m and 2m+1.
  auto load_tile = [&](const int32_t* ip, int rows, int R, double (&t)[2]) {
    t[0] = 0.0; t[1] = 0.0;
    if (R < rows && m < 2) {
      const int gi = __ldg(ip + R);
      if (gi >= 0) {
        t[0] = sigma * rhs[3 * gi + 2 * m];
        if (m == 0) t[1] = sigma * rhs[3 * gi + 1];
      }
    }
  };
  double T[NBL + 1][2];
#pragma unroll
  for (int i = 0; i <= NBL; ++i) {
    T[i][0] = 0.0; T[i][1] = 0.0;
    if (nfwd) load_tile(iperm_of(0), rows_in(0), 8 * i + g, T[i]);
  }
  double TM[4][2], XM[4][2];                               // two-sided: M's running rhs after T; x_M
#pragma unroll
  for (int i = 0; i < 4; ++i) { TM[i][0] = TM[i][1] = 0.0; XM[i][0] = XM[i][1] = 0.0; }
  // Global operands of a visit are fetched ahead so that no load sits on a warp's critical path:
  // band-to-free row indices two visits ahead (Qn -> Q -> Qc), values one visit ahead (P -> C).
  // Forward visit: the four row groups that enter the window.  Backward visit: the four groups
  // of z it starts from (and the row indices for storing x).
  double P[4][2], C[4][2];
  int Qn[4], Q[4], Qc[4];
  auto rows_of = [&](int v, int kb) {
    const int r0 = blk_of(v) * 32 + 8 * kb + g;
    return v < nfwd ? r0 + 8 * (NBL + 1) : r0;
  };
  auto fetch_idx = [&](int v) {
    const int32_t* ip = iperm_of(v);
    const int rows = rows_in(v);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      const int R = rows_of(v, kb);
      Qn[kb] = R < rows ? __ldg(ip + R) : -1;
      FDM_ASSERT(R >= 0 && Qn[kb] >= -1 && Qn[kb] < n, "row %d -> free index %d of %d (visit %d, role %d)", R, Qn[kb], n, v, role);
    }
  };
  auto fetch_val = [&](int v) {   // uses Q = row indices of visit v
    if (v < nfwd) {
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        P[kb][0] = 0.0; P[kb][1] = 0.0;
        if (Q[kb] >= 0 && m < 2) {
          P[kb][0] = rhs[3 * Q[kb] + 2 * m];            // scaled by sigma where consumed: no use here, no stall
          if (m == 0) P[kb][1] = rhs[3 * Q[kb] + 1];
        }
      }
    } else {
      const int r0 = gblock_of(v) * 32;
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        P[kb][0] = 0.0; P[kb][1] = 0.0;
        if (m < 2) {
          const double2 zv = *reinterpret_cast<const double2*>(ysol + (int64_t)(r0 + 8 * kb + g) * 4 + 2 * m);
          P[kb][0] = zv.x;
          if (m == 0) P[kb][1] = zv.y;
        }
      }
    }
  };
  fetch_idx(0);
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) Q[kb] = Qn[kb];
  fetch_val(0);
  if (nvis > 1) fetch_idx(1);
  const int mirror = ((7 - g) << 2) | m;                   // lane holding row 7 - g of a tile (same columns)

  for (int v = 0; v < nvis; ++v) {
    const int j0 = gblock_of(v) * 32;                      // first row of the block in the workspace numbering
    __syncwarp();                      // every lane is done with buffer (v+1)&1 (used at visit v-1)
    if (v + 1 < nvis) issue(v + 1);
    mbar_wait(bars + (v & 1), (phases >> (v & 1)) & 1u);
    phases ^= 1u << (v & 1);
    const double* fs = fsb + (v & 1) * 32 * kSS;
    const int rm = 5 * m, rg = 5 * (g & 3);   // record rotation of the columns this lane reads (forward / backward)
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      const double sc = v < nfwd ? sigma : 1.0;
      C[kb][0] = P[kb][0] * sc; C[kb][1] = P[kb][1] * sc; Qc[kb] = Q[kb]; Q[kb] = Qn[kb];
    }
    if (v + 1 < nvis && v + 1 != nfwd) fetch_val(v + 1);
    if (v + 2 < nvis) fetch_idx(v + 2);
    if (role == 0 && tw && nfwd) {
      if (v == nbT) {
        // T is done: its window holds M's rows.  Set them aside and start B's sweep from B's own first rows.
#pragma unroll
        for (int i = 0; i < 4; ++i) { TM[i][0] = T[i][0]; TM[i][1] = T[i][1]; }
#pragma unroll
        for (int i = 0; i <= NBL; ++i) load_tile(A.h[1].iperm, A.h[1].rows, 8 * i + g, T[i]);
      } else if (v == nbT + nbB) {
        // B is done: its window holds its part of M in reversed numbering (group i there = group 3 - i here, row
        // g there = row 7 - g here).  M's right-hand side = what T left + what B left.
        double R4[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          R4[i][0] = __shfl_sync(kFull, T[i][0], mirror);
          R4[i][1] = __shfl_sync(kFull, T[i][1], mirror);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { T[i][0] = TM[i][0] + R4[3 - i][0]; T[i][1] = TM[i][1] + R4[3 - i][1]; }
        if (NBL == 4) { T[NBL][0] = 0.0; T[NBL][1] = 0.0; }
      }
    }
    if (role == 0 && tw && v == nfwd + nbT + 1) {
      // backward: M and T are done, B starts from x_M in its own numbering
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double x0 = __shfl_sync(kFull, XM[3 - i][0], mirror), x1 = __shfl_sync(kFull, XM[3 - i][1], mirror);
        *reinterpret_cast<double2*>(xt + i * 64 + g * 8 + 2 * m) = make_double2(x0, x1);
      }
      __syncwarp();
    }
    if (role == 1 && nfwd && v == nbT) {
      // top of a pair, forward: T is done and the window holds M; the partner has left B's part of M (reversed
      // numbering) in its ring.  Add it: group i, row g here = group 3 - i, row 7 - g there.
      asm volatile("bar.sync %0, 64;" ::"r"(bar_a) : "memory");
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 o = *reinterpret_cast<const double2*>(xt_other + (3 - i) * 64 + (7 - g) * 8 + 2 * m);
        T[i][0] += o.x; T[i][1] += o.y;
      }
      if (NBL == 4) { T[NBL][0] = 0.0; T[NBL][1] = 0.0; }
    }
    if (role == 2 && v == nfwd) {
      // bottom of a pair: after its forward sweep the window holds its part of M -> the ring, for the partner;
      // then x_M comes back through the same ring before the backward sweep starts
      if (nfwd) {
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<double2*>(xt + i * 64 + g * 8 + 2 * m) = make_double2(T[i][0], T[i][1]);
        __threadfence_block();
        asm volatile("bar.sync %0, 64;" ::"r"(bar_a) : "memory");
      }
      asm volatile("bar.sync %0, 64;" ::"r"(bar_b) : "memory");
    }
    if (v < nfwd) {
      // ---------------------------------------------------------------- forward: L z = sigma g
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const int cb = 8 * kb;
        // z_k = inv(L_kk) T[0]
        *reinterpret_cast<double2*>(zt1 + g * 8 + 2 * m) = make_double2(T[0][0], T[0][1]);
        __syncwarp();
        double z0 = 0.0, z1 = 0.0;
        {
          const double b0 = zt1[m * 8 + g], b1 = zt1[(4 + m) * 8 + g];
          const double a0 = g >= m ? fs[(cb + m) * kSS + ((g - m + rm) & 31)] : 0.0;
          const double a1 = g >= 4 + m ? fs[(cb + 4 + m) * kSS + ((g - 4 - m + rm) & 31)] : 0.0;
          dmma(z0, z1, a0, b0);
          dmma(z0, z1, a1, b1);
        }
        if (!sgn) {
          if (m < 2) *reinterpret_cast<double2*>(ysol + (int64_t)(j0 + cb + g) * 4 + 2 * m) = make_double2(z0, z1);
        } else {                                           // stored as D z (slot 3 keeps D): the backward sweep starts from it
          double* zo = ysol + (int64_t)(j0 + cb + g) * 4;
          const double dk = dsign[(int64_t)(j0 + cb + g) * 4 + 3];
          if (m == 0) *reinterpret_cast<double2*>(zo) = make_double2(dk * z0, dk * z1);
          if (m == 1) zo[2] = dk * z0;
        }
        // T[i] -= L_{k+i,k} z_k
        *reinterpret_cast<double2*>(zt2 + g * 8 + 2 * m) = make_double2(z0, z1);
        __syncwarp();
        const double nb0 = -zt2[m * 8 + g], nb1 = -zt2[(4 + m) * 8 + g];
#pragma unroll
        for (int i = 1; i <= NBL; ++i) {
          const int d0 = 8 * i + g - m, d1 = d0 - 4;
          const double a0 = d0 <= 31 ? fs[(cb + m) * kSS + ((d0 + rm) & 31)] : 0.0;
          const double a1 = d1 <= 31 ? fs[(cb + 4 + m) * kSS + ((d1 + rm) & 31)] : 0.0;
          dmma(T[i][0], T[i][1], a0, nb0);
          dmma(T[i][0], T[i][1], a1, nb1);
        }
#pragma unroll
        for (int i = 0; i < NBL; ++i) { T[i][0] = T[i + 1][0]; T[i][1] = T[i + 1][1]; }
        T[NBL][0] = C[kb][0]; T[NBL][1] = C[kb][1];   // rows of group k+NBL+1 enter the window
      }
      if (v + 1 == nfwd) {   // the first backward visit starts from z written during this visit
        __syncwarp();
        fetch_val(v + 1);
      }
    } else {
      // ---------------------------------------------------------------- backward: L^T x = z
#pragma unroll
      for (int kb = 3; kb >= 0; --kb) {
        const int cb = 8 * kb;
        double acc0 = C[kb][0], acc1 = C[kb][1], acd0 = 0.0, acd1 = 0.0;   // two chains of tensor ops
        // acc -= L_{k+i,k}^T x_{k+i}: A[g][kk] = L(8(k+i)+kk, 8k+g), B[kk][nn] = x_{k+i}(kk, nn)
#pragma unroll
        for (int i = 1; i <= NBL; ++i) {
          const double* xs = xt + ((kb + i) & 3) * 64;
          const int d0 = 8 * i + m - g, d1 = d0 + 4;
          const double a0 = d0 <= 31 ? fs[(cb + g) * kSS + ((d0 + rg) & 31)] : 0.0;
          const double a1 = d1 <= 31 ? fs[(cb + g) * kSS + ((d1 + rg) & 31)] : 0.0;
          dmma(acc0, acc1, a0, -xs[m * 8 + g]);
          dmma(acd0, acd1, a1, -xs[(4 + m) * 8 + g]);
        }
        acc0 += acd0; acc1 += acd1;
        // x_k = inv(L_kk)^T acc: A[g][kk] = inv(L_kk)(kk, g)
        *reinterpret_cast<double2*>(zt1 + g * 8 + 2 * m) = make_double2(acc0, acc1);
        __syncwarp();
        double x0 = 0.0, x1 = 0.0;
        {
          const double b0 = zt1[m * 8 + g], b1 = zt1[(4 + m) * 8 + g];
          const double a0 = m >= g ? fs[(cb + g) * kSS + ((m - g + rg) & 31)] : 0.0;
          const double a1 = 4 + m >= g ? fs[(cb + g) * kSS + ((4 + m - g + rg) & 31)] : 0.0;
          dmma(x0, x1, a0, b0);
          dmma(x0, x1, a1, b1);
        }
        *reinterpret_cast<double2*>(xt + (kb & 3) * 64 + g * 8 + 2 * m) = make_double2(x0, x1);
        if (role == 0 && tw && v == nfwd) { XM[kb][0] = x0; XM[kb][1] = x1; }   // x_M, for B's sweep
        if (role == 1 && v == nfwd)                        // x_M into the partner's ring, in ITS numbering
          *reinterpret_cast<double2*>(xt_other + (3 - kb) * 64 + (7 - g) * 8 + 2 * m) = make_double2(x0, x1);
        if (Qc[kb] >= 0 && m < 2) {
          double* xo = A.x + (b * (int64_t)n + Qc[kb]) * 3 + 2 * m;
          xo[0] = x0;
          if (m == 0) xo[1] = x1;
        }
        __syncwarp();
      }
      if (role == 1 && v == nfwd) {                        // the partner may start its backward sweep
        __threadfence_block();
        asm volatile("bar.sync %0, 64;" ::"r"(bar_b) : "memory");
      }
    }
  }
  if (role) asm volatile("bar.sync %0, 64;" ::"r"(bar_a) : "memory");   // both sweeps have stored their rows of x
  if (role == 2) return;                                   // the top warp reports
  if (A.nan_fail && bad != 0) {                            // fdm_solve_opts.nan_on_failure
    __syncwarp();
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    double* xo = A.x + b * (int64_t)n * 3;
    for (int i = lane; i < 3 * n; i += 32) xo[i] = qnan;
  }
  if (A.info && lane == 0) {
    double* inf = A.info + b * FDM_INFO_STRIDE;
    inf[FDM_INFO_STATUS] = bad == 0 ? (double)FDM_STATUS_OK : bad == 2 ? (double)FDM_STATUS_BREAKDOWN : (double)FDM_STATUS_INDEFINITE;
    inf[FDM_INFO_ITERS] = 0.0;
    inf[FDM_INFO_RELRES] = 0.0;
    inf[FDM_INFO_SIGN] = sgn ? 0.0 : sigma;               // 0: indefinite K, solved with signed pivots
  }
}

// four-column panels need column j+3 to stay within one row of the 32-row window; FDM_CHOL_PANEL2 is the A/B switch
inline bool mma_panel4(const fdm_plan* p) {
  static const bool force2 = std::getenv("FDM_CHOL_PANEL2") != nullptr;
  return p->hbw <= 29 && !force2;
}

// designs (= warps) per CTA of the tensor-core factor kernels; FDM_CHOL_FACWARPS overrides for experiments
// Two-sided factorisation of a batch too small to give every SM a four-design CTA (a warp pair per design): ONE design
// per CTA -- 512 designs are then 512 CTAs over all 148 SMs instead of 128 CTAs on 128 of them (step 0.314 -> 0.305 ms,
// one design 0.239 -> 0.226 ms); from ~600 designs on the four-design CTAs are better again (1024: 0.50 against 0.585).
inline int mma_warps_per_cta(bool single = false) {
  static const int w = [] {
    const char* e = std::getenv("FDM_CHOL_FACWARPS");
    const int v = e ? std::atoi(e) : 0;
    return v >= 1 && v <= kFacWarps ? v : 0;
  }();
  return w ? w : single ? 2 : kFacWarps;
}

// designs (= warps) per CTA of the substitution kernel (FDM_CHOL_SUBWARPS) and extra dynamic shared memory
// requested by the factor kernels to cap their CTAs per SM (FDM_CHOL_FAC_SMEM_PAD, bytes): experiments on
// co-residency of the latency-bound factorisation of one slice with the bandwidth-bound kernels of another
inline int sub_warps_per_cta(int dflt) {
  static const int w = [] {
    const char* e = std::getenv("FDM_CHOL_SUBWARPS");
    const int v = e ? std::atoi(e) : 0;
    return v >= 1 && v <= kSubWarps ? v : 0;
  }();
  return w ? w : dflt;
}
inline size_t fac_smem_pad() {
  static const size_t v = [] {
    const char* e = std::getenv("FDM_CHOL_FAC_SMEM_PAD");
    return e ? (size_t)std::atol(e) : (size_t)0;
  }();
  return v;
}

// Two-sided factorisation: two warps per design, half the dependency chain, same flops.  It pays when the batch
// leaves SMs short of warps (one-sided: one warp per design); with every SM full it only adds the hand-over.
// FDM_CHOL_TWISTED = 0 / 1 forces it off / on (where the plan has the tables).
inline bool use_twisted(const fdm_plan* p, int64_t batch) {
  static const int force = [] {
    const char* e = std::getenv("FDM_CHOL_TWISTED");
    return e ? std::atoi(e) : -1;
  }();
  if (p->tw[0].nblk == 0 || p->hbw < 16 || p->hbw > 29 || !p->d.tw_iperm[0]) return false;
  if (force >= 0) return force != 0;
  return batch <= (int64_t)p->sm_count * 8;
}

inline void fill_halves(WArgs& A, const fdm_plan* p, bool from_q) {
  for (int h = 0; h < 2; ++h) {
    WHalf& H = A.h[h];
    const bool on = p->tw[0].nblk > 0 && p->d.tw_iperm[0];
    H.iperm = on ? p->d.tw_iperm[h] : nullptr;
    H.stage_ptr = on ? p->d.tw_stage_ptr[h] : nullptr;
    H.stage_ent = on ? (from_q ? p->d.tw_stage_eent[h] : p->d.tw_stage_kent[h]) : nullptr;
    H.row_diag = on ? p->d.tw_row_diag[h] : nullptr;
    H.row_sup = on ? p->d.tw_row_sup[h] : nullptr;
    H.row_node = on ? p->d.tw_row_node[h] : nullptr;
    H.nblk = on ? p->tw[h].nblk : 0;
    H.rows = on ? (int)p->tw[h].iperm.size() : 0;
    H.block0 = h == 0 ? 0 : p->tw[0].nblk;
  }
  A.twisted = 0;
  A.sub_pair = 0;
  A.shared_factor = 0;
  A.num_edges = (int)p->E; A.num_nodes = (int)p->N; A.num_fixed = (int)p->Ns;
  A.sigma_f = p->band_iperm.empty() ? 0 : p->band_iperm[0];
  A.sigma_k = p->band_rowkidx.empty() ? 0 : p->band_rowkidx[0];
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

bool chol_warp_from_q_supported(const fdm_plan* p) { return chol_warp_supported(p) && p->hbw <= 30; }

// -------------------------------------------------------------------------------------------
// Designs factored with signed pivots (mixed-sign q: sigma K = L D L^T without pivoting) get their residual checked:
// element growth can cost digits there and nothing else would notice.  One CTA per design, returning at once unless
// the design was solved that way (FDM_INFO_SIGN 0, status OK): r = b - K x from q (the operator of models.py:1064-1079
// applied edge by edge: r_f = p_f - sum_inc q_e (x_f - x_other), b_f = p_f + sum_{e to support} q_e x_s) or from
// K.data and the right-hand side; max ||r|| / ||b|| over the three right-hand sides goes to FDM_INFO_RELRES, and a
// design above 1e-6 is reported as a breakdown (NaN positions under nan_on_failure) instead of being handed out.
// -------------------------------------------------------------------------------------------
namespace {

struct RArgs {
  int n, batch, from_q, nan_fail;
  int64_t nnz, q_stride, xs_stride, loads_stride;
  const double *q, *xs, *loads, *Kdata, *rhs;
  const int32_t *rowptr, *colidx, *free_node, *node_slot, *ninc_ptr, *ninc_edge, *ninc_other;
  double *x, *info;
};

__global__ void __launch_bounds__(256) signed_residual_check_kernel(RArgs A) {
  const int b = blockIdx.x, tid = threadIdx.x, n = A.n;
  double* info = A.info + (int64_t)b * FDM_INFO_STRIDE;
  if (info[FDM_INFO_SIGN] != 0.0 || info[FDM_INFO_STATUS] != (double)FDM_STATUS_OK) return;
  __shared__ double s_red[6];
  if (tid < 6) s_red[tid] = 0.0;
  __syncthreads();
  double* x = A.x + (int64_t)b * n * 3;
  double rr[3] = {0.0, 0.0, 0.0}, bb[3] = {0.0, 0.0, 0.0};
  for (int f = tid; f < n; f += 256) {
    const double xf[3] = {x[3 * f], x[3 * f + 1], x[3 * f + 2]};
    double r[3], rhs[3];
    if (A.from_q) {
      const double* q = A.q + b * A.q_stride;
      const double* xs = A.xs + b * A.xs_stride;
      const int node = __ldg(A.free_node + f);
      const double* p = A.loads + b * A.loads_stride + 3 * (int64_t)node;
#pragma unroll
      for (int c = 0; c < 3; ++c) { r[c] = p[c]; rhs[c] = p[c]; }
      for (int k = __ldg(A.ninc_ptr + node); k < __ldg(A.ninc_ptr + node + 1); ++k) {
        const double qe = q[__ldg(A.ninc_edge + k) >> 1];
        const int slot = __ldg(A.node_slot + __ldg(A.ninc_other + k));
        const double* xo = slot >= 0 ? x + 3 * (int64_t)slot : xs + 3 * (int64_t)(-slot - 1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          r[c] -= qe * (xf[c] - xo[c]);
          if (slot < 0) rhs[c] += qe * xo[c];
        }
      }
    } else {
      const double* Kd = A.Kdata + (int64_t)b * A.nnz;
      const double* g = A.rhs + ((int64_t)b * n + f) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) { r[c] = g[c]; rhs[c] = g[c]; }
      for (int k = __ldg(A.rowptr + f); k < __ldg(A.rowptr + f + 1); ++k) {
        const double v = Kd[k];
        const double* xc = x + 3 * (int64_t)__ldg(A.colidx + k);
#pragma unroll
        for (int c = 0; c < 3; ++c) r[c] -= v * xc[c];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { rr[c] += r[c] * r[c]; bb[c] += rhs[c] * rhs[c]; }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    for (int o = 16; o > 0; o >>= 1) {
      rr[c] += __shfl_xor_sync(kFull, rr[c], o);
      bb[c] += __shfl_xor_sync(kFull, bb[c], o);
    }
    if ((tid & 31) == 0) { atomicAdd(&s_red[c], rr[c]); atomicAdd(&s_red[3 + c], bb[c]); }
  }
  __syncthreads();
  double rel = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) rel = fmax(rel, s_red[3 + c] > 0.0 ? sqrt(s_red[c] / s_red[3 + c]) : 0.0);
  const bool ok = rel <= 1e-6;                                       // false for NaN too
  if (tid == 0) {
    info[FDM_INFO_RELRES] = rel;
    if (!ok) info[FDM_INFO_STATUS] = (double)FDM_STATUS_BREAKDOWN;
  }
  if (!ok && A.nan_fail)
    for (int i = tid; i < 3 * n; i += 256) x[i] = __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace

static int launch_signed_check(const fdm_plan* p, const WArgs& W, int64_t batch, bool from_q, cudaStream_t st) {
  if (!W.info) return FDM_OK;
  RArgs A;
  A.n = W.n; A.batch = (int)batch; A.from_q = from_q ? 1 : 0; A.nan_fail = W.nan_fail;
  A.nnz = W.nnz; A.q_stride = W.q_stride; A.xs_stride = W.xs_stride; A.loads_stride = W.loads_stride;
  A.q = W.q; A.xs = W.xs; A.loads = W.loads; A.Kdata = W.Kdata; A.rhs = W.rhs;
  A.rowptr = p->d.rowptr; A.colidx = p->d.colidx; A.free_node = p->d.free_node; A.node_slot = p->d.node_slot;
  A.ninc_ptr = p->d.ninc_ptr; A.ninc_edge = p->d.ninc_edge; A.ninc_other = p->d.ninc_other;
  A.x = W.x; A.info = W.info;
  signed_residual_check_kernel<<<(unsigned)batch, 256, 0, st>>>(A);
  return FDM_OK;
}

static void fill_common(WArgs& A, const fdm_plan* p, double* x, double* info, void* ws, int64_t batch) {
  A.n = (int)p->Nf;
  A.bw = p->hbw;
  A.nblk = (int)((p->Nf + 31) / 32);
  A.rowkidx = p->d.band_rowkidx;
  A.iperm = p->d.band_iperm;
  A.Kdata = nullptr;
  A.rhs = nullptr;
  A.x = x;
  A.info = info;
  A.ws = (double*)ws;
  A.ws_stride = ws_stride_of(p);
  A.nnz = p->nnz;
  A.batch = batch;
  A.forward = 0;
  A.nan_fail = 0;
  A.q = A.xs = A.loads = nullptr;
  A.q_stride = A.xs_stride = A.loads_stride = 0;
  A.diags_ptr = p->d.diags_ptr; A.diags_idx = p->d.diags_idx;
  A.free_node = p->d.free_node; A.node_slot = p->d.node_slot; A.ninc_ptr = p->d.ninc_ptr;
  A.ninc_edge = p->d.ninc_edge; A.ninc_other = p->d.ninc_other; A.sup_mask = p->d.sup_mask;
  A.stage_ptr = p->d.band_stage_ptr; A.stage_ent = p->d.band_stage_eent;
  A.row_diag = p->d.band_row_diag; A.row_sup = p->d.band_row_sup; A.row_node = p->d.band_row_node;
  fill_halves(A, p, true);
}

// The substitution launch.  Small batches whose factor is two-sided get a warp pair per design (the T and B sweeps
// side by side); the predicate is the factor kernel's own, so both calls of a solve agree (the kernel falls back to
// one warp when the stored factor turns out one-sided).
static int launch_backward(const fdm_plan* p, WArgs& A, int64_t batch, cudaStream_t st, int64_t inflight) {
  // a shared factor (opts.reuse_factor == 2) was laid out by an earlier call with ITS batch: take the instance that
  // reads either layout
  const bool twc = mma_panel4(p) && (use_twisted(p, inflight) || (A.shared_factor && use_twisted(p, 1)));   // `batch`: designs in flight (fdm_solve_opts.concurrent_designs)   // the factor of this batch may be two-sided
  void (*skern)(WArgs) = p->hbw <= 7 ? chol_subst_kernel<1> : p->hbw <= 15 ? chol_subst_kernel<2>
                       : p->hbw <= 23 ? (twc ? chol_subst_kernel<3, true> : chol_subst_kernel<3>)
                                      : (twc ? chol_subst_kernel<4, true> : chol_subst_kernel<4>);
  static const bool no_pair = std::getenv("FDM_CHOL_SUBPAIR") && std::atoi(std::getenv("FDM_CHOL_SUBPAIR")) == 0;
  A.sub_pair = twc && !no_pair ? 1 : 0;
  // A sliced caller (designs in flight beyond this call's batch) runs this kernel beside the factor kernels of other
  // slices: a CTA of 4 warps (78 KB, 21 K registers) fits into the hole ONE retiring factor CTA leaves (84 KB, 32 K
  // registers), the full 11-warp CTA needs a whole SM and would wait until the queue of factor CTAs has drained
  // (csrc/pipeline.cu: prioritise).  Alone, the 11-warp CTA streams best (0.195 ms backward at B = 4096).
  // (the paired instance of a small batch is latency-bound whatever its CTA size; small CTAs start beside the
  //  kernels of the side stream instead of waiting for a whole SM: 512 designs 0.323 -> 0.313 ms)
  int sw = sub_warps_per_cta(inflight > batch || A.sub_pair ? 4 : kSubWarps);
  // (measured: 5, 6, 8 or 11 warps for the adjoint's two sweeps only -- 1.50, 1.56, 1.47, 1.49 ms against 1.46 with 4)
  if (A.sub_pair) sw = std::max(2, sw & ~1);
  const int dpc = A.sub_pair ? sw / 2 : sw;                // designs per CTA
  const size_t smem = (size_t)sw * kSubWarpSmem * sizeof(double);
  FDM_CUDA(cudaFuncSetAttribute(skern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  skern<<<(unsigned)((batch + dpc - 1) / dpc), sw * 32, smem, st>>>(A);
  return FDM_OK;
}

// Is `func` (the host-side handle of a __global__ function, as a CUDA graph's kernel node reports it) one of the
// factorisation kernels of this file?  The pipeline gives those nodes the LOW launch priority and every other kernel
// of a slice the high one: a factor CTA holds a quarter of an SM's registers for ~0.3 ms and leaves HBM idle, so when
// slices run side by side the bandwidth-bound kernels of one slice should take freed CTA slots ahead of the queued
// factor CTAs of the next (csrc/pipeline.cu: prioritise).
bool chol_warp_is_factor_kernel(const void* func) {
  // (the SIGNED second-pass instances are NOT in the list: they return at once for every design whose pivots were
  //  positive and sit between a slice's factorisation and its substitution, so they must not queue behind the
  //  factor CTAs of other slices)
  const void* const fac[] = {
      (const void*)chol_factor_mma_kernel<false, 2, false>, (const void*)chol_factor_mma_kernel<false, 4, false>,
      (const void*)chol_factor_mma_kernel<false, 4, false, true>,
      (const void*)chol_factor_mma_kernel<true, 2, false>,  (const void*)chol_factor_mma_kernel<true, 4, false>,
      (const void*)chol_factor_mma_kernel<true, 4, false, true>,
      (const void*)chol_factor_kernel<7>,  (const void*)chol_factor_kernel<15>,
      (const void*)chol_factor_kernel<23>, (const void*)chol_factor_kernel<31>};
  for (const void* f : fac)
    if (f == func) return true;
  return false;
}

int chol_warp_solve_from_q(const fdm_plan* p, const double* q, const double* xs, const double* loads, double* x,
                           double* info, void* ws, size_t ws_bytes, int64_t batch, int64_t qs, int64_t xss,
                           int64_t ls, const fdm_solve_opts& o, cudaStream_t st) {
  if (!chol_warp_from_q_supported(p)) {
    fdm_set_error("chol_warp_solve_from_q: half bandwidth " + std::to_string(p->hbw) + " > 30");
    return FDM_ERR_UNSUPPORTED;
  }
  if (ws_bytes < chol_warp_workspace_bytes(p, batch)) {
    fdm_set_error("chol_warp_solve_from_q: workspace too small");
    return FDM_ERR_WORKSPACE;
  }
  WArgs A;
  fill_common(A, p, x, info, ws, batch);
  A.q = q; A.xs = xs; A.loads = loads;
  A.q_stride = qs; A.xs_stride = xss; A.loads_stride = ls;
  A.nan_fail = o.nan_on_failure;
  const int64_t inflight = std::max<int64_t>(batch, o.concurrent_designs);
  A.twisted = mma_panel4(p) && use_twisted(p, inflight) ? 1 : 0;
  const int stages = o.stages ? o.stages : 7;
  int fw = mma_warps_per_cta(A.twisted && inflight <= 4 * (int64_t)std::max(p->sm_count, 1));
  if (A.twisted) fw = std::max(2, fw & ~1);                // a design is a warp pair
  const int dpc = A.twisted ? fw / 2 : fw;                 // designs per CTA
  const size_t smem = (size_t)fw * kMmaWarpSmem * sizeof(double) + fac_smem_pad();
  void (*fkern)(WArgs) = A.twisted ? chol_factor_mma_kernel<true, 4, false, true>
                       : mma_panel4(p) ? chol_factor_mma_kernel<true, 4, false> : chol_factor_mma_kernel<true, 2, false>;
  if (stages & 1) {
    FDM_CUDA(cudaFuncSetAttribute(fkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fkern<<<(unsigned)((batch + dpc - 1) / dpc), fw * 32, smem, st>>>(A);
    // designs with a non-positive pivot (mixed-sign q) are factored again with signed pivots
    void (*skern)(WArgs) = A.twisted ? chol_factor_mma_kernel<true, 4, true, true>
                         : mma_panel4(p) ? chol_factor_mma_kernel<true, 4, true> : chol_factor_mma_kernel<true, 2, true>;
    FDM_CUDA(cudaFuncSetAttribute(skern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    skern<<<(unsigned)((batch + dpc - 1) / dpc), fw * 32, smem, st>>>(A);
  }
  if (stages & 4) {
    int rc = launch_backward(p, A, batch, st, inflight);
    if (rc != FDM_OK) return rc;
    if (stages & 1) launch_signed_check(p, A, batch, true, st);
  }
  return fdm_check_launch("chol_warp_solve_from_q");
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
  A.forward = o.reuse_factor ? 1 : 0;
  A.nan_fail = o.nan_on_failure;
  A.q = A.xs = A.loads = nullptr;
  A.q_stride = A.xs_stride = A.loads_stride = 0;
  A.diags_ptr = A.diags_idx = A.free_node = A.node_slot = nullptr;
  A.ninc_ptr = A.ninc_edge = A.ninc_other = nullptr; A.sup_mask = nullptr;
  A.stage_ptr = p->d.band_stage_ptr; A.stage_ent = p->d.band_stage_kent;
  A.row_diag = A.row_sup = A.row_node = nullptr;
  fill_halves(A, p, false);
  A.shared_factor = o.reuse_factor == 2 ? 1 : 0;
  void (*fkern)(WArgs) = nullptr;
  void (*skern)(WArgs) = nullptr;
  if (p->hbw <= 7) { fkern = chol_factor_kernel<7>; skern = chol_subst_kernel<1>; }
  else if (p->hbw <= 15) { fkern = chol_factor_kernel<15>; skern = chol_subst_kernel<2>; }
  else if (p->hbw <= 23) { fkern = chol_factor_kernel<23>; skern = chol_subst_kernel<3>; }
  else { fkern = chol_factor_kernel<31>; skern = chol_subst_kernel<4>; }
  const int stages = o.stages ? o.stages : 7;
  static const bool force_simt = std::getenv("FDM_CHOL_SIMT") != nullptr;   // A/B switch for profiling
  const bool mma = p->hbw <= 30 && !force_simt;                             // these invert the diagonal blocks themselves
  if (!o.reuse_factor && (stages & 1)) {
    size_t smem = (size_t)kFacWarps * kFacWarpSmem * sizeof(double);
    int fw = kFacWarps, dpc = kFacWarps;
    if (mma) {
      fkern = mma_panel4(p) ? chol_factor_mma_kernel<false, 4, false> : chol_factor_mma_kernel<false, 2, false>;
      A.twisted = mma_panel4(p) && use_twisted(p, std::max<int64_t>(batch, o.concurrent_designs)) ? 1 : 0;
      fw = mma_warps_per_cta(A.twisted && std::max<int64_t>(batch, o.concurrent_designs) <= 4 * (int64_t)std::max(p->sm_count, 1));
      if (A.twisted) fw = std::max(2, fw & ~1);
      dpc = A.twisted ? fw / 2 : fw;
      if (A.twisted) fkern = chol_factor_mma_kernel<false, 4, false, true>;
      smem = (size_t)fw * kMmaWarpSmem * sizeof(double) + fac_smem_pad();
    }
    FDM_CUDA(cudaFuncSetAttribute(fkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fkern<<<(unsigned)((batch + dpc - 1) / dpc), fw * 32, smem, st>>>(A);
    if (mma) {   // designs with a non-positive pivot (mixed-sign q): again, with signed pivots
      void (*skern)(WArgs) = A.twisted ? chol_factor_mma_kernel<false, 4, true, true>
                           : mma_panel4(p) ? chol_factor_mma_kernel<false, 4, true> : chol_factor_mma_kernel<false, 2, true>;
      FDM_CUDA(cudaFuncSetAttribute(skern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      skern<<<(unsigned)((batch + dpc - 1) / dpc), fw * 32, smem, st>>>(A);
    }
  }
  if (!o.reuse_factor && (stages & 2) && !mma) {
    const int64_t items = batch * A.nblk * 4 * 8;
    chol_invdiag_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(A);
  }
  if (stages & 4) {
    int rc = launch_backward(p, A, batch, st, std::max<int64_t>(batch, o.concurrent_designs));
    if (rc != FDM_OK) return rc;
    // (not for a solve with a stored factor: the factor was checked when it was made, and a sliced caller that never
    //  formed K.data passes a placeholder there)
    if (mma && !o.reuse_factor && (stages & 1)) launch_signed_check(p, A, batch, false, st);
  }
  return fdm_check_launch("chol_warp_solve");
}
