// Y = K X with K on the plan's CSR pattern (symmetric, so CSC data == CSR data) and X (Nf,3).
// Replaces `K @ xyz_free` (models.py:1012) and `A @ xk` (sparse.py:301).
//
// Layout: each CTA owns kRowsPerCta consecutive rows.  The CSR slice of those rows
// (values 8 B + column index 4 B per entry; contiguous in memory because rows are) is staged
// into shared memory with two TMA bulk copies (cp.async.bulk.shared::cluster.global, SASS
// UBLKCP) signalled through one mbarrier, so the 12 B/nnz stream is read with full-width
// coalesced transactions regardless of row length.  Rows are then processed one per thread:
// the thread walks its row's entries in shared memory in ascending order (the summation order
// of a sequential CSR product, hence deterministic), gathers the three coordinates of X[col]
// through the read-only path (neighbouring rows share columns, so the gathers hit L1/L2), and
// the results leave through a shared-memory tile so that Y is written with full-width stores.
// Algorithmic bytes: 12 nnz + 52 Nf.
#include <cuda_runtime.h>

#include <cstdlib>

#include "launch.h"

namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerCta = 256;                      // rows per CTA tile (one per thread)
constexpr int kMaxTileNnz = 1536;                     // staged entries per stage (18 KB; with the Y tile 25 KB => 8 CTAs per SM)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
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
// 1-D TMA bulk copy global -> shared; size and both addresses must be multiples of 16 B.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// One tile = rows [r0, r1) whose entries [k0, k1) fit the staging buffers.  Tiles are cut on the
// host side of the launch (fixed rows per CTA); a tile whose nnz exceeds the buffer is walked in
// several stages.
__global__ void __launch_bounds__(kThreads)
spmm_tma_kernel(int Nf, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const double* __restrict__ Kdata, const double* __restrict__ X, double* __restrict__ Y,
                int64_t nnz, const double* Klo, const double* Khi) {
  __shared__ __align__(16) double s_val[kMaxTileNnz + 2];
  __shared__ __align__(16) int32_t s_col[kMaxTileNnz + 4];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_rp[kRowsPerCta + 1];
  __shared__ __align__(16) double s_y[kRowsPerCta * 3];

  const int64_t b = blockIdx.y;
  Kdata += b * nnz;
  X += b * (int64_t)Nf * 3;
  Y += b * (int64_t)Nf * 3;

  const int r0 = blockIdx.x * kRowsPerCta;
  const int r1 = min(Nf, r0 + kRowsPerCta);
  for (int i = threadIdx.x; i <= r1 - r0; i += kThreads) s_rp[i] = __ldg(rowptr + r0 + i);
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t phase = 0;
  int row_lo = 0;                                         // local row where the current stage starts
  while (row_lo < r1 - r0) {
    // rows [row_lo, row_hi) : as many whole rows as fit the staging buffers
    const int kbeg = s_rp[row_lo];
    int row_hi = row_lo + 1;
    // (all threads compute the same bound; rows are short so this scan is cheap)
    {
      int lo = row_lo + 1, hi = r1 - r0;
      while (lo < hi) {                                   // largest row_hi with s_rp[row_hi]-kbeg <= cap
        int mid = (lo + hi + 1) >> 1;
        if (s_rp[mid] - kbeg <= kMaxTileNnz - 4) lo = mid; else hi = mid - 1;
      }
      row_hi = lo;
    }
    const int kend = s_rp[row_hi];
    const int cnt = kend - kbeg;
    // 16-byte aligned windows around [kbeg, kend) for the bulk copies.  K.data is caller memory:
    // the window must stay inside [Klo, Khi) (the whole batch array); colidx is plan memory,
    // allocated with 16 B of slack at the end.
    const uintptr_t va = (uintptr_t)(Kdata + kbeg), va0 = va & ~uintptr_t(15);
    const uintptr_t ve = ((uintptr_t)(Kdata + kend) + 15) & ~uintptr_t(15);
    const uintptr_t ca = (uintptr_t)(colidx + kbeg), ca0 = ca & ~uintptr_t(15);
    const uintptr_t ce = ((uintptr_t)(colidx + kend) + 15) & ~uintptr_t(15);
    const int voff = (int)((va - va0) >> 3), coff = (int)((ca - ca0) >> 2);
    const bool tma_ok = cnt > 0 && cnt <= kMaxTileNnz - 4 && va0 >= (uintptr_t)Klo && ve <= (uintptr_t)Khi;
    if (tma_ok) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(&s_bar, (uint32_t)((ve - va0) + (ce - ca0)));
        tma_load_1d(s_val, (const void*)va0, (uint32_t)(ve - va0), &s_bar);
        tma_load_1d(s_col, (const void*)ca0, (uint32_t)(ce - ca0), &s_bar);
      }
      mbar_wait(&s_bar, phase);
      phase ^= 1;
    }
    for (int r = row_lo + threadIdx.x; r < row_hi; r += kThreads) {
      const int a = s_rp[r] - kbeg, e = s_rp[r + 1] - kbeg;
      double ax = 0.0, ay = 0.0, az = 0.0;
      // four entries per trip: all twelve gathers are issued before the first FMA needs one
      // (entries past the row end are clamped to the row's first column with a zero value)
      for (int k = a; k < e; k += 4) {
        double v[4], x0[4], x1[4], x2[4];
        int c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool in = k + u < e;
          const int kk = in ? k + u : a;
          if (tma_ok) { v[u] = s_val[voff + kk]; c[u] = s_col[coff + kk]; }
          else { v[u] = __ldg(Kdata + kbeg + kk); c[u] = __ldg(colidx + kbeg + kk); }
          if (!in) v[u] = 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double* x = X + 3 * (int64_t)c[u];
          x0[u] = __ldg(x); x1[u] = __ldg(x + 1); x2[u] = __ldg(x + 2);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { ax += v[u] * x0[u]; ay += v[u] * x1[u]; az += v[u] * x2[u]; }
      }
      double* y = s_y + 3 * (r - row_lo);
      y[0] = ax; y[1] = ay; y[2] = az;
    }
    __syncthreads();
    {
      double* out = Y + 3 * (int64_t)(r0 + row_lo);
      for (int i = threadIdx.x; i < 3 * (row_hi - row_lo); i += kThreads) out[i] = s_y[i];
    }
    __syncthreads();                                      // staging buffers are reused
    row_lo = row_hi;
  }
}

// The same product for a BATCH of designs on one topology: a thread owns one row of the tile and keeps its extent and
// its column indices (<= kBatchRow, in registers) for the whole loop over `dpc` designs, two designs in flight at a
// time.  A row's entries are consecutive in K.data, so the warp's loads of a design's values cover a contiguous
// range with every sector used; X is gathered through the read-only path (a design's X is 20 KB); Y leaves as three
// doubles per thread, contiguous across the warp.  No shared memory, no barrier.  Each row sums its entries in
// ascending order in groups of four exactly as spmm_tma_kernel does => identical bits.
constexpr int kBatchRow = 12;                         // entries per row the registers cover (rows are checked on the host)

__global__ void __launch_bounds__(kThreads, 4)
spmm_batch_kernel(int Nf, int64_t nnz, int64_t batch, int dpc, const int32_t* __restrict__ rowptr,
                  const int32_t* __restrict__ colidx, const double* __restrict__ Kdata, const double* __restrict__ X,
                  double* __restrict__ Y) {
  const int r = blockIdx.x * kThreads + threadIdx.x;
  if (r >= Nf) return;
  const int a = __ldg(rowptr + r), len = __ldg(rowptr + r + 1) - a;
  int c[kBatchRow];
#pragma unroll
  for (int j = 0; j < kBatchRow; ++j) c[j] = __ldg(colidx + a + (j < len ? j : 0));   // past the end: the first column, value 0
  const int64_t b_lo = (int64_t)blockIdx.y * dpc, b_hi = min(batch, b_lo + dpc);
  for (int64_t b = b_lo; b < b_hi; b += 2) {
    const int64_t b1 = min(b + 1, b_hi - 1);
    const double* K0 = Kdata + b * nnz + a;
    const double* K1 = Kdata + b1 * nnz + a;
    const double* X0 = X + b * (int64_t)Nf * 3;
    const double* X1 = X + b1 * (int64_t)Nf * 3;
    double s0[3] = {0.0, 0.0, 0.0}, s1[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < kBatchRow; k += 4) {
      if (k >= len) break;
      double v0[4], v1[4], x0[4][3], x1[4][3];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool in = k + u < len;
        v0[u] = in ? __ldg(K0 + k + u) : 0.0;
        v1[u] = in ? __ldg(K1 + k + u) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double* p0 = X0 + 3 * (int64_t)c[k + u];
        const double* p1 = X1 + 3 * (int64_t)c[k + u];
        x0[u][0] = __ldg(p0); x0[u][1] = __ldg(p0 + 1); x0[u][2] = __ldg(p0 + 2);
        x1[u][0] = __ldg(p1); x1[u][1] = __ldg(p1 + 1); x1[u][2] = __ldg(p1 + 2);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s0[0] += v0[u] * x0[u][0]; s0[1] += v0[u] * x0[u][1]; s0[2] += v0[u] * x0[u][2];
        s1[0] += v1[u] * x1[u][0]; s1[1] += v1[u] * x1[u][1]; s1[2] += v1[u] * x1[u][2];
      }
    }
    double* y0 = Y + (b * (int64_t)Nf + r) * 3;
    y0[0] = s0[0]; y0[1] = s0[1]; y0[2] = s0[2];
    if (b + 1 < b_hi) {
      double* y1 = Y + ((b + 1) * (int64_t)Nf + r) * 3;
      y1[0] = s1[0]; y1[1] = s1[1]; y1[2] = s1[2];
    }
  }
}

}  // namespace

int launch_spmm(const fdm_plan* p, const double* Kdata, const double* X, double* Y, int64_t batch,
                cudaStream_t st) {
  const unsigned tiles = (unsigned)((p->Nf + kRowsPerCta - 1) / kRowsPerCta);
  if (batch >= 8 && p->spmm_batch_ok && !(std::getenv("FDM_SPMM_BATCH") && std::atoi(std::getenv("FDM_SPMM_BATCH")) == 0)) {
    // about eight resident waves of CTAs (4 per SM: the kernel's launch bound), each walking an even number of designs
    const int64_t target = (int64_t)std::max(p->sm_count, 1) * 4 * 8;
    int dpc = (int)std::min<int64_t>(64, std::max<int64_t>(2, 2 * ((batch * tiles + 2 * target - 1) / (2 * target))));
    if (const char* e = std::getenv("FDM_SPMM_DPC")) dpc = std::max(2, std::atoi(e) & ~1);
    spmm_batch_kernel<<<dim3(tiles, (unsigned)((batch + dpc - 1) / dpc)), kThreads, 0, st>>>(
        (int)p->Nf, p->nnz, batch, dpc, p->d.rowptr, p->d.colidx, Kdata, X, Y);
    return fdm_check_launch("spmm_batch");
  }
  FDM_BATCH_LOOP(batch, b0, nb) {
    spmm_tma_kernel<<<dim3(tiles, (unsigned)nb), kThreads, 0, st>>>(
        (int)p->Nf, p->d.rowptr, p->d.colidx, Kdata + b0 * p->nnz, X + b0 * p->Nf * 3,
        Y + b0 * p->Nf * 3, p->nnz, Kdata, Kdata + batch * p->nnz);
  }
  return fdm_check_launch("spmm");
}
