// K1 (assembly), positions, K4 (state + VJP), K3 (adjoint gradients), loss helper.
// All HBM-bound gather/segmented-sum kernels: FP64 values, int32 indices, one design per
// blockIdx.y.  Every sum runs over a precomputed incidence list in ascending edge id, so
// results are deterministic and independent of the launch geometry.
#include <cuda_runtime.h>

#include "plan.h"
#include <cstdlib>

#include "launch.h"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ double3 ld3(const double* __restrict__ p, int64_t row) {
  const double* q = p + 3 * row;
  return make_double3(__ldg(q), __ldg(q + 1), __ldg(q + 2));
}
__device__ __forceinline__ void st3(double* p, int64_t row, double3 v) {
  double* q = p + 3 * row;
  q[0] = v.x; q[1] = v.y; q[2] = v.z;
}

// ---------------------------------------------------------------------------------------
// K1: K.data and rhs.   models.py:1064-1079, 828-856, 949-976
// One CTA owns kAsmRows consecutive free rows, i.e. ONE CONTIGUOUS slice [k0, k1) of K.data:
//   phase 1  thread per stored entry  : s_val[k] = -q[entry_edge[k]]            (off-diagonal)
//   phase 2  thread per row           : s_val[diag] = sum_inc q_e  over `diags` (ascending edge id, the
//                                       order of the reference's BCSR product, structures.py:316-350)
//                                       s_rhs[row] = loads[node] (+ sum_{e to support s} q_e x_s for the
//                                       rows flagged in sup_mask: they walk the node's incidence list)
//   phase 3  coalesced stores of the staged slice of K.data and of rhs (every line written once, full width)
// Index streams (entry_edge, diags, rowptr, free_node) are read once, coalesced; q is gathered.
// Algorithmic bytes per design: 16E + 12nnz + 52Nf + 24Ns.
// ---------------------------------------------------------------------------------------
constexpr int kAsmRows = 256;
constexpr int kAsmCap = 2560;   // staged entries per tile (20 KB); tiles with more fall back to direct stores

__global__ void __launch_bounds__(kThreads, 8)   // 8 CTAs per SM: the kernel lives on loads in flight
assemble_kernel(int nnz, int Nf, int N, int Ns,
                const int32_t* __restrict__ rowptr, const int32_t* __restrict__ entry_edge,
                const int32_t* __restrict__ diag_pos, const int32_t* __restrict__ free_node,
                const int32_t* __restrict__ diags_ptr, const int32_t* __restrict__ diags_idx,
                const uint32_t* __restrict__ sup_mask, const int32_t* __restrict__ node_slot,
                const int32_t* __restrict__ ninc_ptr, const int32_t* __restrict__ ninc_edge,
                const int32_t* __restrict__ ninc_other,
                const double* __restrict__ q, const double* __restrict__ xs,
                const double* __restrict__ loads, double* __restrict__ Kdata,
                double* __restrict__ rhs, int64_t q_stride, int64_t xs_stride, int64_t loads_stride) {
  __shared__ __align__(16) double s_val[kAsmCap];
  __shared__ __align__(16) double s_rhs[kAsmRows * 3];
  const int64_t b = blockIdx.y;
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * kAsmRows, r1 = min(Nf, r0 + kAsmRows);
  const int k0 = __ldg(rowptr + r0), k1 = __ldg(rowptr + r1), cnt = k1 - k0;
  const bool staged = cnt <= kAsmCap;
  q += b * q_stride;
  double* Kd = Kdata ? Kdata + b * (int64_t)nnz : nullptr;
  if (Kd) {
    // four entries per trip so that the index loads, then the q gathers, are in flight together
    for (int kb = tid; kb < cnt; kb += 4 * kThreads) {
      int e[4];
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = kb + u * kThreads;
        e[u] = k < cnt ? __ldg(entry_edge + k0 + k) : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = e[u] >= 0 ? -__ldg(q + e[u]) : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = kb + u * kThreads;
        if (e[u] >= 0) {
          if (staged) s_val[k] = v[u]; else Kd[k0 + k] = v[u];
        }
      }
    }
  }
  const int t = r0 + tid;
  if (t < r1) {
    if (Kd) {
      double diag = 0.0;
      const int d0 = __ldg(diags_ptr + t), d1 = __ldg(diags_ptr + t + 1);
      for (int k = d0; k < d1; k += 4) {   // same ascending order of additions, gathers issued four at a time
        int ed[4];
        double qv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) ed[u] = k + u < d1 ? __ldg(diags_idx + k + u) : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) qv[u] = ed[u] >= 0 ? __ldg(q + ed[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (ed[u] >= 0) diag += qv[u];
      }
      const int dp = __ldg(diag_pos + t);
      if (staged) s_val[dp - k0] = diag; else Kd[dp] = diag;
    }
    if (rhs) {
      const int n = __ldg(free_node + t);
      double3 acc = make_double3(0.0, 0.0, 0.0);
      if ((__ldg(sup_mask + (t >> 5)) >> (t & 31)) & 1u) {
        const double* xsb = xs + b * xs_stride;
        const int i0 = __ldg(ninc_ptr + n), i1 = __ldg(ninc_ptr + n + 1);
        for (int k = i0; k < i1; ++k) {
          const int slot = __ldg(node_slot + __ldg(ninc_other + k));
          if (slot < 0) {
            const double qe = __ldg(q + (__ldg(ninc_edge + k) >> 1));
            const double3 x = ld3(xsb, -slot - 1);
            acc.x += qe * x.x; acc.y += qe * x.y; acc.z += qe * x.z;
          }
        }
      }
      const double3 p = ld3(loads + b * loads_stride, n);
      s_rhs[3 * tid] = p.x + acc.x; s_rhs[3 * tid + 1] = p.y + acc.y; s_rhs[3 * tid + 2] = p.z + acc.z;
    }
  }
  __syncthreads();
  if (Kd && staged)
    for (int k = tid; k < cnt; k += kThreads) Kd[k0 + k] = s_val[k];
  if (rhs) {
    double* out = rhs + (b * (int64_t)Nf + r0) * 3;
    for (int i = tid; i < (r1 - r0) * 3; i += kThreads) out[i] = s_rhs[i];
  }
}

// ---------------------------------------------------------------------------------------
// K1 for a BATCH of designs on one topology (the vmap of loss_plotter.py:98-103), as streaming kernels.  The
// per-design kernel above is bound by instruction issue and by its q gathers -- every stored entry is an 8-byte load
// at a random place of the design's q, 32 sectors per warp request, behind a chain of index loads -- not by the
// K.data it writes.  Here a CTA owns WHOLE designs, one after another, and reads every index stream once per launch.
//
// assemble_K_batch_kernel (K.data):
//   * shared memory holds two regions [diag (Nf) | q (E) | 0.0], one per design in flight.  A design's q arrives as
//     one coalesced asynchronous copy (cp.async) while the previous design is processed;
//   * pass 1, a thread per row: the row's incident edges (<= 8, as byte offsets into the region, padded with the
//     offset of the 0.0 slot; two 16-byte shared loads) -> the diagonal, summed in list order as `diags @ q` does,
//     stored NEGATED in the region;
//   * pass 2, a thread per stored entry: the entry's region offset sits in a REGISTER (NPT entries per thread, fixed
//     for the whole launch); value = -region[offset] (-q on an off-diagonal, -(-sum) on the diagonal: negation is
//     exact), stored in entry order: full-width stores, every byte of K.data written once.
// rhs: an element of a row without supports is load + 0.0 -- registers for the whole launch when all designs share
//   the loads -- so assemble_rhs_stream_kernel is full-width stores only; the few rows next to supports (load + sum
//   of q_e * x_support in incidence order) are written afterwards by assemble_K_batch_kernel, which has q in shared
//   memory, or by assemble_rhs_pairs_kernel.
// Same additions in the same order as assemble_kernel => bit-identical output (adding the +0.0 pads never changes a
// sum that starts from +0.0).  Topologies that do not fit (plan: asm_batch_ok) keep the per-design kernel.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 lds_v4(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ double flip_sign(double v) {
  return __hiloint2double(__double2hiint(v) ^ (int)0x80000000, __double2loint(v));
}

constexpr int kAsmNR = 4;                      // rows per thread: Nf <= 4 * 256
constexpr int kAsmNS = 2;                      // supports per row the registers cover

// SUP: the CTA also writes the rhs rows next to supports (q is in shared memory here): thread t owns the t-th of them
// (<= 256 such rows, <= kAsmNS supports per row; plan: asm_sup_in_K), indices in registers -- and the load and the
// support coordinates too when all designs share them; the rhs has been filled before on the same stream.
template <int NPT, bool SUP>
__global__ void __launch_bounds__(kThreads, 3)
assemble_K_batch_kernel(int E, int Nf, int nnz, int maxdeg, int64_t batch, const int32_t* __restrict__ asm_code,
                        const int32_t* __restrict__ diags_ptr, const int32_t* __restrict__ diags_idx,
                        const double* __restrict__ q, double* __restrict__ Kdata, int64_t q_stride, int nsup_rows,
                        const int32_t* __restrict__ rsup_rows, const int32_t* __restrict__ rsup_ptr,
                        const int32_t* __restrict__ rsup_edge, const int32_t* __restrict__ rsup_slot,
                        const int32_t* __restrict__ free_node, const double* __restrict__ xs,
                        const double* __restrict__ loads, double* __restrict__ rhs, int64_t xs_stride,
                        int64_t loads_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NfP = (Nf + 1) & ~1;                                // q starts 16-byte aligned
  const int region = NfP + ((E + 2) & ~1);                      // doubles per region: diag (| pad) | q | 0.0 (| pad)
  double* s_reg = reinterpret_cast<double*>(smem_raw);          // [2][region]
  int4* s_de = reinterpret_cast<int4*>(s_reg + 2 * region);     // [Nf][2]: eight byte offsets into a region
  const int tid = threadIdx.x;
  // designs blockIdx.x, blockIdx.x + gridDim.x, ...: the grid is one resident wave, so every SM gets the same number
  // of designs to within one
  const int64_t b_lo = blockIdx.x, b_step = gridDim.x, b_hi = batch;
  if (b_lo >= b_hi) return;
  const uint32_t reg0 = (uint32_t)__cvta_generic_to_shared(s_reg);
  const uint32_t de0 = (uint32_t)__cvta_generic_to_shared(s_de) + 32u * tid;
  const bool vec = (E & 1) == 0 && (q_stride & 1) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0;
  auto fetch_q = [&](int64_t b, int buf) {
    const double* src = q + b * q_stride;
    const uint32_t dst = reg0 + 8u * (buf * region + NfP);
    if (vec) {
      for (int i = tid; i < E / 2; i += kThreads)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 2 * i) : "memory");
    } else {
      for (int i = tid; i < E; i += kThreads)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * i), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch_q(b_lo, 0);
  // ---- once per CTA: the index streams ----
  const int zero_off = 8 * (NfP + E);
  if (tid < 2) s_reg[tid * region + NfP + E] = 0.0;
#pragma unroll
  for (int i = 0; i < kAsmNR; ++i) {
    const int r = tid + i * kThreads;
    if (r < Nf) {
      int o[8];
      const int d0 = __ldg(diags_ptr + r), d1 = __ldg(diags_ptr + r + 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = d0 + j < d1 ? 8 * (NfP + __ldg(diags_idx + d0 + j)) : zero_off;
      s_de[2 * r] = make_int4(o[0], o[1], o[2], o[3]);
      s_de[2 * r + 1] = make_int4(o[4], o[5], o[6], o[7]);
    }
  }
  int off[NPT];                                                 // entry -> byte offset into a region, -1 past the end
#pragma unroll
  for (int u = 0; u < NPT; ++u) {
    const int k = tid + u * kThreads;
    off[u] = -1;
    if (k < nnz) {
      const int c = __ldg(asm_code + k);
      off[u] = c >= 0 ? 8 * (NfP + c) : 8 * (-c - 2);
    }
  }
  // the row next to supports this thread owns, if any: its place in the rhs, its node, its (q offset, support slot) terms
  int srow = -1, snode = 0, pe[kAsmNS], ps[kAsmNS];
  double3 xe[kAsmNS], pp = make_double3(0.0, 0.0, 0.0);
  const bool sup_shared = xs_stride == 0 && loads_stride == 0;
  if (SUP) {
#pragma unroll
    for (int j = 0; j < kAsmNS; ++j) { pe[j] = zero_off; ps[j] = -1; xe[j] = make_double3(0.0, 0.0, 0.0); }
    if (tid < nsup_rows) {
      srow = __ldg(rsup_rows + tid);
      snode = __ldg(free_node + srow);
      const int k0 = __ldg(rsup_ptr + srow), k1 = __ldg(rsup_ptr + srow + 1);
#pragma unroll
      for (int j = 0; j < kAsmNS; ++j)
        if (k0 + j < k1) { pe[j] = 8 * (NfP + __ldg(rsup_edge + k0 + j)); ps[j] = __ldg(rsup_slot + k0 + j); }
      if (sup_shared) {
        pp = ld3(loads, snode);
#pragma unroll
        for (int j = 0; j < kAsmNS; ++j)
          if (ps[j] >= 0) xe[j] = ld3(xs, ps[j]);
      }
    }
  }
  // launched as a programmatic dependent of the rhs kernel (launch_assemble): everything above ran while that grid was
  // finishing; the rows next to supports written below must land after its stores (a no-op in an ordinary launch)
  if (SUP) asm volatile("griddepcontrol.wait;" ::: "memory");
  int buf = 0;
  for (int64_t b = b_lo; b < b_hi; b += b_step, buf ^= 1) {
    if (SUP && !sup_shared && srow >= 0) {   // this design's load and support coordinates: issued before the wait for q
      pp = ld3(loads + b * loads_stride, snode);
#pragma unroll
      for (int j = 0; j < kAsmNS; ++j)
        if (ps[j] >= 0) xe[j] = ld3(xs + b * xs_stride, ps[j]);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();            // q of design b is in place for everyone; everyone is done with the other region
    if (b + b_step < b_hi) fetch_q(b + b_step, buf ^ 1);
    const uint32_t rb = reg0 + 8u * (buf * region);
    if (SUP && srow >= 0) {
      double3 acc = make_double3(0.0, 0.0, 0.0);
#pragma unroll
      for (int j = 0; j < kAsmNS; ++j)
        if (ps[j] >= 0) {
          const double qe = lds_f64(rb + pe[j]);
          acc.x += qe * xe[j].x; acc.y += qe * xe[j].y; acc.z += qe * xe[j].z;
        }
      st3(rhs + b * (int64_t)Nf * 3, srow, make_double3(pp.x + acc.x, pp.y + acc.y, pp.z + acc.z));
    }
#pragma unroll
    for (int i = 0; i < kAsmNR; ++i) {
      if (tid + i * kThreads < Nf) {
        // `maxdeg` (the longest list of the topology) is uniform: the pad slots beyond it cost nothing
        const int4 a = lds_v4(de0 + 32u * kThreads * i);
        double acc = 0.0;
        acc += lds_f64(rb + a.x);
        if (maxdeg > 1) acc += lds_f64(rb + a.y);
        if (maxdeg > 2) acc += lds_f64(rb + a.z);
        if (maxdeg > 3) acc += lds_f64(rb + a.w);
        if (maxdeg > 4) {
          const int4 c = lds_v4(de0 + 32u * kThreads * i + 16u);
          acc += lds_f64(rb + c.x);
          if (maxdeg > 5) acc += lds_f64(rb + c.y);
          if (maxdeg > 6) acc += lds_f64(rb + c.z);
          if (maxdeg > 7) acc += lds_f64(rb + c.w);
        }
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(rb + 8u * (tid + i * kThreads)), "d"(flip_sign(acc)) : "memory");
      }
    }
    __syncthreads();            // the diagonals are in place
    double* Kd = Kdata + b * (int64_t)nnz + tid;
#pragma unroll
    for (int u = 0; u < NPT; ++u)
      if (off[u] >= 0) Kd[u * kThreads] = flip_sign(lds_f64(rb + off[u]));
  }
}

// rhs, step 1: every element as load + 0.0 -- final for the rows without supports, provisional for the others (step 2
// overwrites them later on the same stream).  rhs_src[i] = index of flat element i's load in a design's loads (plan).
//
// assemble_rhs_tma_kernel -- shared loads (loads_stride == 0), 16-byte aligned rhs: every design's image is the same
// 3 Nf doubles.  A CTA builds the image of TWO consecutive designs in shared memory once (2 * 3 Nf doubles: 16-byte
// granular whatever the parity of 3 Nf) and one thread then issues a bulk asynchronous copy (TMA, shared -> global) per
// pair of its designs: no registers, no store instructions, the copy engine streams the bytes.
// assemble_rhs_stream_kernel -- per-design loads: plain stores, one design at a time.
constexpr int kRhsThreads = 512;

__global__ void __launch_bounds__(kRhsThreads, 1)
assemble_rhs_tma_kernel(int Nf, int64_t batch, int dpc, const int32_t* __restrict__ rhs_src,
                        const double* __restrict__ loads, double* __restrict__ rhs) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* img = reinterpret_cast<double*>(smem_raw);      // [2 * 3 Nf]
  // programmatic dependent launch: the K.data kernel behind this one may start its CTAs now (they run their prologue
  // and wait, `griddepcontrol.wait`, for this grid to finish before they touch the rhs)
  asm volatile("griddepcontrol.launch_dependents;");
  const int tid = threadIdx.x;
  const int n3 = 3 * Nf;
  const int64_t b_lo = (int64_t)blockIdx.x * dpc;         // dpc is even: b_lo * n3 * 8 is a multiple of 16
  const int64_t b_hi = min(batch, b_lo + dpc);
  const int64_t b_pairs = b_lo + ((b_hi - b_lo) & ~int64_t(1));   // [b_lo, b_pairs): whole pairs of designs
  for (int i = tid; i < n3; i += kRhsThreads) {
    const double v = __ldg(loads + __ldg(rhs_src + i)) + 0.0;
    img[i] = v;
    img[i + n3] = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the generic-proxy writes above, seen by the copy engine
  __syncthreads();
  if (b_pairs < b_hi) {                                   // an odd design left over at the end of the batch
    double* out = rhs + b_pairs * (int64_t)n3;
    for (int i = tid; i < n3; i += kRhsThreads) out[i] = img[i];
  }
  if (tid == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(img);
    for (int64_t b = b_lo; b < b_pairs; b += 2)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(rhs + b * (int64_t)n3), "r"(src),
                   "r"((uint32_t)(16 * n3))
                   : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the image must outlive the reads
  }
}

__global__ void __launch_bounds__(kRhsThreads, 2)
assemble_rhs_stream_kernel(int Nf, int64_t batch, int dpc, const int32_t* __restrict__ rhs_src,
                           const double* __restrict__ loads, double* __restrict__ rhs, int64_t loads_stride) {
  const int tid = threadIdx.x;
  const int n3 = 3 * Nf;
  const int64_t b_lo = (int64_t)blockIdx.x * dpc, b_hi = min(batch, b_lo + dpc);
  for (int64_t b = b_lo; b < b_hi; ++b) {
    double* out = rhs + b * (int64_t)n3;
    for (int i = tid; i < n3; i += kRhsThreads) out[i] = __ldg(loads + b * loads_stride + __ldg(rhs_src + i)) + 0.0;
  }
}

// rhs, step 2 when assemble_K_batch_kernel<., true> does not run: the (design, support row, component) items, four
// per thread per trip so that their short chains of index loads and their gathers are in flight together.
__global__ void __launch_bounds__(kRhsThreads, 1)
assemble_rhs_pairs_kernel(int Nf, int nsup_rows, int64_t batch, int dpp, const int32_t* __restrict__ rsup_rows,
                          const int32_t* __restrict__ rsup_ptr, const int32_t* __restrict__ rsup_edge,
                          const int32_t* __restrict__ rsup_slot, const int32_t* __restrict__ free_node,
                          const double* __restrict__ q, const double* __restrict__ xs, const double* __restrict__ loads,
                          double* __restrict__ rhs, int64_t q_stride, int64_t xs_stride, int64_t loads_stride) {
  const int tid = threadIdx.x;
  const int n3 = 3 * Nf, npairs = 3 * nsup_rows;
  const int64_t b0 = (int64_t)blockIdx.x * dpp;
  const int nitems = (int)min((int64_t)dpp, batch - b0) * npairs;
  for (int base = tid; base < nitems; base += 4 * kRhsThreads) {
    int r[4], c[4], k0[4], k1[4];
    int64_t b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int it = min(base + u * kRhsThreads, nitems - 1);       // past the end: the last item again, not stored
      const int db = it / npairs, pi = it - db * npairs;
      b[u] = b0 + db;
      r[u] = __ldg(rsup_rows + pi / 3);
      c[u] = pi % 3;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { k0[u] = __ldg(rsup_ptr + r[u]); k1[u] = __ldg(rsup_ptr + r[u] + 1); }
    double qe[4][2], xe[4][2], pp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const bool in = k0[u] + j < k1[u];
        qe[u][j] = in ? __ldg(q + b[u] * q_stride + __ldg(rsup_edge + k0[u] + j)) : 0.0;
        xe[u][j] = in ? __ldg(xs + b[u] * xs_stride + 3 * __ldg(rsup_slot + k0[u] + j) + c[u]) : 0.0;
      }
      pp[u] = __ldg(loads + b[u] * loads_stride + 3 * __ldg(free_node + r[u]) + c[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (k0[u] + j < k1[u]) acc += qe[u][j] * xe[u][j];
      for (int k = k0[u] + 2; k < k1[u]; ++k)                        // rows with more than two supports
        acc += __ldg(q + b[u] * q_stride + __ldg(rsup_edge + k)) * __ldg(xs + b[u] * xs_stride + 3 * __ldg(rsup_slot + k) + c[u]);
      if (base + u * kRhsThreads < nitems) rhs[b[u] * (int64_t)n3 + 3 * r[u] + c[u]] = pp[u] + acc;
    }
  }
}

// ---------------------------------------------------------------------------------------
// positions  (models.py:195-220)  and transpose
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
positions_kernel(int N, int Nf, int Ns, const int32_t* __restrict__ node_slot,
                 double* __restrict__ x_free, double* __restrict__ xs, double* __restrict__ xyz,
                 int64_t xs_stride, int transpose) {
  const int64_t b = blockIdx.y;
  const int n = blockIdx.x * kThreads + threadIdx.x;
  if (n >= N) return;
  const int slot = __ldg(node_slot + n);
  double* xf = x_free ? x_free + b * (int64_t)Nf * 3 : nullptr;
  double* xsb = xs ? xs + b * xs_stride : nullptr;
  double* xb = xyz + b * (int64_t)N * 3;
  if (!transpose) {
    st3(xb, n, slot >= 0 ? ld3(xf, slot) : ld3(xsb, -slot - 1));
  } else {
    const double3 v = ld3(xb, n);
    if (slot >= 0) { if (xf) st3(xf, slot, v); }
    else if (xsb) st3(xsb, -slot - 1, v);
  }
}

// ---------------------------------------------------------------------------------------
// K4 forward: equilibrium_state (models.py:753-794).  Bytes per design: 56E + 72N.
//   thread t < E: vectors, lengths, forces       thread t < N: residuals
// residual_n = loads_n - sum_inc q_e (xyz_n - xyz_other)   (sign of C^T folds into the difference)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
state_kernel(int E, int N, const int32_t* __restrict__ edge_nodes,
             const int32_t* __restrict__ ninc_ptr, const int32_t* __restrict__ ninc_edge,
             const int32_t* __restrict__ ninc_other,
             const double* __restrict__ q, const double* __restrict__ xyz,
             const double* __restrict__ loads, double* __restrict__ vectors,
             double* __restrict__ lengths, double* __restrict__ forces,
             double* __restrict__ residuals, int64_t q_stride, int64_t loads_stride) {
  const int64_t b = blockIdx.y;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  q += b * q_stride;
  xyz += b * (int64_t)N * 3;
  if (t < E && (vectors || lengths || forces)) {
    const int2 th = __ldg(reinterpret_cast<const int2*>(edge_nodes) + t);
    const double3 a = ld3(xyz, th.x), c = ld3(xyz, th.y);
    const double3 v = make_double3(c.x - a.x, c.y - a.y, c.z - a.z);
    if (vectors) st3(vectors + b * (int64_t)E * 3, t, v);
    if (lengths || forces) {
      const double len = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
      if (lengths) lengths[b * (int64_t)E + t] = len;
      if (forces) forces[b * (int64_t)E + t] = __ldg(q + t) * len;
    }
  }
  if (t < N && residuals) {
    const double3 xn = ld3(xyz, t);
    double3 r = ld3(loads + b * loads_stride, t);
    const int k0 = __ldg(ninc_ptr + t), k1 = __ldg(ninc_ptr + t + 1);
    for (int k = k0; k < k1; ++k) {
      const double qe = __ldg(q + (__ldg(ninc_edge + k) >> 1));
      const double3 xo = ld3(xyz, __ldg(ninc_other + k));
      r.x -= qe * (xn.x - xo.x); r.y -= qe * (xn.y - xo.y); r.z -= qe * (xn.z - xo.z);
    }
    st3(residuals + b * (int64_t)N * 3, t, r);
  }
}

// ---------------------------------------------------------------------------------------
// K4 forward mode (JVP of models.py:780-794): tangents of the state for tangents (q_dot, xyz_dot, loads_dot) at the
// primal (q, xyz).  T tangents may share one primal (primal strides 0).
//   per edge e=(t,h):  v = x_h - x_t, vd = xd_h - xd_t, l = |v|:  vectors_dot = vd, lengths_dot = v.vd / l,
//                      forces_dot = q_dot l + q lengths_dot
//   per node n:        residuals_dot_n = loads_dot_n - sum_inc [ q_dot_e (x_n - x_o) + q_e (xd_n - xd_o) ]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
state_jvp_kernel(int E, int N, const int32_t* __restrict__ edge_nodes, const int32_t* __restrict__ ninc_ptr,
                 const int32_t* __restrict__ ninc_edge, const int32_t* __restrict__ ninc_other,
                 const double* __restrict__ q, const double* __restrict__ xyz, const double* __restrict__ q_dot,
                 const double* __restrict__ xyz_dot, const double* __restrict__ loads_dot,
                 double* __restrict__ vectors_dot, double* __restrict__ lengths_dot, double* __restrict__ forces_dot,
                 double* __restrict__ residuals_dot, int64_t q_stride, int64_t xyz_stride, int64_t q_dot_stride,
                 int64_t loads_dot_stride) {
  const int64_t b = blockIdx.y;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  q += b * q_stride;
  xyz += b * xyz_stride;
  if (q_dot) q_dot += b * q_dot_stride;
  xyz_dot += b * (int64_t)N * 3;
  if (t < E && (vectors_dot || lengths_dot || forces_dot)) {
    const int2 th = __ldg(reinterpret_cast<const int2*>(edge_nodes) + t);
    const double3 a = ld3(xyz, th.x), c = ld3(xyz, th.y), ad = ld3(xyz_dot, th.x), cd = ld3(xyz_dot, th.y);
    const double3 v = make_double3(c.x - a.x, c.y - a.y, c.z - a.z);
    const double3 vd = make_double3(cd.x - ad.x, cd.y - ad.y, cd.z - ad.z);
    if (vectors_dot) st3(vectors_dot + b * (int64_t)E * 3, t, vd);
    if (lengths_dot || forces_dot) {
      const double len = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
      const double ld = (v.x * vd.x + v.y * vd.y + v.z * vd.z) / len;   // NaN on a zero-length edge, like jnp.linalg.norm's JVP
      if (lengths_dot) lengths_dot[b * (int64_t)E + t] = ld;
      if (forces_dot) forces_dot[b * (int64_t)E + t] = (q_dot ? __ldg(q_dot + t) : 0.0) * len + __ldg(q + t) * ld;
    }
  }
  if (t < N && residuals_dot) {
    const double3 xn = ld3(xyz, t), xdn = ld3(xyz_dot, t);
    double3 r = loads_dot ? ld3(loads_dot + b * loads_dot_stride, t) : make_double3(0.0, 0.0, 0.0);
    const int k0 = __ldg(ninc_ptr + t), k1 = __ldg(ninc_ptr + t + 1);
    for (int k = k0; k < k1; ++k) {
      const int e = __ldg(ninc_edge + k) >> 1, o = __ldg(ninc_other + k);
      const double qe = __ldg(q + e), qd = q_dot ? __ldg(q_dot + e) : 0.0;
      const double3 xo = ld3(xyz, o), xdo = ld3(xyz_dot, o);
      r.x -= qd * (xn.x - xo.x) + qe * (xdn.x - xdo.x);
      r.y -= qd * (xn.y - xo.y) + qe * (xdn.y - xdo.y);
      r.z -= qd * (xn.z - xo.z) + qe * (xdn.z - xdo.z);
    }
    st3(residuals_dot + b * (int64_t)N * 3, t, r);
  }
}

// ---------------------------------------------------------------------------------------
// K4 reverse (autodiff transpose of models.py:780-794; SURVEY.md Appendix B "direct" terms)
//   per edge e=(t,h):  v = x_h - x_t, l = |v|, Crb = rbar_h - rbar_t
//       q_bar_e  = fbar_e * l - Crb . v
//       vtot_e   = vbar_e + (lbar_e + q_e fbar_e) v / l - q_e Crb
//   per node n:        xyz_bar_n = xyz_cot_n + sum_inc (+vtot_e if n is head, -vtot_e if tail)
//                      loads_bar_n = loads_cot_n + rbar_n
// Node threads recompute vtot of their incident edges, so no (E,3) temporary and no atomics.
// ---------------------------------------------------------------------------------------
struct EdgeCot {
  const double *rbar, *lbar, *fbar, *vbar;
};

__device__ __forceinline__ double3 edge_vtot(const EdgeCot& c, int e, int t, int h,
                                             const double* __restrict__ xyz, double qe,
                                             double* qbar_out) {
  const double3 a = ld3(xyz, t), d = ld3(xyz, h);
  const double3 v = make_double3(d.x - a.x, d.y - a.y, d.z - a.z);
  double3 crb = make_double3(0.0, 0.0, 0.0);
  if (c.rbar) {
    const double3 rt = ld3(c.rbar, t), rh = ld3(c.rbar, h);
    crb = make_double3(rh.x - rt.x, rh.y - rt.y, rh.z - rt.z);
  }
  double3 vt = c.vbar ? ld3(c.vbar, e) : make_double3(0.0, 0.0, 0.0);
  vt.x -= qe * crb.x; vt.y -= qe * crb.y; vt.z -= qe * crb.z;
  const double fb = c.fbar ? __ldg(c.fbar + e) : 0.0;
  const double ltot = (c.lbar ? __ldg(c.lbar + e) : 0.0) + qe * fb;
  double len = 0.0;
  if (ltot != 0.0 || (qbar_out && fb != 0.0)) len = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  if (ltot != 0.0) {
    const double s = ltot / len;  // len == 0 -> NaN, as jnp.linalg.norm's gradient
    vt.x += s * v.x; vt.y += s * v.y; vt.z += s * v.z;
  }
  if (qbar_out) *qbar_out = fb * len - (crb.x * v.x + crb.y * v.y + crb.z * v.z);
  return vt;
}

__global__ void __launch_bounds__(kThreads)
state_vjp_kernel(int E, int N, const int32_t* __restrict__ edge_nodes,
                 const int32_t* __restrict__ ninc_ptr, const int32_t* __restrict__ ninc_edge,
                 const int32_t* __restrict__ ninc_other,
                 const double* __restrict__ q, const double* __restrict__ xyz,
                 const double* __restrict__ xyz_cot, const double* __restrict__ rbar,
                 const double* __restrict__ lbar, const double* __restrict__ fbar,
                 const double* __restrict__ loads_cot, const double* __restrict__ vbar,
                 double* __restrict__ q_bar, double* __restrict__ xyz_bar,
                 double* __restrict__ loads_bar, int64_t q_stride) {
  const int64_t b = blockIdx.y;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  q += b * q_stride;
  xyz += b * (int64_t)N * 3;
  EdgeCot c;
  c.rbar = rbar ? rbar + b * (int64_t)N * 3 : nullptr;
  c.lbar = lbar ? lbar + b * (int64_t)E : nullptr;
  c.fbar = fbar ? fbar + b * (int64_t)E : nullptr;
  c.vbar = vbar ? vbar + b * (int64_t)E * 3 : nullptr;
  if (t < E && q_bar) {
    const int2 th = __ldg(reinterpret_cast<const int2*>(edge_nodes) + t);
    double qb;
    edge_vtot(c, t, th.x, th.y, xyz, __ldg(q + t), &qb);
    q_bar[b * (int64_t)E + t] = qb;
  }
  if (t < N) {
    if (xyz_bar) {
      double3 acc = xyz_cot ? ld3(xyz_cot + b * (int64_t)N * 3, t) : make_double3(0.0, 0.0, 0.0);
      if (c.rbar || c.lbar || c.fbar || c.vbar) {
        const int k0 = __ldg(ninc_ptr + t), k1 = __ldg(ninc_ptr + t + 1);
        for (int k = k0; k < k1; ++k) {
          const int code = __ldg(ninc_edge + k);
          const int e = code >> 1, o = __ldg(ninc_other + k);
          const bool is_head = code & 1;
          const double3 vt = edge_vtot(c, e, is_head ? o : t, is_head ? t : o, xyz, __ldg(q + e), nullptr);
          const double s = is_head ? 1.0 : -1.0;
          acc.x += s * vt.x; acc.y += s * vt.y; acc.z += s * vt.z;
        }
      }
      st3(xyz_bar + b * (int64_t)N * 3, t, acc);
    }
    if (loads_bar) {
      double3 l = loads_cot ? ld3(loads_cot + b * (int64_t)N * 3, t) : make_double3(0.0, 0.0, 0.0);
      if (c.rbar) { const double3 r = ld3(c.rbar, t); l.x += r.x; l.y += r.y; l.z += r.z; }
      st3(loads_bar + b * (int64_t)N * 3, t, l);
    }
  }
}

// ---------------------------------------------------------------------------------------
// K3: gradients of the implicit solve (sparse.py:299-308 + transposes of models.py:1069-1077,
// 856, 973-976).  Bytes per design: 24E + 24Nf + 48N + 24Ns.
//   thread t < E: q_bar_e (+)= -(lam_h - lam_t).(x_h - x_t), lam = 0 on supports
//   thread t < N: free  -> loads_bar_n (+)= lam[slot]
//                 fixed -> xyz_fixed_bar_s (+)= sum_{inc, other free} q_e lam[other]; loads_bar_n (+)= 0
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
adjoint_grads_kernel(int E, int N, int Nf, int Ns, const int32_t* __restrict__ edge_nodes,
                     const int32_t* __restrict__ node_slot, const int32_t* __restrict__ ninc_ptr,
                     const int32_t* __restrict__ ninc_edge, const int32_t* __restrict__ ninc_other,
                     const double* __restrict__ q, const double* __restrict__ xyz,
                     const double* __restrict__ lam, double* __restrict__ q_bar,
                     double* __restrict__ loads_bar, double* __restrict__ xs_bar,
                     int64_t q_stride, int accumulate) {
  const int64_t b = blockIdx.y;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  q += b * q_stride;
  xyz += b * (int64_t)N * 3;
  lam += b * (int64_t)Nf * 3;
  const double3 zero = make_double3(0.0, 0.0, 0.0);
  if (t < E && q_bar) {
    const int2 th = __ldg(reinterpret_cast<const int2*>(edge_nodes) + t);
    const int st = __ldg(node_slot + th.x), sh = __ldg(node_slot + th.y);
    const double3 lt = st >= 0 ? ld3(lam, st) : zero, lh = sh >= 0 ? ld3(lam, sh) : zero;
    const double3 a = ld3(xyz, th.x), c = ld3(xyz, th.y);
    const double g = -((lh.x - lt.x) * (c.x - a.x) + (lh.y - lt.y) * (c.y - a.y) + (lh.z - lt.z) * (c.z - a.z));
    double* out = q_bar + b * (int64_t)E + t;
    *out = accumulate ? *out + g : g;
  }
  if (t < N) {
    const int slot = __ldg(node_slot + t);
    if (slot >= 0) {
      if (loads_bar) {
        double* out = loads_bar + (b * (int64_t)N + t) * 3;
        const double3 l = ld3(lam, slot);
        if (accumulate) { out[0] += l.x; out[1] += l.y; out[2] += l.z; }
        else { out[0] = l.x; out[1] = l.y; out[2] = l.z; }
      }
    } else {
      if (loads_bar && !accumulate) st3(loads_bar + b * (int64_t)N * 3, t, zero);
      if (xs_bar) {
        double3 acc = zero;
        const int k0 = __ldg(ninc_ptr + t), k1 = __ldg(ninc_ptr + t + 1);
        for (int k = k0; k < k1; ++k) {
          const int so = __ldg(node_slot + __ldg(ninc_other + k));
          if (so >= 0) {
            const double qe = __ldg(q + (__ldg(ninc_edge + k) >> 1));
            const double3 l = ld3(lam, so);
            acc.x += qe * l.x; acc.y += qe * l.y; acc.z += qe * l.z;
          }
        }
        double* out = xs_bar + (b * (int64_t)Ns + (-slot - 1)) * 3;
        if (accumulate) { out[0] += acc.x; out[1] += acc.y; out[2] += acc.z; }
        else { out[0] = acc.x; out[1] = acc.y; out[2] = acc.z; }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// The cotangent structure of the reference's custom_vjp, materialised (sparse.py:299-306): for every stored
// entry k of the pattern, Kbar.data[k] = sign * sum_xyz a[row_k] * b[col_k].  Thread per COLUMN c (the pattern is
// CSC = CSR, symmetric): b[c] in registers, its entries are contiguous.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
kbar_pattern_kernel(int Nf, int64_t nnz, const int32_t* __restrict__ colptr, const int32_t* __restrict__ rowidx,
                    const double* __restrict__ a, const double* __restrict__ bvec, double* __restrict__ out, double sign) {
  const int64_t b = blockIdx.y;
  const int c = blockIdx.x * kThreads + threadIdx.x;
  if (c >= Nf) return;
  a += b * (int64_t)Nf * 3;
  bvec += b * (int64_t)Nf * 3;
  out += b * nnz;
  const double3 bc = ld3(bvec, c);
  const int k0 = __ldg(colptr + c), k1 = __ldg(colptr + c + 1);
  for (int k = k0; k < k1; ++k) {
    const double3 ar = ld3(a, __ldg(rowidx + k));
    out[k] = sign * (ar.x * bc.x + ar.y * bc.y + ar.z * bc.z);
  }
}

// Transpose of the assembly of K.data (models.py:1069-1077): q_bar_e = [tail free] Kbar[diag_tail] +
// [head free] Kbar[diag_head] - [both free] (Kbar[(tail, head)] + Kbar[(head, tail)]).  Thread per edge; the two
// off-diagonal entries are found by scanning the (short) columns of its end nodes.
__global__ void __launch_bounds__(kThreads)
stiffness_vjp_kernel(int E, int64_t nnz, const int32_t* __restrict__ edge_nodes, const int32_t* __restrict__ node_slot,
                     const int32_t* __restrict__ colptr, const int32_t* __restrict__ rowidx,
                     const int32_t* __restrict__ diag_pos, const double* __restrict__ Kbar, double* __restrict__ q_bar) {
  const int64_t b = blockIdx.y;
  const int e = blockIdx.x * kThreads + threadIdx.x;
  if (e >= E) return;
  Kbar += b * nnz;
  const int2 th = __ldg(reinterpret_cast<const int2*>(edge_nodes) + e);
  const int i = __ldg(node_slot + th.x), j = __ldg(node_slot + th.y);
  double g = 0.0;
  if (i >= 0) g += Kbar[__ldg(diag_pos + i)];
  if (j >= 0) g += Kbar[__ldg(diag_pos + j)];
  if (i >= 0 && j >= 0) {
    for (int k = __ldg(colptr + i); k < __ldg(colptr + i + 1); ++k)
      if (__ldg(rowidx + k) == j) { g -= Kbar[k]; break; }
    for (int k = __ldg(colptr + j); k < __ldg(colptr + j + 1); ++k)
      if (__ldg(rowidx + k) == i) { g -= Kbar[k]; break; }
  }
  q_bar[b * (int64_t)E + e] = g;
}

// ---------------------------------------------------------------------------------------
// small vector helpers for the load kernels
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double3 v_sub(double3 a, double3 b) { return make_double3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ double3 v_add(double3 a, double3 b) { return make_double3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ double3 v_scale(double s, double3 a) { return make_double3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ double v_dot(double3 a, double3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double3 v_cross(double3 a, double3 b) {
  return make_double3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// cotangent of a through y = a / |a| :  (ybar - y (y . ybar)) / |a|
__device__ __forceinline__ double3 unit_bar(double3 y, double norm, double3 ybar) {
  return v_scale(1.0 / norm, v_sub(ybar, v_scale(v_dot(y, ybar), y)));
}
// jnp.allclose(|e . w|, 1.0) with the default tolerances (geometry.py:609, 648)
__device__ __forceinline__ bool near_one(double c) { return fabs(fabs(c) - 1.0) <= 1e-8 + 1e-5; }

// Local coordinate system of a line (geometry.py:581-618): U = unit(x_h - x_t), V = unit(vperp x U) with
// vperp = Z (Y when U is parallel to Z), W = unit(+-(U x V)) signed so that W . vperp >= 0.
struct LineFrame {
  double3 u, v, w, vperp;
  double len, nv, nw, s;
};
__device__ __forceinline__ LineFrame line_frame(double3 xt, double3 xh) {
  LineFrame f;
  const double3 d = v_sub(xh, xt);
  f.len = sqrt(v_dot(d, d));
  f.u = v_scale(1.0 / f.len, d);
  f.vperp = near_one(f.u.z) ? make_double3(0.0, 1.0, 0.0) : make_double3(0.0, 0.0, 1.0);
  const double3 v0 = v_cross(f.vperp, f.u);
  f.nv = sqrt(v_dot(v0, v0));
  f.v = v_scale(1.0 / f.nv, v0);
  const double3 w0 = v_cross(f.u, f.v);
  f.s = v_dot(w0, f.vperp) < 0.0 ? -1.0 : 1.0;
  f.nw = sqrt(v_dot(w0, w0));
  f.w = v_scale(f.s / f.nw, w0);
  return f;
}
// total load of an edge: |d| * (w_u U + w_v V + w_w W) in the local frame, |d| * w in the global one
__device__ __forceinline__ double3 edge_total(const LineFrame& f, double3 wl, bool local) {
  const double3 g = local ? v_add(v_add(v_scale(wl.x, f.u), v_scale(wl.y, f.v)), v_scale(wl.z, f.w)) : wl;
  return v_scale(f.len, g);
}
// cotangent of d = x_h - x_t given the cotangent `ob` of the edge total
__device__ __forceinline__ double3 edge_total_dbar(const LineFrame& f, double3 wl, double3 ob, bool local) {
  if (!local) return v_scale(v_dot(wl, ob), f.u);                      // only the length depends on d
  const double3 g = v_add(v_add(v_scale(wl.x, f.u), v_scale(wl.y, f.v)), v_scale(wl.z, f.w));
  const double lbar = v_dot(g, ob);
  const double3 gb = v_scale(f.len, ob);
  double3 ub = v_scale(wl.x, gb), vb = v_scale(wl.y, gb);
  const double3 wb = v_scale(wl.z, gb);
  const double3 w0b = v_scale(f.s, unit_bar(f.w, f.nw, wb));           // w = unit(s w0), w0 = u x v
  ub = v_add(ub, v_cross(f.v, w0b));
  vb = v_add(vb, v_cross(w0b, f.u));
  const double3 v0b = unit_bar(f.v, f.nv, vb);                          // v = unit(v0), v0 = vperp x u
  ub = v_add(ub, v_cross(v0b, f.vperp));
  return v_add(unit_bar(f.u, f.len, ub), v_scale(lbar, f.u));           // u = unit(d), len = |d|
}

// ---------------------------------------------------------------------------------------
// shape-dependent loads: tributary edge line loads (loads.py:296-331, 340-465), global frame or the
// edge's local frame (follower loads)
//   forward, thread per node:  out_n = loads_n + 1/2 sum_inc total_e                (ascending edge id)
//   reverse, thread per edge:  w_bar_e = |d| * (frame^T) 1/2 (c_t + c_h)
//            thread per node:  xyz_bar_n = sum_inc (+-) dbar_e
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
edge_loads_kernel(int N, int E, const int32_t* __restrict__ ninc_ptr, const int32_t* __restrict__ ninc_edge,
                  const int32_t* __restrict__ ninc_other, const double* __restrict__ xyz,
                  const double* __restrict__ loads_nodes, const double* __restrict__ edges_load,
                  double* __restrict__ out, int64_t ls, int64_t es, int local) {
  const int64_t b = blockIdx.y;
  const int n = blockIdx.x * kThreads + threadIdx.x;
  if (n >= N) return;
  xyz += b * (int64_t)N * 3;
  edges_load += b * es;
  double3 acc = make_double3(0.0, 0.0, 0.0);
  const double3 xn = ld3(xyz, n);
  const int k0 = __ldg(ninc_ptr + n), k1 = __ldg(ninc_ptr + n + 1);
  for (int k = k0; k < k1; ++k) {
    const int code = __ldg(ninc_edge + k), e = code >> 1;
    const double3 xo = ld3(xyz, __ldg(ninc_other + k));
    const LineFrame f = line_frame((code & 1) ? xo : xn, (code & 1) ? xn : xo);   // code & 1: n is the head
    acc = v_add(acc, edge_total(f, ld3(edges_load, e), local != 0));
  }
  const double3 p = ld3(loads_nodes + b * ls, n);
  st3(out + b * (int64_t)N * 3, n, make_double3(p.x + 0.5 * acc.x, p.y + 0.5 * acc.y, p.z + 0.5 * acc.z));
}

__global__ void __launch_bounds__(kThreads)
edge_loads_vjp_kernel(int N, int E, const int32_t* __restrict__ edge_nodes, const int32_t* __restrict__ ninc_ptr,
                      const int32_t* __restrict__ ninc_edge, const int32_t* __restrict__ ninc_other,
                      const double* __restrict__ xyz, const double* __restrict__ edges_load,
                      const double* __restrict__ cot, double* __restrict__ xyz_bar,
                      double* __restrict__ edges_load_bar, int64_t es, int local) {
  const int64_t b = blockIdx.y;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  xyz += b * (int64_t)N * 3;
  cot += b * (int64_t)N * 3;
  edges_load += b * es;
  if (t < E && edges_load_bar) {
    const int2 th = __ldg(reinterpret_cast<const int2*>(edge_nodes) + t);
    const LineFrame f = line_frame(ld3(xyz, th.x), ld3(xyz, th.y));
    const double3 gb = v_scale(0.5 * f.len, v_add(ld3(cot, th.x), ld3(cot, th.y)));
    st3(edges_load_bar + b * (int64_t)E * 3, t,
        local ? make_double3(v_dot(f.u, gb), v_dot(f.v, gb), v_dot(f.w, gb)) : gb);
  }
  if (t < N && xyz_bar) {
    double3 acc = make_double3(0.0, 0.0, 0.0);
    const double3 xn = ld3(xyz, t), cn = ld3(cot, t);
    const int k0 = __ldg(ninc_ptr + t), k1 = __ldg(ninc_ptr + t + 1);
    for (int k = k0; k < k1; ++k) {
      const int code = __ldg(ninc_edge + k), e = code >> 1, o = __ldg(ninc_other + k);
      const bool is_head = code & 1;
      const double3 xo = ld3(xyz, o);
      const LineFrame f = line_frame(is_head ? xo : xn, is_head ? xn : xo);
      const double3 ob = v_scale(0.5, v_add(cn, ld3(cot, o)));
      const double3 db = edge_total_dbar(f, ld3(edges_load, e), ob, local != 0);
      acc = is_head ? v_add(acc, db) : v_sub(acc, db);     // d = x_head - x_tail
    }
    st3(xyz_bar + b * (int64_t)N * 3, t, acc);
  }
}

// ---------------------------------------------------------------------------------------
// tributary face loads in the global frame (loads.py:35-71, 187-288)
//   centroid_f = mean of the face's vertices                                    (thread per face)
//   load_e     = sum_{f in first two faces of e} A(x_t, x_h, centroid_f) w_f,  A = |(x_h-x_t) x (c-x_t)| / 2
//   out_n      = loads_n + 1/2 sum_{e incident to n} load_e                     (thread per node)
// reverse:  per face: w_bar_f, centroid_bar_f over the face's edges; per node: direct area terms of its
// incident edges + centroid_bar_f / deg_f of the faces around it.  No atomics, fixed summation orders.
// ---------------------------------------------------------------------------------------
struct TriGrad { double area; double3 gh, gc; };   // dA/dx_h, dA/dc  (dA/dx_t = -(gh + gc))

__device__ __forceinline__ TriGrad tri_area(double3 a, double3 b, double3 c, bool want_grad) {
  const double3 u = make_double3(b.x - a.x, b.y - a.y, b.z - a.z), v = make_double3(c.x - a.x, c.y - a.y, c.z - a.z);
  const double3 n = make_double3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
  const double len = sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
  TriGrad g;
  g.area = 0.5 * len;
  g.gh = g.gc = make_double3(0.0, 0.0, 0.0);
  if (want_grad) {
    const double s = 0.5 / len;   // zero-area triangle -> NaN, like the norm's gradient in the reference
    const double3 h = make_double3(n.x * s, n.y * s, n.z * s);
    g.gh = make_double3(v.y * h.z - v.z * h.y, v.z * h.x - v.x * h.z, v.x * h.y - v.y * h.x);   // v x n^ / 2
    g.gc = make_double3(h.y * u.z - h.z * u.y, h.z * u.x - h.x * u.z, h.x * u.y - h.y * u.x);   // n^ x u / 2
  }
  return g;
}

// Local coordinate system of a polygon (geometry.py:621-657, normal_polygon :380-425): W = unit normal
// of the centred polygon, V = unit(W x vperp) with vperp = X (Y when W is parallel to X), U = unit(+-(W x V))
// signed so that U . vperp >= 0; a (near) zero normal falls back to the identity frame (loads.py:161-185).
constexpr int kMaxFaceDeg = 8;
struct PolyFrame {
  double3 u, v, w, vperp, c;
  double nn, nv, nu, s;
  bool identity;
};
__device__ __forceinline__ PolyFrame poly_frame(const double3* p, int deg, double3* r) {
  PolyFrame f;
  f.c = make_double3(0.0, 0.0, 0.0);
  for (int i = 0; i < deg; ++i) f.c = v_add(f.c, p[i]);
  f.c = v_scale(1.0 / deg, f.c);
  for (int i = 0; i < deg; ++i) r[i] = v_sub(p[i], f.c);
  double3 n0 = make_double3(0.0, 0.0, 0.0);
  for (int i = 0; i < deg; ++i) n0 = v_add(n0, v_scale(0.5, v_cross(r[i == 0 ? deg - 1 : i - 1], r[i])));
  f.identity = fabs(n0.x) <= 1e-8 && fabs(n0.y) <= 1e-8 && fabs(n0.z) <= 1e-8;
  if (f.identity) {
    f.u = make_double3(1.0, 0.0, 0.0); f.v = make_double3(0.0, 1.0, 0.0); f.w = make_double3(0.0, 0.0, 1.0);
    f.vperp = f.u; f.nn = f.nv = f.nu = f.s = 1.0;
    return f;
  }
  f.nn = sqrt(v_dot(n0, n0));
  f.w = v_scale(1.0 / f.nn, n0);
  f.vperp = near_one(f.w.x) ? make_double3(0.0, 1.0, 0.0) : make_double3(1.0, 0.0, 0.0);
  const double3 v0 = v_cross(f.w, f.vperp);
  f.nv = sqrt(v_dot(v0, v0));
  f.v = v_scale(1.0 / f.nv, v0);
  const double3 u0 = v_cross(f.w, f.v);
  f.s = v_dot(u0, f.vperp) < 0.0 ? -1.0 : 1.0;
  f.nu = sqrt(v_dot(u0, u0));
  f.u = v_scale(f.s / f.nu, u0);
  return f;
}

// per face: centroid and the face load in global coordinates (gl = fl, or fl rotated by the face frame)
__global__ void __launch_bounds__(kThreads)
face_centroids_kernel(int F, int D, int N, const int32_t* __restrict__ faces, const double* __restrict__ xyz,
                      const double* __restrict__ faces_load, double* __restrict__ cen, double* __restrict__ gl,
                      int64_t fs, int local) {
  const int64_t b = blockIdx.y;
  const int f = blockIdx.x * kThreads + threadIdx.x;
  if (f >= F) return;
  xyz += b * (int64_t)N * 3;
  double3 p[kMaxFaceDeg], r[kMaxFaceDeg];
  int deg = 0;
  for (int i = 0; i < D; ++i) {
    const int v = __ldg(faces + (int64_t)f * D + i);
    if (v >= 0 && deg < kMaxFaceDeg) p[deg++] = ld3(xyz, v);
  }
  const PolyFrame fr = poly_frame(p, deg, r);
  st3(cen + b * (int64_t)F * 3, f, fr.c);
  const double3 fl = ld3(faces_load + b * fs, f);
  const double3 g = local ? v_add(v_add(v_scale(fl.x, fr.u), v_scale(fl.y, fr.v)), v_scale(fl.z, fr.w)) : fl;
  st3(gl + b * (int64_t)F * 3, f, g);
}

__global__ void __launch_bounds__(kThreads)
face_loads_kernel(int N, int F, const int32_t* __restrict__ edges_faces, const int32_t* __restrict__ ninc_ptr,
                  const int32_t* __restrict__ ninc_edge, const int32_t* __restrict__ ninc_other,
                  const double* __restrict__ xyz, const double* __restrict__ loads_in,
                  const double* __restrict__ gl, const double* __restrict__ cen, double* __restrict__ out,
                  int64_t ls) {
  const int64_t b = blockIdx.y;
  const int n = blockIdx.x * kThreads + threadIdx.x;
  if (n >= N) return;
  xyz += b * (int64_t)N * 3;
  cen += b * (int64_t)F * 3;
  gl += b * (int64_t)F * 3;
  double3 acc = make_double3(0.0, 0.0, 0.0);
  const double3 xn = ld3(xyz, n);
  const int k0 = __ldg(ninc_ptr + n), k1 = __ldg(ninc_ptr + n + 1);
  for (int k = k0; k < k1; ++k) {
    const int code = __ldg(ninc_edge + k), e = code >> 1;
    const double3 xo = ld3(xyz, __ldg(ninc_other + k));
    const double3 xt = (code & 1) ? xo : xn, xh = (code & 1) ? xn : xo;   // code & 1: n is the head
    for (int slot = 0; slot < 2; ++slot) {
      const int f = __ldg(edges_faces + 2 * (int64_t)e + slot);
      if (f >= 0) {
        const double a = tri_area(xt, xh, ld3(cen, f), false).area;
        const double3 w = ld3(gl, f);
        acc.x += a * w.x; acc.y += a * w.y; acc.z += a * w.z;
      }
    }
  }
  const double3 p = ld3(loads_in + b * ls, n);
  st3(out + b * (int64_t)N * 3, n, make_double3(p.x + 0.5 * acc.x, p.y + 0.5 * acc.y, p.z + 0.5 * acc.z));
}

// per face: centroid_bar, faces_load_bar and (local frames) the cotangents of the face's own vertices
__global__ void __launch_bounds__(kThreads)
face_loads_vjp_faces_kernel(int F, int D, int N, const int32_t* __restrict__ faces,
                            const int32_t* __restrict__ faces_edges, const int32_t* __restrict__ edge_nodes,
                            const double* __restrict__ xyz, const double* __restrict__ faces_load,
                            const double* __restrict__ gl, const double* __restrict__ cen,
                            const double* __restrict__ cot, double* __restrict__ cen_bar,
                            double* __restrict__ fl_bar, double* __restrict__ vert_bar, int64_t fs, int local) {
  const int64_t b = blockIdx.y;
  const int f = blockIdx.x * kThreads + threadIdx.x;
  if (f >= F) return;
  xyz += b * (int64_t)N * 3;
  cot += b * (int64_t)N * 3;
  const double3 c = ld3(cen + b * (int64_t)F * 3, f);
  const double3 w = ld3(gl + b * (int64_t)F * 3, f);
  double3 cb = make_double3(0.0, 0.0, 0.0), glb = make_double3(0.0, 0.0, 0.0);
  for (int i = 0; i < D; ++i) {
    const int e = __ldg(faces_edges + (int64_t)f * D + i);
    if (e < 0) continue;
    const int2 th = __ldg(reinterpret_cast<const int2*>(edge_nodes) + e);
    const double3 hc = v_scale(0.5, v_add(ld3(cot, th.x), ld3(cot, th.y)));
    const TriGrad g = tri_area(ld3(xyz, th.x), ld3(xyz, th.y), c, true);
    glb = v_add(glb, v_scale(g.area, hc));
    cb = v_add(cb, v_scale(v_dot(w, hc), g.gc));                        // cotangent of the triangle area times dA/dc
  }
  if (cen_bar) st3(cen_bar + b * (int64_t)F * 3, f, cb);
  if (!local) {
    if (fl_bar) st3(fl_bar + b * (int64_t)F * 3, f, glb);
    return;
  }
  // local frame: gl = fl.x U + fl.y V + fl.z W  ->  fl_bar = frame . gl_bar, and gl_bar flows into the vertices
  double3 p[kMaxFaceDeg], r[kMaxFaceDeg], rb[kMaxFaceDeg];
  int deg = 0;
  for (int i = 0; i < D; ++i) {
    const int v = __ldg(faces + (int64_t)f * D + i);
    if (v >= 0 && deg < kMaxFaceDeg) p[deg++] = ld3(xyz, v);
  }
  const PolyFrame fr = poly_frame(p, deg, r);
  const double3 fl = ld3(faces_load + b * fs, f);
  if (fl_bar) st3(fl_bar + b * (int64_t)F * 3, f, make_double3(v_dot(fr.u, glb), v_dot(fr.v, glb), v_dot(fr.w, glb)));
  double* vb_out = vert_bar + (b * (int64_t)F + f) * D * 3;
  for (int i = 0; i < deg; ++i) rb[i] = make_double3(0.0, 0.0, 0.0);
  if (!fr.identity) {
    double3 ub = v_scale(fl.x, glb), vb = v_scale(fl.y, glb), wb = v_scale(fl.z, glb);
    const double3 u0b = v_scale(fr.s, unit_bar(fr.u, fr.nu, ub));       // u = unit(s u0), u0 = w x v
    wb = v_add(wb, v_cross(fr.v, u0b));
    vb = v_add(vb, v_cross(u0b, fr.w));
    const double3 v0b = unit_bar(fr.v, fr.nv, vb);                      // v = unit(v0), v0 = w x vperp
    wb = v_add(wb, v_cross(fr.vperp, v0b));
    const double3 n0b = unit_bar(fr.w, fr.nn, wb);                      // w = unit(n0), n0 = 1/2 sum r_{i-1} x r_i
    for (int i = 0; i < deg; ++i) {
      const int im = i == 0 ? deg - 1 : i - 1;
      rb[im] = v_add(rb[im], v_scale(0.5, v_cross(r[i], n0b)));
      rb[i] = v_add(rb[i], v_scale(0.5, v_cross(n0b, r[im])));
    }
  }
  double3 mean = make_double3(0.0, 0.0, 0.0);
  for (int i = 0; i < deg; ++i) mean = v_add(mean, rb[i]);
  mean = v_scale(1.0 / deg, mean);
  int k = 0;
  for (int i = 0; i < D; ++i) {                                         // r_i = p_i - mean(p)
    const bool valid = __ldg(faces + (int64_t)f * D + i) >= 0 && k < deg;
    const double3 o = valid ? v_sub(rb[k], mean) : make_double3(0.0, 0.0, 0.0);
    vb_out[3 * i] = o.x; vb_out[3 * i + 1] = o.y; vb_out[3 * i + 2] = o.z;
    k += valid;
  }
}

__global__ void __launch_bounds__(kThreads)
face_loads_vjp_nodes_kernel(int N, int F, int D, const int32_t* __restrict__ faces,
                            const int32_t* __restrict__ edges_faces, const int32_t* __restrict__ node_faces_ptr,
                            const int32_t* __restrict__ node_faces, const int32_t* __restrict__ ninc_ptr,
                            const int32_t* __restrict__ ninc_edge, const int32_t* __restrict__ ninc_other,
                            const double* __restrict__ xyz, const double* __restrict__ gl,
                            const double* __restrict__ cen, const double* __restrict__ cot,
                            const double* __restrict__ cen_bar, const double* __restrict__ vert_bar,
                            double* __restrict__ xyz_bar, int local) {
  const int64_t b = blockIdx.y;
  const int n = blockIdx.x * kThreads + threadIdx.x;
  if (n >= N) return;
  xyz += b * (int64_t)N * 3;
  cot += b * (int64_t)N * 3;
  cen += b * (int64_t)F * 3;
  cen_bar += b * (int64_t)F * 3;
  gl += b * (int64_t)F * 3;
  double3 acc = make_double3(0.0, 0.0, 0.0);
  const double3 xn = ld3(xyz, n), cn = ld3(cot, n);
  const int k0 = __ldg(ninc_ptr + n), k1 = __ldg(ninc_ptr + n + 1);
  for (int k = k0; k < k1; ++k) {
    const int code = __ldg(ninc_edge + k), e = code >> 1, o = __ldg(ninc_other + k);
    const bool is_head = code & 1;
    const double3 xo = ld3(xyz, o);
    const double3 hc = v_scale(0.5, v_add(cn, ld3(cot, o)));
    for (int slot = 0; slot < 2; ++slot) {
      const int f = __ldg(edges_faces + 2 * (int64_t)e + slot);
      if (f < 0) continue;
      const TriGrad g = tri_area(is_head ? xo : xn, is_head ? xn : xo, ld3(cen, f), true);
      const double abar = v_dot(ld3(gl, f), hc);
      // dA/dx_head = gh ; dA/dx_tail = -(gh + gc)
      const double3 gn = is_head ? g.gh : v_scale(-1.0, v_add(g.gh, g.gc));
      acc = v_add(acc, v_scale(abar, gn));
    }
  }
  const int f0 = __ldg(node_faces_ptr + n), f1 = __ldg(node_faces_ptr + n + 1);
  for (int k = f0; k < f1; ++k) {
    const int f = __ldg(node_faces + k);
    int cnt = 0, pos = -1;
    for (int i = 0; i < D; ++i) {
      const int v = __ldg(faces + (int64_t)f * D + i);
      cnt += v >= 0;
      if (v == n && pos < 0) pos = i;
    }
    acc = v_add(acc, v_scale(1.0 / cnt, ld3(cen_bar, f)));
    if (local && pos >= 0) acc = v_add(acc, ld3(vert_bar + (b * (int64_t)F + f) * D * 3, pos));
  }
  st3(xyz_bar + b * (int64_t)N * 3, n, acc);
}

// ---------------------------------------------------------------------------------------
// loss helper: g = x - x0, loss_b = 0.5 sum g^2   (one CTA per design, fixed-order reduction)
// ---------------------------------------------------------------------------------------
// out = a x + b y + c z (y, z optional), grid-stride: the few vector combinations the host-side rules need between
// kernels (tangent right-hand sides, fixed-point adjoint updates) without leaving this library
__global__ void __launch_bounds__(256) lincomb_kernel(double* __restrict__ out, const double* __restrict__ x,
                                                       const double* __restrict__ y, const double* __restrict__ z, double a,
                                                       double b, double c, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    double v = a * x[i];
    if (y) v += b * y[i];
    if (z) v += c * z[i];
    out[i] = v;
  }
}

// out[i] = sum_b in[b, i] in a fixed order (deterministic, launch-geometry independent): the local part of the
// loss / shared-gradient all-reduce (distributed.py).  n == 1: one CTA, thread t adds b = t, t + T, ... then a
// fixed tree; otherwise one thread per column, b ascending (coalesced across the warp).
__global__ void __launch_bounds__(1024) batch_sum_scalar_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t B) {
  double acc = 0.0;
  for (int64_t b = threadIdx.x; b < B; b += 1024) acc += in[b];
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = red[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) out[0] = acc;
  }
}
__global__ void __launch_bounds__(256) batch_sum_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t B, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int64_t b = 0;
  for (; b + 3 < B; b += 4) {
    a0 += in[b * n + i]; a1 += in[(b + 1) * n + i]; a2 += in[(b + 2) * n + i]; a3 += in[(b + 3) * n + i];
  }
  for (; b < B; ++b) a0 += in[b * n + i];
  out[i] = (a0 + a1) + (a2 + a3);
}

// fdm_solve_opts.nan_on_failure for the solvers whose own kernels do not do it: design b (blockIdx.y) with a
// status other than OK gets x = NaN
__global__ void __launch_bounds__(256) nan_failed_kernel(const double* __restrict__ info, double* __restrict__ x, int64_t n) {
  const int64_t b = blockIdx.y;
  if (__ldg(info + b * FDM_INFO_STRIDE + FDM_INFO_STATUS) == 0.0) return;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) x[b * n + i] = __longlong_as_double(0x7ff8000000000000LL);
}

__global__ void __launch_bounds__(1024)
loss_sqdist_kernel(const double* __restrict__ x, const double* __restrict__ x0,
                   double* __restrict__ g, double* __restrict__ loss, int64_t n, int64_t x0_stride) {
  const int64_t b = blockIdx.x;
  x += b * n;
  x0 += b * x0_stride;
  if (g) g += b * n;
  // fixed order: thread t accumulates elements t, t + T, t + 2T, ... in that order (four loads in flight per trip)
  double acc = 0.0;
  const int T = blockDim.x;
  int64_t i = threadIdx.x;
  for (; i + 3 * T < n; i += 4 * T) {
    const double d0 = x[i] - __ldg(x0 + i), d1 = x[i + T] - __ldg(x0 + i + T);
    const double d2 = x[i + 2 * T] - __ldg(x0 + i + 2 * T), d3 = x[i + 3 * T] - __ldg(x0 + i + 3 * T);
    if (g) { g[i] = d0; g[i + T] = d1; g[i + 2 * T] = d2; g[i + 3 * T] = d3; }
    acc += d0 * d0; acc += d1 * d1; acc += d2 * d2; acc += d3 * d3;
  }
  for (; i < n; i += T) {
    const double d = x[i] - __ldg(x0 + i);
    if (g) g[i] = d;
    acc += d * d;
  }
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0 && loss) loss[b] = 0.5 * acc;
  }
}

// The same loss for a batch, persistent CTAs: CTA c takes designs c, c + gridDim.x, ...; a thread keeps its NPT
// elements of x0 in registers when the designs share the target (x0_stride == 0) and loads design b + step's x before
// it reduces design b, so the reduction's barrier never leaves the memory pipe idle.  Same per-thread order of
// accumulation (elements t, t + 256, ...) and the same reduction tree as loss_sqdist_kernel => identical bits.
template <int NPT>
__global__ void __launch_bounds__(256)
loss_sqdist_batch_kernel(const double* __restrict__ x, const double* __restrict__ x0, double* __restrict__ g,
                         double* __restrict__ loss, int n, int64_t batch, int64_t x0_stride) {
  __shared__ double red[2][8];
  const int tid = threadIdx.x;
  double t0[NPT], cur[NPT], nxt[NPT];
#pragma unroll
  for (int k = 0; k < NPT; ++k) {
    const int i = tid + k * 256;
    t0[k] = x0_stride == 0 && i < n ? __ldg(x0 + i) : 0.0;
    cur[k] = i < n && (int64_t)blockIdx.x < batch ? x[(int64_t)blockIdx.x * n + i] : 0.0;
    nxt[k] = 0.0;
  }
  int buf = 0;
  for (int64_t b = blockIdx.x; b < batch; b += gridDim.x, buf ^= 1) {
    const int64_t bn = b + gridDim.x;
    if (bn < batch) {
#pragma unroll
      for (int k = 0; k < NPT; ++k)
        if (tid + k * 256 < n) nxt[k] = x[bn * n + tid + k * 256];
    }
    if (x0_stride != 0) {
#pragma unroll
      for (int k = 0; k < NPT; ++k)
        if (tid + k * 256 < n) t0[k] = __ldg(x0 + b * x0_stride + tid + k * 256);
    }
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < NPT; ++k)
      if (tid + k * 256 < n) {
        const double d = cur[k] - t0[k];
        if (g) g[b * n + tid + k * 256] = d;
        acc += d * d;
      }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) red[buf][tid >> 5] = acc;
    __syncthreads();            // red[buf] complete; red[buf ^ 1] is rewritten only after the next barrier
    if (tid < 32) {
      acc = tid < 8 ? red[buf][tid] : 0.0;
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (tid == 0 && loss) loss[b] = 0.5 * acc;
    }
#pragma unroll
    for (int k = 0; k < NPT; ++k) cur[k] = nxt[k];
  }
}

inline dim3 grid_for(int64_t work, int64_t batch) {
  return dim3((unsigned)((work + kThreads - 1) / kThreads), (unsigned)batch, 1);
}

}  // namespace

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
int launch_assemble(const fdm_plan* p, const double* q, const double* xs, const double* loads,
                    double* Kdata, double* rhs, int64_t batch, int64_t qs, int64_t xss, int64_t ls,
                    cudaStream_t st) {
  const DeviceArrays& d = p->d;
  const unsigned tiles = (unsigned)((p->Nf + kAsmRows - 1) / kAsmRows);
  // a batch on one topology: whole designs per CTA, index streams read once per launch (when the plan says it fits)
  if (batch >= 8 && p->asm_batch_ok && !(std::getenv("FDM_ASM_BATCH") && std::atoi(std::getenv("FDM_ASM_BATCH")) == 0)) {
    const int sms = std::max(p->sm_count, 1);
    const int nsup = (int)p->rsup_rows.size();
    const bool sup_in_K = Kdata && rhs && nsup && p->asm_sup_in_K;
    bool rhs_tma = false;
    if (rhs) {
      const int64_t img_bytes = 16 * 3 * (int64_t)p->Nf;
      if (ls == 0 && (reinterpret_cast<uintptr_t>(rhs) & 15) == 0 && img_bytes <= 96 * 1024) {
        rhs_tma = true;
        FDM_CUDA(cudaFuncSetAttribute(assemble_rhs_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        int dpc = 2 * (int)std::max<int64_t>(1, (batch + 2 * sms - 1) / (2 * sms));       // even; one CTA per SM
        if (const char* e = std::getenv("FDM_RHS_DPC")) dpc = std::max(2, std::atoi(e) & ~1);
        assemble_rhs_tma_kernel<<<(unsigned)((batch + dpc - 1) / dpc), kRhsThreads, (size_t)img_bytes, st>>>(
            (int)p->Nf, batch, dpc, d.rhs_src, loads, rhs);
        if (int rc = fdm_check_launch("assemble_rhs_tma")) return rc;
      } else {
        const int dpc = (int)std::max<int64_t>(1, (batch + 4 * sms - 1) / (4 * sms));
        assemble_rhs_stream_kernel<<<(unsigned)((batch + dpc - 1) / dpc), kRhsThreads, 0, st>>>((int)p->Nf, batch, dpc,
                                                                                                d.rhs_src, loads, rhs, ls);
        if (int rc = fdm_check_launch("assemble_rhs_stream")) return rc;
      }
    }
    if (Kdata) {
      const int smem = (int)p->asm_smem_bytes;
      auto kern = p->nnz <= 8 * kThreads ? (sup_in_K ? assemble_K_batch_kernel<8, true> : assemble_K_batch_kernel<8, false>)
                                         : (sup_in_K ? assemble_K_batch_kernel<16, true> : assemble_K_batch_kernel<16, false>);
      FDM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      const int per_sm = std::max(1, std::min(3, (227 * 1024) / (smem + 1024)));          // resident CTAs per SM
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)std::min<int64_t>(batch, (int64_t)sms * per_sm));
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = (size_t)smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      static const bool pdl_off = std::getenv("FDM_ASM_PDL") && std::atoi(std::getenv("FDM_ASM_PDL")) == 0;
      cfg.attrs = attr;
      cfg.numAttrs = sup_in_K && rhs_tma && !pdl_off ? 1 : 0;      // only behind the bulk-store rhs kernel
      FDM_CUDA(cudaLaunchKernelEx(&cfg, kern, (int)p->E, (int)p->Nf, (int)p->nnz, p->asm_maxdeg, batch,
                                  (const int32_t*)d.asm_code, (const int32_t*)d.diags_ptr, (const int32_t*)d.diags_idx, q,
                                  Kdata, qs, nsup, (const int32_t*)d.rsup_rows, (const int32_t*)d.rsup_ptr,
                                  (const int32_t*)d.rsup_edge, (const int32_t*)d.rsup_slot, (const int32_t*)d.free_node, xs,
                                  loads, rhs, xss, ls));
      if (int rc = fdm_check_launch("assemble_K_batch")) return rc;
    }
    if (rhs && nsup && !sup_in_K) {
      const int dpp = std::max(1, (16 * kRhsThreads) / (3 * nsup));                       // about four trips per CTA
      assemble_rhs_pairs_kernel<<<(unsigned)((batch + dpp - 1) / dpp), kRhsThreads, 0, st>>>(
          (int)p->Nf, nsup, batch, dpp, d.rsup_rows, d.rsup_ptr, d.rsup_edge, d.rsup_slot, d.free_node, q, xs, loads, rhs, qs,
          xss, ls);
      if (int rc = fdm_check_launch("assemble_rhs_pairs")) return rc;
    }
    return FDM_OK;
  }
  FDM_BATCH_LOOP(batch, b0, nb) {
    assemble_kernel<<<dim3(tiles, (unsigned)nb), kThreads, 0, st>>>(
        (int)p->nnz, (int)p->Nf, (int)p->N, (int)p->Ns, d.rowptr, d.entry_edge, d.diag_pos, d.free_node,
        d.diags_ptr, d.diags_idx, d.sup_mask, d.node_slot, d.ninc_ptr, d.ninc_edge, d.ninc_other, q + b0 * qs,
        xs ? xs + b0 * xss : nullptr, loads ? loads + b0 * ls : nullptr,
        Kdata ? Kdata + b0 * p->nnz : nullptr, rhs ? rhs + b0 * p->Nf * 3 : nullptr, qs, xss, ls);
  }
  return fdm_check_launch("assemble");
}

int launch_positions(const fdm_plan* p, double* xf, double* xs, double* xyz, int64_t batch,
                     int64_t xss, int transpose, cudaStream_t st) {
  FDM_BATCH_LOOP(batch, b0, nb) {
    positions_kernel<<<grid_for(p->N, nb), kThreads, 0, st>>>(
        (int)p->N, (int)p->Nf, (int)p->Ns, p->d.node_slot, xf ? xf + b0 * p->Nf * 3 : nullptr,
        xs ? xs + b0 * xss : nullptr, xyz + b0 * p->N * 3, xss, transpose);
  }
  return fdm_check_launch("positions");
}

int launch_state(const fdm_plan* p, const double* q, const double* xyz, const double* loads,
                 double* vectors, double* lengths, double* forces, double* residuals,
                 int64_t batch, int64_t qs, int64_t ls, cudaStream_t st) {
  const DeviceArrays& d = p->d;
  const int64_t work = std::max(p->E, p->N);
  FDM_BATCH_LOOP(batch, b0, nb) {
    state_kernel<<<grid_for(work, nb), kThreads, 0, st>>>(
        (int)p->E, (int)p->N, d.edge_nodes, d.ninc_ptr, d.ninc_edge, d.ninc_other, q + b0 * qs,
        xyz + b0 * p->N * 3, loads ? loads + b0 * ls : nullptr,
        vectors ? vectors + b0 * p->E * 3 : nullptr, lengths ? lengths + b0 * p->E : nullptr,
        forces ? forces + b0 * p->E : nullptr, residuals ? residuals + b0 * p->N * 3 : nullptr, qs, ls);
  }
  return fdm_check_launch("state");
}

int launch_state_jvp(const fdm_plan* p, const double* q, const double* xyz, const double* q_dot, const double* xyz_dot,
                     const double* loads_dot, double* vd, double* ld, double* fd, double* rd, int64_t batch, int64_t qs,
                     int64_t xs, int64_t qds, int64_t lds, cudaStream_t st) {
  const DeviceArrays& d = p->d;
  const int64_t work = std::max(p->E, p->N);
  FDM_BATCH_LOOP(batch, b0, nb) {
    state_jvp_kernel<<<grid_for(work, nb), kThreads, 0, st>>>(
        (int)p->E, (int)p->N, d.edge_nodes, d.ninc_ptr, d.ninc_edge, d.ninc_other, q + b0 * qs, xyz + b0 * xs,
        q_dot ? q_dot + b0 * qds : nullptr, xyz_dot + b0 * p->N * 3, loads_dot ? loads_dot + b0 * lds : nullptr,
        vd ? vd + b0 * p->E * 3 : nullptr, ld ? ld + b0 * p->E : nullptr, fd ? fd + b0 * p->E : nullptr,
        rd ? rd + b0 * p->N * 3 : nullptr, qs, xs, qds, lds);
  }
  return fdm_check_launch("state_jvp");
}

int launch_state_vjp(const fdm_plan* p, const double* q, const double* xyz, const double* xyz_cot,
                     const double* rbar, const double* lbar, const double* fbar,
                     const double* loads_cot, const double* vbar, double* q_bar, double* xyz_bar,
                     double* loads_bar, int64_t batch, int64_t qs, cudaStream_t st) {
  const DeviceArrays& d = p->d;
  const int64_t work = std::max(p->E, p->N);
  const int64_t N3 = p->N * 3, E = p->E;
  FDM_BATCH_LOOP(batch, b0, nb) {
    state_vjp_kernel<<<grid_for(work, nb), kThreads, 0, st>>>(
        (int)p->E, (int)p->N, d.edge_nodes, d.ninc_ptr, d.ninc_edge, d.ninc_other, q + b0 * qs,
        xyz + b0 * N3, xyz_cot ? xyz_cot + b0 * N3 : nullptr, rbar ? rbar + b0 * N3 : nullptr,
        lbar ? lbar + b0 * E : nullptr, fbar ? fbar + b0 * E : nullptr,
        loads_cot ? loads_cot + b0 * N3 : nullptr, vbar ? vbar + b0 * E * 3 : nullptr,
        q_bar ? q_bar + b0 * E : nullptr, xyz_bar ? xyz_bar + b0 * N3 : nullptr,
        loads_bar ? loads_bar + b0 * N3 : nullptr, qs);
  }
  return fdm_check_launch("state_vjp");
}

int launch_adjoint_grads(const fdm_plan* p, const double* q, const double* xyz, const double* lam,
                         double* q_bar, double* loads_bar, double* xs_bar, int64_t batch, int64_t qs,
                         int accumulate, cudaStream_t st) {
  const DeviceArrays& d = p->d;
  const int64_t work = std::max(p->E, p->N);
  FDM_BATCH_LOOP(batch, b0, nb) {
    adjoint_grads_kernel<<<grid_for(work, nb), kThreads, 0, st>>>(
        (int)p->E, (int)p->N, (int)p->Nf, (int)p->Ns, d.edge_nodes, d.node_slot, d.ninc_ptr,
        d.ninc_edge, d.ninc_other, q + b0 * qs, xyz + b0 * p->N * 3, lam + b0 * p->Nf * 3,
        q_bar ? q_bar + b0 * p->E : nullptr, loads_bar ? loads_bar + b0 * p->N * 3 : nullptr,
        xs_bar ? xs_bar + b0 * p->Ns * 3 : nullptr, qs, accumulate);
  }
  return fdm_check_launch("adjoint_grads");
}

int launch_kbar_pattern(const fdm_plan* p, const double* a, const double* bvec, double* out, double sign, int64_t batch,
                        cudaStream_t st) {
  const DeviceArrays& d = p->d;
  FDM_BATCH_LOOP(batch, b0, nb) {
    kbar_pattern_kernel<<<grid_for(p->Nf, nb), kThreads, 0, st>>>((int)p->Nf, p->nnz, d.rowptr, d.colidx, a + b0 * p->Nf * 3,
                                                                 bvec + b0 * p->Nf * 3, out + b0 * p->nnz, sign);
  }
  return fdm_check_launch("kbar_pattern");
}

int launch_stiffness_vjp(const fdm_plan* p, const double* Kbar, double* q_bar, int64_t batch, cudaStream_t st) {
  const DeviceArrays& d = p->d;
  FDM_BATCH_LOOP(batch, b0, nb) {
    stiffness_vjp_kernel<<<grid_for(p->E, nb), kThreads, 0, st>>>((int)p->E, p->nnz, d.edge_nodes, d.node_slot, d.rowptr,
                                                                 d.colidx, d.diag_pos, Kbar + b0 * p->nnz, q_bar + b0 * p->E);
  }
  return fdm_check_launch("stiffness_vjp");
}

int launch_edge_loads(const fdm_plan* p, const double* xyz, const double* loads_nodes, const double* edges_load,
                      double* loads_out, int64_t batch, int64_t ls, int64_t es, int local, cudaStream_t st) {
  const DeviceArrays& d = p->d;
  FDM_BATCH_LOOP(batch, b0, nb) {
    edge_loads_kernel<<<grid_for(p->N, nb), kThreads, 0, st>>>(
        (int)p->N, (int)p->E, d.ninc_ptr, d.ninc_edge, d.ninc_other, xyz + b0 * p->N * 3, loads_nodes + b0 * ls,
        edges_load + b0 * es, loads_out + b0 * p->N * 3, ls, es, local);
  }
  return fdm_check_launch("edge_loads");
}

int launch_edge_loads_vjp(const fdm_plan* p, const double* xyz, const double* edges_load, const double* cot,
                          double* xyz_bar, double* edges_load_bar, int64_t batch, int64_t es, int local,
                          cudaStream_t st) {
  const DeviceArrays& d = p->d;
  const int64_t work = std::max(p->E, p->N);
  FDM_BATCH_LOOP(batch, b0, nb) {
    edge_loads_vjp_kernel<<<grid_for(work, nb), kThreads, 0, st>>>(
        (int)p->N, (int)p->E, d.edge_nodes, d.ninc_ptr, d.ninc_edge, d.ninc_other, xyz + b0 * p->N * 3,
        edges_load + b0 * es, cot + b0 * p->N * 3, xyz_bar ? xyz_bar + b0 * p->N * 3 : nullptr,
        edges_load_bar ? edges_load_bar + b0 * p->E * 3 : nullptr, es, local);
  }
  return fdm_check_launch("edge_loads_vjp");
}

int launch_face_loads(const fdm_plan* p, const int32_t* faces, int64_t F, int64_t D, const int32_t* edges_faces,
                      const double* xyz, const double* loads_in, const double* faces_load, double* cen, double* gl,
                      double* out, int64_t batch, int64_t ls, int64_t fs, int local, cudaStream_t st) {
  const DeviceArrays& d = p->d;
  FDM_BATCH_LOOP(batch, b0, nb) {
    face_centroids_kernel<<<grid_for(F, nb), kThreads, 0, st>>>((int)F, (int)D, (int)p->N, faces, xyz + b0 * p->N * 3,
                                                               faces_load + b0 * fs, cen + b0 * F * 3, gl + b0 * F * 3,
                                                               fs, local);
    face_loads_kernel<<<grid_for(p->N, nb), kThreads, 0, st>>>(
        (int)p->N, (int)F, edges_faces, d.ninc_ptr, d.ninc_edge, d.ninc_other, xyz + b0 * p->N * 3, loads_in + b0 * ls,
        gl + b0 * F * 3, cen + b0 * F * 3, out + b0 * p->N * 3, ls);
  }
  return fdm_check_launch("face_loads");
}

int launch_face_loads_vjp(const fdm_plan* p, const int32_t* faces, int64_t F, int64_t D, const int32_t* edges_faces,
                          const int32_t* faces_edges, const int32_t* nf_ptr, const int32_t* nf, const double* xyz,
                          const double* faces_load, const double* cen, const double* gl, const double* cot,
                          double* cen_bar, double* vert_bar, double* xyz_bar, double* fl_bar, int64_t batch, int64_t fs,
                          int local, cudaStream_t st) {
  const DeviceArrays& d = p->d;
  FDM_BATCH_LOOP(batch, b0, nb) {
    face_loads_vjp_faces_kernel<<<grid_for(F, nb), kThreads, 0, st>>>(
        (int)F, (int)D, (int)p->N, faces, faces_edges, d.edge_nodes, xyz + b0 * p->N * 3, faces_load + b0 * fs,
        gl + b0 * F * 3, cen + b0 * F * 3, cot + b0 * p->N * 3, cen_bar + b0 * F * 3,
        fl_bar ? fl_bar + b0 * F * 3 : nullptr, vert_bar ? vert_bar + b0 * F * D * 3 : nullptr, fs, local);
    if (xyz_bar)
      face_loads_vjp_nodes_kernel<<<grid_for(p->N, nb), kThreads, 0, st>>>(
          (int)p->N, (int)F, (int)D, faces, edges_faces, nf_ptr, nf, d.ninc_ptr, d.ninc_edge, d.ninc_other,
          xyz + b0 * p->N * 3, gl + b0 * F * 3, cen + b0 * F * 3, cot + b0 * p->N * 3, cen_bar + b0 * F * 3,
          vert_bar ? vert_bar + b0 * F * D * 3 : nullptr, xyz_bar + b0 * p->N * 3, local);
  }
  return fdm_check_launch("face_loads_vjp");
}

int launch_lincomb(double* out, const double* x, const double* y, const double* z, double a, double b, double c, int64_t n,
                   cudaStream_t st) {
  const int64_t blocks = std::min<int64_t>((n + 255) / 256, 148 * 16);
  lincomb_kernel<<<(unsigned)blocks, 256, 0, st>>>(out, x, y, z, a, b, c, n);
  return fdm_check_launch("lincomb");
}

int launch_batch_sum(const double* in, double* out, int64_t batch, int64_t n, cudaStream_t st) {
  if (n == 1) batch_sum_scalar_kernel<<<1, 1024, 0, st>>>(in, out, batch);
  else batch_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, batch, n);
  return fdm_check_launch("batch_sum");
}

int launch_nan_failed(const double* info, double* x, int64_t n, int64_t batch, cudaStream_t st) {
  FDM_BATCH_LOOP(batch, b0, nb) {
    nan_failed_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)nb), 256, 0, st>>>(info + b0 * FDM_INFO_STRIDE,
                                                                                       x + b0 * n, n);
  }
  return fdm_check_launch("nan_failed");
}

int launch_loss_sqdist(const double* x, const double* x0, double* g, double* loss, int64_t n,
                       int64_t batch, int64_t x0_stride, cudaStream_t st) {
  if (batch >= 8 && n >= 256 && n <= 12 * 256 && !(std::getenv("FDM_LOSS_BATCH") && std::atoi(std::getenv("FDM_LOSS_BATCH")) == 0)) {
    int dev = 0, sms = 0;
    FDM_CUDA(cudaGetDevice(&dev));
    FDM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int per_sm = 2;                                   // resident CTAs per SM at this kernel's register count
    const unsigned ctas = (unsigned)std::min<int64_t>(batch, (int64_t)sms * per_sm);
    if (n <= 4 * 256) loss_sqdist_batch_kernel<4><<<ctas, 256, 0, st>>>(x, x0, g, loss, (int)n, batch, x0_stride);
    else if (n <= 8 * 256) loss_sqdist_batch_kernel<8><<<ctas, 256, 0, st>>>(x, x0, g, loss, (int)n, batch, x0_stride);
    else if (n <= 10 * 256) loss_sqdist_batch_kernel<10><<<ctas, 256, 0, st>>>(x, x0, g, loss, (int)n, batch, x0_stride);
    else loss_sqdist_batch_kernel<12><<<ctas, 256, 0, st>>>(x, x0, g, loss, (int)n, batch, x0_stride);
    return fdm_check_launch("loss_sqdist_batch");
  }
  const int threads = n >= 256 ? 256 : (int)std::max<int64_t>(32, ((n + 31) / 32) * 32);
  loss_sqdist_kernel<<<(unsigned)batch, threads, 0, st>>>(x, x0, g, loss, n, x0_stride);
  return fdm_check_launch("loss_sqdist");
}
