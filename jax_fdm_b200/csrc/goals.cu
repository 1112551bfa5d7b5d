// Loss = sum over error terms of goals evaluated on the equilibrium state, fused with the cotangents of
// the state.  Replaces, for the goals that read one element of the state,
//   Error.__call__            losses/errors.py:100-135   (vmap of the per-element error, summed per collection)
//   SquaredError / AbsoluteError / PredictionError / LogMaxError .error      errors.py:232-431
//   Mean* / RootMeanSquaredError .errors                                      errors.py:262-300, 340-400
//   Loss.__call__             losses/loss.py:60-90
//   NodePointGoal, NodeResidualForceGoal, NodeResidualVectorGoal, EdgeLengthGoal, EdgeForceGoal,
//   EdgeLoadPathGoal, NetworkLoadPathGoal, Node{X,Y,Z}CoordinateGoal .prediction   (goals/node/point.py,
//   node/residual.py, node/coordinates.py, edge/length.py, edge/force.py, edge/loadpath.py, network/loadpath.py)
// and the reverse pass JAX derives from them.
//
// One CTA per design.  The goal "program" (shared by the batch) is a table of terms sorted by group
// (= one Error instance): kind, element index, error kind, weight, target.  Pass 1 reduces every
// group's sum S_g with a fixed tree (deterministic); loss = sum_g outer_g * f_g(inner_g * S_g), f = id or
// sqrt.  Pass 2 scatters d loss / d state into xyz_bar, lengths_bar, forces_bar, residuals_bar (zeroed
// here first; atomicAdd because several terms may read the same element -- two addends commute exactly,
// three or more on one element are summed in an unspecified order).
#include <cuda_runtime.h>

#include "launch.h"

namespace {

constexpr int kThreads = 256;

enum GoalKind : int {
  G_NODE_POINT = 0,        // xyz[i]                 (3) vs target (3)
  G_NODE_RESIDUAL_FORCE,   // |residuals[i]|         (1)
  G_NODE_RESIDUAL_VECTOR,  // residuals[i]           (3)
  G_EDGE_LENGTH,           // lengths[e]             (1)
  G_EDGE_FORCE,            // forces[e]              (1)
  G_EDGE_LOADPATH,         // |forces[e]| lengths[e] (1)
  G_NETWORK_LOADPATH,      // sum_e |forces[e] lengths[e]| (1)
  G_NODE_X,                // xyz[i, 0], G_NODE_X + 1: xyz[i, 1], + 2: xyz[i, 2]   (1)
  G_NODE_Y,
  G_NODE_Z,
  G_NODE_LINE,             // xyz[i] vs its projection on the line through target[0:3], target[3:6]   (3)
  G_NODE_PLANE,            // xyz[i] vs its projection on the plane (origin target[0:3], normal target[3:6])  (3)
  G_NODE_SEGMENT,          // xyz[i] vs the closest point of the segment target[0:3] - target[3:6]     (3)
  G_NODE_RESIDUAL_DIRECTION,  // residuals[i] / |residuals[i]| vs target / |target|                    (3)
  G_NODE_RESIDUAL_PLANE,      // the same unit vector vs its projection on the plane through 0 with normal target (3)
  G_EDGE_DIRECTION,           // unit edge vector (xyz[head] - xyz[tail]) vs target / |target|          (3)
  G_EDGE_ANGLE,               // angle between the edge vector and target[1:4] vs target[0]             (1)
  G_EDGES_LENGTH_EQUAL,       // var(lengths[S]) / mean(lengths[S]) over the edge set S of the term (index = aggregate slot)
  G_EDGES_FORCE_EQUAL,        // var(forces[S]) / mean(|forces[S]|)
  G_NODES_COLINEAR,           // colinearity energy of the ordered node sequence S (geometry.py:660-694)
};
constexpr int kMaxAgg = 16;    // aggregate goals per loss

// scatter into a cotangent array; a NULL array is one no term of the program reads (fdm_goal_kind_arrays): skipped
__device__ __forceinline__ void gadd(double* base, int idx, double v) {
  if (base) atomicAdd(base + idx, v);
}
enum ErrKind : int { E_SQUARED = 0, E_ABSOLUTE, E_PREDICTION, E_LOGMAX };

struct GoalArgs {
  int T, G, N, E;
  const int32_t *kind, *index, *err, *group_ptr;   // (T), (T), (T), (G+1)
  const int32_t* edge_nodes;                       // plan: (E,2) tail, head
  const int32_t *agg_ptr, *agg_index;              // (T+1), (sum of set sizes): element sets of the aggregate goals
  const double *weight, *target;                   // (T), (T,6)
  const int32_t* gfinal;                           // (G) 0 identity, 1 sqrt
  const double *ginner, *gouter;                   // (G)
  const double *xyz, *lengths, *forces, *residuals;
  double *loss, *gsum;                             // (B), (B,G)
  const double* loss_cot;                          // (B) upstream cotangent of the loss, or null (= 1)
  int num_agg;                                     // aggregate goals in the program (0: their serial term scans are skipped)
  double *xyz_bar, *lengths_bar, *forces_bar, *residuals_bar;
  int need_loadpath;
};

__device__ __forceinline__ double block_sum(double v, double* red) {
  // fixed tree: warp shuffles then one warp over the partials
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = (threadIdx.x < kThreads / 32) ? red[threadIdx.x] : 0.0;
  if (w == 0) {
    for (int o = 4; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (l == 0) red[8] = s;
  }
  __syncthreads();
  return red[8];
}

__device__ __forceinline__ double err_value(int ek, double w, double p, double t) {
  const double d = p - t;
  switch (ek) {
    case E_SQUARED: return w * d * d;
    case E_ABSOLUTE: return w * fabs(d);
    case E_PREDICTION: return w * p;
    default: return w * log1p(fmax(d, 0.0));
  }
}
__device__ __forceinline__ double err_deriv(int ek, double w, double p, double t) {
  const double d = p - t;
  switch (ek) {
    case E_SQUARED: return 2.0 * w * d;
    case E_ABSOLUTE: return d > 0.0 ? w : (d < 0.0 ? -w : 0.0);
    case E_PREDICTION: return w;
    default: return d > 0.0 ? w / (1.0 + d) : 0.0;
  }
}


// Goals whose reference value is a projection of the prediction (goals/node/line.py, plane.py, segment.py with
// geometry.py:260-358).  gp = goal point; the Jacobian of p -> gp is the symmetric projector P = s * d d^T (line,
// segment interior; d unit), I - n n^T (plane), 0 (segment end points); JAX differentiates through it, so the
// cotangent of p is (I - P) e for the error derivative e.
struct ProjGoal {
  double gp[3];
  double d[3];    // unit direction (line / segment interior) or unit normal (plane)
  int mode;       // 0: P = d d^T, 1: P = I - d d^T, 2: P = 0
};
// normalize_vector (geometry.py:138-168): the zero vector stays zero
__device__ __forceinline__ double unit3(const double* v, double* u) {
  const bool zero = fabs(v[0]) <= 1e-8 && fabs(v[1]) <= 1e-8 && fabs(v[2]) <= 1e-8;
  const double nn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const double inv = zero ? 1.0 : 1.0 / nn;
  for (int c = 0; c < 3; ++c) u[c] = zero ? 0.0 : v[c] * inv;
  return zero ? 0.0 : nn;
}

__device__ __forceinline__ ProjGoal proj_goal(int kind, const double* p, const double* t) {
  ProjGoal r;
  if (kind == G_NODE_RESIDUAL_DIRECTION || kind == G_EDGE_DIRECTION) {   // residual.py:86-145, edge/direction.py:49-69
    unit3(t, r.gp);
    r.d[0] = r.d[1] = r.d[2] = 0.0;
    r.mode = 2;
    return r;
  }
  if (kind == G_NODE_RESIDUAL_PLANE) {              // residual.py:148-219: p - (p . n) n
    unit3(t, r.d);
    const double pn = p[0] * r.d[0] + p[1] * r.d[1] + p[2] * r.d[2];
    for (int c = 0; c < 3; ++c) r.gp[c] = p[c] - pn * r.d[c];
    r.mode = 1;
    return r;
  }
  if (kind == G_NODE_PLANE) {
    const double nn = sqrt(t[3] * t[3] + t[4] * t[4] + t[5] * t[5]);
    const bool zero = fabs(t[3]) <= 1e-8 && fabs(t[4]) <= 1e-8 && fabs(t[5]) <= 1e-8;   // normalize_vector, geometry.py:138-168
    const double inv = zero ? 1.0 : 1.0 / nn;
    for (int c = 0; c < 3; ++c) r.d[c] = zero ? 0.0 : t[3 + c] * inv;
    const double e = r.d[0] * (p[0] - t[0]) + r.d[1] * (p[1] - t[1]) + r.d[2] * (p[2] - t[2]);
    const double k = e / (r.d[0] * r.d[0] + r.d[1] * r.d[1] + r.d[2] * r.d[2]);
    for (int c = 0; c < 3; ++c) r.gp[c] = p[c] - r.d[c] * k;
    r.mode = 1;
    return r;
  }
  const double ab[3] = {t[3] - t[0], t[4] - t[1], t[5] - t[2]};
  const double l2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
  const double s = ((p[0] - t[0]) * ab[0] + (p[1] - t[1]) * ab[1] + (p[2] - t[2]) * ab[2]) / l2;
  const double il = 1.0 / sqrt(l2);
  for (int c = 0; c < 3; ++c) { r.gp[c] = t[c] + s * ab[c]; r.d[c] = ab[c] * il; }
  r.mode = 0;
  if (kind == G_NODE_SEGMENT) {
    double d1 = 0.0, d2 = 0.0;
    for (int c = 0; c < 3; ++c) { d1 += (t[c] - r.gp[c]) * (t[c] - r.gp[c]); d2 += (t[3 + c] - r.gp[c]) * (t[3 + c] - r.gp[c]); }
    if (d1 > l2 || d2 > l2) {                       // outside: the nearer end point
      for (int c = 0; c < 3; ++c) r.gp[c] = d1 < d2 ? t[c] : t[3 + c];
      r.mode = 2;
    }
  }
  return r;
}

__global__ void __launch_bounds__(kThreads) goals_loss_kernel(GoalArgs A) {
  __shared__ double red[16];
  __shared__ double s_mult[64];
  const int64_t b = blockIdx.x;
  const double* xyz = A.xyz + b * (int64_t)A.N * 3;
  // lengths / forces / residuals may be NULL when no goal of the program reads them (a loss of node goals only)
  const double* res = A.residuals ? A.residuals + b * (int64_t)A.N * 3 : nullptr;
  const double* len = A.lengths ? A.lengths + b * (int64_t)A.E : nullptr;
  const double* frc = A.forces ? A.forces + b * (int64_t)A.E : nullptr;
  // reverse mode: any cotangent array given; an array no term of the program reads may be NULL (not zeroed, not written)
  const bool bwd = A.xyz_bar || A.residuals_bar || A.lengths_bar || A.forces_bar;
  double* xb = A.xyz_bar ? A.xyz_bar + b * (int64_t)A.N * 3 : nullptr;
  double* rb = A.residuals_bar ? A.residuals_bar + b * (int64_t)A.N * 3 : nullptr;
  double* lb = A.lengths_bar ? A.lengths_bar + b * (int64_t)A.E : nullptr;
  double* fb = A.forces_bar ? A.forces_bar + b * (int64_t)A.E : nullptr;
  if (bwd) {
    for (int i = threadIdx.x; i < 3 * A.N; i += kThreads) {
      if (xb) xb[i] = 0.0;
      if (rb) rb[i] = 0.0;
    }
    for (int i = threadIdx.x; i < A.E; i += kThreads) {
      if (lb) lb[i] = 0.0;
      if (fb) fb[i] = 0.0;
    }
  }
  double loadpath = 0.0;
  if (A.need_loadpath) {
    double acc = 0.0;
    for (int e = threadIdx.x; e < A.E; e += kThreads) acc += fabs(len[e] * frc[e]);
    loadpath = block_sum(acc, red);
  }
  // aggregate goals (goals/edge/length.py:49-75, force.py:49-75): p = var(v) / M over the term's edge set, v = lengths
  // (M = mean v) or forces (M = mean |v|); evaluated by the whole CTA, kept per aggregate slot for both passes
  __shared__ double s_aggp[kMaxAgg], s_aggmu[kMaxAgg], s_aggV[kMaxAgg], s_aggM[kMaxAgg];
  if (A.agg_ptr && A.num_agg > 0) {
    for (int t = 0; t < A.T; ++t) {
      const int k = A.kind[t];
      if (k < G_EDGES_LENGTH_EQUAL) continue;
      const int a0 = A.agg_ptr[t], a1 = A.agg_ptr[t + 1], n = a1 - a0;
      if (k == G_NODES_COLINEAR) {
        // colinearity_points: sum_i |d_{i+1} - d_i|^2 / (0.5 (|d_i|^2 + |d_{i+1}|^2)) / max(n - 2, 1), d_i = P_{i+1} - P_i
        double acc = 0.0;
        for (int j = a0 + threadIdx.x; j + 2 < a1; j += kThreads) {
          const double* P0 = xyz + 3 * A.agg_index[j];
          const double* P1 = xyz + 3 * A.agg_index[j + 1];
          const double* P2 = xyz + 3 * A.agg_index[j + 2];
          double dd = 0.0, l0 = 0.0, l1 = 0.0;
          for (int c = 0; c < 3; ++c) {
            const double d0 = P1[c] - P0[c], d1 = P2[c] - P1[c];
            dd += (d1 - d0) * (d1 - d0); l0 += d0 * d0; l1 += d1 * d1;
          }
          acc += dd / (0.5 * (l0 + l1));
        }
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) s_aggp[A.index[t]] = acc / (n - 2 > 1 ? n - 2 : 1);
        continue;
      }
      const double* v = k == G_EDGES_LENGTH_EQUAL ? len : frc;
      double s1 = 0.0, sa = 0.0;
      for (int j = a0 + threadIdx.x; j < a1; j += kThreads) { const double x = v[A.agg_index[j]]; s1 += x; sa += fabs(x); }
      s1 = block_sum(s1, red);
      sa = block_sum(sa, red);
      const double mu = s1 / n;
      double s2 = 0.0;
      for (int j = a0 + threadIdx.x; j < a1; j += kThreads) { const double x = v[A.agg_index[j]] - mu; s2 += x * x; }
      s2 = block_sum(s2, red);
      if (threadIdx.x == 0) {
        const int slot = A.index[t];
        const double V = s2 / n, M = k == G_EDGES_LENGTH_EQUAL ? mu : sa / n;
        s_aggp[slot] = V / M; s_aggmu[slot] = mu; s_aggV[slot] = V; s_aggM[slot] = M;
      }
    }
    __syncthreads();
  }
  auto predict = [&](int k, int i, const double* t, double (&p)[3]) -> int {   // returns the number of components
    switch (k) {
      case G_NODE_POINT: case G_NODE_LINE: case G_NODE_PLANE: case G_NODE_SEGMENT:
        p[0] = xyz[3 * i]; p[1] = xyz[3 * i + 1]; p[2] = xyz[3 * i + 2]; return 3;
      case G_NODE_RESIDUAL_VECTOR: p[0] = res[3 * i]; p[1] = res[3 * i + 1]; p[2] = res[3 * i + 2]; return 3;
      case G_NODE_RESIDUAL_DIRECTION: case G_NODE_RESIDUAL_PLANE: unit3(res + 3 * i, p); return 3;
      case G_EDGE_DIRECTION: case G_EDGE_ANGLE: {      // goals/edge/direction.py:23-47, angle.py:74-101, geometry.py:46-99
        const int a = A.edge_nodes[2 * i], h = A.edge_nodes[2 * i + 1];
        const double v[3] = {xyz[3 * h] - xyz[3 * a], xyz[3 * h + 1] - xyz[3 * a + 1], xyz[3 * h + 2] - xyz[3 * a + 2]};
        if (k == G_EDGE_DIRECTION) { unit3(v, p); return 3; }
        double u[3], w[3];
        unit3(v, u); unit3(t + 1, w);
        p[1] = fmin(1.0, fmax(-1.0, u[0] * w[0] + u[1] * w[1] + u[2] * w[2]));   // the clipped cosine, kept for pass 2
        p[0] = acos(p[1]);
        return 1;
      }
      case G_NODE_RESIDUAL_FORCE:
        p[0] = sqrt(res[3 * i] * res[3 * i] + res[3 * i + 1] * res[3 * i + 1] + res[3 * i + 2] * res[3 * i + 2]);
        return 1;
      case G_EDGE_LENGTH: p[0] = len[i]; return 1;
      case G_EDGE_FORCE: p[0] = frc[i]; return 1;
      case G_EDGE_LOADPATH: p[0] = fabs(frc[i]) * len[i]; return 1;
      case G_NODE_X: case G_NODE_Y: case G_NODE_Z: p[0] = xyz[3 * i + (k - G_NODE_X)]; return 1;
      case G_EDGES_LENGTH_EQUAL: case G_EDGES_FORCE_EQUAL: case G_NODES_COLINEAR: p[0] = s_aggp[i]; return 1;
      default: p[0] = loadpath; return 1;
    }
  };
  // ---- pass 1: group sums, loss
  double loss = 0.0;
  for (int g = 0; g < A.G; ++g) {
    double acc = 0.0;
    for (int t = A.group_ptr[g] + threadIdx.x; t < A.group_ptr[g + 1]; t += kThreads) {
      double p[3];
      const int k = A.kind[t];
      const int nc = predict(k, A.index[t], A.target + 6 * t, p);
      if (k >= G_NODE_LINE && k <= G_EDGE_DIRECTION) {
        const ProjGoal pg = proj_goal(k, p, A.target + 6 * t);
        for (int c = 0; c < 3; ++c) acc += err_value(A.err[t], A.weight[t], p[c], pg.gp[c]);
      } else {
        for (int c = 0; c < nc; ++c) acc += err_value(A.err[t], A.weight[t], p[c], A.target[6 * t + c]);
      }
    }
    const double S = block_sum(acc, red);
    const double u = A.ginner[g] * S;
    const bool root = A.gfinal[g] == 1;
    loss += A.gouter[g] * (root ? sqrt(u) : u);
    if (threadIdx.x == 0) {
      if (A.gsum) A.gsum[b * A.G + g] = S;
      if (g < 64) s_mult[g] = (A.loss_cot ? A.loss_cot[b] : 1.0) * A.gouter[g] * A.ginner[g] * (root ? 0.5 / sqrt(u) : 1.0);   // loss_bar * d loss / d S_g
    }
  }
  if (threadIdx.x == 0 && A.loss) A.loss[b] = loss;
  if (!bwd) return;
  __syncthreads();
  // ---- pass 2: cotangents of the state
  for (int g = 0; g < A.G; ++g) {
    const double mult = s_mult[g];
    for (int t = A.group_ptr[g] + threadIdx.x; t < A.group_ptr[g + 1]; t += kThreads) {
      const int k = A.kind[t], i = A.index[t], ek = A.err[t];
      const double w = A.weight[t];
      double p[3];
      const int nc = predict(k, i, A.target + 6 * t, p);
      double d[3];
      if (k >= G_NODE_LINE && k <= G_EDGE_DIRECTION) {
        const ProjGoal pg = proj_goal(k, p, A.target + 6 * t);
        for (int c = 0; c < 3; ++c) d[c] = mult * err_deriv(ek, w, p[c], pg.gp[c]);
        const double de = pg.d[0] * d[0] + pg.d[1] * d[1] + pg.d[2] * d[2];
        double cp[3];
        for (int c = 0; c < 3; ++c) {                     // (I - P) d
          const double Pd = pg.mode == 0 ? pg.d[c] * de : pg.mode == 1 ? d[c] - pg.d[c] * de : 0.0;
          cp[c] = d[c] - Pd;
        }
        if (k == G_EDGE_DIRECTION) {                      // through v -> v / |v|, v = xyz[head] - xyz[tail]
          const int a = A.edge_nodes[2 * i], h = A.edge_nodes[2 * i + 1];
          const double v[3] = {xyz[3 * h] - xyz[3 * a], xyz[3 * h + 1] - xyz[3 * a + 1], xyz[3 * h + 2] - xyz[3 * a + 2]};
          double u[3];
          const double nn = unit3(v, u);
          if (nn > 0.0) {
            const double pc = p[0] * cp[0] + p[1] * cp[1] + p[2] * cp[2];
            for (int c = 0; c < 3; ++c) {
              const double gv = (cp[c] - p[c] * pc) / nn;
              gadd(xb, 3 * h + c, gv);
              gadd(xb, 3 * a + c, -gv);
            }
          }
        } else if (k >= G_NODE_RESIDUAL_DIRECTION) {      // through r -> r / |r|: (I - p p^T) / |r|
          double u[3];
          const double nn = unit3(res + 3 * i, u);
          if (nn > 0.0) {
            const double pc = p[0] * cp[0] + p[1] * cp[1] + p[2] * cp[2];
            for (int c = 0; c < 3; ++c) gadd(rb, 3 * i + c, (cp[c] - p[c] * pc) / nn);
          }
        } else {
          for (int c = 0; c < 3; ++c) gadd(xb, 3 * i + c, cp[c]);
        }
        continue;
      }
      for (int c = 0; c < nc; ++c) d[c] = mult * err_deriv(ek, w, p[c], A.target[6 * t + c]);
      switch (k) {
        case G_NODE_POINT:
          for (int c = 0; c < 3; ++c) gadd(xb, 3 * i + c, d[c]);
          break;
        case G_NODE_RESIDUAL_VECTOR:
          for (int c = 0; c < 3; ++c) gadd(rb, 3 * i + c, d[c]);
          break;
        case G_NODE_RESIDUAL_FORCE:
          if (p[0] > 0.0)
            for (int c = 0; c < 3; ++c) gadd(rb, 3 * i + c, d[0] * res[3 * i + c] / p[0]);
          break;
        case G_NODE_X: case G_NODE_Y: case G_NODE_Z: gadd(xb, 3 * i + (k - G_NODE_X), d[0]); break;
        case G_EDGE_ANGLE: {
          const int a = A.edge_nodes[2 * i], h = A.edge_nodes[2 * i + 1];
          const double v[3] = {xyz[3 * h] - xyz[3 * a], xyz[3 * h + 1] - xyz[3 * a + 1], xyz[3 * h + 2] - xyz[3 * a + 2]};
          double u[3], wv[3];
          const double nn = unit3(v, u);
          unit3(A.target + 6 * t + 1, wv);
          const double c = p[1], s2 = 1.0 - c * c;
          if (nn > 0.0 && s2 > 0.0) {
            const double f = -d[0] / (sqrt(s2) * nn);     // d acos / d cos, through u = v / |v|
            for (int cc = 0; cc < 3; ++cc) {
              const double gv = f * (wv[cc] - u[cc] * c);
              gadd(xb, 3 * h + cc, gv);
              gadd(xb, 3 * a + cc, -gv);
            }
          }
          break;
        }
        case G_EDGE_LENGTH: gadd(lb, i, d[0]); break;
        case G_EDGE_FORCE: gadd(fb, i, d[0]); break;
        case G_EDGE_LOADPATH: {
          const double f = frc[i];
          gadd(fb, i, d[0] * (f > 0.0 ? 1.0 : (f < 0.0 ? -1.0 : 0.0)) * len[i]);
          gadd(lb, i, d[0] * fabs(f));
          break;
        }
        default: break;   // network load path: every edge, by the whole CTA below
      }
    }
    __syncthreads();
    // aggregate and network load path terms of this group: all threads scatter over their element sets (the
    // serial scan of the group's terms is skipped when the program has none of them)
    if (A.num_agg == 0 && !A.need_loadpath) continue;
    for (int t = A.group_ptr[g]; t < A.group_ptr[g + 1]; ++t) {
      const int kk = A.kind[t];
      if (kk == G_NODES_COLINEAR) {
        const int slot = A.index[t], a0 = A.agg_ptr[t], a1 = A.agg_ptr[t + 1], n = a1 - a0;
        const double d0s = mult * err_deriv(A.err[t], A.weight[t], s_aggp[slot], A.target[6 * t]) / (n - 2 > 1 ? n - 2 : 1);
        for (int j = a0 + threadIdx.x; j + 2 < a1; j += kThreads) {
          const int i0 = A.agg_index[j], i1 = A.agg_index[j + 1], i2 = A.agg_index[j + 2];
          double d0[3], d1[3], dd = 0.0, l0 = 0.0, l1 = 0.0;
          for (int c = 0; c < 3; ++c) {
            d0[c] = xyz[3 * i1 + c] - xyz[3 * i0 + c]; d1[c] = xyz[3 * i2 + c] - xyz[3 * i1 + c];
            dd += (d1[c] - d0[c]) * (d1[c] - d0[c]); l0 += d0[c] * d0[c]; l1 += d1[c] * d1[c];
          }
          const double lbar = 0.5 * (l0 + l1);
          for (int c = 0; c < 3; ++c) {
            const double gD = 2.0 * (d1[c] - d0[c]) / lbar;       // d T / d (d1 - d0)
            const double g0 = -gD - dd / (lbar * lbar) * d0[c];   // d T / d d0  (lbar = (|d0|^2 + |d1|^2) / 2)
            const double g1 = gD - dd / (lbar * lbar) * d1[c];    // d T / d d1
            gadd(xb, 3 * i0 + c, -d0s * g0);
            gadd(xb, 3 * i1 + c, d0s * (g0 - g1));
            gadd(xb, 3 * i2 + c, d0s * g1);
          }
        }
        continue;
      }
      if (kk == G_EDGES_LENGTH_EQUAL || kk == G_EDGES_FORCE_EQUAL) {
        const int slot = A.index[t], a0 = A.agg_ptr[t], a1 = A.agg_ptr[t + 1], n = a1 - a0;
        const double d0 = mult * err_deriv(A.err[t], A.weight[t], s_aggp[slot], A.target[6 * t]);
        const double mu = s_aggmu[slot], V = s_aggV[slot], M = s_aggM[slot];
        const double* v = kk == G_EDGES_LENGTH_EQUAL ? len : frc;
        double* vb = kk == G_EDGES_LENGTH_EQUAL ? lb : fb;
        for (int j = a0 + threadIdx.x; j < a1; j += kThreads) {
          const int e = A.agg_index[j];
          const double x = v[e];
          const double dM = kk == G_EDGES_LENGTH_EQUAL ? 1.0 : (x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : 0.0));
          gadd(vb, e, d0 * (2.0 * (x - mu) / (n * M) - V * dM / (M * M * n)));
        }
        continue;
      }
      if (A.kind[t] != G_NETWORK_LOADPATH) continue;
      const double d0 = mult * err_deriv(A.err[t], A.weight[t], loadpath, A.target[6 * t]);
      for (int e = threadIdx.x; e < A.E; e += kThreads) {
        const double pl = len[e] * frc[e];
        const double sg = pl > 0.0 ? 1.0 : (pl < 0.0 ? -1.0 : 0.0);
        gadd(fb, e, d0 * sg * len[e]);
        gadd(lb, e, d0 * sg * frc[e]);
      }
    }
    __syncthreads();
  }
}

}  // namespace

int launch_goals_loss(const fdm_plan* p, int n_terms, const int32_t* kind, const int32_t* index, const int32_t* err,
                      const double* weight, const double* target, int n_groups, const int32_t* group_ptr,
                      const int32_t* gfinal, const double* ginner, const double* gouter, int need_loadpath,
                      const int32_t* agg_ptr, const int32_t* agg_index, int num_agg,
                      const double* xyz, const double* lengths, const double* forces, const double* residuals,
                      double* loss, double* gsum, const double* loss_cot, double* xyz_bar, double* lengths_bar,
                      double* forces_bar, double* residuals_bar, int64_t batch, cudaStream_t st) {
  GoalArgs A;
  A.T = n_terms; A.G = n_groups; A.N = (int)p->N; A.E = (int)p->E;
  A.kind = kind; A.index = index; A.err = err; A.group_ptr = group_ptr; A.edge_nodes = p->d.edge_nodes;
  A.agg_ptr = agg_ptr; A.agg_index = agg_index; A.num_agg = num_agg;
  A.weight = weight; A.target = target; A.gfinal = gfinal; A.ginner = ginner; A.gouter = gouter;
  A.xyz = xyz; A.lengths = lengths; A.forces = forces; A.residuals = residuals;
  A.loss = loss; A.gsum = gsum; A.loss_cot = loss_cot;
  A.xyz_bar = xyz_bar; A.lengths_bar = lengths_bar; A.forces_bar = forces_bar; A.residuals_bar = residuals_bar;
  A.need_loadpath = need_loadpath;
  goals_loss_kernel<<<(unsigned)batch, kThreads, 0, st>>>(A);
  return fdm_check_launch("goals_loss");
}

// Which state arrays does a goal kind differentiate into?  bit 0 xyz, 1 lengths, 2 forces, 3 residuals.  A caller of
// fdm_goals_loss may pass NULL for the cotangent arrays no term of its program touches (they are then neither zeroed
// nor written, and fdm_state_vjp takes NULL as zero).
extern "C" int fdm_goal_kind_arrays(int kind) {
  switch (kind) {
    case G_NODE_POINT: case G_NODE_X: case G_NODE_Y: case G_NODE_Z: case G_NODE_LINE: case G_NODE_PLANE:
    case G_NODE_SEGMENT: case G_EDGE_DIRECTION: case G_EDGE_ANGLE: case G_NODES_COLINEAR: return 1;
    case G_EDGE_LENGTH: case G_EDGES_LENGTH_EQUAL: return 2;
    case G_EDGE_FORCE: case G_EDGES_FORCE_EQUAL: return 4;
    case G_EDGE_LOADPATH: case G_NETWORK_LOADPATH: return 2 | 4;
    case G_NODE_RESIDUAL_FORCE: case G_NODE_RESIDUAL_VECTOR: case G_NODE_RESIDUAL_DIRECTION:
    case G_NODE_RESIDUAL_PLANE: return 8;
  }
  return 15;
}
