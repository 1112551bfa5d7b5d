// Host-side plan construction: topology -> static index arrays.
//
// Reproduces, without scipy, the arrays the reference derives once per topology
// (reference file:line under src/jax_fdm/equilibrium/structures/):
//   graphs.py:176-211          signed incidence C (implicit here: edge_nodes + incidence lists)
//   structures.py:152-188      indices_free / indices_fixed / indices_freefixed
//   structures.py:248-288      index_array (CSC, sorted indices, explicit 0 on the diagonal)
//   structures.py:290-314      diag_indices
//   structures.py:316-350      diags (per free node: incident edge ids, ascending)
// plus two orderings the reference does not have: a bandwidth-reducing permutation for the
// batched band Cholesky and a partition of the free nodes into compact aggregates for the
// two-level PCG.
#include "plan.h"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <queue>

static thread_local std::string g_err;
void fdm_set_error(const std::string& msg) { g_err = msg; }
extern "C" const char* fdm_last_error(void) { return g_err.c_str(); }

int fdm_build_topology(fdm_plan& p, const int64_t* edges, int64_t E, int64_t N,
                       const int64_t* supports) {
  if (!edges || !supports || E <= 0 || N <= 0) {
    fdm_set_error("plan: need at least one edge and one node");
    return E <= 0 ? FDM_ERR_NO_SUPPORT : FDM_ERR_INVALID;
  }
  if (E >= (int64_t(1) << 30) || N >= (int64_t(1) << 31) - 1) {
    fdm_set_error("plan: E must be < 2^30 and N < 2^31-1");
    return FDM_ERR_UNSUPPORTED;
  }
  p.N = N;
  p.E = E;

  // ---- free / fixed maps (structures.py:152-188) ----
  p.node_slot.assign(N, 0);
  p.indices_free.clear();
  p.indices_fixed.clear();
  for (int64_t n = 0; n < N; ++n) {
    if (supports[n] == 0) {
      p.node_slot[n] = (int32_t)p.indices_free.size();
      p.indices_free.push_back(n);
    } else {
      p.node_slot[n] = -(int32_t)p.indices_fixed.size() - 1;
      p.indices_fixed.push_back(n);
    }
  }
  p.Nf = (int64_t)p.indices_free.size();
  p.Ns = (int64_t)p.indices_fixed.size();
  if (p.Ns == 0) {
    fdm_set_error("plan: structure has no supports (fdm.py:491-522 asserts at least one)");
    return FDM_ERR_NO_SUPPORT;
  }
  if (p.Nf == 0) {
    fdm_set_error("plan: structure has no free nodes");
    return FDM_ERR_INVALID;
  }
  // indices_freefixed = argsort(concat(free, fixed)): position of node n in the stacked order
  p.indices_freefixed.assign(N, 0);
  for (int64_t i = 0; i < p.Nf; ++i) p.indices_freefixed[p.indices_free[i]] = i;
  for (int64_t s = 0; s < p.Ns; ++s) p.indices_freefixed[p.indices_fixed[s]] = p.Nf + s;
  p.free_node.resize(p.Nf);
  p.fixed_node.resize(p.Ns);
  for (int64_t i = 0; i < p.Nf; ++i) p.free_node[i] = (int32_t)p.indices_free[i];
  for (int64_t s = 0; s < p.Ns; ++s) p.fixed_node[s] = (int32_t)p.indices_fixed[s];

  // ---- edges and node incidence lists (graphs.py:176-211 in list form) ----
  p.edge_nodes.resize(2 * E);
  p.ninc_ptr.assign(N + 1, 0);
  for (int64_t e = 0; e < E; ++e) {
    int64_t t = edges[2 * e], h = edges[2 * e + 1];
    if (t < 0 || h < 0 || t >= N || h >= N) {
      fdm_set_error("plan: edge endpoint out of range");
      return FDM_ERR_INVALID;
    }
    if (t == h) {
      fdm_set_error("plan: self loop (edge " + std::to_string(e) + ")");
      return FDM_ERR_MULTI_EDGE;
    }
    p.edge_nodes[2 * e] = (int32_t)t;
    p.edge_nodes[2 * e + 1] = (int32_t)h;
    p.ninc_ptr[t + 1]++;
    p.ninc_ptr[h + 1]++;
  }
  for (int64_t n = 0; n < N; ++n) p.ninc_ptr[n + 1] += p.ninc_ptr[n];
  p.ninc_edge.resize(2 * E);
  p.ninc_other.resize(2 * E);
  {
    std::vector<int32_t> fill(p.ninc_ptr.begin(), p.ninc_ptr.end() - 1);
    for (int64_t e = 0; e < E; ++e) {  // ascending e => each list is ascending in edge id
      int32_t t = p.edge_nodes[2 * e], h = p.edge_nodes[2 * e + 1];
      p.ninc_edge[fill[t]] = (int32_t)(e << 1);
      p.ninc_other[fill[t]++] = h;
      p.ninc_edge[fill[h]] = (int32_t)((e << 1) | 1);
      p.ninc_other[fill[h]++] = t;
    }
  }

  // ---- diags (structures.py:316-350): per free node, incident edges ascending ----
  p.diags_indptr.assign(p.Nf + 1, 0);
  for (int64_t i = 0; i < p.Nf; ++i) {
    int64_t n = p.indices_free[i];
    p.diags_indptr[i + 1] = p.diags_indptr[i] + (p.ninc_ptr[n + 1] - p.ninc_ptr[n]);
  }
  p.dnnz = p.diags_indptr[p.Nf];
  p.diags_indices.resize(p.dnnz);
  for (int64_t i = 0; i < p.Nf; ++i) {
    int64_t n = p.indices_free[i];
    int32_t o = p.diags_indptr[i];
    for (int32_t k = p.ninc_ptr[n]; k < p.ninc_ptr[n + 1]; ++k)
      p.diags_indices[o++] = p.ninc_edge[k] >> 1;
  }

  // ---- sup_mask: free rows with at least one support neighbour (they carry a load-term sum) ----
  p.sup_mask.assign((size_t)(p.Nf + 31) / 32, 0u);
  for (int64_t i = 0; i < p.Nf; ++i) {
    int64_t n = p.indices_free[i];
    for (int32_t k = p.ninc_ptr[n]; k < p.ninc_ptr[n + 1]; ++k)
      if (p.node_slot[p.ninc_other[k]] < 0) {
        p.sup_mask[(size_t)i >> 5] |= 1u << (i & 31);
        break;
      }
  }

  // ---- index_array (structures.py:248-288), diag_indices (:290-314) ----
  // column i (== row i, the pattern is symmetric) holds, sorted by row index, the diagonal with
  // label 0 and every free neighbour j with label e+1 of the connecting edge.
  p.index_indptr.assign(p.Nf + 1, 0);
  for (int64_t i = 0; i < p.Nf; ++i) {
    int64_t n = p.indices_free[i];
    int32_t cnt = 1;
    for (int32_t k = p.ninc_ptr[n]; k < p.ninc_ptr[n + 1]; ++k)
      if (p.node_slot[p.ninc_other[k]] >= 0) ++cnt;
    p.index_indptr[i + 1] = p.index_indptr[i] + cnt;
  }
  p.nnz = p.index_indptr[p.Nf];
  if (p.nnz >= (int64_t(1) << 31) - 1) {
    fdm_set_error("plan: nnz must be < 2^31-1");
    return FDM_ERR_UNSUPPORTED;
  }
  p.index_indices.resize(p.nnz);
  p.index_data.resize(p.nnz);
  p.entry_edge.resize(p.nnz);
  p.diag_indices.resize(p.Nf);
  p.diag_pos.resize(p.Nf);
  std::vector<std::pair<int32_t, int32_t>> row;
  for (int64_t i = 0; i < p.Nf; ++i) {
    int64_t n = p.indices_free[i];
    row.clear();
    row.emplace_back((int32_t)i, -1);
    for (int32_t k = p.ninc_ptr[n]; k < p.ninc_ptr[n + 1]; ++k) {
      int32_t j = p.node_slot[p.ninc_other[k]];
      if (j >= 0) row.emplace_back(j, p.ninc_edge[k] >> 1);
    }
    std::sort(row.begin(), row.end());
    int32_t o = p.index_indptr[i];
    for (size_t r = 0; r < row.size(); ++r) {
      if (r > 0 && row[r].first == row[r - 1].first) {
        fdm_set_error("plan: parallel edges between the same pair of nodes (edge " +
                      std::to_string(row[r].second) + ")");
        return FDM_ERR_MULTI_EDGE;
      }
      p.index_indices[o + r] = row[r].first;
      p.entry_edge[o + r] = row[r].second;
      p.index_data[o + r] = (int64_t)row[r].second + 1;  // e+1 off-diagonal, 0 on the diagonal
      if (row[r].first == (int32_t)i) {
        p.diag_indices[i] = o + (int64_t)r;
        p.diag_pos[i] = o + (int32_t)r;
      }
    }
  }
  // parallel edges with a support at one end do not show in the free-free pattern; catch them too
  {
    std::vector<int32_t> seen;
    for (int64_t n = 0; n < N; ++n) {
      seen.assign(p.ninc_other.begin() + p.ninc_ptr[n], p.ninc_other.begin() + p.ninc_ptr[n + 1]);
      std::sort(seen.begin(), seen.end());
      if (std::adjacent_find(seen.begin(), seen.end()) != seen.end()) {
        fdm_set_error("plan: parallel edges at node " + std::to_string(n));
        return FDM_ERR_MULTI_EDGE;
      }
    }
  }
  return FDM_OK;
}

// ---------------------------------------------------------------------------------------
// graph helpers on the free-node adjacency (K's pattern without the diagonal)
// ---------------------------------------------------------------------------------------
namespace {

struct Adj {
  const fdm_plan& p;
  explicit Adj(const fdm_plan& pl) : p(pl) {}
  template <class F>
  void for_each(int32_t i, F f) const {
    for (int32_t k = p.index_indptr[i]; k < p.index_indptr[i + 1]; ++k) {
      int32_t j = p.index_indices[k];
      if (j != i) f(j);
    }
  }
  int32_t degree(int32_t i) const { return p.index_indptr[i + 1] - p.index_indptr[i] - 1; }
};

// BFS restricted to nodes with mask[v] == tag; returns visitation order, fills level[].
void bfs(const Adj& g, int32_t start, const std::vector<int32_t>& mask, int32_t tag,
         std::vector<int32_t>& level, std::vector<int32_t>& order, bool sort_by_degree) {
  order.clear();
  order.push_back(start);
  level[start] = 0;
  std::vector<int32_t> nb;
  for (size_t head = 0; head < order.size(); ++head) {
    int32_t v = order[head];
    nb.clear();
    g.for_each(v, [&](int32_t w) {
      if (mask[w] == tag && level[w] < 0) {
        level[w] = level[v] + 1;
        nb.push_back(w);
      }
    });
    if (sort_by_degree)
      std::sort(nb.begin(), nb.end(), [&](int32_t a, int32_t b) {
        int32_t da = g.degree(a), db = g.degree(b);
        return da != db ? da < db : a < b;
      });
    order.insert(order.end(), nb.begin(), nb.end());
  }
}

// pseudo-peripheral node of the component of `seed` inside {mask == tag}
int32_t pseudo_peripheral(const Adj& g, int32_t seed, const std::vector<int32_t>& mask, int32_t tag,
                          std::vector<int32_t>& level, std::vector<int32_t>& order) {
  int32_t cur = seed, ecc = -1;
  for (int it = 0; it < 8; ++it) {
    bfs(g, cur, mask, tag, level, order, false);
    int32_t last = order.back(), e = level[last];
    // among the deepest level pick the smallest degree
    for (auto v : order)
      if (level[v] == e && g.degree(v) < g.degree(last)) last = v;
    for (auto v : order) level[v] = -1;
    if (e <= ecc) break;
    ecc = e;
    cur = last;
  }
  return cur;
}

int32_t bandwidth_of(const fdm_plan& p, const std::vector<int32_t>& perm) {
  int32_t bw = 0;
  for (int64_t i = 0; i < p.Nf; ++i)
    for (int32_t k = p.index_indptr[i]; k < p.index_indptr[i + 1]; ++k)
      bw = std::max(bw, std::abs(perm[i] - perm[p.index_indices[k]]));
  return bw;
}

}  // namespace

// Reverse Cuthill-McKee over every component; keep the natural order when it is at least as good.
void fdm_build_band_order(fdm_plan& p) {
  const int32_t n = (int32_t)p.Nf;
  Adj g(p);
  std::vector<int32_t> mask(n, 0), level(n, -1), order, cm;
  cm.reserve(n);
  for (int32_t s = 0; s < n; ++s) {
    if (mask[s] != 0) continue;
    int32_t start = pseudo_peripheral(g, s, mask, 0, level, order);
    bfs(g, start, mask, 0, level, order, true);
    for (auto v : order) {
      mask[v] = 1;
      level[v] = -1;
      cm.push_back(v);
    }
  }
  std::vector<int32_t> rcm_perm(n), nat(n);
  for (int32_t k = 0; k < n; ++k) rcm_perm[cm[n - 1 - k]] = k;
  std::iota(nat.begin(), nat.end(), 0);
  int32_t bw_nat = bandwidth_of(p, nat), bw_rcm = bandwidth_of(p, rcm_perm);
  if (bw_rcm < bw_nat) {
    p.band_perm = rcm_perm;
    p.hbw = bw_rcm;
  } else {
    p.band_perm = nat;
    p.hbw = bw_nat;
  }
  p.band_iperm.resize(n);
  for (int32_t i = 0; i < n; ++i) p.band_iperm[p.band_perm[i]] = i;
  // band_kidx[j*(hbw+1) + d]: where A(j+d, j) (band order, lower triangle) sits in K.data; -1 = zero.
  // Built only when the band is narrow enough to be useful (<= 255); wider => iterative solvers.
  p.band_kidx.clear();
  if (p.hbw <= 255 && (int64_t)n * (p.hbw + 1) < (int64_t(1) << 28)) {
    const int32_t w = p.hbw + 1;
    p.band_kidx.assign((size_t)n * w, -1);
    for (int32_t i = 0; i < n; ++i) {
      int32_t r = p.band_perm[i];
      for (int32_t k = p.index_indptr[i]; k < p.index_indptr[i + 1]; ++k) {
        int32_t c = p.band_perm[p.index_indices[k]];
        if (c <= r) p.band_kidx[(size_t)c * w + (r - c)] = k;
      }
    }
  }
  // band_rowkidx[R*32 + t]: where A(R, R-t) sits in K.data (t = 0..31), for the warp-per-design
  // Cholesky (hbw <= 31).  Rows are padded to (ceil(n/32)+1)*32 with identity rows:
  // -1 = structural zero (or column < 0), -2 = one (diagonal of a padding row).
  p.band_rowkidx.clear();
  p.band_rowedge.clear();
  p.band_stage_ptr.clear();
  p.band_stage_kent.clear();
  p.band_stage_eent.clear();
  p.band_row_diag.clear();
  p.band_row_sup.clear();
  p.band_row_node.clear();
  if (p.hbw <= 31) {
    const int64_t rows = ((int64_t)(n + 31) / 32 + 1) * 32;
    p.band_rowkidx.assign((size_t)rows * 32, -1);
    for (int64_t r = n; r < rows; ++r) p.band_rowkidx[(size_t)r * 32] = -2;
    for (int32_t i = 0; i < n; ++i) {
      int32_t r = p.band_perm[i];
      for (int32_t k = p.index_indptr[i]; k < p.index_indptr[i + 1]; ++k) {
        int32_t c = p.band_perm[p.index_indices[k]];
        if (c <= r) p.band_rowkidx[(size_t)r * 32 + (r - c)] = k;
      }
    }
    // same table in terms of q: edge id of the off-diagonal entry, -3 = diagonal (sum over `diags`),
    // -1 / -2 as above.  Lets the batched factorisation assemble its rows straight from q.
    p.band_rowedge.assign(p.band_rowkidx.size(), -1);
    for (size_t i = 0; i < p.band_rowkidx.size(); ++i) {
      const int32_t k = p.band_rowkidx[i];
      p.band_rowedge[i] = k >= 0 ? (p.entry_edge[k] >= 0 ? p.entry_edge[k] : -3) : k;
    }
    // The tables are mostly zeros (a row of K has degree + 1 entries, the band 32 slots): per block of 32
    // rows keep only the non-zero entries as (r*32 + slot, index) pairs, slot = (r - t) & 31 = column & 31,
    // in groups of 32 so that a warp reads a group with one coalesced load.
    const int64_t blocks = rows / 32;
    p.band_stage_ptr.assign((size_t)blocks + 1, 0);
    p.band_stage_kent.clear();
    p.band_stage_eent.clear();
    for (int64_t b = 0; b < blocks; ++b) {
      p.band_stage_ptr[(size_t)b] = (int32_t)(p.band_stage_kent.size() / 64);
      for (int32_t r = 0; r < 32; ++r)
        for (int32_t t = 0; t < 32; ++t) {
          const size_t i = (size_t)(b * 32 + r) * 32 + t;
          if (p.band_rowkidx[i] == -1) continue;
          const int32_t pos = r * 32 + ((r - t) & 31);
          const int32_t e = p.band_rowedge[i];
          p.band_stage_kent.push_back(pos);
          p.band_stage_kent.push_back(p.band_rowkidx[i]);
          p.band_stage_eent.push_back(pos);
          p.band_stage_eent.push_back(e == -3 ? -1 : e);
        }
      while (p.band_stage_kent.size() % 64) {
        p.band_stage_kent.push_back(0); p.band_stage_kent.push_back(-4);
        p.band_stage_eent.push_back(0); p.band_stage_eent.push_back(-4);
      }
    }
    p.band_stage_ptr[(size_t)blocks] = (int32_t)(p.band_stage_kent.size() / 64);
    // per band row, padded and vector-loadable: what the row needs from q, x_s and loads when the factorisation
    // assembles it in place -- the same terms, in the same order, as the assembly kernel adds them
    p.band_row_diag.assign((size_t)rows * 8, -1);
    p.band_row_sup.assign((size_t)rows * 8, -1);
    p.band_row_node.assign((size_t)rows, -1);
    for (int32_t R = 0; R < n; ++R) {
      const int32_t f = p.band_iperm[R];
      const int32_t node = p.free_node[f];
      p.band_row_node[(size_t)R] = node;
      const int32_t d0 = p.diags_indptr[f], d1 = p.diags_indptr[f + 1];
      if (d1 - d0 > 8) p.band_row_diag[(size_t)R * 8] = -2;
      else for (int32_t k = d0; k < d1; ++k) p.band_row_diag[(size_t)R * 8 + (k - d0)] = p.diags_indices[k];
      int32_t cnt = 0;
      bool overflow = false;
      for (int32_t k = p.ninc_ptr[node]; k < p.ninc_ptr[node + 1]; ++k) {
        const int32_t slot = p.node_slot[p.ninc_other[k]];
        if (slot >= 0) continue;
        if (cnt == 4) { overflow = true; break; }
        p.band_row_sup[(size_t)R * 8 + 2 * cnt] = p.ninc_edge[k] >> 1;
        p.band_row_sup[(size_t)R * 8 + 2 * cnt + 1] = -slot - 1;
        ++cnt;
      }
      if (overflow) {
        for (int j = 0; j < 8; ++j) p.band_row_sup[(size_t)R * 8 + j] = -1;
        p.band_row_sup[(size_t)R * 8] = -2;
      }
    }
  }
}

// Tables of one half of the two-sided factorisation (plan.h: BandHalf).  `vmap[R]` = band position of the half's
// row R, or -1 for an identity padding row; rows in [excl_lo, excl_hi) have their mutual entries, diagonals and
// right-hand sides left out (the other half owns them).  One trailing identity block is appended: the kernels
// always stage one block ahead.
static void build_half_tables(const fdm_plan& p, const std::vector<int32_t>& vmap, int32_t excl_lo, int32_t excl_hi,
                              BandHalf& out) {
  const int32_t n = (int32_t)p.Nf;
  const int64_t rows = (int64_t)vmap.size() + 32;
  std::vector<int32_t> inv(n, -1);
  for (size_t R = 0; R < vmap.size(); ++R)
    if (vmap[R] >= 0 && vmap[R] < n) inv[vmap[R]] = (int32_t)R;
  auto excluded = [&](int64_t R) { return R >= excl_lo && R < excl_hi; };
  std::vector<int32_t> rowk((size_t)rows * 32, -1);
  out.iperm.assign((size_t)rows, -1);
  out.row_diag.assign((size_t)rows * 8, -1);
  out.row_sup.assign((size_t)rows * 8, -1);
  out.row_node.assign((size_t)rows, -1);
  for (int64_t R = 0; R < rows; ++R) {
    const int32_t bp = R < (int64_t)vmap.size() ? vmap[(size_t)R] : -1;
    if (bp < 0 || bp >= n) {
      rowk[(size_t)R * 32] = -2;
      continue;
    }
    const int32_t f = p.band_iperm[bp];
    for (int32_t k = p.index_indptr[f]; k < p.index_indptr[f + 1]; ++k) {
      const int32_t C = inv[p.band_perm[p.index_indices[k]]];
      if (C < 0 || C > R) continue;                        // the other half's column, or the upper triangle
      if (excluded(R) && excluded(C)) continue;
      rowk[(size_t)R * 32 + (R - C)] = k;
    }
    if (excluded(R)) continue;                             // no diagonal, no right-hand side, no output row
    out.iperm[(size_t)R] = f;
    const int32_t node = p.free_node[f];
    out.row_node[(size_t)R] = node;
    const int32_t d0 = p.diags_indptr[f], d1 = p.diags_indptr[f + 1];
    if (d1 - d0 > 8) out.row_diag[(size_t)R * 8] = -2;
    else for (int32_t k = d0; k < d1; ++k) out.row_diag[(size_t)R * 8 + (k - d0)] = p.diags_indices[k];
    int32_t cnt = 0;
    bool overflow = false;
    for (int32_t k = p.ninc_ptr[node]; k < p.ninc_ptr[node + 1]; ++k) {
      const int32_t slot = p.node_slot[p.ninc_other[k]];
      if (slot >= 0) continue;
      if (cnt == 4) { overflow = true; break; }
      out.row_sup[(size_t)R * 8 + 2 * cnt] = p.ninc_edge[k] >> 1;
      out.row_sup[(size_t)R * 8 + 2 * cnt + 1] = -slot - 1;
      ++cnt;
    }
    if (overflow) {
      for (int j = 0; j < 8; ++j) out.row_sup[(size_t)R * 8 + j] = -1;
      out.row_sup[(size_t)R * 8] = -2;
    }
  }
  const int64_t blocks = rows / 32;
  out.stage_ptr.assign((size_t)blocks + 1, 0);
  out.stage_kent.clear();
  out.stage_eent.clear();
  for (int64_t b = 0; b < blocks; ++b) {
    out.stage_ptr[(size_t)b] = (int32_t)(out.stage_kent.size() / 64);
    for (int32_t r = 0; r < 32; ++r)
      for (int32_t t = 0; t < 32; ++t) {
        const int32_t k = rowk[(size_t)(b * 32 + r) * 32 + t];
        if (k == -1) continue;
        const int32_t pos = r * 32 + ((r - t) & 31);
        const int32_t e = k >= 0 ? (p.entry_edge[k] >= 0 ? p.entry_edge[k] : -1) : k;   // diagonal: -1 (assembled per row)
        out.stage_kent.push_back(pos);
        out.stage_kent.push_back(k);
        out.stage_eent.push_back(pos);
        out.stage_eent.push_back(e);
      }
    while (out.stage_kent.size() % 64) {
      out.stage_kent.push_back(0); out.stage_kent.push_back(-4);
      out.stage_eent.push_back(0); out.stage_eent.push_back(-4);
    }
  }
  out.stage_ptr[(size_t)blocks] = (int32_t)(out.stage_kent.size() / 64);
}

// Two-sided factorisation tables (four-column panel kernels only: hbw <= 29; at least two blocks per side).
// Tables of the batched kernels.  assemble_K_batch_kernel (kernels_basic.cu) walks whole designs with one CTA: two
// [diagonal | q] regions and the rows' incident-edge lists (<= 8 per row) live in shared memory, a thread's rows (<= 4)
// and stored entries (<= 16) in registers; asm_smem_bytes <= 96 KB keeps two CTAs per SM.  assemble_rhs_batch_kernel
// keeps a thread's <= 6 rhs elements in registers (the same Nf <= 1024).  spmm_batch_kernel keeps a row's columns
// (<= 12) in registers.
void fdm_check_assemble_batch(fdm_plan& p) {
  p.asm_code.assign((size_t)p.nnz, -1);
  for (int64_t r = 0; r < p.Nf; ++r)
    for (int32_t k = p.index_indptr[r]; k < p.index_indptr[r + 1]; ++k)
      p.asm_code[(size_t)k] = p.entry_edge[(size_t)k] >= 0 ? p.entry_edge[(size_t)k] : -(int32_t)r - 2;
  p.rsup_ptr.assign((size_t)p.Nf + 1, 0);
  p.rsup_rows.clear();
  p.rsup_edge.clear();
  p.rsup_slot.clear();
  bool rows_ok = true;
  p.asm_maxdeg = 0;
  for (int64_t i = 0; i < p.Nf; ++i) {
    const int32_t node = p.free_node[i];
    for (int32_t k = p.ninc_ptr[node]; k < p.ninc_ptr[node + 1]; ++k) {
      const int32_t slot = p.node_slot[p.ninc_other[k]];
      if (slot < 0) {
        p.rsup_edge.push_back(p.ninc_edge[k] >> 1);
        p.rsup_slot.push_back(-slot - 1);
      }
    }
    p.rsup_ptr[(size_t)i + 1] = (int32_t)p.rsup_edge.size();
    if (p.rsup_ptr[(size_t)i + 1] > p.rsup_ptr[(size_t)i]) p.rsup_rows.push_back((int32_t)i);
    if (p.diags_indptr[i + 1] - p.diags_indptr[i] > 8) rows_ok = false;
    p.asm_maxdeg = std::max(p.asm_maxdeg, (int)(p.diags_indptr[i + 1] - p.diags_indptr[i]));
  }
  p.rhs_src.assign((size_t)p.Nf * 3, 0);
  p.asm_sup_in_K = p.rsup_rows.size() <= 256;
  for (int64_t i = 0; i < p.Nf; ++i) {
    for (int c = 0; c < 3; ++c) p.rhs_src[(size_t)i * 3 + c] = 3 * p.free_node[i] + c;
    if (p.rsup_ptr[(size_t)i + 1] - p.rsup_ptr[(size_t)i] > 2) p.asm_sup_in_K = false;
  }
  p.asm_smem_bytes = 16 * (((p.Nf + 1) & ~int64_t(1)) + ((p.E + 2) & ~int64_t(1))) + 32 * p.Nf + 16;
  p.asm_batch_ok = rows_ok && p.Nf <= 4 * 256 && p.asm_smem_bytes <= 96 * 1024 && p.nnz <= 16 * 256;
  p.spmm_batch_ok = true;
  for (int64_t i = 0; i < p.Nf; ++i)
    if (p.index_indptr[i + 1] - p.index_indptr[i] > 12) p.spmm_batch_ok = false;
}

void fdm_build_twisted(fdm_plan& p) {
  p.tw[0] = BandHalf();
  p.tw[1] = BandHalf();
  const int32_t n = (int32_t)p.Nf;
  const int32_t nblk = (n + 31) / 32;
  if (p.band_stage_ptr.empty() || p.hbw > 29 || nblk < 5) return;
  const int32_t nbT = (nblk - 1) / 2, nbB = nblk - 1 - nbT, m = nbT * 32, npad = nblk * 32;
  std::vector<int32_t> top((size_t)m + 32), bot((size_t)nbB * 32 + 32);
  for (int32_t R = 0; R < m + 32; ++R) top[(size_t)R] = R;                         // [T | M]; positions >= n pad
  for (int32_t R = 0; R < nbB * 32; ++R) bot[(size_t)R] = npad - 1 - R;            // B reversed
  for (int32_t R = 0; R < 32; ++R) bot[(size_t)(nbB * 32 + R)] = m + 31 - R;       // M reversed
  build_half_tables(p, top, -1, -1, p.tw[0]);
  build_half_tables(p, bot, nbB * 32, nbB * 32 + 32, p.tw[1]);
  p.tw[0].nblk = nbT + 1;
  p.tw[1].nblk = nbB;
}

// Partition of the free nodes into `num_parts` compact aggregates: farthest-point seeding on
// graph distance + Voronoi assignment (multi-source BFS).  Each new seed is the node farthest
// from all previous seeds; its BFS only walks where it improves the distance, so the whole
// construction costs a few sweeps of the graph.  Compact (ball-shaped) aggregates are what the
// piecewise-constant coarse space of the two-level preconditioner wants: on the 200x200
// cable-net they give ~260 PCG iterations against ~380 for BFS-order recursive bisection and
// ~1250 for plain Jacobi.  Ties are broken by node index, so the result is deterministic.
void fdm_build_partition(fdm_plan& p, int num_parts) {
  const int32_t n = (int32_t)p.Nf;
  num_parts = std::max(1, std::min(num_parts, n));
  p.part_of_row.assign(n, 0);
  p.num_parts = num_parts;
  if (num_parts == 1) return;
  Adj g(p);
  const int32_t INF = 0x3fffffff;
  std::vector<int32_t> dist(n, INF), owner(n, -1), frontier, next;
  auto grow = [&](int32_t seed, int32_t id) {
    dist[seed] = 0;
    owner[seed] = id;
    frontier.assign(1, seed);
    int32_t lvl = 0;
    while (!frontier.empty()) {
      ++lvl;
      next.clear();
      for (int32_t v : frontier)
        g.for_each(v, [&](int32_t w) {
          if (dist[w] > lvl) {
            dist[w] = lvl;
            owner[w] = id;
            next.push_back(w);
          }
        });
      frontier.swap(next);
    }
  };
  int32_t seed = 0;
  for (int32_t k = 0; k < num_parts; ++k) {
    grow(seed, k);
    // farthest node from the current seed set (unreached components count as infinitely far)
    int32_t best = -1, bestd = -1;
    for (int32_t v = 0; v < n; ++v)
      if (dist[v] > bestd) { bestd = dist[v]; best = v; }
    seed = best;
    if (bestd == 0) {  // fewer distinct nodes than parts (cannot happen: num_parts <= n)
      p.num_parts = k + 1;
      break;
    }
  }
  // More connected components than parts: the components no seed reached are handed, whole, to the
  // part that is smallest at that moment (an aggregate may then span several components; the coarse
  // space stays piecewise constant and every row has an owner).
  {
    std::vector<int64_t> size(p.num_parts, 0);
    for (int32_t v = 0; v < n; ++v)
      if (owner[v] >= 0) ++size[owner[v]];
    for (int32_t v = 0; v < n; ++v) {
      if (owner[v] >= 0) continue;
      const int32_t id = (int32_t)(std::min_element(size.begin(), size.end()) - size.begin());
      owner[v] = id;
      frontier.assign(1, v);
      int64_t cnt = 1;
      while (!frontier.empty()) {
        next.clear();
        for (int32_t u : frontier)
          g.for_each(u, [&](int32_t w) {
            if (owner[w] < 0) {
              owner[w] = id;
              ++cnt;
              next.push_back(w);
            }
          });
        frontier.swap(next);
      }
      size[id] += cnt;
    }
  }
  // compact ids (every seed owns at least itself, so all ids 0..num_parts-1 are present)
  p.part_of_row = owner;
}
