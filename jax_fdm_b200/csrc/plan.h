// Internal plan layout shared by the host builder (plan.cpp) and the kernel launchers.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/fdm_b200.h"

struct DeviceArrays {
  // topology, node order
  int32_t* node_slot = nullptr;    // (N)  >=0: free index, <0: -(fixed index)-1
  int32_t* edge_nodes = nullptr;   // (E,2) tail, head
  int32_t* ninc_ptr = nullptr;     // (N+1) incidence lists of every node
  int32_t* ninc_edge = nullptr;    // (2E)  (edge << 1) | is_head, ascending edge id per node
  int32_t* ninc_other = nullptr;   // (2E)  other endpoint (node id)
  int32_t* free_node = nullptr;    // (Nf)  node id of each free index   (= indices_free)
  int32_t* fixed_node = nullptr;   // (Ns)  node id of each fixed index  (= indices_fixed)
  // CSR/CSC pattern of K in the reference's order (symmetric pattern)
  int32_t* rowptr = nullptr;       // (Nf+1) = index_array.indptr
  int32_t* colidx = nullptr;       // (nnz)  = index_array.indices
  int32_t* entry_edge = nullptr;   // (nnz)  = index_array.data - 1   (-1 on the diagonal)
  int32_t* diag_pos = nullptr;     // (Nf)   = diag_indices
  int32_t* diags_ptr = nullptr;    // (Nf+1) = diags.indptr  (per free row: incident edges, ascending id)
  int32_t* diags_idx = nullptr;    // (dnnz) = diags.indices
  uint32_t* sup_mask = nullptr;    // (ceil(Nf/32)) bit r set: free row r has a support neighbour
  int32_t* asm_code = nullptr;     // (nnz)  entry -> edge id, or -(row) - 2 on the diagonal      (assemble_design_kernel)
  int32_t* rsup_ptr = nullptr;     // (Nf+1) per free row: its edges to supports, in incidence order
  int32_t* rsup_rows = nullptr;    // the free rows that have at least one such edge
  int32_t* rhs_src = nullptr;      // (3 Nf) flat rhs element -> 3 * node + component (index into a design's loads)
  int32_t* rsup_edge = nullptr;    // (rsup) edge id
  int32_t* rsup_slot = nullptr;    // (rsup) support slot (row of xyz_fixed)
  // band ordering (batched Cholesky)
  int32_t* band_perm = nullptr;    // (Nf) free index -> band position
  int32_t* band_iperm = nullptr;   // (Nf) band position -> free index
  int32_t* band_kidx = nullptr;    // (Nf, hbw+1) column-oriented: [j*(hbw+1)+d] = position in K.data of A(j+d, j), -1 = zero
  int32_t* band_rowkidx = nullptr; // ((ceil(Nf/32)+1)*32, 32) row-oriented: [R*32+t] = position of A(R, R-t), -1 zero, -2 one
  // the same two tables without their zeros: per 32-row block, groups of 32 (position, index) pairs
  int32_t* band_stage_ptr = nullptr;   // (ceil(Nf/32)+2) first group of every block
  int32_t* band_stage_kent = nullptr;  // (groups*32, 2): r*32 + slot, position in K.data | -2 one | -4 padding
  // per band row (padded rows), for the factorisation that assembles its rows from q (fdm_solve_from_q):
  int32_t* band_row_diag = nullptr;    // (rows, 8)  incident edge ids in ascending order, -1 padded; [0] = -2: list too long, walk `diags`
  int32_t* band_row_sup = nullptr;     // (rows, 8)  up to 4 (edge id, fixed index) pairs in incidence order, -1 padded; [0] = -2: walk the incidence list
  int32_t* band_row_node = nullptr;    // (rows)     node id of the row, -1 for padding rows
  int32_t* band_stage_eent = nullptr;  // (groups*32, 2): r*32 + slot, edge id | -1 zero (diagonal) | -2 one | -4 padding
  // two-sided ("twisted") factorisation: the same tables for the top half [T | M] and the bottom half
  // [reversed B | reversed M] of the band order (see BandHalf)
  int32_t* tw_iperm[2] = {nullptr, nullptr};
  int32_t* tw_stage_ptr[2] = {nullptr, nullptr};
  int32_t* tw_stage_kent[2] = {nullptr, nullptr};
  int32_t* tw_stage_eent[2] = {nullptr, nullptr};
  int32_t* tw_row_diag[2] = {nullptr, nullptr};
  int32_t* tw_row_sup[2] = {nullptr, nullptr};
  int32_t* tw_row_node[2] = {nullptr, nullptr};
  // two-level persistent PCG (solver_pcg_persistent.cu); built lazily by the host
  void* pcg2 = nullptr;
};

// One half of the two-sided band factorisation.  The band order is cut into T = blocks [0, nbT), a middle block M
// (32 rows, >= hbw, so T and B do not couple) and B = the blocks after M.  Half 0 factors [T | M] top-down; half 1
// factors B bottom-up, i.e. the order [reversed B | reversed M] top-down with the M x M entries and M's right-hand
// side left out (half 0 owns them).  Both leave their Schur complement on M; half 0 adds half 1's and factors M.
// Same flops and fill as the one-sided factorisation, half the dependency chain.  Tables as the one-sided ones,
// over the half's own row numbering; iperm = -1 for padding rows and, in half 1, for M's rows.
struct BandHalf {
  std::vector<int32_t> iperm, stage_ptr, stage_kent, stage_eent, row_diag, row_sup, row_node;
  int32_t nblk = 0;       // blocks this half eliminates (half 0: nbT + 1 including M; half 1: nbB)
};

struct fdm_plan {
  int64_t N = 0, E = 0, Nf = 0, Ns = 0, nnz = 0, dnnz = 0;
  int device = -1;
  bool on_device = false;
  int sm_count = 0;

  // reference-typed host arrays
  std::vector<int64_t> indices_free, indices_fixed, indices_freefixed, index_data, diag_indices;
  std::vector<int32_t> index_indices, index_indptr, diags_indptr, diags_indices;

  // kernel-side host arrays
  std::vector<int32_t> node_slot, edge_nodes, ninc_ptr, ninc_edge, ninc_other;
  std::vector<int32_t> free_node, fixed_node, entry_edge, diag_pos;
  std::vector<uint32_t> sup_mask;
  std::vector<int32_t> asm_code, rsup_ptr, rsup_rows, rsup_edge, rsup_slot, rhs_src;   // batched assembly (fdm_check_assemble_batch)

  // band ordering
  std::vector<int32_t> band_perm, band_iperm, band_kidx, band_rowkidx;
  std::vector<int32_t> band_rowedge;   // host only: band_rowkidx in terms of edge ids (-3 diagonal), source of band_stage_eent
  std::vector<int32_t> band_stage_ptr, band_stage_kent, band_stage_eent;
  std::vector<int32_t> band_row_diag, band_row_sup, band_row_node;
  int32_t hbw = 0;
  bool spmm_batch_ok = false;  // every row has <= 12 stored entries (spmm_batch_kernel keeps a row's columns in registers)
  bool asm_batch_ok = false;   // a design's q, diagonals and index streams fit one CTA's shared memory (asm_smem_bytes)
  int64_t asm_smem_bytes = 0;
  int asm_maxdeg = 0;          // longest incident-edge list of a free row
  bool asm_sup_in_K = false;   // rows next to supports: <= 256 of them, <= 2 supports each (assemble_K_batch_kernel<., true>)
  BandHalf tw[2];         // empty when the two-sided factorisation does not apply (hbw > 29 or fewer than 5 blocks)

  // aggregates for the two-level preconditioner
  std::vector<int32_t> part_of_row;  // (Nf) free index -> aggregate
  int32_t num_parts = 0;

  DeviceArrays d;
};

// error text (thread-local), shared by all translation units
void fdm_set_error(const std::string& msg);

// host builders (plan.cpp)
int fdm_build_topology(fdm_plan& p, const int64_t* edges, int64_t E, int64_t N, const int64_t* supports);
void fdm_build_band_order(fdm_plan& p);
void fdm_build_partition(fdm_plan& p, int num_parts);
void fdm_build_twisted(fdm_plan& p);
void fdm_check_assemble_batch(fdm_plan& p);
