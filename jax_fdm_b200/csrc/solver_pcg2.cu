// Persistent two-level PCG (placeholder until the cooperative kernel lands).
#include <cuda_runtime.h>

#include "launch.h"

size_t pcg2_workspace_bytes(const fdm_plan*) { return 0; }
bool pcg2_supported(const fdm_plan*) { return false; }
int pcg2_solve(const fdm_plan*, const double*, const double*, double*, double*, void*, size_t, int64_t,
               const fdm_solve_opts&, cudaStream_t) {
  fdm_set_error("pcg2_solve: not built");
  return FDM_ERR_UNSUPPORTED;
}
void pcg2_destroy(fdm_plan*) {}
