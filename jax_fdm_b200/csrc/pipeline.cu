// fdm_fa: value-and-gradient of a loss over a batch of designs as ONE native object -- what
// `jit(value_and_grad(loss_fn))(params)` is to the reference (optimization/optimizers/optimizer.py:365-370)
// and `vmap(...)(q)` over designs to visualization/plotters/loss_plotter.py:98-103.
//
// The object owns every device buffer, the chunk streams, the events and one CUDA graph per chunk, so that a
// step costs the host a handful of driver calls: the batch is cut into contiguous chunks; chunk c copies its
// q in (one cudaMemcpyAsync), replays its graph (solve from q -> positions -> state -> loss and cotangents ->
// adjoint solve with the stored factor -> gradients) and copies loss and q_bar out (one cudaMemcpyAsync
// each), all on its own stream: the H2D copy of one chunk, the kernels of another and the D2H copy of a third
// overlap on the copy engines and the SMs; inside the slice graphs the factor kernels carry the LOW launch priority
// (prioritise) and positions + state run on a side stream (launch_chunk), so that one slice's bandwidth-bound kernels
// run beside the next slice's issue-bound factorisation.  Two losses: the benchmark's quadratic distance to x0
// (fdm_loss_sqdist) and a compiled goal program (fdm_goals_loss + fdm_state_vjp).
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "launch.h"

extern "C" size_t fdm_workspace_bytes(const fdm_plan* p, int64_t batch, int32_t solver);

namespace {

struct Chunk {
  int64_t off = 0, B = 0;
  bool lowlat = false;     // a small slice at either end of a tapered batch (latency matters more than throughput)
  cudaStream_t st = nullptr;
  cudaStream_t side = nullptr;                          // positions + state beside the loss and the adjoint solve
  cudaEvent_t done = nullptr, ev_fork = nullptr, ev_join = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphExec_t exec_park[2] = {nullptr, nullptr};   // the same + the parking of the losses in staging row 0 / 1
  void* ws = nullptr;
  size_t ws_bytes = 0;
};

struct Buf {
  double* p = nullptr;
  int64_t per = 0;   // doubles per design (0: shared by the batch)
  double* at(int64_t b) const { return p ? p + b * per : nullptr; }
};

}  // namespace

__global__ void park_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

struct fdm_fa {
  const fdm_plan* plan = nullptr;
  int device = 0;          // kept here: the plan may be gone by the time the object is destroyed
  int64_t B = 0;
  fdm_fa_opts o{};
  bool fused = false, capturable = true, has_prog = false;
  int solver = 0;
  fdm_goal_program prog{};
  std::vector<void*> prog_mem;
  // inputs
  Buf q, xs, loads, x0;
  // forward
  Buf K, rhs, x, xyz, vectors, lengths, forces, residuals, info_f;
  // reverse
  Buf g, lam, loss, q_bar, loads_bar, xs_bar, info_a, xyz_cot, len_cot, for_cot, res_cot, xyz_bar;
  std::vector<void*> allocs;
  std::vector<Chunk> chunks;
  cudaEvent_t fork = nullptr;
  cudaEvent_t loss_ev[2] = {nullptr, nullptr};   // recorded after a host step's loss copy (copy_losses), by staging row
  bool loss_ev_armed[2] = {false, false};
  int64_t host_steps = 0;
  int prog_arrays = 15;                  // state arrays the goal program differentiates into (fdm_goal_kind_arrays)
  Buf loss_stage;                        // (2, B) staging rows of the losses of host steps
  cudaStream_t own = nullptr;      // run_host's ordering stream
  cudaGraphExec_t step_exec = nullptr;   // the whole device-resident step (fork / chunks / join) as one graph
  int64_t launches_per_step = 0;
};

namespace {

int alloc(fdm_fa* f, Buf& b, int64_t per, int64_t count, bool zero = true) {
  b.per = per;
  const size_t bytes = (size_t)std::max<int64_t>(per * count, 2) * sizeof(double);
  FDM_CUDA(cudaMalloc((void**)&b.p, bytes));
  f->allocs.push_back(b.p);
  if (zero) FDM_CUDA(cudaMemset(b.p, 0, bytes));
  return FDM_OK;
}

// the kernel sequence of one forward + adjoint pass over designs [c.off, c.off + c.B) on stream st
int launch_chunk(fdm_fa* f, const Chunk& c, cudaStream_t st, int64_t* count) {
  const fdm_plan* p = f->plan;
  const int64_t o = c.off, B = c.B, E = p->E, N = p->N, Nf = p->Nf, Ns = p->Ns;
  int rc;
  int64_t n = 0;
  fdm_solve_opts so;
  std::memset(&so, 0, sizeof(so));
  so.solver = f->solver;
  so.max_iters = f->o.max_iters;
  so.tol = f->o.tol;
  so.nan_on_failure = 1;
  so.concurrent_designs = (int32_t)std::min<int64_t>(f->B, 1 << 30);   // the slices run side by side
  // a tapered slice (fdm_fa_create) runs while the machine is filling up or draining: it is factored two-sided (half
  // the dependency chain) whatever the size of the whole batch
  if (c.lowlat) so.concurrent_designs = (int32_t)c.B;
  if (f->fused) {
    rc = fdm_solve_from_q(p, f->q.at(o), f->xs.p, f->loads.p, f->x.at(o), f->info_f.at(o), c.ws, c.ws_bytes, B, E, 0, 0,
                          &so, st);
    n += 4;                                   // factor, signed second pass, substitution, residual check of signed designs
  } else {
    rc = fdm_assemble(p, f->q.at(o), f->xs.p, f->loads.p, f->K.at(o), f->rhs.at(o), B, E, 0, 0, st);
    if (rc != FDM_OK) return rc;
    rc = fdm_solve(p, f->K.at(o), f->rhs.at(o), f->x.at(o), f->info_f.at(o), c.ws, c.ws_bytes, B, &so, st);
    n += 2 + 4;                               // assembly (K.data, rhs) + the band Cholesky's four (other solvers differ)
  }
  if (rc != FDM_OK) return rc;
  // The quadratic loss reads x_free only: positions and state leave the critical path (solve -> loss -> adjoint solve)
  // and run on the slice's side stream, joined again before the gradient kernel, which reads xyz.  For a small batch
  // every kernel of the chain is latency-bound and the step is their sum: 512 designs 0.331 -> 0.323 ms (0.314 with the
  // small substitution CTAs that start beside the state kernel), 4096 designs in 8 slices 1.478 -> 1.456 ms.
  static const bool no_fork = std::getenv("FDM_FA_FORK_STATE") && std::atoi(std::getenv("FDM_FA_FORK_STATE")) == 0;
  const bool fork = !f->has_prog && !no_fork && c.side;
  cudaStream_t ss = st;
  if (fork) {
    FDM_CUDA(cudaEventRecord(c.ev_fork, st));
    FDM_CUDA(cudaStreamWaitEvent(c.side, c.ev_fork, 0));
    ss = c.side;
  }
  rc = fdm_positions(p, f->x.at(o), f->xs.p, f->xyz.at(o), B, 0, 0, ss);
  if (rc != FDM_OK) return rc;
  ++n;
  // a goal program of node goals only reads the positions: the state is computed only when the caller wants it
  const bool prog_needs_state = f->has_prog && (f->prog_arrays & ~1) != 0;
  const bool have_state = f->o.with_state || prog_needs_state;
  if (have_state) {
    rc = fdm_state(p, f->q.at(o), f->xyz.at(o), f->loads.p, f->vectors.at(o), f->lengths.at(o), f->forces.at(o),
                   f->residuals.at(o), B, E, 0, ss);
    if (rc != FDM_OK) return rc;
    ++n;
  }
  if (fork) FDM_CUDA(cudaEventRecord(c.ev_join, c.side));
  so.reuse_factor = 1;
  if (!f->has_prog) {
    // L_b = 1/2 |x_free - x0|^2, g = x_free - x0
    rc = fdm_loss_sqdist(f->x.at(o), f->x0.p, f->g.at(o), f->loss.at(o), Nf * 3, B, 0, st);
    if (rc != FDM_OK) return rc;
    rc = fdm_solve(p, f->fused ? f->x.at(o) /* unused */ : f->K.at(o), f->g.at(o), f->lam.at(o), f->info_a.at(o), c.ws,
                   c.ws_bytes, B, &so, st);
    if (rc != FDM_OK) return rc;
    if (fork) FDM_CUDA(cudaStreamWaitEvent(st, c.ev_join, 0));
    rc = fdm_adjoint_grads(p, f->q.at(o), f->xyz.at(o), f->lam.at(o), f->q_bar.at(o), f->loads_bar.at(o), f->xs_bar.at(o),
                           B, E, 0, st);
    if (rc != FDM_OK) return rc;
    n += 3;
  } else {
    // goal program: loss and the cotangents of the state, their transposes through the state (direct terms of
    // q_bar, loads_bar and the cotangent of xyz), the split of xyz_bar into free / fixed rows, the adjoint solve
    // and the implicit terms accumulated on top
    // cotangent arrays of the state no goal of the program differentiates into are left out altogether
    double* xc = (f->prog_arrays & 1) ? f->xyz_cot.at(o) : nullptr;
    double* lc = (f->prog_arrays & 2) ? f->len_cot.at(o) : nullptr;
    double* fc = (f->prog_arrays & 4) ? f->for_cot.at(o) : nullptr;
    double* rc_ = (f->prog_arrays & 8) ? f->res_cot.at(o) : nullptr;
    rc = fdm_goals_loss(p, &f->prog, f->xyz.at(o), prog_needs_state ? f->lengths.at(o) : nullptr,
                        prog_needs_state ? f->forces.at(o) : nullptr, prog_needs_state ? f->residuals.at(o) : nullptr,
                        f->loss.at(o), nullptr, nullptr, xc, lc, fc, rc_, B, st);
    if (rc != FDM_OK) return rc;
    double* xyz_bar = f->xyz_bar.at(o);
    if (prog_needs_state) {
      rc = fdm_state_vjp(p, f->q.at(o), f->xyz.at(o), xc, rc_, lc, fc, nullptr, nullptr, f->q_bar.at(o), xyz_bar,
                         f->loads_bar.at(o), B, E, st);
      if (rc != FDM_OK) return rc;
    } else {
      // only xyz was read: its cotangent IS xyz_bar and there are no direct terms on q and the loads
      xyz_bar = xc;
      FDM_CUDA(cudaMemsetAsync(f->q_bar.at(o), 0, (size_t)B * E * 8, st));
      FDM_CUDA(cudaMemsetAsync(f->loads_bar.at(o), 0, (size_t)B * N * 3 * 8, st));
    }
    rc = fdm_positions(p, f->g.at(o), f->xs_bar.at(o), xyz_bar, B, Ns * 3, 1, st);
    if (rc != FDM_OK) return rc;
    rc = fdm_solve(p, f->fused ? f->x.at(o) : f->K.at(o), f->g.at(o), f->lam.at(o), f->info_a.at(o), c.ws, c.ws_bytes, B,
                   &so, st);
    if (rc != FDM_OK) return rc;
    rc = fdm_adjoint_grads(p, f->q.at(o), f->xyz.at(o), f->lam.at(o), f->q_bar.at(o), f->loads_bar.at(o), f->xs_bar.at(o),
                           B, E, 1, st);
    if (rc != FDM_OK) return rc;
    n += prog_needs_state ? 5 : 4;
  }
  (void)N;
  if (count) *count += n;
  return FDM_OK;
}

// Launch priorities of a captured slice (or step) graph: the factorisation kernels LOW, every other kernel HIGH.
// Slices run side by side on their own streams; a factor CTA keeps its registers for ~0.3 ms while HBM idles, and
// without priorities the queued factor CTAs of the next slice take every freed slot before the bandwidth-bound
// substitution / state / gradient kernels of the slice that has just been factored get one.  With them those kernels
// start as soon as slots free up and stream while the other slices factor.  FDM_FA_PRIORITY=0: off (A/B switch).
int prioritise(cudaGraph_t graph) {
  static const bool off = std::getenv("FDM_FA_PRIORITY") && std::atoi(std::getenv("FDM_FA_PRIORITY")) == 0;
  if (off) return FDM_OK;
  int least = 0, greatest = 0;
  FDM_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  if (least == greatest) return FDM_OK;
  size_t n = 0;
  FDM_CUDA(cudaGraphGetNodes(graph, nullptr, &n));
  std::vector<cudaGraphNode_t> nodes(n);
  if (n) FDM_CUDA(cudaGraphGetNodes(graph, nodes.data(), &n));
  for (cudaGraphNode_t node : nodes) {
    cudaGraphNodeType type;
    FDM_CUDA(cudaGraphNodeGetType(node, &type));
    if (type != cudaGraphNodeTypeKernel) continue;
    cudaKernelNodeParams kp;
    FDM_CUDA(cudaGraphKernelNodeGetParams(node, &kp));
    cudaKernelNodeAttrValue v;
    std::memset(&v, 0, sizeof(v));
    v.priority = chol_warp_is_factor_kernel(kp.func) ? least : greatest;
    FDM_CUDA(cudaGraphKernelNodeSetAttribute(node, cudaKernelNodeAttributePriority, &v));
  }
  return FDM_OK;
}

int run_chunk(fdm_fa* f, Chunk& c, cudaStream_t st) {
  if (c.exec) {
    FDM_CUDA(cudaGraphLaunch(c.exec, st));
    return FDM_OK;
  }
  return launch_chunk(f, c, st, nullptr);
}

int capture_chunks(fdm_fa* f) {
  f->launches_per_step = 0;
  for (Chunk& c : f->chunks) {
    int64_t n = 0;
    int rc = launch_chunk(f, c, c.st, &n);                // warm-up outside capture (also sets function attributes)
    if (rc != FDM_OK) return rc;
    f->launches_per_step += n;
    FDM_CUDA(cudaStreamSynchronize(c.st));
    if (!f->capturable || !f->o.use_graph) continue;
    if (c.exec) { cudaGraphExecDestroy(c.exec); c.exec = nullptr; }
    cudaGraph_t graph = nullptr;
    FDM_CUDA(cudaStreamBeginCapture(c.st, cudaStreamCaptureModeThreadLocal));
    rc = launch_chunk(f, c, c.st, nullptr);
    cudaError_t e = cudaStreamEndCapture(c.st, &graph);
    if (rc != FDM_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) { fdm_set_error(std::string("fdm_fa: graph capture: ") + cudaGetErrorString(e)); return FDM_ERR_CUDA; }
    if (prioritise(graph) != FDM_OK) { cudaGraphDestroy(graph); return FDM_ERR_CUDA; }
    e = cudaGraphInstantiate(&c.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { fdm_set_error(std::string("fdm_fa: graph instantiate: ") + cudaGetErrorString(e)); return FDM_ERR_CUDA; }
    // the same slice followed by the parking of its losses in staging row 0 / 1 (asynchronous host steps, see
    // copy_losses): inside the graph the parking kernel is free, launched after the graph it costs 18 us per slice
    for (int par = 0; par < 2; ++par) {
      if (c.exec_park[par]) { cudaGraphExecDestroy(c.exec_park[par]); c.exec_park[par] = nullptr; }
      graph = nullptr;
      FDM_CUDA(cudaStreamBeginCapture(c.st, cudaStreamCaptureModeThreadLocal));
      rc = launch_chunk(f, c, c.st, nullptr);
      park_kernel<<<(unsigned)((c.B + 255) / 256), 256, 0, c.st>>>(f->loss.at(c.off),
                                                                   f->loss_stage.p + (int64_t)par * f->B + c.off, c.B);
      e = cudaStreamEndCapture(c.st, &graph);
      if (rc != FDM_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (e != cudaSuccess) { fdm_set_error(std::string("fdm_fa: graph capture: ") + cudaGetErrorString(e)); return FDM_ERR_CUDA; }
      if (prioritise(graph) != FDM_OK) { cudaGraphDestroy(graph); return FDM_ERR_CUDA; }
      e = cudaGraphInstantiate(&c.exec_park[par], graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) { fdm_set_error(std::string("fdm_fa: graph instantiate: ") + cudaGetErrorString(e)); return FDM_ERR_CUDA; }
    }
  }
  // the device-resident step as ONE graph: fork from the capturing stream to the chunk streams and join back
  if (f->capturable && f->o.use_graph && f->chunks.size() > 1) {
    if (f->step_exec) { cudaGraphExecDestroy(f->step_exec); f->step_exec = nullptr; }
    cudaGraph_t graph = nullptr;
    FDM_CUDA(cudaStreamBeginCapture(f->own, cudaStreamCaptureModeThreadLocal));
    int rc = FDM_OK;
    cudaError_t e = cudaEventRecord(f->fork, f->own);
    for (Chunk& c : f->chunks) {
      if (e == cudaSuccess) e = cudaStreamWaitEvent(c.st, f->fork, 0);
      if (e == cudaSuccess && rc == FDM_OK) rc = launch_chunk(f, c, c.st, nullptr);
      if (e == cudaSuccess) e = cudaEventRecord(c.done, c.st);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(f->own, c.done, 0);
    }
    cudaError_t e2 = cudaStreamEndCapture(f->own, &graph);
    if (rc != FDM_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess || e2 != cudaSuccess) {
      fdm_set_error(std::string("fdm_fa: step graph capture: ") + cudaGetErrorString(e != cudaSuccess ? e : e2));
      if (graph) cudaGraphDestroy(graph);
      return FDM_ERR_CUDA;
    }
    if (prioritise(graph) != FDM_OK) { cudaGraphDestroy(graph); return FDM_ERR_CUDA; }
    e = cudaGraphInstantiate(&f->step_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { fdm_set_error(std::string("fdm_fa: step graph instantiate: ") + cudaGetErrorString(e)); return FDM_ERR_CUDA; }
  }
  return FDM_OK;
}

void destroy(fdm_fa* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  for (Chunk& c : f->chunks) {
    if (c.st) cudaStreamSynchronize(c.st);
    if (c.exec) cudaGraphExecDestroy(c.exec);
    for (cudaGraphExec_t g : c.exec_park)
      if (g) cudaGraphExecDestroy(g);
    if (c.done) cudaEventDestroy(c.done);
    if (c.ev_fork) cudaEventDestroy(c.ev_fork);
    if (c.ev_join) cudaEventDestroy(c.ev_join);
    if (c.side) cudaStreamDestroy(c.side);
    if (c.st) cudaStreamDestroy(c.st);
    if (c.ws) cudaFree(c.ws);
  }
  if (f->step_exec) cudaGraphExecDestroy(f->step_exec);
  if (f->fork) cudaEventDestroy(f->fork);
  for (cudaEvent_t e : f->loss_ev)
    if (e) cudaEventDestroy(e);
  if (f->own) cudaStreamDestroy(f->own);
  for (void* q : f->allocs) cudaFree(q);
  for (void* q : f->prog_mem) cudaFree(q);
  delete f;
}

template <class T>
int upload_arr(fdm_fa* f, const T* h, size_t count, const T** d) {
  void* m = nullptr;
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  FDM_CUDA(cudaMalloc(&m, bytes));
  f->prog_mem.push_back(m);
  if (count) FDM_CUDA(cudaMemcpy(m, h, count * sizeof(T), cudaMemcpyHostToDevice));
  *d = (const T*)m;
  return FDM_OK;
}

}  // namespace

extern "C" {

int fdm_fa_create(const fdm_plan* p, int64_t batch, const fdm_fa_opts* opts, fdm_fa** out) {
  if (!out) return FDM_ERR_INVALID;
  *out = nullptr;
  if (!p || !p->on_device || batch <= 0) {
    fdm_set_error("fdm_fa_create: plan is null or host-only, or batch <= 0");
    return FDM_ERR_INVALID;
  }
  FDM_CUDA(cudaSetDevice(p->device));
  fdm_fa* f = new (std::nothrow) fdm_fa();
  if (!f) return FDM_ERR_INVALID;
  f->plan = p;
  f->device = p->device;
  f->B = batch;
  if (opts) f->o = *opts;
  else { f->o.with_state = 1; f->o.use_graph = 1; }
  if (f->o.chunks <= 0) f->o.chunks = 1;
  if (f->o.tol <= 0) f->o.tol = 1e-12;
  const int64_t nch = std::min<int64_t>(f->o.chunks, batch);
  const int64_t E = p->E, N = p->N, Nf = p->Nf, Ns = p->Ns, nnz = p->nnz;
  // which solver serves this plan and chunk size (chunks differ by at most one design: same choice for all)
  const int64_t cb = batch / nch;
  fdm_solve_opts probe;
  std::memset(&probe, 0, sizeof(probe));
  probe.solver = f->o.solver;
  f->fused = chol_warp_from_q_supported(p) && cb > 1 &&
             (f->o.solver == FDM_SOLVER_AUTO || f->o.solver == FDM_SOLVER_CHOLESKY_BATCHED);
  f->solver = f->fused ? FDM_SOLVER_CHOLESKY_BATCHED : f->o.solver;
  // the multi-kernel PCG polls convergence from the host: not capturable
  {
    const bool chol = chol_warp_supported(p) || (chol_supported(p) && (cb > 1 || !pcg2_supported(p)));
    const int eff = f->solver != FDM_SOLVER_AUTO ? f->solver
                    : chol ? FDM_SOLVER_CHOLESKY_BATCHED : pcg2_supported(p) ? FDM_SOLVER_PCG_PERSISTENT : FDM_SOLVER_PCG;
    f->capturable = eff != FDM_SOLVER_PCG;
  }
  int rc = FDM_OK;
  auto A = [&](Buf& b, int64_t per, int64_t count) { if (rc == FDM_OK) rc = alloc(f, b, per, count); };
  A(f->q, E, batch); A(f->xs, Ns * 3, 1); A(f->loads, N * 3, 1); A(f->x0, Nf * 3, 1);
  if (!f->fused) { A(f->K, nnz, batch); A(f->rhs, Nf * 3, batch); }
  A(f->x, Nf * 3, batch); A(f->xyz, N * 3, batch);
  A(f->info_f, FDM_INFO_STRIDE, batch); A(f->info_a, FDM_INFO_STRIDE, batch);
  A(f->g, Nf * 3, batch); A(f->lam, Nf * 3, batch); A(f->loss, 1, batch); A(f->loss_stage, 1, 2 * batch);
  A(f->q_bar, E, batch); A(f->loads_bar, N * 3, batch); A(f->xs_bar, Ns * 3, batch);
  if (f->o.with_state) {
    A(f->vectors, E * 3, batch); A(f->lengths, E, batch); A(f->forces, E, batch); A(f->residuals, N * 3, batch);
  }
  if (rc == FDM_OK && cudaStreamCreateWithFlags(&f->own, cudaStreamNonBlocking) != cudaSuccess) rc = FDM_ERR_CUDA;
  if (rc == FDM_OK && cudaEventCreateWithFlags(&f->fork, cudaEventDisableTiming) != cudaSuccess) rc = FDM_ERR_CUDA;
  for (int i = 0; i < 2; ++i)
    if (rc == FDM_OK && cudaEventCreateWithFlags(&f->loss_ev[i], cudaEventDisableTiming) != cudaSuccess) rc = FDM_ERR_CUDA;
  // slice sizes: nch equal slices.  FDM_FA_TAPER = k (measurement only) halves the first and the last slice k times
  // ([64, 64, 128, 256, ..., 256, 128, 64, 64] for k = 2 on slices of 256) and factors those small slices two-sided:
  // measured SLOWER end to end (2.02 -> 2.06 ms at 16 slices) -- the host step is bound by the device's ramp-up, not
  // by the latency of its first and last slice, and small slices ramp up more slowly
  std::vector<int64_t> sizes;
  for (int64_t i = 0; i < nch; ++i) sizes.push_back(batch / nch + (i < batch % nch ? 1 : 0));
  {
    const char* e = std::getenv("FDM_FA_TAPER");
    const int taper = e ? std::atoi(e) : 0;
    const int64_t base = batch / nch;
    for (int k = 0; k < taper && nch >= 4; ++k) {
      const int64_t last = sizes.back(), first = sizes.front();
      if (last < 64 || first < 64) break;
      sizes.back() = last - last / 2;
      sizes.push_back(last / 2);
      sizes.front() = first - first / 2;
      sizes.insert(sizes.begin(), first / 2);
    }
    f->chunks.resize(sizes.size());
    for (size_t i = 0; i < sizes.size(); ++i) f->chunks[i].lowlat = sizes[i] < base;
  }
  int64_t off = 0;
  for (size_t i = 0; i < f->chunks.size() && rc == FDM_OK; ++i) {
    Chunk& c = f->chunks[i];
    c.off = off;
    c.B = sizes[i];
    off += c.B;
    if (cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming) != cudaSuccess) { rc = FDM_ERR_CUDA; break; }
    c.ws_bytes = std::max<size_t>(fdm_workspace_bytes(p, c.B, f->solver), 16);
    if (cudaMalloc(&c.ws, c.ws_bytes) != cudaSuccess) { rc = FDM_ERR_CUDA; break; }
  }
  if (rc != FDM_OK) {
    if (rc == FDM_ERR_CUDA) fdm_set_error(std::string("fdm_fa_create: ") + cudaGetErrorString(cudaGetLastError()));
    destroy(f);
    return rc;
  }
  *out = f;
  return FDM_OK;
}

void fdm_fa_destroy(fdm_fa* f) { destroy(f); }

int fdm_fa_set_constants(fdm_fa* f, const double* h_xyz_fixed, const double* h_loads, const double* h_x0) {
  if (!f || !h_xyz_fixed || !h_loads) {
    fdm_set_error("fdm_fa_set_constants: bad arguments");
    return FDM_ERR_INVALID;
  }
  const fdm_plan* p = f->plan;
  FDM_CUDA(cudaSetDevice(p->device));
  FDM_CUDA(cudaMemcpy(f->xs.p, h_xyz_fixed, (size_t)p->Ns * 3 * 8, cudaMemcpyHostToDevice));
  FDM_CUDA(cudaMemcpy(f->loads.p, h_loads, (size_t)p->N * 3 * 8, cudaMemcpyHostToDevice));
  if (h_x0) FDM_CUDA(cudaMemcpy(f->x0.p, h_x0, (size_t)p->Nf * 3 * 8, cudaMemcpyHostToDevice));
  return FDM_OK;
}

int fdm_fa_set_goal_program(fdm_fa* f, const fdm_goal_program* h) {
  if (!f || !h) {
    fdm_set_error("fdm_fa_set_goal_program: bad arguments");
    return FDM_ERR_INVALID;
  }
  const fdm_plan* p = f->plan;
  FDM_CUDA(cudaSetDevice(p->device));
  for (void* q : f->prog_mem) cudaFree(q);
  f->prog_mem.clear();
  fdm_goal_program d = *h;
  const size_t T = (size_t)h->num_terms, G = (size_t)h->num_groups;
  int rc = FDM_OK;
  if (rc == FDM_OK) rc = upload_arr(f, h->kind, T, &d.kind);
  if (rc == FDM_OK) rc = upload_arr(f, h->index, T, &d.index);
  if (rc == FDM_OK) rc = upload_arr(f, h->error, T, &d.error);
  if (rc == FDM_OK) rc = upload_arr(f, h->weight, T, &d.weight);
  if (rc == FDM_OK) rc = upload_arr(f, h->target, T * 6, &d.target);
  if (rc == FDM_OK) rc = upload_arr(f, h->group_ptr, G + 1, &d.group_ptr);
  if (rc == FDM_OK) rc = upload_arr(f, h->group_final, G, &d.group_final);
  if (rc == FDM_OK) rc = upload_arr(f, h->group_inner, G, &d.group_inner);
  if (rc == FDM_OK) rc = upload_arr(f, h->group_outer, G, &d.group_outer);
  if (rc == FDM_OK && h->agg_ptr) {
    rc = upload_arr(f, h->agg_ptr, T + 1, &d.agg_ptr);
    if (rc == FDM_OK) rc = upload_arr(f, h->agg_index, (size_t)h->agg_ptr[T], &d.agg_index);
  }
  if (rc != FDM_OK) return rc;
  f->prog = d;
  const int64_t B = f->B, E = p->E, N = p->N;
  auto A = [&](Buf& b, int64_t per) { if (rc == FDM_OK && !b.p) rc = alloc(f, b, per, B); };
  A(f->vectors, E * 3); A(f->lengths, E); A(f->forces, E); A(f->residuals, N * 3);
  A(f->xyz_cot, N * 3); A(f->len_cot, E); A(f->for_cot, E); A(f->res_cot, N * 3); A(f->xyz_bar, N * 3);
  if (rc != FDM_OK) return rc;
  f->has_prog = true;
  f->prog_arrays = 0;
  for (int32_t t = 0; t < h->num_terms; ++t) f->prog_arrays |= fdm_goal_kind_arrays(h->kind[t]);
  if (h->num_terms == 0) f->prog_arrays = 1;             // keep the launch well-formed: an all-zero xyz cotangent
  for (Chunk& c : f->chunks) {                                         // recaptured by the next run
    for (cudaGraphExec_t& g : c.exec_park)
      if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    if (c.exec) { cudaGraphExecDestroy(c.exec); c.exec = nullptr; }
  }
  if (f->step_exec) { cudaGraphExecDestroy(f->step_exec); f->step_exec = nullptr; }
  f->launches_per_step = 0;
  return FDM_OK;
}

int fdm_fa_set_q(fdm_fa* f, const double* h_q) {
  if (!f || !h_q) return FDM_ERR_INVALID;
  FDM_CUDA(cudaSetDevice(f->plan->device));
  FDM_CUDA(cudaMemcpy(f->q.p, h_q, (size_t)f->B * f->plan->E * 8, cudaMemcpyHostToDevice));
  return FDM_OK;
}

static int ensure_captured(fdm_fa* f) {
  if (f->launches_per_step) return FDM_OK;
  return capture_chunks(f);
}

int fdm_fa_step_device(fdm_fa* f, void* stream) {
  if (!f) return FDM_ERR_INVALID;
  int rc = ensure_captured(f);
  if (rc != FDM_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (f->step_exec) {
    FDM_CUDA(cudaGraphLaunch(f->step_exec, s));
    return FDM_OK;
  }
  if (f->chunks.size() == 1) return run_chunk(f, f->chunks[0], s);
  FDM_CUDA(cudaEventRecord(f->fork, s));
  for (Chunk& c : f->chunks) {
    FDM_CUDA(cudaStreamWaitEvent(c.st, f->fork, 0));
    rc = run_chunk(f, c, c.st);
    if (rc != FDM_OK) return rc;
    FDM_CUDA(cudaEventRecord(c.done, c.st));
    FDM_CUDA(cudaStreamWaitEvent(s, c.done, 0));
  }
  return FDM_OK;
}

// The per-design losses of ALL slices leave in one copy on the last slice's stream, once every slice's kernels are
// done: a 2 KB device-to-host copy per slice between the big ones costs the link 0.3 ms per step (16 slices, 61 MB each
// way: 1.51 -> 1.80 ms, tools/link_bandwidth.py).  Each slice parks its losses in a staging row (device-to-device, on
// its own stream) so that a step enqueued right behind this one cannot overwrite them before the copy has read them;
// two staging rows alternate, and a step waits for the copy that read its row two steps ago.  The parking costs the
// synchronous step 0.3 ms (2.04 -> 2.33 ms: a kernel launch between a graph launch and a copy on every stream), so
// fdm_fa_run_host, which waits before it returns, copies from the loss array itself (`park` = false).
static int copy_losses(fdm_fa* f, double* h_loss, int parity, bool park) {
  Chunk& last = f->chunks.back();
  for (Chunk& c : f->chunks)
    if (&c != &last) FDM_CUDA(cudaStreamWaitEvent(last.st, c.done, 0));
  FDM_CUDA(cudaMemcpyAsync(h_loss, park ? f->loss_stage.p + (int64_t)parity * f->B : f->loss.p, (size_t)f->B * 8, cudaMemcpyDeviceToHost, last.st));
  FDM_CUDA(cudaEventRecord(f->loss_ev[parity], last.st));
  f->loss_ev_armed[parity] = true;
  return FDM_OK;
}

static int enqueue_host(fdm_fa* f, const double* h_q, double* h_loss, double* h_q_bar, bool park) {
  if (!f || !h_q || !h_loss) {
    fdm_set_error("fdm_fa_enqueue_host: bad arguments");
    return FDM_ERR_INVALID;
  }
  int rc = ensure_captured(f);
  if (rc != FDM_OK) return rc;
  const int64_t E = f->plan->E;
  const int parity = (int)(f->host_steps++ & 1);
  for (Chunk& c : f->chunks) {
    FDM_CUDA(cudaMemcpyAsync(f->q.at(c.off), h_q + c.off * E, (size_t)c.B * E * 8, cudaMemcpyHostToDevice, c.st));
    if (park && f->loss_ev_armed[parity]) FDM_CUDA(cudaStreamWaitEvent(c.st, f->loss_ev[parity], 0));
    if (park && c.exec_park[parity]) {
      FDM_CUDA(cudaGraphLaunch(c.exec_park[parity], c.st));
      FDM_CUDA(cudaEventRecord(c.done, c.st));
      if (h_q_bar)
        FDM_CUDA(cudaMemcpyAsync(h_q_bar + c.off * E, f->q_bar.at(c.off), (size_t)c.B * E * 8, cudaMemcpyDeviceToHost, c.st));
      continue;
    }
    rc = run_chunk(f, c, c.st);
    if (rc != FDM_OK) return rc;
    // (a kernel, not a device-to-device memcpy: small memcpy operations of any kind between the big copies of a
    //  stream cost the link, see copy_losses)
    if (park) park_kernel<<<(unsigned)((c.B + 255) / 256), 256, 0, c.st>>>(f->loss.at(c.off),
                                                                 f->loss_stage.p + (int64_t)parity * f->B + c.off, c.B);
    FDM_CUDA(cudaEventRecord(c.done, c.st));
    if (h_q_bar)
      FDM_CUDA(cudaMemcpyAsync(h_q_bar + c.off * E, f->q_bar.at(c.off), (size_t)c.B * E * 8, cudaMemcpyDeviceToHost, c.st));
  }
  return copy_losses(f, h_loss, parity, park);
}

// asynchronous: the caller may enqueue the next step before this one has finished, so the losses are parked
int fdm_fa_enqueue_host(fdm_fa* f, const double* h_q, double* h_loss, double* h_q_bar) {
  return enqueue_host(f, h_q, h_loss, h_q_bar, true);
}

// measurement only: exactly the copies of fdm_fa_enqueue_host (same chunking, same streams), no kernels --
// what the host <-> device link can do for this call
int fdm_fa_enqueue_copies_only(fdm_fa* f, const double* h_q, double* h_loss, double* h_q_bar) {
  if (!f || !h_q || !h_loss) return FDM_ERR_INVALID;
  const int64_t E = f->plan->E;
  for (Chunk& c : f->chunks) {
    FDM_CUDA(cudaMemcpyAsync(f->q.at(c.off), h_q + c.off * E, (size_t)c.B * E * 8, cudaMemcpyHostToDevice, c.st));
    FDM_CUDA(cudaEventRecord(c.done, c.st));
    if (h_q_bar)
      FDM_CUDA(cudaMemcpyAsync(h_q_bar + c.off * E, f->q_bar.at(c.off), (size_t)c.B * E * 8, cudaMemcpyDeviceToHost, c.st));
  }
  return copy_losses(f, h_loss, 0, false);
}

int fdm_fa_wait(fdm_fa* f) {
  if (!f) return FDM_ERR_INVALID;
  for (Chunk& c : f->chunks) FDM_CUDA(cudaStreamSynchronize(c.st));
  return FDM_OK;
}

int fdm_fa_run_host(fdm_fa* f, const double* h_q, double* h_loss, double* h_q_bar) {
  int rc = enqueue_host(f, h_q, h_loss, h_q_bar, false);   // synchronous: nothing can overwrite the losses before the wait
  if (rc != FDM_OK) return rc;
  return fdm_fa_wait(f);
}

int fdm_fa_device_array(fdm_fa* f, int which, void** d_out, int64_t* per_design) {
  if (!f || !d_out) return FDM_ERR_INVALID;
  const Buf* b = nullptr;
  switch (which) {
    case FDM_FA_Q: b = &f->q; break;
    case FDM_FA_X_FREE: b = &f->x; break;
    case FDM_FA_XYZ: b = &f->xyz; break;
    case FDM_FA_LOSS: b = &f->loss; break;
    case FDM_FA_Q_BAR: b = &f->q_bar; break;
    case FDM_FA_LOADS_BAR: b = &f->loads_bar; break;
    case FDM_FA_XYZ_FIXED_BAR: b = &f->xs_bar; break;
    case FDM_FA_INFO_FORWARD: b = &f->info_f; break;
    case FDM_FA_INFO_ADJOINT: b = &f->info_a; break;
    case FDM_FA_LENGTHS: b = &f->lengths; break;
    case FDM_FA_FORCES: b = &f->forces; break;
    case FDM_FA_RESIDUALS: b = &f->residuals; break;
    case FDM_FA_VECTORS: b = &f->vectors; break;
    case FDM_FA_XYZ_FIXED: b = &f->xs; break;
    case FDM_FA_LOADS: b = &f->loads; break;
  }
  if (!b || !b->p) {
    fdm_set_error("fdm_fa_device_array: unknown or unallocated array");
    return FDM_ERR_INVALID;
  }
  *d_out = b->p;
  if (per_design) *per_design = b->per;
  return FDM_OK;
}

int64_t fdm_fa_info(const fdm_fa* f, int which) {
  if (!f) return FDM_ERR_INVALID;
  switch (which) {
    case FDM_FA_INFO_CHUNKS: return (int64_t)f->chunks.size();
    case FDM_FA_INFO_LAUNCHES: return f->launches_per_step;
    case FDM_FA_INFO_FUSED_ASSEMBLY: return f->fused ? 1 : 0;
    case FDM_FA_INFO_GRAPH: return (f->capturable && f->o.use_graph) ? 1 : 0;
    case FDM_FA_INFO_BATCH: return f->B;
  }
  return FDM_ERR_INVALID;
}

void* fdm_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    fdm_set_error(std::string("fdm_host_alloc: ") + cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}

void fdm_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
