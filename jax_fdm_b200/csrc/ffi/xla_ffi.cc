// XLA FFI handlers over the C ABI of include/fdm_b200.h: what `jax.ffi.register_ffi_target`
// binds so that the kernels run as XLA custom calls inside jit / grad / vmap
// (jax_fdm_b200/jax_ffi.py).  Purely mechanical: every handler unpacks buffers, takes the CUDA
// stream XLA hands it and forwards to ONE fdm_* call -- no allocation, no synchronisation, no
// state kept between calls, errors returned as ffi::Error (never thrown).
//
// NOT part of libfdm_b200.so: it needs the XLA FFI headers (`jax.ffi.include_dir()`), which do
// not exist in this image (no jax / jaxlib).  Build where JAX is installed:
//   g++ -O2 -std=c++17 -fPIC -shared -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I include -I /usr/local/cuda/include jax_fdm_b200/csrc/ffi/xla_ffi.cc \
//       -L jax_fdm_b200/lib -lfdm_b200 -Wl,-rpath,'$ORIGIN' -o jax_fdm_b200/lib/libfdm_b200_xla.so
//
// The plan (static topology, built once on the host by fdm_plan_create) travels as an int64
// attribute holding the fdm_plan*; it is immutable during calls.  A leading batch axis on q /
// Kdata / rhs (jax.vmap with vmap_method="broadcast_all") maps onto the `batch` argument.
#include <cstdint>
#include <string>

#include <cuda_runtime_api.h>

#include "fdm_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using F64R = ffi::Result<ffi::Buffer<ffi::F64>>;
using U8R = ffi::Result<ffi::Buffer<ffi::U8>>;

namespace {

inline fdm_plan* plan_of(int64_t handle) { return reinterpret_cast<fdm_plan*>(static_cast<intptr_t>(handle)); }

inline ffi::Error status(int rc) {
  if (rc == FDM_OK) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string("fdm_b200: ") + fdm_last_error());
}

// batch = product of the leading dimensions beyond the per-design rank
template <class B>
inline int64_t batch_of(const B& buf, size_t base_rank) {
  auto d = buf.dimensions();
  int64_t b = 1;
  for (size_t i = 0; i + base_rank < d.size(); ++i) b *= d[i];
  return b;
}
inline int64_t stride_of(const F64& buf, size_t base_rank, int64_t per_design) {
  return buf.dimensions().size() > base_rank ? per_design : 0;
}

// ---- K1: (q, xyz_fixed, loads) -> (Kdata, rhs)          models.py:1064-1079, 828-856, 949-976
ffi::Error AssembleImpl(cudaStream_t stream, int64_t plan, F64 q, F64 xyz_fixed, F64 loads, F64R Kdata, F64R rhs) {
  fdm_plan* p = plan_of(plan);
  const int64_t E = fdm_plan_size(p, FDM_SIZE_EDGES), N = fdm_plan_size(p, FDM_SIZE_NODES),
                Ns = fdm_plan_size(p, FDM_SIZE_FIXED);
  return status(fdm_assemble(p, q.typed_data(), xyz_fixed.typed_data(), loads.typed_data(), Kdata->typed_data(),
                             rhs->typed_data(), batch_of(q, 1), E, stride_of(xyz_fixed, 2, Ns * 3),
                             stride_of(loads, 2, N * 3), stream));
}

// ---- K2: (Kdata, rhs) -> (x, info, workspace)           sparse.py:207-230, 296
// `workspace` is an extra uint8 result sized in Python with fdm_workspace_bytes: XLA owns it; the
// backward pass receives the forward pass's workspace as an operand to reuse the factor.
ffi::Error SolveImpl(cudaStream_t stream, int64_t plan, int64_t solver, int64_t max_iters, double tol,
                     int64_t reuse_factor, F64 Kdata, F64 rhs, F64R x, F64R info, U8R workspace) {
  fdm_plan* p = plan_of(plan);
  fdm_solve_opts o{};
  o.solver = (int32_t)solver;
  o.max_iters = (int32_t)max_iters;
  o.tol = tol;
  o.reuse_factor = (int32_t)reuse_factor;
  return status(fdm_solve(p, Kdata.typed_data(), rhs.typed_data(), x->typed_data(), info->typed_data(),
                          workspace->typed_data(), workspace->size_bytes(), batch_of(rhs, 2), &o, stream));
}

// Adjoint / repeated solve with the factor (or coarse inverse) a previous fdm_b200_solve left in `workspace_in`.
// XLA aliases operand 2 onto result 2 (input_output_aliases={2: 2} in jax_ffi.py), so both name one buffer;
// without the alias they differ and the factor is copied over first (device-to-device, on the stream).
ffi::Error SolveReuseImpl(cudaStream_t stream, int64_t plan, int64_t solver, int64_t max_iters, double tol, F64 Kdata,
                          F64 rhs, ffi::Buffer<ffi::U8> workspace_in, F64R x, F64R info, U8R workspace) {
  fdm_plan* p = plan_of(plan);
  if (workspace_in.size_bytes() != workspace->size_bytes())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "fdm_b200_solve_reuse: workspace operand and result differ in size");
  if (workspace_in.typed_data() != workspace->typed_data() &&
      cudaMemcpyAsync(workspace->typed_data(), workspace_in.typed_data(), workspace->size_bytes(),
                      cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "fdm_b200_solve_reuse: workspace copy failed");
  fdm_solve_opts o{};
  o.solver = (int32_t)solver;
  o.max_iters = (int32_t)max_iters;
  o.tol = tol;
  o.reuse_factor = 1;
  return status(fdm_solve(p, Kdata.typed_data(), rhs.typed_data(), x->typed_data(), info->typed_data(),
                          workspace->typed_data(), workspace->size_bytes(), batch_of(rhs, 2), &o, stream));
}

// ---- K1 + K2 fused: (q, xyz_fixed, loads) -> (x_free, info, workspace)     models.py:222-256
ffi::Error SolveFromQImpl(cudaStream_t stream, int64_t plan, int64_t solver, F64 q, F64 xyz_fixed, F64 loads, F64R x,
                          F64R info, U8R workspace) {
  fdm_plan* p = plan_of(plan);
  const int64_t E = fdm_plan_size(p, FDM_SIZE_EDGES), N = fdm_plan_size(p, FDM_SIZE_NODES),
                Ns = fdm_plan_size(p, FDM_SIZE_FIXED);
  fdm_solve_opts o{};
  o.solver = (int32_t)solver;
  return status(fdm_solve_from_q(p, q.typed_data(), xyz_fixed.typed_data(), loads.typed_data(), x->typed_data(),
                                 info->typed_data(), workspace->typed_data(), workspace->size_bytes(), batch_of(q, 1), E,
                                 stride_of(xyz_fixed, 2, Ns * 3), stride_of(loads, 2, N * 3), &o, stream));
}

// ---- tributary edge loads and their VJP (tmax > 1)      models.py:290-344, loads.py:296-331
ffi::Error EdgeLoadsImpl(cudaStream_t stream, int64_t plan, int64_t is_local, F64 xyz, F64 loads_nodes, F64 edges_load,
                         F64R loads_out) {
  fdm_plan* p = plan_of(plan);
  const int64_t E = fdm_plan_size(p, FDM_SIZE_EDGES), N = fdm_plan_size(p, FDM_SIZE_NODES);
  return status(fdm_edge_loads(p, xyz.typed_data(), loads_nodes.typed_data(), edges_load.typed_data(),
                               loads_out->typed_data(), batch_of(xyz, 2), stride_of(loads_nodes, 2, N * 3),
                               stride_of(edges_load, 2, E * 3), (int)is_local, stream));
}
ffi::Error EdgeLoadsVjpImpl(cudaStream_t stream, int64_t plan, int64_t is_local, F64 xyz, F64 edges_load, F64 cot,
                            F64R xyz_bar, F64R edges_load_bar) {
  fdm_plan* p = plan_of(plan);
  return status(fdm_edge_loads_vjp(p, xyz.typed_data(), edges_load.typed_data(), cot.typed_data(), xyz_bar->typed_data(),
                                   edges_load_bar->typed_data(), batch_of(xyz, 2),
                                   stride_of(edges_load, 2, fdm_plan_size(p, FDM_SIZE_EDGES) * 3), (int)is_local, stream));
}

// ---- benchmark loss: (x, x0) -> (g, loss)
ffi::Error LossSqdistImpl(cudaStream_t stream, int64_t plan, F64 x, F64 x0, F64R g, F64R loss) {
  (void)plan;
  const int64_t B = batch_of(x, 2);
  const int64_t per = (int64_t)x.element_count() / (B > 0 ? B : 1);
  return status(fdm_loss_sqdist(x.typed_data(), x0.typed_data(), g->typed_data(), loss->typed_data(), per, B,
                                stride_of(x0, 2, per), stream));
}

// ---- SpMM: (Kdata, X) -> Y                               models.py:1012, sparse.py:301
ffi::Error SpmmImpl(cudaStream_t stream, int64_t plan, F64 Kdata, F64 X, F64R Y) {
  return status(fdm_spmm(plan_of(plan), Kdata.typed_data(), X.typed_data(), Y->typed_data(), batch_of(X, 2), stream));
}

// ---- positions: (x_free, xyz_fixed) -> xyz               models.py:195-220
ffi::Error PositionsImpl(cudaStream_t stream, int64_t plan, F64 x_free, F64 xyz_fixed, F64R xyz) {
  fdm_plan* p = plan_of(plan);
  return status(fdm_positions(p, x_free.typed_data(), xyz_fixed.typed_data(), xyz->typed_data(), batch_of(x_free, 2),
                              stride_of(xyz_fixed, 2, fdm_plan_size(p, FDM_SIZE_FIXED) * 3), 0, stream));
}

// ---- K4: (q, xyz, loads) -> (vectors, lengths, forces, residuals)   models.py:753-794
ffi::Error StateImpl(cudaStream_t stream, int64_t plan, F64 q, F64 xyz, F64 loads, F64R vectors, F64R lengths,
                     F64R forces, F64R residuals) {
  fdm_plan* p = plan_of(plan);
  const int64_t E = fdm_plan_size(p, FDM_SIZE_EDGES), N = fdm_plan_size(p, FDM_SIZE_NODES);
  return status(fdm_state(p, q.typed_data(), xyz.typed_data(), loads.typed_data(), vectors->typed_data(),
                          lengths->typed_data(), forces->typed_data(), residuals->typed_data(), batch_of(xyz, 2), E,
                          stride_of(loads, 2, N * 3), stream));
}

// ---- K4 transpose: cotangents of the state -> (q_bar, xyz_bar, loads_bar)
ffi::Error StateVjpImpl(cudaStream_t stream, int64_t plan, F64 q, F64 xyz, F64 xyz_cot, F64 residuals_cot,
                        F64 lengths_cot, F64 forces_cot, F64 loads_cot, F64 vectors_cot, F64R q_bar, F64R xyz_bar,
                        F64R loads_bar) {
  fdm_plan* p = plan_of(plan);
  return status(fdm_state_vjp(p, q.typed_data(), xyz.typed_data(), xyz_cot.typed_data(), residuals_cot.typed_data(),
                              lengths_cot.typed_data(), forces_cot.typed_data(), loads_cot.typed_data(),
                              vectors_cot.typed_data(), q_bar->typed_data(), xyz_bar->typed_data(),
                              loads_bar->typed_data(), batch_of(xyz, 2), fdm_plan_size(p, FDM_SIZE_EDGES), stream));
}

// ---- K3: (q, xyz, lam) -> (q_bar, loads_bar, xyz_fixed_bar)   sparse.py:299-308 + transposes of models.py:1069-1077, 856, 973-976
ffi::Error AdjointGradsImpl(cudaStream_t stream, int64_t plan, F64 q, F64 xyz, F64 lam, F64R q_bar, F64R loads_bar,
                            F64R xyz_fixed_bar) {
  fdm_plan* p = plan_of(plan);
  return status(fdm_adjoint_grads(p, q.typed_data(), xyz.typed_data(), lam.typed_data(), q_bar->typed_data(),
                                  loads_bar->typed_data(), xyz_fixed_bar->typed_data(), batch_of(xyz, 2),
                                  fdm_plan_size(p, FDM_SIZE_EDGES), /*accumulate=*/0, stream));
}

// ---- forward mode of the state: (q, xyz, q_dot, xyz_dot, loads_dot) -> tangents of the state   jax.jvp of models.py:780-794
ffi::Error StateJvpImpl(cudaStream_t stream, int64_t plan, F64 q, F64 xyz, F64 q_dot, F64 xyz_dot, F64 loads_dot,
                        F64R vectors_dot, F64R lengths_dot, F64R forces_dot, F64R residuals_dot) {
  fdm_plan* p = plan_of(plan);
  const int64_t E = fdm_plan_size(p, FDM_SIZE_EDGES), N = fdm_plan_size(p, FDM_SIZE_NODES);
  // T tangents may share one primal: a primal without the leading axis has stride 0
  return status(fdm_state_jvp(p, q.typed_data(), xyz.typed_data(), q_dot.typed_data(), xyz_dot.typed_data(),
                              loads_dot.typed_data(), vectors_dot->typed_data(), lengths_dot->typed_data(),
                              forces_dot->typed_data(), residuals_dot->typed_data(), batch_of(xyz_dot, 2),
                              stride_of(q, 1, E), stride_of(xyz, 2, N * 3), stride_of(q_dot, 1, E),
                              stride_of(loads_dot, 2, N * 3), stream));
}

// ---- Kbar on the pattern and its pull-back to q_bar      sparse.py:299-306, transpose of models.py:1069-1077
ffi::Error KbarPatternImpl(cudaStream_t stream, int64_t plan, double sign, F64 a, F64 b, F64R Kbar) {
  return status(fdm_kbar_pattern(plan_of(plan), a.typed_data(), b.typed_data(), Kbar->typed_data(), sign, batch_of(a, 2),
                                 stream));
}
ffi::Error StiffnessVjpImpl(cudaStream_t stream, int64_t plan, F64 Kbar, F64R q_bar) {
  return status(fdm_stiffness_vjp(plan_of(plan), Kbar.typed_data(), q_bar->typed_data(), batch_of(Kbar, 1), stream));
}

// ---- tributary face loads (mesh structures)              models.py:333-344, loads.py:35-288
using S32 = ffi::Buffer<ffi::S32>;
ffi::Error FaceLoadsImpl(cudaStream_t stream, int64_t plan, int64_t is_local, S32 faces, S32 edges_faces, F64 xyz,
                         F64 loads_nodes, F64 faces_load, F64R centroids, F64R gl, F64R loads_out) {
  fdm_plan* p = plan_of(plan);
  const int64_t N = fdm_plan_size(p, FDM_SIZE_NODES);
  const int64_t F = (int64_t)faces.dimensions()[0], D = (int64_t)faces.dimensions()[1];
  return status(fdm_face_loads(p, faces.typed_data(), F, D, edges_faces.typed_data(), xyz.typed_data(),
                               loads_nodes.typed_data(), faces_load.typed_data(), centroids->typed_data(), gl->typed_data(),
                               loads_out->typed_data(), batch_of(xyz, 2), stride_of(loads_nodes, 2, N * 3),
                               stride_of(faces_load, 2, F * 3), (int)is_local, stream));
}

}  // namespace

#define FDM_STREAM .Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("plan")

XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_assemble, AssembleImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_solve, SolveImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Attr<int64_t>("solver").Attr<int64_t>("max_iters")
                                  .Attr<double>("tol").Attr<int64_t>("reuse_factor")
                                  .Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_solve_reuse, SolveReuseImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Attr<int64_t>("solver").Attr<int64_t>("max_iters")
                                  .Attr<double>("tol")
                                  .Arg<F64>().Arg<F64>().Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_solve_from_q, SolveFromQImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Attr<int64_t>("solver")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_edge_loads, EdgeLoadsImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Attr<int64_t>("is_local")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_edge_loads_vjp, EdgeLoadsVjpImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Attr<int64_t>("is_local")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_loss_sqdist, LossSqdistImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_spmm, SpmmImpl, ffi::Ffi::Bind() FDM_STREAM.Arg<F64>().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_positions, PositionsImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Arg<F64>().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_state, StateImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_state_vjp, StateVjpImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_adjoint_grads, AdjointGradsImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_state_jvp, StateJvpImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_kbar_pattern, KbarPatternImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Attr<double>("sign").Arg<F64>().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_stiffness_vjp, StiffnessVjpImpl, ffi::Ffi::Bind() FDM_STREAM.Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(fdm_b200_face_loads, FaceLoadsImpl,
                              ffi::Ffi::Bind() FDM_STREAM.Attr<int64_t>("is_local")
                                  .Arg<S32>().Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
