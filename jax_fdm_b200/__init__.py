"""
jax_fdm_b200 -- B200-native equilibrium hot path of JAX FDM (forward force-density solve
and its implicit-differentiation adjoint) behind the `jax_fdm.equilibrium` API surface.

    from jax_fdm_b200.equilibrium import EquilibriumStructureSparse, EquilibriumModelSparse, fdm

Hand-written sm_100a CUDA kernels behind a plain C ABI (include/fdm_b200.h); torch is used
for device memory, streams and torch.distributed only.
"""

__version__ = "0.1.0"
