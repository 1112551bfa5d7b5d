"""
CPU: the C-ABI library loads, exports every symbol include/fdm_b200.h declares, and its
native plan builder reproduces the reference's topology arrays bit-for-bit (golden fixtures
made by the reference's own code, tests/golden/make_golden.py) -- no compute calls.
"""

import ctypes as C
import os
import re

import numpy as np
import pytest

from jax_fdm_b200 import _lib as L
from jax_fdm_b200 import datasets
from jax_fdm_b200.equilibrium.structures import EquilibriumStructureSparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["arch10", "readme10", "cablenet12", "pillow30", "olympia", "stuttgart21", "liew_dome", "random40"]


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    header = open(os.path.join(ROOT, "include", "fdm_b200.h")).read()
    declared = set(re.findall(r"\b(fdm_[a-z_0-9]+)\s*\(", header))
    declared -= {"fdm_plan", "fdm_solve_opts"}
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.fdm_version()


@pytest.mark.parametrize("name", CASES)
def test_plan_arrays_bit_exact_vs_reference(name):
    g = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    s = EquilibriumStructureSparse(g["nodes"], g["edges"], g["supports"], device=None)
    assert np.array_equal(s.edges_indexed, g["ref_edges_indexed"])
    assert np.array_equal(s.indices_free, g["ref_indices_free"])
    assert np.array_equal(s.indices_fixed, g["ref_indices_fixed"])
    assert np.array_equal(s.indices_freefixed, g["ref_indices_freefixed"])
    ia = s.index_array
    assert np.array_equal(ia.data, g["ref_index_data"]) and ia.data.dtype == g["ref_index_data"].dtype
    assert np.array_equal(ia.indices, g["ref_index_indices"]) and ia.indices.dtype == g["ref_index_indices"].dtype
    assert np.array_equal(ia.indptr, g["ref_index_indptr"]) and ia.indptr.dtype == g["ref_index_indptr"].dtype
    assert np.array_equal(s.diag_indices, g["ref_diag_indices"])
    assert np.array_equal(s.diags_indices, g["ref_diags_indices"])
    assert np.array_equal(s.diags_indptr, g["ref_diags_indptr"])
    # orderings the reference does not have: permutation / partition sanity
    bp = s.band_perm
    assert np.array_equal(np.sort(bp), np.arange(s.num_free))
    rows = np.repeat(np.arange(s.num_free), np.diff(ia.indptr))
    assert np.abs(bp[rows] - bp[ia.indices]).max() == s.half_bandwidth
    part = s.part_of_row
    assert part.min() == 0 and part.max() == min(148, s.num_free) - 1
    assert np.bincount(part).min() >= 1


def test_plan_sizes_baseline_configs():
    # nnz verified values
    for prob, nf, nnz in [(datasets.cablenet_readme_problem(10), 117, 541),
                          (datasets.pillow_problem(30), 841, 4089),
                          (datasets.cablenet_problem(200), 40397, 201181)]:
        s = EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=None)
        assert (s.num_free, s.nnz) == (nf, nnz)
    assert s.half_bandwidth <= 203
    s30 = EquilibriumStructureSparse(*datasets.pillow_problem(30)[:3], device=None)
    assert s30.half_bandwidth == 29


def test_rejects_what_the_reference_cannot_represent():
    nodes = np.arange(4)
    sup = np.array([1, 0, 0, 1])
    with pytest.raises(L.FdmError) as e:       # parallel edges: index_array would sum e+1 (structures.py:248-288)
        EquilibriumStructureSparse(nodes, [[0, 1], [1, 2], [2, 1], [2, 3]], sup, device=None)
    assert e.value.code == -3
    with pytest.raises(L.FdmError) as e:       # self loop
        EquilibriumStructureSparse(nodes, [[0, 1], [1, 1], [2, 3]], sup, device=None)
    assert e.value.code == -3
    with pytest.raises(L.FdmError) as e:       # no supports (fdm.py:491-522)
        EquilibriumStructureSparse(nodes, [[0, 1], [1, 2], [2, 3]], np.zeros(4, int), device=None)
    assert e.value.code == -4
    with pytest.raises(KeyError):
        EquilibriumStructureSparse(nodes, [[0, 7]], sup, device=None)


def test_compute_on_host_plan_is_refused():
    s = EquilibriumStructureSparse(*datasets.arch_problem(10)[:3], device=None)
    rc = L.lib().fdm_spmm(s.plan, C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), 1, C.c_void_p(0))
    assert rc == -1 and b"host-only" in L.lib().fdm_last_error()


def test_partition_with_more_components_than_parts():
    """300 disjoint support-free-free-support cables: 300 components of the free-node graph against 148 aggregates.
    Every row must have an owner in [0, parts) (rows left at -1 used to index out of bounds in pcg2_build)."""
    k = 300
    nodes = np.arange(4 * k)
    edges = np.concatenate([np.array([[4 * c, 4 * c + 1], [4 * c + 1, 4 * c + 2], [4 * c + 2, 4 * c + 3]]) for c in range(k)])
    sup = np.tile(np.array([1, 0, 0, 1]), k)
    s = EquilibriumStructureSparse(nodes, edges, sup, device=None)
    part = s.part_of_row
    assert s.num_free == 2 * k and part.min() == 0 and part.max() == 147
    assert np.bincount(part, minlength=148).min() >= 1
    assert np.array_equal(part[0::2], part[1::2])        # a component is never split by the fallback
