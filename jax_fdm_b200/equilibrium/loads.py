"""
Mirror of `jax_fdm.equilibrium.loads` (loads.py:35-465): shape-dependent tributary loads.

The two entry points the model calls -- `nodes_load_from_edges`, `nodes_load_from_faces` (models.py:349-408) --
run in this library's kernels (fdm_edge_loads / fdm_face_loads, with their VJPs) on CUDA tensors.  The reference
builds them from per-element helpers that it `vmap`s (`edge_load_lcs`, `face_load_lcs`, `edge_tributary_faces_load`,
...); those are restated here as small batched tensor expressions with the same names, argument meaning and
results, for code that composes them itself.  They are NOT on the device path (`EquilibriumModel.nodes_load` never
calls them); they run wherever their tensors live and the CPU tests pin them to the oracle.
"""
import numpy as np
import torch

__all__ = [
    "calculate_edges_load", "calculate_faces_load", "edge_load_lcs", "edge_tributary_face_area",
    "edge_tributary_faces_load", "edges_tributary_edges_load", "edges_tributary_faces_load", "face_load_lcs", "face_xyz",
    "nodes_load_from_edges", "nodes_load_from_faces", "nodes_tributary_edges_load",
]


def _idx(a, like):
    return torch.as_tensor(np.asarray(a), dtype=torch.int64, device=like.device)


def _unit(v):
    return v / torch.linalg.norm(v, dim=-1, keepdim=True)


def _allclose_abs1(x):
    """jnp.allclose(|x|, 1.0) with its defaults (rtol 1e-5, atol 1e-8)."""
    return (x.abs() - 1.0).abs() <= 1e-8 + 1e-5


def _line_lcs(lines):
    """geometry.py:583-618 for a stack of lines (..., 2, 3): rows U, V, W of every edge frame."""
    ez = lines.new_tensor([0.0, 0.0, 1.0])
    ey = lines.new_tensor([0.0, 1.0, 0.0])
    u = _unit(lines[..., 1, :] - lines[..., 0, :])
    vperp = torch.where(_allclose_abs1(u[..., 2:3]), ey, ez)
    v = _unit(torch.linalg.cross(vperp, u))
    w = torch.linalg.cross(u, v)
    w = _unit(torch.where((w * vperp).sum(-1, keepdim=True) < 0.0, -w, w))
    return torch.stack((u, v, w), dim=-2)


def _normal_polygon(poly):
    """geometry.py:383-425 for a stack of polygons (..., D, 3): unit normal, zero for a degenerate polygon."""
    r = poly - poly.mean(dim=-2, keepdim=True)
    n = (0.5 * torch.linalg.cross(torch.roll(r, 1, dims=-2), r)).sum(dim=-2)
    zero = (n.abs() <= 1e-8).all(dim=-1, keepdim=True)
    return torch.where(zero, torch.zeros_like(n), n / torch.linalg.norm(torch.where(zero, torch.ones_like(n), n), dim=-1, keepdim=True))


def _polygon_lcs(poly):
    """geometry.py:621-657: rows U, V, W of every face frame."""
    ex = poly.new_tensor([1.0, 0.0, 0.0])
    ey = poly.new_tensor([0.0, 1.0, 0.0])
    w = _normal_polygon(poly)
    vperp = torch.where(_allclose_abs1(w[..., 0:1]), ey, ex)
    v = _unit(torch.linalg.cross(w, vperp))
    u = torch.linalg.cross(w, v)
    u = _unit(torch.where((u * vperp).sum(-1, keepdim=True) < 0.0, -u, u))
    return torch.stack((u, v, w), dim=-2)


# ==========================================================================
# Face loads
# ==========================================================================
def nodes_load_from_faces(xyz, faces_load, structure, is_local=False):
    """loads.py:35-71 -- on CUDA tensors one fused kernel (fdm_face_loads) with its VJP."""
    if xyz.is_cuda:
        from .models import _FacesLoad
        return _FacesLoad.apply(xyz, torch.zeros_like(xyz), faces_load, structure, is_local)
    faces_load = calculate_faces_load(xyz, structure.faces_indexed, faces_load, is_local)
    return nodes_tributary_edges_load(edges_tributary_faces_load(xyz, faces_load, structure), structure)


def calculate_faces_load(xyz, faces, faces_load, is_local):
    """loads.py:74-106."""
    if not is_local:
        return faces_load
    return face_load_lcs(xyz, faces, faces_load)


def face_xyz(xyz, face):
    """loads.py:109-145: the face polygon with -1 padding replaced by the face's first vertex."""
    face = _idx(face, xyz)
    pts = xyz[face.clamp(min=0)]
    return torch.where((face >= 0).unsqueeze(-1), pts, pts[..., :1, :])


def face_load_lcs(xyz, face, face_load):
    """loads.py:147-184 for one face or a stack of faces: the load given in the face frame, in global coordinates
    (identity frame on a degenerate normal).  As there, the polygon is `xyz[face]` (a -1 pad reads the last node)."""
    face = _idx(face, xyz)
    fxyz = xyz[face]
    normal = _normal_polygon(fxyz)
    is_zero = (normal.abs() <= 1e-8).all(dim=-1)
    lcs = torch.where(is_zero[..., None, None], torch.eye(3, dtype=xyz.dtype, device=xyz.device), _polygon_lcs(fxyz))
    return (face_load.unsqueeze(-2) @ lcs).squeeze(-2)


def edges_tributary_faces_load(xyz, faces_load, structure):
    """loads.py:186-217."""
    cfv = structure.connectivity_faces_vertices
    cfv = cfv.toarray() if hasattr(cfv, "toarray") else np.asarray(cfv)
    centroids = torch.as_tensor(cfv, dtype=xyz.dtype, device=xyz.device) @ xyz
    return edge_tributary_faces_load(structure.edges_indexed, structure.edges_faces_indexed, xyz, faces_load, centroids)


def edge_tributary_faces_load(edge, edge_faces, xyz, faces_load, face_centroids):
    """loads.py:219-265 for one edge or a stack: sum over the (<= 2) incident faces of area(edge, centroid) * load."""
    edge, f = _idx(edge, xyz), _idx(edge_faces, xyz)
    line = xyz[edge]
    fc = f.clamp(min=0)
    areas = edge_tributary_face_area(line.unsqueeze(-3), face_centroids[fc])
    areas = torch.where(f >= 0, areas, torch.zeros_like(areas))
    return (areas.unsqueeze(-1) * faces_load[fc]).sum(dim=-2)


def edge_tributary_face_area(line, centroid):
    """loads.py:267-288: area of the triangle (line[0], line[1], centroid)."""
    a, b = line[..., 0, :], line[..., 1, :]
    return 0.5 * torch.linalg.norm(torch.linalg.cross(b - a, centroid - a), dim=-1)


# ==========================================================================
# Edge loads
# ==========================================================================
def nodes_load_from_edges(xyz, edges_load, structure, is_local=False):
    """loads.py:296-331 -- on CUDA tensors one fused kernel (fdm_edge_loads) with its VJP."""
    if xyz.is_cuda:
        from .models import _NodesLoad
        return _NodesLoad.apply(xyz, torch.zeros_like(xyz), edges_load, structure, is_local)
    edges_load = calculate_edges_load(xyz, structure.edges_indexed, edges_load, is_local)
    return nodes_tributary_edges_load(edges_tributary_edges_load(xyz, edges_load, structure), structure)


def calculate_edges_load(xyz, edges, edges_load, is_local):
    """loads.py:334-365."""
    if not is_local:
        return edges_load
    return edge_load_lcs(xyz, edges, edges_load)


def edge_load_lcs(xyz, edge, edge_load):
    """loads.py:368-397 for one edge or a stack: the load given in the edge frame, in global coordinates."""
    lcs = _line_lcs(xyz[_idx(edge, xyz)])
    return (edge_load.unsqueeze(-2) @ lcs).squeeze(-2)


def edges_tributary_edges_load(xyz, edges_load, structure):
    """loads.py:400-432: line loads scaled by the edge length."""
    ei = _idx(structure.edges_indexed, xyz)
    return edges_load * torch.linalg.norm(xyz[ei[:, 1]] - xyz[ei[:, 0]], dim=-1, keepdim=True)


# ==========================================================================
# Node helpers
# ==========================================================================
def nodes_tributary_edges_load(edges_load, structure):
    """loads.py:435-465: half of every edge load to each end node."""
    ei = _idx(structure.edges_indexed, edges_load)
    out = torch.zeros((structure.num_nodes, 3), dtype=edges_load.dtype, device=edges_load.device)
    out = out.index_add(0, ei[:, 0], 0.5 * edges_load)
    return out.index_add(0, ei[:, 1], 0.5 * edges_load)
