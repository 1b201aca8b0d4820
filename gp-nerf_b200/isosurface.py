"""Iso-surface extraction for the mesh branch (BaseRender.py:266-272, demo_render.py:366-376).

The reference hands the α cube to PyMCubes (`mcubes.marching_cubes(cube, mesh_th)`) and wraps the result in a
`trimesh.Trimesh`.  Neither package is part of the reference tree or of this image, so `Renderer.render_mesh` uses
them when they can be imported and this module otherwise: marching *tetrahedra* (every grid cell split into six
tetrahedra around its main diagonal, at most two triangles per tetrahedron, vertices placed on cell edges / face and
body diagonals by linear interpolation of the two end values – the same placement rule marching cubes uses on its
edges).  Same call contract as `mcubes.marching_cubes`: `(vertices [n,3] float64 in index coordinates of the cube,
triangles [m,3] int)`; the surface is closed and consistently oriented (normals point towards lower values), but the
triangulation differs from PyMCubes' 256-case table, which cannot be pinned here.  Vectorised torch on the cube's
device (the GPU inside `render_mesh`); no per-cell Python loops.
"""
from __future__ import annotations

import numpy as np
import torch

# corners of a cell and the six tetrahedra around the diagonal 0–6
_CORNERS = ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1))
_TETS = ((0, 5, 1, 6), (0, 1, 2, 6), (0, 2, 3, 6), (0, 3, 7, 6), (0, 7, 4, 6), (0, 4, 5, 6))
_EDGES = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))
# per inside-mask (bit i = vertex i above the iso value): edge ids of the polygon, cyclic order
_POLY = {1: (0, 1, 2), 2: (0, 3, 4), 4: (1, 3, 5), 8: (2, 4, 5),
         3: (1, 2, 4, 3), 5: (0, 2, 5, 3), 6: (0, 4, 5, 1)}
for _m in list(_POLY):
    _POLY[15 - _m] = _POLY[_m]


def marching_tetrahedra(cube, iso):
    """cube [X,Y,Z] (numpy or torch, any float dtype) → (vertices float64 [n,3], triangles int64 [m,3]) as numpy."""
    v = torch.as_tensor(cube)
    if v.dim() != 3:
        raise ValueError("cube must be a 3-D array")
    v = v.to(torch.float64)
    dev = v.device
    X, Y, Z = v.shape
    if min(X, Y, Z) < 2:
        return np.zeros((0, 3)), np.zeros((0, 3), dtype=np.int64)
    gx, gy, gz = torch.meshgrid(torch.arange(X - 1, device=dev), torch.arange(Y - 1, device=dev),
                                torch.arange(Z - 1, device=dev), indexing="ij")
    base = torch.stack([gx, gy, gz], -1).reshape(-1, 3)                       # cell origins
    # cells the surface cannot cross are dropped up front (the α cube is >90 % empty)
    lin = lambda p: (p[:, 0] * Y + p[:, 1]) * Z + p[:, 2]                     # noqa: E731
    flat = v.reshape(-1)
    corner_vals = torch.stack([flat[lin(base + torch.tensor(c, device=dev))] for c in _CORNERS], 1)
    above = corner_vals > iso
    live = above.any(1) & ~above.all(1)
    base, corner_vals, above = base[live], corner_vals[live], above[live]
    corner_ids = torch.stack([lin(base + torch.tensor(c, device=dev)) for c in _CORNERS], 1)
    tri_a, tri_b, tri_w, tri_ref = [], [], [], []          # per triangle corner: the edge's two grid vertices + weight
    offs = torch.tensor(_CORNERS, device=dev, dtype=torch.float64)
    for tet in _TETS:
        t = torch.tensor(tet, device=dev)
        tv, ta, tid = corner_vals[:, t], above[:, t], corner_ids[:, t]
        tpos = base[:, None, :].to(torch.float64) + offs[t][None]
        mask = (ta.long() * torch.tensor([1, 2, 4, 8], device=dev)).sum(1)
        for m, poly in _POLY.items():
            sel = torch.nonzero(mask == m).reshape(-1)
            if sel.numel() == 0:
                continue
            e = torch.tensor([_EDGES[k] for k in poly], device=dev)            # [k,2] tet-local vertex pairs
            va, vb = tv[sel][:, e[:, 0]], tv[sel][:, e[:, 1]]
            w = (iso - va) / (vb - va)
            ia, ib = tid[sel][:, e[:, 0]], tid[sel][:, e[:, 1]]
            # reference direction for the orientation: from the vertices above the iso value to those below
            a_sel = ta[sel].to(torch.float64)
            p_sel = tpos[sel]
            c_in = (p_sel * a_sel[..., None]).sum(1) / a_sel.sum(1, keepdim=True)
            c_out = (p_sel * (1 - a_sel)[..., None]).sum(1) / (1 - a_sel).sum(1, keepdim=True)
            ref = c_out - c_in
            for tri in ((0, 1, 2),) if len(poly) == 3 else ((0, 1, 2), (0, 2, 3)):
                tri_a.append(ia[:, tri]); tri_b.append(ib[:, tri]); tri_w.append(w[:, tri]); tri_ref.append(ref)
    if not tri_a:
        return np.zeros((0, 3)), np.zeros((0, 3), dtype=np.int64)
    A, B, Wt, Ref = torch.cat(tri_a), torch.cat(tri_b), torch.cat(tri_w), torch.cat(tri_ref)
    # one vertex per grid edge (lo id, hi id); the weight is re-expressed from the lower id so that both users agree
    swap = A > B
    lo, hi = torch.where(swap, B, A), torch.where(swap, A, B)
    Wt = torch.where(swap, 1.0 - Wt, Wt)
    key = lo * (X * Y * Z) + hi
    uniq, inv = torch.unique(key.reshape(-1), return_inverse=True)
    first = torch.full((uniq.numel(),), key.numel(), device=dev, dtype=torch.long)
    first.scatter_reduce_(0, inv, torch.arange(key.numel(), device=dev), reduce="amin")
    lo_u, hi_u, w_u = lo.reshape(-1)[first], hi.reshape(-1)[first], Wt.reshape(-1)[first]

    def pos(idx):
        return torch.stack([idx // (Y * Z), (idx // Z) % Y, idx % Z], 1).to(torch.float64)
    verts = pos(lo_u) + w_u[:, None] * (pos(hi_u) - pos(lo_u))
    tris = inv.reshape(-1, 3)
    # consistent orientation: normals along `ref` (towards lower values)
    p0, p1, p2 = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    nrm = torch.cross(p1 - p0, p2 - p0, dim=1)
    flip = (nrm * Ref).sum(1) < 0
    tris = torch.where(flip[:, None], tris[:, [0, 2, 1]], tris)
    # degenerate triangles (a vertex exactly on a grid point shared by two edges) are dropped
    ok = (tris[:, 0] != tris[:, 1]) & (tris[:, 1] != tris[:, 2]) & (tris[:, 0] != tris[:, 2])
    return verts.cpu().numpy(), tris[ok].cpu().numpy().astype(np.int64)
