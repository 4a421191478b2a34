"""Binary-STL reader, convex-hull vertex extraction and mesh inertia.

The reference never touches meshes itself: `mujoco==3.2.3` (un-vendored pip
dependency, /root/reference/Code/mujoco/install.sh:27) compiles
`our_robot/our_robot.xml:33-38` (`<mesh file=... scale=".0008 ...">`) into
per-body mass / centre of mass / inertia and a convex hull per collision mesh.
This module restates that compile step in numpy/scipy.

Mesh inertia algorithms (named switch `inertia_mode`, see DESIGN.md "parity risks"):
  legacy : MuJoCo 3.2.3 default (`exactmeshinertia=false`). Pyramids from the
           area-weighted face centroid, |volume| per face (exact only for
           convex meshes, over-counts otherwise).            [3P-recalled]
  exact  : signed tetrahedra about the origin (divergence theorem).
  convex : exact algorithm on the convex hull.
"""
from __future__ import annotations

import struct

import numpy as np

_STL_REC = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])


def read_stl(path: str) -> np.ndarray:
    """Return triangles [F,3,3] (float64) of a binary STL file."""
    with open(path, "rb") as fh:
        blob = fh.read()
    if len(blob) < 84:
        raise ValueError(f"{path}: not a binary STL (too short)")
    (nface,) = struct.unpack("<I", blob[80:84])
    if 84 + 50 * nface != len(blob):
        raise ValueError(f"{path}: binary STL size mismatch ({nface} faces, {len(blob)} bytes)")
    rec = np.frombuffer(blob, dtype=_STL_REC, count=nface, offset=84)
    return rec["v"].astype(np.float64)


def hull_vertices(points: np.ndarray) -> np.ndarray:
    """Vertices of the convex hull of `points` [N,3], in qhull vertex order.

    The support function of a convex mesh only ever returns hull vertices, so
    the hull vertex set is all plane-vs-mesh collision needs.
    """
    from scipy.spatial import ConvexHull

    pts = np.unique(np.asarray(points, dtype=np.float64), axis=0)
    hull = ConvexHull(pts)
    return pts[np.sort(hull.vertices)].copy()


def hull_triangles(points: np.ndarray) -> np.ndarray:
    """Outward-oriented triangles [F,3,3] of the convex hull of `points`."""
    from scipy.spatial import ConvexHull

    pts = np.unique(np.asarray(points, dtype=np.float64), axis=0)
    hull = ConvexHull(pts)
    tris = pts[hull.simplices]
    centre = pts[hull.vertices].mean(axis=0)
    nrm = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0])
    flip = np.einsum("ij,ij->i", nrm, tris.mean(axis=1) - centre) < 0
    tris[flip] = tris[flip][:, ::-1]
    return tris


def _tet_second_moment(d: np.ndarray, e: np.ndarray, f: np.ndarray, vol: np.ndarray) -> np.ndarray:
    """Sum over tetrahedra (0,d,e,f) of the second-moment matrix  ∫ x xᵀ dV."""
    s = d + e + f
    # ∫ x xᵀ over tet with apex at origin = vol/20 * (d dᵀ + e eᵀ + f fᵀ + s sᵀ)
    c = (
        np.einsum("i,ij,ik->jk", vol, d, d)
        + np.einsum("i,ij,ik->jk", vol, e, e)
        + np.einsum("i,ij,ik->jk", vol, f, f)
        + np.einsum("i,ij,ik->jk", vol, s, s)
    )
    return c / 20.0


def mesh_mass_properties(tris: np.ndarray, mode: str = "legacy"):
    """Volume, centre of mass and unit-density inertia (about the COM) of a triangle mesh.

    Returns (volume, com[3], inertia[3,3]); multiply the inertia by mass/volume.
    """
    tris = np.asarray(tris, dtype=np.float64)
    if mode == "convex":
        verts = tris.reshape(-1, 3)
        return mesh_mass_properties(hull_triangles(verts), "exact")
    a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
    cross = np.cross(b - a, c - a)
    area2 = np.linalg.norm(cross, axis=1)
    good = area2 > 1e-30
    a, b, c, cross, area2 = a[good], b[good], c[good], cross[good], area2[good]
    area = 0.5 * area2
    nrm = cross / area2[:, None]
    cen = (a + b + c) / 3.0
    if mode == "legacy":
        apex = (area[:, None] * cen).sum(axis=0) / area.sum()
    elif mode == "exact":
        apex = np.zeros(3)
    else:
        raise ValueError(f"unknown inertia mode {mode!r}")
    # pyramid volumes with apex `apex`
    vol = np.einsum("ij,ij->i", cen - apex, nrm) * area / 3.0
    if mode == "legacy":
        vol = np.abs(vol)
    volume = vol.sum()
    com = (vol[:, None] * (0.75 * cen + 0.25 * apex)).sum(axis=0) / volume
    # second moments of the same pyramids, taken about the COM
    d, e, f = a - apex, b - apex, c - apex
    cov_apex = _tet_second_moment(d, e, f, vol)          # about the apex
    first = (vol[:, None] * (0.75 * cen + 0.25 * apex - apex)).sum(axis=0)  # ∫ (x-apex) dV
    r = com - apex
    # shift the second-moment matrix from the apex to the COM
    cov = cov_apex - np.outer(first, r) - np.outer(r, first) + volume * np.outer(r, r)
    inertia = np.trace(cov) * np.eye(3) - cov
    return float(volume), com, inertia
