"""Structured polyMesh generators (a blockMesh subset) for the synthetic and
tutorial-shaped cases.

The reference builds all of its meshes with OpenFOAM's ``blockMesh`` from
``constant/polyMesh/blockMeshDict`` (e.g. tutorials/SandiaD/constant/polyMesh/
blockMeshDict:22-48: three hex blocks, one cell thick, wedge half-angle 5 deg).
OpenFOAM is not available here, so this module produces the same kind of
polyMesh (points / faces / owner / neighbour / patches, upper-triangular face
order, owner-outward face normals) directly with numpy.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

# enums of include/pdfb200.h
PATCH_GENERIC, PATCH_EMPTY, PATCH_WEDGE = 0, 1, 2
BC_SLIP, BC_EMPTY, BC_OPEN, BC_INLET_OUTLET = 0, 1, 2, 3
BC_BY_NAME = {"slip": BC_SLIP, "empty": BC_EMPTY, "open": BC_OPEN, "inletOutlet": BC_INLET_OUTLET}


@dataclass
class Patch:
    name: str
    start: int
    size: int
    geom: int = PATCH_GENERIC      # polyPatch class: generic / empty / wedge
    handler: int = BC_SLIP         # mcBoundary type from boundaryHandlers{}
    reflecting: int = 1            # mcOpenBoundary "reflecting" switch


@dataclass
class PolyMesh:
    points: np.ndarray             # [nPoints,3] float64
    face_start: np.ndarray         # [nFaces+1] int32
    face_points: np.ndarray        # [sum] int32
    owner: np.ndarray              # [nFaces] int32
    neighbour: np.ndarray          # [nInternalFaces] int32
    patches: List[Patch]
    n_cells: int
    geometric_d: Tuple[int, int, int] = (1, 1, 1)
    solution_d: Tuple[int, int, int] = (1, 1, 1)
    shape: Tuple[int, int, int] = (0, 0, 0)   # structured (nx,ny,nz), informational
    extra: Dict = field(default_factory=dict)

    @property
    def n_faces(self) -> int:
        return int(self.owner.shape[0])

    @property
    def n_internal_faces(self) -> int:
        return int(self.neighbour.shape[0])

    @property
    def n_boundary_faces(self) -> int:
        return self.n_faces - self.n_internal_faces

    @property
    def n_points(self) -> int:
        return int(self.points.shape[0])

    def patch(self, name: str) -> Patch:
        for p in self.patches:
            if p.name == name:
                return p
        raise KeyError(name)

    def boundary_slice(self, name: str) -> slice:
        """Indices of a patch in boundary-face arrays (face - nInternalFaces)."""
        p = self.patch(name)
        s = p.start - self.n_internal_faces
        return slice(s, s + p.size)

    def face_centres_simple(self) -> np.ndarray:
        """Average of the face points (exact for the planar quads generated here)."""
        n = np.diff(self.face_start)
        assert np.all(n == 4)
        fp = self.face_points.reshape(-1, 4)
        return self.points[fp].mean(axis=1)

    def cell_centres_simple(self) -> np.ndarray:
        """Average of the 8 cell vertices of the structured hexes generated here."""
        return self.extra["cell_centres"]


def _structured(xs: np.ndarray, ys: np.ndarray, zs: np.ndarray,
                side_patches: Dict[str, Sequence[Tuple[str, int, int, int]]],
                point_map: Optional[Callable[[np.ndarray], np.ndarray]] = None,
                side_split: Optional[Dict[str, Callable[[np.ndarray], np.ndarray]]] = None,
                geometric_d=(1, 1, 1), solution_d=(1, 1, 1)) -> PolyMesh:
    """Hex mesh on the tensor grid xs x ys x zs.

    side_patches maps a side ("xmin","xmax","ymin","ymax","zmin","zmax") to a
    list of (name, geom, handler, reflecting); with more than one entry the
    side is split by side_split[side](face_centres) -> sub-patch index.
    """
    nx, ny, nz = len(xs) - 1, len(ys) - 1, len(zs) - 1
    npx, npy = nx + 1, ny + 1

    def pid(i, j, k):
        return (i + npx * (j + npy * k)).astype(np.int64)

    def cid(i, j, k):
        return (i + nx * (j + ny * k)).astype(np.int64)

    K, J, I = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    pts = np.stack([xs[I.ravel()], ys[J.ravel()], zs[K.ravel()]], axis=1).astype(np.float64)

    ck, cj, ci = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()          # cell order: i fastest
    cells = cid(ci, cj, ck)

    # internal faces: per owner cell, neighbours ascending (+x, +y, +z)
    def xface(i, j, k):   # face at x-index i (normal +x)
        return np.stack([pid(i, j, k), pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i, j, k + 1)], axis=1)

    def yface(i, j, k):   # face at y-index j (normal +y)
        return np.stack([pid(i, j, k), pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j, k)], axis=1)

    def zface(i, j, k):   # face at z-index k (normal +z)
        return np.stack([pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k)], axis=1)

    has_x = ci < nx - 1
    has_y = cj < ny - 1
    has_z = ck < nz - 1
    # order key: owner*3 + dir keeps (+x,+y,+z) ascending-neighbour order per owner
    own_list, nei_list, key_list, fp_list = [], [], [], []
    for d, (mask, fn, off) in enumerate(((has_x, xface, (1, 0, 0)), (has_y, yface, (0, 1, 0)), (has_z, zface, (0, 0, 1)))):
        i, j, k = ci[mask], cj[mask], ck[mask]
        own = cells[mask]
        nei = cid(i + off[0], j + off[1], k + off[2])
        fp = fn(i + off[0], j + off[1], k + off[2]) if d == 0 else (fn(i, j + off[1], k) if d == 1 else fn(i, j, k + off[2]))
        own_list.append(own); nei_list.append(nei); key_list.append(own * 3 + d); fp_list.append(fp)
    own_i = np.concatenate(own_list); nei_i = np.concatenate(nei_list)
    key = np.concatenate(key_list); fp_i = np.concatenate(fp_list)
    order = np.argsort(key, kind="stable")
    own_i, nei_i, fp_i = own_i[order], nei_i[order], fp_i[order]
    n_int = own_i.shape[0]

    # boundary faces per side, outward normals
    def side_faces(side):
        if side == "xmin":
            k, j = np.meshgrid(np.arange(nz), np.arange(ny), indexing="ij"); j, k = j.ravel(), k.ravel(); i = np.zeros_like(j)
            f = np.stack([pid(i, j, k), pid(i, j, k + 1), pid(i, j + 1, k + 1), pid(i, j + 1, k)], axis=1)
            return f, cid(i, j, k)
        if side == "xmax":
            k, j = np.meshgrid(np.arange(nz), np.arange(ny), indexing="ij"); j, k = j.ravel(), k.ravel(); i = np.full_like(j, nx)
            return xface(i, j, k), cid(i - 1, j, k)
        if side == "ymin":
            k, i = np.meshgrid(np.arange(nz), np.arange(nx), indexing="ij"); i, k = i.ravel(), k.ravel(); j = np.zeros_like(i)
            f = np.stack([pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j, k + 1), pid(i, j, k + 1)], axis=1)
            return f, cid(i, j, k)
        if side == "ymax":
            k, i = np.meshgrid(np.arange(nz), np.arange(nx), indexing="ij"); i, k = i.ravel(), k.ravel(); j = np.full_like(i, ny)
            return yface(i, j, k), cid(i, j - 1, k)
        if side == "zmin":
            j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij"); i, j = i.ravel(), j.ravel(); k = np.zeros_like(i)
            f = np.stack([pid(i, j, k), pid(i, j + 1, k), pid(i + 1, j + 1, k), pid(i + 1, j, k)], axis=1)
            return f, cid(i, j, k)
        if side == "zmax":
            j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij"); i, j = i.ravel(), j.ravel(); k = np.full_like(i, nz)
            return zface(i, j, k), cid(i, j, k - 1)
        raise KeyError(side)

    if point_map is not None:
        pts = point_map(pts)

    patches: List[Patch] = []
    fps = [fp_i]; owns = [own_i]
    start = n_int
    for side in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax"):
        spec = side_patches[side]
        f, own = side_faces(side)
        if len(spec) == 1:
            sub = np.zeros(f.shape[0], dtype=np.int64)
        else:
            fc = pts[f].mean(axis=1)
            sub = np.asarray(side_split[side](fc), dtype=np.int64)
        for s, (name, geom, handler, refl) in enumerate(spec):
            m = sub == s
            n = int(m.sum())
            patches.append(Patch(name, start, n, geom, handler, refl))
            fps.append(f[m]); owns.append(own[m])
            start += n

    face_points = np.concatenate(fps).astype(np.int32)
    owner = np.concatenate(owns).astype(np.int32)
    n_faces = owner.shape[0]
    face_start = (np.arange(n_faces + 1, dtype=np.int64) * 4).astype(np.int32)
    cc = pts[np.stack([pid(ci, cj, ck), pid(ci + 1, cj, ck), pid(ci, cj + 1, ck), pid(ci + 1, cj + 1, ck),
                       pid(ci, cj, ck + 1), pid(ci + 1, cj, ck + 1), pid(ci, cj + 1, ck + 1), pid(ci + 1, cj + 1, ck + 1)], axis=1)].mean(axis=1)
    return PolyMesh(points=pts, face_start=face_start, face_points=face_points.ravel(), owner=owner,
                    neighbour=nei_i.astype(np.int32), patches=patches, n_cells=nx * ny * nz,
                    geometric_d=tuple(geometric_d), solution_d=tuple(solution_d), shape=(nx, ny, nz),
                    extra={"cell_centres": cc})


def graded(n: int, lo: float, hi: float, ratio: float = 1.0) -> np.ndarray:
    """blockMesh simpleGrading: n cells from lo to hi, last/first width = ratio."""
    if n == 1 or abs(ratio - 1.0) < 1e-14:
        return np.linspace(lo, hi, n + 1)
    r = ratio ** (1.0 / (n - 1))
    w = r ** np.arange(n)
    x = np.concatenate([[0.0], np.cumsum(w)])
    return lo + (hi - lo) * x / x[-1]


def box_mesh(nx: int, ny: int, nz: int, lo=(0.0, 0.0, 0.0), hi=(2.0, 1.0, 1.0),
             xmin=("inlet", BC_INLET_OUTLET), xmax=("outlet", BC_INLET_OUTLET),
             walls=BC_SLIP) -> PolyMesh:
    """The synthetic 3-D box of SURVEY §8(d): x faces open, y/z faces slip."""
    xs = np.linspace(lo[0], hi[0], nx + 1); ys = np.linspace(lo[1], hi[1], ny + 1); zs = np.linspace(lo[2], hi[2], nz + 1)
    sp = {
        "xmin": [(xmin[0], PATCH_GENERIC, xmin[1], 1)], "xmax": [(xmax[0], PATCH_GENERIC, xmax[1], 1)],
        "ymin": [("ymin", PATCH_GENERIC, walls, 1)], "ymax": [("ymax", PATCH_GENERIC, walls, 1)],
        "zmin": [("zmin", PATCH_GENERIC, walls, 1)], "zmax": [("zmax", PATCH_GENERIC, walls, 1)],
    }
    return _structured(xs, ys, zs, sp)


def slab_mesh(nx: int, ny: int, lo=(0.0, 0.0), hi=(6.0, 2.0), thickness=2.0,
              xmin=("xmin", BC_SLIP), xmax=("xmax", BC_SLIP), walls=BC_SLIP) -> PolyMesh:
    """2-D case: one cell thick in z with empty front/back, centred on z=0
    (the shape of tests/positionCorrectionTest, blockMeshDict:33)."""
    xs = np.linspace(lo[0], hi[0], nx + 1); ys = np.linspace(lo[1], hi[1], ny + 1)
    zs = np.array([-0.5 * thickness, 0.5 * thickness])
    sp = {
        "xmin": [(xmin[0], PATCH_GENERIC, xmin[1], 1)], "xmax": [(xmax[0], PATCH_GENERIC, xmax[1], 1)],
        "ymin": [("ymin", PATCH_GENERIC, walls, 1)], "ymax": [("ymax", PATCH_GENERIC, walls, 1)],
        "zmin": [("back", PATCH_EMPTY, BC_EMPTY, 1)], "zmax": [("front", PATCH_EMPTY, BC_EMPTY, 1)],
    }
    return _structured(xs, ys, zs, sp, geometric_d=(1, 1, -1), solution_d=(1, 1, -1))


def wedge_mesh(xs: np.ndarray, rs: np.ndarray, half_angle_deg: float = 5.0,
               inlet_split: Optional[Sequence[Tuple[str, float]]] = None,
               inlet_handler=BC_INLET_OUTLET, outlet_handler=BC_INLET_OUTLET) -> PolyMesh:
    """Axisymmetric wedge, one cell thick (all four tutorial configs, SURVEY
    Appendix E): x stream-wise, y ~ r, the wedge spans +-z.

    inlet_split: [(name, r_upper[, handler]), ...] splits the x-min side by radius
    (e.g. jet / pilot / coflow, or jet / bluffBody (slip) / coflow). Patches:
    inlet(s), "outflow", "axis", "outerWall", "back", "front".
    """
    th = np.deg2rad(half_angle_deg)

    def pmap(p):
        q = p.copy()
        r = p[:, 1]
        q[:, 1] = r * np.cos(th)
        q[:, 2] = np.where(p[:, 2] < 0, -1.0, 1.0) * r * np.sin(th)
        return q

    if inlet_split is None:
        inlet_split = [("inlet", float("inf"))]
    bounds = np.array([it[1] for it in inlet_split])
    sp = {
        "xmin": [(it[0], PATCH_GENERIC, it[2] if len(it) > 2 else inlet_handler, 1) for it in inlet_split],
        "xmax": [("outflow", PATCH_GENERIC, outlet_handler, 1)],
        "ymin": [("axis", PATCH_GENERIC, BC_SLIP, 1)], "ymax": [("outerWall", PATCH_GENERIC, BC_SLIP, 1)],
        "zmin": [("back", PATCH_WEDGE, BC_EMPTY, 1)], "zmax": [("front", PATCH_WEDGE, BC_EMPTY, 1)],
    }
    split = {"xmin": lambda fc: np.searchsorted(bounds, fc[:, 1] / np.cos(th), side="left")}
    return _structured(np.asarray(xs, float), np.asarray(rs, float), np.array([-1.0, 1.0]), sp, point_map=pmap,
                       side_split=split, geometric_d=(1, 1, -1), solution_d=(1, 1, 1))
