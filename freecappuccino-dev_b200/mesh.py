"""Host-side mesh layer: the arrays of the reference's ``geometry`` module, fixture generators,
polyMesh readers and the ``src-par``-style partitioner.

Reference: src/mesh/geometry.f90:12-86 (array names), :118-406 (native reader), :416-664 (derived
geometry), src-par/geometry.f90:140-240,615-633 (``process`` patches, halo buffers).

Everything here is numpy on the host; index arrays are 1-based int32 exactly as the Fortran host
would hand them to the C-ABI (include/fcp.h).
"""
from __future__ import annotations

import dataclasses
import os
import re
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# patch type codes shared with include/fcp.h (FCP_BC_*)
BC_WALL, BC_INLET, BC_OUTLET, BC_SYMMETRY, BC_PRESSURE, BC_PERIODIC, BC_EMPTY, BC_PROCESS = range(8)
BC_NAMES = ["wall", "inlet", "outlet", "symmetry", "pressure", "periodic", "empty", "process"]
BC_CODE = {n: i for i, n in enumerate(BC_NAMES)}


@dataclasses.dataclass
class Mesh:
    """Mirror of the reference ``geometry`` module. Faces: inner first, then patch by patch."""
    numCells: int
    numInnerFaces: int
    numBoundaryFaces: int
    owner: np.ndarray          # int32 [numFaces], 1-based
    neighbour: np.ndarray      # int32 [numInnerFaces], 1-based
    arx: np.ndarray
    ary: np.ndarray
    arz: np.ndarray
    xf: np.ndarray
    yf: np.ndarray
    zf: np.ndarray
    facint: np.ndarray
    Df: np.ndarray
    xc: np.ndarray             # [numCells] (or [numTotal] with ghost copies in the src-par layout)
    yc: np.ndarray
    zc: np.ndarray
    vol: np.ndarray
    bcname: List[str]
    bctype: np.ndarray         # int32 [numBoundaries]
    nfaces: np.ndarray         # int32 [numBoundaries]
    startFace: np.ndarray      # int32 [numBoundaries]; face index = startFace + i, i = 1..nfaces (1-based result)
    # optional topology (needed only to (re)compute geometry)
    points: Optional[np.ndarray] = None       # [numNodes,3]
    face_nodes: Optional[np.ndarray] = None   # int32 [numFaces, nomax] 1-based, 0-padded
    face_nnodes: Optional[np.ndarray] = None  # int32 [numFaces]
    # src-par layout (partitioned meshes only)
    peer_rank: Optional[np.ndarray] = None    # int32 [numBoundaries], -1 if not a process patch
    peer_patch: Optional[np.ndarray] = None   # int32 [numBoundaries], index of the matching patch on the peer
    fpro: Optional[np.ndarray] = None         # [npro] face interpolation factor on halo faces
    cell_global: Optional[np.ndarray] = None  # int64 [numCells] 0-based global cell id
    face_global: Optional[np.ndarray] = None  # int64 [numFaces] 0-based global face id
    # periodic pairs (geometry.f90:82,251-257): a 'periodic' patch names its twin (listed as 'empty') by the twin's startFace
    startFaceTwin: Optional[np.ndarray] = None  # int32 [numBoundaries], 0-based offset like startFace; -1 for other patches
    DfPeriodic: Optional[np.ndarray] = None   # [numBoundaryFaces] partitions only: the "Df(i)" the GLOBAL mesh uses for each periodic face (quirk Q21)

    @property
    def numFaces(self) -> int:
        return self.numInnerFaces + self.numBoundaryFaces

    @property
    def numTotal(self) -> int:
        return self.numCells + self.numBoundaryFaces

    @property
    def numBoundaries(self) -> int:
        return len(self.bcname)

    @property
    def iBndValueStart(self) -> np.ndarray:
        # geometry.f90:282-290  iBndValueStart = numCells + (startFace - numInnerFaces)
        return (self.numCells + self.startFace - self.numInnerFaces).astype(np.int32)

    @property
    def numPeriodic(self) -> int:
        """geometry.f90:253  number of periodic faces, each pair counted once."""
        return int(self.nfaces[self.bctype == BC_PERIODIC].sum())

    def twin_start(self) -> np.ndarray:
        """startFaceTwin per patch (-1 where the patch is not periodic)."""
        if self.startFaceTwin is None:
            return np.full(self.numBoundaries, -1, dtype=np.int32)
        return np.ascontiguousarray(self.startFaceTwin, dtype=np.int32)

    @property
    def npro(self) -> int:
        return int(self.nfaces[self.bctype == BC_PROCESS].sum())

    def patch_faces(self, ib: int) -> np.ndarray:
        """0-based face indices of patch ib."""
        return np.arange(self.startFace[ib], self.startFace[ib] + self.nfaces[ib])

    def boundary_values_of(self, fn) -> np.ndarray:
        """Evaluate fn(x,y,z) at cell centres then boundary-face centres -> field of length numTotal."""
        F = self.numInnerFaces
        out = np.empty(self.numTotal)
        out[: self.numCells] = fn(self.xc[: self.numCells], self.yc[: self.numCells], self.zc[: self.numCells])
        out[self.numCells:] = fn(self.xf[F:], self.yf[F:], self.zf[F:])
        return out


# ---------------------------------------------------------------------------------------------
# derived geometry (vectorised restatement of the algorithm of geometry.f90:416-664; summation order
# differs from the sequential Fortran loop, so results agree with it to rounding, not to the bit)
# ---------------------------------------------------------------------------------------------
def compute_geometry(points: np.ndarray, face_nodes: np.ndarray, face_nnodes: np.ndarray,
                     owner: np.ndarray, neighbour: np.ndarray, numCells: int) -> Dict[str, np.ndarray]:
    nF = owner.shape[0]
    F = neighbour.shape[0]
    third = 1.0 / 3.0
    ar = np.zeros((nF, 3))
    cf = np.zeros((nF, 3))
    asum = np.zeros(nF)
    vol = np.zeros(numCells)
    cc = np.zeros((numCells, 3))
    den = np.zeros(numCells)
    own0 = owner.astype(np.int64) - 1
    nb0 = neighbour.astype(np.int64) - 1
    nomax = face_nodes.shape[1]
    p1 = points[face_nodes[:, 0] - 1]
    for i in range(nomax - 2):
        sel = np.nonzero(face_nnodes - 2 > i)[0]
        if sel.size == 0:
            break
        a1 = p1[sel]
        a2 = points[face_nodes[sel, i + 1] - 1]
        a3 = points[face_nodes[sel, i + 2] - 1]
        n = 0.5 * np.cross(a2 - a1, a3 - a1)
        c = third * (a3 + a2 + a1)
        are = np.sqrt((n * n).sum(1))
        ar[sel] += n
        cf[sel] += are[:, None] * c
        asum[sel] += are
        riSi = (c * n).sum(1)
        o = own0[sel]
        vol += np.bincount(o, weights=third * riSi, minlength=numCells)
        den += np.bincount(o, weights=riSi, minlength=numCells)
        for d in range(3):
            cc[:, d] += np.bincount(o, weights=0.75 * riSi * c[:, d], minlength=numCells)
        inner = sel < F
        if inner.any():
            nn = nb0[sel[inner]]
            r2 = -riSi[inner]
            vol += np.bincount(nn, weights=third * r2, minlength=numCells)
            den += np.bincount(nn, weights=r2, minlength=numCells)
            for d in range(3):
                cc[:, d] += np.bincount(nn, weights=0.75 * r2 * c[inner, d], minlength=numCells)
    cf /= asum[:, None]
    cc /= den[:, None]
    dP = cf[:F] - cc[own0[:F]]
    dN = cf[:F] - cc[nb0]
    djp = np.sqrt((dP * dP).sum(1))
    djn = np.sqrt((dN * dN).sum(1))
    facint = djp / (djp + djn)
    dpn = cc[nb0] - cc[own0[:F]]
    S = ar[:F]
    Df = (S * S).sum(1) / (S * dpn).sum(1)
    return dict(arx=ar[:, 0].copy(), ary=ar[:, 1].copy(), arz=ar[:, 2].copy(),
                xf=cf[:, 0].copy(), yf=cf[:, 1].copy(), zf=cf[:, 2].copy(),
                xc=cc[:, 0].copy(), yc=cc[:, 1].copy(), zc=cc[:, 2].copy(), vol=vol, facint=facint, Df=Df)


def facint_line_plane(mesh: "Mesh") -> np.ndarray:
    """The MPI tree's interpolation factor (quirk Q9, src-par/geometry.f90:780-819 with find_intersection_point :1143-1215): the line P -> N is cut
    with the plane through the FIRST THREE vertices of the face, facint = |P j'| / |P N| -- where the serial tree (geometry.f90:581-606, what
    `compute_geometry` returns) takes dist(f,P) / (dist(f,P) + dist(f,N)) with the face centre f.  The two agree where f lies on the line P N
    (orthogonal meshes) and differ on skewed ones.  Needs the topology (points, face_nodes); inner faces only -- a partition takes its
    process-face factors `fpro` from the global mesh's values (`partition`) and hands them to the library with `Context.set_process_facint`."""
    if mesh.points is None or mesh.face_nodes is None:
        raise ValueError("facint_line_plane needs the mesh topology (points, face_nodes)")
    F = mesh.numInnerFaces
    own0 = mesh.owner[:F].astype(np.int64) - 1
    nb0 = mesh.neighbour.astype(np.int64) - 1
    fn = mesh.face_nodes[:F].astype(np.int64) - 1
    (x1, y1, z1), (x2, y2, z2), (x3, y3, z3) = (mesh.points[fn[:, k]].T for k in range(3))
    x4, y4, z4 = mesh.xc[own0], mesh.yc[own0], mesh.zc[own0]
    x5, y5, z5 = mesh.xc[nb0], mesh.yc[nb0], mesh.zc[nb0]
    tiny = float(np.float32(1e-30))             # `real(dp), parameter :: tiny = 1e-30` is a single-precision literal (geometry.f90:38)
    num = (x2 * (y3 * z4 - y4 * z3) - x1 * (y3 * z4 - y4 * z3) - x3 * (y2 * z4 - y4 * z2) + x1 * (y2 * z4 - y4 * z2) + x3 * (y1 * z4 - y4 * z1)
           - x2 * (y1 * z4 - y4 * z1) + x4 * (y2 * z3 - y3 * z2) - x1 * (y2 * z3 - y3 * z2) - x4 * (y1 * z3 - y3 * z1) + x2 * (y1 * z3 - y3 * z1)
           + x4 * (y1 * z2 - y2 * z1) - x3 * (y1 * z2 - y2 * z1))
    den = (x2 * (y3 * (z5 - z4) - (y5 - y4) * z3) - x1 * (y3 * (z5 - z4) - (y5 - y4) * z3) - x3 * (y2 * (z5 - z4) - (y5 - y4) * z2)
           + x1 * (y2 * (z5 - z4) - (y5 - y4) * z2) + x3 * (y1 * (z5 - z4) - (y5 - y4) * z1) - x2 * (y1 * (z5 - z4) - (y5 - y4) * z1)
           + (x5 - x4) * (y2 * z3 - y3 * z2) - (x5 - x4) * (y1 * z3 - y3 * z1) + (x5 - x4) * (y1 * z2 - y2 * z1) + tiny)
    t = -num / den
    xj, yj, zj = x4 + (x5 - x4) * t, y4 + (y5 - y4) * t, z4 + (z5 - z4) * t
    dpn = np.sqrt((x5 - x4) ** 2 + (y5 - y4) ** 2 + (z5 - z4) ** 2)
    djn = np.sqrt((xj - x4) ** 2 + (yj - y4) ** 2 + (zj - z4) ** 2)
    return djn / dpn


def mesh_from_topology(points, face_nodes, face_nnodes, owner, neighbour, numCells,
                       patches: Sequence[Tuple[str, str, int, int]], geometry: Optional[Dict[str, np.ndarray]] = None) -> Mesh:
    """patches: (name, type, nFaces, startFace) with startFace the 0-based offset of the reference's
    ``boundary`` file (Appendix D of SURVEY.md)."""
    owner = np.ascontiguousarray(owner, dtype=np.int32)
    neighbour = np.ascontiguousarray(neighbour, dtype=np.int32)
    g = geometry if geometry is not None else compute_geometry(points, face_nodes, face_nnodes, owner, neighbour, numCells)
    F = neighbour.shape[0]
    mesh = Mesh(numCells=numCells, numInnerFaces=F, numBoundaryFaces=owner.shape[0] - F, owner=owner, neighbour=neighbour,
                bcname=[p[0] for p in patches], bctype=np.array([BC_CODE[p[1]] for p in patches], dtype=np.int32),
                nfaces=np.array([p[2] for p in patches], dtype=np.int32),
                startFace=np.array([p[3] for p in patches], dtype=np.int32),
                points=points, face_nodes=face_nodes, face_nnodes=face_nnodes, **g)
    if any(len(p) > 4 for p in patches):
        mesh.startFaceTwin = np.array([p[4] if len(p) > 4 else -1 for p in patches], dtype=np.int32)
    return mesh


# ---------------------------------------------------------------------------------------------
# structured hexahedral generator (cells i-fastest, OpenFOAM upper-triangular face order)
# ---------------------------------------------------------------------------------------------
def bump_nodes(n: int, coef: float = 0.2, length: float = 1.0) -> np.ndarray:
    """n+1 node coordinates clustered toward both ends (coef<1), a stand-in for gmsh 'Bump'."""
    t = np.linspace(0.0, 1.0, n + 1)
    if coef == 1.0:
        return length * t
    a = np.arctanh(np.sqrt(1.0 - coef))
    s = 0.5 * (1.0 + np.tanh(a * (2.0 * t - 1.0)) / np.tanh(a))
    s[0], s[-1] = 0.0, 1.0
    return length * s


def hex_mesh(xs: np.ndarray, ys: np.ndarray, zs: np.ndarray,
             patch_types: Optional[Dict[str, str]] = None, distort: float = 0.0) -> Mesh:
    """Tensor-product hex mesh on node coordinates xs, ys, zs.

    Patches (in this order): 'top' (+y, the lid), 'bottom' (-y), 'left' (-x), 'right' (+x),
    'back' (-z), 'front' (+z); types default to wall (2-D cases pass back/front = 'empty'/'symmetry').
    distort>0 moves interior nodes by a smooth deterministic displacement (fraction of the local
    spacing) to make the mesh non-orthogonal for parity tests.
    """
    nx, ny, nz = len(xs) - 1, len(ys) - 1, len(zs) - 1
    types = dict(top="wall", bottom="wall", left="wall", right="wall", back="wall", front="wall")
    if patch_types:
        types.update(patch_types)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    if distort > 0.0:
        lx, ly, lz = xs[-1] - xs[0], ys[-1] - ys[0], zs[-1] - zs[0]
        sx, sy, sz = (X - xs[0]) / lx, (Y - ys[0]) / ly, (Z - zs[0]) / lz
        hx = np.gradient(xs)[:, None, None]
        hy = np.gradient(ys)[None, :, None]
        hz = np.gradient(zs)[None, None, :]
        bub = np.sin(np.pi * sx) * np.sin(np.pi * sy) * (np.sin(np.pi * sz) if nz > 1 else 1.0)
        X = X + distort * hx * bub * np.sin(7.0 * sy + 3.0 * sz + 0.3)
        Y = Y + distort * hy * bub * np.sin(5.0 * sx + 4.0 * sz + 1.1)
        if nz > 1:
            Z = Z + distort * hz * bub * np.sin(6.0 * sx + 5.0 * sy + 2.2)

    def nid(i, j, k):
        return (i + (nx + 1) * (j + (ny + 1) * k) + 1).astype(np.int32)

    npts = (nx + 1) * (ny + 1) * (nz + 1)
    points = np.empty((npts, 3))
    order = np.arange(npts)
    ii = order % (nx + 1)
    jj = (order // (nx + 1)) % (ny + 1)
    kk = order // ((nx + 1) * (ny + 1))
    points[:, 0] = X[ii, jj, kk]
    points[:, 1] = Y[ii, jj, kk]
    points[:, 2] = Z[ii, jj, kk]

    def cid(i, j, k):
        return i + nx * (j + ny * k)

    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    I, J, K = I.ravel(order="F"), J.ravel(order="F"), K.ravel(order="F")  # cell order: i fastest

    fl_nodes, fl_own, fl_nb, fl_key = [], [], [], []
    # +x faces
    m = I < nx - 1
    i, j, k = I[m], J[m], K[m]
    fl_nodes.append(np.stack([nid(i + 1, j, k), nid(i + 1, j + 1, k), nid(i + 1, j + 1, k + 1), nid(i + 1, j, k + 1)], 1))
    fl_own.append(cid(i, j, k)); fl_nb.append(cid(i + 1, j, k)); fl_key.append(cid(i, j, k) * 3 + 0)
    # +y faces
    m = J < ny - 1
    i, j, k = I[m], J[m], K[m]
    fl_nodes.append(np.stack([nid(i, j + 1, k), nid(i, j + 1, k + 1), nid(i + 1, j + 1, k + 1), nid(i + 1, j + 1, k)], 1))
    fl_own.append(cid(i, j, k)); fl_nb.append(cid(i, j + 1, k)); fl_key.append(cid(i, j, k) * 3 + 1)
    # +z faces
    m = K < nz - 1
    i, j, k = I[m], J[m], K[m]
    fl_nodes.append(np.stack([nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)], 1))
    fl_own.append(cid(i, j, k)); fl_nb.append(cid(i, j, k + 1)); fl_key.append(cid(i, j, k) * 3 + 2)
    key = np.concatenate(fl_key)
    perm = np.argsort(key, kind="stable")
    in_nodes = np.concatenate(fl_nodes)[perm]
    in_own = np.concatenate(fl_own)[perm]
    in_nb = np.concatenate(fl_nb)[perm]
    F = in_own.shape[0]

    bnodes, bown, patches = [], [], []
    start = F

    def add_patch(name, nodes, own):
        nonlocal start
        bnodes.append(nodes); bown.append(own)
        patches.append((name, types[name], own.shape[0], start))
        start += own.shape[0]

    i, k = np.meshgrid(np.arange(nx), np.arange(nz), indexing="ij"); i, k = i.ravel(order="F"), k.ravel(order="F")
    j = np.full_like(i, ny - 1)
    add_patch("top", np.stack([nid(i, j + 1, k), nid(i, j + 1, k + 1), nid(i + 1, j + 1, k + 1), nid(i + 1, j + 1, k)], 1), cid(i, j, k))
    j = np.zeros_like(i)
    add_patch("bottom", np.stack([nid(i, j, k), nid(i + 1, j, k), nid(i + 1, j, k + 1), nid(i, j, k + 1)], 1), cid(i, j, k))
    j, k = np.meshgrid(np.arange(ny), np.arange(nz), indexing="ij"); j, k = j.ravel(order="F"), k.ravel(order="F")
    i = np.zeros_like(j)
    add_patch("left", np.stack([nid(i, j, k), nid(i, j, k + 1), nid(i, j + 1, k + 1), nid(i, j + 1, k)], 1), cid(i, j, k))
    i = np.full_like(j, nx - 1)
    add_patch("right", np.stack([nid(i + 1, j, k), nid(i + 1, j + 1, k), nid(i + 1, j + 1, k + 1), nid(i + 1, j, k + 1)], 1), cid(i, j, k))
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij"); i, j = i.ravel(order="F"), j.ravel(order="F")
    k = np.zeros_like(i)
    add_patch("back", np.stack([nid(i, j, k), nid(i, j + 1, k), nid(i + 1, j + 1, k), nid(i + 1, j, k)], 1), cid(i, j, k))
    k = np.full_like(i, nz - 1)
    add_patch("front", np.stack([nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)], 1), cid(i, j, k))

    face_nodes = np.ascontiguousarray(np.concatenate([in_nodes] + bnodes), dtype=np.int32)
    owner = (np.concatenate([in_own] + bown) + 1).astype(np.int32)
    neighbour = (in_nb + 1).astype(np.int32)
    face_nnodes = np.full(owner.shape[0], 4, dtype=np.int32)
    mesh = mesh_from_topology(points, face_nodes, face_nnodes, owner, neighbour, nx * ny * nz, patches)
    # periodic pairs: the 'periodic' patch names the opposite patch (which the boundary file lists as 'empty',
    # examples/channel395/README.md) through startFaceTwin; faces of opposite patches are generated in the same order
    opposite = dict(top="bottom", bottom="top", left="right", right="left", back="front", front="back")
    if any(t == "periodic" for t in types.values()):
        names = [p[0] for p in patches]
        twin = np.full(len(patches), -1, dtype=np.int32)
        for ib, (name, typ, _, _) in enumerate(patches):
            if typ == "periodic":
                it = names.index(opposite[name])
                assert patches[it][1] == "empty", "the twin of a periodic patch is listed as 'empty'"
                twin[ib] = patches[it][3]
        mesh.startFaceTwin = twin
    return mesh


def face_mapping(mesh: Mesh) -> Mesh:
    """geometry.f90:1848-1997 (quirk Q18): pair every face of a periodic patch with the twin-patch face whose OWNER CELL
    CENTRE is nearest (L1 distance, first minimum) to the periodic face's centre, then overwrite the twin patch's `owner`
    and permute its xf,yf,zf,arx,ary,arz in place so that twin face i pairs with face i.  Returns the same mesh object."""
    if mesh.startFaceTwin is None:
        return mesh
    n = mesh.numCells
    for ib in range(mesh.numBoundaries):
        if mesh.bctype[ib] != BC_PERIODIC:
            continue
        nf, s, st = int(mesh.nfaces[ib]), int(mesh.startFace[ib]), int(mesh.startFaceTwin[ib])
        f = np.arange(s, s + nf)
        cand = mesh.owner[st:st + nf].astype(np.int64) - 1
        tmp = np.empty(nf, dtype=np.int32)
        itmp = np.empty(nf, dtype=np.int64)
        for b in range(0, nf, 512):
            fb = f[b:b + 512]
            d = (np.abs(mesh.xf[fb, None] - mesh.xc[None, cand]) + np.abs(mesh.yf[fb, None] - mesh.yc[None, cand])
                 + np.abs(mesh.zf[fb, None] - mesh.zc[None, cand]))
            k = np.argmin(d, axis=1)          # first minimum, like the strict `<` of the reference loop
            itmp[b:b + 512] = k
            tmp[b:b + 512] = cand[k] + 1
        mesh.owner[st:st + nf] = tmp
        for arr in (mesh.xf, mesh.yf, mesh.zf, mesh.arx, mesh.ary, mesh.arz):
            arr[st:st + nf] = arr[st + itmp].copy()
    assert n == mesh.numCells
    return mesh


def cavity_mesh(n: int, nz: Optional[int] = None, length: float = 1.0, bump: float = 1.0, distort: float = 0.0,
                two_d_type: str = "empty") -> Mesh:
    """Lid-driven cavity fixtures: n x n x nz. nz=None -> n (3-D, six walls); nz=1 -> 2-D slab with
    back/front patches of type two_d_type (examples/cavity uses 'empty', test/ uses 'symmetry')."""
    nz = n if nz is None else nz
    xs = bump_nodes(n, bump, length)
    ys = bump_nodes(n, bump, length)
    if nz == 1:
        zs = np.array([0.0, 0.05 * length if length == 1.0 else 0.1 * length])
        return hex_mesh(xs, ys, zs, dict(back=two_d_type, front=two_d_type), distort)
    zs = bump_nodes(nz, bump, length)
    return hex_mesh(xs, ys, zs, None, distort)


def polyhedral_mesh(nx: int, ny: Optional[int] = None, nz: Optional[int] = None, distort: float = 0.15,
                    patch_types: Optional[Dict[str, str]] = None) -> Mesh:
    """Synthetic POLYHEDRAL mesh (BASELINE config 5 stand-in): a distorted nx x ny x nz hex grid whose cells are agglomerated
    pairwise along x in a brick pattern staggered in both y and z (pair start parity = (j+k) mod 2).  Every interior
    polyhedron has 10 quadrilateral faces and 10 DISTINCT neighbours (no two cells share more than one face, so the CSR
    pattern has no duplicate columns); cells at the x ends stay single hexahedra (6 faces) -- a mixed-cell mesh with
    variable row lengths, like examples/elbow3D or transientConductionFlange.  Inner faces are renumbered in OpenFOAM's
    upper-triangular order (owner ascending, then neighbour), boundary faces keep their patches."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    h = hex_mesh(np.linspace(0.0, 1.0, nx + 1), np.linspace(0.0, 1.0, ny + 1), np.linspace(0.0, 1.0, nz + 1), patch_types, distort)
    c = np.arange(h.numCells)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
    s = (j + k) % 2
    head = np.where((i - s) % 2 == 0, i, i - 1)          # x index of the first hex of the pair (may be -1 -> single at the left end)
    head = np.where(head < 0, 0, head)
    key = head + nx * (j + ny * k)                       # id of the pair's head hex; singles map to themselves
    uniq, poly = np.unique(key, return_inverse=True)     # ascending head id keeps the cells roughly i-fastest
    ncell = uniq.size
    Fi = h.numInnerFaces
    own = poly[h.owner.astype(np.int64) - 1]
    nb = poly[h.neighbour.astype(np.int64) - 1]
    keep = own[:Fi] != nb                                # drop the face between the two hexes of a pair
    fo, fn, fnodes = own[:Fi][keep], nb[keep], h.face_nodes[:Fi][keep].copy()
    swap = fo > fn                                       # owner must be the lower cell id: flip the face
    fo2, fn2 = np.where(swap, fn, fo), np.where(swap, fo, fn)
    fnodes[swap] = fnodes[swap][:, ::-1]
    order = np.lexsort((fn2, fo2))
    fo2, fn2, fnodes = fo2[order], fn2[order], fnodes[order]
    pair = fo2.astype(np.int64) * ncell + fn2
    assert np.unique(pair).size == pair.size, "two polyhedra share more than one face"
    nF_in = fo2.size
    removed = Fi - nF_in
    face_nodes = np.ascontiguousarray(np.concatenate([fnodes, h.face_nodes[Fi:]]), dtype=np.int32)
    owner = (np.concatenate([fo2, own[Fi:]]) + 1).astype(np.int32)
    neighbour = (fn2 + 1).astype(np.int32)
    patches = [(h.bcname[ib], BC_NAMES[h.bctype[ib]], int(h.nfaces[ib]), int(h.startFace[ib]) - removed) for ib in range(h.numBoundaries)]
    return mesh_from_topology(h.points, face_nodes, np.full(owner.shape[0], 4, dtype=np.int32), owner, neighbour, ncell, patches)


# ---------------------------------------------------------------------------------------------
# polyMesh readers (Appendix D of SURVEY.md; geometry.f90:118-406 native, :846-1717 OpenFOAM)
# ---------------------------------------------------------------------------------------------
def _strip_foam(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//.*", "", text)
    text = re.sub(r"FoamFile\s*\{.*?\}", "", text, flags=re.S)
    return text


def read_boundary_simplified(path: str) -> List[Tuple[str, str, int, int]]:
    out = []
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t or t[0].startswith("#"):
                continue
            if t[1] == "periodic":        # geometry.f90:251-257: a 5th integer, the startFace of the twin patch
                out.append((t[0], t[1], int(t[2]), int(t[3]), int(t[4])))
            else:
                out.append((t[0], t[1], int(t[2]), int(t[3])))
    return out


def read_polymesh_openfoam(dirname: str) -> Mesh:
    """OpenFOAM ASCII points/faces/owner/neighbour + the reference's simplified 4-column boundary."""
    def body(name):
        with open(os.path.join(dirname, name)) as fh:
            raw = fh.read()
        return raw, _strip_foam(raw)

    raw_owner, t_owner = body("owner")
    mnote = re.search(r"nCells:\s*(\d+)", raw_owner)
    _, t_pts = body("points")
    _, t_faces = body("faces")
    _, t_nb = body("neighbour")
    pts = np.array(re.findall(r"\(\s*([-+0-9.eE]+)\s+([-+0-9.eE]+)\s+([-+0-9.eE]+)\s*\)", t_pts), dtype=float)
    own = np.array(t_owner.replace("(", " ").replace(")", " ").split()[1:], dtype=np.int64)
    nb = np.array(t_nb.replace("(", " ").replace(")", " ").split()[1:], dtype=np.int64)
    fl = re.findall(r"(\d+)\s*\(([^()]*)\)", t_faces)
    nomax = max(int(a) for a, _ in fl)
    fn = np.zeros((len(fl), nomax), dtype=np.int32)
    fnn = np.zeros(len(fl), dtype=np.int32)
    for r, (a, b) in enumerate(fl):
        ids = [int(v) + 1 for v in b.split()]
        fnn[r] = int(a)
        fn[r, : len(ids)] = ids
    numCells = int(mnote.group(1)) if mnote else int(own.max()) + 1
    patches = read_boundary_simplified(os.path.join(dirname, "boundary"))
    return face_mapping(mesh_from_topology(pts, fn, fnn, (own + 1).astype(np.int32), (nb + 1).astype(np.int32), numCells, patches))   # geometry.f90:110


def read_polymesh_native(dirname: str) -> Mesh:
    """Native format: size, points, faces ('n i1..in', 1-based), owner, neighbour, boundary."""
    with open(os.path.join(dirname, "size")) as fh:
        sizes = [int(line.split()[0]) for line in fh if line.strip()]
    numNodes, numCells, numInner, numBnd, numFaces = sizes[:5]
    pts = np.loadtxt(os.path.join(dirname, "points")).reshape(-1, 3)[:numNodes]
    own = np.loadtxt(os.path.join(dirname, "owner"), dtype=np.int64).ravel()[:numFaces]
    nb = np.loadtxt(os.path.join(dirname, "neighbour"), dtype=np.int64).ravel()[:numInner]
    rows = []
    with open(os.path.join(dirname, "faces")) as fh:
        for line in fh:
            t = line.split()
            if t:
                rows.append([int(v) for v in t])
    nomax = max(r[0] for r in rows)
    fn = np.zeros((len(rows), nomax), dtype=np.int32)
    fnn = np.array([r[0] for r in rows], dtype=np.int32)
    for r, row in enumerate(rows):
        fn[r, : row[0]] = row[1: 1 + row[0]]
    patches = read_boundary_simplified(os.path.join(dirname, "boundary"))
    return face_mapping(mesh_from_topology(pts, fn, fnn, own.astype(np.int32), nb.astype(np.int32), numCells, patches))   # geometry.f90:110


def write_polymesh_native(mesh: Mesh, dirname: str) -> None:
    """Writes the native format the serial reference reads (geometry.f90:118-406)."""
    os.makedirs(dirname, exist_ok=True)
    with open(os.path.join(dirname, "size"), "w") as fh:
        fh.write(f"{mesh.points.shape[0]} numNodes\n{mesh.numCells} numCells\n{mesh.numInnerFaces} numInnerFaces\n"
                 f"{mesh.numBoundaryFaces} numBoundaryFaces\n{mesh.numFaces} numFaces\n")
    np.savetxt(os.path.join(dirname, "points"), mesh.points, fmt="%.17g")
    with open(os.path.join(dirname, "faces"), "w") as fh:
        for n, row in zip(mesh.face_nnodes, mesh.face_nodes):
            fh.write(str(int(n)) + " " + " ".join(str(int(v)) for v in row[:n]) + "\n")
    np.savetxt(os.path.join(dirname, "owner"), mesh.owner, fmt="%d")
    np.savetxt(os.path.join(dirname, "neighbour"), mesh.neighbour, fmt="%d")
    with open(os.path.join(dirname, "boundary"), "w") as fh:
        fh.write("# name type nFaces startFace\n")
        for ib in range(mesh.numBoundaries):
            twin = f" {mesh.startFaceTwin[ib]}" if mesh.bctype[ib] == BC_PERIODIC else ""
            fh.write(f"{mesh.bcname[ib]} {BC_NAMES[mesh.bctype[ib]]} {mesh.nfaces[ib]} {mesh.startFace[ib]}{twin}\n")


# ---------------------------------------------------------------------------------------------
# src-par on-disk layout (src-par/geometry.f90:118-260): processorK/constant/polyMesh/{points,faces,owner,neighbour}
# in OpenFOAM ASCII format, the simplified four-column `boundary` (process patches typed 'process') and `process`
# (header line + the neighbour rank of every process patch, in patch order).
# ---------------------------------------------------------------------------------------------
_FOAM_HEADER = """FoamFile
{{
    version     2.0;
    format      ascii;
    class       {cls};
    note        "nPoints: {npts} nCells: {ncells} nFaces: {nfaces} nInternalFaces: {ninner}";
    location    "constant/polyMesh";
    object      {obj};
}}

"""


def write_partition_srcpar(gmesh: Mesh, parts: List[Mesh], root: str) -> None:
    """Writes the partitions made by ``partition(gmesh, ...)`` as the directory tree the reference's MPI build reads.
    Cut faces keep the node order of the global face when this rank owns the face's owner cell and the reversed order
    otherwise, so that the area vector computed from the nodes points out of the partition (src-par/geometry.f90:826-871)."""
    if gmesh.points is None or gmesh.face_nodes is None:
        raise ValueError("write_partition_srcpar needs the global mesh topology (points, face_nodes)")
    own0 = gmesh.owner.astype(np.int64) - 1
    for r, part in enumerate(parts):
        d = os.path.join(root, f"processor{r}", "constant", "polyMesh")
        os.makedirs(d, exist_ok=True)
        gf = part.face_global
        fn = gmesh.face_nodes[gf].astype(np.int64)
        nn = gmesh.face_nnodes[gf].astype(np.int64)
        # a face is flipped when its local owner is not the global owner
        flipped = part.cell_global[part.owner.astype(np.int64) - 1] != own0[gf]
        used = np.unique(fn[fn > 0])
        g2l = np.zeros(gmesh.points.shape[0] + 1, dtype=np.int64)
        g2l[used] = np.arange(1, used.size + 1)
        hdr = dict(npts=used.size, ncells=part.numCells, nfaces=part.numFaces, ninner=part.numInnerFaces)
        with open(os.path.join(d, "points"), "w") as fh:
            fh.write(_FOAM_HEADER.format(cls="vectorField", obj="points", **hdr))
            fh.write(f"{used.size}\n(\n")
            for x, y, z in gmesh.points[used - 1]:
                fh.write(f"({x:.17g} {y:.17g} {z:.17g})\n")
            fh.write(")\n")
        with open(os.path.join(d, "faces"), "w") as fh:
            fh.write(_FOAM_HEADER.format(cls="faceList", obj="faces", **hdr))
            fh.write(f"{gf.size}\n(\n")
            for k in range(gf.size):
                ids = g2l[fn[k, : nn[k]]] - 1
                if flipped[k]:
                    ids = np.concatenate([ids[:1], ids[:0:-1]])   # reversed orientation, same first node: the triangle fan of the
                                                                   # geometry routine (geometry.f90:416-470) stays the same on both ranks
                fh.write(f"{nn[k]}(" + " ".join(str(int(v)) for v in ids) + ")\n")
            fh.write(")\n")
        for name, arr in (("owner", part.owner), ("neighbour", part.neighbour)):
            with open(os.path.join(d, name), "w") as fh:
                fh.write(_FOAM_HEADER.format(cls="labelList", obj=name, **hdr))
                fh.write(f"{arr.size}\n(\n" + "\n".join(str(int(v) - 1) for v in arr) + "\n)\n")
        with open(os.path.join(d, "boundary"), "w") as fh:
            fh.write("# bcName bcType nFaces startFace\n")
            for ib in range(part.numBoundaries):
                fh.write(f"{part.bcname[ib]} {BC_NAMES[part.bctype[ib]]} {part.nfaces[ib]} {part.startFace[ib]}\n")
        with open(os.path.join(d, "process"), "w") as fh:
            fh.write("# neighbour process of every 'process' patch, in patch order\n")
            for ib in range(part.numBoundaries):
                if part.bctype[ib] == BC_PROCESS:
                    fh.write(f"{int(part.peer_rank[ib])}\n")


def read_partition_srcpar(root: str, rank: int) -> Mesh:
    """Reads processor<rank>/constant/polyMesh like src-par/geometry.f90:118-260: geometry from the node coordinates,
    `peer_rank` from the `process` file (k-th row = neighbour of the k-th process patch).  Ghost copies of xc, yc, zc, vol
    are left at zero: fcp_comm_init's first exchanges fill them (src-par/geometry.f90:769-773)."""
    d = os.path.join(root, f"processor{rank}", "constant", "polyMesh")
    m = read_polymesh_openfoam(d)
    with open(os.path.join(d, "process")) as fh:
        nbr = [int(t.split()[0]) for t in fh.read().splitlines()[1:] if t.strip()]
    peer = np.full(m.numBoundaries, -1, dtype=np.int32)
    k = 0
    for ib in range(m.numBoundaries):
        if m.bctype[ib] == BC_PROCESS:
            peer[ib] = nbr[k]
            k += 1
    if k != len(nbr):
        raise ValueError(f"process file lists {len(nbr)} connections, boundary file has {k} process patches")
    m.peer_rank = peer
    m.peer_patch = np.full(m.numBoundaries, -1, dtype=np.int32)
    nT = m.numTotal
    for name in ("xc", "yc", "zc", "vol"):
        full = np.zeros(nT)
        full[: m.numCells] = getattr(m, name)[: m.numCells]
        setattr(m, name, full)
    return m


def load_mesh_npz(path: str) -> Mesh:
    d = np.load(path, allow_pickle=False)
    patches = [(str(n), str(t), int(c), int(s)) for n, t, c, s in
               zip(d["patch_name"], d["patch_type"], d["patch_nfaces"], d["patch_start"])]
    return mesh_from_topology(d["points"], d["face_nodes"], d["face_nnodes"], d["owner"], d["neighbour"],
                              int(d["numCells"]), patches)


def save_mesh_npz(mesh: Mesh, path: str) -> None:
    np.savez_compressed(path, points=mesh.points, face_nodes=mesh.face_nodes, face_nnodes=mesh.face_nnodes,
                        owner=mesh.owner, neighbour=mesh.neighbour, numCells=mesh.numCells,
                        patch_name=np.array(mesh.bcname), patch_type=np.array([BC_NAMES[t] for t in mesh.bctype]),
                        patch_nfaces=mesh.nfaces, patch_start=mesh.startFace)


# ---------------------------------------------------------------------------------------------
# partitioner: global mesh + cell->rank map  ->  per-rank meshes in the src-par layout
# (src-par/geometry.f90:140-240: `process` patches, one per neighbour rank, same face order on both
#  sides; ghost value of a process face lives in that face's boundary slot, exchange.f90:110-127;
#  ghost xc,yc,zc,vol copies, geometry.f90:769-773; fpro, :826-871)
# ---------------------------------------------------------------------------------------------
def slab_partition(mesh: Mesh, nranks: int) -> np.ndarray:
    """Contiguous blocks of cells (z-slabs for an i-fastest structured mesh)."""
    return (np.arange(mesh.numCells, dtype=np.int64) * nranks // mesh.numCells).astype(np.int32)


def partition(mesh: Mesh, cell_rank: np.ndarray) -> List[Mesh]:
    nranks = int(cell_rank.max()) + 1
    F = mesh.numInnerFaces
    own0 = mesh.owner.astype(np.int64) - 1
    nb0 = mesh.neighbour.astype(np.int64) - 1
    ro = cell_rank[own0]                     # rank of owner, all faces
    rn = cell_rank[nb0]                      # rank of neighbour, inner faces
    parts: List[Mesh] = []
    # patch index bookkeeping so that peer_patch can be filled after all ranks are built
    proc_patch_index: List[Dict[int, int]] = []
    for r in range(nranks):
        cells = np.nonzero(cell_rank == r)[0]
        g2l = np.full(mesh.numCells, -1, dtype=np.int64)
        g2l[cells] = np.arange(cells.size)
        inner = np.nonzero((ro[:F] == r) & (rn == r))[0]
        f_idx = [inner]
        f_own = [g2l[own0[inner]]]
        f_flip = [np.zeros(inner.size, dtype=bool)]
        names, types, counts = [], [], []
        for ib in range(mesh.numBoundaries):
            pf = mesh.patch_faces(ib)
            pf = pf[ro[pf] == r]
            f_idx.append(pf); f_own.append(g2l[own0[pf]]); f_flip.append(np.zeros(pf.size, dtype=bool))
            names.append(mesh.bcname[ib]); types.append(int(mesh.bctype[ib])); counts.append(pf.size)
        # periodic pairs: a rank keeps face i of a periodic patch together with face i of its twin (the partition must not separate a pair)
        twin_local = np.full(mesh.numBoundaries, -1, dtype=np.int64)
        if mesh.startFaceTwin is not None:
            for ib in range(mesh.numBoundaries):
                if mesh.bctype[ib] != BC_PERIODIC:
                    continue
                it = [jb for jb in range(mesh.numBoundaries) if jb != ib and mesh.startFace[jb] == mesh.startFaceTwin[ib] and mesh.nfaces[jb] == mesh.nfaces[ib]][0]
                keep_p = ro[mesh.patch_faces(ib)] == r
                keep_t = ro[mesh.patch_faces(it)] == r
                if not np.array_equal(keep_p, keep_t):
                    raise ValueError(f"partition separates periodic pairs of patch {mesh.bcname[ib]}: keep both cells of a pair on one rank")
                twin_local[ib] = it
        peers = {}
        cutO = np.nonzero((ro[:F] == r) & (rn != r))[0]      # we own P: outward normal = S
        cutN = np.nonzero((ro[:F] != r) & (rn == r))[0]      # we own N: outward normal = -S
        other = np.concatenate([rn[cutO], ro[:F][cutN]])
        faces = np.concatenate([cutO, cutN])
        flip = np.concatenate([np.zeros(cutO.size, dtype=bool), np.ones(cutN.size, dtype=bool)])
        loc = np.concatenate([g2l[own0[cutO]], g2l[nb0[cutN]]])
        ppi: Dict[int, int] = {}
        for q in np.unique(other):
            sel = np.nonzero(other == q)[0]
            sel = sel[np.argsort(faces[sel], kind="stable")]   # global face order on both sides
            ppi[int(q)] = len(names)
            f_idx.append(faces[sel]); f_own.append(loc[sel]); f_flip.append(flip[sel])
            names.append(f"procBoundary{r}to{int(q)}"); types.append(BC_PROCESS); counts.append(sel.size)
            peers[len(names) - 1] = int(q)
        proc_patch_index.append(ppi)
        fidx = np.concatenate(f_idx)
        fown = np.concatenate(f_own)
        fflip = np.concatenate(f_flip)
        sgn = np.where(fflip, -1.0, 1.0)
        nFi = inner.size
        starts = nFi + np.concatenate([[0], np.cumsum(counts)[:-1]]) if counts else np.zeros(0)
        nloc = cells.size
        nB = fidx.size - nFi
        # ghost copies of cell-centre data in the boundary slots of process faces
        xc = np.zeros(nloc + nB); yc = np.zeros(nloc + nB); zc = np.zeros(nloc + nB); vol = np.zeros(nloc + nB)
        xc[:nloc], yc[:nloc], zc[:nloc], vol[:nloc] = mesh.xc[cells], mesh.yc[cells], mesh.zc[cells], mesh.vol[cells]
        peer_rank = np.full(len(names), -1, dtype=np.int32)
        fpro = []
        for ib, q in peers.items():
            s = int(starts[ib]); c = counts[ib]
            gf = fidx[s: s + c]
            gl = np.where(fflip[s: s + c], own0[gf], nb0[gf])   # the off-rank cell (global id)
            sl = slice(nloc + s - nFi, nloc + s - nFi + c)
            xc[sl], yc[sl], zc[sl], vol[sl] = mesh.xc[gl], mesh.yc[gl], mesh.zc[gl], mesh.vol[gl]
            lam = mesh.facint[gf]
            fpro.append(np.where(fflip[s: s + c], 1.0 - lam, lam))   # weight of the ghost cell seen from this side
            peer_rank[ib] = q
        part = Mesh(numCells=nloc, numInnerFaces=nFi, numBoundaryFaces=nB,
                    owner=(fown + 1).astype(np.int32), neighbour=(g2l[nb0[inner]] + 1).astype(np.int32),
                    arx=mesh.arx[fidx] * sgn, ary=mesh.ary[fidx] * sgn, arz=mesh.arz[fidx] * sgn,
                    xf=mesh.xf[fidx].copy(), yf=mesh.yf[fidx].copy(), zf=mesh.zf[fidx].copy(),
                    facint=mesh.facint[inner].copy(), Df=mesh.Df[inner].copy(),
                    xc=xc, yc=yc, zc=zc, vol=vol, bcname=names, bctype=np.array(types, dtype=np.int32),
                    nfaces=np.array(counts, dtype=np.int32), startFace=np.asarray(starts, dtype=np.int32),
                    peer_rank=peer_rank, peer_patch=np.full(len(names), -1, dtype=np.int32),
                    fpro=np.concatenate(fpro) if fpro else np.zeros(0),
                    cell_global=cells.astype(np.int64), face_global=fidx.astype(np.int64))
        if (twin_local >= 0).any():
            tw = np.full(len(names), -1, dtype=np.int32)
            for ib in range(mesh.numBoundaries):
                if twin_local[ib] >= 0:
                    tw[ib] = int(starts[twin_local[ib]])
            part.startFaceTwin = tw
            # quirk Q21: the reference reads Df(i), i = the face's ordinal in its (global) patch = the Df of GLOBAL inner face i
            dfp = np.zeros(nB)
            for ib in range(mesh.numBoundaries):
                if twin_local[ib] < 0:
                    continue
                gpf = mesh.patch_faces(ib)
                ordinal = {int(fg): i for i, fg in enumerate(gpf)}
                for jb in (ib, int(twin_local[ib])):
                    s0 = int(starts[jb]); cn = counts[jb]
                    gfaces = fidx[s0: s0 + cn]
                    base = gpf if jb == ib else mesh.patch_faces(jb)
                    pos = {int(fg): i for i, fg in enumerate(base)}
                    dfp[s0 - nFi: s0 - nFi + cn] = [mesh.Df[pos[int(fg)]] for fg in gfaces]
            part.DfPeriodic = dfp
        parts.append(part)
    for r, part in enumerate(parts):
        for ib in range(part.numBoundaries):
            q = int(part.peer_rank[ib])
            if q >= 0:
                part.peer_patch[ib] = proc_patch_index[q][r]
    return parts


def process_face_flipped(gmesh: Mesh, part: Mesh) -> np.ndarray:
    """Per process face of `part` (patch order): 1 when the partition's cell is the face's NEIGHBOUR in the unpartitioned mesh, i.e. the face's owner
    lives on the peer rank.  Input of `fcp_set_process_orientation` (the k-omega SST pair takes 1/sigma from a face's owner cell)."""
    out = []
    for ib in range(part.numBoundaries):
        if part.bctype[ib] != BC_PROCESS:
            continue
        pf = part.patch_faces(ib)
        local_owner = part.cell_global[part.owner[pf].astype(np.int64) - 1]
        global_owner = gmesh.owner[part.face_global[pf]].astype(np.int64) - 1
        out.append((local_owner != global_owner).astype(np.int32))
    return np.concatenate(out) if out else np.zeros(0, dtype=np.int32)


def drop_empty_patches(parts: List[Mesh]) -> List[Mesh]:
    """What a real src-par decomposition hands a rank: the physical patches that have no face on the rank are simply absent from its boundary
    file (``partition`` keeps them with zero faces, which hides every bug that derives a collective from the LOCAL patch table).  The faces do
    not move; only the patch tables shrink, and ``peer_patch`` is re-indexed on both sides."""
    import copy
    maps = []
    out = []
    for p in parts:
        if p.startFaceTwin is not None and (np.asarray(p.startFaceTwin) >= 0).any():
            raise ValueError("drop_empty_patches: not for partitions with periodic pairs")
        keep = [ib for ib in range(p.numBoundaries) if p.nfaces[ib] > 0 or p.bctype[ib] == BC_PROCESS]
        maps.append({old: new for new, old in enumerate(keep)})
        q = copy.copy(p)
        q.bcname = [p.bcname[ib] for ib in keep]
        q.bctype = np.ascontiguousarray(p.bctype[keep], dtype=np.int32)
        q.nfaces = np.ascontiguousarray(p.nfaces[keep], dtype=np.int32)
        q.startFace = np.ascontiguousarray(p.startFace[keep], dtype=np.int32)
        q.peer_rank = np.ascontiguousarray(p.peer_rank[keep], dtype=np.int32)
        q.peer_patch = np.ascontiguousarray(p.peer_patch[keep], dtype=np.int32)
        q.startFaceTwin = None
        out.append(q)
    for q in out:
        for ib in range(q.numBoundaries):
            if q.peer_rank[ib] >= 0:
                q.peer_patch[ib] = maps[int(q.peer_rank[ib])][int(q.peer_patch[ib])]
    return out


def localize_matrix(gmesh: Mesh, gcsr, a_global: np.ndarray, part: Mesh, pcsr) -> Tuple[np.ndarray, np.ndarray]:
    """Restrict a global CSR matrix (values ``a_global`` on the pattern ``gcsr`` of ``gmesh``) to one partition in the
    src-par layout: returns (a_local[nnz_local], apr[npro]) -- src-par/sparse_matrix.f90:25,173 (``apr`` holds the one
    off-rank coefficient of every ``process`` face, patch order).  ``gcsr``/``pcsr`` expose ia, diag, icell_jcell,
    jcell_icell (1-based), e.g. oracle.orc_py.Csr or the arrays returned by Context.csr_pattern()."""
    Fi = part.numInnerFaces
    a_loc = np.zeros(int(pcsr.ia[-1]) - 1)
    fg = part.face_global[:Fi]
    a_loc[pcsr.icell_jcell[:Fi] - 1] = a_global[gcsr.icell_jcell[fg] - 1]
    a_loc[pcsr.jcell_icell[:Fi] - 1] = a_global[gcsr.jcell_icell[fg] - 1]
    if part.numPeriodic:                      # the twin entries of periodic pairs: positions F + l, l counting periodic faces in patch order
        gl = {}
        l = 0
        for ib in range(gmesh.numBoundaries):
            if gmesh.bctype[ib] == BC_PERIODIC:
                for fgl in gmesh.patch_faces(ib):
                    gl[int(fgl)] = gmesh.numInnerFaces + l
                    l += 1
        l = 0
        for ib in range(part.numBoundaries):
            if part.bctype[ib] == BC_PERIODIC:
                for fl in part.patch_faces(ib):
                    k = gl[int(part.face_global[fl])]
                    a_loc[pcsr.icell_jcell[Fi + l] - 1] = a_global[gcsr.icell_jcell[k] - 1]
                    a_loc[pcsr.jcell_icell[Fi + l] - 1] = a_global[gcsr.jcell_icell[k] - 1]
                    l += 1
    a_loc[pcsr.diag - 1] = a_global[gcsr.diag[part.cell_global] - 1]
    apr = []
    own0 = gmesh.owner.astype(np.int64) - 1
    for ib in range(part.numBoundaries):
        if part.bctype[ib] != BC_PROCESS:
            continue
        pf = part.patch_faces(ib)
        gfaces = part.face_global[pf]
        local_is_owner = own0[gfaces] == part.cell_global[part.owner[pf] - 1]
        apr.append(np.where(local_is_owner, a_global[gcsr.icell_jcell[gfaces] - 1], a_global[gcsr.jcell_icell[gfaces] - 1]))
    return a_loc, (np.concatenate(apr) if apr else np.zeros(0))


# ---------------------------------------------------------------------------------------------
# fast tensor-product generator (bench sizes): same topology, numbering and patch order as hex_mesh(), geometry written
# down analytically for an axis-aligned box grid instead of going through points/face_nodes (agrees with the triangle-fan
# formulas of geometry.f90:416-664 to rounding).  Optionally one brick of a px x py x pz block decomposition in the
# src-par layout: sides on internal cuts become `process` patches (appended after the physical patches, one per
# neighbour rank, faces in the same order on both sides).
# ---------------------------------------------------------------------------------------------
def hex_mesh_fast(xs: np.ndarray, ys: np.ndarray, zs: np.ndarray, patch_types: Optional[Dict[str, str]] = None,
                  peer: Optional[Dict[str, int]] = None) -> Mesh:
    """peer: side name -> neighbour rank for sides that are internal cuts of a block decomposition."""
    nx, ny, nz = len(xs) - 1, len(ys) - 1, len(zs) - 1
    types = dict(top="wall", bottom="wall", left="wall", right="wall", back="wall", front="wall")
    if patch_types:
        types.update(patch_types)
    peer = peer or {}
    dx, dy, dz = np.diff(xs), np.diff(ys), np.diff(zs)
    xm, ym, zm = 0.5 * (xs[1:] + xs[:-1]), 0.5 * (ys[1:] + ys[:-1]), 0.5 * (zs[1:] + zs[:-1])
    n = nx * ny * nz
    cid = np.arange(n, dtype=np.int64)
    ci = cid % nx
    cj = (cid // nx) % ny
    ck = cid // (nx * ny)
    # inner faces: for every cell (ascending) its +x, +y, +z face when it exists
    valid = np.stack([ci < nx - 1, cj < ny - 1, ck < nz - 1], axis=1)          # [n,3]
    sel = np.nonzero(valid.ravel())[0]
    fc = sel // 3                                                             # owner cell of each inner face
    fd = (sel % 3).astype(np.int8)                                            # direction
    del valid, sel
    fi, fj, fk = ci[fc], cj[fc], ck[fc]
    stride = np.array([1, nx, nx * ny], dtype=np.int64)
    nb = fc + stride[fd]
    Fi = fc.size
    ax = np.where(fd == 0, dy[fj] * dz[fk], 0.0)
    ay = np.where(fd == 1, dx[fi] * dz[fk], 0.0)
    az = np.where(fd == 2, dx[fi] * dy[fj], 0.0)
    fx = np.where(fd == 0, xs[fi + 1], xm[fi])
    fy = np.where(fd == 1, ys[fj + 1], ym[fj])
    fz = np.where(fd == 2, zs[fk + 1], zm[fk])
    # distance owner centre -> face and face -> neighbour centre along the face normal
    dP = np.where(fd == 0, 0.5 * dx[fi], np.where(fd == 1, 0.5 * dy[fj], 0.5 * dz[fk]))
    dN = np.where(fd == 0, 0.5 * dx[np.minimum(fi + 1, nx - 1)], np.where(fd == 1, 0.5 * dy[np.minimum(fj + 1, ny - 1)], 0.5 * dz[np.minimum(fk + 1, nz - 1)]))
    facint = dP / (dP + dN)
    area = ax + ay + az
    Df = (area * area) / (area * (dP + dN))
    del fi, fj, fk, dP, dN

    def side(name):
        """owner cells (face order of hex_mesh), area vector, face centre, for one side of the box"""
        if name in ("top", "bottom"):
            i, k = np.meshgrid(np.arange(nx), np.arange(nz), indexing="ij")
            i, k = i.ravel(order="F"), k.ravel(order="F")
            j = np.full_like(i, ny - 1 if name == "top" else 0)
            s = 1.0 if name == "top" else -1.0
            return i + nx * (j + ny * k), (0 * dx[i], s * dx[i] * dz[k], 0 * dx[i]), (xm[i], np.full(i.size, ys[-1] if name == "top" else ys[0]), zm[k])
        if name in ("left", "right"):
            j, k = np.meshgrid(np.arange(ny), np.arange(nz), indexing="ij")
            j, k = j.ravel(order="F"), k.ravel(order="F")
            i = np.full_like(j, nx - 1 if name == "right" else 0)
            s = 1.0 if name == "right" else -1.0
            return i + nx * (j + ny * k), (s * dy[j] * dz[k], 0 * dy[j], 0 * dy[j]), (np.full(j.size, xs[-1] if name == "right" else xs[0]), ym[j], zm[k])
        i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        i, j = i.ravel(order="F"), j.ravel(order="F")
        k = np.full_like(i, nz - 1 if name == "front" else 0)
        s = 1.0 if name == "front" else -1.0
        return i + nx * (j + ny * k), (0 * dx[i], 0 * dx[i], s * dx[i] * dy[j]), (xm[i], ym[j], np.full(i.size, zs[-1] if name == "front" else zs[0]))

    order = ["top", "bottom", "left", "right", "back", "front"]
    phys = [s for s in order if s not in peer]
    proc = [s for s in order if s in peer]
    bown, bar, bcf, names, btypes, counts, peers = [], [], [], [], [], [], []
    for s in phys + proc:
        o, a, c = side(s)
        bown.append(o); bar.append(a); bcf.append(c)
        names.append(s if s not in peer else f"procBoundary_{s}_to{peer[s]}")
        btypes.append(BC_CODE[types[s]] if s not in peer else BC_PROCESS)
        counts.append(o.size)
        peers.append(peer.get(s, -1))
    starts = Fi + np.concatenate([[0], np.cumsum(counts)[:-1]])
    owner = (np.concatenate([fc] + bown) + 1).astype(np.int32)
    neighbour = (nb + 1).astype(np.int32)
    cat = lambda first, k: np.concatenate([first] + [b[k] for b in bar])  # noqa: E731
    arx, ary, arz = cat(ax, 0), cat(ay, 1), cat(az, 2)
    xf = np.concatenate([fx] + [c[0] for c in bcf])
    yf = np.concatenate([fy] + [c[1] for c in bcf])
    zf = np.concatenate([fz] + [c[2] for c in bcf])
    xc, yc, zc = xm[ci], ym[cj], zm[ck]
    vol = dx[ci] * dy[cj] * dz[ck]
    # patch_index of the matching patch on the peer: the peer lists its process patches in the same `order`, after its own
    # physical patches; the caller (block_partition_mesh) fills peer_patch because it knows the peer's patch table
    mesh = Mesh(numCells=n, numInnerFaces=Fi, numBoundaryFaces=int(sum(counts)), owner=owner, neighbour=neighbour,
                arx=arx, ary=ary, arz=arz, xf=xf, yf=yf, zf=zf, facint=facint, Df=Df, xc=xc, yc=yc, zc=zc, vol=vol,
                bcname=names, bctype=np.array(btypes, dtype=np.int32), nfaces=np.array(counts, dtype=np.int32),
                startFace=np.asarray(starts, dtype=np.int32), peer_rank=np.array(peers, dtype=np.int32),
                peer_patch=np.full(len(names), -1, dtype=np.int32))
    if BC_PERIODIC in btypes:          # periodic pairs: the twin is the opposite side, listed as 'empty' (faces of opposite sides share their order)
        opposite = dict(top="bottom", bottom="top", left="right", right="left", back="front", front="back")
        twin = np.full(len(names), -1, dtype=np.int32)
        for ib, nm in enumerate(names):
            if btypes[ib] == BC_PERIODIC:
                it = names.index(opposite[nm])
                assert btypes[it] == BC_EMPTY, "the twin of a periodic patch is listed as 'empty'"
                twin[ib] = starts[it]
        mesh.startFaceTwin = twin
    return mesh


def polyhedral_mesh_fast(nx: int, ny: Optional[int] = None, nz: Optional[int] = None, patch_types: Optional[Dict[str, str]] = None,
                         zrange: Optional[Tuple[int, int]] = None, peer: Optional[Dict[str, int]] = None) -> Mesh:
    """The brick-pattern polyhedral mesh of `polyhedral_mesh` (10-faced cells, pairs of hexahedra merged along x with the pair start staggered in
    y and z) WITHOUT distortion and without going through points and face-node lists: the hexahedral arrays of `hex_mesh_fast` are merged
    directly -- cell volume = sum, cell centre = volume-weighted mean (what the pyramid formula of geometry.f90:416-530 gives for a cell with planar
    faces), the face between the two halves dropped, face interpolation factors and Df re-evaluated from the new centres (geometry.f90:581-606,
    648-664) -- so that BASELINE config 5's size (~20 M polyhedra) is generated in about two minutes.  Same topology and face order as
    `polyhedral_mesh(nx, ny, nz, distort=0)`.
    zrange = (k0, k1) + peer = {"back": rank below, "front": rank above}: only the z-slab of hexahedron layers k0 <= k < k1 of that mesh, in the
    src-par layout (`polyhedral_partition_fast`): the pairs are merged along x and staggered with the GLOBAL layer index, so a z-cut separates no
    pair and the slabs together are exactly the global mesh; the cut faces become `process` patches, ordered alike on both sides."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    k0, k1 = zrange if zrange is not None else (0, nz)
    h = hex_mesh_fast(np.linspace(0.0, 1.0, nx + 1), np.linspace(0.0, 1.0, ny + 1), np.linspace(0.0, 1.0, nz + 1)[k0: k1 + 1], patch_types, peer)
    nh, Fi = h.numCells, h.numInnerFaces
    c = np.arange(nh, dtype=np.int64)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny) + k0
    head = np.where((i - (j + k) % 2) % 2 == 0, i, i - 1)
    head = np.where(head < 0, 0, head)
    key = head + nx * (j + ny * (k - k0))               # non-decreasing in c: hexes of a pair are consecutive
    del i, j, k, head, c
    first = np.empty(nh, dtype=bool)
    first[0] = True
    np.not_equal(key[1:], key[:-1], out=first[1:])
    poly = np.cumsum(first) - 1                          # polyhedron of every hexahedron
    ncell = int(poly[-1]) + 1
    del key, first
    vol = np.bincount(poly, weights=h.vol[:nh], minlength=ncell)
    xc = np.bincount(poly, weights=h.vol[:nh] * h.xc[:nh], minlength=ncell) / vol
    yc = np.bincount(poly, weights=h.vol[:nh] * h.yc[:nh], minlength=ncell) / vol
    zc = np.bincount(poly, weights=h.vol[:nh] * h.zc[:nh], minlength=ncell) / vol
    own = poly[h.owner.astype(np.int64) - 1]
    nb = poly[h.neighbour.astype(np.int64) - 1]
    keep = np.nonzero(own[:Fi] != nb)[0]
    assert (own[keep] < nb[keep]).all()
    order = keep[np.argsort(own[keep] * ncell + nb[keep], kind="stable")]     # OpenFOAM's upper-triangular order
    del keep
    fo, fn = own[order], nb[order]
    pair = fo * ncell + fn
    assert (pair[1:] > pair[:-1]).all(), "two polyhedra share more than one face"
    del pair
    nF_in = order.size
    removed = Fi - nF_in
    pick = lambda a: np.concatenate([a[order], a[Fi:]])  # noqa: E731
    arx, ary, arz, xf, yf, zf = (pick(a) for a in (h.arx, h.ary, h.arz, h.xf, h.yf, h.zf))
    dPx, dPy, dPz = xf[:nF_in] - xc[fo], yf[:nF_in] - yc[fo], zf[:nF_in] - zc[fo]
    dNx, dNy, dNz = xf[:nF_in] - xc[fn], yf[:nF_in] - yc[fn], zf[:nF_in] - zc[fn]
    djp = np.sqrt(dPx * dPx + dPy * dPy + dPz * dPz)
    djn = np.sqrt(dNx * dNx + dNy * dNy + dNz * dNz)
    facint = djp / (djp + djn)
    del dPx, dPy, dPz, dNx, dNy, dNz, djp, djn
    sx, sy, sz = arx[:nF_in], ary[:nF_in], arz[:nF_in]
    Df = (sx * sx + sy * sy + sz * sz) / (sx * (xc[fn] - xc[fo]) + sy * (yc[fn] - yc[fo]) + sz * (zc[fn] - zc[fo]))
    owner = (np.concatenate([fo, own[Fi:]]) + 1).astype(np.int32)
    neighbour = (fn + 1).astype(np.int32)
    return Mesh(numCells=ncell, numInnerFaces=nF_in, numBoundaryFaces=h.numBoundaryFaces, owner=owner, neighbour=neighbour,
                arx=arx, ary=ary, arz=arz, xf=xf, yf=yf, zf=zf, facint=facint, Df=Df, xc=xc, yc=yc, zc=zc, vol=vol,
                bcname=list(h.bcname), bctype=h.bctype.copy(), nfaces=h.nfaces.copy(), startFace=(h.startFace - removed).astype(np.int32),
                peer_rank=None if peer is None else h.peer_rank.copy(), peer_patch=None if peer is None else h.peer_patch.copy())


def polyhedral_partition_fast(nx: int, nranks: int, rank: int, ny: Optional[int] = None, nz: Optional[int] = None,
                              patch_types: Optional[Dict[str, str]] = None) -> Mesh:
    """Rank `rank`'s z-slab of `polyhedral_mesh_fast(nx, ny, nz)` generated directly in the src-par layout (no global mesh is ever built: BASELINE
    config 5 at ~20 M cells over 8 ranks needs 1/8 of the host memory and time per rank).  Ghost copies of xc, yc, zc, vol and the process-face
    facint / Df are filled by fcp_comm_init (src-par/geometry.f90:769-819)."""
    nz = nx if nz is None else nz
    k0, k1 = (nz * rank) // nranks, (nz * (rank + 1)) // nranks
    if k1 <= k0:
        raise ValueError("polyhedral_partition_fast: more ranks than layers")
    peer = {}
    if rank > 0:
        peer["back"] = rank - 1
    if rank < nranks - 1:
        peer["front"] = rank + 1
    return polyhedral_mesh_fast(nx, ny, nz, patch_types, (k0, k1), peer)


def block_dims(nranks: int) -> Tuple[int, int, int]:
    """z-slabs, like a decomposePar 'simple (1 1 n)' run of the src-par tree."""
    return (1, 1, nranks)


def block_partition_mesh(n: Tuple[int, int, int], dims: Tuple[int, int, int], rank: int, length: float = 1.0,
                         patch_types: Optional[Dict[str, str]] = None) -> Mesh:
    """Rank `rank`'s brick of a uniform nx x ny x nz box split in px x py x pz blocks (rank = bx + px*(by + py*bz)),
    generated directly in the src-par layout (no global mesh is ever built).  Ghost copies of xc,yc,zc,vol are filled by
    fcp_comm_init's exchange (src-par/geometry.f90:769-773), so they are left at zero here."""
    nx, ny, nz = n
    px, py, pz = dims
    bx, by, bz = rank % px, (rank // px) % py, rank // (px * py)

    def cut(nn, p, b):
        lo = (nn * b) // p
        hi = (nn * (b + 1)) // p
        return lo, hi
    (i0, i1), (j0, j1), (k0, k1) = cut(nx, px, bx), cut(ny, py, by), cut(nz, pz, bz)
    xs = np.linspace(0.0, length, nx + 1)[i0: i1 + 1]
    ys = np.linspace(0.0, length, ny + 1)[j0: j1 + 1]
    zs = np.linspace(0.0, length, nz + 1)[k0: k1 + 1]
    rk = lambda a, b, c: a + px * (b + py * c)  # noqa: E731
    peer = {}
    if bx > 0: peer["left"] = rk(bx - 1, by, bz)
    if bx < px - 1: peer["right"] = rk(bx + 1, by, bz)
    if by > 0: peer["bottom"] = rk(bx, by - 1, bz)
    if by < py - 1: peer["top"] = rk(bx, by + 1, bz)
    if bz > 0: peer["back"] = rk(bx, by, bz - 1)
    if bz < pz - 1: peer["front"] = rk(bx, by, bz + 1)
    m = hex_mesh_fast(xs, ys, zs, patch_types, peer)
    m.cell_offset = (i0, j0, k0)
    m.global_dims = (nx, ny, nz)
    return m
