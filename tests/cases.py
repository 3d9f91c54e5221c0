"""Shared, seeded test inputs (deterministic synthetic fields)."""
import os

import numpy as np

import fcb200
from fcb200 import mesh as M

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_mesh():
    """The reference's own 400-cell test mesh (test/testFieldOperations/polyMesh), from the committed fixture."""
    return M.load_mesh_npz(os.path.join(GOLDEN, "cavity20_polymesh.npz"))


def meshes():
    """name -> mesh for the parity matrix: orthogonal, graded, distorted (non-orthogonal), 2-D slab, mixed patches."""
    out = {}
    out["ref400"] = golden_mesh()
    out["hex6"] = M.cavity_mesh(6)
    out["hex12_graded"] = M.cavity_mesh(12, bump=0.3)
    out["hex10_distorted"] = M.cavity_mesh(10, distort=0.25)
    out["slab39_empty"] = M.cavity_mesh(39, nz=1, bump=0.2)
    xs = np.linspace(0, 2.0, 15); ys = np.linspace(0, 1.0, 9); zs = np.linspace(0, 0.5, 6)
    out["channel_inout"] = M.hex_mesh(xs, ys, zs, dict(left="inlet", right="outlet", back="symmetry", front="symmetry"), distort=0.15)
    out["channel_pressure"] = M.hex_mesh(xs, ys, zs, dict(left="inlet", right="pressure", back="empty", front="empty"), distort=0.1)
    out["poly_10faces"] = M.polyhedral_mesh(8, 6, 5, distort=0.15, patch_types=dict(left="inlet", right="outlet"))   # 10-faced cells + hexes
    out["hex_many_faces"] = many_faced_mesh()
    out["tiny3"] = M.hex_mesh(np.linspace(0, 1, 4), np.linspace(0, 1, 3), np.linspace(0, 1, 2))   # 3x2x1 = 6 cells (< one warp)
    # periodic pairs (row f3): a channel395-style box, periodic in x and z ('right'/'front' periodic, their twins 'left'/'back' listed as
    # empty, examples/channel395/README.md), walls in y, graded in y, distorted inside; and a one-pair inlet/outlet-free duct with symmetry sides
    out["channel_periodic"] = periodic_channel()
    out["duct_periodic_x"] = M.hex_mesh(np.linspace(0, 1.5, 8), M.bump_nodes(5, 0.4), np.linspace(0, 0.6, 5),
                                        dict(left="empty", right="periodic", back="symmetry", front="symmetry"), distort=0.1)
    out["duct_periodic_first"] = M.hex_mesh(np.linspace(0, 1.5, 7), np.linspace(0, 1, 5), np.linspace(0, 0.6, 4),
                                            dict(left="periodic", right="empty", top="symmetry"), distort=0.1)   # the periodic patch precedes its twin
    return out


def split_boundary_quads(m, patch, strips):
    """Cut boundary quads of `patch` into strips (strips: owner cell, 0-based -> number of strips): the cell keeps its shape and gains faces.
    Ragged face lists for the kernels: longer than the shared-memory list stages, longer than the compact face-kind word describes."""
    F = m.numInnerFaces
    pts = [m.points]
    npts = m.points.shape[0]
    nodes, owner, patches = [m.face_nodes[:F, :4]], [m.owner[:F]], []
    start = F
    for ib, name in enumerate(m.bcname):
        cnt = 0
        for f in m.patch_faces(ib):
            a, b, c, d = (int(x) - 1 for x in m.face_nodes[f, :4])
            own = int(m.owner[f])
            s = strips.get(own - 1, 1) if name == patch else 1
            if s == 1:
                nodes.append(m.face_nodes[f:f + 1, :4]); owner.append([own]); cnt += 1
                continue
            t = np.linspace(0.0, 1.0, s + 1)[1:-1, None]
            P = m.points[a] + t * (m.points[b] - m.points[a])
            Q = m.points[d] + t * (m.points[c] - m.points[d])
            pts += [P, Q]
            pid = [a] + list(range(npts, npts + s - 1)) + [b]
            qid = [d] + list(range(npts + s - 1, npts + 2 * (s - 1))) + [c]
            npts += 2 * (s - 1)
            for i in range(s):
                nodes.append(np.array([[pid[i] + 1, pid[i + 1] + 1, qid[i + 1] + 1, qid[i] + 1]], dtype=np.int32)); owner.append([own]); cnt += 1
        patches.append((name, M.BC_NAMES[int(m.bctype[ib])], cnt, start))
        start += cnt
    face_nodes = np.ascontiguousarray(np.concatenate(nodes), dtype=np.int32)
    owner = np.concatenate([np.asarray(o, dtype=np.int32) for o in owner])
    return M.mesh_from_topology(np.concatenate(pts), face_nodes, np.full(owner.shape[0], 4, dtype=np.int32), owner, m.neighbour, m.numCells, patches)


def many_faced_mesh():
    """5 x 4 x 3 hexahedra with inlet and outlet; the wall quads of three top cells are cut into 2, 7 and 12 strips: cells with 7, 12 (longer than
    the 10-entry list stage) and 17 faces (more than the 14 the compact face-kind word holds)."""
    m = M.hex_mesh(np.linspace(0, 1.0, 6), np.linspace(0, 0.8, 5), np.linspace(0, 0.6, 4), dict(left="inlet", right="outlet"))
    top = [int(c) - 1 for c in m.owner[m.patch_faces(m.bcname.index("top"))]]
    return split_boundary_quads(m, "top", {top[1]: 2, top[3]: 7, top[7]: 12})


def periodic_channel(nx=9, ny=7, nz=6, distort=0.2):
    return M.hex_mesh(np.linspace(0, 2.0, nx + 1), M.bump_nodes(ny, 0.3), np.linspace(0, 1.0, nz + 1),
                      dict(left="empty", right="periodic", back="empty", front="periodic"), distort=distort)


def fields(m, seed=12345):
    """u,v,w,p,pp,den,apu,apv,apw of length numTotal, smooth + a little seeded noise."""
    rng = np.random.default_rng(seed)
    pi = np.pi
    f = {}
    f["u"] = m.boundary_values_of(lambda x, y, z: np.sin(pi * x) * np.cos(pi * y) + 0.1 * z)
    f["v"] = m.boundary_values_of(lambda x, y, z: -np.cos(pi * x) * np.sin(pi * y) + 0.05 * np.sin(pi * z))
    f["w"] = m.boundary_values_of(lambda x, y, z: 0.1 * np.sin(pi * x) * np.sin(pi * z))
    f["p"] = m.boundary_values_of(lambda x, y, z: 0.25 * np.cos(2 * pi * x) * np.cos(2 * pi * y) + 0.1 * z * z)
    f["pp"] = 1e-3 * m.boundary_values_of(lambda x, y, z: np.sin(2 * pi * x) * np.sin(pi * y) * np.cos(pi * z))
    f["den"] = m.boundary_values_of(lambda x, y, z: 1.0 + 0.05 * np.sin(pi * x * y))
    h = m.vol[: m.numCells].mean() ** (1.0 / 3.0)
    base = 0.8 / (6.0 * 0.01 * h)
    f["apu"] = base * m.boundary_values_of(lambda x, y, z: 1.0 + 0.1 * np.sin(3 * x + y))
    f["apv"] = base * m.boundary_values_of(lambda x, y, z: 1.0 + 0.1 * np.cos(2 * y + z))
    f["apw"] = base * m.boundary_values_of(lambda x, y, z: 1.0 + 0.1 * np.sin(x + 2 * z))
    for k in f:
        f[k] = f[k] + 1e-3 * rng.standard_normal(m.numTotal) * (np.abs(f[k]).mean() + 1e-30)
    return f


def poisson_system(m, orc):
    """-lap(p) = 8 pi^2 sin(2 pi x) sin(2 pi y), Dirichlet 0 (applications/Poisson/poisson.f90:63-104), via the oracle's laplacian."""
    csr = orc.Csr(m)
    n = m.numCells
    pi = np.pi
    su = 8 * pi * pi * np.sin(2 * pi * m.xc[:n]) * np.sin(2 * pi * m.yc[:n]) * m.vol[:n]
    mu = -np.ones(m.numTotal)
    phi = np.zeros(m.numTotal)
    a = orc.laplacian(m, csr, mu, phi, su)
    return csr, a, su
