"""CPU: on-disk formats either side of the path: the serial native polyMesh and
the src-par `processorK/constant/polyMesh` tree (+ `process` file) written and read back."""
import os

import numpy as np
import pytest

import fcb200  # noqa: F401
from fcb200 import mesh as M


def test_native_polymesh_round_trip(tmp_path):
    m = M.cavity_mesh(5, distort=0.2)
    M.write_polymesh_native(m, str(tmp_path / "polyMesh"))
    r = M.read_polymesh_native(str(tmp_path / "polyMesh"))
    assert (r.numCells, r.numInnerFaces, r.numBoundaryFaces) == (m.numCells, m.numInnerFaces, m.numBoundaryFaces)
    assert np.array_equal(r.owner, m.owner) and np.array_equal(r.neighbour, m.neighbour)
    for k in ("arx", "ary", "arz", "xf", "yf", "zf", "vol", "facint", "Df"):
        np.testing.assert_allclose(getattr(r, k), getattr(m, k), rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("P", [2, 3])
def test_srcpar_partition_tree_round_trip(tmp_path, P):
    g = M.cavity_mesh(6, distort=0.15)
    parts = M.partition(g, M.slab_partition(g, P))
    M.write_partition_srcpar(g, parts, str(tmp_path))
    for r, part in enumerate(parts):
        d = tmp_path / f"processor{r}" / "constant" / "polyMesh"
        for f in ("points", "faces", "owner", "neighbour", "boundary", "process"):
            assert (d / f).exists()
        back = M.read_partition_srcpar(str(tmp_path), r)
        assert (back.numCells, back.numInnerFaces, back.numBoundaryFaces) == (part.numCells, part.numInnerFaces, part.numBoundaryFaces)
        assert np.array_equal(back.owner, part.owner) and np.array_equal(back.neighbour, part.neighbour)
        assert list(back.bctype) == list(part.bctype) and list(back.nfaces) == list(part.nfaces) and list(back.startFace) == list(part.startFace)
        assert np.array_equal(back.peer_rank, part.peer_rank)
        n = part.numCells
        # geometry recomputed from the written nodes = the partition's (area vectors of cut faces point OUT of the partition)
        for k in ("arx", "ary", "arz", "xf", "yf", "zf"):
            np.testing.assert_allclose(getattr(back, k), getattr(part, k), rtol=0, atol=1e-13)
        for k in ("xc", "yc", "zc", "vol"):
            np.testing.assert_allclose(getattr(back, k)[:n], getattr(part, k)[:n], rtol=0, atol=1e-13)
        np.testing.assert_allclose(back.facint, part.facint, rtol=0, atol=1e-12)
        np.testing.assert_allclose(back.Df, part.Df, rtol=1e-12)


def test_polyhedral_mesh_fast_equals_the_point_based_generator():
    """`polyhedral_mesh_fast` merges the hexahedral arrays directly (no points, no face-node lists); topology and face order must be identical to
    `polyhedral_mesh(distort=0)` and every geometric array equal to rounding."""
    import numpy as np
    from fcb200 import mesh as M
    for dims in ((6, 6, 6), (7, 5, 4), (8, 9, 3), (5, 4, 7)):
        a, b = M.polyhedral_mesh(*dims, distort=0.0), M.polyhedral_mesh_fast(*dims)
        assert (a.numCells, a.numInnerFaces, a.numBoundaryFaces) == (b.numCells, b.numInnerFaces, b.numBoundaryFaces)
        assert np.array_equal(a.owner, b.owner) and np.array_equal(a.neighbour, b.neighbour)
        assert np.array_equal(a.startFace, b.startFace) and np.array_equal(a.nfaces, b.nfaces) and np.array_equal(a.bctype, b.bctype)
        for k in ("arx", "ary", "arz", "xf", "yf", "zf", "facint", "Df", "vol"):
            np.testing.assert_allclose(getattr(a, k), getattr(b, k), rtol=1e-11, atol=1e-13, err_msg=k)
        for k in ("xc", "yc", "zc"):
            np.testing.assert_allclose(getattr(a, k)[: a.numCells], getattr(b, k)[: a.numCells], rtol=1e-11, atol=1e-13, err_msg=k)
        rows = np.bincount(np.concatenate([b.owner[: b.numInnerFaces], b.neighbour]) - 1, minlength=b.numCells)
        assert rows.max() == 10 or min(dims) < 3          # interior polyhedra have ten neighbours


def test_polyhedral_partition_fast_equals_the_partitioned_global_mesh():
    """`mesh.polyhedral_partition_fast` (bench.py --workload poly: every rank generates only its own z-slab of BASELINE config 5's mesh) against
    `mesh.partition` of the global `polyhedral_mesh_fast` mesh: same cells, inner faces, geometry and -- what the halo exchange relies on -- the
    same order of the process faces on both sides of every cut."""
    import numpy as np
    from fcb200 import mesh as M
    nx, ny, nz, P = 8, 6, 10, 3
    g = M.polyhedral_mesh_fast(nx, ny, nz)
    n = g.numCells
    layer = np.floor(g.zc[:n] * nz).astype(int)
    cuts = [(nz * r) // P for r in range(P + 1)]
    cr = np.zeros(n, np.int32)
    for r in range(P):
        cr[(layer >= cuts[r]) & (layer < cuts[r + 1])] = r
    ref = M.partition(g, cr)
    mine = [M.polyhedral_partition_fast(nx, P, r, ny, nz) for r in range(P)]
    for r in range(P):
        a, b = mine[r], ref[r]
        assert (a.numCells, a.numInnerFaces, a.numBoundaryFaces) == (b.numCells, b.numInnerFaces, b.numBoundaryFaces)
        Fi = a.numInnerFaces
        assert np.array_equal(a.owner[:Fi], b.owner[:Fi]) and np.array_equal(a.neighbour, b.neighbour)
        for k in ("xc", "yc", "zc", "vol"):
            np.testing.assert_allclose(getattr(a, k)[:a.numCells], getattr(b, k)[:b.numCells], rtol=0, atol=1e-15)
        for k in ("arx", "ary", "arz", "xf", "yf", "zf"):
            np.testing.assert_allclose(getattr(a, k)[:Fi], getattr(b, k)[:Fi], rtol=0, atol=1e-15)
        np.testing.assert_allclose(a.facint, b.facint[:Fi], rtol=1e-14)
        np.testing.assert_allclose(a.Df, b.Df[:Fi], rtol=1e-13)
        # boundary faces patch by patch (the two generators list the patches in different orders): owner cells, areas and centres
        def patches(msh):
            out = {}
            for ib in range(msh.numBoundaries):
                if msh.nfaces[ib] == 0:
                    continue
                pf = msh.patch_faces(ib)
                key = ("proc", int(msh.peer_rank[ib])) if msh.bctype[ib] == M.BC_PROCESS else ("phys", msh.bcname[ib])
                out[key] = (msh.owner[pf], msh.arx[pf], msh.ary[pf], msh.arz[pf], msh.xf[pf], msh.yf[pf], msh.zf[pf])
            return out
        pa, pb = patches(a), patches(b)
        assert set(pa) == set(pb), (sorted(pa), sorted(pb))
        for key in pa:
            assert np.array_equal(pa[key][0], pb[key][0]), key
            for x, y in zip(pa[key][1:], pb[key][1:]):
                np.testing.assert_allclose(x, y, rtol=0, atol=1e-15)
    # both sides of a cut list its faces in the same order: face i of rank r's patch towards q and face i of q's patch towards r are the same face
    for r in range(P - 1):
        a, b = mine[r], mine[r + 1]
        fa = a.patch_faces(int(np.nonzero(a.peer_rank == r + 1)[0][0]))
        fb = b.patch_faces(int(np.nonzero(b.peer_rank == r)[0][0]))
        assert fa.size == fb.size
        np.testing.assert_allclose(a.xf[fa], b.xf[fb], rtol=0, atol=1e-15)
        np.testing.assert_allclose(a.yf[fa], b.yf[fb], rtol=0, atol=1e-15)
        np.testing.assert_allclose(a.arz[fa], -b.arz[fb], rtol=0, atol=1e-15)


def test_facint_line_plane_variant_of_the_mpi_tree():
    """Quirk Q9: `mesh.facint_line_plane` (src-par/geometry.f90:780-819, the MATLAB-generated determinant quotient of find_intersection_point) against an
    independent restatement -- the parameter of the point where the line P -> N meets the plane through the face's first three vertices is
    n.(p1 - P) / n.(N - P), and |P j'| / |P N| is that parameter -- and against the serial tree's factor: equal on an orthogonal graded mesh (the
    face centre lies on the line), different on a skewed one."""
    import numpy as np
    from fcb200 import mesh as M
    for m, same in ((M.cavity_mesh(7, bump=0.35), True), (M.cavity_mesh(7, distort=0.25), False)):
        F = m.numInnerFaces
        lam = M.facint_line_plane(m)
        own, nb = m.owner[:F].astype(np.int64) - 1, m.neighbour.astype(np.int64) - 1
        P = np.stack([m.xc[own], m.yc[own], m.zc[own]], 1)
        N = np.stack([m.xc[nb], m.yc[nb], m.zc[nb]], 1)
        fn = m.face_nodes[:F].astype(np.int64) - 1
        p1, p2, p3 = (m.points[fn[:, k]] for k in range(3))
        nrm = np.cross(p2 - p1, p3 - p1)
        t = (nrm * (p1 - P)).sum(1) / (nrm * (N - P)).sum(1)
        assert ((t > 0) & (t < 1)).all()
        np.testing.assert_allclose(lam, t, rtol=1e-11, atol=0)
        d = np.abs(lam - m.facint).max()
        assert (d < 1e-12) if same else (d > 1e-4), d
