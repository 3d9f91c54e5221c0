"""CPU: the halo layout of the communication layer (csrc/comm.cu: halo_plan_build, exported host-only as fcp_comm_plan)
for ranks with ONE, TWO and MORE neighbours.  The device kernels index their buffers with exactly these arrays, so an
exchange emulated in numpy from the plans of all ranks must put, into every ghost slot, the value of the cell across the
face -- on z-slabs (interior ranks have two process patches) and on 2x2x2 / 3x1x2 brick partitions (3 neighbours).
(Round 1 shipped a layout bug that only showed with more than one process patch per rank; this test pins it.)"""
import ctypes as C

import numpy as np
import pytest

import fcb200  # noqa: F401
from fcb200 import lib as L
from fcb200 import mesh as M

CHUNK = 2048


def plan(m, rank, nranks):
    lib = L.lib()
    lib.fcp_comm_plan.argtypes = [C.POINTER(L.MeshDesc), L._pi, C.c_int, C.c_int] + [L._pi] * 10
    keep = {}
    md = L.MeshDesc()
    md.numCells, md.numInnerFaces, md.numBoundaryFaces, md.numBoundaries = m.numCells, m.numInnerFaces, m.numBoundaryFaces, m.numBoundaries
    for n in ("owner", "neighbour", "bctype", "nfaces", "startFace"):
        keep[n] = np.ascontiguousarray(getattr(m, n), dtype=np.int32)
        setattr(md, n, L._i(keep[n]))
    for n in ("arx", "ary", "arz", "xf", "yf", "zf", "facint", "Df", "xc", "yc", "zc", "vol"):
        keep[n] = np.ascontiguousarray(getattr(m, n), dtype=np.float64)
        setattr(md, n, L._d(keep[n]))
    npro = m.npro
    nch = max((m.numCells + CHUNK - 1) // CHUNK, 1)
    nb = m.numBoundaries
    out = dict(npatch=np.zeros(1, np.int32), peer=np.zeros(nb, np.int32), off=np.zeros(nb, np.int32), cnt=np.zeros(nb, np.int32),
               cell=np.zeros(max(npro, 1), np.int32), slot=np.zeros(max(npro, 1), np.int32), cptr=np.zeros(nch + 1, np.int32),
               cface=np.zeros(max(npro, 1), np.int32), order=np.zeros(nch, np.int32), gord=np.full(max(m.numBoundaryFaces, 1), -7, np.int32))
    pr = np.ascontiguousarray(m.peer_rank, np.int32)
    rc = lib.fcp_comm_plan(C.byref(md), L._i(pr), rank, nranks, *[L._i(out[k]) for k in ("npatch", "peer", "off", "cnt", "cell", "slot", "cptr", "cface", "order", "gord")])
    assert rc == 0, L.lib().fcp_last_error()
    k = int(out["npatch"][0])
    for key in ("peer", "off", "cnt"):
        out[key] = out[key][:k]
    out["npro"] = npro
    return out


@pytest.mark.parametrize("dims,n", [((1, 1, 2), (6, 5, 8)), ((1, 1, 4), (6, 5, 8)), ((1, 1, 8), (4, 4, 16)), ((2, 2, 2), (6, 6, 6)), ((3, 1, 2), (9, 4, 6)),
                                    ((1, 1, 4), (40, 40, 8))])
def test_halo_plan_and_emulated_exchange(dims, n):
    P = dims[0] * dims[1] * dims[2]
    meshes = [M.block_partition_mesh(n, dims, r) for r in range(P)]
    plans = [plan(meshes[r], r, P) for r in range(P)]
    f = lambda x, y, z: 1.0 + 2.0 * x - 3.0 * y + 5.0 * z + x * y     # noqa: E731
    for r, (m, pl) in enumerate(zip(meshes, plans)):
        proc = [ib for ib in range(m.numBoundaries) if m.bctype[ib] == M.BC_PROCESS]
        assert list(pl["peer"]) == [int(m.peer_rank[ib]) for ib in proc]
        assert list(pl["cnt"]) == [int(m.nfaces[ib]) for ib in proc]
        assert list(pl["off"]) == list(np.concatenate([[0], np.cumsum(pl["cnt"])[:-1]]).astype(int)), "patch offsets must be cumulative"
        assert len(set(pl["peer"])) == len(pl["peer"]), "one patch per neighbour in a brick partition"
        npro = pl["npro"]
        # per-face arrays
        faces = np.concatenate([m.patch_faces(ib) for ib in proc]) if proc else np.zeros(0, int)
        assert np.array_equal(pl["slot"][:npro], m.numCells + faces - m.numInnerFaces)
        assert np.array_equal(pl["cell"][:npro], m.owner[faces] - 1)
        # chunk grouping, launch order, ghost map
        cptr, cface = pl["cptr"], pl["cface"][:npro]
        assert cptr[0] == 0 and cptr[-1] == npro and sorted(cface) == list(range(npro))
        for k in range(len(cptr) - 1):
            assert np.all(pl["cell"][cface[cptr[k]:cptr[k + 1]]] // CHUNK == k)
        order = pl["order"]
        chunks = order & 0x7FFFFFFF
        halo = order < 0
        assert sorted(chunks) == list(range(len(order)))
        assert np.array_equal(halo, np.diff(cptr)[chunks] > 0) and (not halo.any() or not halo[np.argmin(halo):].any() or halo.all()), "halo chunks first"
        g = pl["gord"][: m.numBoundaryFaces]
        assert np.array_equal(np.flatnonzero(g >= 0), np.sort(pl["slot"][:npro] - m.numCells)) and np.array_equal(g[pl["slot"][:npro] - m.numCells], np.arange(npro))
    # emulated exchange: rank A face i -> rank B = peer, ordinal off_B[patch toward A] + t (what nccl_swap_face_ints / the NCCL group deliver)
    fields = []
    for m in meshes:
        phi = np.full(m.numTotal, np.nan)
        phi[: m.numCells] = f(m.xc[: m.numCells], m.yc[: m.numCells], m.zc[: m.numCells])
        fields.append(phi)
    for a, (ma, pa) in enumerate(zip(meshes, plans)):
        for j, b in enumerate(pa["peer"]):
            pb = plans[b]
            jb = list(pb["peer"]).index(a)
            assert pb["cnt"][jb] == pa["cnt"][j]
            t = np.arange(pa["cnt"][j])
            src_cell = pa["cell"][pa["off"][j] + t]
            rord = pb["off"][jb] + t
            fields[b][pb["slot"][rord]] = fields[a][src_cell]
            # LL addressing of the fused push: B reads its own LL slot ghost_ord_B[slot - n], which must be the ordinal A wrote to
            assert np.array_equal(pb["gord"][pb["slot"][rord] - meshes[b].numCells], rord)
    for b, (m, pl) in enumerate(zip(meshes, plans)):
        npro = pl["npro"]
        slots = pl["slot"][:npro]
        faces = slots - m.numCells + m.numInnerFaces
        own = m.owner[faces] - 1
        # uniform mesh: the cell across the face is the mirror image of the owner cell through the face centre
        xo, yo, zo = 2 * m.xf[faces] - m.xc[own], 2 * m.yf[faces] - m.yc[own], 2 * m.zf[faces] - m.zc[own]
        np.testing.assert_allclose(fields[b][slots], f(xo, yo, zo), rtol=0, atol=1e-12)
        # physical boundary slots untouched
        rest = np.setdiff1d(np.arange(m.numCells, m.numTotal), slots)
        assert np.all(np.isnan(fields[b][rest]))


@pytest.mark.parametrize("P", [2, 3, 4, 6])
def test_oracle_virtual_ranks_with_interior_partitions(orc, P):
    """The checker of the multi-GPU parity test (orc_exchange / orc_dpcg_par, src-par/dpcg.f90 + exchange.f90) on slab
    partitions whose interior ranks have TWO process patches: same iteration count and solution as the serial DPCG."""
    import cases
    g = M.cavity_mesh(12, distort=0.2)
    parts = M.partition(g, M.slab_partition(g, P))
    gcsr, ga, gsu = cases.poisson_system(g, orc)
    csrs = [orc.Csr(p) for p in parts]
    loc = [M.localize_matrix(g, gcsr, ga, p, c) for p, c in zip(parts, csrs)]
    fi_l = [np.zeros(p.numTotal) for p in parts]
    rhs_l = [gsu[p.cell_global].copy() for p in parts]
    rep = orc.dpcg_par(parts, csrs, [l[0] for l in loc], [l[1] for l in loc], fi_l, rhs_l, 500, 1e-30, 1e-10, orc.SUM_SEQ)
    xg = np.zeros(g.numCells)
    rg = orc.solve(orc.DPCG, gcsr.ia, gcsr.ja, ga, gcsr.diag, xg, gsu, 500, 1e-30, 1e-10)
    assert abs(rep.iters - rg.iters) <= 1
    for r, p in enumerate(parts):
        np.testing.assert_allclose(fi_l[r][: p.numCells], xg[p.cell_global], rtol=0, atol=1e-12)
    # the plan of the communication layer agrees with the partitioner's patch table on these (unstructured-path) partitions too
    for r, p in enumerate(parts):
        pl = plan(p, r, P)
        assert list(pl["off"]) == list(np.concatenate([[0], np.cumsum(pl["cnt"])[:-1]]).astype(int))
        assert list(pl["peer"]) == [int(x) for x in p.peer_rank[p.peer_rank >= 0]]
