"""CPU, world_size 2, gloo: the host-side logic of the N>1 path -- src-par layout produced by the partitioners, peer
tables, face ordering on both sides of a process patch -- checked by doing the halo exchange of src-par/exchange.f90 with
torch.distributed send/recv between two real processes and comparing the ghost slots with the global field."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    import fcb200  # noqa: F401
    from fcb200 import mesh as M
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 8
        if mode == "generic":
            g = M.cavity_mesh(n, distort=0.1)
            me = M.partition(g, M.slab_partition(g, world))[rank]
            gid = me.cell_global
        else:
            me = M.block_partition_mesh((n, n, n), M.block_dims(world), rank)
            i0, j0, k0 = me.cell_offset
            nzl = me.numCells // (n * n)
            kk, jj, ii = np.meshgrid(np.arange(nzl) + k0, np.arange(n) + j0, np.arange(n) + i0, indexing="ij")
            gid = (ii + n * (jj + n * kk)).ravel()
        gphi = np.random.default_rng(7).standard_normal(n ** 3)
        phi = np.zeros(me.numTotal)
        phi[: me.numCells] = gphi[gid]
        # exchange.f90:48-127 : buffer(ipro) = phi(owner(iface)) ; sendrecv with neighbProcNo ; unpack into the boundary slots
        reqs, recvs = [], []
        for ib in range(me.numBoundaries):
            if me.bctype[ib] != M.BC_PROCESS:
                continue
            pf = me.patch_faces(ib)
            send = torch.from_numpy(phi[me.owner[pf] - 1].copy())
            recv = torch.empty_like(send)
            reqs.append(dist.isend(send, int(me.peer_rank[ib])))
            reqs.append(dist.irecv(recv, int(me.peer_rank[ib])))
            recvs.append((pf, recv))
        for r in reqs:
            r.wait()
        for pf, recv in recvs:
            phi[me.numCells + pf - me.numInnerFaces] = recv.numpy()
        # check: ghost value of a process face = global value of the cell on the other side, located geometrically
        ok = True
        h = 1.0 / n
        for pf, _ in recvs:
            if mode == "generic":
                own0 = g.owner.astype(np.int64) - 1; nb0 = g.neighbour.astype(np.int64) - 1
                gf = me.face_global[pf]
                mine = me.cell_global[me.owner[pf] - 1]
                other = np.where(own0[gf] == mine, nb0[gf], own0[gf])
            else:
                # the cell across a z-cut face sits one h further along the outward normal
                s = np.sign(me.arz[pf])
                xo, yo, zo = me.xf[pf], me.yf[pf], me.zf[pf] + 0.5 * h * s
                other = (np.floor(xo / h) + n * (np.floor(yo / h) + n * np.floor(zo / h))).astype(np.int64)
            ok &= np.array_equal(phi[me.numCells + pf - me.numInnerFaces], gphi[other])
        # global_sum (global_sum_mpi.f90): rank-ordered add of the per-rank partials
        part = torch.tensor([float(phi[: me.numCells].sum())], dtype=torch.float64)
        gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, part)
        s = gathered[0].item()
        for t in gathered[1:]:
            s = s + t.item()
        ok &= abs(s - gphi.sum()) < 1e-10
        q.put((rank, bool(ok), me.npro))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["generic", "block"])
def test_two_rank_halo_exchange(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + (0 if mode == "generic" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(npro == 64 for _, _, npro in res), res
