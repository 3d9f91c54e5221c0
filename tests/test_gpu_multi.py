"""GPU (>= 2 B200): partitioned run through NCCL, one rank per GPU, against the oracle's virtual-rank drivers.
Launches tests/mgpu_worker.py under torch.distributed.run; skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    from conftest import EMU
    if EMU:          # FCP_TEST_EMU=1: ranks are processes sharing the emulated device memory (tests/emu)
        return 8
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except (OSError, subprocess.TimeoutExpired):
        return 0


@pytest.mark.parametrize("part", ["slab", "brick", "poly", "periodic", "inout"])
@pytest.mark.parametrize("comm", ["p2p", "nccl"])
@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_partitioned_parity(nranks, comm, part):
    """comm = p2p: CUDA-IPC windows, halo pushes and rank-ordered reductions fused into the Krylov kernels (NVLink stores);
    comm = nccl: the send/recv + all-gather path.  Both must give the oracle's bits."""
    if part == "brick" and nranks < 4:
        pytest.skip("brick partitions start at 4 ranks")
    if part == "inout" and nranks > 4:
        pytest.skip("the inlet/outlet channel fixture has 5 cell layers in z: 2 or 4 slabs")
    if part == "periodic" and nranks > 4:
        pytest.skip("the periodic channel fixture has 8 cell layers across: 2 or 4 y-slabs")
    if _ngpus() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + nranks + (10 if comm == "nccl" else 0) + (20 if part == "brick" else 40 if part == "poly" else 60 if part == "periodic" else 80 if part == "inout" else 0)), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, FCP_COMM=comm, FCP_TEST_PART="slab" if part in ("poly", "periodic", "inout") else part, FCP_TEST_MESH=part if part in ("poly", "periodic", "inout") else "hex"))
    ok = r.stdout.count("MGPU_OK")   # (two ranks may print on one line)
    assert r.returncode == 0 and ok == nranks, r.stdout[-3000:] + r.stderr[-3000:]
