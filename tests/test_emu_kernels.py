"""CPU: the product's CUDA sources executed under the emulation of tests/emu (fibers for CUDA threads, processes for
ranks), checked against the oracle.  This is a LOGIC check of csrc/*.cu for containers without a GPU -- loop order,
reduction trees, tickets, barriers, the peer-memory protocol -- and nothing more: no performance meaning, not a product
path (the product only loads csrc/libfcp_b200.so).  The complete `-m gpu` suite runs the same way with
`FCP_TEST_EMU=1 python -m pytest tests -m gpu`; here only a slice that finishes in about a minute."""
import os
import subprocess
import sys

import pytest

from conftest import HAS_GPU, EMU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(HAS_GPU and not EMU, reason="a real GPU is present: the -m gpu suite is the check")


def _run(args, timeout=900, **env):
    e = dict(os.environ, FCP_TEST_EMU="1", **env)
    return subprocess.run([sys.executable] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=e)


def test_smoke_under_emulation():
    r = _run(["-c", "import sys; sys.path.insert(0, 'tests'); import emu_hook; emu_hook.activate(); import __graft_entry__ as g; g.smoke()"])
    assert r.returncode == 0 and "bit-identical to the oracle" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_parity_slice_under_emulation():
    r = _run(["-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_gpu_parity.py", "-k", "hex6 or tiny3 or poly_10faces or golden"])
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("switches", [dict(FCP_FACE_PF="2", FCP_FACE_CL="0", FCP_FACE_OG="0", FCP_FACE_OCC="3"),
                                      dict(FCP_FACE_PF="1,0,2,1", FCP_FACE_CL="1", FCP_FACE_OG="1", FCP_ASM_W="1")])
def test_face_kernel_switches_under_emulation(switches):
    """the non-default variants of the face kernels (three list stages + L2 prefetch, plain / compact lists, face-ordered / owner-ordered geometry,
    faces per assembly round), one value for all kernels or one per kernel: same bits as the oracle, on hexahedra and on the mesh whose cells have
    more faces than the list stages and the compact face-kind word hold"""
    r = _run(["-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_gpu_parity.py", "tests/test_gpu_scalar.py", "-k",
              "(hex6 or hex_many_faces) and (grad or assemble or calcp or pcorr or fvx)"], **switches)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_rows_f3_f4_slice_under_emulation():
    """periodic pairs, the scalar transport template with both k-epsilon and SST, the SGS models and the chained LES channel steps"""
    r = _run(["-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_gpu_scalar.py", "tests/test_gpu_zz_les_channel_loop.py", "-k",
              "channel_periodic or tiny3 or les_channel"])
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("comm", ["p2p", "nccl"])
def test_four_ranks_under_emulation(comm):
    """4 slab partitions (interior ranks own two process patches): halo plan, CUDA-IPC windows, the fused flag-in-data pushes,
    in-kernel rank-ordered all-reduce, start-up self-check -- bit-exact against the oracle's virtual ranks."""
    cmd = ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1", "--master-port",
           str(29610 + (1 if comm == "nccl" else 0)), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = _run(cmd, FCP_COMM=comm)
    assert r.returncode == 0 and r.stdout.count("MGPU_OK") == 4, r.stdout[-3000:] + r.stderr[-3000:]


def test_bench_control_flow_four_ranks_under_emulation():
    """bench.py's own main() on 4 ranks (block partition per rank, communicator, timed loops, statistics gather, JSON line);
    the numbers are meaningless here, the line's structure and the solver's convergence are what is checked."""
    import json
    cmd = ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1", "--master-port", "29620",
           os.path.join(ROOT, "tests", "emu_bench_worker.py"), "--gpus", "4", "--cells", "16", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"]
    r = _run(cmd)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, r.stdout[-3000:] + r.stderr[-3000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 4 and d["config"]["comm"] == "p2p" and d["gpu_launches"] > 0
    assert 0 < d["pcg"]["iters"] < 500 and d["pcg"]["resl"] < 1e-8 * d["pcg"]["res0"] * 1.0001
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}


def test_bench_falls_back_to_nccl_when_the_peer_memory_path_fails():
    """A peer-memory failure on ONE rank during the warm-up (injected) must make ALL ranks rebuild their contexts over NCCL and still print one
    valid line that says so -- round 1 lost its 4- and 8-GPU measurements to exactly this kind of failure."""
    import json
    cmd = ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29622",
           os.path.join(ROOT, "tests", "emu_bench_worker.py"), "--gpus", "2", "--cells", "16", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"]
    r = _run(cmd, FCP_EMU_BENCH_FAULT="1")
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, r.stdout[-3000:] + r.stderr[-3000:]
    d = json.loads(lines[0])
    assert d["config"]["comm"].startswith("nccl (fallback") and "repeating over NCCL" in r.stderr
    assert 0 < d["pcg"]["iters"] < 500 and d["pcg"]["resl"] < 1e-8 * d["pcg"]["res0"] * 1.0001


def test_late_round1_additions_under_emulation():
    """'gauss-seidel', the barrier-free IC(0)/ILU(0) sweeps and the full-size property tests at their emulation sizes (with the oracle comparison
    those sizes add): none of them has run on hardware yet."""
    r = _run(["-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_gpu_gauss_seidel.py", "tests/test_gpu_sweep_flags.py",
              "tests/test_zz_gpu_full_size.py"])
    assert r.returncode == 0 and " passed" in r.stdout and "skipped" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
