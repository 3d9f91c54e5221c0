"""TEST INFRASTRUCTURE ONLY: runs bench.py's own main() against the emulation library (tests/emu) so that its control flow
-- per-rank block partition, communicator set-up, timed loops, the cross-rank reduction of the statistics and the JSON line --
can be checked here for 1..8 ranks on a small mesh.  The numbers it prints are meaningless (CPU emulation); bench.py itself
refuses to run without a CUDA device."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import emu_hook  # noqa: E402

assert emu_hook.wanted() and not torch.cuda.is_available(), "emulation worker: set FCP_TEST_EMU=1 on a machine without a GPU"
emu_hook.activate()
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None

import bench  # noqa: E402


def _unpinned_copy(a):
    v = np.array(a, dtype=np.float64, copy=True)
    return None, v


bench.pinned_copy = _unpinned_copy
if os.environ.get("FCP_EMU_BENCH_FAULT"):      # fault injection: the first solve over the peer-memory path fails on rank 1 like a timed-out wait would
    from fcb200 import lib as _L
    _orig = _L.Context.calcp_simple

    def _faulty(self, *a, **k):
        if self.comm_mode() == "p2p" and int(os.environ.get("RANK", "0")) == 1:
            raise _L.FcpError("injected: peer-memory wait timed out")
        return _orig(self, *a, **k)
    _L.Context.calcp_simple = _faulty
if len(sys.argv) > 1 and sys.argv[1] == "rows":          # tools/bench_rows.py (per-kernel timings of the rows beyond the headline step)
    sys.argv.pop(1)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_rows
    bench_rows.main()
else:
    bench.main()
