"""CPU: BASELINE config 2 through the oracle chain -- the reference's own examples/pitzDaily case (mesh fixture
tests/golden/pitzDaily_polymesh.npz, settings of input-simple.nml, fields of 0/): calcuvw (muscl, Venkatakrishnan, BiCGStab-ILU(0)),
calcp_simple (weighted, IC(0)-CG), realizable k-epsilon with wall functions.  The reference publishes no numbers for this case; what is
checked is what any correct implementation must show: global mass conservation through adjustMassFlow, positive k and epsilon, falling
outer residuals, and the separation bubble behind the step (Pitz & Daily measured reattachment near 7 step heights; run to the input
file's tolerance of 1e-6 -- 2000 iterations, 66 s, done once while writing this test -- the oracle chain reattaches at 8.2 H)."""
import numpy as np

import pitz_loop as P
from fcb200 import mesh as M

H_STEP = 0.0254


def test_pitz_daily_simple_iterations(orc):
    m = P.mesh()
    assert (m.numCells, m.numInnerFaces) == (12225, 24170)
    c = orc.Csr(m)
    f, flomas = P.initial_state(m, orc)
    up, sp = P.oracle_params(orc, orc.SUM_SEQ)
    a = np.zeros(c.nnz)
    n, Fi = m.numCells, m.numInnerFaces
    res = []
    for it in range(250):
        ur, pr, kr, er = P.oracle_iteration(orc, m, c, up, sp, f, a, flomas, orc.SUM_SEQ)
        res.append((ur[0].resor, ur[1].resor, pr.res0))
    res = np.array(res)
    assert np.all(np.isfinite(res)) and res[-1, 0] < 0.05 * res[1, 0] and res[-1, 2] < 0.05 * res[1, 2]
    assert f["te"][:n].min() > 0 and f["ed"][:n].min() > 0 and 10 < f["vis"][:n].max() / P.I["viscos"] < 5000
    # outlet mass flow equals the inlet mass flow (adjustMassFlow), and every cell is nearly divergence free
    out = m.patch_faces(m.bcname.index("out"))
    assert abs(f["flmass"][out].sum() - flomas) < 1e-12 * flomas
    net = np.zeros(n)
    np.add.at(net, m.owner[:Fi] - 1, f["flmass"][:Fi]); np.add.at(net, m.neighbour - 1, -f["flmass"][:Fi]); np.add.at(net, m.owner[Fi:] - 1, f["flmass"][Fi:])
    assert np.abs(net).max() < 0.02 * flomas
    # the recirculation bubble on the lower wall behind the step
    pf = m.patch_faces(m.bcname.index("lowerWall"))
    floor = (m.xf[pf] > 0) & (m.yf[pf] < -0.02) & (np.abs(m.ary[pf]) > 0.9 * np.sqrt(m.arx[pf] ** 2 + m.ary[pf] ** 2 + m.arz[pf] ** 2))
    x, u = m.xf[pf][floor], f["u"][m.owner[pf][floor] - 1]
    assert (u < 0).sum() > 10, "no reverse flow behind the step"
    xr = x[u < 0].max() / H_STEP
    assert 1.5 < xr < 10.0, xr                          # still growing after 250 iterations (8.2 H when converged)
