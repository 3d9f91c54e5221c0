"""CPU: the whole oracle chain (calcuvw + calcp_simple restatements, BiCGStab-ILU(0), IC(0)-CG) run as the reference's
SIMPLE loop on its own examples/cavity case converges to the lid-driven-cavity solution of Ghia, Ghia & Shin (1982): an
external known answer for rows a4-a7, a15-a17 and f1 taken together."""
import numpy as np

import simple_loop as S


def test_cavity_re100_matches_ghia(orc):
    m = S.cavity(39)
    c = orc.Csr(m)
    f = S.initial_state(m)
    prm = S.oracle_params(orc, orc.SUM_SEQ)
    a = np.zeros(c.nnz)
    dP = np.zeros((m.numTotal, 3))
    res = []
    for it in range(400):
        ureps, prep = S.oracle_iteration(orc, m, c, prm, f, a, dP, orc.SUM_SEQ)
        res.append(ureps[0].resor)
    assert res[-1] < 1e-4 * max(res[:5]), "outer iterations did not converge"
    y, u = S.centreline_u(m, f["u"])
    ui = np.interp(S.GHIA_Y, y, u)
    err = np.abs(ui - S.GHIA_U)
    assert err.max() < 0.005, (err.max(), list(zip(S.GHIA_Y, ui, S.GHIA_U)))
    # discrete continuity: the corrected face fluxes of the converged field are divergence free
    Fi = m.numInnerFaces
    div = np.zeros(m.numCells)
    np.add.at(div, m.owner[:Fi] - 1, f["flmass"][:Fi]); np.add.at(div, m.neighbour - 1, -f["flmass"][:Fi])
    assert np.abs(div).max() < 5e-3 * np.abs(f["flmass"]).max()
    assert abs(f["w"][: m.numCells]).max() < 1e-12      # 2-D case: nothing drives w
