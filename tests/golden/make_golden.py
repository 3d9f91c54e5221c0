"""Generates the committed fixtures under tests/golden/ (run here, where /root/reference is mounted).

  python tests/golden/make_golden.py

* cavity20_polymesh.npz : topology of the reference's own test mesh test/testFieldOperations/polyMesh
  (OpenFOAM lid-driven cavity, 400 cells, 1640 faces, 760 inner; patches movingTop wall, fixedWalls wall,
  backForth symmetry), read with our OpenFOAM reader.  The reference's gradient smoke test runs on it
  (test/testFieldOperations/testFieldOperations.f90:137-160).
* spsolve_5x5.json : the two 5x5 systems and the 2-decimal solutions printed beside the solver output by
  test/test_linear_solvers_spsolve.f90:13-53,77-98,129-138 (matrix literals are default-real in the Fortran source,
  i.e. float32 values promoted to double; both forms are stored).
* pitzDaily_polymesh.npz : the reference's own examples/pitzDaily mesh (BASELINE config 2; 12 225 hexahedra, OpenFOAM polyMesh inside
  examples/pitzDaily/polyMesh.zip with the reference's simplified `boundary` file: in inlet, out outlet, upperWall / lowerWall wall, sides
  symmetry), read with our OpenFOAM reader.
* oracle_pins.npz : outputs of the CPU oracle on the 400-cell mesh for seeded inputs.  These are REGRESSION pins of
  the oracle itself (not reference outputs: the Fortran reference cannot be built in this image).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import fcb200  # noqa: E402
from fcb200 import mesh as M  # noqa: E402
from oracle import orc_py as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def pitz_daily():
    import tempfile
    import zipfile
    with tempfile.TemporaryDirectory() as tmp:
        zipfile.ZipFile(os.path.join(REF, "examples/pitzDaily/polyMesh.zip")).extractall(tmp)
        m = M.read_polymesh_openfoam(os.path.join(tmp, "polyMesh"))
    assert (m.numCells, m.numFaces, m.numInnerFaces) == (12225, 49180, 24170)
    M.save_mesh_npz(m, os.path.join(HERE, "pitzDaily_polymesh.npz"))


def main():
    pitz_daily()
    m = M.read_polymesh_openfoam(os.path.join(REF, "test/testFieldOperations/polyMesh"))
    assert (m.numCells, m.numFaces, m.numInnerFaces) == (400, 1640, 760)
    M.save_mesh_npz(m, os.path.join(HERE, "cavity20_polymesh.npz"))

    a1 = [6.80, -6.05, -0.45, 8.32, -9.67, -2.11, -3.30, 2.58, 2.71, -5.14, 5.66, 5.36, -2.70, 4.35, -7.26,
          5.97, -4.44, 0.27, -7.17, 6.08, 8.23, 1.08, 9.04, 2.14, -6.87]
    b1 = [4.02, 6.19, -8.22, -7.57, -3.03]
    x1 = [-0.80, -0.70, 0.59, 1.32, 0.57]
    a2 = [3.14, 0.17, -0.90, 1.65, -0.72, 0.17, 0.79, 0.83, -0.65, 0.28, -0.90, 0.83, 4.53, -3.70, 1.60,
          1.65, -0.65, -3.70, 5.32, -1.37, -0.72, 0.28, 1.60, -1.37, 1.98]
    b2 = [-7.29, 9.25, 5.99, -1.94, -8.30]
    x2 = [-6.02, 15.62, 3.02, 3.25, -8.78]
    f32 = lambda v: [float(np.float32(t)) for t in v]  # noqa: E731
    js = dict(source="test/test_linear_solvers_spsolve.f90:13-53,77-98,129-138",
              ja=[1, 2, 3, 4, 5] * 5, ioffset=[1, 6, 11, 16, 21, 26], diag=[1, 7, 13, 19, 25],
              itr_max=5, tol_rel=1e-7, tol_abs=1e-10,
              nonsymmetric=dict(a=a1, b=b1, x=x1, a_f32=f32(a1), b_f32=f32(b1), solvers=["bicgstab"]),
              spd=dict(a=a2, b=b2, x=x2, a_f32=f32(a2), b_f32=f32(b2), solvers=["iccg", "dpcg"]))
    with open(os.path.join(HERE, "spsolve_5x5.json"), "w") as fh:
        json.dump(js, fh, indent=1)

    # oracle regression pins on the 400-cell mesh
    rng = np.random.default_rng(12345)
    csr = O.Csr(m)
    phi = m.boundary_values_of(lambda x, y, z: np.sin(3 * x) * np.cos(2 * y) + 0.3 * z) + 0.01 * rng.standard_normal(m.numTotal)
    pins = dict(phi=phi, ia=csr.ia, ja=csr.ja, diag=csr.diag, kpn=csr.icell_jcell, knp=csr.jcell_icell)
    pins["grad_gauss"] = O.grad_gauss(m, phi)
    for w, nm in ((False, "lsq"), (True, "lsq_dm")):
        D = O.create_matrix_lsq(m, w)
        pins["Dmat_" + nm] = D
        pins["grad_" + nm] = O.grad_lsq(m, w, D, phi)
    mu = np.ones(m.numTotal)
    su = np.zeros(m.numCells)
    pins["lap_a"] = O.laplacian(m, csr, mu, phi, su)
    pins["lap_su"] = su
    np.savez_compressed(os.path.join(HERE, "oracle_pins.npz"), **pins)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
