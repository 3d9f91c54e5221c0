// host/poisson_app.cpp -- the reference's Poisson application (applications/Poisson/poisson.f90:50-104) written against
// the C++ mirror of its module API; every arithmetic step runs on the GPU through the C-ABI.
//   -lap(p) = 8 pi^2 sin(2 pi x) sin(2 pi y), p = 0 on the boundary; solved with iccg and dpcg; prints h and the L_inf error.
// build: g++ -O2 -std=c++17 host/poisson_app.cpp -o host/poisson_app -Lfreecappuccino-dev_b200/csrc -lfcp_b200 -Wl,-rpath,...
#include <algorithm>
#include <cstring>
#include "fcp_host.hpp"
using namespace fcp;

int main(int argc, char **argv) {
  const int n = argc > 1 ? std::atoi(argv[1]) : 32;
  const double pi = 3.14159265358979323846;
  // one layer of very thick cells: the front/back Dirichlet coefficients (area/distance = dx*dy/(lz/2)) vanish against the
  // in-plane ones (dy*lz/dx), i.e. the quasi-2D problem of the reference application
  box_mesh(n, n, 1, 1.0, 1.0, 1.0e3);
  create_CSR_matrix();
  using namespace geometry;
  using namespace sparse_matrix;
  // source term (poisson.f90:63)
  for (int i = 0; i < numCells; ++i) su[i] = 8 * pi * pi * std::sin(2 * pi * xc[i]) * std::sin(2 * pi * yc[i]) * vol[i];
  std::vector<dp> p(numTotal, 0.0), mu(numCells, -1.0), su0 = su;
  laplacian(mu, p);                       // poisson.f90:78  call laplacian(sv,p) with sv = -1
  double err[2];
  int k = 0;
  for (const char *solver : {"iccg", "dpcg"}) {
    std::fill(p.begin(), p.end(), 0.0);
    dp res0;
    csrsolve(solver, p, su, res0, 1000, 1e-30, 1e-13, "p");
    double e = 0;
    for (int i = 0; i < numCells; ++i) e = std::max(e, std::fabs(p[i] - std::sin(2 * pi * xc[i]) * std::sin(2 * pi * yc[i])));
    err[k++] = e;
    std::printf(" %s: h = %11.4e  Linf error = %11.4e\n", solver, 1.0 / n, e);   // poisson.f90:104
  }
  finalize();
  std::printf("POISSON_APP_DONE %d %.6e %.6e\n", n, err[0], err[1]);
  return 0;
}
