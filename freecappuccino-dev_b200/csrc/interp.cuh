// interp.cuh -- the face_value family of interpolation.f90:28-650 as one device function, shared by the momentum predictor
// (fvm_uvw.cu) and the scalar transport template (fvm_scalar.cu)
#pragma once
#include "fcp_internal.h"

#define TINY30 ((double)1e-30f)                 // interpolation.f90:607  `1e-30` (default-real literal)
#define S13 ((double)(1.f / 3.f))               // `1./3.`
#define S23 ((double)(2.f / 3.f))               // `2./3.`

// face_value(ijp, ijn, xf, yf, zf, lambda, u, dUdxi, scheme), interpolation.f90:28-113: "p" is the first cell argument
__device__ __forceinline__ double face_value_dev(int scheme, double up, double un, const double (&gp)[3], const double (&gn)[3], double xp, double yp,
                                                 double zp, double xn, double yn, double zn, double xf, double yf, double zf, double lambda) {
  if (scheme == 0) return up + (un - up) * lambda;
  if (scheme == 1 || scheme == 3) {
    const double gc = gp[0] * (xf - xp) + gp[1] * (yf - yp) + gp[2] * (zf - zp) + gn[0] * (xf - xn) + gn[1] * (yf - yn) + gn[2] * (zf - zn);
    const double vf_central = 0.5 * (up + un + gc);
    if (scheme == 1) return vf_central;
    const double theta = S23;
    const double gu = gp[0] * (xf - xp) + gp[1] * (yf - yp) + gp[2] * (zf - zp);
    return theta * vf_central + (1.0 - theta) * (up + gu);
  }
  if (scheme == 2) {
    const double gu = gp[0] * (xf - xp) + gp[1] * (yf - yp) + gp[2] * (zf - zp);
    return up + gu;
  }
  const double fxp = 1.0 - lambda;
  const double xpn = xn - xp, ypn = yn - yp, zpn = zn - zp;
  const double r = (2 * gp[0] * xpn + 2 * gp[1] * ypn + 2 * gp[2] * zpn) / (un - up + TINY30) - 1.0;
  double psi;
  switch (scheme) {
    case 4: psi = fmax(0., fmin(fmin(2 * r, 0.5 * r + 0.5), 2.0)); break;
    case 5: psi = fmax(0., fmin(fmin(fmin(2 * r, 0.75 * r + 0.25), 0.25 * r + 0.75), 2.0)); break;
    case 6: psi = fmax(0., fmin(fmin(2 * r, 2. / 3. * r + 1. / 3.0), 2.0)); break;
    case 7: psi = fmax(0., fmin(fmin(2 * r, 0.75 * r + 0.25), 4.0)); break;
    case 8: psi = fmax(0., fmin(fmin(1.5 * r, 0.75 * r + 0.25), 2.5)); break;
    case 9: psi = fmax(0., (r + fabs(r)) * (3 * r + 1.0) / (2 * ((r + 1.0) * (r + 1.0)))); break;
    case 10: psi = fmax(0., fmin((r + fabs(r)) / (r + 1.0), 2.0)); break;
    case 11: psi = fmax(0., 3 * r * (r + 1.0) / (2 * (r * r + r + 1.0))); break;
    case 12: psi = fmax(0., fmin(r, 1.0)); break;
    case 13: psi = fmax(0., fmin(2 * r, 1.0)); break;
    case 14: psi = fmax(0., fmin(10 * r, 1.0)); break;
    case 15: psi = fmax(0., fmin(r, 4.0)); break;
    case 16: psi = 0.5 * r + 0.5; break;
    case 17: psi = S23 * r + S13; break;
    case 18: psi = 0.75 * r + 0.25; break;
    case 19: psi = fmax(0., fmin(fmin(fmin(2 * r, S13 * r + S23), S23 * r + S13), 2.0)); break;
    default: psi = 1.0; break;
  }
  return up + fxp * psi * (un - up);
}

