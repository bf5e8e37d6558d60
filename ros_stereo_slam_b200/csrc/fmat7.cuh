// fmat7.cuh -- 7-point fundamental-matrix solver + cubic, restating what
// cv::findFundamentalMat(FM_RANSAC) runs per minimal sample (reference call sites:
// src/tracking.cpp:34,75).  Same rules as cvmath.cuh (host+device, fixed operation order).
#pragma once
#include "cvmath.cuh"

namespace vo {

#ifndef VO_PI
#define VO_PI 3.1415926535897932384626433832795
#endif

// cv::solveCubic for 4 coefficients c[0] x^3 + c[1] x^2 + c[2] x + c[3]; returns the number
// of real roots written to r (OpenCV's conventions incl. degenerate cases; -1 = any x).
VO_HD int solve_cubic(const double* coef, double* r) {
  int n = 0;
  double a0 = coef[0], a1 = coef[1], a2 = coef[2], a3 = coef[3];
  double x0 = 0., x1 = 0., x2 = 0.;
  if (a0 == 0) {
    if (a1 == 0) {
      if (a2 == 0)
        n = a3 == 0 ? -1 : 0;
      else {
        x0 = -a3 / a2;
        n = 1;
      }
    } else {
      double d = a2 * a2 - 4 * a1 * a3;
      if (d >= 0) {
        d = sqrt(d);
        double q1 = (-a2 + d) * 0.5;
        double q2 = (a2 + d) * -0.5;
        if (fabs(q1) > fabs(q2)) {
          x0 = q1 / a1;
          x1 = a3 / q1;
        } else {
          x0 = q2 / a1;
          x1 = a3 / q2;
        }
        n = d > 0 ? 2 : 1;
      }
    }
  } else {
    a0 = 1. / a0;
    a1 *= a0;
    a2 *= a0;
    a3 *= a0;
    double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    double Qcubed = Q * Q * Q;
    double d = (a1 * a1 * (a2 * a2 - 4 * a1 * a3) + 2 * a2 * (9 * a1 * a3 - 2 * a2 * a2) - 27 * a3 * a3) * (1. / 108);
    if (d > 0) {
      double theta = acos(R / sqrt(Qcubed));
      double sqrtQ = sqrt(Q);
      double t0 = -2 * sqrtQ;
      double t1 = theta * (1. / 3);
      double t2 = a1 * (1. / 3);
      x0 = t0 * cos(t1) - t2;
      x1 = t0 * cos(t1 + (2. * VO_PI / 3)) - t2;
      x2 = t0 * cos(t1 + (4. * VO_PI / 3)) - t2;
      n = 3;
    } else if (d == 0) {
      if (R >= 0) {
        x0 = -2 * pow(R, 1. / 3) - a1 / 3;
        x1 = pow(R, 1. / 3) - a1 / 3;
      } else {
        x0 = 2 * pow(-R, 1. / 3) - a1 / 3;
        x1 = -pow(-R, 1. / 3) - a1 / 3;
      }
      x2 = 0;
      n = x0 == x1 ? 1 : 2;
      x1 = x0 == x1 ? 0 : x1;
    } else {
      double e;
      d = sqrt(-d);
      e = pow(d + fabs(R), 1. / 3);
      if (R > 0) e = -e;
      x0 = (e + Q / e) - a1 * (1. / 3);
      n = 1;
    }
  }
  r[0] = x0;
  r[1] = x1;
  r[2] = x2;
  return n;
}

// FMEstimatorCallback::run7Point, split at its SVD like EPnP (see jacobi_warp.cuh):
//   front: Hartley normalisation + the 7x9 design matrix written into the 9x9 buffer `v`
//          (rows 7, 8 zeroed); returns false for a degenerate sample (0 models);
//   SVD  : SVDecomp(A (7x9), W, U, Vt, MODIFY_A + FULL_UV).  m < n, so OpenCV runs the Jacobi on
//          the 7 rows of A itself (length 9), completes rows 7 and 8 of the 9x9 with its
//          pseudo-random Gram-Schmidt vectors, and returns that 9x9 as Vt (U is never used);
//   back : cubic in the null-space parameter, up to 3 models, de-normalisation.
struct FmatNorm {
  double m1cx, m1cy, m2cx, m2cy, scale1, scale2;
};

VO_HDN bool fmat_7point_front(const float* m1 /*14*/, const float* m2 /*14*/, double* v /*81*/, FmatNorm& nm) {
  int i;
  // Hartley normalisation of both point sets (centroid to origin, mean distance sqrt(2))
  double m1cx = 0, m1cy = 0, m2cx = 0, m2cy = 0, t, scale1 = 0, scale2 = 0;
  for (i = 0; i < 7; i++) {
    m1cx += (double)m1[i * 2];
    m1cy += (double)m1[i * 2 + 1];
    m2cx += (double)m2[i * 2];
    m2cy += (double)m2[i * 2 + 1];
  }
  t = 1. / 7;
  m1cx *= t; m1cy *= t; m2cx *= t; m2cy *= t;
  for (i = 0; i < 7; i++) {
    double dx = m1[i * 2] - m1cx, dy = m1[i * 2 + 1] - m1cy;
    scale1 += sqrt(dx * dx + dy * dy);
    dx = m2[i * 2] - m2cx; dy = m2[i * 2 + 1] - m2cy;
    scale2 += sqrt(dx * dx + dy * dy);
  }
  scale1 *= t;
  scale2 *= t;
  if (scale1 < FLT_EPSILON || scale2 < FLT_EPSILON) return false;
  scale1 = sqrt(2.) / scale1;
  scale2 = sqrt(2.) / scale2;
  nm.m1cx = m1cx; nm.m1cy = m1cy; nm.m2cx = m2cx; nm.m2cy = m2cy; nm.scale1 = scale1; nm.scale2 = scale2;

  for (i = 0; i < 7; i++) {
    double x0 = (m1[i * 2] - m1cx) * scale1, y0 = (m1[i * 2 + 1] - m1cy) * scale1;
    double x1 = (m2[i * 2] - m2cx) * scale2, y1 = (m2[i * 2 + 1] - m2cy) * scale2;
    v[i * 9 + 0] = x1 * x0;
    v[i * 9 + 1] = x1 * y0;
    v[i * 9 + 2] = x1;
    v[i * 9 + 3] = y1 * x0;
    v[i * 9 + 4] = y1 * y0;
    v[i * 9 + 5] = y1;
    v[i * 9 + 6] = x0;
    v[i * 9 + 7] = y0;
    v[i * 9 + 8] = 1;
  }
  for (i = 63; i < 81; i++) v[i] = 0;
  return true;
}

VO_HDN int fmat_7point_back(double* v /*81: Vt*/, const FmatNorm& nm, double* fmatrix /*27*/) {
  double c[4], r[3] = {0, 0, 0};
  double *f1, *f2;
  double t0, t1, t2;
  int i, k, n;
  const double m1cx = nm.m1cx, m1cy = nm.m1cy, m2cx = nm.m2cx, m2cy = nm.m2cy, scale1 = nm.scale1, scale2 = nm.scale2;
  f1 = v + 7 * 9;
  f2 = v + 8 * 9;

  for (i = 0; i < 9; i++) f1[i] -= f2[i];

  t0 = f2[4] * f2[8] - f2[5] * f2[7];
  t1 = f2[3] * f2[8] - f2[5] * f2[6];
  t2 = f2[3] * f2[7] - f2[4] * f2[6];

  c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;

  c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
         f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
         f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
         f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);

  t0 = f1[4] * f1[8] - f1[5] * f1[7];
  t1 = f1[3] * f1[8] - f1[5] * f1[6];
  t2 = f1[3] * f1[7] - f1[4] * f1[6];

  c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;

  c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
         f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
         f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
         f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);

  n = solve_cubic(c, r);
  if (n < 1 || n > 3) return n;

  for (k = 0; k < n; k++, fmatrix += 9) {
    double lambda = r[k], mu = 1.;
    double s = f1[8] * r[k] + f2[8];
    if (fabs(s) > DBL_EPSILON) {
      mu = 1. / s;
      lambda *= mu;
      fmatrix[8] = 1.;
    } else
      fmatrix[8] = 0.;
    for (i = 0; i < 8; i++) fmatrix[i] = f1[i] * lambda + f2[i] * mu;

    // de-normalise: F <- T2^T * F * T1, then scale so that F(3,3) = 1
    const double T1[9] = {scale1, 0, -scale1 * m1cx, 0, scale1, -scale1 * m1cy, 0, 0, 1};
    const double T2[9] = {scale2, 0, -scale2 * m2cx, 0, scale2, -scale2 * m2cy, 0, 0, 1};
    double tmp[9], out[9];
    for (int rr = 0; rr < 3; rr++)
      for (int cc = 0; cc < 3; cc++) {
        double acc = 0;
        for (int kk = 0; kk < 3; kk++) acc += T2[kk * 3 + rr] * fmatrix[kk * 3 + cc];
        tmp[rr * 3 + cc] = acc;
      }
    for (int rr = 0; rr < 3; rr++)
      for (int cc = 0; cc < 3; cc++) {
        double acc = 0;
        for (int kk = 0; kk < 3; kk++) acc += tmp[rr * 3 + kk] * T1[kk * 3 + cc];
        out[rr * 3 + cc] = acc;
      }
    if (fabs(out[8]) > FLT_EPSILON) {
      double sc = 1. / out[8];
      for (i = 0; i < 9; i++) out[i] *= sc;
    }
    for (i = 0; i < 9; i++) fmatrix[i] = out[i];
  }
  return n;
}

// up to 3 F (row-major 3x3 each, written consecutively); returns the number of models
VO_HDN int fmat_7point(const float* m1 /*14*/, const float* m2 /*14*/, double* fmatrix /*27*/) {
  double v[81], w[7];
  FmatNorm nm;
  if (!fmat_7point_front(m1, m2, v, nm)) return 0;
  jacobi_svd<9, 7, 9, false>(v, w, v);   // U (= V of the transposed problem) is not needed
  return fmat_7point_back(v, nm, fmatrix);
}

}  // namespace vo
