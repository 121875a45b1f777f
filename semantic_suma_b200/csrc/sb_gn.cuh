// sb_gn.cuh -- fp64 Gauss-Newton arithmetic: SE3 exp/log (core/lie_algebra.cpp:4-71), 6x6 LDL^T solve
// (Eigen::LDLT at core/LieGaussNewton.cpp:60) and one optimiser step (LieGaussNewton.cpp:53-79,
// Objective::increment Objective.h:45-48). __host__ __device__: the same code runs in the last block of the fused
// Jacobian kernel and in the host-side sb_gn_step, and both give the same bits (-fmad=false / -ffp-contract=off).
#pragma once

#include "sb_math.cuh"

namespace sbg {

SB_HD void se3_exp(const double* x, double* T) {
  for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
  const double v0 = x[0], v1 = x[1], v2 = x[2], w0 = x[3], w1 = x[4], w2 = x[5];
  double theta = sqrt((w0 * w0 + w1 * w1) + w2 * w2);
  if (theta > 1e-10) {
    double K[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};  // row-major skew(omega)
    double K2[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        K2[r * 3 + c] = (K[r * 3 + 0] * K[0 * 3 + c] + K[r * 3 + 1] * K[1 * 3 + c]) + K[r * 3 + 2] * K[2 * 3 + c];
    double s, c;
    sbm::sincos_(theta, &s, &c);
    double alpha = s / theta;
    double beta = (1.0 - c) / (theta * theta);
    double gamma = beta;
    double delta = (theta - s) / ((theta * theta) * theta);
    double V[9];
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) {
        double I = (r == cc) ? 1.0 : 0.0;
        T[cc * 4 + r] = (I + alpha * K[r * 3 + cc]) + beta * K2[r * 3 + cc];
        V[r * 3 + cc] = (I + gamma * K[r * 3 + cc]) + delta * K2[r * 3 + cc];
      }
    T[12] = (V[0] * v0 + V[1] * v1) + V[2] * v2;
    T[13] = (V[3] * v0 + V[4] * v1) + V[5] * v2;
    T[14] = (V[6] * v0 + V[7] * v1) + V[8] * v2;
  } else {
    T[12] = v0;
    T[13] = v1;
    T[14] = v2;
  }
}

// LDL^T with diagonal pivoting on the lower triangle of a column-major 6x6; x = A^-1 b
SB_HD void ldlt_solve6(const double* Ain, const double* bin, double* x) {
  double A[36];
  int perm[6];
  for (int c = 0; c < 6; ++c)
    for (int r = 0; r < 6; ++r) A[c * 6 + r] = (r >= c) ? Ain[c * 6 + r] : Ain[r * 6 + c];
  for (int i = 0; i < 6; ++i) perm[i] = i;
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double best = fabs(A[k * 6 + k]);
    for (int i = k + 1; i < 6; ++i) {
      double v = fabs(A[i * 6 + i]);
      if (v > best) {
        best = v;
        piv = i;
      }
    }
    if (piv != k) {
      for (int i = 0; i < 6; ++i) {
        double t = A[k * 6 + i];
        A[k * 6 + i] = A[piv * 6 + i];
        A[piv * 6 + i] = t;
      }
      for (int i = 0; i < 6; ++i) {
        double t = A[i * 6 + k];
        A[i * 6 + k] = A[i * 6 + piv];
        A[i * 6 + piv] = t;
      }
      int t = perm[k];
      perm[k] = perm[piv];
      perm[piv] = t;
    }
    double dk = A[k * 6 + k];
    if (dk == 0.0) continue;
    for (int i = k + 1; i < 6; ++i) A[k * 6 + i] = A[k * 6 + i] / dk;
    for (int j = k + 1; j < 6; ++j)
      for (int i = j; i < 6; ++i) {
        A[j * 6 + i] = A[j * 6 + i] - (A[k * 6 + i] * dk) * A[k * 6 + j];
        A[i * 6 + j] = A[j * 6 + i];
      }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) y[i] = bin[perm[i]];
  for (int i = 0; i < 6; ++i)
    for (int k = 0; k < i; ++k) y[i] = y[i] - A[k * 6 + i] * y[k];
  for (int i = 0; i < 6; ++i) y[i] = (A[i * 6 + i] == 0.0) ? 0.0 : y[i] / A[i * 6 + i];
  for (int i = 5; i >= 0; --i)
    for (int k = i + 1; k < 6; ++k) y[i] = y[i] - A[i * 6 + k] * y[k];
  for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
}

// raw fixed-point sums -> the reference's 48-value layout (Frame2Model.cpp:214-227)
SB_HD void unpack48(const long long* raw, double* out48) {
  const double s = 1.0 / 1073741824.0;
  int k = 0;
  for (int c = 0; c < 6; ++c)
    for (int r = c; r < 6; ++r) {
      double v = (double)raw[k++] * s;
      out48[c * 6 + r] = v;
      out48[r * 6 + c] = v;
    }
  for (int r = 0; r < 6; ++r) out48[36 + r] = (double)raw[k++] * s;
  out48[42] = (double)raw[29];
  out48[43] = (double)raw[27] * s;
  out48[44] = (double)raw[30];
  out48[45] = (double)raw[28] * s;
  out48[46] = (double)raw[31];
  out48[47] = 0.0;
}

// LieGaussNewton::step: returns 0 when a stop criterion fired (the increment is applied in either case)
SB_HD int gn_step(const double* out48, double last_error, double eps, double delta_thr, double* pose, double* dx) {
  int result = 1;
  double current_error = out48[43];
  double nb[6];
  for (int i = 0; i < 6; ++i) nb[i] = -out48[36 + i];
  ldlt_solve6(out48, nb, dx);
  double linf = 0.0, maxc = out48[36];
  for (int i = 0; i < 6; ++i) {
    if (fabs(dx[i]) > linf) linf = fabs(dx[i]);
    if (out48[36 + i] > maxc) maxc = out48[36 + i];
  }
  if (linf < delta_thr) result = 0;
  if (fabs(maxc) < eps) result = 0;
  if (current_error < last_error && fabs(current_error - last_error) < eps) result = 0;
  double E[16], Pn[16];
  se3_exp(dx, E);
  sbm::mat4_mul<double>(E, pose, Pn);
  for (int i = 0; i < 16; ++i) pose[i] = Pn[i];
  return result;
}

// rigid inverse [R^T | -R^T t] in fp64 (Eigen's general inverse in the reference; rule fixed here)
SB_HD void rigid_inverse_d(const double* M, double* Mi) {
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) Mi[c * 4 + r] = M[r * 4 + c];
  for (int r = 0; r < 3; ++r) {
    double v = (M[r * 4 + 0] * M[12] + M[r * 4 + 1] * M[13]) + M[r * 4 + 2] * M[14];
    Mi[12 + r] = -v;
  }
  Mi[3] = Mi[7] = Mi[11] = 0.0;
  Mi[15] = 1.0;
}
// fp32 pose -> fp32 inverse, evaluated in fp64 and rounded once
SB_HD void rigid_inverse_f(const float* M, float* Mi) {
  double Md[16], Id[16];
  for (int i = 0; i < 16; ++i) Md[i] = (double)M[i];
  rigid_inverse_d(Md, Id);
  for (int i = 0; i < 16; ++i) Mi[i] = (float)Id[i];
  Mi[3] = Mi[7] = Mi[11] = 0.0f;
  Mi[15] = 1.0f;
}

}  // namespace sbg
