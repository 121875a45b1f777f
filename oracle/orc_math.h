/* orc_math.h -- TEST INFRASTRUCTURE (CPU oracle). Not part of the shipped product.
 *
 * Scalar fp32/fp64 arithmetic used by the oracle's restatement of the SuMa++ shaders.
 *
 * Why hand-written transcendentals: GLSL leaves the precision of atan/asin/acos/sin/exp/log
 * implementation-defined, so "the reference's value" does not exist bit-wise. The oracle pins
 * one concrete evaluation (Cephes-style single-precision polynomials, strict left-to-right
 * IEEE-754 fp32 +,*,/,sqrt, no FMA contraction). Every function here is a pure sequence of
 * correctly-rounded IEEE operations, so any conforming machine (x86 SSE, sm_100a with
 * -fmad=false) reproduces it bit-for-bit.
 *
 * Build flags that this file relies on: -ffp-contract=off, no -ffast-math.
 */
#ifndef ORC_MATH_H_
#define ORC_MATH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_PI_F 3.14159265358979323846f
#define ORC_INV_PI_F 0.31830988618379067154f
#define ORC_PIO2_F 1.57079632679489661923f
#define ORC_PIO4_F 0.78539816339744830962f
#define ORC_RAD2DEG_F 57.29577951308232f

static inline float orc_bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t orc_f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* atan on |t| <= tan(pi/8) */
static inline float orc_atan_poly(float t) {
  float z = t * t;
  float p = 8.05374449538e-2f;
  p = p * z - 1.38776856032e-1f;
  p = p * z + 1.99777106478e-1f;
  p = p * z - 3.33329491539e-1f;
  return (p * z) * t + t;
}

/* GLSL atan(y, x). atan(0,0) := 0 (GLSL: undefined). */
static inline float orc_atan2f(float y, float x) {
  float ax = fabsf(x), ay = fabsf(y);
  int swap = ay > ax;
  float hi = swap ? ay : ax;
  float lo = swap ? ax : ay;
  if (hi == 0.0f) return 0.0f;
  float t = lo / hi;
  float r;
  if (t > 0.4142135623730950f)
    r = ORC_PIO4_F + orc_atan_poly((t - 1.0f) / (t + 1.0f));
  else
    r = orc_atan_poly(t);
  if (swap) r = ORC_PIO2_F - r;
  if (x < 0.0f) r = ORC_PI_F - r;
  if (y < 0.0f) r = -r;
  return r;
}

/* GLSL asin(x); |x| > 1 -> NaN (sqrt of a negative). */
static inline float orc_asinf(float x) {
  float a = fabsf(x);
  if (a < 1.0e-4f) return x;
  float z, w;
  int flag = a > 0.5f;
  if (flag) {
    z = 0.5f * (1.0f - a);
    w = sqrtf(z);
  } else {
    w = a;
    z = w * w;
  }
  float p = 4.2163199048e-2f;
  p = p * z + 2.4181311049e-2f;
  p = p * z + 4.5470025998e-2f;
  p = p * z + 7.4953002686e-2f;
  p = p * z + 1.6666752422e-1f;
  float r = (p * z) * w + w;
  if (flag) {
    r = r + r;
    r = ORC_PIO2_F - r;
  }
  return x < 0.0f ? -r : r;
}

static inline float orc_acosf(float x) {
  if (x < -0.5f) return ORC_PI_F - 2.0f * orc_asinf(sqrtf(0.5f * (1.0f + x)));
  if (x > 0.5f) return 2.0f * orc_asinf(sqrtf(0.5f * (1.0f - x)));
  return ORC_PIO2_F - orc_asinf(x);
}

static inline float orc_sinf(float xx) {
  if (xx != xx) return xx;
  float x = fabsf(xx);
  int sign = xx < 0.0f;
  if (x > 8192.0f) return 0.0f;
  int j = (int)(1.27323954473516f * x);
  float y = (float)j;
  if (j & 1) {
    j += 1;
    y += 1.0f;
  }
  j &= 7;
  if (j > 3) {
    sign = !sign;
    j -= 4;
  }
  x = ((x - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
  float z = x * x, r, p;
  if (j == 1 || j == 2) {
    p = 2.443315711809948e-5f;
    p = p * z - 1.388731625493765e-3f;
    p = p * z + 4.166664568298827e-2f;
    r = (p * z) * z;
    r = r - 0.5f * z;
    r = r + 1.0f;
  } else {
    p = -1.9515295891e-4f;
    p = p * z + 8.3321608736e-3f;
    p = p * z - 1.6666654611e-1f;
    r = (p * z) * x + x;
  }
  return sign ? -r : r;
}

static inline float orc_pow2i(int n) { /* 2^n, n in [-126,127] */
  return orc_bits2f((uint32_t)(n + 127) << 23);
}

static inline float orc_expf(float x) {
  if (x != x) return x;
  if (x > 88.0f) return INFINITY;
  if (x < -87.0f) return 0.0f;
  float fn = floorf(1.44269504088896341f * x + 0.5f);
  int n = (int)fn;
  x = x - fn * 0.693359375f;
  x = x - fn * (-2.12194440e-4f);
  float z = x * x;
  float p = 1.9875691500e-4f;
  p = p * x + 1.3981999507e-3f;
  p = p * x + 8.3334519073e-3f;
  p = p * x + 4.1665795894e-2f;
  p = p * x + 1.6666665459e-1f;
  p = p * x + 5.0000001201e-1f;
  float r = (p * z + x) + 1.0f;
  int n1 = n >> 1, n2 = n - n1;
  return (r * orc_pow2i(n1)) * orc_pow2i(n2);
}

static inline float orc_logf(float x) {
  if (x != x) return x;
  if (x < 0.0f) return NAN;
  if (x == 0.0f) return -INFINITY;
  if (x == INFINITY) return x;
  int e = 0;
  uint32_t b = orc_f2bits(x);
  if (((b >> 23) & 0xffu) == 0u) { /* denormal: scale up */
    x = x * 8388608.0f;
    b = orc_f2bits(x);
    e = -23;
  }
  e += (int)((b >> 23) & 0xffu) - 126;
  float m = orc_bits2f((b & 0x807fffffu) | 0x3f000000u); /* [0.5,1) */
  if (m < 0.707106781186547524f) {
    e -= 1;
    m = (m + m) - 1.0f;
  } else {
    m = m - 1.0f;
  }
  float z = m * m;
  float p = 7.0376836292e-2f;
  p = p * m - 1.1514610310e-1f;
  p = p * m + 1.1676998740e-1f;
  p = p * m - 1.2420140846e-1f;
  p = p * m + 1.4249322787e-1f;
  p = p * m - 1.6668057665e-1f;
  p = p * m + 2.0000714765e-1f;
  p = p * m - 2.4999993993e-1f;
  p = p * m + 3.3333331174e-1f;
  float y = (p * m) * z;
  float fe = (float)e;
  y = y + (-2.12194440e-4f) * fe;
  y = y + (-0.5f) * z;
  float r = m + y;
  r = r + 0.693359375f * fe;
  return r;
}

/* GLSL round(): ties away from zero is the rule the oracle fixes (GLSL leaves ties open). */
static inline float orc_roundf(float x) { return roundf(x); }

/* ---- fp64 sin/cos for SE3::exp (lie_algebra.cpp:19-25): fixed polynomial so host and device GN agree ---- */
static inline void orc_sincos(double xx, double* s, double* c) {
  static const double S[6] = {1.58962301576546568060E-10, -2.50507477628578072866E-8, 2.75573136213857245213E-6,
                              -1.98412698295895385996E-4, 8.33333333332211858878E-3,  -1.66666666666666307295E-1};
  static const double C[6] = {-1.13585365213876817300E-11, 2.08757008419747316778E-9, -2.75573141792967388112E-7,
                              2.48015872888517045348E-5,   -1.38888888888730564116E-3, 4.16666666666665929218E-2};
  double x = fabs(xx);
  if (!(x < 1.0e9)) { /* outside the range the path ever produces (also NaN) */
    *s = xx - xx;
    *c = xx - xx;
    return;
  }
  int64_t j = (int64_t)(x * 1.27323954473516268615);
  double y = (double)j;
  if (j & 1) {
    j += 1;
    y += 1.0;
  }
  j &= 7;
  int ssign = xx < 0.0, csign = 0;
  if (j > 3) {
    ssign = !ssign;
    csign = !csign;
    j -= 4;
  }
  if (j > 1) csign = !csign;
  double z = ((x - y * 7.85398125648498535156E-1) - y * 3.77489470793079817668E-8) - y * 2.69515142907905952645E-15;
  double zz = z * z;
  double ps = S[0];
  for (int i = 1; i < 6; ++i) ps = ps * zz + S[i];
  double pc = C[0];
  for (int i = 1; i < 6; ++i) pc = pc * zz + C[i];
  double vs = z + (z * zz) * ps;
  double vc = (1.0 - 0.5 * zz) + (zz * zz) * pc;
  double sv, cv;
  if (j == 1 || j == 2) {
    sv = vc;
    cv = vs;
  } else {
    sv = vs;
    cv = vc;
  }
  *s = ssign ? -sv : sv;
  *c = csign ? -cv : cv;
}

/* ---- small vector helpers with the evaluation order the oracle fixes ---- */
typedef struct { float x, y, z; } orc_v3;

static inline orc_v3 orc_mk3(float x, float y, float z) { orc_v3 r = {x, y, z}; return r; }
static inline float orc_dot3(orc_v3 a, orc_v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float orc_len3(orc_v3 a) { return sqrtf(orc_dot3(a, a)); }
static inline orc_v3 orc_sub3(orc_v3 a, orc_v3 b) { return orc_mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline orc_v3 orc_add3(orc_v3 a, orc_v3 b) { return orc_mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline orc_v3 orc_scale3(float s, orc_v3 a) { return orc_mk3(s * a.x, s * a.y, s * a.z); }
static inline orc_v3 orc_divs3(orc_v3 a, float s) { return orc_mk3(a.x / s, a.y / s, a.z / s); }
static inline orc_v3 orc_neg3(orc_v3 a) { return orc_mk3(-a.x, -a.y, -a.z); }
static inline orc_v3 orc_cross3(orc_v3 a, orc_v3 b) {
  return orc_mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline orc_v3 orc_normalize3(orc_v3 a) { return orc_divs3(a, orc_len3(a)); }

/* column-major 4x4 (Eigen/GL storage): M[c*4+r] */
static inline orc_v3 orc_xform_point(const float* M, orc_v3 p) { /* (M * (p,1)).xyz */
  orc_v3 r;
  r.x = ((M[0] * p.x + M[4] * p.y) + M[8] * p.z) + M[12];
  r.y = ((M[1] * p.x + M[5] * p.y) + M[9] * p.z) + M[13];
  r.z = ((M[2] * p.x + M[6] * p.y) + M[10] * p.z) + M[14];
  return r;
}
static inline orc_v3 orc_xform_dir(const float* M, orc_v3 p) { /* (M * (p,0)).xyz */
  orc_v3 r;
  r.x = (M[0] * p.x + M[4] * p.y) + M[8] * p.z;
  r.y = (M[1] * p.x + M[5] * p.y) + M[9] * p.z;
  r.z = (M[2] * p.x + M[6] * p.y) + M[10] * p.z;
  return r;
}
/* C = A*B in fp32 */
static inline void orc_mat4_mul_f(const float* A, const float* B, float* C) {
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r)
      C[c * 4 + r] = ((A[0 * 4 + r] * B[c * 4 + 0] + A[1 * 4 + r] * B[c * 4 + 1]) + A[2 * 4 + r] * B[c * 4 + 2]) +
                     A[3 * 4 + r] * B[c * 4 + 3];
}
static inline void orc_mat4_mul_d(const double* A, const double* B, double* C) {
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r)
      C[c * 4 + r] = ((A[0 * 4 + r] * B[c * 4 + 0] + A[1 * 4 + r] * B[c * 4 + 1]) + A[2 * 4 + r] * B[c * 4 + 2]) +
                     A[3 * 4 + r] * B[c * 4 + 3];
}
/* rigid inverse [R^T | -R^T t] evaluated in fp64, rounded once to fp32.
 * (The reference calls Eigen's general 4x4 inverse / GLSL inverse(); both are implementation-defined in
 * rounding, the oracle fixes this rule.) */
static inline void orc_rigid_inverse_f(const float* M, float* Mi) {
  double R[9], t[3];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) R[c * 3 + r] = (double)M[c * 4 + r];
  for (int r = 0; r < 3; ++r) t[r] = (double)M[12 + r];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) Mi[c * 4 + r] = (float)R[r * 3 + c];
  for (int r = 0; r < 3; ++r) {
    double v = (R[r * 3 + 0] * t[0] + R[r * 3 + 1] * t[1]) + R[r * 3 + 2] * t[2];
    Mi[12 + r] = (float)(-v);
  }
  Mi[3] = Mi[7] = Mi[11] = 0.0f;
  Mi[15] = 1.0f;
}
static inline void orc_rigid_inverse_d(const double* M, double* Mi) {
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) Mi[c * 4 + r] = M[r * 4 + c];
  for (int r = 0; r < 3; ++r) {
    double v = (M[r * 4 + 0] * M[12] + M[r * 4 + 1] * M[13]) + M[r * 4 + 2] * M[14];
    Mi[12 + r] = -v;
  }
  Mi[3] = Mi[7] = Mi[11] = 0.0;
  Mi[15] = 1.0;
}

#endif /* ORC_MATH_H_ */
