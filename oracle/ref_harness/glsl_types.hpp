// glsl_types.hpp -- TEST INFRASTRUCTURE (oracle/_ref): the GLSL 3.30 value types and built-ins the reference's hot-path
// shaders use, so that the shader text -- mechanically rewritten by glsl2cpp.py -- compiles as C++ and runs on the CPU.
//
// GLSL leaves the precision of its built-ins (and the association order of dot / matrix products) to the
// implementation. Two implementations live behind one switch:
//   -DREF_MATH_PINNED   the rules the oracle pins (oracle/orc_math.h: Cephes-style fp32 polynomials, left-to-right
//                       fp32 sums, rigid pose inverse). With it the transpiled shaders must agree with oracle/ BIT FOR BIT.
//   (default)           "precise GL": every built-in evaluated in fp64 (libm) and rounded once to fp32 -- an
//                       independent, at-least-as-accurate legal GL. The distance between the two is what the
//                       implementation-defined precision of GL is worth on this path (tests/test_ref_shaders.py).
// Plain +,-,*,/ on floats and vectors are single IEEE fp32 operations in both (build with -ffp-contract=off).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#ifdef REF_MATH_PINNED
extern "C" {
#include "../orc_math.h"
}
#endif

namespace glsl {

struct vec2; struct vec3; struct vec4;
template <int K> struct VecOf;
template <> struct VecOf<2> { typedef vec2 type; };
template <> struct VecOf<3> { typedef vec3 type; };
template <> struct VecOf<4> { typedef vec4 type; };

// swizzle proxy living in a union with the parent's components: reads convert to vecK, writes scatter
template <int N, int... I>
struct Swz {
  float d[N];
  typedef typename VecOf<sizeof...(I)>::type V;
  operator V() const { return V(d[I]...); }
  Swz& operator=(const V& v) { const int ix[] = {I...}; for (int k = 0; k < (int)sizeof...(I); ++k) d[ix[k]] = v[k]; return *this; }
  Swz& operator=(const Swz& o) { V v = o; return *this = v; }
  Swz& operator+=(const V& v) { V t = *this; return *this = t + v; }
  Swz& operator-=(const V& v) { V t = *this; return *this = t - v; }
  Swz& operator*=(float s) { V t = *this; return *this = t * s; }
  Swz& operator/=(float s) { V t = *this; return *this = t / s; }
};

struct ivec2 { int x, y; ivec2() : x(0), y(0) {} ivec2(int a, int b) : x(a), y(b) {} };

struct vec2 {
  union {
    struct { float x, y; };
    struct { float r, g; };
    float d[2];
#include "swizzles2.inc"
  };
  vec2() : x(0), y(0) {}
  vec2(float a) : x(a), y(a) {}
  vec2(float a, float b) : x(a), y(b) {}
  vec2(const ivec2& v) : x((float)v.x), y((float)v.y) {}
  vec2(const vec2& o) : x(o.x), y(o.y) {}
  vec2& operator=(const vec2& o) { x = o.x; y = o.y; return *this; }
  float operator[](int i) const { return d[i]; }
  float& operator[](int i) { return d[i]; }
};

struct vec3 {
  union {
    struct { float x, y, z; };
    struct { float r, g, b; };
    float d[3];
#include "swizzles3.inc"
  };
  vec3() : x(0), y(0), z(0) {}
  vec3(float a) : x(a), y(a), z(a) {}
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  vec3(const vec2& a, float c) : x(a.x), y(a.y), z(c) {}
  vec3(float a, const vec2& b) : x(a), y(b.x), z(b.y) {}
  explicit vec3(const vec4& v);
  vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
  vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
  float operator[](int i) const { return d[i]; }
  float& operator[](int i) { return d[i]; }
};

struct vec4 {
  union {
    struct { float x, y, z, w; };
    struct { float r, g, b, a; };
    float d[4];
#include "swizzles4.inc"
  };
  vec4() : x(0), y(0), z(0), w(0) {}
  vec4(float s) : x(s), y(s), z(s), w(s) {}
  vec4(float a, float b, float c, float e) : x(a), y(b), z(c), w(e) {}
  vec4(const vec3& v, float e) : x(v.x), y(v.y), z(v.z), w(e) {}
  vec4(float a, const vec3& v) : x(a), y(v.x), z(v.y), w(v.z) {}
  vec4(const vec2& a, const vec2& b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
  vec4(const vec2& a, float c, float e) : x(a.x), y(a.y), z(c), w(e) {}
  vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
  vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
  float operator[](int i) const { return d[i]; }
  float& operator[](int i) { return d[i]; }
  // GLSL scalar constructors take the first component: int(texture(index_map, ...)) (update_surfels.vert:303)
  explicit operator int() const { return (int)x; }
  explicit operator float() const { return x; }
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

struct bvec2 { bool x, y; };
struct bvec3 { bool x, y, z; };
struct bvec4 { bool x, y, z, w; };
inline bool all(bvec2 b) { return b.x && b.y; }
inline bool all(bvec3 b) { return b.x && b.y && b.z; }
inline bool all(bvec4 b) { return b.x && b.y && b.z && b.w; }
inline bool any(bvec3 b) { return b.x || b.y || b.z; }

// ---- component-wise arithmetic: one fp32 IEEE operation per component ----
#define GLSL_VEC_OPS(V, N)                                                                                        \
  inline V operator+(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; } \
  inline V operator-(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; } \
  inline V operator*(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.d[i]; return r; } \
  inline V operator/(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b.d[i]; return r; } \
  inline V operator+(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + s; return r; }       \
  inline V operator-(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - s; return r; }       \
  inline V operator*(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }       \
  inline V operator/(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / s; return r; }       \
  inline V operator+(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r.d[i] = s + a.d[i]; return r; }       \
  inline V operator-(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r.d[i] = s - a.d[i]; return r; }       \
  inline V operator*(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }       \
  inline V operator/(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r.d[i] = s / a.d[i]; return r; }       \
  inline V operator-(const V& a) { V r; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }                   \
  inline V& operator+=(V& a, const V& b) { a = a + b; return a; }                                                   \
  inline V& operator-=(V& a, const V& b) { a = a - b; return a; }                                                   \
  inline V& operator*=(V& a, float s) { a = a * s; return a; }                                                      \
  inline V& operator/=(V& a, float s) { a = a / s; return a; }
GLSL_VEC_OPS(vec2, 2)
GLSL_VEC_OPS(vec3, 3)
GLSL_VEC_OPS(vec4, 4)
// the double-typed scalar overloads keep expressions such as `2.0f * v` unambiguous when an int sneaks in
inline vec2 operator*(int s, const vec2& a) { return (float)s * a; }
inline vec3 operator*(int s, const vec3& a) { return (float)s * a; }
inline vec4 operator*(int s, const vec4& a) { return (float)s * a; }

inline bvec3 lessThan(const vec3& a, const vec3& b) { return bvec3{a.x < b.x, a.y < b.y, a.z < b.z}; }
inline bvec3 greaterThanEqual(const vec3& a, const vec3& b) { return bvec3{a.x >= b.x, a.y >= b.y, a.z >= b.z}; }
inline bvec2 lessThan(const vec2& a, const vec2& b) { return bvec2{a.x < b.x, a.y < b.y}; }

// ---- scalar built-ins ----
namespace impl {
#ifdef REF_MATH_PINNED
inline float atan2_(float y, float x) { return orc_atan2f(y, x); }
inline float asin_(float x) { return orc_asinf(x); }
inline float acos_(float x) { return orc_acosf(x); }
inline float sin_(float x) { return orc_sinf(x); }
inline float cos_(float x) { return orc_sinf(x + ORC_PIO2_F); }  // not reached by the hot path (centerize() is dead code)
inline float exp_(float x) { return orc_expf(x); }
inline float log_(float x) { return orc_logf(x); }
inline float sqrt_(float x) { return sqrtf(x); }
inline float degrees_(float r) { return ORC_RAD2DEG_F * r; }
inline float radians_(float d) { return d * (float)(3.14159265358979323846 / 180.0); }
inline float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline float dot4(const float* a, const float* b) { return ((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]) + a[3] * b[3]; }
inline float dot2(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1]; }
#else
inline float atan2_(float y, float x) { return (float)std::atan2((double)y, (double)x); }
inline float asin_(float x) { return (float)std::asin((double)x); }
inline float acos_(float x) { return (float)std::acos((double)x); }
inline float sin_(float x) { return (float)std::sin((double)x); }
inline float cos_(float x) { return (float)std::cos((double)x); }
inline float exp_(float x) { return (float)std::exp((double)x); }
inline float log_(float x) { return (float)std::log((double)x); }
inline float sqrt_(float x) { return (float)std::sqrt((double)x); }
inline float degrees_(float r) { return (float)((double)r * (180.0 / 3.14159265358979323846)); }
inline float radians_(float d) { return (float)((double)d * (3.14159265358979323846 / 180.0)); }
inline float dot3(const float* a, const float* b) { return (float)((double)a[0] * b[0] + (double)a[1] * b[1] + (double)a[2] * b[2]); }
inline float dot4(const float* a, const float* b) { return (float)((double)a[0] * b[0] + (double)a[1] * b[1] + (double)a[2] * b[2] + (double)a[3] * b[3]); }
inline float dot2(const float* a, const float* b) { return (float)((double)a[0] * b[0] + (double)a[1] * b[1]); }
#endif
}  // namespace impl

inline float abs(float x) { return std::fabs(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float floor(float x) { return std::floor(x); }
inline float fract(float x) { return x - std::floor(x); }
inline float round(float x) { return ::roundf(x); }  // ties away from zero: the rule the oracle fixes (GLSL leaves ties open)
inline float min(float a, float b) { return b < a ? b : a; }  // GLSL: y if y < x else x
inline float max(float a, float b) { return a < b ? b : a; }  // GLSL: y if x < y else x
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float sqrt(float x) { return impl::sqrt_(x); }
inline float atan(float y, float x) { return impl::atan2_(y, x); }
inline float asin(float x) { return impl::asin_(x); }
inline float acos(float x) { return impl::acos_(x); }
inline float sin(float x) { return impl::sin_(x); }
inline float cos(float x) { return impl::cos_(x); }
inline float exp(float x) { return impl::exp_(x); }
inline float log(float x) { return impl::log_(x); }
inline float degrees(float r) { return impl::degrees_(r); }
inline float radians(float d) { return impl::radians_(d); }

#define GLSL_MAP1(V, N, F) inline V F(const V& a) { V r; for (int i = 0; i < N; ++i) r.d[i] = F(a.d[i]); return r; }
GLSL_MAP1(vec2, 2, abs) GLSL_MAP1(vec3, 3, abs) GLSL_MAP1(vec4, 4, abs)
GLSL_MAP1(vec2, 2, floor) GLSL_MAP1(vec3, 3, floor) GLSL_MAP1(vec4, 4, floor)
GLSL_MAP1(vec2, 2, fract) GLSL_MAP1(vec3, 3, fract) GLSL_MAP1(vec4, 4, fract)
inline vec3 clamp(const vec3& a, float lo, float hi) { return vec3(clamp(a.x, lo, hi), clamp(a.y, lo, hi), clamp(a.z, lo, hi)); }
inline vec3 mix(const vec3& a, const vec3& b, float t) { return vec3(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t)); }
inline vec4 mix(const vec4& a, const vec4& b, float t) { return vec4(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t), mix(a.w, b.w, t)); }

inline float dot(const vec2& a, const vec2& b) { return impl::dot2(a.d, b.d); }
inline float dot(const vec3& a, const vec3& b) { return impl::dot3(a.d, b.d); }
inline float dot(const vec4& a, const vec4& b) { return impl::dot4(a.d, b.d); }
inline float length(const vec2& a) { return impl::sqrt_(dot(a, a)); }
inline float length(const vec3& a) { return impl::sqrt_(dot(a, a)); }
inline float length(const vec4& a) { return impl::sqrt_(dot(a, a)); }
inline vec3 cross(const vec3& a, const vec3& b) {
  return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
#ifdef REF_MATH_PINNED
inline vec3 normalize(const vec3& a) { return a / length(a); }
inline vec4 normalize(const vec4& a) { return a / length(a); }
#else
inline vec3 normalize(const vec3& a) {
  double l = std::sqrt((double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z);
  return vec3((float)(a.x / l), (float)(a.y / l), (float)(a.z / l));
}
inline vec4 normalize(const vec4& a) {
  double l = std::sqrt((double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w);
  return vec4((float)(a.x / l), (float)(a.y / l), (float)(a.z / l), (float)(a.w / l));
}
#endif

// ---- matrices: column-major, m[c] is a column (GLSL) ----
struct mat4;
struct mat3 {
  vec3 c[3];
  mat3() {}
  explicit mat3(const mat4& m);
  mat3(const vec3& a, const vec3& b, const vec3& e) { c[0] = a; c[1] = b; c[2] = e; }
  vec3& operator[](int i) { return c[i]; }
  const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
  vec4 c[4];
  mat4() { for (int i = 0; i < 4; ++i) c[i] = vec4(0); }
  explicit mat4(float s) { for (int i = 0; i < 4; ++i) { c[i] = vec4(0); c[i][i] = s; } }
  mat4(const vec4& a, const vec4& b, const vec4& e, const vec4& f) { c[0] = a; c[1] = b; c[2] = e; c[3] = f; }
  explicit mat4(const mat3& m) {
    for (int i = 0; i < 3; ++i) c[i] = vec4(m[i], 0.0f);
    c[3] = vec4(0, 0, 0, 1);
  }
  explicit mat4(const float* colmajor) { for (int i = 0; i < 4; ++i) c[i] = vec4(colmajor[4 * i], colmajor[4 * i + 1], colmajor[4 * i + 2], colmajor[4 * i + 3]); }
  vec4& operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
  void store(float* colmajor) const { for (int i = 0; i < 4; ++i) for (int r = 0; r < 4; ++r) colmajor[4 * i + r] = c[i][r]; }
};
inline mat3::mat3(const mat4& m) { c[0] = vec3(m[0]); c[1] = vec3(m[1]); c[2] = vec3(m[2]); }
inline mat3 operator-(const mat3& m) { return mat3(-m[0], -m[1], -m[2]); }
inline mat3 transpose(const mat3& m) {
  return mat3(vec3(m[0].x, m[1].x, m[2].x), vec3(m[0].y, m[1].y, m[2].y), vec3(m[0].z, m[1].z, m[2].z));
}
inline vec3 operator*(const mat3& m, const vec3& v) {
  vec3 r;
  for (int i = 0; i < 3; ++i) {
    float row[3] = {m[0][i], m[1][i], m[2][i]};
    r[i] = impl::dot3(row, v.d);
  }
  return r;
}
inline vec4 operator*(const mat4& m, const vec4& v) {
  vec4 r;
  for (int i = 0; i < 4; ++i) {
    float row[4] = {m[0][i], m[1][i], m[2][i], m[3][i]};
    r[i] = impl::dot4(row, v.d);
  }
  return r;
}
inline mat4 operator*(const mat4& a, const mat4& b) {
  mat4 r;
  for (int c = 0; c < 4; ++c) r[c] = a * b[c];
  return r;
}
// GLSL inverse(mat4): only ever applied to rigid poses on this path (update_surfels.vert:194)
inline mat4 inverse(const mat4& m) {
  float a[16], b[16];
  m.store(a);
#ifdef REF_MATH_PINNED
  orc_rigid_inverse_f(a, b);
#else
  // general 4x4 inverse by cofactors in fp64, rounded once
  double s[16], inv[16];
  for (int i = 0; i < 16; ++i) s[i] = a[i];
  inv[0] = s[5] * s[10] * s[15] - s[5] * s[11] * s[14] - s[9] * s[6] * s[15] + s[9] * s[7] * s[14] + s[13] * s[6] * s[11] - s[13] * s[7] * s[10];
  inv[4] = -s[4] * s[10] * s[15] + s[4] * s[11] * s[14] + s[8] * s[6] * s[15] - s[8] * s[7] * s[14] - s[12] * s[6] * s[11] + s[12] * s[7] * s[10];
  inv[8] = s[4] * s[9] * s[15] - s[4] * s[11] * s[13] - s[8] * s[5] * s[15] + s[8] * s[7] * s[13] + s[12] * s[5] * s[11] - s[12] * s[7] * s[9];
  inv[12] = -s[4] * s[9] * s[14] + s[4] * s[10] * s[13] + s[8] * s[5] * s[14] - s[8] * s[6] * s[13] - s[12] * s[5] * s[10] + s[12] * s[6] * s[9];
  inv[1] = -s[1] * s[10] * s[15] + s[1] * s[11] * s[14] + s[9] * s[2] * s[15] - s[9] * s[3] * s[14] - s[13] * s[2] * s[11] + s[13] * s[3] * s[10];
  inv[5] = s[0] * s[10] * s[15] - s[0] * s[11] * s[14] - s[8] * s[2] * s[15] + s[8] * s[3] * s[14] + s[12] * s[2] * s[11] - s[12] * s[3] * s[10];
  inv[9] = -s[0] * s[9] * s[15] + s[0] * s[11] * s[13] + s[8] * s[1] * s[15] - s[8] * s[3] * s[13] - s[12] * s[1] * s[11] + s[12] * s[3] * s[9];
  inv[13] = s[0] * s[9] * s[14] - s[0] * s[10] * s[13] - s[8] * s[1] * s[14] + s[8] * s[2] * s[13] + s[12] * s[1] * s[10] - s[12] * s[2] * s[9];
  inv[2] = s[1] * s[6] * s[15] - s[1] * s[7] * s[14] - s[5] * s[2] * s[15] + s[5] * s[3] * s[14] + s[13] * s[2] * s[7] - s[13] * s[3] * s[6];
  inv[6] = -s[0] * s[6] * s[15] + s[0] * s[7] * s[14] + s[4] * s[2] * s[15] - s[4] * s[3] * s[14] - s[12] * s[2] * s[7] + s[12] * s[3] * s[6];
  inv[10] = s[0] * s[5] * s[15] - s[0] * s[7] * s[13] - s[4] * s[1] * s[15] + s[4] * s[3] * s[13] + s[12] * s[1] * s[7] - s[12] * s[3] * s[5];
  inv[14] = -s[0] * s[5] * s[14] + s[0] * s[6] * s[13] + s[4] * s[1] * s[14] - s[4] * s[2] * s[13] - s[12] * s[1] * s[6] + s[12] * s[2] * s[5];
  inv[3] = -s[1] * s[6] * s[11] + s[1] * s[7] * s[10] + s[5] * s[2] * s[11] - s[5] * s[3] * s[10] - s[9] * s[2] * s[7] + s[9] * s[3] * s[6];
  inv[7] = s[0] * s[6] * s[11] - s[0] * s[7] * s[10] - s[4] * s[2] * s[11] + s[4] * s[3] * s[10] + s[8] * s[2] * s[7] - s[8] * s[3] * s[6];
  inv[11] = -s[0] * s[5] * s[11] + s[0] * s[7] * s[9] + s[4] * s[1] * s[11] - s[4] * s[3] * s[9] - s[8] * s[1] * s[7] + s[8] * s[3] * s[5];
  inv[15] = s[0] * s[5] * s[10] - s[0] * s[6] * s[9] - s[4] * s[1] * s[10] + s[4] * s[2] * s[9] + s[8] * s[1] * s[6] - s[8] * s[2] * s[5];
  double det = s[0] * inv[0] + s[1] * inv[4] + s[2] * inv[8] + s[3] * inv[12];
  for (int i = 0; i < 16; ++i) b[i] = (float)(inv[i] / det);
#endif
  return mat4(b);
}

// ---- samplers: RGBA32F rectangle textures (unnormalised coordinates), CLAMP_TO_BORDER with border 0,
//      NEAREST or LINEAR (GL 3.3 spec 3.8.8 / 3.8.11 order of the bilinear blend); texture buffers for the poses ----
struct sampler2DRect {
  const float* data = nullptr;  // [H][W][C]
  int W = 0, H = 0, C = 4;
  bool linear = false;
  // LINEAR lookups of the generic software GL in FRAGMENT stages (full/sgl.hpp): the texture coordinate comes out of the
  // rasteriser's fp32 interpolation, ~1e-4 texel off the pixel centre it means. Hardware filter weights have 8 fractional
  // bits, so within 1/512 texel of a texel centre a GPU returns that texel itself; so does this flag. (The K5 lookups of
  // the geometry stage keep the full-precision weights DESIGN.md section 2 pins.)
  bool snap8 = false;
  vec4 texel(int i, int j) const {
    if (i < 0 || j < 0 || i >= W || j >= H || !data) return vec4(0.0f);
    const float* p = data + ((size_t)j * W + i) * C;
    return C == 4 ? vec4(p[0], p[1], p[2], p[3]) : vec4(p[0], 0.0f, 0.0f, 1.0f);
  }
};
inline ivec2 textureSize(const sampler2DRect& s) { return ivec2(s.W, s.H); }
inline vec4 texture(const sampler2DRect& s, const vec2& c) {
  if (!s.linear) return s.texel((int)std::floor(c.x), (int)std::floor(c.y));
  // full-precision weights (hardware uses 8 fractional bits; DESIGN.md section 2 lists this as a fixed rule)
  float u = c.x - 0.5f, v = c.y - 0.5f;
  float fu = std::floor(u), fv = std::floor(v);
  int i0 = (int)fu, j0 = (int)fv;
  float a = u - fu, b = v - fv;
  if (s.snap8) {
    int a8 = (int)std::lround((double)a * 256.0), b8 = (int)std::lround((double)b * 256.0);
    if ((a8 == 0 || a8 == 256) && (b8 == 0 || b8 == 256)) return s.texel(i0 + (a8 == 256), j0 + (b8 == 256));
  }
  vec4 t00 = s.texel(i0, j0), t10 = s.texel(i0 + 1, j0), t01 = s.texel(i0, j0 + 1), t11 = s.texel(i0 + 1, j0 + 1);
  vec4 r;
#ifdef REF_MATH_PINNED
  // the association the oracle fixes (oracle/orc_core.c orc_sample): ((w00 t00 + w10 t10) + w01 t01) + w11 t11
  float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
  for (int k = 0; k < 4; ++k) r[k] = ((w00 * t00[k] + w10 * t10[k]) + w01 * t01[k]) + w11 * t11[k];
#else
  double da = a, db = b;
  for (int k = 0; k < 4; ++k)
    r[k] = (float)((1.0 - da) * (1.0 - db) * t00[k] + da * (1.0 - db) * t10[k] + (1.0 - da) * db * t01[k] + da * db * t11[k]);
#endif
  return r;
}
struct samplerBuffer {
  const float* data = nullptr;  // RGBA32F texels
  int n = 0;
};
inline vec4 texelFetch(const samplerBuffer& s, int i) {
  if (i < 0 || i >= s.n) return vec4(0.0f);
  return vec4(s.data[4 * i], s.data[4 * i + 1], s.data[4 * i + 2], s.data[4 * i + 3]);
}

// what every stage inherits
struct StageBase {
  vec4 gl_Position;
  int gl_VertexID = 0;
  struct GlIn { vec4 gl_Position; } gl_in[1];
  bool discarded_ = false;
};

}  // namespace glsl
