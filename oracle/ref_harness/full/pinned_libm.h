// pinned_libm.h -- TEST INFRASTRUCTURE (oracle/_ref/libsuma_ref_full.so). Force-included into core/lie_algebra.cpp ONLY:
// SE3::exp / SE3::log call libm's sin / cos, whose last-bit rounding is the C library's. For bit-for-bit comparisons of
// whole runs the full build evaluates them with the fp64 polynomial the oracle pins (oracle/orc_math.h orc_sincos) --
// the same treatment REF_MATH_PINNED gives the GLSL built-ins. The unpinned behaviour is what libsuma_ref_host.so
// measures (tests/test_ref_host.py: 1e-13 of the matrix scale).
#pragma once
#include <cmath>
extern "C" {
#include "../../orc_math.h"
}
inline double refpin_sin(double x) { double s, c; orc_sincos(x, &s, &c); return s; }
inline double refpin_cos(double x) { double s, c; orc_sincos(x, &s, &c); return c; }
namespace std { using ::refpin_sin; using ::refpin_cos; }
#define sin refpin_sin
#define cos refpin_cos
