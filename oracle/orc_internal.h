/* orc_internal.h -- TEST INFRASTRUCTURE (CPU oracle): helpers shared by the oracle's translation units. */
#ifndef ORC_INTERNAL_H_
#define ORC_INTERNAL_H_

#include "orc_math.h"

#define ORC_DEPTH_CLEAR 0xFFFFFFu /* glClear depth 1.0 in a 24-bit buffer (DEPTH_STENCIL rbo, Preprocessing.cpp:56) */

/* window z in [0,1] -> 24-bit fixed point, round to nearest (GL 4.x spec, section "Controlling the viewport") */
static inline uint32_t orc_depth24(float zw) { return (uint32_t)llrint((double)zw * 16777215.0); }

/* movable classes, color_map.glsl:8-17 as used in gen_vertexmap.vert:96-100 etc.:
 * car 10, bicycle 11, bus 13, motorcycle 15, truck 18, other-vehicle 20, person 30, bicyclist 31, motorcyclist 32 */
static inline int orc_is_movable(float l) {
  return l == 10.0f || l == 11.0f || l == 13.0f || l == 15.0f || l == 18.0f || l == 20.0f || l == 30.0f ||
         l == 31.0f || l == 32.0f;
}

/* exact accumulation: every fp32 term is converted to Q33.30 fixed point (round to nearest even) and summed as
 * int64, which makes the reduction independent of thread/block/GPU order. */
static inline int64_t orc_fix30(float t) { return (int64_t)llrintf(t * 1073741824.0f); }

/* spherical projection used by render_surfels.geom:45-57, gen_indexmap.vert:37-52, update_surfels.vert:77-93,
 * gen_surfels.geom:91-107 (all identical up to the half-pixel snapping done by the callers) */
static inline void orc_project01(orc_v3 p, float fov_up, float fov, float min_depth, float max_depth, float* x,
                                 float* y, float* z) {
  float depth = orc_len3(p);
  float yaw = orc_atan2f(p.y, p.x);
  float pitch = -orc_asinf(p.z / depth);
  *x = 0.5f * ((-yaw * ORC_INV_PI_F) + 1.0f);
  *y = 1.0f - (ORC_RAD2DEG_F * pitch + fov_up) / fov;
  *z = (depth - min_depth) / (max_depth - min_depth);
}

/* pack(), color.glsl:32-38 with round() = ties away from zero */
static inline float orc_pack_rgb(float r, float g, float b) {
  int rgb = (int)roundf(r * 255.0f);
  rgb = (rgb << 8) + (int)roundf(g * 255.0f);
  rgb = (rgb << 8) + (int)roundf(b * 255.0f);
  return (float)rgb;
}

/* Host threads (OpenMP). Every parallel loop of the oracle is written so that its result does not depend on the
 * thread count: integer sums, per-pixel / per-surfel independent work, and depth-tested rasterisation into per-thread
 * targets that are merged in buffer order (orc_map.c). orc_set_threads(1) runs the plain sequential code. */
int orc_threads(void);

#endif
