// ref_pipeline.cpp -- TEST INFRASTRUCTURE (oracle/_ref). The reference's OWN shader programs (transpiled from
// /root/reference/src/shader by glsl2cpp.py into oracle/_ref/gen/) driven by a minimal software GL: exactly the draw
// calls, attachments, uniforms and fixed-function state that core/Preprocessing.cpp, core/Frame2Model.cpp and
// core/SurfelMap.cpp issue (cited per function, paths relative to /root/reference/src). Nothing here is hand-restated
// shader arithmetic: per-point / per-pixel / per-surfel values come out of the reference's shader text.
//
// What this file does fix are the GL rules that the specification leaves to the implementation (the same list as
// DESIGN.md section 2 -- these are common to oracle/ and oracle/_ref by construction, everything else is independent):
//   points       1-pixel points land in the pixel containing the window-space centre; clipped against -w..w
//   depth        24-bit DEPTH_STENCIL renderbuffers: d = round(z_window * (2^24-1)); GL_LESS / GL_LEQUAL; buffer order
//   triangles    vertices snapped to 1/256 pixel, pixel-centre sampling, exact integer edge functions, top-left style
//                tie rule, fp32 barycentric interpolation of z and texCoords, fragments outside 0<=z<=1 clipped
//   blending     GL_ONE,GL_ONE in fp32, fragments in primitive order
//   feedback     transform feedback records primitives in input order
//   textures     RGBA32F rectangle textures, CLAMP_TO_BORDER (0,0,0,0), NEAREST or LINEAR (glsl_types.hpp)
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "glsl_types.hpp"
#include "ref_Frame2Model_jacobians_frag.hpp"
#include "ref_Frame2Model_jacobians_geom.hpp"
#include "ref_Frame2Model_jacobians_vert.hpp"
#include "ref_copy_surfels_geom.hpp"
#include "ref_copy_surfels_vert.hpp"
#include "ref_floodfill_frag.hpp"
#include "ref_gen_indexmap_frag.hpp"
#include "ref_gen_indexmap_vert.hpp"
#include "ref_gen_normalmap_frag.hpp"
#include "ref_gen_surfels_geom.hpp"
#include "ref_gen_surfels_vert.hpp"
#include "ref_gen_vertexmap_frag.hpp"
#include "ref_gen_vertexmap_vert.hpp"
#include "ref_init_radiusConf_frag.hpp"
#include "ref_init_radiusConf_vert.hpp"
#include "ref_render_compose_frag.hpp"
#include "ref_render_surfels_frag.hpp"
#include "ref_render_surfels_geom.hpp"
#include "ref_render_surfels_vert.hpp"
#include "ref_update_surfels_frag.hpp"
#include "ref_update_surfels_geom.hpp"
#include "ref_update_surfels_vert.hpp"

extern "C" {
#include "../orc.h"  // orc_params / orc_surfel: the parameter block and the 64-byte Surfel record shared with the oracle
}

using namespace glsl;

namespace {

// ------------------------------------------------------------------------------------------------ mini GL
typedef std::vector<float> Image;  // [H][W][4]
const uint32_t kDepthClear = 0xFFFFFFu;

inline uint32_t depth24(float zw) { return (uint32_t)llrint((double)zw * 16777215.0); }

struct Target {
  int W = 0, H = 0;
  float* color[3] = {nullptr, nullptr, nullptr};  // RGBA32F attachments (COLOR0..2), may be null = not attached
  std::vector<uint32_t> depth;
  void clear_depth() { depth.assign((size_t)W * H, kDepthClear); }
  void clear_color() {
    for (int a = 0; a < 3; ++a)
      if (color[a]) memset(color[a], 0, sizeof(float) * 4 * (size_t)W * H);
  }
};

// GL_POINTS of size 1: window coordinates from clip coordinates (w = 1 everywhere on this path)
inline bool point_to_pixel(const vec4& pos, int W, int H, int* px, int* py, float* zw) {
  if (!(pos.x >= -pos.w && pos.x <= pos.w && pos.y >= -pos.w && pos.y <= pos.w && pos.z >= -pos.w && pos.z <= pos.w))
    return false;  // clip volume (NaN fails)
  float xw = (0.5f * pos.x + 0.5f) * (float)W, yw = (0.5f * pos.y + 0.5f) * (float)H;
  int ix = (int)std::floor(xw), iy = (int)std::floor(yw);
  if (ix < 0 || iy < 0 || ix >= W || iy >= H) return false;
  *px = ix; *py = iy; *zw = 0.5f * pos.z + 0.5f;
  return true;
}

inline void store4(float* img, size_t pix, const vec4& v) { img[4 * pix] = v.x; img[4 * pix + 1] = v.y; img[4 * pix + 2] = v.z; img[4 * pix + 3] = v.w; }

struct RVert { int64_t X, Y; float z, tx, ty; };
inline int64_t fdiv(int64_t a, int64_t b) { int64_t q = a / b, r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q; }
inline int64_t cdiv(int64_t a, int64_t b) { return -fdiv(-a, b); }
inline int64_t edge(const RVert& P, const RVert& Q, int64_t X, int64_t Y) { return (Q.X - P.X) * (Y - P.Y) - (Q.Y - P.Y) * (X - P.X); }
inline bool edge_inside(int64_t w, const RVert& P, const RVert& Q) {
  if (w != 0) return w > 0;
  int64_t dx = Q.X - P.X, dy = Q.Y - P.Y;
  return dy > 0 || (dy == 0 && dx > 0);
}
inline RVert to_window(const vec4& pos, const vec2& tc, int W, int H) {
  RVert v;
  float xw = (0.5f * pos.x + 0.5f) * (float)W, yw = (0.5f * pos.y + 0.5f) * (float)H;
  v.X = (int64_t)llrintf(xw * 256.0f); v.Y = (int64_t)llrintf(yw * 256.0f);
  v.z = 0.5f * pos.z + 0.5f; v.tx = tc.x; v.ty = tc.y;
  return v;
}
// one triangle; `frag` is called per covered pixel with interpolated (z_window, texCoords) and decides what to write
template <class F>
void raster_triangle(RVert A, RVert B, RVert C, int W, int H, F frag) {
  int64_t area = (B.X - A.X) * (C.Y - A.Y) - (B.Y - A.Y) * (C.X - A.X);
  if (area == 0) return;
  if (area < 0) { std::swap(B, C); area = -area; }
  int64_t minX = std::min(A.X, std::min(B.X, C.X)), maxX = std::max(A.X, std::max(B.X, C.X));
  int64_t minY = std::min(A.Y, std::min(B.Y, C.Y)), maxY = std::max(A.Y, std::max(B.Y, C.Y));
  int64_t i0 = std::max<int64_t>(cdiv(minX - 128, 256), 0), i1 = std::min<int64_t>(fdiv(maxX - 128, 256), W - 1);
  int64_t j0 = std::max<int64_t>(cdiv(minY - 128, 256), 0), j1 = std::min<int64_t>(fdiv(maxY - 128, 256), H - 1);
  float farea = (float)area;
  for (int64_t j = j0; j <= j1; ++j)
    for (int64_t i = i0; i <= i1; ++i) {
      int64_t X = i * 256 + 128, Y = j * 256 + 128;
      int64_t wA = edge(B, C, X, Y), wB = edge(C, A, X, Y), wC = edge(A, B, X, Y);
      if (!edge_inside(wA, B, C) || !edge_inside(wB, C, A) || !edge_inside(wC, A, B)) continue;
      float fB = (float)wB / farea, fC = (float)wC / farea, fA = (1.0f - fB) - fC;
      vec2 tc((fA * A.tx + fB * B.tx) + fC * C.tx, (fA * A.ty + fB * B.ty) + fC * C.ty);
      float z = (fA * A.z + fB * B.z) + fC * C.z;
      frag((int)i, (int)j, z, tc);
    }
}

sampler2DRect tex(const float* img, int W, int H, bool linear = false) {
  sampler2DRect s; s.data = img; s.W = W; s.H = H; s.C = 4; s.linear = linear;
  return s;
}
sampler2DRect tex1(const float* img, int W, int H) {
  sampler2DRect s; s.data = img; s.W = W; s.H = H; s.C = 1; s.linear = false;
  return s;
}

inline double deg2rad(double deg) { return deg * 3.14159265358979323846 / 180.0; }  // rv/Math.h:44-47 (double)
// glow::radians(float) / glow::Radians (SurfelMap.cpp:395, 407): glow is not vendored in the reference; taken as
// deg * (float)(pi/180), the same assumption the oracle makes
inline float glow_radians(float d) { return d * (float)(3.14159265358979323846 / 180.0); }

template <class A, class B>
void copy_block(A& dst, const B& src) {  // matching interface blocks of consecutive stages (same member list)
  static_assert(sizeof(A) == sizeof(B), "interface blocks differ");
  memcpy((void*)&dst, (const void*)&src, sizeof(A));
}

}  // namespace

// =============================================================================== K1-K3  Preprocessing::process
// core/Preprocessing.cpp:120-339 (uniforms :76-117)
extern "C" void ref_preprocess(const orc_params* p, const float* pts4, const float* labels, const float* probs, uint32_t n,
                               uint32_t timestamp, float* vertex_map, float* normal_map, float* semantic_map) {
  const int W = p->data_width, H = p->data_height;
  const size_t P = (size_t)W * H;
  // --- pass 1: depth_program_ = gen_vertexmap.vert + .frag into semanticbuffer_ {COLOR0 = frame.vertex_map, COLOR1 =
  // frame.semantic_map}, depth test GL_LESS, clear colour 0 (Preprocessing.cpp:150-184)
  ref_gen_vertexmap_vert::Shader vs;
  vs.width = (float)W; vs.height = (float)H;                                   // :94-95
  vs.fov_up = std::fabs(p->data_fov_up); vs.fov_down = std::fabs(p->data_fov_down);  // :91-97
  vs.min_depth = p->min_depth; vs.max_depth = p->max_depth;                    // :98-99
  vs.isfirst = timestamp < 10;                                                 // :176-179
  Target fb; fb.W = W; fb.H = H; fb.color[0] = vertex_map; fb.color[1] = semantic_map;
  fb.clear_color(); fb.clear_depth();                                          // :174
  ref_gen_vertexmap_frag::Shader fs;
  for (uint32_t i = 0; i < n; ++i) {
    vs.position = vec4(pts4[4 * i], pts4[4 * i + 1], pts4[4 * i + 2], pts4[4 * i + 3]);
    // attribute 1 / 2: stride 4 bytes, pointer offset 16 / 20 bytes into the label / prob buffers (:142-145):
    // vertex i fetches labels[i+4], probs[i+5]; past the end = 0 (robust buffer access)
    uint32_t li = p->label_offset_quirk ? i + 4 : i, pi = p->label_offset_quirk ? i + 5 : i;
    vs.label = (labels && li < n) ? labels[li] : 0.0f;
    vs.prob = (probs && pi < n) ? probs[pi] : 0.0f;
    vs.main();
    int px, py; float zw;
    if (!point_to_pixel(vs.gl_Position, W, H, &px, &py, &zw)) continue;
    size_t pix = (size_t)py * W + px;
    uint32_t d = depth24(zw);
    if (!(d < fb.depth[pix])) continue;  // GL_LESS (:158)
    fb.depth[pix] = d;
    fs.vertex_coord = vs.vertex_coord; fs.vert_label = vs.vert_label; fs.vert_label_prob = vs.vert_label_prob;
    fs.main();
    store4(vertex_map, pix, fs.color);
    store4(semantic_map, pix, fs.semantic_map);
  }
  // --- pass 2: normal_program_ = empty.vert + quad.geom + gen_normalmap.frag over the whole image into
  // {COLOR0 = frame.normal_map, COLOR1 = erode_semantic_map}; sampler NEAREST / CLAMP_TO_BORDER (:238-277, :69-71)
  Image eroded(P * 4, 0.0f);
  {
    ref_gen_normalmap_frag::Shader f;
    f.vertex_map = tex(vertex_map, W, H); f.semantic_map = tex(semantic_map, W, H);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        f.texCoords = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);  // quad.geom texCoords at the pixel centre
        f.main();
        size_t pix = (size_t)y * W + x;
        store4(normal_map, pix, f.normal);
        store4(eroded.data(), pix, f.eroded_semantic_map);
      }
  }
  // --- pass 3: floodfill_program_ into refine_semantic_map, then frame.semantic_map.copy(refine) (:281-327)
  {
    Image refined(P * 4, 0.0f);
    ref_floodfill_frag::Shader f;
    f.vertex_map = tex(vertex_map, W, H); f.semantic_map = tex(eroded.data(), W, H);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        f.texCoords = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
        f.main();
        store4(refined.data(), (size_t)y * W + x, f.refined_semantic_map);
      }
    memcpy(semantic_map, refined.data(), sizeof(float) * 4 * P);
  }
}

// =============================================================================== K5  Frame2Model::jacobianProducts
// core/Frame2Model.cpp:136-261 (constructor :12-63, updateParameters :65-110). out48 = the 2x8 RGB32F blend texture as
// downloaded (:212-213): JtJ col-major 6x6 | Jtf | valid, F, outlier, inlier_residual, invalid, 0.
extern "C" void ref_icp_jacobian(const orc_params* p, const float* data_v, const float* data_n, const float* data_s,
                                 const float* model_v, const float* model_n, const float* model_s, const double pose[16],
                                 int32_t iteration, float max_distance, float max_angle_deg, int32_t entries_per_kernel,
                                 float out48[48]) {
  const int W = p->data_width, H = p->data_height, Wm = p->model_width, Hm = p->model_height;
  Image zero_d((size_t)W * H * 4, 0.0f), zero_m((size_t)Wm * Hm * 4, 0.0f);
  ref_Frame2Model_jacobians_vert::Shader vs;
  ref_Frame2Model_jacobians_geom::Shader gs;
  ref_Frame2Model_jacobians_frag::Shader fs;
  const bool lin = p->bilinear_sampling != 0;  // :102-109: ONE sampler object bound to all six units
  gs.vertex_model = tex(model_v, Wm, Hm, lin); gs.normal_model = tex(model_n, Wm, Hm, lin);
  gs.vertex_data = tex(data_v, W, H, lin); gs.normal_data = tex(data_n, W, H, lin);
  gs.semantic_model = tex(model_s ? model_s : zero_m.data(), Wm, Hm, lin);
  gs.semantic_data = tex(data_s ? data_s : zero_d.data(), W, H, lin);
  gs.angle_thresh = (float)std::cos(deg2rad((double)max_angle_deg));  // :66  float = std::cos(Math::deg2rad(double))
  gs.distance_thresh = max_distance;                                    // :67
  gs.weight_function = p->weighting; gs.factor = p->factor;             // :69-80
  gs.fov_up = std::fabs(p->data_fov_up); gs.fov_down = std::fabs(p->data_fov_down);  // :82-83
  gs.entries_per_kernel = entries_per_kernel;                           // :60
  float pf[16];
  for (int i = 0; i < 16; ++i) pf[i] = (float)pose[i];                  // :194 pose_.cast<float>()
  gs.pose = mat4(pf);
  gs.iteration = iteration;                                             // :195
  float blend[8][2][3];                                                 // fbo_blend_: 2 wide, 8 high, RGB32F, cleared to 0
  memset(blend, 0, sizeof(blend));
  // vbo_img_coords_ (:26-30): x-blocks outer, rows inner
  for (uint32_t i = 0; i < (uint32_t)W; i += (uint32_t)entries_per_kernel)
    for (uint32_t j = 0; j < (uint32_t)H; ++j) {
      vs.texCoords = vec2((float)i + 0.5f, (float)j + 0.5f);
      vs.main();
      copy_block(gs.gs_in[0], vs.vs_out);
      gs.gl_in[0].gl_Position = vs.gl_Position;
      gs.emitted_.clear(); gs.prim_end_.clear();
      gs.main();
      for (const auto& e : gs.emitted_) {  // GL_POINTS out, additive blending GL_ONE, GL_ONE (:189-190)
        int px, py; float zw;
        if (!point_to_pixel(e.gl_Position, 2, 8, &px, &py, &zw)) continue;
        fs.values = e.values;
        fs.main();
        for (int c = 0; c < 3; ++c) blend[py][px][c] = blend[py][px][c] + fs.result[c];
      }
    }
  memcpy(out48, blend, sizeof(blend));  // download(PixelFormat::RGB): row-major, bottom row first
}

// =============================================================================== SurfelMap
struct ref_map {
  orc_params p;
  uint32_t timestamp = 0;
  std::vector<orc_surfel> surfels, updated, data_surfels;
  std::vector<float> poses;  // poseBuffer_: maxPoses_ x mat4 (SurfelMap.h:205-208)
  Image old_v, old_n, old_s, new_v, new_n, new_s, comp_v, comp_s, comp_n;
  std::vector<float> index_map;  // R32F (SurfelMap.cpp:15)
  Image index_v, index_n, radius_map, centerized, integrated;
  Target render_fb;
  float* color2 = nullptr;  // whatever is attached to COLOR2 of renderFramebuffer_ (never detached: SURVEY Q4)
  // derived uniforms, SurfelMap.cpp:336-457
  float pixel_size, p_unstable, log_prior, log_unstable, radconf_angle_thresh, update_angle_thresh;
};

static const uint32_t kMaxPoses = 10000;      // SurfelMap.h:205
static const uint32_t kComposeSurfelAge = 100;  // SurfelMap.h:144

extern "C" ref_map* ref_map_create(const orc_params* p) {
  ref_map* m = new ref_map();
  m->p = *p;
  size_t Pm = (size_t)p->model_width * p->model_height, Pd = (size_t)p->data_width * p->data_height;
  for (Image* im : {&m->old_v, &m->old_n, &m->old_s, &m->new_v, &m->new_n, &m->new_s, &m->comp_v, &m->comp_n, &m->comp_s})
    im->assign(Pm * 4, 0.0f);
  for (Image* im : {&m->index_v, &m->index_n, &m->radius_map, &m->centerized, &m->integrated}) im->assign(Pd * 4, 0.0f);
  m->index_map.assign(Pd, 0.0f);
  m->poses.assign((size_t)kMaxPoses * 16, 0.0f);
  for (uint32_t t = 0; t < kMaxPoses; ++t)
    for (int i = 0; i < 4; ++i) m->poses[16 * t + 5 * i] = 1.0f;  // reset(): identity (SurfelMap.cpp:479-481)
  // setParameters, SurfelMap.cpp:339-350, 395, 407
  float vfov = std::abs(p->data_fov_up) + std::abs(p->data_fov_down), hfov = 360.0f;
  float vpix = std::tan(0.5f * deg2rad(vfov) / (uint32_t)p->data_height) * 1.0f;
  float hpix = std::tan(0.5f * deg2rad(hfov) / (uint32_t)p->data_width) * 1.0f;
  m->pixel_size = std::max(vpix, hpix);
  float p_stable = p->p_stable, p_prior = p->p_prior;
  m->p_unstable = 1.0f - p_stable;
  m->log_prior = std::log(p_prior / (1.0 - p_prior));
  m->log_unstable = std::log(m->p_unstable / (1.0 - m->p_unstable));
  m->radconf_angle_thresh = std::cos(glow_radians(p->max_angle));    // :395 std::cos(radians(float))
  m->update_angle_thresh = std::sin(glow_radians(p->map_max_angle));  // :407 std::sin(Radians(float))
  m->render_fb.W = p->model_width; m->render_fb.H = p->model_height;
  return m;
}
extern "C" void ref_map_destroy(ref_map* m) { delete m; }
extern "C" uint32_t ref_map_size(const ref_map* m) { return (uint32_t)m->surfels.size(); }
extern "C" uint32_t ref_map_timestamp(const ref_map* m) { return m->timestamp; }
extern "C" uint32_t ref_map_download(const ref_map* m, orc_surfel* dst, uint32_t cap) {
  uint32_t k = std::min<uint32_t>(cap, (uint32_t)m->surfels.size());
  memcpy(dst, m->surfels.data(), sizeof(orc_surfel) * k);
  return k;
}
extern "C" void ref_map_upload(ref_map* m, const orc_surfel* src, uint32_t n, uint32_t timestamp) {
  m->surfels.assign(src, src + n);
  m->timestamp = timestamp;
}
extern "C" void ref_map_set_pose(ref_map* m, uint32_t t, const float pose[16]) {
  if (t < kMaxPoses) memcpy(&m->poses[16 * (size_t)t], pose, 64);
}
extern "C" void ref_map_derived(const ref_map* m, float out[6]) {
  out[0] = m->pixel_size; out[1] = m->p_unstable; out[2] = m->log_prior; out[3] = m->log_unstable;
  out[4] = m->radconf_angle_thresh; out[5] = m->update_angle_thresh;
}

namespace {

samplerBuffer pose_texture(const ref_map* m) {
  samplerBuffer s; s.data = m->poses.data(); s.n = (int)kMaxPoses * 4;
  return s;
}

// Eigen's pose.inverse() (SurfelMap.cpp:496, 873 ...): general 4x4 inverse of a rigid pose; precision is the library's.
mat4 eigen_inverse(const float pose[16]) { return inverse(mat4(pose)); }

// one glDrawArrays(GL_POINTS, 0, surfels_.size()) with render_program_ into renderFramebuffer_ (COLOR0..2 + depth)
void draw_surfels(ref_map* m, const float inv_pose[16], float conf_threshold, int timestamp_threshold, bool render_old,
                  bool lequal) {
  const orc_params& p = m->p;
  Target& fb = m->render_fb;
  const int W = fb.W, H = fb.H;
  ref_render_surfels_vert::Shader vs;
  ref_render_surfels_geom::Shader gs;
  ref_render_surfels_frag::Shader fs;
  vs.poseBuffer = pose_texture(m);
  vs.inv_pose = mat4(inv_pose);
  gs.fov_up = std::abs(p.model_fov_up); gs.fov_down = std::abs(p.model_fov_down);  // SurfelMap.cpp:437-438
  gs.min_depth = p.model_min_depth; gs.max_depth = p.model_max_depth;            // :439-440
  gs.use_stability = p.use_stability != 0;                                       // :442
  gs.conf_threshold = conf_threshold; gs.timestamp_threshold = timestamp_threshold; gs.render_old_surfels = render_old;
  for (size_t k = 0; k < m->surfels.size(); ++k) {
    const orc_surfel& s = m->surfels[k];
    vs.gl_VertexID = (int)k;
    vs.position_radius = vec4(s.x, s.y, s.z, s.radius);
    vs.normal_confidence = vec4(s.nx, s.ny, s.nz, s.confidence);
    vs.timestamp = (int)s.timestamp;
    vs.surfel_color_weight_count = vec3(s.color, s.weight, s.count);
    vs.surfel_semantic_map = vec4(s.r, s.g, s.b, s.w);
    vs.main();
    copy_block(gs.gs_in[0], vs.vs_out);
    gs.emitted_.clear(); gs.prim_end_.clear();
    gs.main();
    if (gs.emitted_.size() != 4) continue;
    // triangle strip of 4 vertices = 2 triangles; flat data (vertex / normal / semantic / confidence) is identical on all
    // four vertices (render_surfels.geom:95-99 sets them once)
    RVert q[4];
    for (int i = 0; i < 4; ++i) q[i] = to_window(gs.emitted_[i].gl_Position, gs.emitted_[i].texCoords, W, H);
    const auto e0 = gs.emitted_[0];
    auto frag = [&](int i, int j, float z, const vec2& tc) {
      fs.texCoords = tc; fs.vertex = e0.vertex; fs.normal = e0.normal; fs.semantic = e0.semantic; fs.confidence = e0.confidence;
      fs.discarded_ = false;
      fs.main();
      if (fs.discarded_) return;
      if (!(z >= 0.0f && z <= 1.0f)) return;  // clipped against the near / far planes
      size_t pix = (size_t)j * W + i;
      uint32_t d = depth24(z);
      if (!(lequal ? d <= fb.depth[pix] : d < fb.depth[pix])) return;
      fb.depth[pix] = d;
      if (fb.color[0]) store4(fb.color[0], pix, fs.vertexmap);
      if (fb.color[1]) store4(fb.color[1], pix, fs.normalmap);
      if (fb.color[2]) store4(fb.color[2], pix, fs.semantic_rangemap);
    };
    raster_triangle(q[0], q[1], q[2], W, H, frag);
    raster_triangle(q[1], q[2], q[3], W, H, frag);
  }
}

void attach(ref_map* m, Image& v, Image& n, Image* s) {
  m->render_fb.color[0] = v.data(); m->render_fb.color[1] = n.data();
  if (s) m->color2 = s->data();          // COLOR2 keeps its last attachment when a caller attaches only COLOR0/1
  m->render_fb.color[2] = m->color2;
}
void clear(ref_map* m) { m->render_fb.clear_color(); m->render_fb.clear_depth(); }

}  // namespace

// SurfelMap::render, SurfelMap.cpp:847-1021
extern "C" void ref_map_render(ref_map* m, const float pose_old[16], const float pose_new[16], float conf_thr, float* frame_v,
                               float* frame_n, float* frame_s) {
  const int W = m->p.model_width, H = m->p.model_height;
  const size_t P = (size_t)W * H;
  float inv_old[16], inv_new[16];
  eigen_inverse(pose_old).store(inv_old); eigen_inverse(pose_new).store(inv_new);
  if (m->p.compose_rendering) {
    int thr = (int)(m->timestamp - kComposeSurfelAge);  // :873 unsigned arithmetic, then GlUniform<int>
    attach(m, m->old_v, m->old_n, &m->old_s); clear(m);
    draw_surfels(m, inv_old, conf_thr, thr, true, false);                       // :876-882
    attach(m, m->new_v, m->new_n, &m->new_s); clear(m);
    draw_surfels(m, inv_new, conf_thr, thr, false, false);                      // :884-891
    attach(m, m->comp_v, m->comp_n, &m->comp_s); clear(m);
    draw_surfels(m, inv_old, conf_thr, thr, true, false);                       // :893-902 (depth func still GL_LESS)
    draw_surfels(m, inv_new, conf_thr, thr, false, false);                      // :904-908 "not clearing!"
    // compose_program_ over the whole image into frame (:913-941)
    Image fv(P * 4), fn(P * 4), fsm(P * 4);
    m->render_fb.color[0] = fv.data(); m->render_fb.color[1] = fn.data(); m->render_fb.color[2] = fsm.data();
    ref_render_compose_frag::Shader f;
    f.old_vertexmap = tex(m->old_v.data(), W, H); f.old_normalmap = tex(m->old_n.data(), W, H);
    f.new_vertexmap = tex(m->new_v.data(), W, H); f.new_normalmap = tex(m->new_n.data(), W, H);
    // NOTE the sampler *uniforms* of compose_program_ are set in the constructor (SurfelMap.cpp:243-248); the texture
    // units bound here are 0,1 = old v/n; 2,3 = new v/n; 4 = new semantic; 6 = old semantic (:918-929)
    f.new_semanticmap = tex(m->new_s.data(), W, H); f.old_semanticmap = tex(m->old_s.data(), W, H);
    f.max_distance = m->p.max_loop_closure_distance;  // :446
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        f.texCoords = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
        f.main();
        size_t pix = (size_t)y * W + x;
        store4(fv.data(), pix, f.vertexmap); store4(fn.data(), pix, f.normalmap); store4(fsm.data(), pix, f.semanticmap);
      }
    memcpy(frame_v, fv.data(), P * 16); memcpy(frame_n, fn.data(), P * 16); memcpy(frame_s, fsm.data(), P * 16);
    // COLOR2 now points at frame.semantic_map, a texture of the CALLER's Frame (:915), and stays attached: the later
    // two-attachment renders (render_active/_inactive/_composed) put their semantic output there, so the semantic images
    // of SurfelMap's own old/new/composed frames keep what render() wrote.
    m->color2 = nullptr;  // writes to the caller's frame are not observable through SurfelMap's own frames
    m->render_fb.color[2] = nullptr;
  } else {
    Image fv(P * 4), fn(P * 4), fsm(P * 4);
    m->render_fb.color[0] = fv.data(); m->render_fb.color[1] = fn.data(); m->render_fb.color[2] = fsm.data();
    clear(m);
    draw_surfels(m, inv_old, conf_thr, 0, false, false);  // :983-1003
    memcpy(frame_v, fv.data(), P * 16); memcpy(frame_n, fn.data(), P * 16); memcpy(frame_s, fsm.data(), P * 16);
    m->new_v = fv; m->new_n = fn; m->new_s = fsm; m->old_v = fv; m->old_n = fn; m->old_s = fsm;  // :1016-1017
    m->color2 = nullptr; m->render_fb.color[2] = nullptr;
  }
}
// render_active / render_inactive / render_composed, SurfelMap.cpp:1023-1165: COLOR0/1 only
extern "C" void ref_map_render_active(ref_map* m, const float pose[16], float conf_thr) {
  float inv[16]; eigen_inverse(pose).store(inv);
  attach(m, m->new_v, m->new_n, nullptr); clear(m);
  draw_surfels(m, inv, conf_thr, (int)(m->timestamp - kComposeSurfelAge), false, false);
}
extern "C" void ref_map_render_inactive(ref_map* m, const float pose[16], float conf_thr) {
  float inv[16]; eigen_inverse(pose).store(inv);
  attach(m, m->old_v, m->old_n, nullptr); clear(m);
  draw_surfels(m, inv, conf_thr, (int)(m->timestamp - kComposeSurfelAge), true, false);
}
extern "C" void ref_map_render_composed(ref_map* m, const float pose_old[16], const float pose_new[16], float conf_thr) {
  float io[16], in_[16]; eigen_inverse(pose_old).store(io); eigen_inverse(pose_new).store(in_);
  attach(m, m->comp_v, m->comp_n, nullptr); clear(m);
  int thr = (int)(m->timestamp - kComposeSurfelAge);
  draw_surfels(m, io, conf_thr, thr, true, true);    // glDepthFunc(GL_LEQUAL), :1126
  draw_surfels(m, in_, conf_thr, thr, false, true);  // "not clearing!"
}
extern "C" void ref_map_get_frame(const ref_map* m, int which, float* v, float* n, float* s) {
  const Image& V = which == 0 ? m->old_v : which == 1 ? m->new_v : m->comp_v;
  const Image& N = which == 0 ? m->old_n : which == 1 ? m->new_n : m->comp_n;
  const Image& S = which == 0 ? m->old_s : which == 1 ? m->new_s : m->comp_s;
  if (v) memcpy(v, V.data(), V.size() * 4);
  if (n) memcpy(n, N.data(), N.size() * 4);
  if (s) memcpy(s, S.data(), S.size() * 4);
}

// SurfelMap::update, SurfelMap.cpp:492-584 WITHOUT updateActiveSubmaps (:556; host-side paging logic, not a shader)
extern "C" void ref_map_update(ref_map* m, const float pose[16], const float* frame_v, const float* frame_n,
                               const float* frame_s) {
  const orc_params& p = m->p;
  const int W = p.data_width, H = p.data_height;
  const size_t P = (size_t)W * H;
  if (m->timestamp < kMaxPoses) memcpy(&m->poses[16 * (size_t)m->timestamp], pose, 64);  // :494-495
  float inv_pose[16];
  eigen_inverse(pose).store(inv_pose);  // :497
  const float fov_up = std::abs(p.data_fov_up), fov_down = std::abs(p.data_fov_down);
  // ---- renderIndexmap (:586-604): indexMap_program_ into {COLOR0 indexMap_ (R32F), COLOR1, COLOR2} + depth GL_LESS
  {
    ref_gen_indexmap_vert::Shader vs;
    ref_gen_indexmap_frag::Shader fs;
    vs.poseBuffer = pose_texture(m);
    vs.fov_up = fov_up; vs.fov_down = fov_down; vs.min_depth = p.min_depth; vs.max_depth = p.max_depth;  // SurfelMap.cpp:283-290
    vs.width = (float)W; vs.height = (float)H;
    vs.pose = mat4(pose); vs.inv_pose = mat4(inv_pose);
    std::fill(m->index_map.begin(), m->index_map.end(), 0.0f);
    std::fill(m->index_v.begin(), m->index_v.end(), 0.0f); std::fill(m->index_n.begin(), m->index_n.end(), 0.0f);
    std::vector<uint32_t> depth(P, kDepthClear);
    for (size_t k = 0; k < m->surfels.size(); ++k) {
      const orc_surfel& s = m->surfels[k];
      vs.gl_VertexID = (int)k;
      vs.surfel_position_radius = vec4(s.x, s.y, s.z, s.radius);
      vs.surfel_normal_confidence = vec4(s.nx, s.ny, s.nz, s.confidence);
      vs.surfel_timestamp = (int)s.timestamp;
      vs.surfel_color_weight_count = vec3(s.color, s.weight, s.count);
      vs.sfl_semantic_map = vec4(s.r, s.g, s.b, s.w);
      vs.main();
      int px, py; float zw;
      if (!point_to_pixel(vs.gl_Position, W, H, &px, &py, &zw)) continue;
      size_t pix = (size_t)py * W + px;
      uint32_t d = depth24(zw);
      if (!(d < depth[pix])) continue;
      depth[pix] = d;
      fs.index = vs.index; fs.vertex = vs.vertex; fs.normal = vs.normal;
      fs.main();
      m->index_map[pix] = fs.indexmap;
      store4(m->index_v.data(), pix, fs.vertexmap); store4(m->index_n.data(), pix, fs.normalmap);
    }
  }
  // ---- generateDataSurfels (:606-619): radConf_program_ over vbo_img_coords_ into {centerizedVertexMap_, radiusConfidenceMap_}
  // vbo_img_coords_ (SurfelMap.cpp:86-93): x outer, y inner, (x + 0.5, y + 0.5)
  {
    ref_init_radiusConf_vert::Shader vs;
    ref_init_radiusConf_frag::Shader fs;
    vs.vertex_map = tex(frame_v, W, H); vs.normal_map = tex(frame_n, W, H);
    vs.fov_up = fov_up; vs.fov_down = fov_down; vs.min_depth = p.min_depth; vs.max_depth = p.max_depth;
    vs.pixel_size = m->pixel_size; vs.confidence_mode = p.confidence_mode;
    vs.min_radius = p.min_radius; vs.max_radius = p.max_radius; vs.angle_thresh = m->radconf_angle_thresh;  // :384-397
    std::fill(m->radius_map.begin(), m->radius_map.end(), 0.0f);
    std::fill(m->centerized.begin(), m->centerized.end(), 0.0f);
    for (int x = 0; x < W; ++x)
      for (int y = 0; y < H; ++y) {
        vs.img_coords = vec2((float)x + 0.5f, (float)y + 0.5f);
        vs.centerized_vertex = vec4(0.0f);  // a fresh invocation: outputs not written are undefined; 0 here
        vs.main();
        int px, py; float zw;
        if (!point_to_pixel(vs.gl_Position, W, H, &px, &py, &zw)) continue;
        fs.valid = vs.valid; fs.centerized_vertex = vs.centerized_vertex; fs.radius = vs.radius; fs.confidence = vs.confidence;
        fs.main();
        size_t pix = (size_t)py * W + px;  // depth test enabled, GL_LESS, z = 0 -> every pixel is hit exactly once
        store4(m->centerized.data(), pix, fs.centerized_vertex_map);
        store4(m->radius_map.data(), pix, fs.radius_confidence_map);
      }
  }
  // ---- updateSurfels part 1 (:621-643): update_program_ with transform feedback -> updated_surfels_, and
  // updateFramebuffer_ {COLOR0 = measurementIntegrated_} + depth GL_LESS
  {
    ref_update_surfels_vert::Shader vs;
    ref_update_surfels_geom::Shader gs;
    ref_update_surfels_frag::Shader fs;
    vs.poseBuffer = pose_texture(m);
    vs.vertex_map = tex(frame_v, W, H); vs.normal_map = tex(frame_n, W, H);
    vs.radiusConfidence_map = tex(m->radius_map.data(), W, H);
    vs.index_map = tex1(m->index_map.data(), W, H);
    vs.semantic_map_in = tex(frame_s, W, H);
    vs.fov_up = fov_up; vs.fov_down = fov_down; vs.min_depth = p.min_depth; vs.max_depth = p.max_depth;  // :399-402
    vs.width = (float)W; vs.height = (float)H; vs.pixel_size = m->pixel_size;                            // :403-405
    vs.distance_thresh = p.map_max_distance; vs.angle_thresh = m->update_angle_thresh;                   // :406-407
    vs.confidence_mode = p.confidence_mode; vs.unstable_age = p.unstable_age;                            // :408, :413
    vs.p_stable = p.p_stable; vs.p_unstable = m->p_unstable; vs.p_prior = p.p_prior;
    vs.log_prior = m->log_prior; vs.log_unstable = m->log_unstable;
    vs.sigma_angle = p.sigma_angle; vs.sigma_distance = p.sigma_distance;
    vs.confidence_threshold = p.confidence_threshold;
    vs.min_radius = 0.0f;  // :422 -- update_program_'s min_radius is never overwritten from the parameters
    vs.max_weight = p.max_weight; vs.weighting_scheme = p.weighting_scheme; vs.averaging_scheme = p.averaging_scheme;
    vs.update_always = p.update_always != 0; vs.active_timestamps = p.active_timestamps;
    vs.use_stability = p.use_stability != 0;
    vs.pose = mat4(pose); vs.inv_pose = mat4(inv_pose); vs.timestamp = (int)m->timestamp;  // :627-629
    std::fill(m->integrated.begin(), m->integrated.end(), 0.0f);
    std::vector<uint32_t> depth(P, kDepthClear);
    m->updated.clear();
    for (size_t k = 0; k < m->surfels.size(); ++k) {
      const orc_surfel& s = m->surfels[k];
      vs.gl_VertexID = (int)k;
      vs.surfel_position_radius = vec4(s.x, s.y, s.z, s.radius);
      vs.surfel_normal_confidence = vec4(s.nx, s.ny, s.nz, s.confidence);
      vs.surfel_timestamp = (int)s.timestamp;
      vs.surfel_color_weight_count = vec3(s.color, s.weight, s.count);
      vs.surfel_semantic_map = vec4(s.r, s.g, s.b, s.w);
      vs.main();
      copy_block(gs.gs_in[0], vs.vs_out);
      gs.gl_in[0].gl_Position = vs.gl_Position;
      gs.emitted_.clear(); gs.prim_end_.clear();
      gs.main();
      for (const auto& e : gs.emitted_) {
        orc_surfel o;  // transform-feedback varyings, SurfelMap.cpp:38-40
        o.x = e.sfl_position_radius.x; o.y = e.sfl_position_radius.y; o.z = e.sfl_position_radius.z; o.radius = e.sfl_position_radius.w;
        o.nx = e.sfl_normal_confidence.x; o.ny = e.sfl_normal_confidence.y; o.nz = e.sfl_normal_confidence.z; o.confidence = e.sfl_normal_confidence.w;
        o.timestamp = (uint32_t)e.sfl_timestamp;
        o.color = e.sfl_color_weight_count.x; o.weight = e.sfl_color_weight_count.y; o.count = e.sfl_color_weight_count.z;
        o.r = e.sfl_semantic_map.x; o.g = e.sfl_semantic_map.y; o.b = e.sfl_semantic_map.z; o.w = e.sfl_semantic_map.w;
        m->updated.push_back(o);
        int px, py; float zw;
        if (!point_to_pixel(e.gl_Position, W, H, &px, &py, &zw)) continue;
        size_t pix = (size_t)py * W + px;
        uint32_t d = depth24(zw);
        if (!(d < depth[pix])) continue;
        depth[pix] = d;
        fs.main();
        store4(m->integrated.data(), pix, fs.color);
      }
    }
  }
  // ---- updateSurfels part 2 (:645-664): initialize_program_ (gen_surfels) with transform feedback -> data_surfels_
  {
    ref_gen_surfels_vert::Shader vs;
    ref_gen_surfels_geom::Shader gs;
    {
      vs.vertex_map = tex(frame_v, W, H); vs.normal_map = tex(frame_n, W, H);
      vs.radiusConfidence_map = tex(m->radius_map.data(), W, H);
      vs.measurementIntegrated_map = tex(m->integrated.data(), W, H);
      gs.vertex_map = vs.vertex_map; gs.normal_map = vs.normal_map; gs.radiusConfidence_map = vs.radiusConfidence_map;
      gs.semantic_map = tex(frame_s, W, H);                 // unit 6 = frame.semantic_map (:363, :523-524)
      gs.model_semantic_map = tex(m->new_s.data(), p.model_width, p.model_height);  // unit 9: nothing is bound there; the value is unused
      gs.log_prior = m->log_prior; vs.log_prior = m->log_prior;
      gs.fov_up = fov_up; gs.fov_down = fov_down; gs.min_depth = p.min_depth; gs.max_depth = p.max_depth;
      gs.width = (float)W; gs.height = (float)H; gs.pixel_size = m->pixel_size;
      gs.pose = mat4(pose); gs.inv_pose = mat4(inv_pose); gs.timestamp = (int)m->timestamp;  // :650-652
      vs.pose = gs.pose; vs.timestamp = gs.timestamp;
    }
    m->data_surfels.clear();
    for (int x = 0; x < W; ++x)
      for (int y = 0; y < H; ++y) {
        vs.img_coords = vec2((float)x + 0.5f, (float)y + 0.5f);
        vs.main();
        copy_block(gs.gs_in[0], vs.vs_out);
        gs.emitted_.clear(); gs.prim_end_.clear();
        gs.main();
        for (const auto& e : gs.emitted_) {
          orc_surfel o;
          o.x = e.sfl_position_radius.x; o.y = e.sfl_position_radius.y; o.z = e.sfl_position_radius.z; o.radius = e.sfl_position_radius.w;
          o.nx = e.sfl_normal_confidence.x; o.ny = e.sfl_normal_confidence.y; o.nz = e.sfl_normal_confidence.z; o.confidence = e.sfl_normal_confidence.w;
          o.timestamp = (uint32_t)e.sfl_timestamp;
          o.color = e.sfl_color_weight_count.x; o.weight = e.sfl_color_weight_count.y; o.count = e.sfl_color_weight_count.z;
          o.r = e.sfl_semantic_map.x; o.g = e.sfl_semantic_map.y; o.b = e.sfl_semantic_map.z; o.w = e.sfl_semantic_map.w;
          m->data_surfels.push_back(o);
        }
      }
  }
  // ---- copySurfels (:666-698): copy_program_, updated_surfels_ then data_surfels_ -> surfels_
  {
    ref_copy_surfels_vert::Shader vs;
    ref_copy_surfels_geom::Shader gs;
    vs.poseBuffer = pose_texture(m);
    // submap origin (0,0): the harness never shifts submaps; extent :673-678
    vs.submap_center = vec2(0.0f, 0.0f);
    vs.submap_extent = 2.0f * (float)p.submap_dimension * p.submap_extent + p.submap_extent;
    std::vector<orc_surfel> out;
    for (const std::vector<orc_surfel>* src : {&m->updated, &m->data_surfels})
      for (size_t k = 0; k < src->size(); ++k) {
        const orc_surfel& s = (*src)[k];
        vs.gl_VertexID = (int)k;
        vs.position_radius = vec4(s.x, s.y, s.z, s.radius);
        vs.normal_confidence = vec4(s.nx, s.ny, s.nz, s.confidence);
        vs.in_timestamp = (int)s.timestamp;
        vs.surfel_color_weight_count = vec3(s.color, s.weight, s.count);
        vs.sfl_semantic_map = vec4(s.r, s.g, s.b, s.w);
        vs.main();
        copy_block(gs.gs_in[0], vs.vs_out);
        gs.emitted_.clear(); gs.prim_end_.clear();
        gs.main();
        for (const auto& e : gs.emitted_) {
          orc_surfel o;
          o.x = e.sfl_position_radius.x; o.y = e.sfl_position_radius.y; o.z = e.sfl_position_radius.z; o.radius = e.sfl_position_radius.w;
          o.nx = e.sfl_normal_confidence.x; o.ny = e.sfl_normal_confidence.y; o.nz = e.sfl_normal_confidence.z; o.confidence = e.sfl_normal_confidence.w;
          o.timestamp = (uint32_t)e.sfl_timestamp;
          o.color = e.sfl_color_weight_count.x; o.weight = e.sfl_color_weight_count.y; o.count = e.sfl_color_weight_count.z;
          o.r = e.sfl_semantic_map.x; o.g = e.sfl_semantic_map.y; o.b = e.sfl_semantic_map.z; o.w = e.sfl_semantic_map.w;
          out.push_back(o);
        }
      }
    m->surfels.swap(out);
  }
  m->timestamp += 1;  // :577
}
extern "C" void ref_map_get_update_debug(const ref_map* m, uint32_t* index_map, float* radius_map, uint8_t* integrated,
                                         uint32_t* n_updated, uint32_t* n_new) {
  size_t Pd = (size_t)m->p.data_width * m->p.data_height;
  if (index_map) for (size_t i = 0; i < Pd; ++i) index_map[i] = (uint32_t)m->index_map[i];
  if (radius_map) memcpy(radius_map, m->radius_map.data(), Pd * 16);
  if (integrated) for (size_t i = 0; i < Pd; ++i) integrated[i] = m->integrated[4 * i] > 0.5f;
  if (n_updated) *n_updated = (uint32_t)m->updated.size();
  if (n_new) *n_new = (uint32_t)m->data_surfels.size();
}

extern "C" const char* ref_math_mode() {
#ifdef REF_MATH_PINNED
  return "pinned";
#else
  return "precise";
#endif
}
