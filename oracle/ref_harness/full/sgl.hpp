// sgl.hpp -- TEST INFRASTRUCTURE (oracle/_ref/libsuma_ref_full.so). A software OpenGL 3.3 subset: exactly the state and
// draw behaviour the reference's core classes (core/Preprocessing.cpp, Frame2Model.cpp, SurfelMap.cpp, SurfelMapping.cpp)
// use, so that those files -- compiled where they lie against the stand-in `glow` headers of this directory -- drive the
// reference's own shaders (transpiled by ../glsl2cpp.py, adapters generated under -DSGL_REFLECT) on the CPU.
//
// Nothing here knows about SuMa: programs, attributes, uniforms, samplers, framebuffers and transform feedback are
// resolved generically from what the reference's C++ sets up through glow. The rules the GL specification leaves to the
// implementation are the ones listed in DESIGN.md section 2 and implemented identically in ../ref_pipeline.cpp:
//   points       1-pixel points: the pixel containing the window-space centre; clip volume -w..w
//   depth        24-bit: d = round(z_window * (2^24-1)); GL_LESS / GL_LEQUAL; primitives in buffer order
//   triangles    vertices snapped to 1/256 px, pixel-centre sampling, exact integer edge functions, one tie rule, fp32
//                barycentric attributes ((fA a + fB b) + fC c; a varying equal on all three vertices stays exactly that
//                value -- plane-equation hardware), fragments outside 0 <= z <= 1 clipped
//   blending     GL_ONE, GL_ONE in fp32 in primitive order
//   feedback     records primitives in input order (interleaved attributes)
//   textures     sampler object state overrides texture state; no mipmaps, level of detail 0 => magnification filter
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "glsl_types.hpp"

namespace sgl {

// ------------------------------------------------------------------------------------------------ stage adapters
struct UVal {
  enum Kind { F, I, B, U, V2, V3, V4, M4 } kind = F;
  float f[16] = {0};
  int i = 0;
  int as_int() const { return (kind == F) ? (int)f[0] : i; }
  float as_float() const { return (kind == F) ? f[0] : (float)i; }
};
inline void assign_uniform(float& d, const UVal& v) { d = v.as_float(); }
inline void assign_uniform(int& d, const UVal& v) { d = v.as_int(); }
inline void assign_uniform(unsigned& d, const UVal& v) { d = (unsigned)v.as_int(); }
inline void assign_uniform(bool& d, const UVal& v) { d = v.as_int() != 0; }
inline void assign_uniform(glsl::vec2& d, const UVal& v) { d = glsl::vec2(v.f[0], v.f[1]); }
inline void assign_uniform(glsl::vec3& d, const UVal& v) { d = glsl::vec3(v.f[0], v.f[1], v.f[2]); }
inline void assign_uniform(glsl::vec4& d, const UVal& v) { d = glsl::vec4(v.f[0], v.f[1], v.f[2], v.f[3]); }
inline void assign_uniform(glsl::mat4& d, const UVal& v) { d = glsl::mat4(v.f); }  // column-major, as glUniformMatrix4fv

struct AttrVal {
  float f[4] = {0, 0, 0, 1};
  int i[4] = {0, 0, 0, 1};
  bool integer = false;
};
inline void load_attr(float& d, const AttrVal& a) { d = a.integer ? (float)a.i[0] : a.f[0]; }
inline void load_attr(int& d, const AttrVal& a) { d = a.integer ? a.i[0] : (int)a.f[0]; }
inline void load_attr(glsl::vec2& d, const AttrVal& a) { d = glsl::vec2(a.f[0], a.f[1]); }
inline void load_attr(glsl::vec3& d, const AttrVal& a) { d = glsl::vec3(a.f[0], a.f[1], a.f[2]); }
inline void load_attr(glsl::vec4& d, const AttrVal& a) { d = glsl::vec4(a.f[0], a.f[1], a.f[2], a.f[3]); }

struct Units {
  virtual ~Units() {}
  virtual glsl::sampler2DRect rect(int unit) const = 0;
  virtual glsl::samplerBuffer buffer(int unit) const = 0;
};

// one interface variable (or interface block) of a stage: name, storage, size; ncomp > 0 for float / vecN (interpolated)
struct Var {
  std::string name;
  void* ptr;
  size_t size;
  int ncomp;
  int loc;
};
template <class T>
T* first_elem(T& x) { return &x; }
template <class T, size_t N>
T* first_elem(T (&x)[N]) { return &x[0]; }

struct AnyStage {
  virtual ~AnyStage() {}
  virtual glsl::StageBase& base() = 0;
  virtual void run() = 0;
  virtual bool set_uniform(const std::string& name, const UVal& v) = 0;
  virtual void bind_samplers(const Units& u) = 0;
  virtual void set_attribute(int loc, const AttrVal& a) = 0;
  virtual std::vector<Var> ins() = 0;
  virtual std::vector<Var> outs() = 0;
  virtual bool emits_triangle_strips() = 0;  // layout(triangle_strip, ...) out
  virtual int n_emitted() = 0;
  virtual void clear_emitted() = 0;
  virtual const std::vector<int>& prim_ends() = 0;
  virtual const glsl::vec4& emitted_position(int i) = 0;
  virtual const char* emitted_data(int i) = 0;
  virtual std::vector<Var> emitted_vars() = 0;  // ptr = offset inside one emitted record
};

typedef std::function<AnyStage*()> StageFactory;
inline std::map<std::string, StageFactory>& registry() {
  static std::map<std::string, StageFactory> r;
  return r;
}
struct Registrar {
  Registrar(const char* file, StageFactory f) { registry()[file] = f; }
};

// ------------------------------------------------------------------------------------------------ GL objects
enum : unsigned {
  GL_POINTS = 0x0000, GL_TRIANGLES = 0x0004, GL_TRIANGLE_STRIP = 0x0005,
  GL_DEPTH_BUFFER_BIT = 0x0100, GL_COLOR_BUFFER_BIT = 0x4000, GL_STENCIL_BUFFER_BIT = 0x0400,
  GL_NEVER = 0x0200, GL_LESS = 0x0201, GL_EQUAL = 0x0202, GL_LEQUAL = 0x0203, GL_GREATER = 0x0204, GL_NOTEQUAL = 0x0205,
  GL_GEQUAL = 0x0206, GL_ALWAYS = 0x0207,
  GL_ZERO = 0, GL_ONE = 1, GL_SRC_ALPHA = 0x0302, GL_ONE_MINUS_SRC_ALPHA = 0x0303,
  GL_DEPTH_TEST = 0x0B71, GL_BLEND = 0x0BE2, GL_CULL_FACE = 0x0B44, GL_RASTERIZER_DISCARD = 0x8C89,
  GL_PROGRAM_POINT_SIZE = 0x8642, GL_POINT_SMOOTH = 0x0B10, GL_LINE_SMOOTH = 0x0B20, GL_MULTISAMPLE = 0x809D,
  GL_VIEWPORT = 0x0BA2, GL_DEPTH_FUNC = 0x0B74, GL_COLOR_CLEAR_VALUE = 0x0C22, GL_POINT_SIZE = 0x0B11,
  GL_DRAW_FRAMEBUFFER_BINDING = 0x8CA6, GL_READ_FRAMEBUFFER_BINDING = 0x8CAA, GL_FRAMEBUFFER_BINDING = 0x8CA6,
  GL_FRAMEBUFFER = 0x8D40, GL_DRAW_FRAMEBUFFER = 0x8CA9, GL_READ_FRAMEBUFFER = 0x8CA8,
  GL_COLOR_ATTACHMENT0 = 0x8CE0, GL_COLOR_ATTACHMENT1 = 0x8CE1, GL_COLOR_ATTACHMENT2 = 0x8CE2, GL_COLOR_ATTACHMENT3 = 0x8CE3,
  GL_NONE = 0, GL_TEXTURE0 = 0x84C0, GL_FLOAT = 0x1406, GL_INT = 0x1404, GL_UNSIGNED_INT = 0x1405, GL_NO_ERROR = 0,
  GL_FRONT = 0x0404, GL_BACK = 0x0405, GL_FRONT_AND_BACK = 0x0408, GL_CW = 0x0900, GL_CCW = 0x0901,
};
enum { kMaxUnits = 16, kMaxAttribs = 16, kMaxColor = 4 };
const uint32_t kDepthClear = 0xFFFFFFu;

struct Buffer {
  std::vector<uint8_t> bytes;  // the bytes written so far (grown on demand); reads past them return 0
  size_t capacity = 0;         // size of the GL buffer object's data store
  void ensure(size_t n) { if (bytes.size() < n) bytes.resize(n, 0); }
};
struct Texture {
  int w = 0, h = 0, c = 4;      // c floats per texel (1 = R32F, 3 = RGB32F, 4 = RGBA32F)
  std::vector<float> data;      // [h][w][c]
  bool lin_min = false, lin_mag = false;
  std::shared_ptr<Buffer> tbo;  // texture buffer: texels are read from the buffer object
  void alloc(int W, int H, int C) { w = W; h = H; c = C; data.assign((size_t)W * H * C, 0.0f); }
};
struct SamplerObj {
  bool lin_min = false, lin_mag = false;
};
struct Renderbuffer {
  int w = 0, h = 0;
  std::vector<uint32_t> depth;
};
struct Framebuffer {
  int w = 0, h = 0;
  std::shared_ptr<Texture> color[kMaxColor];
  std::shared_ptr<Renderbuffer> depth;
  std::vector<int> draw_buffers{0};  // indices of colour attachments that fragment outputs 0,1,2.. go to (-1 = none)
};
struct VertexArray {
  struct Attr {
    std::shared_ptr<Buffer> buf;
    int size = 4;
    unsigned type = GL_FLOAT;
    bool integer = false;
    size_t stride = 0, offset = 0;
    bool enabled = false;
  } attr[kMaxAttribs];
};
struct TransformFeedback {
  std::vector<std::string> varyings;
  std::shared_ptr<Buffer> buffer;
  bool active = false;
  size_t written_bytes = 0;
  uint32_t primitives = 0;
};
struct Program {
  std::shared_ptr<AnyStage> vs, gs, fs;
  std::string vs_name, gs_name, fs_name;
  std::shared_ptr<TransformFeedback> tf;  // varyings captured at link time
  struct Link { void* dst; const void* src; size_t size; };
  std::vector<Link> vs_to_gs;
  bool linked = false;
  bool drawable = false;  // all stages transpiled
};

struct Context : Units {
  int viewport[4] = {0, 0, 0, 0};
  bool depth_test = false, blend = false, rasterizer_discard = false;
  unsigned depth_func = GL_LESS, blend_src = GL_ONE, blend_dst = GL_ZERO;
  float clear_color[4] = {0, 0, 0, 0};
  float point_size = 1.0f;
  std::shared_ptr<Framebuffer> fbo;  // null = default framebuffer (no storage: draws are dropped)
  std::shared_ptr<Program> program;
  std::shared_ptr<VertexArray> vao;
  std::shared_ptr<TransformFeedback> tf;
  int active_unit = 0;
  std::shared_ptr<Texture> unit_tex[kMaxUnits];
  std::shared_ptr<SamplerObj> unit_sampler[kMaxUnits];
  std::map<int, std::weak_ptr<Framebuffer>> fbo_ids;  // glGetIntegerv(binding) / glBindFramebuffer(id) round trip
  std::map<const Framebuffer*, int> fbo_id_of;
  int next_fbo_id = 1;
  uint64_t draw_calls = 0;
  bool binding_fragment_stage = false;
  // Attribute fetches past the last element written by GlBuffer::assign but inside the buffer object's data store return
  // what the store holds: with glow's assign (glBufferSubData into a store that only grows) that is the PREVIOUS, larger
  // upload's data. true = return zeros instead (what oracle/ and the CUDA path define for the label / probability
  // attributes whose pointer offset runs 4 / 5 elements past the end).
  bool zero_stale_tail = false;

  glsl::sampler2DRect rect(int unit) const override {
    glsl::sampler2DRect s;
    if (unit < 0 || unit >= kMaxUnits || !unit_tex[unit]) return s;
    const Texture& t = *unit_tex[unit];
    s.data = t.data.data(); s.W = t.w; s.H = t.h; s.C = t.c;
    // level of detail 0 on a texture without mipmaps selects the MAGNIFICATION filter (OpenGL 3.3 section 3.8.11:
    // lambda <= c with c = 0 unless the minification filter uses mipmaps)
    s.linear = unit_sampler[unit] ? unit_sampler[unit]->lin_mag : t.lin_mag;
    s.snap8 = binding_fragment_stage;  // see glsl_types.hpp
    return s;
  }
  glsl::samplerBuffer buffer(int unit) const override {
    glsl::samplerBuffer s;
    if (unit < 0 || unit >= kMaxUnits || !unit_tex[unit] || !unit_tex[unit]->tbo) return s;
    const Buffer& b = *unit_tex[unit]->tbo;
    s.data = reinterpret_cast<const float*>(b.bytes.data());
    s.n = (int)(b.bytes.size() / 16);
    return s;
  }
  int id_of(const std::shared_ptr<Framebuffer>& f) {
    if (!f) return 0;
    auto it = fbo_id_of.find(f.get());
    if (it != fbo_id_of.end()) return it->second;
    int id = next_fbo_id++;
    fbo_id_of[f.get()] = id;
    fbo_ids[id] = f;
    return id;
  }
};
inline Context& ctx() {
  static Context c;
  return c;
}

// ------------------------------------------------------------------------------------------------ rasteriser
inline uint32_t depth24(float zw) { return (uint32_t)llrint((double)zw * 16777215.0); }
inline bool depth_pass(unsigned func, uint32_t d, uint32_t cur) {
  switch (func) {
    case GL_LESS: return d < cur;
    case GL_LEQUAL: return d <= cur;
    case GL_ALWAYS: return true;
    case GL_EQUAL: return d == cur;
    case GL_GREATER: return d > cur;
    case GL_GEQUAL: return d >= cur;
    case GL_NOTEQUAL: return d != cur;
    default: return false;
  }
}
struct RVert { int64_t X, Y; float z; };
inline int64_t fdiv(int64_t a, int64_t b) { int64_t q = a / b, r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q; }
inline int64_t cdiv(int64_t a, int64_t b) { return -fdiv(-a, b); }
inline int64_t edge(const RVert& P, const RVert& Q, int64_t X, int64_t Y) { return (Q.X - P.X) * (Y - P.Y) - (Q.Y - P.Y) * (X - P.X); }
inline bool edge_inside(int64_t w, const RVert& P, const RVert& Q) {
  if (w != 0) return w > 0;
  int64_t dx = Q.X - P.X, dy = Q.Y - P.Y;
  return dy > 0 || (dy == 0 && dx > 0);
}

struct Pipeline {
  Context& c;
  Program& p;
  Framebuffer* fb;
  int W, H, vx, vy;
  std::vector<Var> fs_ins, fs_outs, last_vars;  // last_vars: emitted_vars of the GS, or outs of the VS
  struct FsLink { void* dst; size_t src_off; size_t size; int ncomp; };
  std::vector<FsLink> to_fs;  // from one emitted record (GS) / from VS out storage (no GS)
  struct TfField { size_t src_off; const void* src_abs; size_t size; };
  std::vector<TfField> tf_fields;
  TransformFeedback* tf = nullptr;

  Pipeline(Context& cx, Program& pr) : c(cx), p(pr) {
    fb = c.fbo.get();
    vx = c.viewport[0]; vy = c.viewport[1]; W = c.viewport[2]; H = c.viewport[3];
    if (p.fs) { fs_ins = p.fs->ins(); fs_outs = p.fs->outs(); }
    last_vars = p.gs ? p.gs->emitted_vars() : p.vs->outs();
    for (const Var& in : fs_ins)
      for (const Var& o : last_vars)
        if (o.name == in.name && o.size == in.size) {
          to_fs.push_back(FsLink{in.ptr, (size_t)o.ptr, in.size, in.ncomp});
          break;
        }
    if (c.tf && c.tf->active) {
      tf = c.tf.get();
      for (const std::string& name : tf->varyings) {
        bool found = false;
        for (const Var& o : last_vars)
          if (o.name == name) {
            tf_fields.push_back(TfField{(size_t)o.ptr, o.ptr, o.size});
            found = true;
            break;
          }
        if (!found) throw std::runtime_error("transform feedback varying not found: " + name);
      }
    }
  }

  // base pointer of the last vertex-processing stage's outputs for emitted vertex i (GS) or the current vertex (VS)
  const char* record(int i) const { return p.gs ? p.gs->emitted_data(i) : nullptr; }
  const void* field(const char* rec, size_t off_or_ptr) const { return rec ? (const void*)(rec + off_or_ptr) : (const void*)off_or_ptr; }

  void capture(const char* rec) {  // one point primitive
    if (!tf || !tf->buffer) return;
    size_t need = 0;
    for (const TfField& f : tf_fields) need += f.size;
    if (tf->written_bytes + need > tf->buffer->capacity) return;  // full: the primitive is dropped (not counted)
    tf->buffer->ensure(tf->written_bytes + need);
    uint8_t* dst = tf->buffer->bytes.data() + tf->written_bytes;
    for (const TfField& f : tf_fields) {
      memcpy(dst, field(rec, f.src_off), f.size);
      dst += f.size;
    }
    tf->written_bytes += need;
    tf->primitives += 1;
  }

  void write_fragment(int px, int py, float zw) {  // fragment shader inputs are set; runs it and writes its outputs
    glsl::StageBase& b = p.fs->base();
    b.discarded_ = false;
    p.fs->run();
    if (b.discarded_) return;
    if (px < 0 || py < 0 || px >= fb->w || py >= fb->h) return;
    size_t pix = (size_t)py * fb->w + px;
    if (c.depth_test && fb->depth) {
      if (!(zw >= 0.0f && zw <= 1.0f)) return;
      uint32_t d = depth24(zw);
      if (!depth_pass(c.depth_func, d, fb->depth->depth[pix])) return;
      fb->depth->depth[pix] = d;
    }
    for (const Var& o : fs_outs) {
      int loc = o.loc < 0 ? 0 : o.loc;
      if (loc >= (int)fb->draw_buffers.size()) continue;
      int att = fb->draw_buffers[loc];
      if (att < 0 || att >= kMaxColor || !fb->color[att]) continue;
      Texture& t = *fb->color[att];
      const float* src = (const float*)o.ptr;
      int n = o.ncomp > 0 ? o.ncomp : 1;
      float v[4] = {0, 0, 0, 1};
      for (int k = 0; k < n && k < 4; ++k) v[k] = src[k];
      float* dst = t.data.data() + pix * t.c;
      for (int k = 0; k < t.c; ++k) {
        if (c.blend) {
          float s = c.blend_src == GL_ONE ? v[k] : 0.0f, d = c.blend_dst == GL_ONE ? dst[k] : 0.0f;
          dst[k] = d + s;  // GL_FUNC_ADD; destination first, as in ../ref_pipeline.cpp (blend[py][px][c] + result[c])
        } else {
          dst[k] = v[k];
        }
      }
    }
  }

  bool to_pixel(const glsl::vec4& pos, int* px, int* py, float* zw) const {
    if (!(pos.x >= -pos.w && pos.x <= pos.w && pos.y >= -pos.w && pos.y <= pos.w && pos.z >= -pos.w && pos.z <= pos.w))
      return false;
    float xw = (0.5f * pos.x + 0.5f) * (float)W, yw = (0.5f * pos.y + 0.5f) * (float)H;  // w = 1 on this path
    int ix = (int)std::floor(xw), iy = (int)std::floor(yw);
    if (ix < 0 || iy < 0 || ix >= W || iy >= H) return false;
    *px = vx + ix; *py = vy + iy; *zw = 0.5f * pos.z + 0.5f;
    return true;
  }

  void point(const glsl::vec4& pos, const char* rec) {
    if (c.rasterizer_discard || !fb || !p.fs) return;
    int px, py; float zw;
    if (!to_pixel(pos, &px, &py, &zw)) return;
    for (const FsLink& l : to_fs) memcpy(l.dst, field(rec, l.src_off), l.size);
    write_fragment(px, py, zw);
  }

  void triangle(int i0, int i1, int i2) {  // emitted vertices of the geometry stage
    if (c.rasterizer_discard || !fb || !p.fs) return;
    const int idx[3] = {i0, i1, i2};
    RVert v[3];
    const char* rec[3];
    for (int k = 0; k < 3; ++k) {
      const glsl::vec4& pos = p.gs->emitted_position(idx[k]);
      float xw = (0.5f * pos.x + 0.5f) * (float)W, yw = (0.5f * pos.y + 0.5f) * (float)H;
      v[k].X = (int64_t)llrintf(xw * 256.0f); v[k].Y = (int64_t)llrintf(yw * 256.0f);
      v[k].z = 0.5f * pos.z + 0.5f;
      rec[k] = p.gs->emitted_data(idx[k]);
    }
    RVert A = v[0], B = v[1], C = v[2];
    const char *rA = rec[0], *rB = rec[1], *rC = rec[2];
    int64_t area = (B.X - A.X) * (C.Y - A.Y) - (B.Y - A.Y) * (C.X - A.X);
    if (area == 0) return;
    if (area < 0) { std::swap(B, C); std::swap(rB, rC); area = -area; }
    int64_t minX = std::min(A.X, std::min(B.X, C.X)), maxX = std::max(A.X, std::max(B.X, C.X));
    int64_t minY = std::min(A.Y, std::min(B.Y, C.Y)), maxY = std::max(A.Y, std::max(B.Y, C.Y));
    int64_t x0 = std::max<int64_t>(cdiv(minX - 128, 256), 0), x1 = std::min<int64_t>(fdiv(maxX - 128, 256), W - 1);
    int64_t y0 = std::max<int64_t>(cdiv(minY - 128, 256), 0), y1 = std::min<int64_t>(fdiv(maxY - 128, 256), H - 1);
    float farea = (float)area;
    for (int64_t j = y0; j <= y1; ++j)
      for (int64_t i = x0; i <= x1; ++i) {
        int64_t X = i * 256 + 128, Y = j * 256 + 128;
        int64_t wA = edge(B, C, X, Y), wB = edge(C, A, X, Y), wC = edge(A, B, X, Y);
        if (!edge_inside(wA, B, C) || !edge_inside(wB, C, A) || !edge_inside(wC, A, B)) continue;
        float fB = (float)wB / farea, fC = (float)wC / farea, fA = (1.0f - fB) - fC;
        float z = (fA * A.z + fB * B.z) + fC * C.z;
        for (const FsLink& l : to_fs) {
          const float* a = (const float*)(rA + l.src_off);
          const float* b = (const float*)(rB + l.src_off);
          const float* cc = (const float*)(rC + l.src_off);
          if (l.ncomp <= 0) { memcpy(l.dst, rA + l.src_off, l.size); continue; }
          float* d = (float*)l.dst;
          for (int k = 0; k < l.ncomp; ++k)
            d[k] = (a[k] == b[k] && b[k] == cc[k]) ? a[k] : (fA * a[k] + fB * b[k]) + fC * cc[k];
        }
        write_fragment(vx + (int)i, vy + (int)j, z);
      }
  }
};

inline AttrVal fetch_attr(const VertexArray::Attr& a, int vertex) {
  AttrVal v;
  if (!a.buf) return v;
  size_t stride = a.stride ? a.stride : (size_t)a.size * 4;
  size_t off = a.offset + stride * (size_t)vertex;
  v.integer = a.integer;
  for (int k = 0; k < a.size && k < 4; ++k) {
    // reads past the end of the buffer object return 0 (robust buffer access; cf. ../ref_pipeline.cpp on the label /
    // probability attributes whose pointer offset runs past the last elements)
    uint32_t word = 0;
    if (off + 4 * (size_t)k + 4 <= a.buf->bytes.size()) memcpy(&word, a.buf->bytes.data() + off + 4 * (size_t)k, 4);
    if (a.integer) {
      v.i[k] = (int)word;
    } else if (a.type == GL_FLOAT) {
      memcpy(&v.f[k], &word, 4);
    } else {
      v.f[k] = (a.type == GL_INT) ? (float)(int)word : (float)word;
    }
  }
  return v;
}

inline void draw_arrays(unsigned mode, int first, int count) {
  Context& c = ctx();
  c.draw_calls += 1;
  if (!c.program || count <= 0) return;
  Program& p = *c.program;
  if (!p.drawable) throw std::runtime_error("draw with a program whose shaders are not part of the transpiled set: " +
                                            p.vs_name + " " + p.gs_name + " " + p.fs_name);
  if (mode != GL_POINTS) throw std::runtime_error("software GL: only GL_POINTS input primitives are used on this path");
  if (!c.vao) throw std::runtime_error("draw without a vertex array");
  Pipeline pl(c, p);
  p.vs->bind_samplers(c);
  if (p.gs) p.gs->bind_samplers(c);
  if (p.fs) {
    c.binding_fragment_stage = true;
    p.fs->bind_samplers(c);
    c.binding_fragment_stage = false;
  }
  const std::vector<Var> vs_outs = p.vs->outs();
  for (int k = 0; k < count; ++k) {
    const int vid = first + k;
    for (int a = 0; a < kMaxAttribs; ++a)
      if (c.vao->attr[a].enabled) p.vs->set_attribute(a, fetch_attr(c.vao->attr[a], vid));
    p.vs->base().gl_VertexID = vid;
    p.vs->run();
    if (!p.gs) {
      pl.capture(nullptr);
      pl.point(p.vs->base().gl_Position, nullptr);
      continue;
    }
    for (const Program::Link& l : p.vs_to_gs) memcpy(l.dst, l.src, l.size);
    p.gs->base().gl_in[0].gl_Position = p.vs->base().gl_Position;
    p.gs->clear_emitted();
    p.gs->run();
    const int n = p.gs->n_emitted();
    if (n == 0) continue;
    const bool strips = p.gs->emits_triangle_strips();
    const std::vector<int>& ends = p.gs->prim_ends();
    int begin = 0;
    for (size_t e = 0; e <= ends.size(); ++e) {
      int end = e < ends.size() ? ends[e] : n;
      if (strips) {
        // triangle t of a strip = vertices (t, t+1, t+2). GL swaps the first two for odd t to keep the winding; with
        // face culling off that only renames the vertices -- the barycentric rule of DESIGN.md section 2 is stated for
        // (t, t+1, t+2), as ../ref_pipeline.cpp, oracle/ and the CUDA path evaluate it
        for (int t = begin; t + 2 < end; ++t) pl.triangle(t, t + 1, t + 2);
      } else {
        for (int t = begin; t < end; ++t) {
          pl.capture(p.gs->emitted_data(t));
          pl.point(p.gs->emitted_position(t), p.gs->emitted_data(t));
        }
      }
      begin = end;
      if (e == ends.size()) break;
    }
  }
}

inline void clear(unsigned mask) {
  Context& c = ctx();
  Framebuffer* fb = c.fbo.get();
  if (!fb) return;
  if (mask & GL_COLOR_BUFFER_BIT)
    for (size_t k = 0; k < fb->draw_buffers.size(); ++k) {
      int att = fb->draw_buffers[k];
      if (att < 0 || att >= kMaxColor || !fb->color[att]) continue;
      Texture& t = *fb->color[att];
      for (size_t i = 0; i < t.data.size(); ++i) t.data[i] = c.clear_color[(i % t.c) < 4 ? (i % t.c) : 3];
    }
  if ((mask & GL_DEPTH_BUFFER_BIT) && fb->depth) fb->depth->depth.assign((size_t)fb->depth->w * fb->depth->h, kDepthClear);
}

// link: resolve the stage objects and the VS -> GS interface (by block / variable name)
inline void link(Program& p) {
  auto make = [](const std::string& name) -> std::shared_ptr<AnyStage> {
    auto it = registry().find(name);
    if (it == registry().end()) return nullptr;
    return std::shared_ptr<AnyStage>(it->second());
  };
  p.vs = p.vs_name.empty() ? nullptr : make(p.vs_name);
  p.gs = p.gs_name.empty() ? nullptr : make(p.gs_name);
  p.fs = p.fs_name.empty() ? nullptr : make(p.fs_name);
  p.drawable = p.vs && (p.gs_name.empty() || p.gs) && (p.fs_name.empty() || p.fs);
  p.vs_to_gs.clear();
  if (p.drawable && p.gs) {
    std::vector<Var> o = p.vs->outs(), i = p.gs->ins();
    for (const Var& in : i)
      for (const Var& out : o)
        if (in.name == out.name && in.size == out.size) {
          p.vs_to_gs.push_back(Program::Link{in.ptr, out.ptr, in.size});
          break;
        }
  }
  p.linked = true;
}

inline void set_uniform(Program& p, const std::string& name, const UVal& v) {
  if (p.vs) p.vs->set_uniform(name, v);
  if (p.gs) p.gs->set_uniform(name, v);
  if (p.fs) p.fs->set_uniform(name, v);  // unknown names are ignored, as glGetUniformLocation == -1 is
}

}  // namespace sgl
