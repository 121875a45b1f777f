// glow_all.hpp -- TEST INFRASTRUCTURE (oracle/_ref/libsuma_ref_full.so). Stand-in for `glow` (github.com/jbehley/glow @
// e66d7f85, the C++ OpenGL wrapper the reference is written against; NOT vendored in the reference and not installed
// here): the classes, enumerators and GL entry points that core/Frame.h, Preprocessing.cpp, Frame2Model.cpp,
// SurfelMap.cpp and SurfelMapping.cpp use, with glow's public interface (names, argument order, value semantics of the
// handles = shared GL object) on top of the software GL of ../sgl.hpp. Written from glow's documented interface and the
// reference's call sites; contains no glow source.
#pragma once
#include <cassert>
#include <cstdint>
#include <iostream>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#include <eigen3/Eigen/Dense>

#include "../sgl.hpp"

// ---- GL types and entry points (global namespace, as <GL/glew.h> provides them) ----
typedef unsigned int GLenum;
typedef unsigned int GLuint;
typedef int GLint;
typedef int GLsizei;
typedef float GLfloat;
typedef unsigned char GLboolean;
typedef void GLvoid;
typedef unsigned int GLbitfield;
using sgl::GL_POINTS; using sgl::GL_TRIANGLES; using sgl::GL_TRIANGLE_STRIP;
using sgl::GL_DEPTH_BUFFER_BIT; using sgl::GL_COLOR_BUFFER_BIT; using sgl::GL_STENCIL_BUFFER_BIT;
using sgl::GL_LESS; using sgl::GL_LEQUAL; using sgl::GL_ALWAYS; using sgl::GL_EQUAL; using sgl::GL_GREATER; using sgl::GL_GEQUAL;
using sgl::GL_ONE; using sgl::GL_ZERO; using sgl::GL_SRC_ALPHA; using sgl::GL_ONE_MINUS_SRC_ALPHA;
using sgl::GL_DEPTH_TEST; using sgl::GL_BLEND; using sgl::GL_CULL_FACE; using sgl::GL_RASTERIZER_DISCARD;
using sgl::GL_PROGRAM_POINT_SIZE; using sgl::GL_POINT_SMOOTH; using sgl::GL_LINE_SMOOTH; using sgl::GL_MULTISAMPLE;
using sgl::GL_VIEWPORT; using sgl::GL_DEPTH_FUNC; using sgl::GL_COLOR_CLEAR_VALUE; using sgl::GL_POINT_SIZE;
using sgl::GL_DRAW_FRAMEBUFFER_BINDING; using sgl::GL_READ_FRAMEBUFFER_BINDING;
using sgl::GL_FRAMEBUFFER; using sgl::GL_DRAW_FRAMEBUFFER; using sgl::GL_READ_FRAMEBUFFER;
using sgl::GL_COLOR_ATTACHMENT0; using sgl::GL_COLOR_ATTACHMENT1; using sgl::GL_COLOR_ATTACHMENT2; using sgl::GL_COLOR_ATTACHMENT3;
using sgl::GL_NONE; using sgl::GL_TEXTURE0; using sgl::GL_FLOAT; using sgl::GL_INT; using sgl::GL_UNSIGNED_INT;
using sgl::GL_FRONT; using sgl::GL_BACK; using sgl::GL_FRONT_AND_BACK; using sgl::GL_CW; using sgl::GL_CCW;
enum : unsigned { GL_TEXTURE1 = GL_TEXTURE0 + 1, GL_TEXTURE2, GL_TEXTURE3, GL_TEXTURE4, GL_TEXTURE5, GL_TEXTURE6, GL_TEXTURE7,
                  GL_TEXTURE8, GL_TEXTURE9, GL_TEXTURE10, GL_TEXTURE11, GL_TEXTURE12 };
#define GL_TRUE 1
#define GL_FALSE 0

inline void glEnable(GLenum cap) {
  sgl::Context& c = sgl::ctx();
  if (cap == GL_DEPTH_TEST) c.depth_test = true;
  else if (cap == GL_BLEND) c.blend = true;
  else if (cap == GL_RASTERIZER_DISCARD) c.rasterizer_discard = true;
}
inline void glDisable(GLenum cap) {
  sgl::Context& c = sgl::ctx();
  if (cap == GL_DEPTH_TEST) c.depth_test = false;
  else if (cap == GL_BLEND) c.blend = false;
  else if (cap == GL_RASTERIZER_DISCARD) c.rasterizer_discard = false;
}
inline void glDepthFunc(GLenum f) { sgl::ctx().depth_func = f; }
inline void glBlendFunc(GLenum s, GLenum d) { sgl::ctx().blend_src = s; sgl::ctx().blend_dst = d; }
inline void glClearColor(float r, float g, float b, float a) {
  float* c = sgl::ctx().clear_color;
  c[0] = r; c[1] = g; c[2] = b; c[3] = a;
}
inline void glViewport(GLint x, GLint y, GLsizei w, GLsizei h) {
  int* v = sgl::ctx().viewport;
  v[0] = x; v[1] = y; v[2] = w; v[3] = h;
}
inline void glPointSize(float s) { sgl::ctx().point_size = s; }
inline void glClear(GLbitfield mask) { sgl::clear(mask); }
inline void glFinish() {}
inline void glFlush() {}
inline void glActiveTexture(GLenum unit) { sgl::ctx().active_unit = (int)(unit - GL_TEXTURE0); }
inline void glDrawArrays(GLenum mode, GLint first, GLsizei count) { sgl::draw_arrays(mode, first, count); }
inline GLenum glGetError() { return 0; }
inline void glGetIntegerv(GLenum what, GLint* out) {
  sgl::Context& c = sgl::ctx();
  if (what == GL_VIEWPORT) for (int i = 0; i < 4; ++i) out[i] = c.viewport[i];
  else if (what == GL_DEPTH_FUNC) out[0] = (GLint)c.depth_func;
  else if (what == GL_DRAW_FRAMEBUFFER_BINDING || what == GL_READ_FRAMEBUFFER_BINDING) out[0] = c.id_of(c.fbo);
  else out[0] = 0;
}
inline void glGetFloatv(GLenum what, GLfloat* out) {
  sgl::Context& c = sgl::ctx();
  if (what == GL_COLOR_CLEAR_VALUE) for (int i = 0; i < 4; ++i) out[i] = c.clear_color[i];
  else if (what == GL_POINT_SIZE) out[0] = c.point_size;
  else out[0] = 0.0f;
}
inline void glGetBooleanv(GLenum what, GLboolean* out) {
  sgl::Context& c = sgl::ctx();
  out[0] = (what == GL_DEPTH_TEST) ? c.depth_test : (what == GL_BLEND) ? c.blend : false;
}
inline void glBindFramebuffer(GLenum, GLuint id) {
  sgl::Context& c = sgl::ctx();
  if (id == 0) { c.fbo.reset(); return; }
  auto it = c.fbo_ids.find((int)id);
  c.fbo = it == c.fbo_ids.end() ? nullptr : it->second.lock();
}
inline void glDrawBuffers(GLsizei n, const GLenum* bufs) {  // state of the bound framebuffer object
  sgl::Context& c = sgl::ctx();
  if (!c.fbo) return;
  c.fbo->draw_buffers.clear();
  for (int i = 0; i < n; ++i) c.fbo->draw_buffers.push_back(bufs[i] == GL_NONE ? -1 : (int)(bufs[i] - GL_COLOR_ATTACHMENT0));
}

namespace glow {

inline void _CheckGlError(const char*, int) {}
#define CheckGlError() glow::_CheckGlError(__FILE__, __LINE__)

// glow/glutil.h: small POD vectors + angle helpers
struct vec2 { float x, y; vec2() : x(0), y(0) {} vec2(float a, float b) : x(a), y(b) {} };
struct vec3 { float x, y, z; vec3() : x(0), y(0), z(0) {} vec3(float a, float b, float c) : x(a), y(b), z(c) {} };
struct vec4 { float x, y, z, w; vec4() : x(0), y(0), z(0), w(0) {} vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {} };
// un-vendored; taken as deg * (float)(pi / 180) like oracle/ and ../ref_pipeline.cpp
// (the capitalised Radians() the reference also calls is rv::Radians, vendored in rv/geometry.h:81 -- same constant)
inline float radians(float deg) { return deg * (float)(3.14159265358979323846 / 180.0); }
inline float degrees(float rad) { return rad * (float)(180.0 / 3.14159265358979323846); }

enum class BufferTarget { ARRAY_BUFFER, ELEMENT_ARRAY_BUFFER, TEXTURE_BUFFER, TRANSFORM_FEEDBACK_BUFFER };
enum class BufferUsage { STREAM_DRAW, STREAM_READ, STREAM_COPY, STATIC_DRAW, STATIC_READ, STATIC_COPY, DYNAMIC_DRAW, DYNAMIC_READ, DYNAMIC_COPY };
enum class TextureFormat { RGBA, RGB, RG, R, R_INTEGER, RG_INTEGER, RGB_INTEGER, RGBA_INTEGER, R_FLOAT, RG_FLOAT, RGB_FLOAT, RGBA_FLOAT, DEPTH, DEPTH_STENCIL };
enum class PixelFormat { R, RG, RGB, RGBA, R_INTEGER, RG_INTEGER, RGB_INTEGER, RGBA_INTEGER, DEPTH, DEPTH_STENCIL };
enum class PixelType { UNSIGNED_BYTE, BYTE, UNSIGNED_SHORT, SHORT, UNSIGNED_INT, INT, HALF_FLOAT, FLOAT };
enum class TexMinOp { LINEAR, NEAREST, NEAREST_MIPMAP_NEAREST, NEAREST_MIPMAP_LINEAR, LINEAR_MIPMAP_NEAREST, LINEAR_MIPMAP_LINEAR };
enum class TexMagOp { LINEAR, NEAREST };
enum class TexWrapOp { CLAMP_TO_EDGE, CLAMP_TO_BORDER, MIRRORED_REPEAT, REPEAT, MIRROR_CLAMP_TO_EDGE };
enum class TexRectMinOp { LINEAR, NEAREST };
enum class TexRectMagOp { LINEAR, NEAREST };
enum class TexRectWrapOp { CLAMP_TO_EDGE, CLAMP_TO_BORDER };
enum class FramebufferTarget { BOTH, READ, DRAW };
enum class FramebufferAttachment { COLOR0, COLOR1, COLOR2, COLOR3, COLOR4, COLOR5, COLOR6, COLOR7, DEPTH, STENCIL, DEPTH_STENCIL };
enum class RenderbufferFormat { RGBA, RGB, RG, R, DEPTH, DEPTH_STENCIL };
enum class ShaderType { VERTEX_SHADER, GEOMETRY_SHADER, FRAGMENT_SHADER, TESS_CONTROL_SHADER, TESS_EVALUTION_SHADER, COMPUTE_SHADER };
enum class AttributeType { BYTE, UNSIGNED_BYTE, SHORT, UNSIGNED_SHORT, INT, UNSIGNED_INT, HALF_FLOAT, FLOAT, DOUBLE };
enum class TransformFeedbackMode { POINTS, LINES, TRIANGLES };
enum class QueryTarget { SAMPLES_PASSED, ANY_SAMPLES_PASSED, PRIMITIVES_GENERATED, TRANSFORM_FEEDBACK_PRIMITIVES_WRITTEN, TIME_ELAPSED };

inline int format_channels(TextureFormat f) {
  switch (f) {
    case TextureFormat::R: case TextureFormat::R_FLOAT: case TextureFormat::R_INTEGER: case TextureFormat::DEPTH: return 1;
    case TextureFormat::RGB: case TextureFormat::RGB_FLOAT: case TextureFormat::RGB_INTEGER: return 3;
    case TextureFormat::RG: case TextureFormat::RG_FLOAT: case TextureFormat::RG_INTEGER: return 2;
    default: return 4;
  }
}
inline int pixel_channels(PixelFormat f) {
  switch (f) {
    case PixelFormat::R: case PixelFormat::R_INTEGER: case PixelFormat::DEPTH: return 1;
    case PixelFormat::RG: case PixelFormat::RG_INTEGER: return 2;
    case PixelFormat::RGB: case PixelFormat::RGB_INTEGER: return 3;
    default: return 4;
  }
}

// how an element of GlBuffer<T> lies in the buffer object: T's bytes, except for the stand-in Eigen matrices, whose GL
// image is their column-major coefficients (what the real Eigen::Matrix4f's bytes are)
template <class T, class Enable = void>
struct Pod {
  static const size_t size = sizeof(T);
  static void store(uint8_t* dst, const T& v) { memcpy(dst, &v, sizeof(T)); }
  static void load(T& v, const uint8_t* src) { memcpy((void*)&v, src, sizeof(T)); }
};
template <class S, int R, int C>
struct Pod<Eigen::Matrix<S, R, C>, void> {
  static const size_t size = sizeof(S) * R * C;
  static void store(uint8_t* dst, const Eigen::Matrix<S, R, C>& v) { memcpy(dst, v.data(), size); }
  static void load(Eigen::Matrix<S, R, C>& v, const uint8_t* src) { memcpy(v.data(), src, size); }
};

template <class T>
class GlBuffer {
 public:
  GlBuffer(BufferTarget target = BufferTarget::ARRAY_BUFFER, BufferUsage usage = BufferUsage::DYNAMIC_DRAW)
      : st_(std::make_shared<State>()), target_(target), usage_(usage) {
    st_->buf = std::make_shared<sgl::Buffer>();
  }
  void assign(const std::vector<T>& data) {
    reserve(data.size());
    st_->buf->ensure(data.size() * Pod<T>::size);
    for (size_t i = 0; i < data.size(); ++i) Pod<T>::store(st_->buf->bytes.data() + i * Pod<T>::size, data[i]);
    st_->size = data.size();
    zero_tail();
  }
  void assign(const GlBuffer<T>& other) {
    if (other.st_ == st_) return;
    const size_t n = other.st_->size * Pod<T>::size;
    reserve(other.st_->size);
    st_->buf->ensure(n);
    other.st_->buf->ensure(n);
    memcpy(st_->buf->bytes.data(), other.st_->buf->bytes.data(), n);
    st_->size = other.st_->size;
    zero_tail();
  }
  void insert(uint32_t offset, const T& value) {
    if (offset >= st_->capacity) return;
    st_->buf->ensure(((size_t)offset + 1) * Pod<T>::size);
    Pod<T>::store(st_->buf->bytes.data() + (size_t)offset * Pod<T>::size, value);
    if (offset >= st_->size) st_->size = offset + 1;
  }
  void replace(uint32_t offset, const std::vector<T>& data) {
    size_t n = data.size();
    if (offset + n > st_->capacity) n = offset < st_->capacity ? st_->capacity - offset : 0;
    st_->buf->ensure((offset + n) * Pod<T>::size);
    for (size_t i = 0; i < n; ++i) Pod<T>::store(st_->buf->bytes.data() + (offset + i) * Pod<T>::size, data[i]);
  }
  void get(std::vector<T>& data) const { get(data, 0, (uint32_t)st_->size); }
  void get(std::vector<T>& data, uint32_t start, uint32_t count) const {
    data.clear();
    st_->buf->ensure(std::min<size_t>((size_t)start + count, st_->capacity) * Pod<T>::size);
    for (uint32_t i = start; i < start + count && i < st_->capacity; ++i) {
      T v;
      Pod<T>::load(v, st_->buf->bytes.data() + (size_t)i * Pod<T>::size);
      data.push_back(v);
    }
  }
  void reserve(uint32_t n) {  // grows the data store, keeps the content
    if (n <= st_->capacity) return;
    st_->capacity = n;
    st_->buf->capacity = (size_t)n * Pod<T>::size;
  }
  void resize(uint32_t n) {
    if (n > st_->capacity) reserve(n);
    st_->size = n;
  }
  size_t size() const { return st_->size; }
  size_t capacity() const { return st_->capacity; }
  size_t memorySize() const { return st_->capacity * Pod<T>::size; }
  void bind() {}
  void release() {}
  BufferTarget target() const { return target_; }
  const std::shared_ptr<sgl::Buffer>& object() const { return st_->buf; }

 private:
  void zero_tail() {  // see sgl::Context::zero_stale_tail
    const size_t used = st_->size * Pod<T>::size;
    if (sgl::ctx().zero_stale_tail && st_->buf->bytes.size() > used)
      memset(st_->buf->bytes.data() + used, 0, st_->buf->bytes.size() - used);
  }
  struct State {
    std::shared_ptr<sgl::Buffer> buf;
    size_t size = 0, capacity = 0;
  };
  std::shared_ptr<State> st_;  // copies of a GlBuffer are handles of the same buffer object
  BufferTarget target_;
  BufferUsage usage_;
};

// common part of GlTexture / GlTextureRectangle / GlTextureBuffer: binds to the active texture unit
class GlTextureBase {
 public:
  void bind() { sgl::ctx().unit_tex[sgl::ctx().active_unit] = tex_; }
  void release() { sgl::ctx().unit_tex[sgl::ctx().active_unit].reset(); }
  uint32_t width() const { return (uint32_t)tex_->w; }
  uint32_t height() const { return (uint32_t)tex_->h; }
  const std::shared_ptr<sgl::Texture>& object() const { return tex_; }
  // texel data in the requested channel layout, rows bottom to top (GL order = our storage order)
  void download(PixelFormat fmt, float* dst) const {
    const int n = pixel_channels(fmt);
    const sgl::Texture& t = *tex_;
    for (size_t p = 0; p < (size_t)t.w * t.h; ++p)
      for (int k = 0; k < n; ++k) dst[p * n + k] = k < t.c ? t.data[p * t.c + k] : (k == 3 ? 1.0f : 0.0f);
  }
  void download(std::vector<vec4>& data) const {
    const sgl::Texture& t = *tex_;
    data.resize((size_t)t.w * t.h);
    for (size_t p = 0; p < data.size(); ++p) {
      float v[4] = {0, 0, 0, 1};
      for (int k = 0; k < t.c && k < 4; ++k) v[k] = t.data[p * t.c + k];
      data[p] = vec4(v[0], v[1], v[2], v[3]);
    }
  }
  void download(std::vector<float>& data) const { data = tex_->data; }
  // upload: FLOAT data in the given channel layout; other pixel types (colour tables of the visualiser) are not on the path
  template <class P>
  void assign(PixelFormat fmt, PixelType type, const P* src) {
    if (type != PixelType::FLOAT || !src) return;
    const int n = pixel_channels(fmt);
    sgl::Texture& t = *tex_;
    const float* f = reinterpret_cast<const float*>(src);
    for (size_t px = 0; px < (size_t)t.w * t.h; ++px)
      for (int k = 0; k < t.c; ++k) t.data[px * t.c + k] = k < n ? f[px * n + k] : (k == 3 ? 1.0f : 0.0f);
  }

 protected:
  std::shared_ptr<sgl::Texture> tex_ = std::make_shared<sgl::Texture>();
};

class GlTextureRectangle : public GlTextureBase {
 public:
  GlTextureRectangle(uint32_t w, uint32_t h, TextureFormat f = TextureFormat::RGBA) { tex_->alloc((int)w, (int)h, format_channels(f)); }
  void setMinifyingOperation(TexRectMinOp op) { tex_->lin_min = op == TexRectMinOp::LINEAR; }
  void setMagnifyingOperation(TexRectMagOp op) { tex_->lin_mag = op == TexRectMagOp::LINEAR; }
  void setWrapOperation(TexRectWrapOp, TexRectWrapOp) {}  // border colour (0,0,0,0) either way; coordinates are unnormalised
  void copy(const GlTextureRectangle& other) {
    if (other.tex_->w == tex_->w && other.tex_->h == tex_->h && other.tex_->c == tex_->c) tex_->data = other.tex_->data;
    else throw std::runtime_error("GlTextureRectangle::copy: size / format mismatch");
  }
  void resize(uint32_t w, uint32_t h) { tex_->alloc((int)w, (int)h, tex_->c); }
  GlTextureRectangle clone() const {
    GlTextureRectangle t((uint32_t)tex_->w, (uint32_t)tex_->h);
    *t.tex_ = *tex_;
    return t;
  }
};

class GlTexture : public GlTextureBase {
 public:
  GlTexture(uint32_t w, TextureFormat f = TextureFormat::RGBA) { tex_->alloc((int)w, 1, format_channels(f)); }
  GlTexture(uint32_t w, uint32_t h, TextureFormat f = TextureFormat::RGBA) { tex_->alloc((int)w, (int)h, format_channels(f)); }
  GlTexture(uint32_t w, uint32_t h, uint32_t, TextureFormat f) { tex_->alloc((int)w, (int)h, format_channels(f)); }
  void setMinifyingOperation(TexMinOp op) { tex_->lin_min = op != TexMinOp::NEAREST; }
  void setMagnifyingOperation(TexMagOp op) { tex_->lin_mag = op == TexMagOp::LINEAR; }
  void setWrapOperation(TexWrapOp) {}
  void setWrapOperation(TexWrapOp, TexWrapOp) {}
  void setWrapOperation(TexWrapOp, TexWrapOp, TexWrapOp) {}
  void copy(const GlTexture& other) { tex_->data = other.tex_->data; }
  void resize(uint32_t w, uint32_t h) { tex_->alloc((int)w, (int)h, tex_->c); }
  void generateMipmaps() {}
};

class GlTextureBuffer : public GlTextureBase {
 public:
  template <class T>
  GlTextureBuffer(GlBuffer<T>& buffer, TextureFormat) { tex_->tbo = buffer.object(); }
};

class GlSampler {
 public:
  void setMinifyingOperation(TexMinOp op) { s_->lin_min = op != TexMinOp::NEAREST; }
  void setMagnifyingOperation(TexMagOp op) { s_->lin_mag = op == TexMagOp::LINEAR; }
  void setWrapOperation(TexWrapOp) {}
  void setWrapOperation(TexWrapOp, TexWrapOp) {}
  void setWrapOperation(TexWrapOp, TexWrapOp, TexWrapOp) {}
  void bind(uint32_t unit) { sgl::ctx().unit_sampler[unit] = s_; }
  void release(uint32_t unit) { sgl::ctx().unit_sampler[unit].reset(); }

 private:
  std::shared_ptr<sgl::SamplerObj> s_ = std::make_shared<sgl::SamplerObj>();
};

class GlRenderbuffer {
 public:
  GlRenderbuffer(uint32_t w, uint32_t h, RenderbufferFormat f = RenderbufferFormat::RGBA) : format_(f) {
    rb_->w = (int)w; rb_->h = (int)h;
    rb_->depth.assign((size_t)w * h, sgl::kDepthClear);
  }
  uint32_t width() const { return (uint32_t)rb_->w; }
  uint32_t height() const { return (uint32_t)rb_->h; }
  const std::shared_ptr<sgl::Renderbuffer>& object() const { return rb_; }
  RenderbufferFormat format() const { return format_; }

 private:
  std::shared_ptr<sgl::Renderbuffer> rb_ = std::make_shared<sgl::Renderbuffer>();
  RenderbufferFormat format_;
};

class GlFramebuffer {
 public:
  GlFramebuffer(uint32_t w, uint32_t h, FramebufferTarget = FramebufferTarget::BOTH) { fb_->w = (int)w; fb_->h = (int)h; }
  void attach(FramebufferAttachment a, GlTextureBase& t) { attach_tex(a, t.object()); }
  void attach(FramebufferAttachment a, GlRenderbuffer& rb) {
    if (a == FramebufferAttachment::DEPTH || a == FramebufferAttachment::DEPTH_STENCIL) fb_->depth = rb.object();
  }
  bool valid() const { return true; }
  void bind() { sgl::ctx().id_of(fb_); sgl::ctx().fbo = fb_; }
  void release() { sgl::ctx().fbo.reset(); }
  uint32_t width() const { return (uint32_t)fb_->w; }
  uint32_t height() const { return (uint32_t)fb_->h; }
  void resize(uint32_t w, uint32_t h) { fb_->w = (int)w; fb_->h = (int)h; }

 private:
  void attach_tex(FramebufferAttachment a, const std::shared_ptr<sgl::Texture>& t) {
    int k = (int)a - (int)FramebufferAttachment::COLOR0;
    if (k >= 0 && k < sgl::kMaxColor) fb_->color[k] = t;
  }
  std::shared_ptr<sgl::Framebuffer> fb_ = std::make_shared<sgl::Framebuffer>();
};

class GlShader {
 public:
  static GlShader fromCache(ShaderType type, const std::string& filename) { return GlShader(type, filename); }
  static GlShader fromFile(ShaderType type, const std::string& filename) { return GlShader(type, filename); }
  ShaderType type() const { return type_; }
  const std::string& filename() const { return filename_; }

 private:
  GlShader(ShaderType t, const std::string& f) : type_(t), filename_(f) {}
  ShaderType type_;
  std::string filename_;
};

template <class T>
class GlUniform {
 public:
  GlUniform(const std::string& name, const T& value) : name_(name), value_(value) {}
  const std::string& name() const { return name_; }
  const T& value() const { return value_; }
  GlUniform& operator=(const T& v) { value_ = v; return *this; }

 private:
  std::string name_;
  T value_;
};

inline sgl::UVal to_uval(float v) { sgl::UVal u; u.kind = sgl::UVal::F; u.f[0] = v; return u; }
inline sgl::UVal to_uval(double v) { return to_uval((float)v); }
inline sgl::UVal to_uval(int32_t v) { sgl::UVal u; u.kind = sgl::UVal::I; u.i = v; return u; }
inline sgl::UVal to_uval(uint32_t v) { sgl::UVal u; u.kind = sgl::UVal::U; u.i = (int)v; return u; }
inline sgl::UVal to_uval(bool v) { sgl::UVal u; u.kind = sgl::UVal::B; u.i = v ? 1 : 0; return u; }
inline sgl::UVal to_uval(const vec2& v) { sgl::UVal u; u.kind = sgl::UVal::V2; u.f[0] = v.x; u.f[1] = v.y; return u; }
inline sgl::UVal to_uval(const vec3& v) { sgl::UVal u; u.kind = sgl::UVal::V3; u.f[0] = v.x; u.f[1] = v.y; u.f[2] = v.z; return u; }
inline sgl::UVal to_uval(const vec4& v) { sgl::UVal u; u.kind = sgl::UVal::V4; u.f[0] = v.x; u.f[1] = v.y; u.f[2] = v.z; u.f[3] = v.w; return u; }
inline sgl::UVal to_uval(const Eigen::Matrix4f& m) {
  sgl::UVal u; u.kind = sgl::UVal::M4;
  for (int i = 0; i < 16; ++i) u.f[i] = m.data()[i];  // column-major
  return u;
}

class GlTransformFeedback {
 public:
  template <class T>
  void attach(const std::vector<std::string>& varyings, GlBuffer<T>& buffer) {
    tf_->varyings = varyings;
    tf_->buffer = buffer.object();
  }
  void bind() { sgl::ctx().tf = tf_; }
  void release() { sgl::ctx().tf.reset(); }
  void begin(TransformFeedbackMode) {
    tf_->active = true;
    tf_->written_bytes = 0;
    tf_->primitives = 0;
  }
  uint32_t end() {  // number of primitives written (TRANSFORM_FEEDBACK_PRIMITIVES_WRITTEN)
    tf_->active = false;
    return tf_->primitives;
  }
  const std::shared_ptr<sgl::TransformFeedback>& object() const { return tf_; }

 private:
  std::shared_ptr<sgl::TransformFeedback> tf_ = std::make_shared<sgl::TransformFeedback>();
};

class GlProgram {
 public:
  void attach(const GlShader& s) {
    if (s.type() == ShaderType::VERTEX_SHADER) p_->vs_name = s.filename();
    else if (s.type() == ShaderType::GEOMETRY_SHADER) p_->gs_name = s.filename();
    else if (s.type() == ShaderType::FRAGMENT_SHADER) p_->fs_name = s.filename();
  }
  void attach(const GlTransformFeedback& tf) { p_->tf = tf.object(); }
  void link() { sgl::link(*p_); }
  void bind() { sgl::ctx().program = p_; }
  void release() { sgl::ctx().program.reset(); }
  template <class T>
  void setUniform(const GlUniform<T>& u) { sgl::set_uniform(*p_, u.name(), to_uval(u.value())); }

 private:
  std::shared_ptr<sgl::Program> p_ = std::make_shared<sgl::Program>();
};

class GlVertexArray {
 public:
  template <class T>
  void setVertexAttribute(uint32_t idx, GlBuffer<T>& buffer, int32_t size, AttributeType type, bool /*normalized*/,
                          uint32_t stride, const GLvoid* offset) {
    sgl::VertexArray::Attr& a = va_->attr[idx];
    a.buf = buffer.object();
    a.size = size;
    a.integer = type == AttributeType::INT || type == AttributeType::UNSIGNED_INT;  // glVertexAttribIPointer
    a.type = type == AttributeType::INT ? GL_INT : type == AttributeType::UNSIGNED_INT ? GL_UNSIGNED_INT : GL_FLOAT;
    a.stride = stride;
    a.offset = (size_t)offset;
    a.enabled = true;  // glow enables the array together with the pointer
  }
  void enableVertexAttribute(uint32_t idx) { va_->attr[idx].enabled = true; }
  void disableVertexAttribute(uint32_t idx) { va_->attr[idx].enabled = false; }
  void bind() { sgl::ctx().vao = va_; }
  void release() { sgl::ctx().vao.reset(); }

 private:
  std::shared_ptr<sgl::VertexArray> va_ = std::make_shared<sgl::VertexArray>();
};

class GlQuery {
 public:
  GlQuery(QueryTarget = QueryTarget::TIME_ELAPSED) {}
  void begin() {}
  void end() {}
  bool ready() const { return true; }
  template <class T>
  void value(T& v) const { v = T(0); }
  operator uint32_t() const { return 0; }
  operator int32_t() const { return 0; }
};

template <class T>
class ScopedBinder {
 public:
  explicit ScopedBinder(T& o) : o_(o) { o_.bind(); }
  ~ScopedBinder() { o_.release(); }

 private:
  T& o_;
};

class GlState {
 public:
  static GlState queryAll() { return GlState(); }
  void restore() {}
  bool operator==(const GlState&) const { return true; }
  bool operator!=(const GlState&) const { return false; }
};

}  // namespace glow
