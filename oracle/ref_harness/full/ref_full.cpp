// ref_full.cpp -- TEST INFRASTRUCTURE (oracle/_ref/libsuma_ref_full.so). C entry points into the reference's OWN core
// classes, compiled where they lie under /root/reference/src (../Makefile):
//   core/Preprocessing.cpp   core/Frame2Model.cpp   core/SurfelMap.cpp   core/SurfelMapping.cpp
//   core/LieGaussNewton.cpp  core/lie_algebra.cpp   rv/*.cpp (parameters, Laserscan)   util/kitti_utils.cpp
// driving the reference's own shaders (../glsl2cpp.py -> oracle/_ref/gen) through the stand-in glow / software GL of
// this directory. This file only marshals plain arrays in and out of those classes; it restates none of their logic.
// What is NOT the reference here: glow + the GL implementation (sgl.hpp, glow/glow_all.hpp), Eigen (../host/eigen3),
// Boost (../host/boost), the gtsam-backed pose graph (posegraph_stub.cpp: stores poses and edges, optimises nothing).
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "core/Frame2Model.h"
#include "core/Preprocessing.h"
#include "core/SurfelMap.h"
#include "core/SurfelMapping.h"
#include "rv/ParameterList.h"
#include "rv/PrimitiveParameters.h"

extern "C" {
#include "../../orc.h"  // orc_params / orc_surfel: the parameter block and the 64-byte Surfel record shared with the oracle
}

namespace {

// config/default.xml keys from the oracle's parameter block (names: config/default.xml, the classes' setParameters)
rv::ParameterList make_params(const orc_params* p, const char* xml) {
  rv::ParameterList L;
  if (xml && xml[0]) rv::parseXmlFile(xml, L);  // the reference's own file first (keys the block does not carry)
  auto I = [&](const char* n, int v) { L.insert(rv::IntegerParameter(n, v)); };
  auto F = [&](const char* n, double v) { L.insert(rv::FloatParameter(n, v)); };
  auto B = [&](const char* n, bool v) { L.insert(rv::BooleanParameter(n, v)); };
  I("data_width", p->data_width); I("data_height", p->data_height);
  F("data_fov_up", p->data_fov_up); F("data_fov_down", p->data_fov_down);
  F("min_depth", p->min_depth); F("max_depth", p->max_depth);
  I("model_width", p->model_width); I("model_height", p->model_height);
  F("model_fov_up", p->model_fov_up); F("model_fov_down", p->model_fov_down);
  F("model_min_depth", p->model_min_depth); F("model_max_depth", p->model_max_depth);
  I("max iterations", p->max_iterations);
  F("stopping threshold", p->stopping_threshold); F("delta", p->delta);
  F("icp-max-distance", p->icp_max_distance); F("icp-max-angle", p->icp_max_angle);
  const char* wn = p->weighting == 1 || p->weighting == 4 ? "huber" : p->weighting == 2 ? "turkey" : p->weighting == 3 ? "stability" : "none";
  L.insert(rv::StringParameter("weighting", wn));
  F("factor", p->factor);
  B("initialize_identity", p->initialize_identity != 0);
  B("bilinear_sampling", p->bilinear_sampling != 0);
  B("fallback_mode", p->fallback_mode != 0);
  F("fallback-max-distance", p->fallback_max_distance); F("fallback-max-angle", p->fallback_max_angle);
  B("compose_rendering", p->compose_rendering != 0);
  F("max_loop_closure_distance", p->max_loop_closure_distance);
  F("min_radius", p->min_radius); F("max_radius", p->max_radius); F("max_angle", p->max_angle);
  F("map-max-distance", p->map_max_distance); F("map-max-angle", p->map_max_angle);
  I("unstable_age", p->unstable_age); I("confidence_mode", p->confidence_mode);
  F("confidence_threshold", p->confidence_threshold);
  F("p_stable", p->p_stable); F("p_prior", p->p_prior);
  F("sigma_angle", p->sigma_angle); F("sigma_distance", p->sigma_distance);
  B("use_stability", p->use_stability != 0); I("active_timestamps", p->active_timestamps);
  F("max_weight", p->max_weight);
  I("weighting_scheme", p->weighting_scheme); I("averaging_scheme", p->averaging_scheme);
  B("update_always", p->update_always != 0);
  I("submap-dimension", p->submap_dimension); F("submap-extent", p->submap_extent);
  B("partial-extraction", p->partial_extraction != 0);
  // keys the classes read unconditionally and the block does not carry: the values of config/default.xml (when the
  // reference's file was parsed above these are already present and stay)
  if (!L.hasParam("cutoff_threshold")) F("cutoff_threshold", 10.0);
  if (!L.hasParam("close-loops")) B("close-loops", false);
  return L;
}

void upload(glow::GlTextureRectangle& t, const float* src) {
  if (src) t.assign(glow::PixelFormat::RGBA, glow::PixelType::FLOAT, src);
}
void download(const glow::GlTextureRectangle& t, float* dst) {
  if (dst) t.download(glow::PixelFormat::RGBA, dst);
}
Eigen::Matrix4f m4f(const float* m) {  // column-major
  Eigen::Matrix4f M;
  for (int i = 0; i < 16; ++i) M.data()[i] = m[i];
  return M;
}
Eigen::Matrix4d m4d(const double* m) {
  Eigen::Matrix4d M;
  for (int i = 0; i < 16; ++i) M.data()[i] = m[i];
  return M;
}

}  // namespace

struct reffull {
  orc_params p;
  rv::ParameterList params;
  std::unique_ptr<Preprocessing> prep;
  std::unique_ptr<Frame2Model> f2m;
  std::unique_ptr<SurfelMap> map;
  std::unique_ptr<SurfelMapping> slam;
  std::string error;
};

extern "C" {

reffull* reffull_create(const orc_params* p, const char* default_xml) {
  reffull* h = new reffull();
  h->p = *p;
  try {
    h->params = make_params(p, default_xml);
  } catch (const std::exception& e) {
    h->error = e.what();
  }
  return h;
}
void reffull_destroy(reffull* h) { delete h; }
const char* reffull_error(const reffull* h) { return h->error.c_str(); }
// one parameter of the list the classes are constructed with, as text ("" if absent)
int reffull_param(reffull* h, const char* name, char* out, int cap) {
  if (!h->params.hasParam(name)) return -1;
  std::string v = h->params[name].valueStr();
  strncpy(out, v.c_str(), (size_t)cap - 1);
  out[cap - 1] = 0;
  return 0;
}
void reffull_set_float(reffull* h, const char* name, double v) { h->params.insert(rv::FloatParameter(name, v)); }
void reffull_set_int(reffull* h, const char* name, int v) { h->params.insert(rv::IntegerParameter(name, v)); }
void reffull_set_bool(reffull* h, const char* name, int v) { h->params.insert(rv::BooleanParameter(name, v != 0)); }

// Preprocessing::process (core/Preprocessing.cpp:120-339)
int reffull_preprocess(reffull* h, const float* pts4, const float* labels, const float* probs, uint32_t n, uint32_t timestamp,
                       float* vertex_map, float* normal_map, float* semantic_map) {
  try {
    if (!h->prep) h->prep.reset(new Preprocessing(h->params));
    glow::GlBuffer<rv::Point3f> pts{glow::BufferTarget::ARRAY_BUFFER, glow::BufferUsage::DYNAMIC_READ};
    glow::GlBuffer<float> lab{glow::BufferTarget::ARRAY_BUFFER, glow::BufferUsage::DYNAMIC_READ};
    glow::GlBuffer<float> prb{glow::BufferTarget::ARRAY_BUFFER, glow::BufferUsage::DYNAMIC_READ};
    std::vector<rv::Point3f> P(n);
    for (uint32_t i = 0; i < n; ++i) {
      P[i] = rv::Point3f(pts4[4 * i], pts4[4 * i + 1], pts4[4 * i + 2]);
      P[i].vec[3] = pts4[4 * i + 3];
    }
    std::vector<float> Lb(n, 0.0f), Pb(n, 0.0f);
    if (labels) Lb.assign(labels, labels + n);
    if (probs) Pb.assign(probs, probs + n);
    pts.assign(P); lab.assign(Lb); prb.assign(Pb);
    Frame frame(h->p.data_width, h->p.data_height);
    h->prep->process(pts, frame, lab, prb, timestamp);
    download(frame.vertex_map, vertex_map);
    download(frame.normal_map, normal_map);
    download(frame.semantic_map, semantic_map);
    return 0;
  } catch (const std::exception& e) {
    h->error = e.what();
    return -1;
  }
}

// Frame2Model::jacobianProducts (core/Frame2Model.cpp:136-261) at `pose`, `iteration`. out48 = JtJ (36, column-major) |
// Jtf (6) | valid, F, outlier, inlier_residual, invalid, 0 as the class reports them after its float -> uint32 conversions.
int reffull_icp_jacobian(reffull* h, const float* data_v, const float* data_n, const float* data_s, const float* model_v,
                         const float* model_n, const float* model_s, const double pose[16], int32_t iteration,
                         float max_distance, float max_angle_deg, double out48[48]) {
  try {
    rv::ParameterList params = h->params;
    params.insert(rv::FloatParameter("icp-max-distance", max_distance));
    params.insert(rv::FloatParameter("icp-max-angle", max_angle_deg));
    Frame2Model f(params);
    auto cur = std::make_shared<Frame>(h->p.data_width, h->p.data_height);
    auto mod = std::make_shared<Frame>(h->p.model_width, h->p.model_height);
    upload(cur->vertex_map, data_v); upload(cur->normal_map, data_n); upload(cur->semantic_map, data_s);
    upload(mod->vertex_map, model_v); upload(mod->normal_map, model_n); upload(mod->semantic_map, model_s);
    f.setData(cur, mod);
    f.initialize(m4d(pose));
    // Objective::increment bumps iteration_ (Objective.h:45-48); reach `iteration` the way LieGaussNewton does
    for (int i = 0; i < iteration; ++i) {
      Eigen::VectorXd zero = Eigen::VectorXd::Zero(6);
      f.increment(zero);
    }
    f.initialize(m4d(pose));  // exp(0) * pose == pose, but make the pose exact
    Eigen::MatrixXd JtJ(6, 6), Jtf(6, 1);
    double F = f.jacobianProducts(JtJ, Jtf);
    for (int i = 0; i < 36; ++i) out48[i] = JtJ.data()[i];
    for (int i = 0; i < 6; ++i) out48[36 + i] = Jtf.data()[i];
    out48[42] = f.valid(); out48[43] = F; out48[44] = f.outlier(); out48[45] = f.inlier_residual();
    out48[46] = f.invalid(); out48[47] = 0.0;
    return 0;
  } catch (const std::exception& e) {
    h->error = e.what();
    return -1;
  }
}

uint64_t reffull_draw_calls(void) { return sgl::ctx().draw_calls; }

}  // extern "C"
