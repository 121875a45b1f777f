// ref_full.cpp -- TEST INFRASTRUCTURE (oracle/_ref/libsuma_ref_full.so). C entry points into the reference's OWN core
// classes, compiled where they lie under /root/reference/src (../Makefile):
//   core/Preprocessing.cpp   core/Frame2Model.cpp   core/SurfelMap.cpp   core/SurfelMapping.cpp
//   core/LieGaussNewton.cpp  core/lie_algebra.cpp   rv/*.cpp (parameters, Laserscan)   util/kitti_utils.cpp
// driving the reference's own shaders (../glsl2cpp.py -> oracle/_ref/gen) through the stand-in glow / software GL of
// this directory. This file only marshals plain arrays in and out of those classes; it restates none of their logic.
// What is NOT the reference here: glow + the GL implementation (sgl.hpp, glow/glow_all.hpp), Eigen (../host/eigen3),
// Boost (../host/boost), the gtsam-backed pose graph (posegraph_stub.cpp: stores poses and edges, optimises nothing).
#include <cmath>
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
  if (!L.hasParam("approach")) L.insert(rv::StringParameter("approach", "frame-to-model"));
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

// LieGaussNewton::minimize (core/LieGaussNewton.cpp:13-36) on Frame2Model with the list's parameters; history: the poses
// pushed per iteration (column-major), errors: residual() after the run. Returns iterationCount().
int reffull_icp_minimize(reffull* h, const float* data_v, const float* data_n, const float* data_s, const float* model_v,
                         const float* model_n, const float* model_s, const double T0[16], double pose_out[16],
                         double* history, int hist_cap, int* hist_len) {
  try {
    Frame2Model f(h->params);
    auto cur = std::make_shared<Frame>(h->p.data_width, h->p.data_height);
    auto mod = std::make_shared<Frame>(h->p.model_width, h->p.model_height);
    upload(cur->vertex_map, data_v); upload(cur->normal_map, data_n); upload(cur->semantic_map, data_s);
    upload(mod->vertex_map, model_v); upload(mod->normal_map, model_n); upload(mod->semantic_map, model_s);
    f.setData(cur, mod);
    LieGaussNewton gn;
    gn.setParameters(h->params);
    gn.minimize(f, m4d(T0));
    for (int i = 0; i < 16; ++i) pose_out[i] = gn.pose().data()[i];
    const std::vector<Eigen::Matrix4d>& hs = gn.history();
    if (hist_len) *hist_len = (int)hs.size();
    for (int k = 0; k < (int)hs.size() && k < hist_cap; ++k)
      for (int i = 0; i < 16; ++i) history[16 * (size_t)k + i] = hs[k].data()[i];
    return (int)gn.iterationCount();
  } catch (const std::exception& e) {
    h->error = e.what();
    return -1;
  }
}

uint64_t reffull_draw_calls(void) { return sgl::ctx().draw_calls; }
// 1: attribute reads behind the last uploaded element return 0 instead of the previous upload's data (sgl.hpp)
void reffull_zero_stale_tail(int on) { sgl::ctx().zero_stale_tail = on != 0; }

// ---------------------------------------------------------------------------------------------- SurfelMap (core/SurfelMap.cpp)
// test access to two protected members of SurfelMap (the class itself is untouched): put a given surfel set and pose
// table in place, as a sequence of update() calls would have
struct MapAccess : public SurfelMap {
  using SurfelMap::SurfelMap;
  void upload(const std::vector<Surfel>& s, uint32_t timestamp) {
    surfels_.assign(s);
    timestamp_ = timestamp;
  }
  void set_pose(uint32_t t, const Eigen::Matrix4f& P) {  // what update() does for its own timestamp (SurfelMap.cpp:494-495)
    if (t >= poses_.size()) return;
    poses_[t] = P;
    poseBuffer_.insert(t, P);
  }
};
static SurfelMap* the_map(reffull* h) {
  if (h->slam) return h->slam->getMap().get();
  if (!h->map) h->map.reset(new MapAccess(h->params));
  return h->map.get();
}
static void frame_from(Frame& f, const float* v, const float* n, const float* s) {
  upload(f.vertex_map, v); upload(f.normal_map, n); upload(f.semantic_map, s);
  f.valid = true;
}
static void frame_to(const Frame& f, float* v, float* n, float* s) {
  download(f.vertex_map, v); download(f.normal_map, n); download(f.semantic_map, s);
}
#define REFFULL_TRY(body) try { body; return 0; } catch (const std::exception& e) { h->error = e.what(); return -1; }

uint32_t reffull_map_size(reffull* h) { return the_map(h)->size(); }
uint32_t reffull_map_download(reffull* h, orc_surfel* dst, uint32_t cap) {
  std::vector<Surfel> all = the_map(h)->getAllSurfels();
  static_assert(sizeof(Surfel) == sizeof(orc_surfel), "Surfel record layout (core/Surfel.h:5-15)");
  uint32_t k = std::min<uint32_t>(cap, (uint32_t)all.size());
  if (k) memcpy(dst, all.data(), sizeof(orc_surfel) * k);
  return k;
}
int reffull_map_upload(reffull* h, const orc_surfel* src, uint32_t n, uint32_t timestamp) {
  REFFULL_TRY({
    if (h->slam) throw std::runtime_error("map_upload: only for a stand-alone SurfelMap");
    the_map(h);
    std::vector<Surfel> v(n);
    if (n) memcpy(v.data(), src, sizeof(Surfel) * n);
    static_cast<MapAccess*>(h->map.get())->upload(v, timestamp);
  })
}
int reffull_map_set_pose(reffull* h, uint32_t t, const float pose[16]) {
  REFFULL_TRY({
    if (h->slam) throw std::runtime_error("map_set_pose: only for a stand-alone SurfelMap");
    the_map(h);
    static_cast<MapAccess*>(h->map.get())->set_pose(t, m4f(pose));
  })
}
// SurfelMap::update (SurfelMap.cpp:492-584), including updateActiveSubmaps
int reffull_map_update(reffull* h, const float pose[16], const float* fv, const float* fn, const float* fs) {
  REFFULL_TRY({
    Frame frame(h->p.data_width, h->p.data_height);
    frame_from(frame, fv, fn, fs);
    the_map(h)->update(m4f(pose), frame);
  })
}
int reffull_map_render(reffull* h, const float pose_old[16], const float pose_new[16], float conf_thr, float* fv, float* fn,
                       float* fs) {
  REFFULL_TRY({
    Frame frame(h->p.model_width, h->p.model_height);
    the_map(h)->render(m4f(pose_old), m4f(pose_new), frame, conf_thr);
    frame_to(frame, fv, fn, fs);
  })
}
int reffull_map_render_active(reffull* h, const float pose[16], float conf_thr) { REFFULL_TRY(the_map(h)->render_active(m4f(pose), conf_thr)) }
int reffull_map_render_inactive(reffull* h, const float pose[16], float conf_thr) { REFFULL_TRY(the_map(h)->render_inactive(m4f(pose), conf_thr)) }
int reffull_map_render_composed(reffull* h, const float pose_old[16], const float pose_new[16], float conf_thr) {
  REFFULL_TRY(the_map(h)->render_composed(m4f(pose_old), m4f(pose_new), conf_thr))
}
int reffull_map_get_frame(reffull* h, int which, float* v, float* n, float* s) {
  REFFULL_TRY({
    std::shared_ptr<Frame>& f = which == 0 ? the_map(h)->oldMapFrame() : which == 1 ? the_map(h)->newMapFrame() : the_map(h)->composedFrame();
    frame_to(*f, v, n, s);
  })
}

// ---------------------------------------------------------------------------------------------- SurfelMapping (core/SurfelMapping.cpp)
int reffull_slam_create(reffull* h) { REFFULL_TRY(h->slam.reset(new SurfelMapping(h->params))) }
// SurfelMapping::processScan (SurfelMapping.cpp:175-210); labels / probs may be null (geometric run: zeros)
int reffull_slam_process_scan(reffull* h, const float* pts4, const float* labels, const float* probs, uint32_t n) {
  REFFULL_TRY({
    if (!h->slam) h->slam.reset(new SurfelMapping(h->params));
    rv::Laserscan scan;
    scan.points().resize(n);
    for (uint32_t i = 0; i < n; ++i) scan.points()[i] = rv::Point3f(pts4[4 * i], pts4[4 * i + 1], pts4[4 * i + 2]);
    scan.labels_float.assign(n, 0.0f);
    scan.labels_prob.assign(n, 0.0f);
    if (labels) scan.labels_float.assign(labels, labels + n);
    if (probs) scan.labels_prob.assign(probs, probs + n);
    h->slam->processScan(scan);
  })
}
uint32_t reffull_slam_timestamp(reffull* h) { return h->slam ? h->slam->timestamp() : 0; }
void reffull_slam_pose(reffull* h, double pose[16]) {
  const Eigen::Matrix4d& P = h->slam->getCurrentPose();
  for (int i = 0; i < 16; ++i) pose[i] = P.data()[i];
}
void reffull_slam_last_pose(reffull* h, double pose[16]) {
  const Eigen::Matrix4d& P = h->slam->getLastPose();
  for (int i = 0; i < 16; ++i) pose[i] = P.data()[i];
}
// which: 0 currentFrame, 1 lastFrame, 2 currentModelFrame, 3 lastModelFrame
int reffull_slam_frame(reffull* h, int which, float* v, float* n, float* s) {
  REFFULL_TRY({
    Frame::Ptr f = which == 0 ? h->slam->getCurrentFrame() : which == 1 ? h->slam->getLastFrame()
                 : which == 2 ? h->slam->getCurrentModelFrame() : h->slam->getLastModelFrame();
    frame_to(*f, v, n, s);
  })
}
// one entry of SurfelMapping::getStatistics() (NaN if absent)
double reffull_slam_statistic(reffull* h, const char* name) {
  const SurfelMapping::Stats& st = h->slam->getStatistics();
  auto it = st.find(name);
  return it == st.end() ? std::nan("") : (double)it->second;
}
int reffull_slam_found_loop_candidate(reffull* h) { return h->slam->foundLoopClosureCandidate() ? 1 : 0; }
int reffull_slam_use_loop_candidate(reffull* h) { return h->slam->useLoopClosureCandidate() ? 1 : 0; }
uint32_t reffull_slam_edges(reffull* h, int32_t* from_to, uint32_t cap) {
  const std::vector<Posegraph::Edge>& e = h->slam->getPosegraph()->getEdges();
  for (uint32_t i = 0; i < e.size() && i < cap; ++i) { from_to[2 * i] = e[i].from; from_to[2 * i + 1] = e[i].to; }
  return (uint32_t)e.size();
}

}  // extern "C"
