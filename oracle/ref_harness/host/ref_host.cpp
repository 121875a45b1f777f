// ref_host.cpp -- TEST INFRASTRUCTURE (oracle/_ref/libsuma_ref_host.so). C entry points into the reference's OWN host
// sources, compiled where they lie under /root/reference/src (see ../Makefile):
//   core/lie_algebra.cpp      SE3::exp / SE3::log                                  (hot path row a7)
//   core/LieGaussNewton.cpp   minimize / initialize / step, with core/Objective.h  (a7)
//   util/kitti_utils.cpp      KITTICalibration, KITTI::Odometry::*                 (row f4)
//   rv/*.cpp                  ParameterList, XML parser (config/default.xml), string_utils
// Nothing below restates their arithmetic or control flow; it only marshals plain arrays. What stands in for libraries
// that are not in this image is in this directory: eigen3/Eigen/Dense, boost/*.hpp (see their headers), frame_stub.h.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "core/LieGaussNewton.h"
#include "core/lie_algebra.h"
#include "rv/ParameterList.h"
#include "rv/PrimitiveParameters.h"
#include "util/kitti_utils.h"

namespace {

Eigen::Matrix4d m4_from_colmajor(const double* m) {
  Eigen::Matrix4d M;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) M(r, c) = m[4 * c + r];
  return M;
}
void m4_to_colmajor(const Eigen::Matrix4d& M, double* m) {
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) m[4 * c + r] = M(r, c);
}
Eigen::Matrix4f m4f_from_rowmajor(const float* m) {
  Eigen::Matrix4f M;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) M(r, c) = m[4 * r + c];
  return M;
}

// jacobianProducts supplied by the caller: fills JtJ (6x6 column-major), Jtf (6) and returns the error F for the
// objective's current pose (column-major) and iteration. Everything around it -- increment(), pose(), the iteration
// counter -- is the reference's Objective.
typedef double (*ref_products_fn)(void* user, const double pose[16], int iteration, double JtJ[36], double Jtf[6]);

class CallbackObjective : public Objective {
 public:
  CallbackObjective(ref_products_fn fn, void* user) : fn_(fn), user_(user) {}
  uint32_t num_parameters() const override { return 6; }
  double residual(const Eigen::VectorXd&) override { return 0.0; }
  double jacobianProducts(Eigen::MatrixXd& JtJ, Eigen::MatrixXd& Jtf) override {
    double pose[16], A[36], b[6];
    m4_to_colmajor(pose_, pose);
    double F = fn_(user_, pose, (int)iteration_, A, b);
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) JtJ(r, c) = A[6 * c + r];
    for (int r = 0; r < 6; ++r) Jtf(r, 0) = b[r];
    return F;
  }
  void reset_iteration() { iteration_ = 0; }

 private:
  ref_products_fn fn_;
  void* user_;
};

}  // namespace

extern "C" {

// core/lie_algebra.cpp:4-34 ; T column-major
void ref_se3_exp(const double x[6], double T[16]) {
  Eigen::VectorXd v = Eigen::VectorXd::Zero(6);
  for (int i = 0; i < 6; ++i) v[i] = x[i];
  m4_to_colmajor(SE3::exp(v), T);
}
// core/lie_algebra.cpp:36-71
void ref_se3_log(const double T[16], double x[6]) {
  Eigen::VectorXd v = SE3::log(m4_from_colmajor(T));
  for (int i = 0; i < 6; ++i) x[i] = v[i];
}

// LieGaussNewton::minimize (core/LieGaussNewton.cpp:13-36) on a caller-supplied objective. Parameters go through the
// reference's own setParameters / ParameterList. history: up to hist_cap poses (column-major); returns minimize()'s value.
int ref_gn_minimize(ref_products_fn fn, void* user, const double T0[16], int max_iter, double eps, double delta,
                    double pose_out[16], double* history, int hist_cap, int* hist_len, int* iterations,
                    double* residual) {
  CallbackObjective obj(fn, user);
  LieGaussNewton gn;
  rv::ParameterList params;
  params.insert(rv::IntegerParameter("max iterations", max_iter));
  params.insert(rv::FloatParameter("stopping threshold", eps));
  params.insert(rv::FloatParameter("delta", delta));
  gn.setParameters(params);
  int32_t ret = gn.minimize(obj, m4_from_colmajor(T0));
  m4_to_colmajor(gn.pose(), pose_out);
  const std::vector<Eigen::Matrix4d>& h = gn.history();
  if (hist_len) *hist_len = (int)h.size();
  for (int i = 0; i < (int)h.size() && i < hist_cap; ++i) m4_to_colmajor(h[i], history + 16 * (size_t)i);
  if (iterations) *iterations = (int)gn.iterationCount();
  if (residual) *residual = gn.residual();
  return ret;
}

// rv::parseXmlFile (rv/ParameterList.cpp:127) + the conversion operators of rv::Parameter: value of one parameter of an
// XML configuration as text; returns 0, -1 if absent, -2 on a parse error
int ref_param_lookup(const char* xml_file, const char* name, char* out, int cap) {
  try {
    rv::ParameterList params;
    rv::parseXmlFile(xml_file, params);
    if (!params.hasParam(name)) return -1;
    std::string v = params[name].valueStr();
    strncpy(out, v.c_str(), (size_t)cap - 1);
    out[cap - 1] = 0;
    return 0;
  } catch (...) {
    return -2;
  }
}
// all parameter names of the file, '\n'-separated
int ref_param_names(const char* xml_file, char* out, int cap) {
  try {
    rv::ParameterList params;
    rv::parseXmlFile(xml_file, params);
    std::string all;
    for (rv::ParameterList::const_iterator it = params.begin(); it != params.end(); ++it) all += it->name() + "\n";
    strncpy(out, all.c_str(), (size_t)cap - 1);
    out[cap - 1] = 0;
    return (int)params.size();
  } catch (...) {
    return -2;
  }
}

// ---- util/kitti_utils.cpp; poses row-major 4x4 float as in the KITTI text files ----
// returns the number of poses, or -2 when loadPoses throws (boost::bad_lexical_cast on an entry that is not a number --
// e.g. the empty token a leading or doubled blank produces, since rv::split keeps empty tokens)
int ref_kitti_load_poses(const char* file, float* out, int cap) {
  std::vector<Eigen::Matrix4f> poses;
  try {
    poses = KITTI::Odometry::loadPoses(file);
  } catch (...) {
    return -2;
  }
  for (int i = 0; i < (int)poses.size() && i < cap; ++i)
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) out[16 * (size_t)i + 4 * r + c] = poses[i](r, c);
  return (int)poses.size();
}
int ref_kitti_calibration(const char* file, const char* name, float out[16]) {
  try {
    KITTICalibration calib(file);
    if (!calib.exists(name)) return -1;
    const Eigen::Matrix4f& M = calib[name];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) out[4 * r + c] = M(r, c);
    return 0;
  } catch (...) {
    return -2;
  }
}
void ref_kitti_trajectory_distances(const float* poses, int n, float* dist) {
  std::vector<Eigen::Matrix4f> P;
  for (int i = 0; i < n; ++i) P.push_back(m4f_from_rowmajor(poses + 16 * (size_t)i));
  std::vector<float> d = KITTI::Odometry::trajectoryDistances(P);
  for (size_t i = 0; i < d.size(); ++i) dist[i] = d[i];
}
float ref_kitti_rotation_error(const float pose_error[16]) {
  return KITTI::Odometry::rotationError(m4f_from_rowmajor(pose_error));
}
float ref_kitti_translation_error(const float pose_error[16]) {
  return KITTI::Odometry::translationError(m4f_from_rowmajor(pose_error));
}
int ref_kitti_last_frame(const float* dist, int n, int first_frame, float len) {
  std::vector<float> d(dist, dist + n);
  return KITTI::Odometry::lastFrameFromSegmentLength(d, first_frame, len);
}
// calcSequenceErrors: out = rows of {first_frame, r_err, t_err, len, speed}; returns the number of rows
int ref_kitti_sequence_errors(const float* gt, const float* res, int n, float* out, int cap) {
  std::vector<Eigen::Matrix4f> G, R;
  for (int i = 0; i < n; ++i) {
    G.push_back(m4f_from_rowmajor(gt + 16 * (size_t)i));
    R.push_back(m4f_from_rowmajor(res + 16 * (size_t)i));
  }
  std::vector<KITTI::Odometry::errors> e = KITTI::Odometry::calcSequenceErrors(G, R);
  for (int i = 0; i < (int)e.size() && i < cap; ++i) {
    out[5 * (size_t)i + 0] = (float)e[i].first_frame;
    out[5 * (size_t)i + 1] = e[i].r_err;
    out[5 * (size_t)i + 2] = e[i].t_err;
    out[5 * (size_t)i + 3] = e[i].len;
    out[5 * (size_t)i + 4] = e[i].speed;
  }
  return (int)e.size();
}
// saveStats (kitti_utils.cpp:479-504) writes "stats.txt" = mean t_err, mean r_err into `dir`
void ref_kitti_save_stats(const float* rows, int n, const char* dir) {
  std::vector<KITTI::Odometry::errors> e;
  for (int i = 0; i < n; ++i)
    e.push_back(KITTI::Odometry::errors((int32_t)rows[5 * (size_t)i], rows[5 * (size_t)i + 1], rows[5 * (size_t)i + 2],
                                        rows[5 * (size_t)i + 3], rows[5 * (size_t)i + 4]));
  KITTI::Odometry::saveStats(e, dir);
}

}  // extern "C"
