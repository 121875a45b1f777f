// posegraph_stub.cpp -- TEST INFRASTRUCTURE (oracle/_ref/libsuma_ref_full.so). core/Posegraph.cpp is a thin wrapper of
// gtsam 4.0 (not in this image; the pose-graph optimiser is outside the hot path, SURVEY.md section 8 "out of scope").
// This file implements the interface core/Posegraph.h declares without an optimiser: poses and edges are stored,
// optimize() leaves them as they are. SurfelMapping::processScan uses setInitial / addEdge / pose / size on the path.
#include <map>

#include <eigen3/Eigen/Dense>
#include <memory>
#include <string>
#include <vector>

#include "core/Posegraph.h"

namespace {
std::map<const Posegraph*, std::map<int32_t, Eigen::Matrix4d>>& store() {
  static std::map<const Posegraph*, std::map<int32_t, Eigen::Matrix4d>> s;
  return s;
}
gtsam::NonlinearFactorGraph g_graph;
gtsam::Values g_values;
}  // namespace

Posegraph::Posegraph() {}
Posegraph::Ptr Posegraph::clone() const {
  Posegraph::Ptr c(new Posegraph());
  c->edges_ = edges_;
  store()[c.get()] = store()[this];
  return c;
}
double Posegraph::error() const { return 0.0; }
void Posegraph::clear() {
  edges_.clear();
  store()[this].clear();
}
void Posegraph::setInitial(int32_t id, const Eigen::Matrix4d& initial_estimate) { store()[this][id] = initial_estimate; }
void Posegraph::addEdge(int32_t from, int32_t to, const Eigen::Matrix4d& measurement, const Matrix6d&) {
  edges_.push_back(Edge(from, to, measurement));
}
Eigen::Matrix4d Posegraph::pose(int32_t id) const {
  auto& m = store()[this];
  auto it = m.find(id);
  if (it == m.end()) return Eigen::Matrix4d::Identity();
  return it->second;
}
std::vector<Eigen::Matrix4d> Posegraph::poses() const {
  std::vector<Eigen::Matrix4d> out;
  for (auto& kv : store()[this]) out.push_back(kv.second);
  return out;
}
int32_t Posegraph::size() const { return (int32_t)store()[this].size(); }
void Posegraph::reinitialize() {}
bool Posegraph::optimize(uint32_t) { return true; }
void Posegraph::save(const std::string&) const {}
void Posegraph::load(const std::string&) {}
void Posegraph::setMEstimator(const gtsam::noiseModel::mEstimator::Base::shared_ptr& m) { robustifier_ = m; }
const gtsam::NonlinearFactorGraph& Posegraph::graph() const { return g_graph; }
const gtsam::Values& Posegraph::initial() const { return g_values; }
const gtsam::Values& Posegraph::result() const { return g_values; }
