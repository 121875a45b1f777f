// orc_loop.cpp -- TEST INFRASTRUCTURE (CPU oracle): loop-closure twin. SurfelMapping::checkLoopClosure
// (core/SurfelMapping.cpp:527-795) over the oracle's CPU operators. The control flow is the product's header
// include/suma_b200_loop.hpp (host logic, written once over a backend concept); what differs -- and what the parity test
// compares -- is every operator underneath: render_inactive / render_composed, the Gauss-Newton runs against the old map
// frame and the Jacobian evaluations, here from orc_map.c / orc_core.c.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/suma_b200_loop.hpp"

extern "C" {
#include "orc.h"
}

namespace lp = suma_b200::loop;

struct orc_loop {
  orc_slam* slam;
  lp::State st;
  std::vector<float> mv, mn, ms;  // a model frame downloaded from the map
  double current_pose_old[16];
};

namespace {

struct OrcBackend {
  orc_loop* L;
  const orc_params* p;
  float *cv, *cn, *cs;
  float confidence_threshold() { return orc_slam_conf_threshold(L->slam); }
  void render_inactive(const float* pose, float ct) { orc_map_render_inactive(orc_slam_map(L->slam), pose, ct); }
  void render_composed(const float* po, const float* pn, float ct) { orc_map_render_composed(orc_slam_map(L->slam), po, pn, ct); }
  void fetch(int which) { orc_map_get_frame(orc_slam_map(L->slam), which, L->mv.data(), L->mn.data(), L->ms.data()); }
  int minimize_old(const double* T0, double* pose_out, double* out48) {
    fetch(0);
    return orc_icp_minimize(p, cv, cn, cs, L->mv.data(), L->mn.data(), L->ms.data(), T0, p->icp_max_distance,
                            p->icp_max_angle, pose_out, out48, nullptr, nullptr);
  }
  void evaluate(int model, const double* pose, int iteration, double* out48) {
    fetch(model);
    orc_icp_jacobian(p, cv, cn, cs, L->mv.data(), L->mn.data(), L->ms.data(), pose, iteration, p->icp_max_distance,
                     p->icp_max_angle, 0, p->data_height, out48, nullptr);
  }
};

void hook(orc_slam* s, void* user) {
  orc_loop* L = (orc_loop*)user;
  const orc_params* p = orc_slam_params(s);
  double stats[16];
  // result_new_ (SurfelMapping.cpp:417-423) from the statistics pass update_pose() just ran: stats hold F, inlier, outlier,
  // invalid, inlier_residual (orc_slam.c)
  orc_slam_get_stats(s, stats);
  lp::OptResult& rn = L->st.result_new;
  rn.error = stats[1];
  rn.outlier = (uint32_t)(float)stats[3];
  rn.inlier = (uint32_t)stats[2];
  rn.valid = rn.inlier + rn.outlier;
  rn.invalid = (uint32_t)(float)stats[4];
  rn.residual = rn.error / (rn.inlier + rn.outlier);
  rn.inlier_residual = (float)stats[5] / rn.inlier;
  lp::Poses P;
  memcpy(P.current.m, orc_slam_pose_member(s, 0), 128);
  memcpy(P.current_old.m, orc_slam_pose_member(s, 1), 128);
  memcpy(P.current_new.m, orc_slam_pose_member(s, 2), 128);
  memcpy(P.last_old.m, orc_slam_pose_member(s, 3), 128);
  memcpy(P.last_increment.m, orc_slam_pose_member(s, 6), 128);  // lastIncrement_ = this scan's increment (:473)
  lp::M4 inc;
  memcpy(inc.m, orc_slam_pose_member(s, 5), 128);
  const uint32_t t = orc_slam_timestamp(s);
  L->st.after_update_pose(t, inc, P.current);
  OrcBackend be{L, p, nullptr, nullptr, nullptr};
  orc_slam_current_frame(s, &be.cv, &be.cn, &be.cs);
  lp::check_loop_closure(L->st, be, t, P);
  memcpy(orc_slam_pose_member(s, 1), P.current_old.m, 128);
  memcpy(L->current_pose_old, P.current_old.m, 128);
}

}  // namespace

extern "C" {

orc_loop* orc_loop_attach(orc_slam* s, float search_distance, float min_trajectory_distance, int32_t min_verifications,
                          float residual_threshold, float outlier_threshold, float valid_threshold) {
  orc_loop* L = new orc_loop();
  L->slam = s;
  L->st.reset();
  L->st.p.search_distance = search_distance;
  L->st.p.min_trajectory_distance = min_trajectory_distance;
  L->st.p.min_verifications = min_verifications;
  L->st.p.residual_threshold = residual_threshold;
  L->st.p.outlier_threshold = outlier_threshold;
  L->st.p.valid_threshold = valid_threshold;
  const orc_params* p = orc_slam_params(s);
  size_t Pm = (size_t)p->model_width * p->model_height * 4;
  L->mv.assign(Pm, 0.f); L->mn.assign(Pm, 0.f); L->ms.assign(Pm, 0.f);
  for (int i = 0; i < 16; ++i) L->current_pose_old[i] = (i % 5 == 0) ? 1.0 : 0.0;
  orc_slam_set_hook(s, hook, L);
  return L;
}
void orc_loop_detach(orc_loop* L) {
  if (!L) return;
  orc_slam_set_hook(L->slam, nullptr, nullptr);
  delete L;
}
/* info: [0] loop_count [1] time_without_loop_closure [2] candidates_tested [3] loop_edges_added [4] unverified
 * [5] already_verified [6] found_candidate [7] use_candidate [8] optimisation_requested [9] last_added_candidate
 * [10] n_edges [11] n_poses ; ratios: valid, outlier, rel_error, residual_old, residual_new */
void orc_loop_info(const orc_loop* L, int64_t info[12], double ratios[5], double current_pose_old[16]) {
  const lp::State& S = L->st;
  info[0] = S.loop_count; info[1] = S.time_without_loop_closure; info[2] = S.candidates_tested;
  info[3] = S.loop_edges_added; info[4] = (int64_t)S.unverified.size(); info[5] = S.already_verified;
  info[6] = S.found_candidate; info[7] = S.use_candidate; info[8] = S.optimisation_requested;
  info[9] = S.last_added_candidate; info[10] = (int64_t)S.graph.edges.size(); info[11] = (int64_t)S.graph.poses.size();
  ratios[0] = S.stat_valid_ratio; ratios[1] = S.stat_outlier_ratio; ratios[2] = S.stat_rel_error;
  ratios[3] = S.result_old.residual; ratios[4] = S.result_new.residual;
  memcpy(current_pose_old, L->current_pose_old, 128);
}
/* SurfelMapping::integrateLoopClosures (SurfelMapping.cpp:212-258) at the start of a scan, with the optimiser replaced by
 * the identity (poses16 == NULL: the poses the graph held when the optimisation was requested) or by the caller's poses
 * (n x 16 doubles, column-major). Returns the number of poses handed to SurfelMap::updatePoses, 0 if nothing was pending. */
uint32_t orc_loop_integrate(orc_loop* L, const double* poses16, uint32_t n) {
  lp::State& S = L->st;
  if (!S.optimisation_requested) return 0;
  std::vector<lp::M4> opt;
  if (poses16) {
    opt.resize(n);
    for (uint32_t i = 0; i < n; ++i) memcpy(opt[i].m, poses16 + 16 * (size_t)i, 128);
  } else {
    opt.assign(S.graph.poses.begin(), S.graph.poses.begin() + (long)S.requested_size);
  }
  lp::M4 cur, cold, cnew;
  memcpy(cur.m, orc_slam_pose_member(L->slam, 0), 128);
  memcpy(cold.m, orc_slam_pose_member(L->slam, 1), 128);
  memcpy(cnew.m, orc_slam_pose_member(L->slam, 2), 128);
  std::vector<lp::M4> all = S.integrate(opt, cur, cold, cnew);
  if (all.empty()) return 0;
  memcpy(orc_slam_pose_member(L->slam, 0), cur.m, 128);
  memcpy(orc_slam_pose_member(L->slam, 1), cold.m, 128);
  memcpy(orc_slam_pose_member(L->slam, 2), cnew.m, 128);
  memcpy(L->current_pose_old, cold.m, 128);
  for (size_t t = 0; t < all.size(); ++t) {  // map_->updatePoses(casted_poses), SurfelMap.cpp:485-490
    float pf[16];
    lp::to_float(all[t], pf);
    orc_map_set_pose(orc_slam_map(L->slam), (uint32_t)t, pf);
  }
  return (uint32_t)all.size();
}

uint32_t orc_loop_edges(const orc_loop* L, int32_t* from_to, double* rel16, uint32_t cap) {
  const auto& E = L->st.graph.edges;
  uint32_t k = (uint32_t)E.size() < cap ? (uint32_t)E.size() : cap;
  for (uint32_t i = 0; i < k; ++i) {
    from_to[2 * i] = E[i].from; from_to[2 * i + 1] = E[i].to;
    memcpy(rel16 + 16 * i, E[i].rel.m, 128);
  }
  return (uint32_t)E.size();
}

}  // extern "C"
