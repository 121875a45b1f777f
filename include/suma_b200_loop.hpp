// suma_b200_loop.hpp -- loop-closure detection and verification of SurfelMapping (host logic), restated from
// core/SurfelMapping.cpp:478-795 (checkLoopClosure, getCandidateIndexes, getClosestIndex) and :451-473 (the pose-graph
// bookkeeping of updatePose) of PRBonn/semantic_suma. The reference runs this between updatePose() and updateMap()
// (SurfelMapping.cpp:196-200); its building blocks on the device side are SurfelMap::render_inactive / render_composed
// and extra Gauss-Newton runs / Jacobian evaluations against the old and the composed map frames.
//
// The control flow is written ONCE, over a small backend concept, and instantiated twice: over the CUDA operators inside
// libsuma_b200 (csrc/sb_loop.cu) and over the CPU oracle's operators (oracle/orc_loop.cpp, test infrastructure), so that a
// parity test can compare every decision, counter and candidate of a sequence with a real loop bit for bit.
//
// Out of scope: the pose-graph OPTIMISATION itself (gtsam Levenberg-Marquardt, Posegraph.cpp:90-101).
// PoseGraphLite below only stores what the reference feeds into gtsam (initial poses, odometry and loop edges); a host
// application optimises those with gtsam and hands the result back through State::integrate() -- the body of
// SurfelMapping::integrateLoopClosures (:212-258).
//
// Backend concept (all matrices column-major 4x4):
//   float  confidence_threshold();                                            // getConfidenceThreshold()
//   void   render_inactive(const float pose[16], float ct);                   // map_->render_inactive
//   void   render_composed(const float pose_old[16], const float pose_new[16], float ct);
//   int    minimize_old(const double T0[16], double pose_out[16], double out48[48]);
//                                       // objective_->setData(currentFrame_, map_->oldMapFrame()); gn_->minimize(...);
//                                       // out48 = sums of the last evaluated iteration; returns the iteration count
//   void   evaluate(int model /*0 old map frame, 2 composed frame*/, const double pose[16], int iteration, double out48[48]);
//                                       // objective_->setData(currentFrame_, frame); initialize(pose); jacobianProducts()
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace suma_b200 {
namespace loop {

struct M4 {
  double m[16];
};
inline M4 identity() {
  M4 r;
  for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0 : 0.0;
  return r;
}
inline M4 mul(const M4& a, const M4& b) {  // same association as the library's mat4_mul (sbm::mat4_mul<double>)
  M4 c;
  for (int col = 0; col < 4; ++col)
    for (int r = 0; r < 4; ++r)
      c.m[col * 4 + r] = ((a.m[0 * 4 + r] * b.m[col * 4 + 0] + a.m[1 * 4 + r] * b.m[col * 4 + 1]) + a.m[2 * 4 + r] * b.m[col * 4 + 2]) +
                         a.m[3 * 4 + r] * b.m[col * 4 + 3];
  return c;
}
inline M4 rigid_inverse(const M4& a) {  // poses are rigid: [R^T | -R^T t] (the rule the whole library uses for inverse())
  M4 r;
  for (int c = 0; c < 3; ++c)
    for (int rr = 0; rr < 3; ++rr) r.m[c * 4 + rr] = a.m[rr * 4 + c];
  for (int rr = 0; rr < 3; ++rr) r.m[12 + rr] = -((a.m[rr * 4 + 0] * a.m[12] + a.m[rr * 4 + 1] * a.m[13]) + a.m[rr * 4 + 2] * a.m[14]);
  r.m[3] = r.m[7] = r.m[11] = 0.0;
  r.m[15] = 1.0;
  return r;
}
inline void to_float(const M4& a, float out[16]) {
  for (int i = 0; i < 16; ++i) out[i] = (float)a.m[i];
}
inline double translation_distance(const M4& a, const M4& b) {  // pose_distance(), SurfelMapping.cpp:500-502
  const double dx = a.m[12] - b.m[12], dy = a.m[13] - b.m[13], dz = a.m[14] - b.m[14], dw = a.m[15] - b.m[15];
  return std::sqrt(((dx * dx + dy * dy) + dz * dz) + dw * dw);
}
// SE3::log (lie_algebra.cpp:36-71), used only for the increment comparison at :556
inline void se3_log(const M4& T, double x[6]) {
  const double* M = T.m;
  for (int i = 0; i < 6; ++i) x[i] = 0.0;
  double d = 0.5 * (((M[0] + M[5]) + M[10]) - 1.0);
  double K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (d < 1 - 1e-10) {
    double theta = std::acos(d);
    double f = theta / (2 * std::sin(theta));
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) K[r * 3 + c] = f * (M[c * 4 + r] - M[r * 4 + c]);
    x[3] = K[2 * 3 + 1]; x[4] = K[0 * 3 + 2]; x[5] = K[1 * 3 + 0];
  }
  double theta = std::sqrt((x[3] * x[3] + x[4] * x[4]) + x[5] * x[5]);
  x[0] = M[12]; x[1] = M[13]; x[2] = M[14];
  if (std::fabs(theta) > 1e-10) {
    double half = 0.5 * theta;
    double beta = 1 / (theta * theta) * (1 - theta * std::cos(half) / (2 * std::sin(half)));
    double t[3] = {M[12], M[13], M[14]};
    for (int r = 0; r < 3; ++r) {
      double acc = 0;
      for (int c = 0; c < 3; ++c) {
        double k2 = (K[r * 3 + 0] * K[0 * 3 + c] + K[r * 3 + 1] * K[1 * 3 + c]) + K[r * 3 + 2] * K[2 * 3 + c];
        double Vi = ((r == c ? 1.0 : 0.0) + (-0.5) * K[r * 3 + c]) + beta * k2;
        acc += Vi * t[c];
      }
      x[r] = acc;
    }
  }
}

struct Edge {
  int32_t from, to;
  M4 rel;
};

// what the reference hands to gtsam (Posegraph.h): initial values and between-factors
struct PoseGraphLite {
  std::vector<M4> poses;
  std::vector<Edge> edges;
  void reset() {
    poses.assign(1, identity());
    edges.clear();
  }
  const M4& pose(int32_t j) const { return poses[(size_t)j]; }
  void set_initial(int32_t t, const M4& p) {
    if ((size_t)t >= poses.size()) poses.resize((size_t)t + 1, identity());
    poses[(size_t)t] = p;
  }
  void add_edge(int32_t from, int32_t to, const M4& rel) { edges.push_back(Edge{from, to, rel}); }
};

struct OptResult {  // SurfelMapping.h:111-122
  double error = 10000.0, residual = 10000.0, inlier_residual = 10000.0;
  uint32_t inlier = 0, outlier = 0, valid = 0, invalid = 0;
  float outlier_ratio = 1.0f;
  M4 pose = identity();
};

struct Params {  // SurfelMapping.h:223-235, config/default.xml:70-76
  float search_distance = 20.0f;         // loop-search-distance
  float min_trajectory_distance = 200.0f;  // loop-min-trajectory-distance
  int32_t min_verifications = 3;         // loop-min-verifications
  float residual_threshold = 1.05f;      // loop-residual-threshold
  float outlier_threshold = 1.1f;        // loop-outlier-threshold
  float valid_threshold = 0.9f;          // loop-valid-threshold
  float delta_timestamp = 100.0f;        // loopDetlaTimestamp_ (not configurable in the reference)
};

struct State {
  Params p;
  PoseGraphLite graph;
  std::vector<float> trajectory_distances;
  OptResult result_new, result_old;
  std::vector<Edge> unverified, verified;
  bool already_verified = false;
  uint32_t time_without_loop_closure = 0, loop_count = 0;
  bool found_candidate = false, use_candidate = false, optimisation_requested = false;
  int32_t last_added_candidate = -1;
  uint32_t candidates_tested = 0, loop_edges_added = 0;
  float stat_valid_ratio = 0, stat_outlier_ratio = 0, stat_rel_error = 0;

  void reset() {
    graph.reset();
    trajectory_distances.assign(1, 0.0f);
    result_new = result_old = OptResult();
    unverified.clear();
    verified.clear();
    already_verified = false;
    time_without_loop_closure = loop_count = 0;
    found_candidate = use_candidate = optimisation_requested = false;
    before_id = -1;
    before_loop_count = 0;
    requested_size = 0;
    last_added_candidate = -1;
    candidates_tested = loop_edges_added = 0;
  }

  // updatePose(), SurfelMapping.cpp:460-469: odometry edge and travelled distance of scan `timestamp`
  void after_update_pose(uint32_t timestamp, const M4& increment, const M4& current_pose) {
    float distance = 0;
    if (timestamp > 0) {
      graph.set_initial((int32_t)timestamp, mul(graph.pose((int32_t)timestamp - 1), increment));
      graph.add_edge((int32_t)timestamp - 1, (int32_t)timestamp, increment);
      distance = (float)translation_distance(graph.pose((int32_t)timestamp - 1), current_pose);
      distance += trajectory_distances[timestamp - 1];
    }
    if (trajectory_distances.size() <= timestamp) trajectory_distances.resize((size_t)timestamp + 1, 0.0f);
    trajectory_distances[timestamp] = distance;
  }

  // What optimizeAsync() (:818-825) records when the optimisation is requested: the scan, its pose and the loop count.
  int32_t before_id = -1;
  uint32_t before_loop_count = 0;
  M4 before_pose = identity();
  size_t requested_size = 0;  // poses in the graph at that moment = what the host application optimises

  void request_optimisation(uint32_t timestamp) {
    optimisation_requested = true;
    before_id = (int32_t)timestamp;
    before_loop_count = loop_count;
    before_pose = graph.pose((int32_t)timestamp);
    requested_size = graph.poses.size();
  }

  // integrateLoopClosures(), :212-258, for a host application that optimised the first `requested_size` poses of `graph`
  // elsewhere (gtsam). Later poses and the current pose are carried along by the correction of the pose the request was
  // made at. Returns every pose of the graph afterwards -- what the reference hands to SurfelMap::updatePoses (as float).
  std::vector<M4> integrate(const std::vector<M4>& poses_opt, M4& current, M4& current_old, M4& current_new) {
    std::vector<M4> all;
    if (!optimisation_requested || before_id < 0 || (size_t)before_id >= poses_opt.size()) return all;
    const std::vector<M4> before = graph.poses;  // posegraph_->poses()
    for (size_t i = 0; i < poses_opt.size(); ++i) {
      all.push_back(poses_opt[i]);
      graph.set_initial((int32_t)i, poses_opt[i]);
    }
    loop_count -= before_loop_count;                                                    // :223
    const M4 difference = mul(poses_opt[(size_t)before_id], rigid_inverse(before_pose)); // :224
    for (size_t i = poses_opt.size(); i < before.size(); ++i) {
      const M4 moved = mul(difference, before[i]);
      all.push_back(moved);
      graph.set_initial((int32_t)i, moved);
    }
    optimisation_requested = false;           // currentlyOptimizing_ = false
    current = mul(difference, current);       // :236
    current_new = current_old = current;      // :240
    trajectory_distances.resize(graph.poses.size());
    M4 last = graph.pose(0);
    float distance = 0;
    for (size_t t = 0; t < graph.poses.size(); ++t) {  // :242-250 (float += double)
      distance = (float)((double)distance + translation_distance(last, graph.poses[t]));
      trajectory_distances[t] = distance;
      last = graph.poses[t];
    }
    before_id = -1;
    return all;
  }
};

// getCandidateIndexes / getClosestIndex, SurfelMapping.cpp:478-521 (both search around currentPose_)
inline int32_t closest_index(const State& s, uint32_t timestamp, const M4& current_pose, float radius) {
  int32_t closest = -1;
  float min_distance = radius;
  for (int32_t j = (int32_t)timestamp - (int32_t)s.p.delta_timestamp; j >= 0; --j) {
    float distance = (float)translation_distance(current_pose, s.graph.pose(j));
    float tdistance = s.trajectory_distances[timestamp] - s.trajectory_distances[(size_t)j];
    if (distance < min_distance && tdistance > s.p.min_trajectory_distance) {
      closest = j;
      min_distance = distance;
    }
  }
  return closest;
}

struct Counters {
  uint32_t valid, invalid, inlier, outlier;
  double error, inlier_residual;
};
inline Counters counters_of(const double out48[48]) {  // Frame2Model.cpp:214-227
  Counters c;
  c.valid = (uint32_t)out48[42];
  c.error = out48[43];
  c.outlier = (uint32_t)(float)out48[44];
  c.inlier_residual = (float)out48[45];
  c.inlier = c.valid - c.outlier;
  c.invalid = (uint32_t)(float)out48[46];
  return c;
}

// The poses the reference keeps as members and this function reads / writes.
struct Poses {
  M4 current, current_old, current_new, last_old, last_increment;
};

// SurfelMapping::checkLoopClosure, SurfelMapping.cpp:527-795
template <class Backend>
void check_loop_closure(State& s, Backend& be, uint32_t timestamp, Poses& P) {
  s.found_candidate = false;
  s.result_old = OptResult();
  s.use_candidate = false;
  bool candidate_added = false;
  bool have_min_candidate = false;
  const OptResult& rn = s.result_new;
  const float outlier_ratio_new = rn.outlier / float(rn.outlier + rn.inlier);
  const float valid_ratio_new = float(rn.valid) / float(rn.invalid + rn.valid);
  s.time_without_loop_closure += 1;
  double out48[48];

  // 1. verify the pending loop closure with the old map rendered at the chained old pose (:551-626)
  if (!s.unverified.empty() || s.already_verified) {
    float pose_old_f[16], pose_new_f[16];
    to_float(P.last_old, pose_old_f);
    be.render_inactive(pose_old_f, be.confidence_threshold());
    M4 increment_old;
    be.minimize_old(P.last_increment.m, increment_old.m, out48);
    Counters c = counters_of(out48);
    const float valid_ratio = float(c.valid) / float(c.valid + c.invalid);
    const float outlier_ratio = float(c.outlier) / float(c.outlier + c.inlier);
    double la[6], lb[6];
    se3_log(P.last_increment, la);
    se3_log(increment_old, lb);
    double n2 = 0;
    for (int i = 0; i < 6; ++i) n2 += (la[i] - lb[i]) * (la[i] - lb[i]);
    const float increment_difference = (float)std::sqrt(n2);
    if (valid_ratio > 0.2 && outlier_ratio < 0.85 && increment_difference < 0.1) {
      const M4 chained = mul(P.last_old, increment_old);
      to_float(chained, pose_old_f);
      to_float(P.current_new, pose_new_f);
      be.render_composed(pose_old_f, pose_new_f, be.confidence_threshold());
      const M4 I = identity();
      be.evaluate(2, I.m, 0, out48);
      c = counters_of(out48);
      const float error = (float)c.error;
      const float residual = error / float(c.inlier + c.outlier);
      s.result_old.error = error;
      s.result_old.inlier = c.inlier;
      s.result_old.outlier = c.outlier;
      s.result_old.residual = s.result_old.error / float(c.inlier + c.outlier);
      s.result_old.inlier_residual = c.inlier_residual / c.inlier;
      s.result_old.valid = c.valid;
      s.result_old.invalid = c.invalid;
      const float rel_error_all = (float)(residual / rn.residual);
      s.found_candidate = true;
      P.current_old = chained;
      const bool loop_closure = (rel_error_all < s.p.residual_threshold) || (residual - rn.residual) < 0.1;
      if (loop_closure) {
        s.time_without_loop_closure = 0;
        const int32_t index = closest_index(s, timestamp, P.current, s.p.search_distance);  // :595 (searches around currentPose_)
        if (index > -1) {
          Edge cand;
          cand.from = (int32_t)timestamp;
          cand.to = index;
          s.last_added_candidate = index;
          cand.rel = mul(rigid_inverse(P.current_old), s.graph.pose(index));
          (s.already_verified ? s.verified : s.unverified).push_back(cand);
        }
        s.use_candidate = true;
      }
    }
  }
  // enough verified loop closures (:629-634)
  if (!s.already_verified && (int32_t)s.unverified.size() >= s.p.min_verifications + 1) {
    for (const Edge& e : s.unverified) s.verified.push_back(e);
    s.unverified.clear();
    s.already_verified = true;
  }
  // 2. add the verified loop closures to the graph (:636-656)
  int32_t last_from = -1;
  for (const Edge& e : s.verified) {
    if (last_from != e.from) {
      last_from = e.from;
      s.loop_count += 1;
    }
    s.graph.add_edge(e.from, e.to, e.rel);
    s.loop_edges_added += 1;
  }
  s.verified.clear();
  // :658-663 -- the reference clones the graph and optimises it asynchronously with gtsam; here: a flag for the host
  if (!s.optimisation_requested && (s.loop_count > 6 || (s.loop_count > 0 && s.time_without_loop_closure > 3)))
    s.request_optimisation(timestamp);

  // 3. search a new candidate (:665-776)
  if (s.time_without_loop_closure > 3) {
    s.unverified.clear();
    s.use_candidate = false;
    s.already_verified = false;
    const int32_t to = closest_index(s, timestamp, P.current, s.p.search_distance);
    int32_t loop_timestamp = -1;
    if (to > -1) {
      s.candidates_tested += 1;
      const M4 pose_prior = s.graph.pose(to);
      float prior_f[16], pose_f[16], new_f[16];
      to_float(pose_prior, prior_f);
      be.render_inactive(prior_f, be.confidence_threshold());
      M4 O = mul(rigid_inverse(pose_prior), P.current);
      O.m[14] = 0.0;                       // O(2,3) = 0
      M4 inits[3];
      inits[0] = O;
      inits[1] = O;                        // R(O): rotation only
      inits[1].m[12] = inits[1].m[13] = inits[1].m[14] = 0.0;
      inits[2] = O;
      inits[2].m[12] = 0.5 * O.m[12];
      inits[2].m[13] = 0.5 * O.m[13];
      for (int i = 0; i < 3; ++i) {
        M4 gn_pose;
        const int iterations = be.minimize_old(inits[i].m, gn_pose.m, out48);
        be.evaluate(0, gn_pose.m, iterations, out48);  // objective_->jacobianProducts(JtJ, Jtr) at the final pose (:707)
        Counters c = counters_of(out48);
        const float valid_ratio = float(c.valid) / float(c.valid + c.invalid);
        const float outlier_ratio = float(c.outlier) / float(c.outlier + c.inlier);
        s.found_candidate = true;
        if (valid_ratio > 0.2 && outlier_ratio < 0.85) {
          const M4 pose = mul(pose_prior, gn_pose);
          to_float(pose, pose_f);
          to_float(P.current_new, new_f);
          be.render_composed(pose_f, new_f, be.confidence_threshold());
          const M4 I = identity();
          be.evaluate(2, I.m, 0, out48);
          c = counters_of(out48);
          const float error = (float)c.error;
          const float residual = error / float(c.inlier + c.outlier);
          const float outlier_ratio_old = float(c.outlier) / float(c.outlier + c.inlier);
          const float valid_ratio_old = c.valid / float(c.valid + c.invalid);
          const float rel_error_all = (float)(residual / rn.residual);
          const float rel_valid_ratio = valid_ratio_old / valid_ratio_new;
          const float rel_outlier_ratio = outlier_ratio_old / outlier_ratio_new;
          if (!candidate_added || (residual < s.result_old.residual && outlier_ratio_old < s.result_old.outlier_ratio)) {
            if (rel_valid_ratio >= s.p.valid_threshold && rel_outlier_ratio < s.p.outlier_threshold) {
              candidate_added = true;
              loop_timestamp = to;
              have_min_candidate = true;
              s.result_old.pose = pose;
              s.result_old.error = error;
              s.result_old.inlier = c.inlier;
              s.result_old.outlier = c.outlier;
              s.result_old.outlier_ratio = outlier_ratio_old;
              s.result_old.valid = c.valid;
              s.result_old.residual = s.result_old.error / c.valid;
              s.result_old.inlier_residual = c.inlier_residual / c.inlier;
              s.result_old.invalid = c.invalid;
              const bool loop_closure = (rel_error_all < s.p.residual_threshold) || (residual - rn.residual) < 0.1;
              if (loop_closure) P.current_old = pose;
            }
          }
        }
      }
    }
    if (have_min_candidate) {
      Edge cand;
      cand.from = (int32_t)timestamp;
      cand.to = loop_timestamp;
      cand.rel = mul(rigid_inverse(P.current_old), s.graph.pose(loop_timestamp));
      s.last_added_candidate = loop_timestamp;
      s.unverified.push_back(cand);
    }
  }
  const float valid_ratio_old = float(s.result_old.valid) / float(s.result_old.valid + s.result_old.invalid);
  const float outlier_ratio_old = float(s.result_old.outlier) / float(s.result_old.outlier + s.result_old.inlier);
  s.stat_valid_ratio = valid_ratio_old / valid_ratio_new;
  s.stat_outlier_ratio = outlier_ratio_old / outlier_ratio_new;
  s.stat_rel_error = (float)(s.result_old.residual / rn.residual);
}

}  // namespace loop
}  // namespace suma_b200
