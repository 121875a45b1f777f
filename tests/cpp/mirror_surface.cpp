// Compile-only check (tests/test_library_abi.py): every method of the reference's four classes that SURVEY.md 8b lists
// for the hot path exists on the mirror with the reference's argument meaning. Nothing here is executed.
#include <suma_b200.hpp>

struct FakeCalibration {};
struct FakeColorMap {};

void surface(suma::SurfelMapping& slam, suma::Laserscan& scan, const suma::ParameterList& params) {
  slam.setParameters(params);
  slam.setCalibration(FakeCalibration());
  slam.setColorMap(FakeColorMap());
  slam.processScan(scan);
  uint32_t t = slam.timestamp();
  suma::Matrix4d cur = slam.getCurrentPose(), last = slam.getLastPose();
  suma::Frame::Ptr f0 = slam.getCurrentFrame(), f1 = slam.getLastFrame(), f2 = slam.getCurrentModelFrame(),
                   f3 = slam.getLastModelFrame(), f4 = slam.getOldSurfelMap(), f5 = slam.getNewSurfelMap();
  suma::SurfelMapping::Stats st = slam.getStatistics();
  std::vector<suma::Matrix4d> poses = slam.getOptimizedPoses();
  // "close-loops": checkLoopClosure runs inside processScan; the graph edges go to the caller's optimiser (gtsam)
  sb_loop_params lp;
  sb_default_loop_params(&lp);
  slam.enableLoopClosure(true, &lp);
  bool lc = slam.foundLoopClosureCandidate() || slam.useLoopClosureCandidate() || slam.optimisationRequested();
  std::vector<sb_loop_edge> edges = slam.getLoopEdges();
  sb_loop_info li = slam.getLoopInfo();
  slam.setCurrentPose(cur);
  (void)edges; (void)li;
  std::shared_ptr<suma::SurfelMap> map = slam.getMap();
  suma::Matrix4f P = cur.cast<float>();
  map->update(P, *f0);
  map->render(P, *f2, 0.5f);
  map->render(P, P, *f2, 0.5f);
  map->render_active(P, 0.5f);
  map->render_inactive(P, 0.5f);
  map->render_composed(P, P, 0.5f);
  suma::Frame::Ptr m0 = map->oldMapFrame(), m1 = map->newMapFrame(), m2 = map->composedFrame();
  map->updatePoses(std::vector<suma::Matrix4f>(1, P));
  std::vector<suma::Surfel> all = map->getAllSurfels();
  uint32_t n = map->size();
  map->setColorMap(FakeColorMap());
  map->reset();

  suma::Preprocessing pre(slam.context());
  pre.setParameters(params);
  pre.setCalibration(FakeCalibration());
  pre.process(scan.points, *f0, scan.labels_float, scan.labels_prob, t);

  suma::Frame2Model objective(slam.context());
  objective.setData(f0, m1);
  objective.initialize(last);
  suma::Matrix6d JtJ;
  suma::Vector6d Jtf;
  double F = objective.jacobianProducts(JtJ, Jtf);
  objective.increment(Jtf);
  suma::Matrix4d pose = objective.pose();
  double counters = objective.inlier() + objective.outlier() + objective.valid() + objective.invalid() +
                    objective.inlier_residual() + objective.num_parameters();
  suma::LieGaussNewton gn(slam.context());
  gn.setParameters(params);
  int32_t rc = gn.minimize(objective, last);
  (void)st; (void)poses; (void)lc; (void)all; (void)n; (void)F; (void)pose; (void)counters; (void)rc; (void)m0; (void)m2;
  (void)f1; (void)f3; (void)f4; (void)f5;
  slam.reset();
}
