// A caller written against the reference's class names (SurfelMapping / Frame2Model / LieGaussNewton / SurfelMap),
// compiled against include/suma_b200.hpp. Reads raw scans (uint32 n, then n*4 floats), processes them and prints the
// final pose, the surfel count and the pose of an explicit frame-to-frame minimisation.
#include <cstdio>
#include <vector>

#include "suma_b200.hpp"

using namespace suma;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  int width = std::atoi(argv[2]);
  ParameterList params;
  params.insert("data_width", width);
  params.insert("model_width", width);
  params.insert("max iterations", 10);
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  std::vector<Laserscan> scans;
  for (;;) {
    uint32_t n = 0;
    if (std::fread(&n, 4, 1, f) != 1) break;
    Laserscan s;
    s.points.resize(static_cast<size_t>(n) * 4);
    if (std::fread(s.points.data(), 4, s.points.size(), f) != s.points.size()) return 4;
    scans.push_back(s);
  }
  std::fclose(f);
  SurfelMapping fusion(params);
  for (const Laserscan& s : scans) fusion.processScan(s);
  Matrix4d pose = fusion.getCurrentPose();
  std::printf("pose");
  for (int i = 0; i < 16; ++i) std::printf(" %.17g", pose.m[i]);
  std::printf("\nsurfels %u timestamp %u\n", fusion.getMap()->size(), fusion.timestamp());
  // explicit operator-level use: frame-to-frame ICP between the last two scans
  ContextPtr ctx = fusion.context();
  Preprocessing pre(ctx);
  auto a = std::make_shared<Frame>(ctx, width, 64), b = std::make_shared<Frame>(ctx, width, 64);
  std::vector<float> none;
  pre.process(scans[scans.size() - 2].points, *a, none, none, 100);
  pre.process(scans[scans.size() - 1].points, *b, none, none, 100);
  Frame2Model objective(ctx);
  objective.setData(b, a);
  LieGaussNewton gn(ctx);
  gn.minimize(objective, Matrix4d::Identity());
  std::printf("icp %u", gn.iterationCount());
  for (int i = 0; i < 16; ++i) std::printf(" %.17g", gn.pose().m[i]);
  std::printf("\ninlier %u outlier %u\n", objective.inlier(), objective.outlier());
  return 0;
}
