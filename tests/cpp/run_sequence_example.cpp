// The drop-in flow end to end in C++, as a maintainer would write it against the reference's class names:
// KITTIReader -> SurfelMapping::processScan -> getOptimizedPoses -> savePoses (+ devkit errors if poses.txt exists).
//   run_sequence_example <sequence dir> <out poses.txt> [width]
// Compiled and linked by the CPU test suite (tests/test_library_abi.py); running it needs a B200.
#include <cstdio>
#include <cstdlib>
#include <suma_b200_io.hpp>

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s <sequence dir> <out poses.txt> [width]\n", argv[0]);
    return 2;
  }
  const std::string seq = argv[1];
  suma::ParameterList params;
  if (argc > 3) {
    params.insert("data_width", std::atof(argv[3]));
    params.insert("model_width", std::atof(argv[3]));
  }
  try {
    suma::KITTIReader reader(seq);
    suma::SurfelMapping fusion(params);  // throws std::runtime_error without a CUDA device: there is no CPU fallback
    suma::KITTICalibration calib;
    std::filesystem::path calib_file = std::filesystem::path(seq) / "calib.txt";
    if (std::filesystem::exists(calib_file)) calib.initialize(calib_file.string());
    fusion.setCalibration(calib);
    suma::Laserscan scan;
    while (reader.read(scan)) {
      fusion.processScan(scan);
      suma::SurfelMapping::Stats st = fusion.getStatistics();
      std::printf("scan %u: %u surfels, %d iterations, residual %.6f\n", fusion.timestamp() - 1, fusion.getMap()->size(),
                  (int)st["num_iterations"], st["residual"]);
    }
    const suma::Matrix4f Tr = calib.exists("Tr") ? calib["Tr"] : suma::Matrix4f();
    suma::KITTI::Odometry::savePoses(argv[2], fusion.getOptimizedPoses(), Tr);
    std::filesystem::path gt_file = std::filesystem::path(seq) / "poses.txt";
    if (std::filesystem::exists(gt_file)) {
      auto gt = suma::KITTI::Odometry::loadPoses(gt_file.string());
      auto est = suma::KITTI::Odometry::loadPoses(argv[2]);
      gt.resize(std::min(gt.size(), est.size()));
      est.resize(gt.size());
      auto err = suma::KITTI::Odometry::calcSequenceErrors(gt, est);
      if (!err.empty()) {
        auto mean = suma::KITTI::Odometry::saveStats(err, "");
        std::printf("odometry: %.4f %% translational, %.6f deg/m rotational over %zu segments\n", 100.0 * mean.first,
                    mean.second * 57.29577951308232, err.size());
      }
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
