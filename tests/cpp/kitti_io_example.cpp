// Reads a KITTI-style sequence directory with the C++ reader and evaluates an estimated pose file against poses.txt:
// prints one line of numbers that tests/test_kitti_io.py compares with the Python implementation.
//   kitti_io_example <seq_dir> <estimated_poses.txt>
#include <cstdio>
#include <suma_b200_io.hpp>

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string seq = argv[1];
  suma::KITTIReader reader(seq);
  suma::Laserscan scan;
  double coord_sum = 0.0, label_sum = 0.0, rem_sum = 0.0;
  uint32_t n_scans = 0, n_points = 0;
  while (reader.read(scan)) {
    ++n_scans;
    n_points += scan.size();
    for (uint32_t i = 0; i < scan.size(); ++i) coord_sum += scan.points[4 * i] + scan.points[4 * i + 1] + scan.points[4 * i + 2] + scan.points[4 * i + 3];
    for (float l : scan.labels_float) label_sum += l;
    for (float r : reader.remissions) rem_sum += r;
  }
  reader.seek(1);
  reader.read(scan);
  const uint32_t second = scan.size();
  suma::KITTICalibration calib(seq + "/calib.txt");
  auto gt = suma::KITTI::Odometry::loadPoses(seq + "/poses.txt");
  auto est = suma::KITTI::Odometry::loadPoses(argv[2]);
  auto err = suma::KITTI::Odometry::calcSequenceErrors(gt, est);
  auto stats = suma::KITTI::Odometry::saveStats(err, "");
  std::vector<suma::Matrix4d> back;
  for (const auto& P : est) back.push_back((suma::detail::inverse_affine(calib["Tr"]) * P * calib["Tr"]).cast<double>());
  suma::KITTI::Odometry::savePoses(std::string(argv[2]) + ".roundtrip", back, calib["Tr"]);
  std::printf("%u %u %u %.6f %.1f %.3f %d %zu %.9g %.9g %.6f\n", reader.count(), n_scans, n_points, coord_sum, label_sum, rem_sum,
              calib.exists("Tr") && !calib.exists("P9") ? 1 : 0, err.size(), stats.first, stats.second, (double)second);
  return 0;
}
