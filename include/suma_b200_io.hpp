// suma_b200_io.hpp -- header-only C++ for the data formats on either side of SurfelMapping::processScan
// (SURVEY.md 8f-4): the KITTI scan reader, calibration and pose files, and the odometry devkit's segment errors.
//
// Mirrors, with the same names and argument meaning:
//   io/KITTIReader.h:24-45        rv::KITTIReader        -> suma::KITTIReader      (fills suma::Laserscan)
//   util/kitti_utils.h:12-32      KITTICalibration       -> suma::KITTICalibration
//   util/kitti_utils.h:34-88      KITTI::Odometry::*     -> suma::KITTI::Odometry::* (loadPoses, trajectoryDistances,
//                                 lastFrameFromSegmentLength, rotationError, translationError, calcSequenceErrors,
//                                 saveSequenceErrors, saveStats) plus savePoses (VisualizerWindow.cpp:848-868)
// Differences: labels come from SemanticKITTI `.label` files next to the scans (`../labels/NNNNNN.label`, class id in
// the low 16 bits) instead of a RangeNet++ call inside read() (io/KITTIReader.cpp:172-203; the network is outside the hot
// path); the reader keeps no ring buffer of decoded scans (a visualizer convenience for stepping backwards).
// No GPU, no library call: this header only needs suma_b200.hpp for Matrix4 and Laserscan.
#ifndef SUMA_B200_IO_HPP_
#define SUMA_B200_IO_HPP_

#include <algorithm>
#include <cmath>
#include <filesystem>
#include <fstream>
#include <sstream>

#include "suma_b200.hpp"

namespace suma {

namespace detail {
inline std::vector<std::string> split_ws(const std::string& s) {
  std::vector<std::string> out;
  std::istringstream in(s);
  std::string tok;
  while (in >> tok) out.push_back(tok);
  return out;
}
inline std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
/** inverse of an affine 4x4 (last row 0 0 0 1): [A | t]^-1 = [A^-1 | -A^-1 t], A^-1 by cofactors */
template <typename T>
Matrix4<T> inverse_affine(const Matrix4<T>& M) {
  const T a = M(0, 0), b = M(0, 1), c = M(0, 2), d = M(1, 0), e = M(1, 1), f = M(1, 2), g = M(2, 0), h = M(2, 1), i = M(2, 2);
  const T det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  Matrix4<T> R;
  R(0, 0) = (e * i - f * h) / det; R(0, 1) = (c * h - b * i) / det; R(0, 2) = (b * f - c * e) / det;
  R(1, 0) = (f * g - d * i) / det; R(1, 1) = (a * i - c * g) / det; R(1, 2) = (c * d - a * f) / det;
  R(2, 0) = (d * h - e * g) / det; R(2, 1) = (b * g - a * h) / det; R(2, 2) = (a * e - b * d) / det;
  for (int r = 0; r < 3; ++r) R(r, 3) = -(R(r, 0) * M(0, 3) + R(r, 1) * M(1, 3) + R(r, 2) * M(2, 3));
  return R;
}
}  // namespace detail

/** util/kitti_utils.h:12-32 -- "name: 12 numbers" lines of calib.txt as 4x4 matrices */
class KITTICalibration {
 public:
  KITTICalibration() {}
  explicit KITTICalibration(const std::string& filename) { initialize(filename); }
  void initialize(const std::string& filename) {
    m_.clear();
    std::ifstream in(filename.c_str());
    if (!in.is_open()) throw std::runtime_error(std::string("Unable to open calibration file: ") + filename);
    std::string line;
    while (std::getline(in, line)) {
      size_t colon = line.find(':');
      if (colon == std::string::npos || line.find(':', colon + 1) != std::string::npos) continue;
      std::vector<std::string> entries = detail::split_ws(line.substr(colon + 1));
      if (entries.size() != 12) continue;
      Matrix4f m;
      for (int k = 0; k < 12; ++k) m(k / 4, k % 4) = std::stof(entries[k]);
      m_[detail::trim(line.substr(0, colon))] = m;
    }
  }
  void clear() { m_.clear(); }
  bool exists(const std::string& name) const { return m_.find(name) != m_.end(); }
  const Matrix4f& operator[](const std::string& name) const {
    auto it = m_.find(name);
    if (it == m_.end()) throw std::runtime_error("KITTICalibration: no matrix named " + name);
    return it->second;
  }

 private:
  std::map<std::string, Matrix4f> m_;
};

/** io/KITTIReader.h:24-45 -- all `.bin` files of the directory of `scan_filename` (or of the directory itself), sorted */
class KITTIReader {
 public:
  explicit KITTIReader(const std::string& scan_filename) { initScanFilenames(scan_filename); }
  void reset() { currentScan = 0; }
  bool isSeekable() const { return true; }
  uint32_t count() const { return static_cast<uint32_t>(scan_filenames.size()); }
  void seek(uint32_t scannr) {
    if (scannr >= scan_filenames.size()) throw std::runtime_error("KITTIReader::seek: scan out of range");
    currentScan = scannr;
  }
  /** next scan; false at the end of the sequence */
  bool read(Laserscan& scan) {
    if (currentScan >= scan_filenames.size()) return false;
    return read(currentScan++, scan);
  }
  /** io/KITTIReader.cpp:136-203: N = bytes / 16 records of (x, y, z, remission); points become (x, y, z, 1) */
  bool read(uint32_t scan_idx, Laserscan& scan) {
    if (scan_idx >= scan_filenames.size()) return false;
    std::ifstream in(scan_filenames[scan_idx].c_str(), std::ios::binary);
    if (!in.is_open()) return false;
    in.seekg(0, std::ios::end);
    const uint32_t num_points = static_cast<uint32_t>(static_cast<uint64_t>(in.tellg()) / (4 * sizeof(float)));
    in.seekg(0, std::ios::beg);
    std::vector<float> values(4 * static_cast<size_t>(num_points));
    if (num_points) in.read(reinterpret_cast<char*>(values.data()), static_cast<std::streamsize>(values.size() * sizeof(float)));
    scan.points.assign(4 * static_cast<size_t>(num_points), 1.0f);
    remissions.resize(num_points);
    float max_remission = 0.0f;
    for (uint32_t i = 0; i < num_points; ++i) {
      scan.points[4 * i + 0] = values[4 * i + 0];
      scan.points[4 * i + 1] = values[4 * i + 1];
      scan.points[4 * i + 2] = values[4 * i + 2];
      remissions[i] = values[4 * i + 3];
      max_remission = std::max(max_remission, remissions[i]);
    }
    if (max_remission > 0.0f)
      for (uint32_t i = 0; i < num_points; ++i) remissions[i] /= max_remission;
    scan.labels_float.clear();
    scan.labels_prob.clear();
    const std::filesystem::path bin(scan_filenames[scan_idx]);
    const std::filesystem::path lab = bin.parent_path().parent_path() / "labels" / (bin.stem().string() + ".label");
    std::ifstream lin(lab.string().c_str(), std::ios::binary);
    if (lin.is_open()) {
      std::vector<uint32_t> raw(num_points);
      lin.read(reinterpret_cast<char*>(raw.data()), static_cast<std::streamsize>(raw.size() * sizeof(uint32_t)));
      if (static_cast<size_t>(lin.gcount()) != raw.size() * sizeof(uint32_t))
        throw std::runtime_error("KITTIReader: label file does not match the scan: " + lab.string());
      scan.labels_float.resize(num_points);
      scan.labels_prob.assign(num_points, 1.0f);
      for (uint32_t i = 0; i < num_points; ++i) scan.labels_float[i] = static_cast<float>(raw[i] & 0xFFFFu);
    }
    return true;
  }
  /** remissions of the scan read last, scaled to a maximum of 1 (Laserscan::remissions_ in the reference) */
  std::vector<float> remissions;

 protected:
  void initScanFilenames(const std::string& scan_filename) {
    namespace fs = std::filesystem;
    scan_filenames.clear();
    fs::path p(scan_filename);
    fs::path dir = fs::is_directory(p) ? p : p.parent_path();
    if (fs::is_directory(dir / "velodyne")) dir = dir / "velodyne";
    if (!fs::is_directory(dir)) return;
    for (const auto& e : fs::directory_iterator(dir))
      if (e.path().extension() == ".bin") scan_filenames.push_back(e.path().string());
    std::sort(scan_filenames.begin(), scan_filenames.end());
  }
  uint32_t currentScan{0};
  std::vector<std::string> scan_filenames;
};

namespace KITTI {
namespace Odometry {

struct errors {  // util/kitti_utils.h:40-48
  int32_t first_frame;
  float r_err, t_err, len, speed;
  errors(int32_t first_frame, float r_err, float t_err, float len, float speed)
      : first_frame(first_frame), r_err(r_err), t_err(t_err), len(len), speed(speed) {}
};

static const float lengths[] = {100, 200, 300, 400, 500, 600, 700, 800};  // util/kitti_utils.cpp:75-76

inline std::vector<Matrix4f> loadPoses(const std::string& file_name) {
  std::vector<Matrix4f> poses;
  std::ifstream fp(file_name.c_str());
  std::string line;
  while (std::getline(fp, line)) {
    std::vector<std::string> entries = detail::split_ws(line);
    if (entries.size() < 12) continue;
    Matrix4f P;
    for (int i = 0; i < 12; ++i) P(i / 4, i % 4) = std::stof(entries[i]);
    poses.push_back(P);
  }
  return poses;
}

/** VisualizerWindow.cpp:848-868: one line per pose, rows of T_cam_velo * pose * T_velo_cam (identity = velodyne frame) */
inline void savePoses(const std::string& file_name, const std::vector<Matrix4d>& poses, const Matrix4f& T_cam_velo = Matrix4f()) {
  const Matrix4f T_velo_cam = detail::inverse_affine(T_cam_velo);
  std::ofstream out(file_name.c_str());
  out.precision(9);
  for (size_t i = 0; i < poses.size(); ++i) {
    Matrix4f pose = (T_cam_velo * poses[i].cast<float>()) * T_velo_cam;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) out << ((r == 0 && c == 0) ? "" : " ") << pose(r, c);
    out << std::endl;
  }
}

/** cumulative path length (float accumulation, as the devkit does) */
inline std::vector<float> trajectoryDistances(const std::vector<Matrix4f>& poses) {
  std::vector<float> travelled(poses.empty() ? 0 : 1, 0.0f);
  for (size_t k = 1; k < poses.size(); ++k) {
    float sq = 0.0f;
    for (int axis = 0; axis < 3; ++axis) {
      const float delta = poses[k - 1](axis, 3) - poses[k](axis, 3);
      sq += delta * delta;
    }
    travelled.push_back(travelled.back() + std::sqrt(sq));
  }
  return travelled;
}

/** first frame whose travelled distance exceeds that of `first_frame` by more than `len`; -1 if the run is too short */
inline int32_t lastFrameFromSegmentLength(const std::vector<float>& dist, int32_t first_frame, float len) {
  const float goal = dist[first_frame] + len;
  auto it = std::find_if(dist.begin() + first_frame, dist.end(), [goal](float d) { return d > goal; });
  return it == dist.end() ? -1 : static_cast<int32_t>(it - dist.begin());
}

/** angle of the rotation part: acos((trace - 1) / 2), clamped */
inline float rotationError(const Matrix4f& pose_error) {
  const float trace = pose_error(0, 0) + pose_error(1, 1) + pose_error(2, 2);
  const float cosine = std::min(1.0f, std::max(-1.0f, 0.5f * (trace - 1.0f)));
  return std::acos(cosine);
}

inline float translationError(const Matrix4f& pose_error) {
  return std::sqrt(pose_error(0, 3) * pose_error(0, 3) + pose_error(1, 3) * pose_error(1, 3) + pose_error(2, 3) * pose_error(2, 3));
}

/** util/kitti_utils.cpp:149-191: a start every 10 frames, every segment length the ground truth is long enough for;
 * errors are per metre of segment length */
inline std::vector<errors> calcSequenceErrors(const std::vector<Matrix4f>& poses_gt, const std::vector<Matrix4f>& poses_result) {
  if (poses_gt.size() != poses_result.size()) throw std::runtime_error("calcSequenceErrors: pose counts differ");
  const std::vector<float> dist = trajectoryDistances(poses_gt);
  std::vector<errors> out;
  for (size_t start = 0; start < poses_gt.size(); start += 10) {
    const Matrix4f gt_start_inv = detail::inverse_affine(poses_gt[start]);
    const Matrix4f res_start_inv = detail::inverse_affine(poses_result[start]);
    for (float len : lengths) {
      const int32_t stop = lastFrameFromSegmentLength(dist, static_cast<int32_t>(start), len);
      if (stop < 0) continue;
      const Matrix4f motion_gt = gt_start_inv * poses_gt[stop];
      const Matrix4f motion_res = res_start_inv * poses_result[stop];
      const Matrix4f mismatch = detail::inverse_affine(motion_res) * motion_gt;
      const float frames = static_cast<float>(stop - static_cast<int32_t>(start) + 1);
      out.emplace_back(static_cast<int32_t>(start), rotationError(mismatch) / len, translationError(mismatch) / len, len,
                       len / (0.1f * frames));
    }
  }
  return out;
}

inline void saveSequenceErrors(const std::vector<errors>& err, const std::string& file_name) {
  std::ofstream out(file_name.c_str());
  for (const errors& e : err) out << e.first_frame << " " << e.r_err << " " << e.t_err << " " << e.len << " " << e.speed << std::endl;
}

/** util/kitti_utils.cpp:479-498: mean translational and rotational error; written as "<t> <r>" to <dir>/stats.txt when a
 * directory is given */
inline std::pair<float, float> saveStats(const std::vector<errors>& err, const std::string& dir) {
  float sum_t = 0.0f, sum_r = 0.0f;
  for (const errors& e : err) {
    sum_t += e.t_err;
    sum_r += e.r_err;
  }
  const float n = static_cast<float>(err.size());
  const std::pair<float, float> mean(sum_t / n, sum_r / n);
  if (!dir.empty()) {
    std::ofstream out((dir + "/stats.txt").c_str());
    out << std::fixed << mean.first << " " << mean.second << "\n";
  }
  return mean;
}

}  // namespace Odometry
}  // namespace KITTI
}  // namespace suma

#endif  // SUMA_B200_IO_HPP_
