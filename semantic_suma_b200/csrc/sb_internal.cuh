// sb_internal.cuh -- internal declarations shared by the kernels and the C-ABI layer of libsuma_b200.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/suma_b200.h"
#include "sb_math.cuh"

namespace sb {

constexpr uint32_t kMaxSurfels = 2048u * 2048u;  // SurfelMap.h:87
constexpr uint32_t kMaxPoses = 10000u;           // SurfelMap.h:205
constexpr uint32_t kExtractCap = 500000u;        // SurfelMap.cpp:279
constexpr uint32_t kComposeAge = 100u;           // SurfelMap.h:144
constexpr int kMaxGnIter = 256;
constexpr uint32_t kTilePoolCap = 16u * 1024u * 1024u;  // surfels held by the HBM tile cache (1 GiB)
constexpr uint32_t kMaxTileSlots = 16384u;
constexpr int kGroupCounts = 256;  // group totals of the update pass: kMaxSurfels / 256 threads / 64 blocks per group = 256

// everything the kernels need from sb_params plus the derived constants the reference computes on the host
struct KParams {
  int W, H, Wm, Hm;
  float fov_up, fov, min_depth, max_depth;          // data image (abs values, Preprocessing.cpp:94-95)
  float m_fov_up, m_fov, m_min_depth, m_max_depth;  // model image (SurfelMap.cpp:440-443)
  int weighting, bilinear;
  float factor;
  int use_stability, unstable_age, confidence_mode, active_timestamps;
  int weighting_scheme, averaging_scheme, update_always;
  float confidence_threshold, p_stable, p_unstable, log_prior, log_unstable;
  float sigma_angle, sigma_distance, max_weight;
  float pixel_size, min_radius, max_radius, radconf_angle_thresh, update_angle_thresh, map_max_distance;
  float max_loop_closure_distance;
  int label_offset_quirk;
};

// Frame.h:21-79: three RGBA32F images; one allocation, SoA of images
struct FrameDev {
  int W, H;
  float4* vertex;
  float4* normal;
  float4* semantic;
};

// Surfel store: structure of arrays of 16-byte lanes so that every pass streams exactly the lanes it needs with
// fully coalesced 128-bit accesses (the reference keeps 64-byte AoS records, Surfel.h:5-15):
//   p0 = (x, y, z, radius)   p1 = (nx, ny, nz, confidence)
//   p2 = (timestamp:u32, color, weight, count)   p3 = (r, g, b, w)  semantic label/255 x3 + probability
struct SurfelPlanes {
  float4* p0;
  float4* p1;
  float4* p2;
  float4* p3;
};

// ---- device state of one fused Gauss-Newton run (LieGaussNewton.cpp:13-79) ----
struct GnState {
  double pose[16];       // Tk_
  double last_error;     // last_error
  double out48[48];      // sums of the last evaluated iteration
  int k;                 // k_
  int done;              // loop left (converged or max iterations)
  int history_len;
  int error;             // a bounded spin of the persistent kernel timed out (missing block / peer GPU)
  double history[(kMaxGnIter + 1) * 16];
};

struct IcpArgs {
  const float4* data_v; const float4* data_n; const float4* data_s;
  const float4* model_v; const float4* model_n; const float4* model_s;
  float distance_thresh, angle_thresh;  // Frame2Model.cpp:66-67
  int row_begin, row_end;
  int has_semantics;
  int row_step;  // row permutation stride of the work distribution, coprime to (row_end - row_begin)
  int Wm, Hm;    // size of the model images = textureSize(vertex_model) (Frame2Model_jacobians.geom:70): the rendered
                 // model frame, or the last data frame in the frame-to-frame fallback (SurfelMapping.cpp:442)
};

// peer mailboxes for the multi-GPU one-shot all-reduce (fused into the Jacobian kernel's last block)
struct CommDev {
  int rank, nranks;
  long long* mailbox[8];  // mailbox[r]: rank r's mailbox (peer-mapped): [epoch parity 2][source rank 8][40] int64
  unsigned int* epoch;    // local all-reduce counter (identical on every rank by construction)
};

struct Mat4 {
  float m[16];
};
struct Mat4d {
  double m[16];
};

// Device-resident pose bookkeeping of SurfelMapping::updatePose (SurfelMapping.cpp:372-476): the per-scan pipeline is
// enqueued without a host round trip; the host reads this block back once, at the end of the scan.
struct PoseDev {
  double currentPose[16], lastPose[16], lastIncrement[16], increment[16], T0[16];
  Mat4 P_active, invP_active;  // currentPose_new * increment (render_active), and its inverse
  Mat4 P_cur, invP_cur;        // the scan's final pose (float) and inverse: map update + model rendering
  int fallback;                // track loss detected: the frame-to-frame minimisation runs
  int trackLoss;               // counter
  int icp_iterations;          // k_ of the frame-to-model minimisation
  int gn_error;                // the persistent Gauss-Newton kernel gave up (timeout): the scan's pose is invalid
  float t_err, r_err;
  int pad2[2];
};

// ---------------- kernel ids (launch accounting + the optional CUDA-event profiler) ----------------
enum KernelId {
  K_FILL = 0, K_PROJECT_SCATTER, K_PREPROCESS_TILE, K_ICP_JACOBIAN, K_GN_INIT,
  K_ICP_FUSED, K_POSE_PRODUCTS, K_RENDER_SCATTER, K_RENDER_RESOLVE, K_INDEX_SCATTER, K_RADIUS, K_UPDATE_SURFELS,
  K_GEN_SURFELS, K_EXTRACT_FLAGS, K_SCAN_BLOCKS, K_COMPACT_SCATTER, K_AOS_TO_SOA, K_SOA_TO_AOS, K_ICP_POST, K_COUNT
};
const char* kernel_name(int id);

// per-kernel device time with CUDA events on the launching stream (off by default: zero overhead)
struct Profiler {
  bool on = false;
  struct Rec { int id; cudaEvent_t e0, e1; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
};

// ---------------- launchers (implemented in the .cu files) ----------------
struct Launch {
  cudaStream_t stream;
  uint64_t* counter;  // number of kernels launched
  Profiler* prof;
};

struct ScopedKernel {
  const Launch& L;
  int slot = -1;
  ScopedKernel(const Launch& l, int id) : L(l) {
    ++*L.counter;
    if (L.prof && L.prof->on) {
      Profiler::Rec r{id, L.prof->get(), L.prof->get()};
      cudaEventRecord(r.e0, L.stream);
      L.prof->recs.push_back(r);
      slot = (int)L.prof->recs.size() - 1;
    }
  }
  ~ScopedKernel() {
    if (slot >= 0) cudaEventRecord(L.prof->recs[slot].e1, L.stream);
  }
};

// sb_preprocess.cu
// the two alternating z-buffer key images of the scan projection ([H][W+6] each, see sb_preprocess.cu) and their TMA maps
struct PrepKeys {
  unsigned long long* img[2] = {nullptr, nullptr};
  int cur = 0;
  bool use_tma = false;
  bool tmap_ok[2] = {false, false};
  CUtensorMap tmap[2];
};
size_t preprocess_key_elems(int W, int H);
bool make_key_tensor_map(CUtensorMap* out, const unsigned long long* keys, int W, int H);
// K1-K3 (+ optionally K6b radius map and the re-arming of the index-map keys / integrated flags for the map update)
void launch_preprocess(const Launch& L, const KParams& kp, const float4* pts, const float* labels, const float* probs,
                       uint32_t n, uint32_t timestamp, PrepKeys& keys, FrameDev out, float4* radius_map,
                       unsigned long long* index_keys, uint8_t* integrated, uint32_t* group_counts);

// sb_icp.cu
void launch_icp_jacobian(const Launch& L, const KParams& kp, const IcpArgs& a, const Mat4& pose, int iteration,
                         long long* acc32, long long* slots, unsigned int* ticket, int blocks);
// ---- persistent Gauss-Newton kernel (sb_icp.cu) ----
enum GnMode { GN_PLAIN = 0, GN_MAIN = 1, GN_POST = 2 };
struct GnJob {
  int mode;
  IcpArgs a;    // GN_PLAIN / GN_MAIN: the objective; GN_POST: the statistics pass at identity (all rows)
  IcpArgs fb;   // GN_POST: the frame-to-frame recovery objective (row stripe in multi-GPU mode)
  GnState* st;  // state of the run (GN_PLAIN: prepared by k_gn_init)
  PoseDev* pd;  // GN_MAIN / GN_POST
  long long* stats32;                       // GN_POST: raw sums of the statistics pass, read by the host after the scan
  float* poses; float* poses_inv; uint32_t t_map;  // GN_POST: pose-table entry of this scan
  float* Mtab; uint32_t n_poses;            // GN_MAIN / GN_POST: table inv(P) * poses[t] filled at the end (or null)
  int initialize_identity, fallback_mode;
  uint32_t timestamp;
  int max_iter;
  double eps, delta;
  unsigned int epoch_base;
  int use_cache;  // keep the thread's data pixels in shared memory across iterations
};
void launch_gn_init(const Launch& L, GnState* st, const Mat4d& T0, long long* acc32);
int launch_gn_persistent(const Launch& L, const KParams& kp, const GnJob& job, long long* slots, unsigned int* ticket,
                         unsigned long long* pub, const CommDev* comm, int blocks);
int gn_persistent_max_blocks(int sm_count);
// one-thread bookkeeping kernels: first scan of a sequence and the host-callback exchange
void launch_pose_after_icp(const Launch& L, const GnState* gn, PoseDev* pd, const Mat4d& T0, uint32_t timestamp,
                           int fallback_mode);
void launch_pose_finalize(const Launch& L, const GnState* gn, PoseDev* pd, int had_icp, float* poses, float* poses_inv,
                          uint32_t t);
int icp_grid_blocks(int sm_count);
int icp_balanced_blocks(int pixels, int max_blocks);

// sb_map.cu
struct RenderTargets {
  unsigned long long* key_old;   // may be null
  unsigned long long* key_new;   // may be null
  unsigned long long* key_comp;  // may be null
};
void launch_pose_products(const Launch& L, const Mat4& A, const Mat4* A_dev, const float* poses, float* out,
                          uint32_t count);
void launch_render_scatter(const Launch& L, const KParams& kp, SurfelPlanes s, const uint32_t* n_dev, uint32_t n_upper,
                           const float* M, float conf_thr, int t_thr, int emit_old, int emit_new, int lequal,
                           RenderTargets t);
void launch_render_resolve(const Launch& L, const KParams& kp, SurfelPlanes s, const float* M_old, const float* M_new,
                           RenderTargets t, FrameDev f_old, FrameDev f_new, FrameDev f_comp, FrameDev f_out,
                           FrameDev f_copy, int keep_semantic, int lequal);
void launch_index_scatter(const Launch& L, const KParams& kp, SurfelPlanes s, const uint32_t* n_dev, uint32_t n_upper,
                          const float* M, unsigned long long* keys);
void launch_radius(const Launch& L, const KParams& kp, FrameDev frame, float4* radius_map, unsigned long long* index_keys,
                   uint8_t* integrated, uint32_t* group_counts);
void launch_compact_update(const Launch& L, SurfelPlanes src, const uint8_t* keep, const uint32_t* block_counts,
                           const uint32_t* group_counts, const uint32_t* n_dev, uint32_t n_upper, SurfelPlanes dst,
                           uint32_t cap, uint32_t* count_out, uint32_t* kept_out);
void launch_update_surfels(const Launch& L, const KParams& kp, SurfelPlanes src, SurfelPlanes tmp, const uint32_t* n_dev,
                           uint32_t n_upper, const Mat4& pose, const Mat4& inv_pose, const Mat4* pose_dev,
                           const Mat4* inv_pose_dev, const float* poses,
                           const float* poses_inv, const unsigned long long* index_keys, const float4* radius_map,
                           FrameDev frame, int timestamp, float2 submap_center, float submap_extent,
                           uint8_t* integrated, uint8_t* keep, uint32_t* block_counts, uint32_t* group_counts);
// single-pass update / generate with in-place ordered compaction (decoupled look-back)
void launch_gen_compact(const Launch& L, const KParams& kp, FrameDev frame, const float4* radius_map,
                        const uint8_t* integrated, const float* poses, int timestamp, float2 submap_center,
                        float submap_extent, SurfelPlanes map, unsigned long long* desc, uint32_t* ticket, uint32_t gen,
                        uint32_t cap, uint32_t* counts);
// ordered compaction: items flagged in keep[0..n) of src go to dst[base..) in order; *count_out = base + #kept
void launch_extract_to_pool(const Launch& L, SurfelPlanes map, const uint32_t* n_dev, uint32_t n_upper, const float* poses,
                            float2 center, float extent, uint8_t* keep, uint32_t* block_counts, uint32_t* block_offsets,
                            SurfelPlanes pool, uint32_t pool_cap, uint32_t* pool_top, uint2* rec, uint32_t tile_cap);
void launch_extract_flags(const Launch& L, SurfelPlanes s, const uint32_t* n_dev, uint32_t n_upper, const float* poses,
                          float2 center, float extent, uint8_t* keep, uint32_t* block_counts);
void launch_aos_to_soa(const Launch& L, const sb_surfel* aos, SurfelPlanes s, uint32_t offset, uint32_t n);
void launch_soa_to_aos(const Launch& L, SurfelPlanes s, sb_surfel* aos, uint32_t n);

}  // namespace sb
