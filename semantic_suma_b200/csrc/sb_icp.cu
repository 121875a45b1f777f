// sb_icp.cu -- K5: per-pixel projective association, point-to-plane residual, 6-DoF Jacobian and the reduction to
// J^T W J / J^T W r, plus the device-resident Gauss-Newton loop.
// Replaces Frame2Model::jacobianProducts (core/Frame2Model.cpp:136-261), its shader
// Frame2Model_jacobians.geom:67-247 and LieGaussNewton::minimize/step (core/LieGaussNewton.cpp:13-79).
//
// B200 mapping
//  * the reference sums 16 RGB texels through ROP float blending (order undefined) and reads 48 floats back with two
//    glFinish per iteration. Here every pixel's fp32 terms are converted to Q33.30 fixed point and summed as int64:
//    registers -> warp shuffles -> one 64-bit RED per value and block. Integer addition is associative, so the
//    result is bit-identical for any grid, any row striping and any number of GPUs.
//  * the Gauss-Newton step (6x6 LDL^T, SE3 exp, stop tests) runs in the last block to finish; iteration i+1 reads the
//    new pose from HBM. The whole minimisation is launched back-to-back with no host round trip.
//  * 6 range images of 2 MiB are L2-resident; the pass is latency/reduction bound, not HBM bound.
#include "sb_gn.cuh"
#include "sb_internal.cuh"

namespace sb {
using namespace sbm;

constexpr int kIcpThreads = 256;

int icp_grid_blocks(int sm_count) { return sm_count * 2; }

__device__ __forceinline__ float4 tex_border(const float4* __restrict__ img, int W, int H, int x, int y) {
  if (x < 0 || x >= W || y < 0 || y >= H) return make_float4(0.f, 0.f, 0.f, 0.f);
  return __ldg(img + (size_t)y * W + x);
}

// texture(sampler2DRect, idx) with NEAREST or LINEAR filtering, CLAMP_TO_BORDER, border (0,0,0,0)
struct Bilin {
  int i0, j0;
  float w00, w10, w01, w11;
};
__device__ __forceinline__ Bilin bilin_setup(float u, float v) {
  Bilin b;
  float a = u - 0.5f, c = v - 0.5f;
  float fa = floorf(a), fc = floorf(c);
  b.i0 = (int)fa;
  b.j0 = (int)fc;
  float al = a - fa, be = c - fc;
  b.w00 = (1.0f - al) * (1.0f - be);
  b.w10 = al * (1.0f - be);
  b.w01 = (1.0f - al) * be;
  b.w11 = al * be;
  return b;
}
__device__ __forceinline__ float4 sample_bilin(const float4* __restrict__ img, int W, int H, const Bilin& b) {
  float4 t00 = tex_border(img, W, H, b.i0, b.j0);
  float4 t10 = tex_border(img, W, H, b.i0 + 1, b.j0);
  float4 t01 = tex_border(img, W, H, b.i0, b.j0 + 1);
  float4 t11 = tex_border(img, W, H, b.i0 + 1, b.j0 + 1);
  float4 r;
  r.x = ((b.w00 * t00.x + b.w10 * t10.x) + b.w01 * t01.x) + b.w11 * t11.x;
  r.y = ((b.w00 * t00.y + b.w10 * t10.y) + b.w01 * t01.y) + b.w11 * t11.y;
  r.z = ((b.w00 * t00.z + b.w10 * t10.z) + b.w01 * t01.z) + b.w11 * t11.z;
  r.w = ((b.w00 * t00.w + b.w10 * t10.w) + b.w01 * t01.w) + b.w11 * t11.w;
  return r;
}

// accumulators of one thread
struct Acc {
  long long s[29];  // 21 lower-triangle JtWJ | 6 JtWr | F | F_inlier
  int n_valid, n_outlier, n_invalid;
};

// Frame2Model_jacobians.geom:84-200 for one data pixel
__device__ __forceinline__ void icp_pixel(const KParams& kp, const IcpArgs& a, const float* M, int pix, int iteration,
                                          Acc& acc) {
  float4 Vd = __ldg(a.data_v + pix);
  float4 Nd = __ldg(a.data_n + pix);
  float e_d = Vd.w + Nd.w;
  if (!(e_d > 1.5f)) {
    acc.n_invalid += 1;
    return;
  }
  V3 v_d = xform_point(M, mk3(Vd.x, Vd.y, Vd.z));
  V3 n_d = xform_dir(M, mk3(Nd.x, Nd.y, Nd.z));
  // project2model, .geom:53-65
  float depth = len3(v_d);
  float yaw = atan2f_(v_d.y, v_d.x);
  float pitch = -asinf_(v_d.z / depth);
  float px = 0.5f * ((-yaw * kInvPi) + 1.0f);
  float py = 1.0f - (kRad2Deg * pitch + kp.fov_up) / kp.fov;
  float ix = px * (float)kp.Wm, iy = py * (float)kp.Hm;
  if (!(ix >= 0.0f && ix < (float)kp.Wm && iy >= 0.0f && iy < (float)kp.Hm)) {
    acc.n_invalid += 1;
    return;
  }
  float4 Vm, Nm;
  Bilin b;
  int tx = 0, ty = 0;
  if (kp.bilinear) {
    b = bilin_setup(ix, iy);
    Vm = sample_bilin(a.model_v, kp.Wm, kp.Hm, b);
    Nm = sample_bilin(a.model_n, kp.Wm, kp.Hm, b);
  } else {
    tx = (int)floorf(ix);
    ty = (int)floorf(iy);
    Vm = tex_border(a.model_v, kp.Wm, kp.Hm, tx, ty);
    Nm = tex_border(a.model_n, kp.Wm, kp.Hm, tx, ty);
  }
  float e_m = Vm.w + Nm.w;
  if (!(e_m > 1.5f)) {
    acc.n_invalid += 1;
    return;
  }
  V3 v_m = mk3(Vm.x, Vm.y, Vm.z), n_m = mk3(Nm.x, Nm.y, Nm.z);
  bool inlier = true;
  if (len3(sub3(v_m, v_d)) > a.distance_thresh) inlier = false;
  if (dot3(n_m, n_d) < a.angle_thresh) inlier = false;
  float residual = dot3(n_m, sub3(v_d, v_m));
  V3 cp = cross3(v_d, n_m);
  float weight = 1.0f;
  if (kp.weighting == 4 || kp.weighting == 1) {
    if (fabsf(residual) > kp.factor) weight = kp.factor / fabsf(residual);
  } else if (kp.weighting == 2 && iteration > 0) {
    if (fabsf(residual) > kp.factor) {
      weight = 0.0f;
    } else {
      float alpha = residual / kp.factor;
      weight = 1.0f - alpha * alpha;
      weight = weight * weight;
    }
  }
  if (a.has_semantics) {  // .geom:144-158
    float4 Sm = kp.bilinear ? sample_bilin(a.model_s, kp.Wm, kp.Hm, b) : tex_border(a.model_s, kp.Wm, kp.Hm, tx, ty);
    float4 Sd = __ldg(a.data_s + pix);
    float data_label = Sd.x * 255.0f, data_prob = Sd.w, model_label = Sm.x * 255.0f;
    if (is_movable(model_label)) {
      if (roundf_(data_label) != roundf_(model_label))
        weight = weight * (1.0f - data_prob);
      else
        weight = weight * data_prob;
    }
  }
  acc.n_valid += 1;
  float wr = weight * residual;
  acc.s[27] += fix30(wr * residual);
  if (inlier) {
    const float J[6] = {n_m.x, n_m.y, n_m.z, cp.x, cp.y, cp.z};
    int k = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      float wc = weight * J[c];
#pragma unroll
      for (int r = c; r < 6; ++r) acc.s[k++] += fix30(wc * J[r]);
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) acc.s[21 + r] += fix30(wr * J[r]);
    acc.s[28] += fix30(wr * residual);
  } else {
    acc.n_outlier += 1;
  }
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// block-level reduction of Acc into 32 global int64 accumulators
__device__ __forceinline__ void block_reduce_to_global(Acc& acc, long long* __restrict__ g_acc) {
  __shared__ long long sm[kIcpThreads / 32][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 29; ++i) {
    long long v = warp_sum_ll(acc.s[i]);
    if (lane == 0) sm[warp][i] = v;
  }
  int nv = __reduce_add_sync(0xffffffffu, acc.n_valid);
  int no = __reduce_add_sync(0xffffffffu, acc.n_outlier);
  int ni = __reduce_add_sync(0xffffffffu, acc.n_invalid);
  if (lane == 0) {
    sm[warp][29] = nv;
    sm[warp][30] = no;
    sm[warp][31] = ni;
  }
  __syncthreads();
  if (warp == 0) {
    long long v = 0;
#pragma unroll
    for (int w = 0; w < kIcpThreads / 32; ++w) v += sm[w][lane];
    if (v != 0) atomicAdd((unsigned long long*)(g_acc + lane), (unsigned long long)v);
    __threadfence();
  }
}

__device__ __forceinline__ void icp_accumulate(const KParams& kp, const IcpArgs& a, const float* M, int iteration,
                                               Acc& acc) {
#pragma unroll
  for (int i = 0; i < 29; ++i) acc.s[i] = 0;
  acc.n_valid = acc.n_outlier = acc.n_invalid = 0;
  const int begin = a.row_begin * kp.W, end = a.row_end * kp.W;
  for (int pix = begin + blockIdx.x * blockDim.x + threadIdx.x; pix < end; pix += gridDim.x * blockDim.x)
    icp_pixel(kp, a, M, pix, iteration, acc);
}

// ---- plain evaluation: one K5 pass, result = 32 int64 sums in g_acc (zeroed by the caller) ----
__global__ void __launch_bounds__(kIcpThreads) k_icp_jacobian(KParams kp, IcpArgs a, Mat4 pose, int iteration,
                                                              long long* __restrict__ g_acc) {
  float M[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) M[i] = pose.m[i];
  Acc acc;
  icp_accumulate(kp, a, M, iteration, acc);
  block_reduce_to_global(acc, g_acc);
}

void launch_icp_jacobian(const Launch& L, const KParams& kp, const IcpArgs& a, const Mat4& pose, int iteration,
                         long long* acc32, int blocks) {
  cudaMemsetAsync(acc32, 0, 32 * sizeof(long long), L.stream);
  {
    ScopedKernel sk(L, K_ICP_JACOBIAN);
    k_icp_jacobian<<<blocks, kIcpThreads, 0, L.stream>>>(kp, a, pose, iteration, acc32);
  }
}

// ---- fused Gauss-Newton iteration ----
__global__ void k_gn_init(GnState* st, Mat4d T0, long long* acc32, unsigned int* ticket) {
  int i = threadIdx.x;
  if (i < 16) st->pose[i] = T0.m[i];
  if (i < 32) acc32[i] = 0;
  if (i < 48) st->out48[i] = 0.0;
  if (i == 0) {
    st->last_error = (double)3.402823466e+38f;  // numeric_limits<float>::max(), LieGaussNewton.cpp:48
    st->k = 0;
    st->done = 0;
    st->history_len = 0;
    *ticket = 0;
  }
}

void launch_gn_init(const Launch& L, GnState* st, const Mat4d& T0, long long* acc32, unsigned int* ticket) {
  {
    ScopedKernel sk(L, K_GN_INIT);
    k_gn_init<<<1, 64, 0, L.stream>>>(st, T0, acc32, ticket);
  }
}

// one-shot all-reduce of the 32 sums over the ranks' peer-mapped mailboxes (only thread 0 of the last block runs it):
// every rank stores its sums + an epoch stamp into slot [epoch&1][rank] of every peer, then waits for the stamps.
__device__ void comm_allreduce32(const CommDev& cd, long long* raw) {
  const int epoch = (int)(*cd.epoch) + 1;  // stamps start at 1: a zeroed mailbox never matches
  *cd.epoch = (unsigned int)epoch;
  const int slot = epoch & 1;
  for (int r = 0; r < cd.nranks; ++r) {
    volatile long long* dst = cd.mailbox[r] + ((size_t)(slot * 8 + cd.rank)) * 40;
    for (int i = 0; i < 32; ++i) dst[i] = raw[i];
  }
  __threadfence_system();
  for (int r = 0; r < cd.nranks; ++r) {
    volatile long long* dst = cd.mailbox[r] + ((size_t)(slot * 8 + cd.rank)) * 40;
    dst[32] = (long long)epoch;
  }
  __threadfence_system();
  volatile long long* mine = cd.mailbox[cd.rank];
  for (int i = 0; i < 32; ++i) raw[i] = 0;
  for (int r = 0; r < cd.nranks; ++r) {
    volatile long long* src = mine + ((size_t)(slot * 8 + r)) * 40;
    while (src[32] != (long long)epoch) {
    }
    __threadfence_system();
    for (int i = 0; i < 32; ++i) raw[i] += src[i];
  }
}

__global__ void __launch_bounds__(kIcpThreads, 2) k_icp_fused(KParams kp, IcpArgs a, GnState* __restrict__ st,
                                                           long long* __restrict__ g_acc, unsigned int* ticket,
                                                           int max_iter, double eps, double delta, CommDev cd) {
  // all blocks read the state written by the previous launch
  if (*(volatile int*)&st->done) return;
  float M[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) M[i] = (float)st->pose[i];  // pose_.cast<float>(), Frame2Model.cpp:194
  const int k = st->k;
  Acc acc;
  icp_accumulate(kp, a, M, k, acc);
  block_reduce_to_global(acc, g_acc);
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  long long raw[32];
  for (int i = 0; i < 32; ++i) {
    raw[i] = *(volatile long long*)(g_acc + i);
    g_acc[i] = 0;
  }
  *ticket = 0;
  if (cd.nranks > 1) comm_allreduce32(cd, raw);
  double out48[48], pose[16], dx[6];
  sbg::unpack48(raw, out48);
  for (int i = 0; i < 16; ++i) pose[i] = st->pose[i];
  int hl = st->history_len;
  for (int i = 0; i < 16; ++i) st->history[hl * 16 + i] = pose[i];  // history_.push_back(Tk_)
  ++hl;
  int result = sbg::gn_step(out48, st->last_error, eps, delta, pose, dx);
  for (int i = 0; i < 16; ++i) st->pose[i] = pose[i];
  for (int i = 0; i < 48; ++i) st->out48[i] = out48[i];
  st->last_error = out48[43];
  int kk = k;
  int done = 0;
  if (result == 0) {
    done = 1;
  } else {
    ++kk;
    if (kk >= max_iter) {  // the loop pushes the pose once more and leaves (LieGaussNewton.cpp:24-27)
      for (int i = 0; i < 16; ++i) st->history[hl * 16 + i] = pose[i];
      ++hl;
      done = 1;
    }
  }
  st->k = kk;
  st->history_len = hl;
  __threadfence();
  st->done = done;
}

void launch_icp_fused_iteration(const Launch& L, const KParams& kp, const IcpArgs& a, GnState* st, long long* acc32,
                                unsigned int* ticket, int max_iter, double eps, double delta, const CommDev* comm,
                                int blocks) {
  CommDev cd;
  if (comm) {
    cd = *comm;
  } else {
    memset(&cd, 0, sizeof(cd));
    cd.nranks = 1;
  }
  {
    ScopedKernel sk(L, K_ICP_FUSED);
    k_icp_fused<<<blocks, kIcpThreads, 0, L.stream>>>(kp, a, st, acc32, ticket, max_iter, eps, delta, cd);
  }
}

}  // namespace sb
