// sb_preprocess.cu -- K1-K3: spherical projection of a scan into vertex / normal / semantic range images.
// Replaces Preprocessing::process (core/Preprocessing.cpp:120-339) and its shaders gen_vertexmap.{vert,frag},
// gen_normalmap.frag, floodfill.frag.
//
// B200 mapping: the z-buffered point rasterisation (GL depth test GL_LESS, earliest primitive wins ties) becomes a
// 64-bit atomicMin on a packed (depth24 << 32 | point index) key image; the images (2 MiB each at 64x2048) stay
// L2-resident between the passes, so every pass is a plain coalesced float4 stream + a handful of L2 gathers.
#include "sb_internal.cuh"

namespace sb {
using namespace sbm;

__global__ void k_fill_u64(unsigned long long* p, unsigned long long v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

void launch_fill_u64(const Launch& L, unsigned long long* p, unsigned long long v, size_t n) {
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  {
    ScopedKernel sk(L, K_FILL);
    k_fill_u64<<<blocks, 256, 0, L.stream>>>(p, v, n);
  }
}

// K1a: gen_vertexmap.vert:73-91 -- one thread per point, scatter-min into the key image
__global__ void __launch_bounds__(256) k_project_scatter(KParams kp, const float4* __restrict__ pts, uint32_t n,
                                                         unsigned long long* __restrict__ keys) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p4 = __ldg(pts + i);
  V3 pos = mk3(p4.x, p4.y, p4.z);
  float d = len3(pos);
  float yaw = atan2f_(pos.y, pos.x);
  float pitch = -asinf_(pos.z / d);
  float x = -yaw * kInvPi;
  float y = 1.0f - (2.0f * (kRad2Deg * pitch + kp.fov_up)) / kp.fov;
  float z = 2.0f * ((d - kp.min_depth) / (kp.max_depth - kp.min_depth)) - 1.0f;
  float fx = floorf((0.5f * (x + 1.0f)) * (float)kp.W);
  float fy = floorf((0.5f * (y + 1.0f)) * (float)kp.H);
  if (!(fx >= 0.0f && fx < (float)kp.W && fy >= 0.0f && fy < (float)kp.H)) return;  // clipped
  if (!(z >= -1.0f && z <= 1.0f)) return;
  uint32_t d24 = depth24(0.5f * z + 0.5f);
  if (d24 >= kDepthClear) return;  // GL_LESS against the cleared buffer
  size_t pix = (size_t)(int)fy * kp.W + (size_t)(int)fx;
  atomicMin(keys + pix, ((unsigned long long)d24 << 32) | (unsigned long long)i);
}

// K1b: resolve the winners: gen_vertexmap.vert:92-102 + .frag:16-23
__global__ void __launch_bounds__(256) k_project_resolve(KParams kp, const float4* __restrict__ pts,
                                                         const float* __restrict__ labels,
                                                         const float* __restrict__ probs, uint32_t n, int isfirst,
                                                         unsigned long long* __restrict__ keys,
                                                         float4* __restrict__ vertex, float4* __restrict__ sem_raw) {
  int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= kp.W * kp.H) return;
  unsigned long long key = keys[pix];
  keys[pix] = ~0ull;  // consume and clear: the key image is ready for the next scan without a separate fill
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f), s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (key != ~0ull) {
    uint32_t i = (uint32_t)(key & 0xffffffffull);
    float4 p4 = __ldg(pts + i);
    float label = 0.0f, prob = 0.0f;
    if (labels) {  // Q1: attribute offsets 16 / 20 bytes on stride-4 buffers (Preprocessing.cpp:142-145)
      uint32_t li = kp.label_offset_quirk ? i + 4 : i;
      if (li < n) label = __ldg(labels + li);
    }
    if (probs) {
      uint32_t pi = kp.label_offset_quirk ? i + 5 : i;
      if (pi < n) prob = __ldg(probs + pi);
    }
    if (!(isfirst && is_movable(label))) v = make_float4(p4.x, p4.y, p4.z, 1.0f);
    float l = label / 255.0f;
    s = make_float4(l, l, l, prob);
  }
  vertex[pix] = v;
  sem_raw[pix] = s;
}

__device__ __forceinline__ float4 tex_wrapx(const float4* __restrict__ img, int W, int H, int x, int y) {
  if (y < 0 || y >= H) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (x >= W) x -= W;
  if (x < 0) x += W;
  return __ldg(img + (size_t)y * W + x);
}

// K2: gen_normalmap.frag:41-99 (normals from the +x / +y neighbours, label erosion)
__global__ void __launch_bounds__(256) k_normals_erode(KParams kp, const float4* __restrict__ vertex,
                                                       const float4* __restrict__ sem_raw, float4* __restrict__ normal,
                                                       float4* __restrict__ eroded) {
  int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= kp.W * kp.H) return;
  const int W = kp.W, H = kp.H;
  int y = pix / W, x = pix - y * W;
  float4 P = __ldg(vertex + pix);
  float4 nrm = make_float4(0.f, 0.f, 0.f, 1.f), ero = make_float4(0.f, 0.f, 0.f, 1.f);  // Q2
  if (P.w > 0.0f) {
    float nw = 1.0f;
    float4 U = tex_wrapx(vertex, W, H, x + 1, y);
    float4 V = tex_wrapx(vertex, W, H, x, y + 1);
    float4 S = tex_wrapx(vertex, W, H, x - 1, y);
    float4 T = tex_wrapx(vertex, W, H, x, y - 1);
    if (U.w < 1.0f && V.w < 1.0f) nw = 0.0f;
    if (S.w < 1.0f && T.w < 1.0f) nw = 0.0f;
    if (!(U.w > 0.5f) || !(V.w > 0.5f)) nw = 0.0f;
    float4 sp = __ldg(sem_raw + pix);
    ero = sp;
    float pl = sp.x;
    float ul = tex_wrapx(sem_raw, W, H, x + 1, y).x;
    float vl = tex_wrapx(sem_raw, W, H, x, y + 1).x;
    float sl = tex_wrapx(sem_raw, W, H, x - 1, y).x;
    float tl = tex_wrapx(sem_raw, W, H, x, y - 1).x;
    if ((pl != ul && ul != 0.0f) || (pl != vl && vl != 0.0f) || (pl != sl && sl != 0.0f) || (pl != tl && tl != 0.0f))
      ero = make_float4(0.f, 0.f, 0.f, 1.f);
    if (nw > 0.0f) {
      V3 pp = mk3(P.x, P.y, P.z);
      V3 u = normalize3(sub3(mk3(U.x, U.y, U.z), pp));
      V3 v = normalize3(sub3(mk3(V.x, V.y, V.z), pp));
      V3 w = cross3(u, v);
      float len = len3(w);
      nrm = make_float4(w.x / len, w.y / len, w.z / len, (len > 0.0000001f) ? 1.0f : 0.0f);
    } else {
      nrm = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  normal[pix] = nrm;
  eroded[pix] = ero;
}

// K3: floodfill.frag:34-84
__global__ void __launch_bounds__(256) k_floodfill(KParams kp, const float4* __restrict__ vertex,
                                                   const float4* __restrict__ eroded, float4* __restrict__ semantic) {
  int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= kp.W * kp.H) return;
  const int W = kp.W, H = kp.H;
  int y = pix / W, x = pix - y * W;
  float4 out = __ldg(eroded + pix);
  if (out.x == 0.0f) {
    float4 P = __ldg(vertex + pix);
    float lp = len3(mk3(P.x, P.y, P.z));
    bool done = false;
#pragma unroll
    for (int offset = 1; offset < 3; ++offset) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (done) continue;
        int qx = x + ((k == 0) ? offset : (k == 2 ? -offset : 0));
        int qy = y + ((k == 1) ? offset : (k == 3 ? -offset : 0));
        float4 ql = tex_wrapx(eroded, W, H, qx, qy);
        if (ql.x != 0.0f) {
          float4 Q = tex_wrapx(vertex, W, H, qx, qy);
          float lq = len3(mk3(Q.x, Q.y, Q.z));
          if (fabsf(lp - lq) < 0.007f * lp) {
            out = make_float4(ql.x, ql.y, ql.z, ql.w / (float)(offset + 1));
            done = true;
          }
        }
      }
    }
  }
  semantic[pix] = out;
}

void launch_preprocess(const Launch& L, const KParams& kp, const float4* pts, const float* labels, const float* probs,
                       uint32_t n, uint32_t timestamp, unsigned long long* keys, float4* sem_raw, float4* eroded,
                       FrameDev out) {
  const int P = kp.W * kp.H;
  if (n > 0) {
    {
      ScopedKernel sk(L, K_PROJECT_SCATTER);
      k_project_scatter<<<(n + 255) / 256, 256, 0, L.stream>>>(kp, pts, n, keys);
    }
  }
  const int pb = (P + 255) / 256;
  {
    ScopedKernel sk(L, K_PROJECT_RESOLVE);
    k_project_resolve<<<pb, 256, 0, L.stream>>>(kp, pts, labels, probs, n, timestamp < 10 ? 1 : 0, keys, out.vertex,
                                                sem_raw);
  }
  {
    ScopedKernel sk(L, K_NORMALS_ERODE);
    k_normals_erode<<<pb, 256, 0, L.stream>>>(kp, out.vertex, sem_raw, out.normal, eroded);
  }
  {
    ScopedKernel sk(L, K_FLOODFILL);
    k_floodfill<<<pb, 256, 0, L.stream>>>(kp, out.vertex, eroded, out.semantic);
  }
}

}  // namespace sb
