// sb_map.cu -- K4 (model rendering) and K6 (surfel update / initialise / merge) of the surfel map.
// Replaces SurfelMap::render* (core/SurfelMap.cpp:847-1165; render_surfels.{vert,geom,frag}, render_compose.frag)
// and SurfelMap::update (core/SurfelMap.cpp:492-698; gen_indexmap, init_radiusConf, update_surfels, gen_surfels,
// copy_surfels, extract_surfels).
//
// B200 mapping
//  * GL triangle rasterisation + 24-bit z-buffer  ->  one thread per surfel walks the quad's pixel bounding box
//    with exact integer edge functions and resolves visibility with a 64-bit atomicMin on
//    (depth24 << 40 | pass << 32 | surfel index): "earliest primitive wins" falls out of the key order.
//    One pass over the surfel lanes feeds the old / new / composed views at once (the reference draws the whole
//    buffer four times per render()).
//  * transform feedback (ordered stream compaction)  ->  per-block counts + a single-block scan + an ordered scatter,
//    so surfel order (and therefore every surfel index) is identical to the reference's.
//  * surfels live as four float4 lanes (SoA), every pass streams only the lanes it needs with 128-bit accesses.
#include "sb_gn.cuh"
#include "sb_internal.cuh"

namespace sb {
using namespace sbm;

constexpr int kThreads = 256;

const char* kernel_name(int id) {
  static const char* names[K_COUNT] = {"fill_u64", "project_scatter", "preprocess_tile",
                                       "icp_jacobian", "gn_init", "icp_fused", "pose_products", "render_scatter",
                                       "render_resolve", "index_scatter", "radius", "update_surfels", "gen_surfels",
                                       "extract_flags", "scan_blocks", "compact_scatter", "aos_to_soa", "soa_to_aos"};
  return (id >= 0 && id < K_COUNT) ? names[id] : "?";
}

__device__ __forceinline__ int pose_index(float count) {
  int c = (int)count;
  c = c < 0 ? 0 : c;
  c = c >= (int)kMaxPoses ? (int)kMaxPoses - 1 : c;
  return c;
}

__device__ __forceinline__ void load_mat(const float* __restrict__ table, int idx, float* M) {
  const float4* t = reinterpret_cast<const float4*>(table + 16 * (size_t)idx);
  float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2), d = __ldg(t + 3);
  M[0] = a.x; M[1] = a.y; M[2] = a.z; M[3] = a.w;
  M[4] = b.x; M[5] = b.y; M[6] = b.z; M[7] = b.w;
  M[8] = c.x; M[9] = c.y; M[10] = c.z; M[11] = c.w;
  M[12] = d.x; M[13] = d.y; M[14] = d.z; M[15] = d.w;
}

// out[t] = A * poses[t]  (render_surfels.vert:46 evaluates (inv_pose * surfelPose) * v; the product is shared by all
// surfels created at time t)
__global__ void k_pose_products(Mat4 A, const Mat4* __restrict__ A_dev, const float* __restrict__ poses,
                                float* __restrict__ out, uint32_t count) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  float a[16], b[16], c[16];
  for (int i = 0; i < 16; ++i) a[i] = A_dev ? A_dev->m[i] : A.m[i];
  load_mat(poses, (int)t, b);
  mat4_mul<float>(a, b, c);
  for (int i = 0; i < 16; ++i) out[16 * (size_t)t + i] = c[i];
}

void launch_pose_products(const Launch& L, const Mat4& A, const Mat4* A_dev, const float* poses, float* out,
                          uint32_t count) {
  if (count == 0) return;
  {
    ScopedKernel sk(L, K_POSE_PRODUCTS);
    k_pose_products<<<(count + 127) / 128, 128, 0, L.stream>>>(A, A_dev, poses, out, count);
  }
}

// ------------------------------------------------------------------------------------------------------------
// K4 rasteriser
// ------------------------------------------------------------------------------------------------------------
struct RVert {
  long long X, Y;
  float z, tx, ty;
};

__device__ __forceinline__ long long floor_div256(long long a) { return a >> 8; }  // arithmetic shift = floor
__device__ __forceinline__ long long ceil_div256(long long a) { return -((-a) >> 8); }
__device__ __forceinline__ long long edge_fn(const RVert& P, const RVert& Q, long long X, long long Y) {
  return (Q.X - P.X) * (Y - P.Y) - (Q.Y - P.Y) * (X - P.X);
}

// what a fragment of one surfel writes to: which key images, with which low key bits
struct Emit {
  uint32_t k;
  uint32_t flags;  // bit0 old-class surfel, bit1 new-class surfel, bit2 GL_LEQUAL composed mode
};

__device__ __forceinline__ bool edge_tie(const RVert& P, const RVert& Q) {  // rule for pixels exactly on an edge
  long long dx = Q.X - P.X, dy = Q.Y - P.Y;
  return dy > 0 || (dy == 0 && dx > 0);
}

// Triangle prepared for pixel-centre coverage tests with exact integer edge functions. The edge values at the first
// pixel centre of the bounding box and their per-pixel steps are int64 (no rounding anywhere):
// E(P + (256,0)) = E(P) - 256*dy, E(P + (0,256)) = E(P) + 256*dx.
struct TriSetup {
  long long eA, eB, eC;     // edge(B,C), edge(C,A), edge(A,B) at the first pixel centre
  long long sxA, sxB, sxC;  // step per pixel in x
  long long syA, syB, syC;  // step per pixel in y
  float farea;
  float zA, zB, zC, txA, txB, txC, tyA, tyB, tyC;
  int i0, j0, ni, nj;       // bounding box origin and extent (inclusive), clamped to the image
  uint32_t ties;            // bit0..2: pixels exactly on edge A/B/C count as inside
  uint32_t valid;
};

__device__ __forceinline__ TriSetup tri_prepare(RVert A, RVert B, RVert C, int W, int H) {
  TriSetup t;
  t.valid = 0;
  long long area = (B.X - A.X) * (C.Y - A.Y) - (B.Y - A.Y) * (C.X - A.X);
  if (area == 0) return t;
  if (area < 0) {
    RVert x = B;
    B = C;
    C = x;
    area = -area;
  }
  long long minX = min(A.X, min(B.X, C.X)), maxX = max(A.X, max(B.X, C.X));
  long long minY = min(A.Y, min(B.Y, C.Y)), maxY = max(A.Y, max(B.Y, C.Y));
  long long i0 = ceil_div256(minX - 128), i1 = floor_div256(maxX - 128);
  long long j0 = ceil_div256(minY - 128), j1 = floor_div256(maxY - 128);
  if (i0 < 0) i0 = 0;
  if (j0 < 0) j0 = 0;
  if (i1 > W - 1) i1 = W - 1;
  if (j1 > H - 1) j1 = H - 1;
  if (i0 > i1 || j0 > j1) return t;
  t.farea = (float)area;
  const long long X0 = i0 * 256 + 128, Y0 = j0 * 256 + 128;
  t.eA = edge_fn(B, C, X0, Y0);
  t.eB = edge_fn(C, A, X0, Y0);
  t.eC = edge_fn(A, B, X0, Y0);
  t.sxA = -256 * (C.Y - B.Y); t.sxB = -256 * (A.Y - C.Y); t.sxC = -256 * (B.Y - A.Y);
  t.syA = 256 * (C.X - B.X); t.syB = 256 * (A.X - C.X); t.syC = 256 * (B.X - A.X);
  t.ties = (edge_tie(B, C) ? 1u : 0u) | (edge_tie(C, A) ? 2u : 0u) | (edge_tie(A, B) ? 4u : 0u);
  t.zA = A.z; t.zB = B.z; t.zC = C.z;
  t.txA = A.tx; t.txB = B.tx; t.txC = C.tx;
  t.tyA = A.ty; t.tyB = B.ty; t.tyC = C.ty;
  t.i0 = (int)i0; t.j0 = (int)j0; t.ni = (int)(i1 - i0); t.nj = (int)(j1 - j0);
  t.valid = 1;
  return t;
}

// fragment stage for one covered pixel (render_surfels.frag:19-33 + depth test)
__device__ __forceinline__ void tri_fragment(const TriSetup& t, const Emit& e, const RenderTargets& rt, int W,
                                             long long wB, long long wC, int di, int dj) {
  float fB = (float)wB / t.farea, fC = (float)wC / t.farea;
  float fA = (1.0f - fB) - fC;
  float tx = (fA * t.txA + fB * t.txB) + fC * t.txC;
  float ty = (fA * t.tyA + fB * t.tyB) + fC * t.tyC;
  if (tx * tx + ty * ty > 1.0f) return;  // outside the disc
  float z = (fA * t.zA + fB * t.zB) + fC * t.zC;
  if (!(z >= 0.0f && z <= 1.0f)) return;  // near / far clip
  unsigned long long d = (unsigned long long)depth24(z);
  const bool lequal = (e.flags & 4u) != 0;
  if (!lequal && d >= kDepthClear) return;  // GL_LESS against the cleared depth
  size_t pix = (size_t)(t.j0 + dj) * W + (size_t)(t.i0 + di);
  unsigned long long hi = d << 40;
  // keys only ever decrease: a (possibly stale, L1-cached) read that is already smaller proves the atomic cannot
  // win, so most hidden fragments never reach the L2 atomic unit
  if (e.flags & 1u) {  // old-class surfel
    if (rt.key_old) {
      unsigned long long key = hi | e.k;
      if (key < rt.key_old[pix]) atomicMin(rt.key_old + pix, key);
    }
    if (rt.key_comp) {  // old pass is drawn first: it wins depth ties under GL_LESS, loses them under GL_LEQUAL
      unsigned long long key = lequal ? (hi | (1ull << 32) | (unsigned long long)(0xffffffffu - e.k)) : (hi | e.k);
      if (key < rt.key_comp[pix]) atomicMin(rt.key_comp + pix, key);
    }
  }
  if (e.flags & 2u) {  // new-class surfel
    if (rt.key_new) {
      unsigned long long key = hi | e.k;
      if (key < rt.key_new[pix]) atomicMin(rt.key_new + pix, key);
    }
    if (rt.key_comp) {
      unsigned long long key = lequal ? (hi | (unsigned long long)(0xffffffffu - e.k)) : (hi | (1ull << 32) | e.k);
      if (key < rt.key_comp[pix]) atomicMin(rt.key_comp + pix, key);
    }
  }
}

constexpr int kRenderThreads = 128;
constexpr int kTrisPerWarp = 64;  // two triangles per surfel, 32 surfels per warp pass

// the triangles of one warp pass, structure-of-arrays in shared memory (lanes read different triangles: an array per
// field keeps those reads spread over the banks)
struct WarpTris {
  long long e[3][kTrisPerWarp], sx[3][kTrisPerWarp], sy[3][kTrisPerWarp];
  float farea[kTrisPerWarp], z[3][kTrisPerWarp], tx[3][kTrisPerWarp], ty[3][kTrisPerWarp];
  int i0[kTrisPerWarp], j0[kTrisPerWarp], ni[kTrisPerWarp], nj[kTrisPerWarp];
  uint32_t ties[kTrisPerWarp], k[kTrisPerWarp], flags[kTrisPerWarp];
  int prefix[kTrisPerWarp + 1];  // exclusive prefix sums of the bounding-box pixel counts
};

__device__ __forceinline__ void tris_store(WarpTris& w, int slot, const TriSetup& t, const Emit& e) {
  w.e[0][slot] = t.eA; w.e[1][slot] = t.eB; w.e[2][slot] = t.eC;
  w.sx[0][slot] = t.sxA; w.sx[1][slot] = t.sxB; w.sx[2][slot] = t.sxC;
  w.sy[0][slot] = t.syA; w.sy[1][slot] = t.syB; w.sy[2][slot] = t.syC;
  w.farea[slot] = t.farea;
  w.z[0][slot] = t.zA; w.z[1][slot] = t.zB; w.z[2][slot] = t.zC;
  w.tx[0][slot] = t.txA; w.tx[1][slot] = t.txB; w.tx[2][slot] = t.txC;
  w.ty[0][slot] = t.tyA; w.ty[1][slot] = t.tyB; w.ty[2][slot] = t.tyC;
  w.i0[slot] = t.i0; w.j0[slot] = t.j0; w.ni[slot] = t.ni; w.nj[slot] = t.nj;
  w.ties[slot] = t.ties; w.k[slot] = e.k; w.flags[slot] = e.flags;
  w.prefix[slot + 1] = (t.ni + 1) * (t.nj + 1);
}
__device__ __forceinline__ void tris_load(const WarpTris& w, int slot, TriSetup& t, Emit& e) {
  t.eA = w.e[0][slot]; t.eB = w.e[1][slot]; t.eC = w.e[2][slot];
  t.sxA = w.sx[0][slot]; t.sxB = w.sx[1][slot]; t.sxC = w.sx[2][slot];
  t.syA = w.sy[0][slot]; t.syB = w.sy[1][slot]; t.syC = w.sy[2][slot];
  t.farea = w.farea[slot];
  t.zA = w.z[0][slot]; t.zB = w.z[1][slot]; t.zC = w.z[2][slot];
  t.txA = w.tx[0][slot]; t.txB = w.tx[1][slot]; t.txC = w.tx[2][slot];
  t.tyA = w.ty[0][slot]; t.tyB = w.ty[1][slot]; t.tyC = w.ty[2][slot];
  t.i0 = w.i0[slot]; t.j0 = w.j0[slot]; t.ni = w.ni[slot]; t.nj = w.nj[slot];
  t.ties = w.ties[slot]; t.valid = 1;
  e.k = w.k[slot]; e.flags = w.flags[slot];
}

// Balanced rasterisation of all triangles of the warp pass: the bounding-box pixels of all triangles form one
// flattened list of T items, lane l walks items [l*ch, (l+1)*ch). Quads differ by two orders of magnitude in size
// (a surfel created far away and seen from close by covers hundreds of pixels); with one lane per quad the warp runs
// as long as its largest quad while most lanes idle. Edge functions stay exact: incremental int64 adds inside a row,
// re-based at every row / triangle change.
__device__ __forceinline__ void raster_balanced(const WarpTris& w, int ntris, int total, const RenderTargets& rt, int W,
                                                int lane) {
  const int ch = (total + 31) >> 5;
  int item = lane * ch;
  const int end = min(item + ch, total);
  if (item >= end) return;
  int lo = 0, hi = ntris;  // largest j with prefix[j] <= item
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (w.prefix[mid] <= item) lo = mid; else hi = mid;
  }
  int j = lo;
  while (item < end) {
    TriSetup t;
    Emit e;
    tris_load(w, j, t, e);
    const int first = w.prefix[j], cnt = w.prefix[j + 1] - first;
    const int local = item - first;
    const int wd = t.ni + 1;
    int dj = local / wd, di = local - dj * wd;
    int n_here = min(end - item, cnt - local);
    const bool tieA = t.ties & 1u, tieB = t.ties & 2u, tieC = t.ties & 4u;
    long long rowA = t.eA + (long long)dj * t.syA, rowB = t.eB + (long long)dj * t.syB, rowC = t.eC + (long long)dj * t.syC;
    long long wA = rowA + (long long)di * t.sxA, wB = rowB + (long long)di * t.sxB, wC = rowC + (long long)di * t.sxC;
    for (int q = 0; q < n_here; ++q) {
      bool in = (wA > 0 || (wA == 0 && tieA)) && (wB > 0 || (wB == 0 && tieB)) && (wC > 0 || (wC == 0 && tieC));
      if (in) tri_fragment(t, e, rt, W, wB, wC, di, dj);
      if (++di > t.ni) {
        di = 0;
        ++dj;
        rowA += t.syA; rowB += t.syB; rowC += t.syC;
        wA = rowA; wB = rowB; wC = rowC;
      } else {
        wA += t.sxA; wB += t.sxB; wC += t.sxC;
      }
    }
    item += n_here;
    ++j;
  }
}

// render_surfels.vert:42-54 + .geom:76-122 for every surfel of the map. One lane per surfel for the transform, the
// visibility tests, the corner projection and the triangle setup; the rasterisation of the warp's triangles is then
// shared evenly by its 32 lanes (raster_balanced).
__global__ void __launch_bounds__(kRenderThreads) k_render_scatter(KParams kp, SurfelPlanes s, const uint32_t* __restrict__ n_dev,
                                                                   const float* __restrict__ Mtab, float conf_thr, int t_thr,
                                                                   int emit_old, int emit_new, int lequal, RenderTargets rt) {
  __shared__ WarpTris s_tris[kRenderThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpTris& wt = s_tris[warp];
  wt.prefix[2 * lane + 1] = 0;
  wt.prefix[2 * lane + 2] = 0;
  if (lane == 0) wt.prefix[0] = 0;
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  bool alive = k < *n_dev;
  float4 p0, p1, p2;
  bool is_old = false, is_new = false;
  if (alive) {
    p2 = __ldg(s.p2 + k);
    p1 = __ldg(s.p1 + k);
    alive = !kp.use_stability || p1.w > conf_thr;  // .geom:87 (cheap test first)
  }
  if (alive) {
    int creation = (int)p2.w, ts = (int)__float_as_uint(p2.x);
    is_old = emit_old && (creation < t_thr);                  // .geom:90
    is_new = emit_new && (creation >= t_thr || ts >= t_thr);  // .geom:91
    alive = is_old || is_new;
  }
  TriSetup t0, t1;
  t0.valid = t1.valid = 0;
  if (alive) {
    p0 = __ldg(s.p0 + k);
    float M[16];
    load_mat(Mtab, pose_index(p2.w), M);
    V3 pp = xform_point(M, mk3(p0.x, p0.y, p0.z));
    V3 nn = xform_dir(M, mk3(p1.x, p1.y, p1.z));
    float r = p0.w;
    bool visible = dot3(nn, divs3(neg3(pp), len3(pp))) > 0.01f;  // .geom:84
    float cx, cy, cz;
    project01(pp, kp.m_fov_up, kp.m_fov, kp.m_min_depth, kp.m_max_depth, cx, cy, cz);
    if (visible && cx >= 0.0f && cy >= 0.0f && cz >= 0.0f && cx < 1.0f && cy < 1.0f && cz < 1.0f) {
      V3 u = normalize3(mk3(nn.y - nn.z, -nn.x, nn.x));
      V3 v = normalize3(cross3(nn, u));
      V3 ru = scale3(r, u), rv = scale3(r, v);
      V3 corner[4];
      corner[0] = sub3(sub3(pp, ru), rv);
      corner[1] = sub3(add3(pp, ru), rv);
      corner[2] = add3(sub3(pp, ru), rv);
      corner[3] = add3(add3(pp, ru), rv);
      RVert q[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x, y, z;
        project01(corner[i], kp.m_fov_up, kp.m_fov, kp.m_min_depth, kp.m_max_depth, x, y, z);
        if (cx - x > 0.5f) x += 1.0f;  // .geom:68
        if (x - cx > 0.5f) x -= 1.0f;  // .geom:69
        float xw = (0.5f * (2.0f * x - 1.0f) + 0.5f) * (float)kp.Wm;
        float yw = (0.5f * (2.0f * y - 1.0f) + 0.5f) * (float)kp.Hm;
        q[i].z = 0.5f * (2.0f * z - 1.0f) + 0.5f;
        q[i].X = __float2ll_rn(xw * 256.0f);
        q[i].Y = __float2ll_rn(yw * 256.0f);
        q[i].tx = (i & 1) ? 1.0f : -1.0f;
        q[i].ty = (i & 2) ? 1.0f : -1.0f;
      }
      t0 = tri_prepare(q[0], q[1], q[2], kp.Wm, kp.Hm);
      t1 = tri_prepare(q[1], q[2], q[3], kp.Wm, kp.Hm);
    }
  }
  // slots of this lane's triangles: exclusive prefix of the per-lane triangle counts
  const int cnt = (int)t0.valid + (int)t1.valid;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int ntris = __shfl_sync(0xffffffffu, incl, 31);
  if (ntris == 0) return;
  __syncwarp();
  if (cnt) {
    Emit e;
    e.k = k;
    e.flags = (is_old ? 1u : 0u) | (is_new ? 2u : 0u) | (lequal ? 4u : 0u);
    int slot = incl - cnt;
    if (t0.valid) tris_store(wt, slot++, t0, e);
    if (t1.valid) tris_store(wt, slot, t1, e);
  }
  __syncwarp();
  // prefix sums of the pixel counts (pairs of adjacent slots per lane)
  int a = wt.prefix[2 * lane + 1], b = wt.prefix[2 * lane + 2];
  int sum = a + b, isum = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, isum, o);
    if (lane >= o) isum += v;
  }
  const int total = __shfl_sync(0xffffffffu, isum, 31);
  __syncwarp();
  wt.prefix[2 * lane + 1] = isum - sum + a;
  wt.prefix[2 * lane + 2] = isum;
  __syncwarp();
  raster_balanced(wt, ntris, total, rt, kp.Wm, lane);
}

// ------------------------------------------------------------------------------------------------------------
// EXPERIMENTAL (opt-in: SUMA_B200_RENDER_VARIANT=1|2, not measured yet -- DESIGN.md section 8): the same kernel with a
// leaner per-warp triangle store, so that more blocks fit on an SM. ncu on the default kernel shows 51 % issue-slot use
// at 25 % occupancy (86 registers, 36.9 KB shared memory per block). Here the x/y edge steps are kept as int32 (they
// are 256 * a vertex-coordinate difference of one quad, |.| <= 2^28), the texture coordinates as six sign bits (they
// are +-1), the bounding box in two packed words: 21.5 KB per block. Arithmetic and results are identical.
// ------------------------------------------------------------------------------------------------------------
struct WarpTrisLite {
  long long e[3][kTrisPerWarp];
  int sx[3][kTrisPerWarp], sy[3][kTrisPerWarp];
  float farea[kTrisPerWarp], z[3][kTrisPerWarp];
  uint32_t boxx[kTrisPerWarp], boxy[kTrisPerWarp];  // i0 | ni << 16, j0 | nj << 16
  uint32_t k[kTrisPerWarp];
  uint32_t meta[kTrisPerWarp];  // bits 0-2 ties, 3-5 flags, 6-8 tx sign of A,B,C (set = +1), 9-11 ty sign of A,B,C
  int prefix[kTrisPerWarp + 1];
};

__device__ __forceinline__ void tris_store_lite(WarpTrisLite& w, int slot, const TriSetup& t, const Emit& e) {
  w.e[0][slot] = t.eA; w.e[1][slot] = t.eB; w.e[2][slot] = t.eC;
  w.sx[0][slot] = (int)t.sxA; w.sx[1][slot] = (int)t.sxB; w.sx[2][slot] = (int)t.sxC;
  w.sy[0][slot] = (int)t.syA; w.sy[1][slot] = (int)t.syB; w.sy[2][slot] = (int)t.syC;
  w.farea[slot] = t.farea;
  w.z[0][slot] = t.zA; w.z[1][slot] = t.zB; w.z[2][slot] = t.zC;
  w.boxx[slot] = (uint32_t)t.i0 | ((uint32_t)t.ni << 16);
  w.boxy[slot] = (uint32_t)t.j0 | ((uint32_t)t.nj << 16);
  w.k[slot] = e.k;
  w.meta[slot] = (t.ties & 7u) | ((e.flags & 7u) << 3) | ((t.txA > 0.0f ? 1u : 0u) << 6) | ((t.txB > 0.0f ? 1u : 0u) << 7) |
                 ((t.txC > 0.0f ? 1u : 0u) << 8) | ((t.tyA > 0.0f ? 1u : 0u) << 9) | ((t.tyB > 0.0f ? 1u : 0u) << 10) |
                 ((t.tyC > 0.0f ? 1u : 0u) << 11);
  w.prefix[slot + 1] = (t.ni + 1) * (t.nj + 1);
}
__device__ __forceinline__ void tris_load_lite(const WarpTrisLite& w, int slot, TriSetup& t, Emit& e) {
  t.eA = w.e[0][slot]; t.eB = w.e[1][slot]; t.eC = w.e[2][slot];
  t.sxA = (long long)w.sx[0][slot]; t.sxB = (long long)w.sx[1][slot]; t.sxC = (long long)w.sx[2][slot];
  t.syA = (long long)w.sy[0][slot]; t.syB = (long long)w.sy[1][slot]; t.syC = (long long)w.sy[2][slot];
  t.farea = w.farea[slot];
  t.zA = w.z[0][slot]; t.zB = w.z[1][slot]; t.zC = w.z[2][slot];
  const uint32_t m = w.meta[slot], bx = w.boxx[slot], by = w.boxy[slot];
  t.txA = (m & (1u << 6)) ? 1.0f : -1.0f; t.txB = (m & (1u << 7)) ? 1.0f : -1.0f; t.txC = (m & (1u << 8)) ? 1.0f : -1.0f;
  t.tyA = (m & (1u << 9)) ? 1.0f : -1.0f; t.tyB = (m & (1u << 10)) ? 1.0f : -1.0f; t.tyC = (m & (1u << 11)) ? 1.0f : -1.0f;
  t.i0 = (int)(bx & 0xffffu); t.ni = (int)(bx >> 16);
  t.j0 = (int)(by & 0xffffu); t.nj = (int)(by >> 16);
  t.ties = m & 7u; t.valid = 1;
  e.k = w.k[slot]; e.flags = (m >> 3) & 7u;
}

// raster_balanced over the lean store (same walk, same arithmetic)
__device__ __forceinline__ void raster_balanced_lite(const WarpTrisLite& w, int ntris, int total, const RenderTargets& rt,
                                                     int W, int lane) {
  const int ch = (total + 31) >> 5;
  int item = lane * ch;
  const int end = min(item + ch, total);
  if (item >= end) return;
  int lo = 0, hi = ntris;  // largest j with prefix[j] <= item
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (w.prefix[mid] <= item) lo = mid; else hi = mid;
  }
  int j = lo;
  while (item < end) {
    TriSetup t;
    Emit e;
    tris_load_lite(w, j, t, e);
    const int first = w.prefix[j], cnt = w.prefix[j + 1] - first;
    const int local = item - first;
    const int wd = t.ni + 1;
    int dj = local / wd, di = local - dj * wd;
    int n_here = min(end - item, cnt - local);
    const bool tieA = t.ties & 1u, tieB = t.ties & 2u, tieC = t.ties & 4u;
    long long rowA = t.eA + (long long)dj * t.syA, rowB = t.eB + (long long)dj * t.syB, rowC = t.eC + (long long)dj * t.syC;
    long long wA = rowA + (long long)di * t.sxA, wB = rowB + (long long)di * t.sxB, wC = rowC + (long long)di * t.sxC;
    for (int q = 0; q < n_here; ++q) {
      bool in = (wA > 0 || (wA == 0 && tieA)) && (wB > 0 || (wB == 0 && tieB)) && (wC > 0 || (wC == 0 && tieC));
      if (in) tri_fragment(t, e, rt, W, wB, wC, di, dj);
      if (++di > t.ni) {
        di = 0;
        ++dj;
        rowA += t.syA; rowB += t.syB; rowC += t.syC;
        wA = rowA; wB = rowB; wC = rowC;
      } else {
        wA += t.sxA; wB += t.sxB; wC += t.sxC;
      }
    }
    item += n_here;
    ++j;
  }
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(kRenderThreads, MIN_BLOCKS)
    k_render_scatter_lite(KParams kp, SurfelPlanes s, const uint32_t* __restrict__ n_dev, const float* __restrict__ Mtab,
                          float conf_thr, int t_thr, int emit_old, int emit_new, int lequal, RenderTargets rt) {
  __shared__ WarpTrisLite s_tris[kRenderThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpTrisLite& wt = s_tris[warp];
  wt.prefix[2 * lane + 1] = 0;
  wt.prefix[2 * lane + 2] = 0;
  if (lane == 0) wt.prefix[0] = 0;
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  bool alive = k < *n_dev;
  float4 p0, p1, p2;
  bool is_old = false, is_new = false;
  if (alive) {
    p2 = __ldg(s.p2 + k);
    p1 = __ldg(s.p1 + k);
    alive = !kp.use_stability || p1.w > conf_thr;  // .geom:87 (cheap test first)
  }
  if (alive) {
    int creation = (int)p2.w, ts = (int)__float_as_uint(p2.x);
    is_old = emit_old && (creation < t_thr);                  // .geom:90
    is_new = emit_new && (creation >= t_thr || ts >= t_thr);  // .geom:91
    alive = is_old || is_new;
  }
  TriSetup t0, t1;
  t0.valid = t1.valid = 0;
  if (alive) {
    p0 = __ldg(s.p0 + k);
    float M[16];
    load_mat(Mtab, pose_index(p2.w), M);
    V3 pp = xform_point(M, mk3(p0.x, p0.y, p0.z));
    V3 nn = xform_dir(M, mk3(p1.x, p1.y, p1.z));
    float r = p0.w;
    bool visible = dot3(nn, divs3(neg3(pp), len3(pp))) > 0.01f;  // .geom:84
    float cx, cy, cz;
    project01(pp, kp.m_fov_up, kp.m_fov, kp.m_min_depth, kp.m_max_depth, cx, cy, cz);
    if (visible && cx >= 0.0f && cy >= 0.0f && cz >= 0.0f && cx < 1.0f && cy < 1.0f && cz < 1.0f) {
      V3 u = normalize3(mk3(nn.y - nn.z, -nn.x, nn.x));
      V3 v = normalize3(cross3(nn, u));
      V3 ru = scale3(r, u), rv = scale3(r, v);
      V3 corner[4];
      corner[0] = sub3(sub3(pp, ru), rv);
      corner[1] = sub3(add3(pp, ru), rv);
      corner[2] = add3(sub3(pp, ru), rv);
      corner[3] = add3(add3(pp, ru), rv);
      RVert q[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x, y, z;
        project01(corner[i], kp.m_fov_up, kp.m_fov, kp.m_min_depth, kp.m_max_depth, x, y, z);
        if (cx - x > 0.5f) x += 1.0f;  // .geom:68
        if (x - cx > 0.5f) x -= 1.0f;  // .geom:69
        float xw = (0.5f * (2.0f * x - 1.0f) + 0.5f) * (float)kp.Wm;
        float yw = (0.5f * (2.0f * y - 1.0f) + 0.5f) * (float)kp.Hm;
        q[i].z = 0.5f * (2.0f * z - 1.0f) + 0.5f;
        q[i].X = __float2ll_rn(xw * 256.0f);
        q[i].Y = __float2ll_rn(yw * 256.0f);
        q[i].tx = (i & 1) ? 1.0f : -1.0f;
        q[i].ty = (i & 2) ? 1.0f : -1.0f;
      }
      t0 = tri_prepare(q[0], q[1], q[2], kp.Wm, kp.Hm);
      t1 = tri_prepare(q[1], q[2], q[3], kp.Wm, kp.Hm);
    }
  }
  const int cnt = (int)t0.valid + (int)t1.valid;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int ntris = __shfl_sync(0xffffffffu, incl, 31);
  if (ntris == 0) return;
  __syncwarp();
  if (cnt) {
    Emit e;
    e.k = k;
    e.flags = (is_old ? 1u : 0u) | (is_new ? 2u : 0u) | (lequal ? 4u : 0u);
    int slot = incl - cnt;
    if (t0.valid) tris_store_lite(wt, slot++, t0, e);
    if (t1.valid) tris_store_lite(wt, slot, t1, e);
  }
  __syncwarp();
  int a = wt.prefix[2 * lane + 1], b = wt.prefix[2 * lane + 2];
  int sum = a + b, isum = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, isum, o);
    if (lane >= o) isum += v;
  }
  const int total = __shfl_sync(0xffffffffu, isum, 31);
  __syncwarp();
  wt.prefix[2 * lane + 1] = isum - sum + a;
  wt.prefix[2 * lane + 2] = isum;
  __syncwarp();
  raster_balanced_lite(wt, ntris, total, rt, kp.Wm, lane);
}

static int render_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SUMA_B200_RENDER_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

void launch_render_scatter(const Launch& L, const KParams& kp, SurfelPlanes s, const uint32_t* n_dev, uint32_t n_upper,
                           const float* M, float conf_thr, int t_thr, int emit_old, int emit_new, int lequal,
                           RenderTargets t) {
  if (n_upper == 0) return;
  {
    ScopedKernel sk(L, K_RENDER_SCATTER);
    const unsigned grid = (n_upper + kRenderThreads - 1) / kRenderThreads;
    const int variant = render_variant();
    if (variant == 1) {  // lean triangle store, up to 6 blocks per SM without spills
      k_render_scatter_lite<6><<<grid, kRenderThreads, 0, L.stream>>>(kp, s, n_dev, M, conf_thr, t_thr, emit_old,
                                                                      emit_new, lequal, t);
    } else if (variant == 2) {  // lean triangle store, 8 blocks per SM (64 registers, a few spills)
      k_render_scatter_lite<8><<<grid, kRenderThreads, 0, L.stream>>>(kp, s, n_dev, M, conf_thr, t_thr, emit_old,
                                                                      emit_new, lequal, t);
    } else {
      k_render_scatter<<<grid, kRenderThreads, 0, L.stream>>>(kp, s, n_dev, M, conf_thr, t_thr, emit_old, emit_new,
                                                              lequal, t);
    }
  }
}

struct Px {
  float4 v, n, s;
};

__device__ __forceinline__ Px resolve_px(const SurfelPlanes& s, unsigned long long key, const float* __restrict__ M_old,
                                         const float* __restrict__ M_new, int fixed_pass, int lequal) {
  Px o;
  o.v = o.n = o.s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (key == ~0ull) return o;
  uint32_t low = (uint32_t)(key & 0xffffffffull);
  int pass = (int)((key >> 32) & 1ull);
  uint32_t k = low;
  if (lequal && fixed_pass < 0) {
    k = 0xffffffffu - low;
    pass = 1 - pass;
  }
  if (fixed_pass >= 0) pass = fixed_pass;
  float4 p0 = __ldg(s.p0 + k), p1 = __ldg(s.p1 + k), p2 = __ldg(s.p2 + k);
  float M[16];
  load_mat(pass ? M_new : M_old, pose_index(p2.w), M);
  V3 pp = xform_point(M, mk3(p0.x, p0.y, p0.z));
  V3 nn = xform_dir(M, mk3(p1.x, p1.y, p1.z));
  o.v = make_float4(pp.x, pp.y, pp.z, 1.0f);  // render_surfels.frag:30, .geom:95-97 (flat: surfel centre)
  o.n = make_float4(nn.x, nn.y, nn.z, 1.0f);
  o.s = __ldg(s.p3 + k);
  return o;
}

// resolve the key images into vertex / normal / semantic maps; optionally fuses render_compose.frag:26-48
__global__ void __launch_bounds__(kThreads) k_render_resolve(KParams kp, SurfelPlanes s, const float* __restrict__ M_old,
                                                            const float* __restrict__ M_new, RenderTargets t,
                                                            FrameDev f_old, FrameDev f_new, FrameDev f_comp,
                                                            FrameDev f_out, int keep_semantic, int lequal) {
  int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= kp.Wm * kp.Hm) return;
  // every key image that is read here is cleared again ("consume and clear"): no separate fill before the next view
  Px po, pn;
  unsigned long long ko_raw = ~0ull, kn_raw = ~0ull;
  if (t.key_old) {
    ko_raw = t.key_old[pix];
    t.key_old[pix] = ~0ull;
    po = resolve_px(s, ko_raw, M_old, M_new, 0, 0);
    f_old.vertex[pix] = po.v;
    f_old.normal[pix] = po.n;
    if (!keep_semantic) f_old.semantic[pix] = po.s;
  }
  if (t.key_new) {
    kn_raw = t.key_new[pix];
    t.key_new[pix] = ~0ull;
    pn = resolve_px(s, kn_raw, M_old, M_new, 1, 0);
    f_new.vertex[pix] = pn.v;
    f_new.normal[pix] = pn.n;
    if (!keep_semantic) f_new.semantic[pix] = pn.s;
  }
  if (f_comp.vertex && !t.key_comp && t.key_old && t.key_new) {
    // GL_LESS, old pass drawn before the new pass on an uncleared buffer (SurfelMap.cpp:893-906): the winner of the
    // union of both passes is the smaller of the two per-pass winners, depth first, then pass, then surfel index
    unsigned long long ko = ko_raw, kn = kn_raw;
    if (kn != ~0ull) kn |= (1ull << 32);
    Px pc = resolve_px(s, ko < kn ? ko : kn, M_old, M_new, -1, 0);
    f_comp.vertex[pix] = pc.v;
    f_comp.normal[pix] = pc.n;
    if (!keep_semantic) f_comp.semantic[pix] = pc.s;
  }
  if (t.key_comp) {
    unsigned long long kc = t.key_comp[pix];
    t.key_comp[pix] = ~0ull;
    Px pc = resolve_px(s, kc, M_old, M_new, -1, lequal);
    f_comp.vertex[pix] = pc.v;
    f_comp.normal[pix] = pc.n;
    if (!keep_semantic) f_comp.semantic[pix] = pc.s;
  }
  if (f_out.vertex && t.key_old && t.key_new) {  // render_compose.frag
    bool valid = po.v.w > 0.5f && po.n.w > 0.5f;
    bool new_valid = pn.v.w > 0.5f && pn.n.w > 0.5f;
    Px o = pn;
    if (!new_valid && valid &&
        (pn.v.w < 0.5f ||
         len3(sub3(mk3(pn.v.x, pn.v.y, pn.v.z), mk3(po.v.x, po.v.y, po.v.z))) < kp.max_loop_closure_distance))
      o = po;
    f_out.vertex[pix] = o.v;
    f_out.normal[pix] = o.n;
    f_out.semantic[pix] = o.s;
  }
}

void launch_render_resolve(const Launch& L, const KParams& kp, SurfelPlanes s, const float* M_old, const float* M_new,
                           RenderTargets t, FrameDev f_old, FrameDev f_new, FrameDev f_comp, FrameDev f_out,
                           int keep_semantic, int lequal) {
  int P = kp.Wm * kp.Hm;
  {
    ScopedKernel sk(L, K_RENDER_RESOLVE);
    k_render_resolve<<<(P + kThreads - 1) / kThreads, kThreads, 0, L.stream>>>(kp, s, M_old, M_new, t, f_old, f_new,
                                                                            f_comp, f_out, keep_semantic, lequal);
  }
}

// ------------------------------------------------------------------------------------------------------------
// K6a: gen_indexmap.vert:62-81 -- nearest visible surfel per data pixel
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_index_scatter(KParams kp, SurfelPlanes s, const uint32_t* __restrict__ n_dev,
                                                           const float* __restrict__ Mtab,
                                                           unsigned long long* __restrict__ keys) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= *n_dev) return;
  float4 p0 = __ldg(s.p0 + k), p1 = __ldg(s.p1 + k), p2 = __ldg(s.p2 + k);
  float M[16];
  load_mat(Mtab, pose_index(p2.w), M);
  V3 v = xform_point(M, mk3(p0.x, p0.y, p0.z));
  V3 n = xform_dir(M, mk3(p1.x, p1.y, p1.z));
  if (!(dot3(n, divs3(neg3(v), len3(v))) > 0.01f)) return;
  float x, y, z;
  project01(v, kp.fov_up, kp.fov, kp.min_depth, kp.max_depth, x, y, z);
  float fx = floorf(x * (float)kp.W), fy = floorf(y * (float)kp.H);
  if (!(fx >= 0.0f && fx < (float)kp.W && fy >= 0.0f && fy < (float)kp.H)) return;
  float zn = 2.0f * z - 1.0f;
  if (!(zn >= -1.0f && zn <= 1.0f)) return;
  uint32_t d24 = depth24(0.5f * zn + 0.5f);
  if (d24 >= kDepthClear) return;
  size_t pix = (size_t)(int)fy * kp.W + (size_t)(int)fx;
  unsigned long long key = ((unsigned long long)d24 << 32) | (unsigned long long)k;
  if (key < keys[pix]) atomicMin(keys + pix, key);
}

void launch_index_scatter(const Launch& L, const KParams& kp, SurfelPlanes s, const uint32_t* n_dev, uint32_t n_upper,
                          const float* M, unsigned long long* keys) {
  if (n_upper == 0) return;
  {
    ScopedKernel sk(L, K_INDEX_SCATTER);
    k_index_scatter<<<(n_upper + kThreads - 1) / kThreads, kThreads, 0, L.stream>>>(kp, s, n_dev, M, keys);
  }
}

// K6b: init_radiusConf.vert:41-68
// (+ re-arms the index-map key image and clears the "measurement integrated" flags of the pixel: the prologue of a
// stand-alone map update; in the per-scan pipeline the tiled preprocessing pass has already done all three)
__global__ void __launch_bounds__(kThreads) k_radius(KParams kp, FrameDev f, float4* __restrict__ radius_map,
                                                    unsigned long long* __restrict__ index_keys,
                                                    uint8_t* __restrict__ integrated) {
  int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= kp.W * kp.H) return;
  index_keys[pix] = ~0ull;
  integrated[pix] = 0;
  float4 V = __ldg(f.vertex + pix), N = __ldg(f.normal + pix);
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  V3 v = mk3(V.x, V.y, V.z), n = mk3(N.x, N.y, N.z);
  float d = len3(v);
  V3 view_dir = divs3(neg3(v), d);
  float angle = dot3(n, view_dir);
  if (V.w > 0.5f && N.w > 0.5f && angle > kp.radconf_angle_thresh) {
    float c = angle < 0.5f ? 0.5f : (angle > 1.0f ? 1.0f : angle);
    float radius = ((1.41f * d) * kp.pixel_size) / c;
    float lo = radius > kp.min_radius ? radius : kp.min_radius;
    radius = lo < kp.max_radius ? lo : kp.max_radius;
    o.x = radius;
    o.w = 1.0f;  // Q3: the confidence channel stays 0
  }
  radius_map[pix] = o;
}

void launch_radius(const Launch& L, const KParams& kp, FrameDev frame, float4* radius_map, unsigned long long* index_keys,
                   uint8_t* integrated) {
  int P = kp.W * kp.H;
  {
    ScopedKernel sk(L, K_RADIUS);
    k_radius<<<(P + kThreads - 1) / kThreads, kThreads, 0, L.stream>>>(kp, frame, radius_map, index_keys, integrated);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Single-pass ordered compaction (transform-feedback semantics) with decoupled look-back: every block publishes the
// number of items it keeps, obtains the number kept by all earlier blocks by walking back over the published
// descriptors, and writes its items to their final position -- one read and one write per surfel, no flag / count /
// scan passes. Blocks take their (virtual) index from a ticket so that earlier indices are always already running.
// Descriptor: [63:34] launch generation, [33:32] status (0 none, 1 block aggregate, 2 inclusive prefix), [31:0] value.
// ------------------------------------------------------------------------------------------------------------
struct Lookback {
  unsigned long long* desc;
  uint32_t* ticket;
  uint32_t gen;
};
__device__ __forceinline__ unsigned long long lb_pack(uint32_t gen, uint32_t status, uint32_t value) {
  return ((unsigned long long)gen << 34) | ((unsigned long long)status << 32) | (unsigned long long)value;
}

__device__ __forceinline__ uint32_t block_ticket(const Lookback& lb) {
  __shared__ uint32_t s_vb;
  if (threadIdx.x == 0) s_vb = atomicAdd(lb.ticket, 1u);
  __syncthreads();
  return s_vb;
}

// rank of this thread's item among the kept items of the block, and the block's count
template <int THREADS>
__device__ __forceinline__ uint32_t block_rank(bool kept, uint32_t& block_count) {
  __shared__ uint32_t s_warp[THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned m = __ballot_sync(0xffffffffu, kept);
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  uint32_t before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) {
    uint32_t v = s_warp[w];
    if (w < warp) before += v;
    total += v;
  }
  block_count = total;
  return before + __popc(m & ((1u << lane) - 1u));
}

// exclusive prefix of `count` over the blocks with a smaller virtual index (warp 0 walks back 32 descriptors at a time)
__device__ __forceinline__ uint32_t lookback_exclusive(const Lookback& lb, uint32_t vb, uint32_t count) {
  __shared__ uint32_t s_excl;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    if (vb == 0) {
      if (lane == 0) {
        s_excl = 0;
        atomicExch(lb.desc + 0, lb_pack(lb.gen, 2u, count));
      }
    } else {
      if (lane == 0) atomicExch(lb.desc + vb, lb_pack(lb.gen, 1u, count));
      uint32_t sum = 0;
      long long base = (long long)vb - 1;
      for (;;) {
        long long idx = base - lane;
        uint32_t status = 2u, value = 0;  // virtual entries before block 0: prefix 0
        if (idx >= 0) {
          unsigned long long d;
          for (;;) {
            d = *(volatile unsigned long long*)(lb.desc + idx);
            if ((uint32_t)(d >> 34) == lb.gen && ((d >> 32) & 3ull) != 0ull) break;
            __nanosleep(200);  // polling from ~1000 resident blocks would otherwise eat a large share of L2 bandwidth
          }
          status = (uint32_t)((d >> 32) & 3ull);
          value = (uint32_t)d;
        }
        unsigned pm = __ballot_sync(0xffffffffu, status == 2u);
        int first = pm ? (__ffs(pm) - 1) : 32;  // closest inclusive prefix in this window
        uint32_t contrib = (lane <= first) ? value : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        sum += contrib;
        if (pm) break;
        base -= 32;
      }
      if (lane == 0) {
        s_excl = sum;
        __threadfence();
        atomicExch(lb.desc + vb, lb_pack(lb.gen, 2u, sum + count));
      }
    }
  }
  __syncthreads();
  return s_excl;
}

// ------------------------------------------------------------------------------------------------------------
// K6c: update_surfels.vert:140-333, fused with the K6e predicate of copy_surfels.vert:38-56
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 data_tex(const float4* __restrict__ img, int W, int H, float fx, float fy) {
  if (!(fx >= 0.0f && fx < (float)W && fy >= 0.0f && fy < (float)H)) return make_float4(0.f, 0.f, 0.f, 0.f);
  return __ldg(img + (size_t)(int)fy * W + (size_t)(int)fx);
}

__device__ __forceinline__ V3 slerp(V3 v0, V3 v1, float weight) {  // update_surfels.vert:113-124
  float omega = acosf_(dot3(normalize3(v0), normalize3(v1)));
  float eta = 1.0f / sinf_(omega);
  float w0 = eta * sinf_(weight * omega);
  float w1 = eta * sinf_((1.0f - weight) * omega);
  return add3(scale3(w0, v0), scale3(w1, v1));
}

__device__ __forceinline__ bool submap_keep(const float* __restrict__ poses, float4 p0, float4 p2, float2 center,
                                            float extent) {
  float M[16];
  load_mat(poses, pose_index(p2.w), M);
  V3 w = xform_point(M, mk3(p0.x, p0.y, p0.z));
  if ((int)__float_as_uint(p2.x) < 0 || fabsf(w.x - center.x) > extent || fabsf(w.y - center.y) > extent) return false;
  return true;
}

struct UpdateArgs {
  Mat4 pose;               // current sensor pose (by value: no upload on the critical path)
  Mat4 inv_pose;
  const Mat4* pose_dev;    // or resident on the device (pipeline mode)
  const Mat4* inv_pose_dev;
  const float* poses;      // pose table
  const float* poses_inv;  // inverse pose table
  const unsigned long long* index_keys;
  const float4* radius_map;
  int timestamp;
  float2 submap_center;
  float submap_extent;
};

__global__ void __launch_bounds__(kThreads) k_update_surfels(KParams kp, SurfelPlanes src, SurfelPlanes tmp,
                                                            const uint32_t* __restrict__ n_dev, UpdateArgs ua,
                                                            FrameDev f, uint8_t* __restrict__ integrated,
                                                            uint8_t* __restrict__ keep,
                                                            uint32_t* __restrict__ block_counts) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  bool kept = false;
  if (k < *n_dev) {
    const int W = kp.W, H = kp.H;
    float4 p0 = __ldg(src.p0 + k), p1 = __ldg(src.p1 + k), p2 = __ldg(src.p2 + k), p3 = __ldg(src.p3 + k);
    const int timestamp = ua.timestamp;
    int s_ts = (int)__float_as_uint(p2.x);
    int surfel_age = timestamp - s_ts;
    int creation = (int)p2.w;
    int ci = pose_index(p2.w);
    float SP[16];
    load_mat(ua.poses, ci, SP);
    const float* POSE = ua.pose_dev ? ua.pose_dev->m : ua.pose.m;
    const float* INV = ua.inv_pose_dev ? ua.inv_pose_dev->m : ua.inv_pose.m;
    V3 old_position = xform_point(SP, mk3(p0.x, p0.y, p0.z));
    V3 old_normal = xform_dir(SP, mk3(p1.x, p1.y, p1.z));
    float old_radius = p0.w, old_conf = p1.w, old_weight = p2.z;
    bool valid_out = true;
    if (old_conf < kp.confidence_threshold && kp.use_stability) valid_out = surfel_age < kp.unstable_age;
    float4 o0 = p0, o1 = p1, o2 = p2, o3 = p3;
    o2.y = pack_rgb(0.3f, 0.3f, 0.3f);
    V3 vertex = xform_point(INV, old_position);
    V3 normal = normalize3(xform_dir(INV, old_normal));
    bool visible = dot3(normal, divs3(neg3(vertex), len3(vertex))) > 0.0f;
    float x, y, z;
    project01(vertex, kp.fov_up, kp.fov, kp.min_depth, kp.max_depth, x, y, z);
    float ix = floorf(x * (float)W) + 0.5f, iy = floorf(y * (float)H) + 0.5f;
    float4 Vt = data_tex(f.vertex, W, H, ix, iy);
    float4 Nt = data_tex(f.normal, W, H, ix, iy);
    bool valid = (Vt.w > 0.5f) && (Nt.w > 0.5f);
    bool inside = (ix < (float)W && iy < (float)H && z < 1.0f) && !(ix < 0.0f && iy < 0.0f && z < 0.0f);
    float penalty = 0.0f;
    float update_conf = kp.log_prior;
    size_t dpix = 0;
    bool mark_integrated = false;
    if (valid && inside && visible) {
      float4 St = data_tex(f.semantic, W, H, ix, iy);
      float4 Rt = data_tex(ua.radius_map, W, H, ix, iy);
      float data_label = St.x * 255.0f, data_prob = St.w;
      float model_label = p3.x * 255.0f, model_prob = p3.w;
      bool label_diff = roundf_(data_label) != roundf_(model_label);
      if (label_diff && is_movable(model_label)) penalty = 1.0f;
      V3 v = mk3(Vt.x, Vt.y, Vt.z), n = mk3(Nt.x, Nt.y, Nt.z);
      V3 v_global = xform_point(POSE, v);
      V3 n_global = normalize3(xform_dir(POSE, n));
      V3 view_dir = divs3(neg3(v), len3(v));
      float distance = fabsf(dot3(old_normal, sub3(v_global, old_position)));
      float angle = len3(cross3(n_global, old_normal));
      float new_radius = Rt.x, new_conf = Rt.y;
      dpix = (size_t)(int)iy * W + (size_t)(int)ix;
      if (distance < kp.map_max_distance && angle < kp.update_angle_thresh) {
        // the "measurement integrated" point is drawn only if the geometry stage emits this surfel, i.e. if it is
        // still valid at the end of the pass (update_surfels.geom:35-48): decided below
        float zn = 2.0f * z - 1.0f;
        mark_integrated = zn >= -1.0f && zn <= 1.0f && depth24(0.5f * zn + 0.5f) < kDepthClear;
        float confidence = old_conf + new_conf;
        o1.w = confidence;
        o2.x = __uint_as_float((uint32_t)timestamp);
        float avg_radius = new_radius < old_radius ? new_radius : old_radius;
        avg_radius = avg_radius > 0.0f ? avg_radius : 0.0f;  // update program's min_radius uniform is 0 (SurfelMap.cpp:422)
        o0.w = avg_radius;
        valid_out = true;
        o2.y = pack_rgb(0.0f, 0.7f, 0.0f);
        o2.w = (float)creation;
        float a = angle, d = distance;
        float pr = kp.p_stable;
        if (kp.confidence_mode == 1 || kp.confidence_mode == 3)
          pr = pr * expf_(((-a) * a) / (kp.sigma_angle * kp.sigma_angle));
        if (kp.confidence_mode == 2 || kp.confidence_mode == 3)
          pr = pr * expf_(((-d) * d) / (kp.sigma_distance * kp.sigma_distance));
        pr = pr > kp.p_unstable ? pr : kp.p_unstable;
        pr = pr < 1.0f ? pr : 1.0f;
        update_conf = logf_(pr / (1.0f - pr));
        if ((new_radius < old_radius && timestamp - creation < kp.active_timestamps) || kp.update_always) {
          float w1 = 0.9f, w2 = 0.1f;
          if (kp.weighting_scheme > 0) {
            w1 = old_weight;
            w2 = 1.0f;
            if (kp.weighting_scheme == 2) w2 = dot3(n, view_dir);
            float sw = w1 + w2;
            o2.z = kp.max_weight < sw ? kp.max_weight : sw;
            float sum = w1 + w2;
            w1 = w1 / sum;
            w2 = w2 / sum;
          }
          V3 avg_position = add3(scale3(w1, old_position), scale3(w2, v_global));
          V3 avg_normal = slerp(old_normal, n_global, w1);
          float avg_prob;
          if (label_diff)
            avg_prob = w1 * model_prob + w2 * (1.0f - data_prob);
          else
            avg_prob = w1 * model_prob + w2 * data_prob;
          o3.w = avg_prob;
          if (kp.averaging_scheme == 1) {
            avg_position = add3(old_position, scale3(w2 * distance, old_normal));
            avg_normal = slerp(old_normal, n_global, w1);
          }
          avg_normal = normalize3(avg_normal);
          float SPI[16];
          load_mat(ua.poses_inv, ci, SPI);
          avg_position = xform_point(SPI, avg_position);
          avg_normal = xform_dir(SPI, avg_normal);
          o0 = make_float4(avg_position.x, avg_position.y, avg_position.z, avg_radius);
          o1 = make_float4(avg_normal.x, avg_normal.y, avg_normal.z, confidence);
          o2.y = pack_rgb(1.0f, 0.0f, 1.0f);
        }
      } else {
        unsigned long long key = ua.index_keys[dpix];
        int idx = (key == ~0ull) ? -1 : (int)(uint32_t)(key & 0xffffffffull);
        if (idx == (int)k) {  // closest visible surfel of that pixel
          update_conf = logf_(kp.p_unstable / (1.0f - kp.p_unstable));
          o2.y = pack_rgb(0.0f, 1.0f, 1.0f);
        }
      }
    }
    update_conf = update_conf - penalty;
    if (kp.use_stability) {
      float c = (old_conf + update_conf) - kp.log_prior;
      o1.w = c < 20.0f ? c : 20.0f;
    } else {
      o1.w = old_conf;
    }
    if (o1.w < kp.log_unstable && kp.use_stability) valid_out = false;
    if (valid_out && mark_integrated) integrated[dpix] = 1;
    kept = valid_out && submap_keep(ua.poses, o0, o2, ua.submap_center, ua.submap_extent);
    tmp.p0[k] = o0;
    tmp.p1[k] = o1;
    tmp.p2[k] = o2;
    tmp.p3[k] = o3;
    keep[k] = kept ? 1 : 0;
  }
  int cnt = __syncthreads_count(kept ? 1 : 0);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)cnt;
}

// K6d: gen_surfels.vert:38-52 + .geom:109-145; thread t <-> pixel (x = t / H, y = t % H): x-major order
// (SurfelMap.cpp:88-92), fused with the K6e predicate
__global__ void __launch_bounds__(kThreads) k_gen_surfels(KParams kp, FrameDev f, const float4* __restrict__ radius_map,
                                                         const uint8_t* __restrict__ integrated,
                                                         const float* __restrict__ poses, int timestamp,
                                                         float2 center, float extent, SurfelPlanes tmp,
                                                         uint8_t* __restrict__ keep,
                                                         uint32_t* __restrict__ block_counts) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  bool kept = false;
  if (t < kp.W * kp.H) {
    int x = t / kp.H, y = t - x * kp.H;
    size_t pix = (size_t)y * kp.W + x;
    float4 V = __ldg(f.vertex + pix), N = __ldg(f.normal + pix), R = __ldg(radius_map + pix);
    bool invalid = (V.w < 1.0f) || (N.w < 1.0f);
    invalid = invalid || (R.w < 0.5f);
    bool integ = integrated[pix] != 0;
    V3 v = mk3(V.x, V.y, V.z), n = mk3(N.x, N.y, N.z);
    V3 view_dir = divs3(neg3(v), len3(v));
    bool valid = !invalid && !integ && (dot3(n, view_dir) > 0.01f);
    if (valid) {
      V3 ng = normalize3(n);
      float4 S = __ldg(f.semantic + pix);
      float conf = is_movable(S.x * 255.0f) ? kp.log_prior - 0.5f : kp.log_prior;
      float4 o0 = make_float4(v.x, v.y, v.z, R.x);
      float4 o1 = make_float4(ng.x, ng.y, ng.z, conf);
      float4 o2 = make_float4(__uint_as_float((uint32_t)timestamp), pack_rgb(0.0f, 0.0f, 1.0f), 1.0f, (float)timestamp);
      kept = submap_keep(poses, o0, o2, center, extent);
      tmp.p0[t] = o0;
      tmp.p1[t] = o1;
      tmp.p2[t] = o2;
      tmp.p3[t] = S;
    }
    keep[t] = kept ? 1 : 0;
  }
  int cnt = __syncthreads_count(kept ? 1 : 0);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)cnt;
}

// K6d + K6e in ONE pass: new surfels are appended in x-major pixel order behind the updated ones
__global__ void __launch_bounds__(kThreads) k_gen_compact(KParams kp, FrameDev f, const float4* __restrict__ radius_map,
                                                         const uint8_t* __restrict__ integrated,
                                                         const float* __restrict__ poses, int timestamp, float2 center,
                                                         float extent, SurfelPlanes dst, Lookback lb, uint32_t cap,
                                                         uint32_t* __restrict__ counts) {
  const uint32_t vb = block_ticket(lb);
  int t = (int)(vb * blockDim.x + threadIdx.x);
  bool kept = false;
  float4 o0, o1, o2, o3;
  o0 = o1 = o2 = o3 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < kp.W * kp.H) {
    int x = t / kp.H, y = t - x * kp.H;
    size_t pix = (size_t)y * kp.W + x;
    float4 V = __ldg(f.vertex + pix), N = __ldg(f.normal + pix), R = __ldg(radius_map + pix);
    bool invalid = (V.w < 1.0f) || (N.w < 1.0f);
    invalid = invalid || (R.w < 0.5f);
    bool integ = integrated[pix] != 0;
    V3 v = mk3(V.x, V.y, V.z), n = mk3(N.x, N.y, N.z);
    V3 view_dir = divs3(neg3(v), len3(v));
    bool valid = !invalid && !integ && (dot3(n, view_dir) > 0.01f);
    if (valid) {
      V3 ng = normalize3(n);
      float4 S = __ldg(f.semantic + pix);
      float conf = is_movable(S.x * 255.0f) ? kp.log_prior - 0.5f : kp.log_prior;
      o0 = make_float4(v.x, v.y, v.z, R.x);
      o1 = make_float4(ng.x, ng.y, ng.z, conf);
      o2 = make_float4(__uint_as_float((uint32_t)timestamp), pack_rgb(0.0f, 0.0f, 1.0f), 1.0f, (float)timestamp);
      o3 = S;
      kept = submap_keep(poses, o0, o2, center, extent);
    }
  }
  uint32_t block_count;
  const uint32_t rank = block_rank<kThreads>(kept, block_count);
  const uint32_t excl = lookback_exclusive(lb, vb, block_count);
  const uint32_t base = counts[1];
  if (kept) {
    uint32_t d = base + excl + rank;
    if (d < cap) {
      dst.p0[d] = o0;
      dst.p1[d] = o1;
      dst.p2[d] = o2;
      dst.p3[d] = o3;
    }
  }
  if (vb == gridDim.x - 1 && threadIdx.x == 0) {
    uint32_t total = base + excl + block_count;
    counts[0] = total < cap ? total : cap;  // new size of the map
    counts[3] = excl + block_count;         // kept new surfels
    *lb.ticket = 0;
  }
}

// extract_surfels.vert:44-62: flags of the surfels inside one submap tile
__global__ void __launch_bounds__(kThreads) k_extract_flags(SurfelPlanes s, const uint32_t* __restrict__ n_dev,
                                                           const float* __restrict__ poses, float2 center, float extent,
                                                           uint8_t* __restrict__ keep,
                                                           uint32_t* __restrict__ block_counts) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  bool kept = false;
  if (k < *n_dev) {
    float4 p0 = __ldg(s.p0 + k), p2 = __ldg(s.p2 + k);
    float M[16];
    load_mat(poses, pose_index(p2.w), M);
    V3 w = xform_point(M, mk3(p0.x, p0.y, p0.z));
    kept = !(fabsf(w.x - center.x) > extent || fabsf(w.y - center.y) > extent);
    keep[k] = kept ? 1 : 0;
  }
  int cnt = __syncthreads_count(kept ? 1 : 0);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)cnt;
}

// ------------------------------------------------------------------------------------------------------------
// ordered compaction (transform feedback semantics): block_counts[b] hold the kept items of block b (256 items)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_scan_blocks(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
                                                      const uint32_t* __restrict__ n_dev, uint32_t n_fixed,
                                                      const uint32_t* __restrict__ base_dev, uint32_t cap,
                                                      uint32_t* __restrict__ count_out,
                                                      uint32_t* __restrict__ kept_out) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t s_total;
  uint32_t n = n_dev ? *n_dev : n_fixed;
  uint32_t nblocks = (n + kThreads - 1) / kThreads;
  uint32_t per = (nblocks + 1023) / 1024;
  uint32_t b0 = threadIdx.x * per, b1 = min(b0 + per, nblocks);
  uint32_t local = 0;
  for (uint32_t b = b0; b < b1; ++b) local += counts[b];
  // block-wide exclusive scan of `local`
  uint32_t v = local;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += u;
    }
    warp_tot[lane] = w;  // inclusive
    if (lane == 31) s_total = w;
  }
  __syncthreads();
  uint32_t excl = (v - local) + (warp > 0 ? warp_tot[warp - 1] : 0u);
  uint32_t run = excl;
  for (uint32_t b = b0; b < b1; ++b) {
    offsets[b] = run;
    run += counts[b];
  }
  if (threadIdx.x == 0) {
    uint32_t base = base_dev ? *base_dev : 0u;
    uint32_t tot = base + s_total;
    if (kept_out) *kept_out = s_total;
    *count_out = tot < cap ? tot : cap;
  }
}

__global__ void __launch_bounds__(kThreads) k_compact_scatter(SurfelPlanes src, const uint8_t* __restrict__ keep,
                                                             const uint32_t* __restrict__ offsets,
                                                             const uint32_t* __restrict__ n_dev, uint32_t n_fixed,
                                                             SurfelPlanes dst, const uint32_t* __restrict__ base_dev,
                                                             uint32_t cap, uint32_t rel_cap) {
  __shared__ uint32_t warp_cnt[kThreads / 32];
  uint32_t n = n_dev ? *n_dev : n_fixed;
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  bool kept = (k < n) && keep[k];
  unsigned m = __ballot_sync(0xffffffffu, kept);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_cnt[warp] = __popc(m);
  __syncthreads();
  uint32_t before = 0;
  for (int w = 0; w < warp; ++w) before += warp_cnt[w];
  if (!kept) return;
  uint32_t rank = before + __popc(m & ((1u << lane) - 1u));
  uint32_t base = base_dev ? *base_dev : 0u;
  uint32_t rel = offsets[blockIdx.x] + rank;
  uint32_t d = base + rel;
  if (d >= cap || rel >= rel_cap) return;  // transform feedback drops what does not fit
  dst.p0[d] = src.p0[k];
  dst.p1[d] = src.p1[k];
  dst.p2[d] = src.p2[k];
  dst.p3[d] = src.p3[k];
}

void launch_compact(const Launch& L, SurfelPlanes src, const uint8_t* keep, const uint32_t* block_counts,
                    uint32_t* block_offsets, const uint32_t* n_dev, uint32_t n_upper, SurfelPlanes dst,
                    const uint32_t* base_dev, uint32_t cap, uint32_t* count_out, uint32_t* kept_out) {
  {
    ScopedKernel sk(L, K_SCAN_BLOCKS);
    k_scan_blocks<<<1, 1024, 0, L.stream>>>(block_counts, block_offsets, n_dev, n_upper, base_dev, cap, count_out,
                                          kept_out);
  }
  if (n_upper == 0) return;
  {
    ScopedKernel sk(L, K_COMPACT_SCATTER);
    k_compact_scatter<<<(n_upper + kThreads - 1) / kThreads, kThreads, 0, L.stream>>>(src, keep, block_offsets, n_dev,
                                                                                   n_upper, dst, base_dev, cap,
                                                                                   0xffffffffu);
  }
}

// ---- submap tile cache in HBM: a leaving tile is compacted straight into a device-resident pool (the reference
// downloads it to host RAM, SurfelMap.cpp:708-742). rec = {base, count} of the tile inside the pool.
__global__ void k_pool_alloc(uint2* __restrict__ rec, uint32_t* __restrict__ pool_top, uint32_t pool_cap) {
  uint32_t base = *pool_top, cnt = rec->y;
  if (base > pool_cap) base = pool_cap;
  if (cnt > pool_cap - base) cnt = pool_cap - base;  // pool exhausted: the tail of the tile is lost
  rec->x = base;
  rec->y = cnt;
  *pool_top = base + cnt;
}

void launch_extract_to_pool(const Launch& L, SurfelPlanes map, const uint32_t* n_dev, uint32_t n_upper, const float* poses,
                            float2 center, float extent, uint8_t* keep, uint32_t* block_counts, uint32_t* block_offsets,
                            SurfelPlanes pool, uint32_t pool_cap, uint32_t* pool_top, uint2* rec, uint32_t tile_cap) {
  launch_extract_flags(L, map, n_dev, n_upper, poses, center, extent, keep, block_counts);
  {
    ScopedKernel sk(L, K_SCAN_BLOCKS);
    k_scan_blocks<<<1, 1024, 0, L.stream>>>(block_counts, block_offsets, n_dev, n_upper, nullptr, tile_cap, &rec->y, nullptr);
  }
  {
    ScopedKernel sk(L, K_SCAN_BLOCKS);
    k_pool_alloc<<<1, 1, 0, L.stream>>>(rec, pool_top, pool_cap);
  }
  if (n_upper == 0) return;
  {
    ScopedKernel sk(L, K_COMPACT_SCATTER);
    k_compact_scatter<<<(n_upper + kThreads - 1) / kThreads, kThreads, 0, L.stream>>>(
        map, keep, block_offsets, n_dev, n_upper, pool, &rec->x, pool_cap, tile_cap);
  }
}

void launch_update_surfels(const Launch& L, const KParams& kp, SurfelPlanes src, SurfelPlanes tmp, const uint32_t* n_dev,
                           uint32_t n_upper, const Mat4& pose, const Mat4& inv_pose, const Mat4* pose_dev,
                           const Mat4* inv_pose_dev, const float* poses,
                           const float* poses_inv, const unsigned long long* index_keys, const float4* radius_map,
                           FrameDev frame, int timestamp, float2 submap_center, float submap_extent,
                           uint8_t* integrated, uint8_t* keep, uint32_t* block_counts) {
  if (n_upper == 0) return;
  UpdateArgs ua{pose, inv_pose, pose_dev, inv_pose_dev, poses, poses_inv, index_keys, radius_map, timestamp,
                submap_center, submap_extent};
  {
    ScopedKernel sk(L, K_UPDATE_SURFELS);
    k_update_surfels<<<(n_upper + kThreads - 1) / kThreads, kThreads, 0, L.stream>>>(kp, src, tmp, n_dev, ua, frame,
                                                                                  integrated, keep, block_counts);
  }
}

void launch_gen_surfels(const Launch& L, const KParams& kp, FrameDev frame, const float4* radius_map,
                        const uint8_t* integrated, const float* poses, int timestamp, float2 submap_center,
                        float submap_extent, SurfelPlanes tmp, uint8_t* keep, uint32_t* block_counts) {
  int P = kp.W * kp.H;
  {
    ScopedKernel sk(L, K_GEN_SURFELS);
    k_gen_surfels<<<(P + kThreads - 1) / kThreads, kThreads, 0, L.stream>>>(kp, frame, radius_map, integrated, poses,
                                                                         timestamp, submap_center, submap_extent, tmp,
                                                                         keep, block_counts);
  }
}

void launch_gen_compact(const Launch& L, const KParams& kp, FrameDev frame, const float4* radius_map,
                        const uint8_t* integrated, const float* poses, int timestamp, float2 submap_center,
                        float submap_extent, SurfelPlanes map, unsigned long long* desc, uint32_t* ticket, uint32_t gen,
                        uint32_t cap, uint32_t* counts) {
  int P = kp.W * kp.H;
  Lookback lb{desc, ticket, gen};
  {
    ScopedKernel sk(L, K_GEN_SURFELS);
    k_gen_compact<<<(P + kThreads - 1) / kThreads, kThreads, 0, L.stream>>>(kp, frame, radius_map, integrated, poses,
                                                                           timestamp, submap_center, submap_extent, map,
                                                                           lb, cap, counts);
  }
}

void launch_extract_flags(const Launch& L, SurfelPlanes s, const uint32_t* n_dev, uint32_t n_upper, const float* poses,
                          float2 center, float extent, uint8_t* keep, uint32_t* block_counts) {
  if (n_upper == 0) return;
  {
    ScopedKernel sk(L, K_EXTRACT_FLAGS);
    k_extract_flags<<<(n_upper + kThreads - 1) / kThreads, kThreads, 0, L.stream>>>(s, n_dev, poses, center, extent, keep,
                                                                                 block_counts);
  }
}

// ---- 64-byte AoS records (Surfel.h) <-> SoA lanes, for upload / download at the ABI ----
__global__ void k_aos_to_soa(const float4* __restrict__ aos, SurfelPlanes s, uint32_t offset, uint32_t n) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  s.p0[offset + k] = aos[4 * (size_t)k + 0];
  s.p1[offset + k] = aos[4 * (size_t)k + 1];
  s.p2[offset + k] = aos[4 * (size_t)k + 2];
  s.p3[offset + k] = aos[4 * (size_t)k + 3];
}
__global__ void k_soa_to_aos(SurfelPlanes s, float4* __restrict__ aos, uint32_t n) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  aos[4 * (size_t)k + 0] = s.p0[k];
  aos[4 * (size_t)k + 1] = s.p1[k];
  aos[4 * (size_t)k + 2] = s.p2[k];
  aos[4 * (size_t)k + 3] = s.p3[k];
}
void launch_aos_to_soa(const Launch& L, const sb_surfel* aos, SurfelPlanes s, uint32_t offset, uint32_t n) {
  if (n == 0) return;
  {
    ScopedKernel sk(L, K_AOS_TO_SOA);
    k_aos_to_soa<<<(n + kThreads - 1) / kThreads, kThreads, 0, L.stream>>>(reinterpret_cast<const float4*>(aos), s, offset,
                                                                        n);
  }
}
void launch_soa_to_aos(const Launch& L, SurfelPlanes s, sb_surfel* aos, uint32_t n) {
  if (n == 0) return;
  {
    ScopedKernel sk(L, K_SOA_TO_AOS);
    k_soa_to_aos<<<(n + kThreads - 1) / kThreads, kThreads, 0, L.stream>>>(s, reinterpret_cast<float4*>(aos), n);
  }
}

}  // namespace sb
