// cusim.hpp -- TEST INFRASTRUCTURE, never part of the product: a CPU executor for the CUDA sources of libsuma_b200.
//
// tests/cusim/build_sim.py compiles semantic_suma_b200/csrc/*.cu UNCHANGED (apart from a mechanical rewrite of the
// `kernel<<<grid, block, smem, stream>>>(args)` launch syntax and of inline PTX) with g++ against this header into
// tests/cusim/_build/libsuma_b200_sim.so, which exports the same C ABI. Every CUDA thread is a fiber, every block runs on a
// worker OS thread (blocks of one launch run concurrently, so inter-block hand-overs -- tickets, decoupled look-back, the
// persistent Gauss-Newton kernel -- execute as written), __syncthreads / warp collectives are rendezvous points of the
// fibers. Purpose: run the `-m gpu` parity tests' CUDA code paths against the oracle in a container without a GPU
// (tests/test_cusim.py) -- a check of the kernels' LOGIC; it says nothing about their speed, and a GPU run stays the proof.
// The product (semantic_suma_b200/) never loads this library: only tests/conftest.py swaps it in, on request.
#pragma once

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

// ---- qualifiers ----------------------------------------------------------------------------------------------------
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types --------------------------------------------------------------------------------------------------
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct ushort2 { unsigned short x, y; };
struct uchar4 { unsigned char x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---- runtime types -------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef struct cusimStream* cudaStream_t;
typedef struct cusimEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1, cudaEnableDefault = 0 };
enum cudaDeviceAttr { cudaDevAttrCooperativeLaunch = 95, cudaDevAttrMultiProcessorCount = 16 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp {
  char name[256];
  int multiProcessorCount;
  size_t totalGlobalMem;
  int major, minor;
};

cudaError_t cudaMalloc(void** p, size_t n);
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocHost(void** p, size_t n);
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t st = 0);
cudaError_t cudaMemset(void* d, int v, size_t n);
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st = 0);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetLastError();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int d);
cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long flags, cudaDriverEntryPointQueryResult* qr = nullptr);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
int cusim_occupancy_blocks_per_sm();
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) {
  *n = cusim_occupancy_blocks_per_sm();
  return cudaSuccess;
}

// ---- driver types named by make_key_tensor_map (the sim has no TMA unit: the entry point is "not found") ----------------
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
typedef uint64_t cuuint64_t;
typedef uint32_t cuuint32_t;
struct alignas(64) CUtensorMap { uint64_t opaque[16]; };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT64 = 9 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };

// ---- execution model -----------------------------------------------------------------------------------------------
namespace cusim {
struct Block;
struct Warp;
struct Fiber {
  uint3 tid;          // threadIdx
  Block* blk;
  Warp* warp;
  int lane;
  int state;
  void* sp;           // saved stack pointer while switched out
};
struct BlockIdx {
  uint3 bid;          // blockIdx
  dim3 bdim, gdim;
};
extern thread_local Fiber* cur;      // the CUDA thread running on this OS thread
extern thread_local BlockIdx* blkid;

struct LaunchCfg {
  dim3 grid, block;
  size_t smem;
};
static inline LaunchCfg cfg(dim3 g, dim3 b, size_t smem = 0, cudaStream_t = 0) { return LaunchCfg{g, b, smem}; }
void launch_impl(const char* name, const LaunchCfg& c, bool cooperative, const std::function<void()>& body);
template <class F> static inline void launch(const char* name, const LaunchCfg& c, F&& f) {
  launch_impl(name, c, false, std::function<void()>(f));
}
template <class F> static inline cudaError_t launch_coop(const char* name, const LaunchCfg& c, F&& f) {
  launch_impl(name, c, true, std::function<void()>(f));
  return cudaSuccess;
}
// rendezvous points (cusim_rt.cpp)
void sync_block();
int sync_block_count(int pred);
unsigned warp_exchange(uint64_t payload, const uint64_t** slots);  // returns the mask of participating lanes
unsigned long long globaltimer();
[[noreturn]] void ptx_unavailable(const char* what);
}  // namespace cusim

#define threadIdx (cusim::cur->tid)
#define blockIdx (cusim::blkid->bid)
#define blockDim (cusim::blkid->bdim)
#define gridDim (cusim::blkid->gdim)

// ---- synchronisation and warp collectives ------------------------------------------------------------------------------
static inline void __syncthreads() { cusim::sync_block(); }
static inline int __syncthreads_count(int p) { return cusim::sync_block_count(p); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  const uint64_t* s;
  cusim::warp_exchange(0, &s);
}
namespace cusim {
template <class T> static inline uint64_t pack(T v) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t u = 0;
  memcpy(&u, &v, sizeof(T));
  return u;
}
template <class T> static inline T unpack(uint64_t u) {
  T v;
  memcpy(&v, &u, sizeof(T));
  return v;
}
static inline void full_mask(unsigned m) {
  if (m != 0xffffffffu) ptx_unavailable("warp collective with a partial mask");
}
}  // namespace cusim
template <class T> static inline T __shfl_sync(unsigned m, T v, int src, int width = 32) {
  cusim::full_mask(m);
  const uint64_t* s;
  const int lane = cusim::cur->lane;
  cusim::warp_exchange(cusim::pack(v), &s);
  const int l = (lane & ~(width - 1)) | (src & (width - 1));
  return cusim::unpack<T>(s[l]);
}
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) {
  cusim::full_mask(m);
  const uint64_t* s;
  const int lane = cusim::cur->lane;
  cusim::warp_exchange(cusim::pack(v), &s);
  const int l = lane - (int)d;
  return (l < (lane & ~(width - 1))) ? v : cusim::unpack<T>(s[l]);
}
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) {
  cusim::full_mask(m);
  const uint64_t* s;
  const int lane = cusim::cur->lane;
  cusim::warp_exchange(cusim::pack(v), &s);
  const int l = lane + (int)d;
  return (l > ((lane & ~(width - 1)) | (width - 1))) ? v : cusim::unpack<T>(s[l]);
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) {
  cusim::full_mask(m);
  const uint64_t* s;
  const int lane = cusim::cur->lane;
  cusim::warp_exchange(cusim::pack(v), &s);
  (void)width;
  return cusim::unpack<T>(s[lane ^ x]);
}
static inline unsigned __ballot_sync(unsigned m, int pred) {
  cusim::full_mask(m);
  const uint64_t* s;
  const unsigned act = cusim::warp_exchange(pred ? 1u : 0u, &s);
  unsigned r = 0;
  for (int l = 0; l < 32; ++l)
    if (((act >> l) & 1u) && s[l]) r |= 1u << l;
  return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline int __all_sync(unsigned m, int pred) {
  cusim::full_mask(m);
  const uint64_t* s;
  const unsigned act = cusim::warp_exchange(pred ? 1u : 0u, &s);
  for (int l = 0; l < 32; ++l)
    if (((act >> l) & 1u) && !s[l]) return 0;
  return 1;
}
template <class T> static inline T __reduce_add_sync(unsigned m, T v) {
  cusim::full_mask(m);
  const uint64_t* s;
  const unsigned act = cusim::warp_exchange(cusim::pack(v), &s);
  T r = 0;
  for (int l = 0; l < 32; ++l)
    if ((act >> l) & 1u) r += cusim::unpack<T>(s[l]);
  return r;
}

// CUDA atomics are relaxed; the kernels pair them with __threadfence() for ordering. ThreadSanitizer (CUSIM_TSAN) does not
// model stand-alone fences, so there the atomics themselves carry acquire-release (see cusim_rt.cpp on volatile accesses).
#ifdef CUSIM_TSAN
#define CUSIM_MO __ATOMIC_SEQ_CST
#else
#define CUSIM_MO __ATOMIC_RELAXED
#endif
// ---- memory and atomics ----------------------------------------------------------------------------------------------
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) {
  __asm__ __volatile__("" ::: "memory");
  T v = *p;
  __asm__ __volatile__("" ::: "memory");
  return v;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, CUSIM_MO); }
static inline unsigned atomicAdd(unsigned* p, int v) { return __atomic_fetch_add(p, (unsigned)v, CUSIM_MO); }
template <class T> static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, CUSIM_MO); }
template <class T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, CUSIM_MO); }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, CUSIM_MO, CUSIM_MO);
  return cmp;
}
template <class T> static inline T atomicMin(T* p, T v) {
  T old = __atomic_load_n(p, CUSIM_MO);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, CUSIM_MO, CUSIM_MO)) {
  }
  return old;
}
template <class T> static inline T atomicMax(T* p, T v) {
  T old = __atomic_load_n(p, CUSIM_MO);
  while (v > old && !__atomic_compare_exchange_n(p, &old, v, true, CUSIM_MO, CUSIM_MO)) {
  }
  return old;
}

// ---- scalar intrinsics -----------------------------------------------------------------------------------------------
static inline unsigned __float_as_uint(float f) { return cusim::pack(f) & 0xffffffffu; }
static inline float __uint_as_float(unsigned u) { return cusim::unpack<float>(u); }
static inline int __float_as_int(float f) { return (int)__float_as_uint(f); }
static inline float __int_as_float(int i) { return __uint_as_float((unsigned)i); }
// PTX cvt semantics, not x86's: NaN converts to 0 and out-of-range values saturate (x86 returns 0x8000... for both), so a
// NaN or an overflow that reaches a conversion shows on the executor the way it would on the GPU
static inline long long __float2ll_rn(float f) {
  if (f != f) return 0;
  if (f >= 9223372036854775808.0f) return 0x7fffffffffffffffll;
  if (f <= -9223372036854775808.0f) return (long long)0x8000000000000000ull;
  return llrintf(f);
}
static inline long long __double2ll_rn(double d) {
  if (d != d) return 0;
  if (d >= 9223372036854775808.0) return 0x7fffffffffffffffll;
  if (d <= -9223372036854775808.0) return (long long)0x8000000000000000ull;
  return llrint(d);
}
static inline double __hiloint2double(int hi, int lo) {
  return cusim::unpack<double>(((uint64_t)(uint32_t)hi << 32) | (uint64_t)(uint32_t)lo);
}
static inline int __double2hiint(double d) { return (int)(cusim::pack(d) >> 32); }
static inline int __double2loint(double d) { return (int)(cusim::pack(d) & 0xffffffffu); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

// CUDA's global min / max overload set
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
static inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
static inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline double min(double a, double b) { return fmin(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
