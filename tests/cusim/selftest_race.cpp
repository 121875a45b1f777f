// TEST INFRASTRUCTURE: does the racecheck build of the executor (CUSIM_TSAN) see a missing barrier, and only that?
// Two kernels exchange values through shared memory: one with the __syncthreads() between the store and the neighbour's
// load, one without. ThreadSanitizer must report the second and stay silent on the first; likewise a warp-level exchange
// with and without __syncwarp(). Built and run by tests/cusim/racecheck.py --selftest.
#include "cusim.hpp"

__global__ void k_block(int* out, int with_barrier) {
  __shared__ long long s[128];  // one 8-byte TSan granule per element: two accesses each
  s[threadIdx.x] = (int)threadIdx.x * 3;
  if (with_barrier) __syncthreads();
  out[blockIdx.x * 128 + threadIdx.x] = (int)s[(threadIdx.x + 1) & 127];
}
__global__ void k_warp(int* out, int with_barrier) {
  __shared__ long long s[32];
  s[threadIdx.x] = (int)threadIdx.x * 5;
  if (with_barrier) __syncwarp();
  out[blockIdx.x * 128 + threadIdx.x] = (int)s[(threadIdx.x + 7) & 31];
}

int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : 0;  // bit 0: block kernel without barrier, bit 1: warp kernel without barrier
  int* out = nullptr;
  cudaMalloc(&out, 256 * 128 * sizeof(int));
  int wb = !(which & 1), ww = !(which & 2);
  cusim::launch("k_block", cusim::cfg(64, 128), [&] { k_block(out, wb); });
  // many blocks: ThreadSanitizer's history is finite (slot / epoch recycling under this many fibers), a single racing pair can be
  // missed; a kernel's race is hit over and over
  cusim::launch("k_warp", cusim::cfg(256, 32), [&] { k_warp(out, ww); });
  printf("selftest ran (%d)\n", which);
  return 0;
}
