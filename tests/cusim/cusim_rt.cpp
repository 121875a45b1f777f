// cusim_rt.cpp -- TEST INFRASTRUCTURE (see cusim.hpp): fibers, the block scheduler, the worker pool and the handful of CUDA
// runtime calls libsuma_b200's host code makes, all on the CPU. "Device memory" is host memory, streams are synchronous.
#include "cusim.hpp"

#include <sys/mman.h>
#include <time.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

// CUSIM_TSAN (build_sim.py, CUSIM_TSAN=1): ThreadSanitizer as a RACECHECK of the kernels. Every CUDA thread is a TSan fiber;
// a fiber switch establishes NO happens-before (two CUDA threads are unordered unless the kernel synchronises them), the
// rendezvous points do (__syncthreads: all threads of the block; warp collectives: the lanes of the warp), as do atomics,
// __threadfence and the launch boundaries. Volatile accesses -- how the kernels poll words another block publishes -- are
// recorded as relaxed atomics (build flag tsan-distinguish-volatile + the hooks at the end of this file).
#ifdef CUSIM_TSAN
extern "C" {
void* __tsan_get_current_fiber(void);
void* __tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void* fiber);
void __tsan_switch_to_fiber(void* fiber, unsigned flags);
void __tsan_acquire(void* addr);
void __tsan_release(void* addr);
}
#define TSAN_SWITCH(f) __tsan_switch_to_fiber((f), 1u /* no_sync */)
#define TSAN_ACQUIRE(a) __tsan_acquire((void*)(a))
#define TSAN_RELEASE(a) __tsan_release((void*)(a))
#else
#define TSAN_SWITCH(f) ((void)0)
#define TSAN_ACQUIRE(a) ((void)0)
#define TSAN_RELEASE(a) ((void)0)
#endif

extern "C" void cusim_switch(void** save_sp, void* load_sp);
__asm__(R"(
.text
.globl cusim_switch
.type cusim_switch,@function
cusim_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  subq $8, %rsp
  stmxcsr (%rsp)
  fnstcw 4(%rsp)
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  ldmxcsr (%rsp)
  fldcw 4(%rsp)
  addq $8, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size cusim_switch,.-cusim_switch
)");

namespace cusim {

enum { ST_READY = 0, ST_WARP = 1, ST_BLOCK = 2, ST_DONE = 3 };
constexpr size_t kStack = 128 * 1024;
constexpr int kMaxThreads = 1024;
// the executor's "device": CUSIM_SMS multiprocessors (default 4) with two resident blocks each -- a cooperative grid of up
// to 2 x SMS blocks, every block on its own OS thread. More SMs = the persistent kernel's grid gets closer to a real GPU's
// (e.g. its shared-memory pixel cache only engages when the image fits blocks x threads x depth).
static int sim_sms() {
  static const int n = [] {
    const char* e = getenv("CUSIM_SMS");
    int v = e ? atoi(e) : 4;
    return v < 1 ? 1 : (v > 148 ? 148 : v);
  }();
  return n;
}
static int n_workers() { return sim_sms() * 2 > 16 ? sim_sms() * 2 : 16; }

struct Warp {
  uint64_t slot[2][32];
  unsigned active[2];
  unsigned gen;
  int n;
};
struct Block {
  int nthreads = 0, nwarps = 0;
  Fiber fibers[kMaxThreads];
  int pred[kMaxThreads];
  Warp warps[kMaxThreads / 32];
  void* sched_sp = nullptr;
  int count_result = 0;
  const std::function<void()>* body = nullptr;
  char* stacks = nullptr;
  const char* kernel = "";
  void* tsan_fiber[kMaxThreads] = {};  // CUSIM_TSAN: one TSan context per CUDA thread slot of this worker (reused)
  void* tsan_sched = nullptr;
  char sync_block = 0, sync_start = 0, sync_end = 0;  // addresses the happens-before edges hang on
  char sync_warp[kMaxThreads / 32] = {};
};

thread_local Fiber* cur = nullptr;
thread_local BlockIdx* blkid = nullptr;
static thread_local Block* tblk = nullptr;

static inline void yield_to_scheduler() {
  Fiber* f = cur;
  TSAN_SWITCH(f->blk->tsan_sched);
  cusim_switch(&f->sp, f->blk->sched_sp);
}

static void fiber_main() {
  Fiber* f = cur;
  TSAN_ACQUIRE(&f->blk->sync_start);  // after the launch (arguments, earlier kernels)
  (*f->blk->body)();
  f = cur;
  TSAN_RELEASE(&f->blk->sync_end);
  f->state = ST_DONE;
  yield_to_scheduler();
  abort();  // a finished fiber is never resumed
}

void sync_block() {
  Block* b = cur->blk;
  TSAN_RELEASE(&b->sync_block);
  cur->state = ST_BLOCK;
  yield_to_scheduler();
  TSAN_ACQUIRE(&b->sync_block);
}
int sync_block_count(int p) {
  Fiber* f = cur;
  Block* b = f->blk;
  b->pred[f - b->fibers] = p ? 1 : 0;
  TSAN_RELEASE(&b->sync_block);
  f->state = ST_BLOCK;
  yield_to_scheduler();
  TSAN_ACQUIRE(&b->sync_block);
  return b->count_result;
}
unsigned warp_exchange(uint64_t payload, const uint64_t** slots) {
  Fiber* f = cur;
  Warp* w = f->warp;
  const unsigned g = w->gen & 1u;
  w->slot[g][f->lane] = payload;
  char* ws = &f->blk->sync_warp[w - f->blk->warps];
  TSAN_RELEASE(ws);
  f->state = ST_WARP;
  yield_to_scheduler();
  TSAN_ACQUIRE(ws);
  *slots = w->slot[g];
  return w->active[g];
}
// %globaltimer. The kernels only use it to bound their spins (10 s). The lanes of a warp are fibers that run one after the
// other between rendezvous points, so 32 lanes that all time out wait 32 x 10 s here where a GPU waits 10 s:
// CUSIM_CLOCK_SCALE=<n> lets the executor's clock run n times faster (tests of the time-out paths).
unsigned long long globaltimer() {
  static const unsigned long long scale = [] {
    const char* e = getenv("CUSIM_CLOCK_SCALE");
    long v = e ? atol(e) : 1;
    return (unsigned long long)(v < 1 ? 1 : v);
  }();
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ((unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec) * scale;
}
void ptx_unavailable(const char* what) {
  fprintf(stderr, "cusim: %s is not available in the CPU executor\n", what);
  abort();
}

struct Order { int mode; uint64_t seed; };
static const Order& order() {
  static const Order o = [] {
    Order r{0, 0};
    const char* e = getenv("CUSIM_ORDER");
    if (e && !strcmp(e, "reverse")) r.mode = 1;
    if (e && !strncmp(e, "random", 6)) { r.mode = 2; r.seed = e[6] == ':' ? strtoull(e + 7, nullptr, 10) : 1; }
    return r;
  }();
  return o;
}

static void deadlock(Block* b) {
  int c[4] = {0, 0, 0, 0};
  for (int i = 0; i < b->nthreads; ++i) c[b->fibers[i].state]++;
  fprintf(stderr, "cusim: deadlock in kernel %s, block (%u,%u): %d ready, %d in a warp collective, %d at __syncthreads, %d done\n",
          b->kernel, blkid->bid.x, blkid->bid.y, c[0], c[1], c[2], c[3]);
  abort();
}

static void run_block(Block* b, dim3 bdim) {
  const int n = (int)(bdim.x * bdim.y * bdim.z);
  if (n > kMaxThreads || n <= 0) ptx_unavailable("block size");
  b->nthreads = n;
  b->nwarps = (n + 31) / 32;
  for (int i = 0; i < n; ++i) {
    Fiber& f = b->fibers[i];
    f.tid.x = (unsigned)i % bdim.x;
    f.tid.y = ((unsigned)i / bdim.x) % bdim.y;
    f.tid.z = (unsigned)i / (bdim.x * bdim.y);
    f.blk = b;
    f.warp = &b->warps[i / 32];
    f.lane = i & 31;
    f.state = ST_READY;
    uintptr_t top = ((uintptr_t)(b->stacks + (size_t)(i + 1) * kStack)) & ~(uintptr_t)15;
    uint64_t* s = (uint64_t*)top;
    s[-1] = 0;                       // return address of fiber_main (never used)
    s[-2] = (uint64_t)(uintptr_t)&fiber_main;
    s[-3] = s[-4] = s[-5] = s[-6] = s[-7] = s[-8] = 0;  // rbp rbx r12 r13 r14 r15
    s[-9] = 0x1F80ull | (0x037Full << 32);              // mxcsr | x87 control word
    f.sp = (void*)&s[-9];
  }
  for (int w = 0; w < b->nwarps; ++w) {
    b->warps[w].gen = 0;
    b->warps[w].n = (w == b->nwarps - 1) ? n - 32 * w : 32;
  }
  // Execution order of the warps of a block and of the lanes of a warp between two rendezvous points. CUDA promises none,
  // so a correct kernel gives the same bits under every order: CUSIM_ORDER=reverse / random:<seed> look for code that
  // only works when a warp happens to run in lockstep or the warps in index order.
  int worder[kMaxThreads / 32], lorder[32];
  for (int w = 0; w < b->nwarps; ++w) worder[w] = w;
  for (int l = 0; l < 32; ++l) lorder[l] = l;
  const Order& ord = order();
  if (ord.mode == 1) {
    for (int w = 0; w < b->nwarps; ++w) worder[w] = b->nwarps - 1 - w;
    for (int l = 0; l < 32; ++l) lorder[l] = 31 - l;
  } else if (ord.mode == 2) {
    uint64_t x = ord.seed * 0x9E3779B97F4A7C15ull + ((uint64_t)blkid->bid.x << 20) + blkid->bid.y + 0x1234567ull;
    auto rnd = [&x]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for (int w = b->nwarps - 1; w > 0; --w) { int j = (int)(rnd() % (uint64_t)(w + 1)); int t = worder[w]; worder[w] = worder[j]; worder[j] = t; }
    for (int l = 31; l > 0; --l) { int j = (int)(rnd() % (uint64_t)(l + 1)); int t = lorder[l]; lorder[l] = lorder[j]; lorder[j] = t; }
  }
#ifdef CUSIM_TSAN
  b->tsan_sched = __tsan_get_current_fiber();
  for (int i = 0; i < n; ++i)
    if (!b->tsan_fiber[i]) b->tsan_fiber[i] = __tsan_create_fiber(0);
#endif
  TSAN_RELEASE(&b->sync_start);
  int live = n;
  while (live > 0) {
    bool progressed = false;
    for (int wi = 0; wi < b->nwarps; ++wi) {
      const int w = worder[wi];
      Warp& W = b->warps[w];
      Fiber* lanes = b->fibers + 32 * w;
      for (;;) {
        bool ran = false;
        for (int li = 0; li < 32; ++li) {
          const int l = lorder[li];
          if (l >= W.n) continue;
          Fiber* f = lanes + l;
          if (f->state != ST_READY) continue;
          cur = f;
          TSAN_SWITCH(b->tsan_fiber[f - b->fibers]);
          cusim_switch(&b->sched_sp, f->sp);
          ran = true;
          if (f->state == ST_DONE) --live;
        }
        progressed |= ran;
        unsigned mask = 0;
        int nblock = 0;
        for (int l = 0; l < W.n; ++l) {
          if (lanes[l].state == ST_WARP) mask |= 1u << l;
          if (lanes[l].state == ST_BLOCK) ++nblock;
        }
        if (!mask) break;
        if (nblock) deadlock(b);  // part of a warp at __syncthreads, part in a warp collective
        W.active[W.gen & 1u] = mask;
        W.gen++;
        for (int l = 0; l < W.n; ++l)
          if (lanes[l].state == ST_WARP) lanes[l].state = ST_READY;
        progressed = true;
      }
    }
    if (live > 0) {
      int waiting = 0, cnt = 0;
      for (int i = 0; i < n; ++i)
        if (b->fibers[i].state == ST_BLOCK) {
          ++waiting;
          cnt += b->pred[i];
        }
      if (waiting == live) {
        b->count_result = cnt;
        for (int i = 0; i < n; ++i)
          if (b->fibers[i].state == ST_BLOCK) {
            b->fibers[i].state = ST_READY;
            b->pred[i] = 0;
          }
        progressed = true;
      } else if (!progressed) {
        deadlock(b);
      }
    }
  }
  cur = nullptr;
  TSAN_ACQUIRE(&b->sync_end);
}

// ---- worker pools ------------------------------------------------------------------------------------------------------
// One pool per HOST thread that launches kernels: two contexts driven by two host threads ("two GPUs") run their kernels
// at the same time, which the in-kernel peer exchange of the row-striped Gauss-Newton needs (each rank's kernel waits for
// the words the other ranks' kernels store). Pools and their threads live until the process ends.
struct Job {
  LaunchCfg cfg;
  const std::function<void()>* body = nullptr;
  const char* name = "";
  unsigned nblocks = 0;
  int active_workers = 0;
  std::atomic<unsigned> next{0};
};
struct Pool {
  std::mutex mu;
  std::condition_variable cv_start, cv_done;
  Job job;
  uint64_t gen = 0;
  int running = 0;
};

static int hw_threads() {
  const char* e = getenv("CUSIM_THREADS");
  int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
  if (n < 1) n = 1;
  if (n > n_workers()) n = n_workers();
  return n;
}

static void worker(Pool* P, int id) {
  Block* b = new Block();
  b->stacks = (char*)mmap(nullptr, kStack * kMaxThreads, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (b->stacks == (char*)MAP_FAILED) abort();
  tblk = b;
  BlockIdx bi;
  blkid = &bi;
  uint64_t seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(P->mu);
      P->cv_start.wait(lk, [&] { return P->gen != seen; });
      seen = P->gen;
    }
    Job& J = P->job;
    if (id < J.active_workers) {
      const LaunchCfg& c = J.cfg;
      bi.bdim = c.block;
      bi.gdim = c.grid;
      b->body = J.body;
      b->kernel = J.name;
      for (;;) {
        unsigned i = J.next.fetch_add(1);
        if (i >= J.nblocks) break;
        bi.bid.x = i % c.grid.x;
        bi.bid.y = (i / c.grid.x) % c.grid.y;
        bi.bid.z = i / (c.grid.x * c.grid.y);
        run_block(b, c.block);
      }
    }
    {
      std::lock_guard<std::mutex> lk(P->mu);
      if (--P->running == 0) P->cv_done.notify_all();
    }
  }
}

void launch_impl(const char* name, const LaunchCfg& c, bool cooperative, const std::function<void()>& body) {
  static thread_local Pool* P = nullptr;
  const unsigned nblocks = c.grid.x * c.grid.y * c.grid.z;
  if (nblocks == 0) return;
  if (!P) {
    P = new Pool();  // never destroyed: its detached workers wait on it
    for (int i = 0; i < n_workers(); ++i) std::thread(worker, P, i).detach();
  }
  if (cooperative && nblocks > (unsigned)n_workers()) ptx_unavailable("a cooperative grid larger than the worker pool");
  static const bool trace = getenv("CUSIM_TRACE") != nullptr;  // kernel names as they start and end (finding a hang)
  if (trace) fprintf(stderr, "cusim: launch %s grid %u block %u\n", name, nblocks, c.block.x * c.block.y * c.block.z);
  {
    std::lock_guard<std::mutex> lk(P->mu);
    Job& J = P->job;
    J.cfg = c;
    J.body = &body;
    J.name = name;
    J.nblocks = nblocks;
    J.next.store(0);
    // blocks of a cooperative launch wait for each other: every block needs its own OS thread
    int act = cooperative ? (int)nblocks : hw_threads();
    if ((unsigned)act > nblocks) act = (int)nblocks;
    J.active_workers = act;
    P->running = n_workers();
    ++P->gen;
  }
  P->cv_start.notify_all();
  std::unique_lock<std::mutex> lk(P->mu);
  P->cv_done.wait(lk, [&] { return P->running == 0; });
  if (trace) fprintf(stderr, "cusim: done   %s\n", name);
}

}  // namespace cusim

// ---- the runtime calls of sb_api.cu ------------------------------------------------------------------------------------
struct cusimStream { int dummy; };
struct cusimEvent { double ms; };
static double now_ms() { return (double)cusim::globaltimer() * 1e-6; }

cudaError_t cudaMalloc(void** p, size_t n) {
  void* q = nullptr;
  if (posix_memalign(&q, 512, n ? n : 1)) return cudaErrorMemoryAllocation;
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new cusimStream(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new cusimEvent(); (*e)->ms = 0; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->ms = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->ms - a->ms); return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cusim error"; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) {  // CUSIM_DEVICES "GPUs" (default 1); every host thread that launches is its own
  const char* e = getenv("CUSIM_DEVICES");
  *n = e && atoi(e) > 0 ? atoi(e) : 1;
  return cudaSuccess;
}
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof(*p));
  snprintf(p->name, sizeof(p->name), "cusim CPU executor (not a GPU)");
  p->multiProcessorCount = cusim::sim_sms();
  p->totalGlobalMem = (size_t)16 << 30;
  p->major = 10;
  p->minor = 0;
  return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
  *v = (a == cudaDevAttrCooperativeLaunch) ? 1 : (a == cudaDevAttrMultiProcessorCount ? cusim::sim_sms() : 0);
  return cudaSuccess;
}
int cusim_occupancy_blocks_per_sm() { return 2; }
cudaError_t cudaGetDriverEntryPoint(const char*, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* qr) {
  *fn = nullptr;  // no driver, no TMA unit: the library takes its plain-load instantiation
  if (qr) *qr = cudaDriverEntryPointSymbolNotFound;
  return cudaErrorNotSupported;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  memset(h, 0, sizeof(*h));
  memcpy(h->reserved, &p, sizeof(p));
  return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
  memcpy(p, h.reserved, sizeof(*p));  // same process: the loop-back exchange test
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

#ifdef CUSIM_TSAN
// Volatile accesses of the kernels (polling / publishing words between blocks and ranks). TSan does not model stand-alone
// fences, and the kernels' hand-overs are "__threadfence(); volatile store" on one side and "volatile load; __threadfence()"
// on the other: a volatile store counts as a store-release and a volatile load as a load-acquire ON THAT ADDRESS (stronger
// than the GPU's relaxed access where a kernel deliberately goes without the fence -- the self-validating words -- but no
// ordering is claimed there, and nothing is ordered between threads that do not hand over through the same word).
extern "C" {
unsigned char __tsan_atomic8_load(const volatile unsigned char*, int);
unsigned short __tsan_atomic16_load(const volatile unsigned short*, int);
unsigned int __tsan_atomic32_load(const volatile unsigned int*, int);
unsigned long long __tsan_atomic64_load(const volatile unsigned long long*, int);
unsigned char __tsan_atomic8_fetch_or(volatile unsigned char*, unsigned char, int);
unsigned short __tsan_atomic16_fetch_or(volatile unsigned short*, unsigned short, int);
unsigned int __tsan_atomic32_fetch_or(volatile unsigned int*, unsigned int, int);
unsigned long long __tsan_atomic64_fetch_or(volatile unsigned long long*, unsigned long long, int);
#define NOTSAN __attribute__((no_sanitize("thread")))
enum { kAcquire = 2, kRelease = 3 };  // __ATOMIC_ACQUIRE, __ATOMIC_RELEASE
NOTSAN void __tsan_volatile_read1(void* a) { __tsan_atomic8_load((const volatile unsigned char*)a, kAcquire); }
NOTSAN void __tsan_volatile_read2(void* a) { __tsan_atomic16_load((const volatile unsigned short*)a, kAcquire); }
NOTSAN void __tsan_volatile_read4(void* a) { __tsan_atomic32_load((const volatile unsigned int*)a, kAcquire); }
NOTSAN void __tsan_volatile_read8(void* a) { __tsan_atomic64_load((const volatile unsigned long long*)a, kAcquire); }
NOTSAN void __tsan_volatile_read16(void* a) {
  __tsan_atomic64_load((const volatile unsigned long long*)a, kAcquire);
  __tsan_atomic64_load((const volatile unsigned long long*)a + 1, kAcquire);
}
NOTSAN void __tsan_volatile_write1(void* a) { __tsan_atomic8_fetch_or((volatile unsigned char*)a, 0, kRelease); }
NOTSAN void __tsan_volatile_write2(void* a) { __tsan_atomic16_fetch_or((volatile unsigned short*)a, 0, kRelease); }
NOTSAN void __tsan_volatile_write4(void* a) { __tsan_atomic32_fetch_or((volatile unsigned int*)a, 0, kRelease); }
NOTSAN void __tsan_volatile_write8(void* a) { __tsan_atomic64_fetch_or((volatile unsigned long long*)a, 0, kRelease); }
NOTSAN void __tsan_volatile_write16(void* a) {
  __tsan_atomic64_fetch_or((volatile unsigned long long*)a, 0, kRelease);
  __tsan_atomic64_fetch_or((volatile unsigned long long*)a + 1, 0, kRelease);
}
}
#endif
