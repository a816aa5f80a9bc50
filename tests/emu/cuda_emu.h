// cuda_emu.h -- TEST-ONLY host emulation of the small CUDA subset our kernels use.
//
// The authoring container has nvcc but no GPU, so the kernel sources under cudacam_b200/csrc are also
// compiled with g++ against this header (-DB2C_EMU) and run as one OS thread per CUDA thread:
// __syncthreads / __syncwarp / shuffles / votes are real barriers, so any missing synchronisation in
// a kernel shows up as a wrong answer here too.  It exists to debug kernel logic before spending GPU
// minutes; it is never built into the product library and nothing in cudacam_b200/ can reach it.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

struct uint3e { unsigned x = 0, y = 0, z = 0; };
struct dim3 {
  unsigned x = 1, y = 1, z = 1;
  dim3() = default;
  dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; } __attribute__((aligned(16)));
struct int4 { int x, y, z, w; } __attribute__((aligned(16)));
struct float2 { float x, y; };
struct float4 { float x, y, z, w; } __attribute__((aligned(16)));
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{ a, b, c, d }; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{ a, b }; }
typedef void *cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

namespace emu
{
struct Warp {
  std::barrier<> bar;
  uint64_t xchg[32];
  int n;
  explicit Warp(int n_) : bar(n_), n(n_) {}
};
struct Grid {
  std::barrier<> bar;
  explicit Grid(int nblocks) : bar(nblocks) {}
};
// named barrier (PTX bar.sync / bar.arrive with an id and a thread count): completes when `n` threads have arrived,
// arrivers do not wait
struct NamedBar {
  std::mutex m;
  std::condition_variable cv;
  int count = 0;
  unsigned gen = 0;
  void arrive(int n, bool wait)
  {
    std::unique_lock<std::mutex> l(m);
    const unsigned g = gen;
    if (++count == n) {
      count = 0;
      ++gen;
      cv.notify_all();
    } else if (wait) {
      cv.wait(l, [&] { return gen != g; });
    }
  }
};
struct Block {
  NamedBar named[16];
  dim3 gridDim, blockDim;
  uint3e blockIdx;
  std::unique_ptr<std::barrier<>> bar;
  std::vector<std::unique_ptr<Warp>> warps;
  std::vector<char> smem;
  Grid *grid = nullptr;
  std::atomic<int> or_flag{ 0 };
};
struct Thread {
  Block *blk = nullptr;
  uint3e tid;
  int lane = 0;
  Warp *warp = nullptr;
};
inline thread_local Thread tls;

template <class F>
void run_block(Block &b, F &&body)
{
  const int nt = (int)(b.blockDim.x * b.blockDim.y * b.blockDim.z);
  b.bar = std::make_unique<std::barrier<>>(nt);
  b.warps.clear();
  for (int w = 0; w * 32 < nt; ++w) b.warps.push_back(std::make_unique<Warp>(std::min(32, nt - w * 32)));
  std::vector<std::thread> th;
  th.reserve(nt);
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&b, t, &body] {
      tls.blk = &b;
      tls.tid.x = t % b.blockDim.x;
      tls.tid.y = (t / b.blockDim.x) % b.blockDim.y;
      tls.tid.z = t / (b.blockDim.x * b.blockDim.y);
      tls.lane = t % 32;
      tls.warp = b.warps[t / 32].get();
      body();
    });
  for (auto &x : th) x.join();
}

// Runs `body` (a closure calling the kernel with its arguments) for every block of the grid.
// cooperative = all blocks alive at once (grid_sync allowed); otherwise a few blocks at a time.
template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, bool cooperative, F body)
{
  const int nb = (int)(grid.x * grid.y * grid.z);
  Grid g(cooperative ? nb : 1);
  const int conc = cooperative ? nb : std::min(nb, 4);
  std::atomic<int> next{ 0 };
  std::vector<std::thread> drivers;
  for (int d = 0; d < conc; ++d)
    drivers.emplace_back([&, d] {
      for (;;) {
        const int i = cooperative ? d : next.fetch_add(1);
        if (i >= nb) break;
        Block b;
        b.gridDim = grid;
        b.blockDim = block;
        b.blockIdx.x = i % grid.x;
        b.blockIdx.y = (i / grid.x) % grid.y;
        b.blockIdx.z = i / (grid.x * grid.y);
        b.smem.assign(smem_bytes + 64, 0);
        b.grid = cooperative ? &g : nullptr;
        run_block(b, body);
        if (cooperative) break;
      }
    });
  for (auto &t : drivers) t.join();
}

inline void grid_sync();
inline void named_bar_sync(int id, int n);
inline void named_bar_arrive(int id, int n);
// mbarrier in 8 bytes of shared memory: word 0 = (expected << 16) | pending arrivals, word 1 = phase parity
inline void mbar_init(void *p, int count)
{
  auto *w = static_cast<uint32_t *>(p);
  __atomic_store_n(w, ((uint32_t)count << 16) | (uint32_t)count, __ATOMIC_SEQ_CST);
  __atomic_store_n(w + 1, 0u, __ATOMIC_SEQ_CST);
}
inline void mbar_arrive(void *p)
{
  auto *w = static_cast<uint32_t *>(p);
  const uint32_t old = __atomic_fetch_sub(w, 1u, __ATOMIC_SEQ_CST);
  if ((old & 0xFFFFu) == 1u) {   // last arrival: re-arm, then flip the phase (release)
    __atomic_store_n(w, (old & 0xFFFF0000u) | (old >> 16), __ATOMIC_SEQ_CST);
    __atomic_fetch_xor(w + 1, 1u, __ATOMIC_SEQ_CST);
  }
}
inline void mbar_wait(void *p, unsigned parity)
{
  auto *w = static_cast<uint32_t *>(p);
  while (__atomic_load_n(w + 1, __ATOMIC_SEQ_CST) == parity) std::this_thread::yield();
}
}// namespace emu

#define threadIdx (emu::tls.tid)
#define blockIdx (emu::tls.blk->blockIdx)
#define blockDim (emu::tls.blk->blockDim)
#define gridDim (emu::tls.blk->gridDim)

static inline void __syncthreads() { emu::tls.blk->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::tls.warp->bar.arrive_and_wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void emu::grid_sync()
{
  __syncthreads();
  if (tls.tid.x == 0 && tls.tid.y == 0 && tls.tid.z == 0) tls.blk->grid->bar.arrive_and_wait();
  __syncthreads();
}
inline void emu::named_bar_sync(int id, int n) { tls.blk->named[id].arrive(n, true); }
inline void emu::named_bar_arrive(int id, int n) { tls.blk->named[id].arrive(n, false); }
static inline char *emu_dyn_smem()
{
  auto p = reinterpret_cast<uintptr_t>(emu::tls.blk->smem.data());
  return reinterpret_cast<char *>((p + 63) & ~uintptr_t(63));
}

// ---- warp collectives (all lanes of the warp must call, as with a full mask on the GPU) ----
template <class T>
static inline T emu_xchg(T v, int src)
{
  static_assert(sizeof(T) <= 8, "shuffle type too wide");
  emu::Warp *w = emu::tls.warp;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w->xchg[emu::tls.lane] = raw;
  w->bar.arrive_and_wait();
  T r = v;
  if (src >= 0 && src < w->n) memcpy(&r, &w->xchg[src], sizeof(T));
  w->bar.arrive_and_wait();
  return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu_xchg(v, src & 31); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return emu_xchg(v, emu::tls.lane - (int)d); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { return emu_xchg(v, emu::tls.lane + (int)d); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu_xchg(v, emu::tls.lane ^ m); }
static inline unsigned __ballot_sync(unsigned, int pred)
{
  emu::Warp *w = emu::tls.warp;
  w->xchg[emu::tls.lane] = pred ? 1 : 0;
  w->bar.arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < w->n; ++i) r |= (unsigned)w->xchg[i] << i;
  w->bar.arrive_and_wait();
  return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == ((emu::tls.warp->n == 32) ? 0xffffffffu : ((1u << emu::tls.warp->n) - 1)); }
static inline int __syncthreads_or(int p)
{
  emu::Block *b = emu::tls.blk;
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) b->or_flag.store(0);
  __syncthreads();
  if (p) b->or_flag.store(1);
  __syncthreads();
  const int r = b->or_flag.load();
  __syncthreads();
  return r;
}

// ---- atomics ----
template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicMax(T *p, T v)
{
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T> static inline T atomicMin(T *p, T v)
{
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T> static inline T atomicCAS(T *p, T cmp, T v)
{
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}

// ---- scalar intrinsics ----
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
  const uint64_t v = ((uint64_t)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    const unsigned sel = (s >> (4 * i)) & 0xF;
    unsigned byte = (unsigned)(v >> (8 * (sel & 7))) & 0xFF;
    if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
    r |= byte << (8 * i);
  }
  return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { return (unsigned)((((uint64_t)hi << 32) | lo) >> (s & 31)); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { return (unsigned)(((((uint64_t)hi << 32) | lo) << (s & 31)) >> 32); }
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c)
{
  for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF);
  return c;
}
static inline int __dp4a(int a, int b, int c)
{
  for (int i = 0; i < 4; ++i) c += (int)(int8_t)((unsigned)a >> (8 * i)) * (int)(int8_t)((unsigned)b >> (8 * i));
  return c;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __brev(unsigned v)
{
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline unsigned __float2uint_rz(float f) { return (unsigned)f; }
static inline float __uint2float_rn(unsigned u) { return (float)u; }
static inline float __int2float_rn(int u) { return (float)u; }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
template <class T> static inline void __stcg(T *p, T v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
static inline unsigned __vminu2(unsigned a, unsigned b)
{
  const unsigned lo = std::min(a & 0xFFFFu, b & 0xFFFFu), hi = std::min(a >> 16, b >> 16);
  return lo | (hi << 16);
}
using std::max;
using std::min;
