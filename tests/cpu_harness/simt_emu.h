// Minimal SIMT emulator: runs a CUDA __global__ function, compiled as plain host C++ (g++ -DFM_HOST_EMU), one host thread
// per CUDA thread, one block at a time.  Enough for the library's CUDA-core kernels (LayerNorm, loss head, misc):
// threadIdx/blockIdx/blockDim/gridDim, __syncthreads, __syncwarp, warp shuffles, __shared__ (static storage: blocks run
// sequentially), dynamic shared memory (FM_DYN_SMEM), atomicAdd/Min/Max, __ldg and the few device math intrinsics used.
// TEST INFRASTRUCTURE ONLY: it exists so that kernels written without GPU access get executed before they reach a B200.
#pragma once
#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>      // vector types (uint4, float4, dim3), host-side definitions of __device__/__global__ as nothing

namespace emu {

struct Block {
  explicit Block(int nthreads)
      : n(nthreads), bar(nthreads), slots(nthreads, 0), dyn(nullptr) {
    for (int w = 0; w < (nthreads + 31) / 32; ++w) {
      const int lanes = std::min(32, nthreads - w * 32);
      warp_bar.emplace_back(std::make_unique<std::barrier<>>(lanes));
    }
  }
  int n;
  std::barrier<> bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint64_t> slots;        // shuffle exchange buffer, one slot per thread
  unsigned char* dyn;                 // dynamic shared memory as the kernel sees it (16-byte aligned, NOT 1024-aligned)
  size_t dyn_bytes = 0;
  std::atomic<int> or_flag{0};        // __syncthreads_or
  // ---- state of the emulated async hardware (tc_emu.h): everything below is guarded by hw_mu
  std::mutex hw_mu;
  std::condition_variable hw_cv;
  struct AsyncOp { void* bar; bool is_commit; std::function<void()> run; };
  std::vector<AsyncOp> tma_pending;   // TMA loads not yet performed (performed when somebody waits on their barrier)
  std::deque<AsyncOp> mma_fifo;       // issued tcgen05.mma / tcgen05.commit, in order, not yet performed
  std::vector<float> tmem;            // [128 lanes][512 columns], NaN until written
  uint32_t tmem_alloc_mask = 0;       // one bit per 32 columns
  dim3 bid;
};

struct Ctx {
  dim3 tid, bid, bdim, gdim;
  Block* blk = nullptr;
  int linear = 0;
};
inline thread_local Ctx ctx;

// Blocks normally run one after the other (static __shared__ variables are plain statics).  Persistent kernels whose CTAs
// talk to each other through global memory (serial split-K flags) need all CTAs resident: set concurrent_next before the
// launch; such kernels must keep all shared state in dynamic shared memory.
inline thread_local bool concurrent_next = false;

template <typename F>
void run_block(dim3 grid, dim3 block, dim3 bid, size_t dyn_bytes, F& kernel) {
  const int nthreads = static_cast<int>(block.x * block.y * block.z);
  // the hardware only promises 16-byte alignment of the dynamic segment: start 16 bytes past a 1 KB boundary so that a
  // kernel that forgets to align its swizzled tiles fails here as it would there
  unsigned char* raw = static_cast<unsigned char*>(std::aligned_alloc(1024, ((dyn_bytes + 1023) / 1024 + 2) * 1024));
  std::memset(raw, 0xCD, ((dyn_bytes + 1023) / 1024 + 2) * 1024);
  {
    Block blk(nthreads);
    blk.dyn = raw + 16;
    blk.dyn_bytes = dyn_bytes;
    blk.bid = bid;
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t) {
      th.emplace_back([&, t] {
        ctx.blk = &blk;
        ctx.linear = t;
        ctx.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        ctx.bid = bid;
        ctx.bdim = block;
        ctx.gdim = grid;
        kernel();
        // a thread that has left the kernel must not block the barriers of those still inside
        blk.warp_bar[t / 32]->arrive_and_drop();
        blk.bar.arrive_and_drop();
      });
    }
    for (auto& x : th) x.join();
  }
  std::free(raw);
}

// kernel: callable run by every thread; grid/block as in <<<grid, block, dyn_bytes>>>
template <typename F>
void launch(dim3 grid, dim3 block, size_t dyn_bytes, F&& kernel) {
  const bool concurrent = concurrent_next;
  concurrent_next = false;
  std::vector<std::thread> blocks;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        if (concurrent) blocks.emplace_back([&, bx, by, bz] { run_block(grid, block, dim3(bx, by, bz), dyn_bytes, kernel); });
        else run_block(grid, block, dim3(bx, by, bz), dyn_bytes, kernel);
      }
  for (auto& b : blocks) b.join();
}

template <typename T>
inline T shfl_exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64-bit values");
  Block& b = *ctx.blk;
  const int w = ctx.linear / 32;
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  b.slots[ctx.linear] = raw;
  b.warp_bar[w]->arrive_and_wait();
  T out = v;
  if (src_lane >= 0 && src_lane < 32 && w * 32 + src_lane < b.n) {
    const uint64_t r = b.slots[w * 32 + src_lane];
    std::memcpy(&out, &r, sizeof(T));
  }
  b.warp_bar[w]->arrive_and_wait();
  return out;
}

}  // namespace emu

// ---- CUDA built-ins the kernels use, in terms of the emulator
#define threadIdx (emu::ctx.tid)
#define blockIdx (emu::ctx.bid)
#define blockDim (emu::ctx.bdim)
#define gridDim (emu::ctx.gdim)
#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __restrict__
#define __restrict__
#undef __grid_constant__
#define __grid_constant__
#define FM_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::ctx.blk->dyn)

inline void __syncthreads() { emu::ctx.blk->bar.arrive_and_wait(); }
inline int __syncthreads_or(int pred) {
  emu::Block& b = *emu::ctx.blk;
  if (pred) b.or_flag.store(1);
  b.bar.arrive_and_wait();
  const int r = b.or_flag.load();
  b.bar.arrive_and_wait();
  if (emu::ctx.linear == 0) b.or_flag.store(0);
  b.bar.arrive_and_wait();
  return r;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::ctx.blk->warp_bar[emu::ctx.linear / 32]->arrive_and_wait(); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu::shfl_exchange(v, (emu::ctx.linear % 32) ^ lane_mask); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl_exchange(v, src); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
  const int lane = emu::ctx.linear % 32;
  return emu::shfl_exchange(v, lane >= static_cast<int>(delta) ? lane - static_cast<int>(delta) : -1);   // out of range: own value
}
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_relaxed); }
inline int atomicMin(int* p, int v) { std::atomic_ref<int> a(*p); int o = a.load(); while (v < o && !a.compare_exchange_weak(o, v)) {} return o; }
inline int atomicMax(int* p, int v) { std::atomic_ref<int> a(*p); int o = a.load(); while (v > o && !a.compare_exchange_weak(o, v)) {} return o; }
inline unsigned int atomicExch(unsigned int* p, unsigned int v) { return std::atomic_ref<unsigned int>(*p).exchange(v); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline long long clock64() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count() / 64; }  // the kernels' own 4e9-"cycle" watchdogs fire after ~4 min here
[[noreturn]] inline void __trap() { std::fprintf(stderr, "EMU: __trap() in block (%u,%u,%u) thread %d\n", emu::ctx.bid.x, emu::ctx.bid.y, emu::ctx.bid.z, emu::ctx.linear); std::abort(); }
inline float __uint_as_float(unsigned int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned int __float_as_uint(float f) { unsigned int u; std::memcpy(&u, &f, 4); return u; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __logf(float x) { return std::log(x); }
inline float __expf(float x) { return std::exp(x); }
using std::max;
using std::min;
