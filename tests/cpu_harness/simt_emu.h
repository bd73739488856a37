// Minimal SIMT emulator: runs a CUDA __global__ function, compiled as plain host C++ (g++ -DFM_HOST_EMU), one host thread
// per CUDA thread, one block at a time.  Enough for the library's CUDA-core kernels (LayerNorm, loss head, misc):
// threadIdx/blockIdx/blockDim/gridDim, __syncthreads, __syncwarp, warp shuffles, __shared__ (static storage: blocks run
// sequentially), dynamic shared memory (FM_DYN_SMEM), atomicAdd/Min/Max, __ldg and the few device math intrinsics used.
// TEST INFRASTRUCTURE ONLY: it exists so that kernels written without GPU access get executed before they reach a B200.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
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
  unsigned char* dyn;
};

struct Ctx {
  dim3 tid, bid, bdim, gdim;
  Block* blk = nullptr;
  int linear = 0;
};
inline thread_local Ctx ctx;

// kernel: callable run by every thread; grid/block as in <<<grid, block, dyn_bytes>>>
template <typename F>
void launch(dim3 grid, dim3 block, size_t dyn_bytes, F&& kernel) {
  const int nthreads = static_cast<int>(block.x * block.y * block.z);
  std::vector<unsigned char> dyn(dyn_bytes + 16);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block blk(nthreads);
        blk.dyn = dyn.data();
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t) {
          th.emplace_back([&, t] {
            ctx.blk = &blk;
            ctx.linear = t;
            ctx.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            ctx.bid = dim3(bx, by, bz);
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
#define FM_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>((reinterpret_cast<uintptr_t>(emu::ctx.blk->dyn) + 15) & ~uintptr_t(15))

inline void __syncthreads() { emu::ctx.blk->bar.arrive_and_wait(); }
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
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __logf(float x) { return std::log(x); }
inline float __expf(float x) { return std::exp(x); }
using std::max;
using std::min;
