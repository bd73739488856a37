// Functional model of the sm_100a async hardware the library's tensor-core kernels use, for tests/cpu_harness:
// mbarrier (phases, arrival counts, transaction bytes), TMA tiled 2-D loads (SWIZZLE_128B / NONE, zero fill),
// tcgen05.mma kind::f16 with shared-memory matrix descriptors (K-major and MN-major, SWIZZLE_128B), TMEM
// (alloc / ld 32x32b / dealloc) and tcgen05.commit.  Included by csrc/ptx.cuh when FM_HOST_EMU is defined, where it
// supplies the SAME function names the inline-PTX wrappers have.
//
// What the model is anchored on: csrc/ (the validated tree) produced correct results on a B200 with exactly these
// descriptor encodings (K-major +32 B per K step inside a swizzle atom; MN-major LBO 8192 / SBO 1024, +2048 B per K step;
// the stacked MN-major operand with LBO 16384).  The model below is the reading of the hardware under which all of those
// are correct; it is not a second source of truth about the hardware, it is a way to run kernel LOGIC (schedules,
// barrier protocols, indexing, epilogues) without a GPU.
//
// Completion is LAZY: a TMA load is performed, and an issued MMA executed, only when some thread waits on the mbarrier
// the operation (or the tcgen05.commit after it) signals.  Code that reads shared memory or TMEM without waiting for the
// right barrier therefore sees 0xCD bytes / NaNs instead of accidentally-correct data.
// Checks (abort with a message): TMEM lane quarter of tcgen05.ld vs warp id, TMEM columns inside the allocation, TMA
// destinations / UMMA operands inside the CTA's dynamic shared memory and 128 B / 16 B aligned, barrier wait time-outs
// (= dead-locks, reported with the tag the kernel passes to mbar_wait).
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include "simt_emu.h"

#include <cstdarg>

#include <cuda.h>      // CUtensorMap, CUresult and the CU_TENSOR_MAP_* enums (types only; libcuda is not linked)

namespace emu {

[[noreturn]] inline void die(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  std::fprintf(stderr, "EMU FATAL [block (%u,%u,%u) thread %d]: ", ctx.bid.x, ctx.bid.y, ctx.bid.z, ctx.linear);
  std::vfprintf(stderr, fmt, ap);
  std::fprintf(stderr, "\n");
  va_end(ap);
  std::fflush(stderr);
  std::abort();
}

// ---------------------------------------------------------------------------------------------- tensor maps
struct TensorMap {          // lives inside the 128 opaque bytes of a CUtensorMap
  uint64_t magic;
  unsigned char* base;
  uint64_t dim[2];          // elements: inner, outer
  uint64_t stride1;         // bytes between outer rows
  uint32_t box[2];
  uint32_t esize;           // 2 (bf16) or 4 (fp32)
  uint32_t swizzle128;      // legacy flag: swizzle_bytes == 128
  uint32_t swizzle_bytes;   // 0 (none), 32, 64 or 128
};
static_assert(sizeof(TensorMap) <= sizeof(CUtensorMap), "emulated tensor map must fit the opaque CUtensorMap");
constexpr uint64_t TMAP_MAGIC = 0x454d55544d415031ull;

inline CUresult encode_tiled(CUtensorMap* out, CUtensorMapDataType dt, cuuint32_t rank, void* ptr, const cuuint64_t* dims,
                             const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il,
                             CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  // the argument checks the driver makes (the ones a wrong launcher would trip over)
  if (rank != 2 || il != CU_TENSOR_MAP_INTERLEAVE_NONE || estr[0] != 1 || estr[1] != 1) return CUDA_ERROR_INVALID_VALUE;
  const uint32_t es = dt == CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 ? 2 : dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 0;
  if (es == 0) return CUDA_ERROR_INVALID_VALUE;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (strides[0] & 15) != 0) return CUDA_ERROR_INVALID_VALUE;
  if (box[0] == 0 || box[1] == 0 || box[0] > 256 || box[1] > 256 || dims[0] == 0 || dims[1] == 0) return CUDA_ERROR_INVALID_VALUE;
  if ((box[0] * es) % 16 != 0) return CUDA_ERROR_INVALID_VALUE;
  const uint32_t swb = sw == CU_TENSOR_MAP_SWIZZLE_128B ? 128 : sw == CU_TENSOR_MAP_SWIZZLE_64B ? 64 : sw == CU_TENSOR_MAP_SWIZZLE_32B ? 32 : 0;
  if (sw != CU_TENSOR_MAP_SWIZZLE_NONE && swb == 0) return CUDA_ERROR_INVALID_VALUE;
  if (swb != 0 && box[0] * es > swb) return CUDA_ERROR_INVALID_VALUE;          // the inner box extent must fit the swizzle span
  if (strides[0] < dims[0] * es) return CUDA_ERROR_INVALID_VALUE;
  std::memset(out, 0, sizeof(*out));
  TensorMap* m = reinterpret_cast<TensorMap*>(out);
  m->magic = TMAP_MAGIC; m->base = static_cast<unsigned char*>(ptr); m->dim[0] = dims[0]; m->dim[1] = dims[1];
  m->stride1 = strides[0]; m->box[0] = box[0]; m->box[1] = box[1]; m->esize = es; m->swizzle128 = (sw == CU_TENSOR_MAP_SWIZZLE_128B); m->swizzle_bytes = swb;
  return CUDA_SUCCESS;
}

// ---------------------------------------------------------------------------------------------- shared-memory window
inline unsigned char* smem_origin() {      // the 1 KB boundary below the dynamic segment: shared "address" 0
  return reinterpret_cast<unsigned char*>(reinterpret_cast<uintptr_t>(ctx.blk->dyn) & ~uintptr_t(1023));
}
inline void check_smem(const void* p, size_t bytes, const char* what) {
  const unsigned char* q = static_cast<const unsigned char*>(p);
  if (q < ctx.blk->dyn || q + bytes > ctx.blk->dyn + ctx.blk->dyn_bytes)
    die("%s touches shared memory outside the CTA's dynamic segment (offset %td, %zu bytes, segment %zu bytes)", what,
        q - ctx.blk->dyn, bytes, ctx.blk->dyn_bytes);
}
// 128-byte swizzle: bits [4,7) of the address are XORed with bits [7,10)
inline uintptr_t swz128(uintptr_t a) { return a ^ (((a >> 7) & 7) << 4); }
// 64-byte / 32-byte swizzles: bits [4,6) resp. bit 4 are XORed with bits [7,9) resp. bit 7 (same source bits, fewer of them)
inline uintptr_t swz_bytes(uintptr_t a, uint32_t mode) {
  return mode == 128 ? swz128(a) : mode == 64 ? (a ^ (((a >> 7) & 3) << 4)) : mode == 32 ? (a ^ (((a >> 7) & 1) << 4)) : a;
}

// ---------------------------------------------------------------------------------------------- mbarrier
struct MBar { int32_t tx; uint16_t pending; uint8_t count; uint8_t phase_magic; };   // 8 bytes, in place in shared memory
static_assert(sizeof(MBar) == 8, "mbarrier state must fit the 64-bit object");
inline MBar* mb(uint64_t* bar, const char* what) {
  check_smem(bar, 8, what);
  MBar* m = reinterpret_cast<MBar*>(bar);
  if ((m->phase_magic & 0xFE) != 0xA0) die("%s on an uninitialised mbarrier", what);
  return m;
}
inline void mb_progress(MBar* m) {         // hw_mu held
  if (m->pending == 0 && m->tx == 0) { m->phase_magic ^= 1; m->pending = m->count; ctx.blk->hw_cv.notify_all(); }
}
inline void mb_arrive_locked(uint64_t* bar, int32_t expect_tx, const char* what) {
  MBar* m = mb(bar, what);
  if (m->pending == 0) die("%s: more arrivals than the barrier's count in one phase", what);
  m->tx += expect_tx;
  m->pending -= 1;
  mb_progress(m);
}
inline void mb_complete_tx_locked(uint64_t* bar, int32_t bytes) {
  MBar* m = mb(bar, "complete_tx");
  m->tx -= bytes;
  mb_progress(m);
}
// perform the lazy operations that signal `bar` (hw_mu held); returns true if anything ran
inline bool drain_for(uint64_t* bar) {
  Block& b = *ctx.blk;
  bool any = false;
  for (size_t i = 0; i < b.tma_pending.size();) {
    if (b.tma_pending[i].bar == bar) {
      auto op = std::move(b.tma_pending[i]);
      b.tma_pending.erase(b.tma_pending.begin() + static_cast<long>(i));
      op.run();
      any = true;
    } else {
      ++i;
    }
  }
  size_t upto = b.mma_fifo.size();
  for (size_t i = 0; i < b.mma_fifo.size(); ++i)
    if (b.mma_fifo[i].is_commit && b.mma_fifo[i].bar == bar) { upto = i; break; }
  if (upto < b.mma_fifo.size()) {
    for (size_t i = 0; i <= upto; ++i) {
      auto op = std::move(b.mma_fifo.front());
      b.mma_fifo.pop_front();
      op.run();
    }
    any = true;
  }
  return any;
}

}  // namespace emu

namespace fm {

static unsigned int g_fm_device_error = 0;

inline uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(static_cast<const unsigned char*>(p) - emu::smem_origin()); }
inline bool elect_one() { return (emu::ctx.linear & 31) == 0; }
inline void pdl_launch_dependents() {}
inline void pdl_wait() {}

// ----------------------------------------------------------------------------- mbarrier
inline void mbar_init(uint64_t* bar, uint32_t count) {
  emu::check_smem(bar, 8, "mbarrier.init");
  if (count == 0 || count > 255) emu::die("mbarrier.init with count %u", count);
  std::lock_guard<std::mutex> lk(emu::ctx.blk->hw_mu);
  emu::MBar* m = reinterpret_cast<emu::MBar*>(bar);
  m->tx = 0; m->pending = static_cast<uint16_t>(count); m->count = static_cast<uint8_t>(count); m->phase_magic = 0xA0;
}
inline void fence_mbar_init() {}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::mutex> lk(emu::ctx.blk->hw_mu);
  emu::mb_arrive_locked(bar, static_cast<int32_t>(bytes), "mbarrier.arrive.expect_tx");
}
inline void mbar_arrive(uint64_t* bar) {
  std::lock_guard<std::mutex> lk(emu::ctx.blk->hw_mu);
  emu::mb_arrive_locked(bar, 0, "mbarrier.arrive");
}
inline void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t tag) {
  emu::Block& b = *emu::ctx.blk;
  std::unique_lock<std::mutex> lk(b.hw_mu);
  static const int timeout_s = std::getenv("FM_EMU_TIMEOUT_S") ? std::atoi(std::getenv("FM_EMU_TIMEOUT_S")) : 120;
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(timeout_s);
  for (;;) {
    emu::MBar* m = emu::mb(bar, "mbarrier.try_wait");
    if (static_cast<uint32_t>(m->phase_magic & 1) != (parity & 1)) return;
    if (emu::drain_for(bar)) continue;
    if (b.hw_cv.wait_until(lk, deadline) == std::cv_status::timeout)
      emu::die("mbarrier wait timed out (dead-lock): tag 0x%x parity %u, barrier at smem offset %u: pending %u tx %d phase %u; "
               "%zu TMA loads and %zu MMA/commit operations issued but not yet waited for", tag, parity,
               smem_u32(bar), m->pending, m->tx, m->phase_magic & 1, b.tma_pending.size(), b.mma_fifo.size());
  }
}
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  std::lock_guard<std::mutex> lk(emu::ctx.blk->hw_mu);
  emu::drain_for(bar);
  return static_cast<uint32_t>(emu::mb(bar, "mbarrier.try_wait")->phase_magic & 1) != (parity & 1);
}

// ----------------------------------------------------------------------------- fences: ordering is sequentially consistent here
inline void fence_proxy_async_smem() {}
inline void tc_fence_before_sync() {}
inline void tc_fence_after_sync() {}

// ----------------------------------------------------------------------------- TMA
inline const emu::TensorMap* tmap(const CUtensorMap* m, const char* what) {
  const emu::TensorMap* t = reinterpret_cast<const emu::TensorMap*>(m);
  if (t->magic != emu::TMAP_MAGIC) emu::die("%s with a tensor map that was never encoded", what);
  return t;
}
inline void tma_prefetch_desc(const CUtensorMap* m) { (void)tmap(m, "prefetch.tensormap"); }
inline void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  const emu::TensorMap* t = tmap(m, "cp.async.bulk.prefetch.tensor");
  if (c0 < 0 || c1 < 0 || static_cast<uint64_t>(c0) >= t->dim[0] || static_cast<uint64_t>(c1) >= t->dim[1])
    emu::die("L2 prefetch box starts outside the tensor: (%d, %d) of (%llu, %llu)", c0, c1, (unsigned long long)t->dim[0], (unsigned long long)t->dim[1]);
}
inline void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  const emu::TensorMap t = *tmap(m, "cp.async.bulk.tensor.2d");
  const size_t row_bytes = static_cast<size_t>(t.box[0]) * t.esize, bytes = row_bytes * t.box[1];
  emu::check_smem(smem_dst, bytes, "TMA load destination");
  if ((reinterpret_cast<uintptr_t>(smem_dst) & 127) != 0) emu::die("TMA destination must be 128-byte aligned");
  if (t.swizzle128 && (reinterpret_cast<uintptr_t>(smem_dst) & 1023) != 0 && row_bytes == 128) {
    // legal on hardware (the pattern follows address bits), but every tile of this library is atom-aligned: a mis-set pointer
    emu::die("swizzled TMA destination is not 1024-byte aligned (offset %u)", smem_u32(smem_dst));
  }
  emu::check_smem(bar, 8, "TMA mbarrier");
  unsigned char* dst = static_cast<unsigned char*>(smem_dst);
  emu::Block& b = *emu::ctx.blk;
  std::lock_guard<std::mutex> lk(b.hw_mu);
  b.tma_pending.push_back({bar, false, [=] {
    for (uint32_t r = 0; r < t.box[1]; ++r) {
      const long long orow = static_cast<long long>(c1) + r;
      const bool row_ok = orow >= 0 && static_cast<uint64_t>(orow) < t.dim[1];
      for (uint32_t cb = 0; cb < row_bytes; cb += 16) {               // 16-byte chunks (box rows are multiples of 16 B)
        unsigned char tmp[16];
        for (uint32_t e = 0; e < 16 / t.esize; ++e) {
          const long long icol = static_cast<long long>(c0) + (cb / t.esize) + e;
          const bool ok = row_ok && icol >= 0 && static_cast<uint64_t>(icol) < t.dim[0];
          if (ok) std::memcpy(tmp + e * t.esize, t.base + static_cast<uint64_t>(orow) * t.stride1 + static_cast<uint64_t>(icol) * t.esize, t.esize);
          else std::memset(tmp + e * t.esize, 0, t.esize);
        }
        uintptr_t a = emu::swz_bytes(reinterpret_cast<uintptr_t>(dst) + r * row_bytes + cb, t.swizzle_bytes);
        std::memcpy(reinterpret_cast<void*>(a), tmp, 16);
      }
    }
    emu::mb_complete_tx_locked(bar, static_cast<int32_t>(bytes));
  }});
  b.hw_cv.notify_all();
}

// TMA store (bulk async group).  Performed immediately: program order already puts every generic write of the tile before it
// (the kernel's fence.proxy.async + barrier are no-ops here), and commit / wait_group have nothing left to wait for.
inline void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1, bool reduce_add = false) {
  const emu::TensorMap t = *tmap(m, "cp.async.bulk.tensor.2d (store)");
  if (reduce_add && t.esize != 4) emu::die("cp.reduce.async.bulk.tensor .add is modelled for fp32 tensor maps only");
  const size_t row_bytes = static_cast<size_t>(t.box[0]) * t.esize, bytes = row_bytes * t.box[1];
  emu::check_smem(smem_src, bytes, "TMA store source");
  if ((reinterpret_cast<uintptr_t>(smem_src) & 127) != 0) emu::die("TMA store source must be 128-byte aligned");
  if (t.swizzle_bytes != 0 && (reinterpret_cast<uintptr_t>(smem_src) & (t.swizzle_bytes * 8 - 1)) != 0)
    emu::die("swizzled TMA store source is not aligned to its %u-byte swizzle pattern (offset %u)", t.swizzle_bytes * 8, smem_u32(smem_src));
  const unsigned char* src = static_cast<const unsigned char*>(smem_src);
  std::lock_guard<std::mutex> lk(emu::ctx.blk->hw_mu);
  for (uint32_t r = 0; r < t.box[1]; ++r) {
    const long long orow = static_cast<long long>(c1) + r;
    if (orow < 0 || static_cast<uint64_t>(orow) >= t.dim[1]) continue;           // clipped
    for (uint32_t cb = 0; cb < row_bytes; cb += 16) {
      const uintptr_t a = emu::swz_bytes(reinterpret_cast<uintptr_t>(src) + r * row_bytes + cb, t.swizzle_bytes);
      for (uint32_t e = 0; e < 16 / t.esize; ++e) {
        const long long icol = static_cast<long long>(c0) + (cb / t.esize) + e;
        if (icol < 0 || static_cast<uint64_t>(icol) >= t.dim[0]) continue;       // clipped
        unsigned char* gdst = t.base + static_cast<uint64_t>(orow) * t.stride1 + static_cast<uint64_t>(icol) * t.esize;
        if (reduce_add) {        // element-wise fp32 add, atomic with respect to other CTAs' reductions
          float add;
          std::memcpy(&add, reinterpret_cast<const unsigned char*>(a) + e * t.esize, 4);
          std::atomic_ref<float> g(*reinterpret_cast<float*>(gdst));
          float cur = g.load(std::memory_order_relaxed);
          while (!g.compare_exchange_weak(cur, cur + add, std::memory_order_relaxed)) {}
        } else {
          std::memcpy(gdst, reinterpret_cast<const unsigned char*>(a) + e * t.esize, t.esize);
        }
      }
    }
  }
}
inline void tma_store_2d_s(const CUtensorMap* m, uint32_t s_src, int c0, int c1) { tma_store_2d(m, emu::smem_origin() + s_src, c0, c1); }
inline void tma_reduce_add_2d_s(const CUtensorMap* m, uint32_t s_src, int c0, int c1) { tma_store_2d(m, emu::smem_origin() + s_src, c0, c1, true); }
inline void tma_load_2d_s(uint32_t s_dst, const CUtensorMap* m, uint32_t s_bar, int c0, int c1) {
  tma_load_2d(emu::smem_origin() + s_dst, m, reinterpret_cast<uint64_t*>(emu::smem_origin() + s_bar), c0, c1);
}
inline void mbar_arrive_expect_tx_s(uint32_t s_bar, uint32_t bytes) { mbar_arrive_expect_tx(reinterpret_cast<uint64_t*>(emu::smem_origin() + s_bar), bytes); }
inline void mbar_wait_s(uint32_t s_bar, uint32_t parity, uint32_t tag) { mbar_wait(reinterpret_cast<uint64_t*>(emu::smem_origin() + s_bar), parity, tag); }
inline void bulk_commit() {}
inline void bulk_wait_read0() {}
inline void bulk_wait0() {}
inline void sts128(uint32_t saddr, const uint4& v) {
  unsigned char* p = emu::smem_origin() + saddr;
  emu::check_smem(p, 16, "st.shared.v4");
  std::memcpy(p, &v, 16);
}
inline uint4 lds128(uint32_t saddr) {
  const unsigned char* p = emu::smem_origin() + saddr;
  emu::check_smem(p, 16, "ld.shared.v4");
  uint4 v;
  std::memcpy(&v, p, 16);
  return v;
}

// ----------------------------------------------------------------------------- tcgen05 / TMEM
constexpr int EMU_TMEM_LANES = 128, EMU_TMEM_COLS = 512;
inline void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {      // called by all 32 lanes of one warp
  emu::check_smem(smem_slot, 4, "tcgen05.alloc result slot");
  if (ncols < 32 || ncols > 512 || (ncols & (ncols - 1)) != 0) emu::die("tcgen05.alloc of %u columns (power of two in [32, 512])", ncols);
  __syncwarp();
  if ((emu::ctx.linear & 31) == 0) {
    emu::Block& b = *emu::ctx.blk;
    std::lock_guard<std::mutex> lk(b.hw_mu);
    if (b.tmem.empty()) b.tmem.assign(static_cast<size_t>(EMU_TMEM_LANES) * EMU_TMEM_COLS, std::nanf(""));
    const uint32_t units = ncols / 32;
    uint32_t col = 0xffffffffu;
    for (uint32_t u0 = 0; u0 + units <= 16; u0 += units) {
      const uint32_t mask = (units == 16 ? 0xffffu : ((1u << units) - 1u)) << u0;
      if ((b.tmem_alloc_mask & mask) == 0) { b.tmem_alloc_mask |= mask; col = u0 * 32; break; }
    }
    if (col == 0xffffffffu) emu::die("tcgen05.alloc: no %u free TMEM columns", ncols);
    *smem_slot = col;
  }
  __syncwarp();
}
inline void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  __syncwarp();
  if ((emu::ctx.linear & 31) == 0) {
    emu::Block& b = *emu::ctx.blk;
    std::lock_guard<std::mutex> lk(b.hw_mu);
    const uint32_t col = taddr & 0xffff, units = ncols / 32;
    const uint32_t mask = (units == 16 ? 0xffffu : ((1u << units) - 1u)) << (col / 32);
    if ((taddr >> 16) != 0 || (b.tmem_alloc_mask & mask) != mask) emu::die("tcgen05.dealloc of columns that are not allocated (addr 0x%x, %u)", taddr, ncols);
    b.tmem_alloc_mask &= ~mask;
  }
  __syncwarp();
}
inline void tmem_check_cols(uint32_t col, uint32_t n, const char* what) {
  emu::Block& b = *emu::ctx.blk;
  if (col + n > 512) emu::die("%s: TMEM columns [%u, %u) out of range", what, col, col + n);
  for (uint32_t c = col / 32; c <= (col + n - 1) / 32; ++c)
    if (!(b.tmem_alloc_mask & (1u << c))) emu::die("%s: TMEM columns [%u, %u) are not allocated", what, col, col + n);
}

struct UmmaDesc { uint32_t start, lbo, sbo; };
inline UmmaDesc umma_decode(uint64_t d) {
  if (((d >> 61) & 7) != 2 || ((d >> 46) & 3) != 1) emu::die("UMMA shared-memory descriptor: only SWIZZLE_128B, version 1 is modelled (0x%llx)", (unsigned long long)d);
  if (((d >> 49) & 7) != 0 || ((d >> 52) & 1) != 0) emu::die("UMMA descriptor with base offset / LBO mode set (0x%llx)", (unsigned long long)d);
  return {static_cast<uint32_t>(d & 0x3FFF) << 4, static_cast<uint32_t>((d >> 16) & 0x3FFF) << 4, static_cast<uint32_t>((d >> 32) & 0x3FFF) << 4};
}
// gather a [rows x 16] bf16 operand slice (one K = 16 step) into dense fp32: out[r * 16 + k]
inline void umma_gather(const UmmaDesc& d, bool mn_major, int rows, float* out) {
  unsigned char* org = emu::smem_origin();
  auto rd = [&](uint32_t off) {
    const uintptr_t a = emu::swz128(reinterpret_cast<uintptr_t>(org) + d.start + off);
    emu::check_smem(reinterpret_cast<const void*>(a), 2, "tcgen05.mma operand");
    return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(a));
  };
  if ((d.start & 15) != 0) emu::die("UMMA operand start address not 16-byte aligned");
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < 16; ++k) {
      const uint32_t off = mn_major ? (r / 64) * d.lbo + (k / 8) * d.sbo + (k % 8) * 128 + (r % 64) * 2      // MN contiguous: 64-element
                                    : (r / 8) * d.sbo + (r % 8) * 128 + k * 2;                                // chunks LBO apart; K contiguous
      out[r * 16 + k] = rd(off);
    }
}
inline void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const int M = static_cast<int>((idesc >> 24) & 0x1F) << 4, N = static_cast<int>((idesc >> 17) & 0x3F) << 3;
  const bool a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
  if ((idesc & 0x7FFF) != ((1u << 4) | (1u << 7) | (1u << 10))) emu::die("instruction descriptor: only bf16 x bf16 -> fp32, dense, no negate is modelled (0x%x)", idesc);
  if (M != 128 || N < 16 || N > 256 || N % 16 != 0) emu::die("tcgen05.mma cta_group::1: M must be 128 (lane = row) and N a multiple of 16 in [16, 256]: M=%d N=%d", M, N);
  if ((d_tmem >> 16) != 0) emu::die("tcgen05.mma accumulator address must name lane 0 (0x%x)", d_tmem);
  const uint32_t col = d_tmem & 0xffff;
  tmem_check_cols(col, static_cast<uint32_t>(N), "tcgen05.mma accumulator");
  const UmmaDesc da = umma_decode(a_desc), db = umma_decode(b_desc);
  emu::Block& b = *emu::ctx.blk;
  std::lock_guard<std::mutex> lk(b.hw_mu);
  b.mma_fifo.push_back({nullptr, false, [=] {
    static thread_local std::vector<float> A, B;
    A.resize(128 * 16); B.resize(static_cast<size_t>(N) * 16);
    umma_gather(da, a_mn, 128, A.data());
    umma_gather(db, b_mn, N, B.data());
    float* T = emu::ctx.blk->tmem.data();
    for (int m = 0; m < 128; ++m) {
      const float* a = A.data() + m * 16;
      float* trow = T + static_cast<size_t>(m) * EMU_TMEM_COLS + col;
      for (int n = 0; n < N; ++n) {
        const float* bb = B.data() + n * 16;
        float acc = 0.0f;
        for (int k = 0; k < 16; ++k) acc += a[k] * bb[k];
        trow[n] = accumulate ? trow[n] + acc : acc;
      }
    }
  }});
}
inline void umma_commit(uint64_t* bar) {
  emu::check_smem(bar, 8, "tcgen05.commit mbarrier");
  emu::Block& b = *emu::ctx.blk;
  std::lock_guard<std::mutex> lk(b.hw_mu);
  b.mma_fifo.push_back({bar, true, [=] { emu::mb_arrive_locked(bar, 0, "tcgen05.commit arrival"); }});
  b.hw_cv.notify_all();
}
inline void tmem_ld_wait() {}
inline void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  const int warp = emu::ctx.linear >> 5, lane = emu::ctx.linear & 31;
  const uint32_t lane_base = taddr >> 16, col = taddr & 0xffff;
  if (lane_base != static_cast<uint32_t>((warp & 3) * 32))
    emu::die("tcgen05.ld 32x32b: warp %d may only touch TMEM lanes [%d, %d), address names lane %u", warp, (warp & 3) * 32, (warp & 3) * 32 + 32, lane_base);
  tmem_check_cols(col, 32, "tcgen05.ld");
  const float* src = emu::ctx.blk->tmem.data() + static_cast<size_t>(lane_base + lane) * EMU_TMEM_COLS + col;
  std::memcpy(r, src, 32 * sizeof(float));
}

inline uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// the three raw global-memory accesses of gemm_tc.cuh
inline int ld_acquire_gpu(const int* p) { std::this_thread::yield(); return std::atomic_ref<int>(*const_cast<int*>(p)).load(std::memory_order_acquire); }
inline void st_release_gpu(int* p, int v) { std::atomic_ref<int>(*p).store(v, std::memory_order_release); }
inline float4 ld_global_cg_f4(const void* p) { return *static_cast<const float4*>(p); }

}  // namespace fm

// the two C++ convenience overloads of cuda_runtime.h that only exist under nvcc
template <class T>
inline cudaError_t cudaFuncSetAttribute(T* fn, enum cudaFuncAttribute attr, int value) {
  return ::cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), attr, value);
}
template <class T>
inline cudaError_t cudaMemcpyFromSymbol(void* dst, const T& symbol, size_t count, size_t offset = 0,
                                        enum cudaMemcpyKind kind = cudaMemcpyDeviceToHost) {
  return ::cudaMemcpyFromSymbol(dst, static_cast<const void*>(&symbol), count, offset, kind);
}
