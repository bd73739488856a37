// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Hand-written for this project; no CUTLASS dependency.
// FM_HOST_EMU: tests/cpu_harness compiles this library as host code with g++ and runs the kernels thread-per-thread on
// the CPU (simt_emu.h); everything below that is sm_100a PTX (mbarrier, TMA, tcgen05, TMEM) is then replaced by the
// functional model in tests/cpu_harness/tc_emu.h (same function names), and the two approx math helpers get host
// equivalents.  Test infrastructure only: the shipped library is never built that way.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef FM_HOST_EMU
#include "../../tests/cpu_harness/tc_emu.h"
#else
#define FM_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

namespace fm {

#ifndef FM_HOST_EMU
// Device-side error word: set by the mbarrier watchdog before trapping, so the host can say *which* wait hung.
static __device__ unsigned int g_fm_device_error = 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- programmatic dependent launch
// Every kernel of this library starts with pdl_launch_dependents() (the next kernel of the stream may be scheduled as
// soon as all CTAs of this grid are resident or done) and calls pdl_wait() after its prologue (barrier init, TMEM
// allocation, descriptor prefetch) and BEFORE its first access to global memory: the wait returns once every
// prerequisite grid has completed and flushed.  Both are no-ops unless the launch carried
// cudaLaunchAttributeProgrammaticStreamSerialization (FM_OPT_PDL).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait with a watchdog: a protocol bug must surface as a trapped kernel (launch failure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      atomicExch(&g_fm_device_error, 0x80000000u | tag);
      __threadfence_system();
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------- fences
__device__ __forceinline__ void fence_proxy_async_smem() {  // generic-proxy smem writes -> visible to async proxy (TMA/UMMA)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// L2 prefetch of a tensor tile (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1) : "memory");
}
// TMA store of a shared-memory tile (bulk async group); the tile must have been made visible to the async proxy
// (fence_proxy_async_smem by every writing thread, then a warp / CTA barrier) before the issuing thread gets here.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be overwritten
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }            // writes are complete
// the same operations addressed by 32-bit shared addresses (epilogue staging tiles / per-warp barriers keep one register each)
__device__ __forceinline__ void tma_store_2d_s(const CUtensorMap* m, uint32_t s_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(s_src), "r"(c0), "r"(c1)
               : "memory");
}
// TMA reduce-add of a shared-memory tile into global memory (element type from the tensor map: fp32 here): parallel split-K
__device__ __forceinline__ void tma_reduce_add_2d_s(const CUtensorMap* m, uint32_t s_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(s_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_s(uint32_t s_dst, const CUtensorMap* m, uint32_t s_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(s_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(s_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_s(uint32_t s_bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_s(uint32_t s_bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(s_bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_s(uint32_t s_bar, uint32_t parity, uint32_t tag) {     // same watchdog as mbar_wait
  if (mbar_try_wait_s(s_bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_s(s_bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      atomicExch(&g_fm_device_error, 0x80000000u | tag);
      __threadfence_system();
      __trap();
    }
  }
}
// 128-bit shared-memory accesses by 32-bit shared address (the compiler cannot always prove the address space of a carved
// dynamic-smem pointer and then emits generic LD/ST)
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ----------------------------------------------------------------------------- tcgen05 / TMEM
// Executed by ONE full warp. Writes the TMEM base address to *smem_slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; bf16 inputs, fp32 accumulation. Issued by a single thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (lane_base + t), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// ----------------------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
// | base_offset [49,52) | lbo_mode [52] | layout_type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D:
// c_format=F32 (1<<4) | a_format=BF16 (1<<7) | b_format=BF16 (1<<10) | a_major<<15 | b_major<<16 | (N>>3)<<17 | (M>>4)<<24
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

#endif  // !FM_HOST_EMU

// ----------------------------------------------------------------------------- small math / packing
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Activations of utils.py:36-40 (reference) and their derivatives. act: 0 = gelu (exact erf form), 1 = sqrelu, 2 = relu.
__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == 0) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
  float r = fmaxf(x, 0.0f);
  return act == 1 ? r * r : r;
}
// returns f'(x); *f receives f(x)
__device__ __forceinline__ float act_bwd(float x, int act, float* f) {
  if (act == 0) {
    float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    *f = x * cdf;
    return cdf + x * pdf;
  }
  float r = fmaxf(x, 0.0f);
  if (act == 1) { *f = r * r; return 2.0f * r; }
  *f = r;
  return x > 0.0f ? 1.0f : 0.0f;
}


// Fast GELU / GELU' for GEMM epilogues (outputs are rounded to bf16; the erf approximation below is accurate to
// 1.5e-7 absolute, Abramowitz-Stegun 7.1.26).  Branch-free, 2 MUFU + ~16 FP32 ops for BOTH value and derivative:
//   q(x)   = 0.5 * erfc(|x|/sqrt2) = 0.5 * poly(t) * t * exp(-x^2/2),  t = 1 / (1 + p |x| / sqrt2)
//   Phi(x) = 0.5 + sign(x) (0.5 - q)
//   gelu   = x Phi(x) = max(x, 0) - |x| q          gelu' = Phi(x) + x exp(-x^2/2) / sqrt(2 pi)
#ifndef FM_HOST_EMU
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#else
__device__ __forceinline__ float rcp_approx(float x) { return 1.0f / x; }
__device__ __forceinline__ float ex2_approx(float x) { return exp2f(x); }
#endif
__device__ __forceinline__ float gelu_q(float x, float* e_out) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f));
  const float e = ex2_approx(x * x * (-0.5f * 1.4426950408889634f));          // exp(-x^2/2)
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  *e_out = e;
  return p * t * e;
}
__device__ __forceinline__ float act_fwd_fast(float x, int act) {
  if (act == 0) { float e; const float q = gelu_q(x, &e); return fmaf(-fabsf(x), q, fmaxf(x, 0.0f)); }
  const float r = fmaxf(x, 0.0f);
  return act == 1 ? r * r : r;
}
__device__ __forceinline__ float act_bwd_fast(float x, int act, float* f) {
  if (act == 0) {
    float e;
    const float q = gelu_q(x, &e);
    *f = fmaf(-fabsf(x), q, fmaxf(x, 0.0f));
    const float cdf = 0.5f + copysignf(0.5f - q, x);
    return fmaf(x * 0.39894228040143267794f, e, cdf);
  }
  const float r = fmaxf(x, 0.0f);
  if (act == 1) { *f = r * r; return 2.0f * r; }
  *f = r;
  return x > 0.0f ? 1.0f : 0.0f;
}

template <int ACT> __device__ __forceinline__ float act_fwd_t(float x) { return act_fwd_fast(x, ACT); }
template <int ACT> __device__ __forceinline__ float act_bwd_t(float x, float* f) { return act_bwd_fast(x, ACT, f); }

// ----------------------------------------------------------------------------- packed fp32 (sm_100: FFMA2 / FMUL2 / FADD2)
// Blackwell issues two IEEE fp32 operations per lane with one instruction (fma/mul/add.rn.f32x2 on a 64-bit register
// pair; SASS FFMA2 accepts |x| / -x operand modifiers and immediates).  The GEMM epilogues are bound by issue slots and
// FMA-pipe cycles (K = 768 leaves ~6 k cycles of MMA per 128 x 256 tile for 32 k outputs), so their arithmetic runs on
// pairs of adjacent accumulator columns.  Results are bit-identical to the same expression written with fmaf/mul/add.
// MEASURED (round 2, profiles/r02_validate_next/gemm_trace_next*.txt): the packed forms are NOT faster on B200 - the ACT tile
// epilogue took 14.8k / 11.3k / 8.8k cycles packed against 13.5k / 9.8k / 7.6k scalar, DACT 21k against 19k - so the scalar
// epilogues are the default; -DFM_EPI_F32X2=1 (FM_B200_NVCC_FLAGS) rebuilds the packed ones for an A/B.
#ifndef FM_EPI_F32X2
#define FM_EPI_F32X2 0
#endif
#ifndef FM_HOST_EMU
__device__ __forceinline__ unsigned long long f2_pack(float2 a) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
__device__ __forceinline__ float2 f2_unpack(unsigned long long r) { float2 a; asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r)); return a; }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
#else
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
#endif
__device__ __forceinline__ float2 f2(float c) { return make_float2(c, c); }
__device__ __forceinline__ float f_or_sign(float h, float x) {       // h >= 0: copysign(h, x) as one LOP3
  return __uint_as_float(__float_as_uint(h) | (__float_as_uint(x) & 0x80000000u));
}

// GELU (exact-erf form, same Abramowitz-Stegun evaluation as gelu_q) of two values: 7 FMA-pipe + 2 MUFU + 3 ALU issue
// slots per element for value AND derivative (scalar: 14 + 2 + 2).  WITH_D = false drops the derivative (inference).
template <bool WITH_D>
__device__ __forceinline__ void gelu2(float2 x, float2& f, float2& d) {
  const float2 nax = make_float2(-fabsf(x.x), -fabsf(x.y));                       // folds into FFMA2's -|x| operand modifier
  const float2 ta = fma2(nax, f2(-0.3275911f * 0.70710678118654752440f), f2(1.0f));
  const float2 t = make_float2(rcp_approx(ta.x), rcp_approx(ta.y));
  const float2 ea = mul2(mul2(x, x), f2(-0.5f * 1.4426950408889634f));
  const float2 e = make_float2(ex2_approx(ea.x), ex2_approx(ea.y));                // exp(-x^2/2)
  float2 p = fma2(f2(0.5f * 1.061405429f), t, f2(0.5f * -1.453152027f));
  p = fma2(p, t, f2(0.5f * 1.421413741f));
  p = fma2(p, t, f2(0.5f * -0.284496736f));
  p = fma2(p, t, f2(0.5f * 0.254829592f));
  const float2 q = mul2(mul2(p, t), e);                                            // 0.5 erfc(|x|/sqrt2), in [0, 0.5]
  f = fma2(nax, q, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));               // max(x,0) - |x| q
  if constexpr (WITH_D) {
    const float2 h = fma2(q, f2(-1.0f), f2(0.5f));                                 // 0.5 - q >= 0
    const float2 cdf = add2(make_float2(f_or_sign(h.x, x.x), f_or_sign(h.y, x.y)), f2(0.5f));
    d = fma2(mul2(x, f2(0.39894228040143267794f)), e, cdf);                        // Phi(x) + x phi(x)
  }
}
// v <- act(v) for 32 accumulator columns; d <- act'(v) when WITH_D
template <int ACT, bool WITH_D>
__device__ __forceinline__ void act32(float (&v)[32], float (&d)[32]) {
  if constexpr (ACT == 0 && FM_EPI_F32X2) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float2 f, g = make_float2(0.0f, 0.0f);
      gelu2<WITH_D>(make_float2(v[j], v[j + 1]), f, g);
      v[j] = f.x; v[j + 1] = f.y;
      if constexpr (WITH_D) { d[j] = g.x; d[j + 1] = g.y; }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if constexpr (WITH_D) d[j] = act_bwd_t<ACT>(v[j], &v[j]);
      else v[j] = act_fwd_t<ACT>(v[j]);
    }
  }
}
// v <- a * v (+ c): scale / gate / residual on pairs
__device__ __forceinline__ void scale32(float (&v)[32], float a) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
#if FM_EPI_F32X2
    const float2 r = mul2(make_float2(v[j], v[j + 1]), f2(a));
    v[j] = r.x; v[j + 1] = r.y;
#else
    v[j] *= a; v[j + 1] *= a;
#endif
  }
}
__device__ __forceinline__ void axpy32(float (&v)[32], float a, const float (&c)[32]) {      // v <- a*v + c
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
#if FM_EPI_F32X2
    const float2 r = fma2(f2(a), make_float2(v[j], v[j + 1]), make_float2(c[j], c[j + 1]));
    v[j] = r.x; v[j + 1] = r.y;
#else
    v[j] = fmaf(a, v[j], c[j]); v[j + 1] = fmaf(a, v[j + 1], c[j + 1]);
#endif
  }
}
__device__ __forceinline__ void scale_mul32(float (&v)[32], float a, const float (&c)[32]) { // v <- (a*v) * c
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
#if FM_EPI_F32X2
    const float2 r = mul2(mul2(f2(a), make_float2(v[j], v[j + 1])), make_float2(c[j], c[j + 1]));
    v[j] = r.x; v[j + 1] = r.y;
#else
    v[j] = a * v[j] * c[j]; v[j + 1] = a * v[j + 1] * c[j + 1];
#endif
  }
}
__device__ __forceinline__ float dot32(const float (&v)[32], const float (&c)[32]) {        // sum_j v_j c_j
#if FM_EPI_F32X2
  float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int j = 0; j < 32; j += 2) acc = fma2(make_float2(v[j], v[j + 1]), make_float2(c[j], c[j + 1]), acc);
  return acc.x + acc.y;
#else
  float acc = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j) acc = fmaf(v[j], c[j], acc);
  return acc;
#endif
}

}  // namespace fm
