// tcgen05 attention cores of the gated cross-attention block (gated_cross_attention.py:95-124 of the reference).
//
// Token i of sample b attends only to the 64 latents of image text_time[b,i] (1-based): one live 64-key slab per
// token.  Forward, per CTA = (128-token tile, head):
//     S = Q K_j^T          tcgen05.mma 128x64x64   (Q, K_j staged by TMA, 128B swizzle; S in TMEM)
//     P = softmax(S)       one thread per row straight out of TMEM (tcgen05.ld 32x32b), fp32, no shuffles
//     O += P V_j           tcgen05.mma 128x64x64   (P re-staged bf16 in swizzled smem; V_j MN-major; O in TMEM)
// Rows with text_time == 0 keep P = 0 (exact zero output); rows with text_time > n_media use the uniform
// P = 1/(n_media*64) against every slab — precisely what the reference's fully-masked softmax degenerates to.
// Backward, per CTA = (head, sample), loops slabs and token tiles and keeps dK_j/dV_j in TMEM:
//     S, dP = dO V_j^T -> dS = g P (dP - rowsum(P dP)) -> dQ = dS K_j ;  [dK_j ; dV_j] += [dS ; P]^T [Q | dO]
// the last product is ONE 128x128x128 MMA on MN-major views of the staged tiles (no transposes anywhere).
#pragma once
#include "ptx.cuh"

namespace fm {

// 16-byte chunk c of row r in a [rows][128 B] tile with the 128B TMA/UMMA swizzle (tile base 1024-aligned)
__device__ __forceinline__ uint32_t sw128(int r, int c) { return static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
  uint32_t r0[32], r1[32];
  tmem_ld_32x32(taddr, r0);
  tmem_ld_32x32(taddr + 32, r1);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
}
// write one 64-element row (bf16) of a K-major swizzled operand tile
__device__ __forceinline__ void put_row_bf16(uint8_t* tile, int r, const float (&v)[64]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 u;
    u.x = pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]); u.y = pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]);
    u.z = pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]); u.w = pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]);
    *reinterpret_cast<uint4*>(tile + sw128(r, c)) = u;
  }
}
// coalesced copy of a staged [nrows][128 B] swizzled tile to global rows (8 lanes per row); row r goes to
// dst + r*ld_bytes when keep(r) is true.
template <typename Keep>
__device__ __forceinline__ void flush_rows(const uint8_t* tile, uint8_t* dst, size_t ld_bytes, int nrows, Keep keep) {
  const int cc = threadIdx.x & 7;
  for (int r = threadIdx.x >> 3; r < nrows; r += blockDim.x >> 3) {
    if (keep(r)) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(r) * ld_bytes + cc * 16) = *reinterpret_cast<const uint4*>(tile + sw128(r, cc));
  }
}

struct XTcArgs {
  const int* tt;            // [B, S]
  __nv_bfloat16* o;         // [B*S, H*64]
  int B, S, H, n_media;
};

constexpr int XTC_FWD_SMEM = 16384 + 8192 + 8192 + 16384 + 1024 /*align*/ + 128 /*barriers*/;

__global__ void __launch_bounds__(128) xattn_core_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                const __grid_constant__ CUtensorMap tmKV, const XTcArgs a) {
  pdl_launch_dependents();
  pdl_wait();      // text_time / lse are read right away: no prologue to overlap in these short kernels
  FM_DYN_SMEM(uint8_t, xs_raw);
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xs_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = sm;                 // [128 tok][64 dh]   (later reused to stage O)
  uint8_t* sK = sQ + 16384;         // [64 keys][64 dh]
  uint8_t* sV = sK + 8192;          // [64 keys][64 dh]
  uint8_t* sP = sV + 8192;          // [128 tok][64 keys]
  // One mbarrier per producer step.  S = QK^T and O += PV used to share one barrier: between the PV completion of slab j
  // and the S completion of slab j+1 there is no CTA-wide sync, so a thread still polling for the former could be lapped
  // by a whole phase and wait forever (found by the host emulator, where threads really do get descheduled that long).
  // With separate barriers every phase flip is separated from the next one by a __syncthreads all waiters have passed.
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sP + 16384);
  uint64_t* bar_s = bar_load + 1;
  uint64_t* bar_o = bar_s + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o + 1);
  int* s_red = reinterpret_cast<int*>(tmem_slot + 1);   // [3] jmin, jmax, any_uniform

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.z, h = blockIdx.y, t0 = blockIdx.x * 128;
  const int t = t0 + tid;
  const bool valid = t < a.S;
  const int mytt = valid ? a.tt[b * a.S + t] : 0;
  const int HD = a.H * 64;

  if (tid == 0) {
    mbar_init(bar_load, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1); fence_mbar_init();
    s_red[0] = 0x7fffffff; s_red[1] = -1; s_red[2] = 0;
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem + (static_cast<uint32_t>(warp * 32) << 16);        // this warp's lane quarter, column 0
  const uint32_t tO = tS + 64;

  if (valid && mytt >= 1 && mytt <= a.n_media) { atomicMin(&s_red[0], mytt - 1); atomicMax(&s_red[1], mytt - 1); }
  if (valid && mytt > a.n_media) s_red[2] = 1;
  __syncthreads();
  int jlo = s_red[0], jhi = s_red[1];
  if (s_red[2]) { jlo = 0; jhi = a.n_media - 1; }
  const bool uniform = valid && mytt > a.n_media;
  const float p_uniform = 1.0f / static_cast<float>(a.n_media * 64);

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false, false);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
  uint32_t ph_load = 0, ph_mma = 0;      // bar_s and bar_o complete one phase per slab each: one shared parity
  bool first = true;
  for (int j = jlo; j <= jhi; ++j) {
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_load, first ? (16384 + 8192 + 8192) : (8192 + 8192));
      if (first) tma_load_2d(sQ, &tmQ, bar_load, h * 64, b * a.S + t0);
      const int krow = (b * a.n_media + j) * 64;
      tma_load_2d(sK, &tmKV, bar_load, h * 64, krow);
      tma_load_2d(sV, &tmKV, bar_load, HD + h * 64, krow);
    }
    mbar_wait(bar_load, ph_load, 0x600); ph_load ^= 1;
    if (tid == 0) {
      tc_fence_after_sync();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, umma_smem_desc_sw128(smem_u32(sQ) + k * 32, 0, 1024), umma_smem_desc_sw128(smem_u32(sK) + k * 32, 0, 1024),
                  idesc_s, k > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, ph_mma, 0x601);
    tc_fence_after_sync();
    {
      float s[64];
      tmem_ld64(tS, s);
      if (valid && mytt == j + 1) {
        float m = s[0];
#pragma unroll
        for (int k = 1; k < 64; ++k) m = fmaxf(m, s[k]);
        float l = 0.0f;
#pragma unroll
        for (int k = 0; k < 64; ++k) { s[k] = __expf(s[k] - m); l += s[k]; }
        const float inv = 1.0f / l;
#pragma unroll
        for (int k = 0; k < 64; ++k) s[k] *= inv;
      } else {
        const float fill = uniform ? p_uniform : 0.0f;
#pragma unroll
        for (int k = 0; k < 64; ++k) s[k] = fill;
      }
      put_row_bf16(sP, tid, s);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem + 64, umma_smem_desc_sw128(smem_u32(sP) + k * 32, 0, 1024),
                  umma_smem_desc_sw128(smem_u32(sV) + k * 2048, 8192, 1024), idesc_o, (!first || k > 0) ? 1u : 0u);
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, ph_mma, 0x602); ph_mma ^= 1;
    tc_fence_after_sync();
    first = false;
  }
  // ---- output: O row -> bf16 -> staged in sQ -> coalesced store
  {
    float o[64];
    if (!first) tmem_ld64(tO, o);
    else {
#pragma unroll
      for (int k = 0; k < 64; ++k) o[k] = 0.0f;
    }
    put_row_bf16(sQ, tid, o);
  }
  tc_fence_before_sync();
  __syncthreads();
  {
    const int S = a.S;
    uint8_t* dst = reinterpret_cast<uint8_t*>(a.o + (static_cast<size_t>(b) * a.S + t0) * HD + h * 64);
    flush_rows(sQ, dst, static_cast<size_t>(HD) * 2, 128, [=](int r) { return t0 + r < S; });
  }
  tc_fence_after_sync();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

struct XTcBwdArgs {
  const int* tt;
  const float* gate;         // optional device scalar; dO is multiplied by tanh(*gate)
  const __nv_bfloat16* d_o;  // [B*S, H*64] (read directly only for the uniform-row prelude)
  __nv_bfloat16* dq;         // [B*S, H*64]
  __nv_bfloat16* dkv;        // [B*n_media*64, 2*H*64]
  float q_scale;
  int B, S, H, n_media;
  int tmem_compact;          // 1: 256 TMEM columns (dQ reuses the S columns), so two CTAs of an SM run side by side
};

constexpr int XTC_BWD_SMEM = 4 * 16384 + 2 * 8192 + 1024 /*align*/ + 512 /*barriers, usum*/;

__global__ void __launch_bounds__(128) xattn_core_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                const __grid_constant__ CUtensorMap tmDO,
                                                                const __grid_constant__ CUtensorMap tmKV, const XTcBwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();      // text_time / lse are read right away: no prologue to overlap in these short kernels
  FM_DYN_SMEM(uint8_t, xb_raw);
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xb_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = sm;                  // [128 tok][64 dh]   \ adjacent: stacked MN-major B operand [Q | dO]
  uint8_t* sDO = sQ + 16384;         // [128 tok][64 dh]   /
  uint8_t* sDS = sDO + 16384;        // [128 tok][64 keys] \ adjacent: stacked MN-major A operand [dS ; P]
  uint8_t* sP = sDS + 16384;         // [128 tok][64 keys] /
  uint8_t* sK = sP + 16384;          // [64 keys][64 dh]
  uint8_t* sV = sK + 8192;
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(sV + 8192);
  uint64_t* bar_q = bar_kv + 1;
  uint64_t* bar_mma = bar_q + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  float* usum = reinterpret_cast<float*>(tmem_slot + 2);   // [64]

  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, b = blockIdx.y;
  const int HD = a.H * 64;
  const float g = a.gate ? tanhf(__ldg(a.gate)) : 1.0f;
  const int nkeys = a.n_media * 64;

  if (tid == 0) {
    mbar_init(bar_kv, 1); mbar_init(bar_q, 1); mbar_init(bar_mma, 1); fence_mbar_init();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmKV);
  }
  if (tid < 64) usum[tid] = 0.0f;
  // TMEM map (fp32 columns).  Wide: S 0 | dP 64 | dQ 128 | [dK;dV] 256 -> 512 columns, i.e. the whole SM: a second resident CTA
  // blocks in tcgen05.alloc until the first one is done.  Compact: dQ takes over the S columns (S is dead once every thread has
  // read its row, which the __syncthreads before the dQ MMA guarantees; the next tile's S MMA is issued after the
  // __syncthreads that follows the dQ read-out) and [dK;dV], which accumulates across token tiles, sits at 128 -> 256 columns.
  const uint32_t ncols = a.tmem_compact ? 256u : 512u;
  const uint32_t cDQ = a.tmem_compact ? 0u : 128u, cDKV = a.tmem_compact ? 128u : 256u;
  if (warp == 0) tmem_alloc(tmem_slot, ncols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  const uint32_t tS = tmem + lane_base, tDP = tS + 64, tDQ = tS + cDQ, tDKV = tS + cDKV;

  // ---- prelude: rows that get no gradient through q (tt == 0 or tt > n_media), uniform-row dV term
  for (int tb = 0; tb < a.S; tb += 128) {
    const int t = tb + tid;
    if (t < a.S) {
      const int mytt = a.tt[b * a.S + t];
      if (mytt < 1 || mytt > a.n_media) {
        const size_t roff = static_cast<size_t>(b * a.S + t) * HD + h * 64;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(a.dq + roff + c * 8) = z;
        if (mytt > a.n_media) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(a.d_o + roff + c * 8);
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_bf16x2(w4[e]);
              atomicAdd(&usum[c * 8 + e * 2], f.x * g / nkeys);
              atomicAdd(&usum[c * 8 + e * 2 + 1], f.y * g / nkeys);
            }
          }
        }
      }
    }
  }
  __syncthreads();

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false, false);     // S = Q K^T, dP = dO V^T
  constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64, false, true);     // dQ = dS K      (B = K MN-major)
  constexpr uint32_t idesc_kv = umma_idesc_bf16(128, 128, true, true);     // [dS;P]^T [Q|dO]
  uint32_t ph_kv = 0, ph_q = 0, ph_mma = 0;

  for (int j = 0; j < a.n_media; ++j) {
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_kv, 8192 + 8192);
      const int krow = (b * a.n_media + j) * 64;
      tma_load_2d(sK, &tmKV, bar_kv, h * 64, krow);
      tma_load_2d(sV, &tmKV, bar_kv, HD + h * 64, krow);
    }
    mbar_wait(bar_kv, ph_kv, 0x610); ph_kv ^= 1;
    bool first_tile = true;
    for (int tb = 0; tb < a.S; tb += 128) {
      const int t = tb + tid;
      const bool active = (t < a.S) && (a.tt[b * a.S + t] == j + 1);
      if (!__syncthreads_or(active)) continue;
      if (tid == 0) {
        mbar_arrive_expect_tx(bar_q, 16384 + 16384);
        tma_load_2d(sQ, &tmQ, bar_q, h * 64, b * a.S + tb);
        tma_load_2d(sDO, &tmDO, bar_q, h * 64, b * a.S + tb);
      }
      mbar_wait(bar_q, ph_q, 0x611); ph_q ^= 1;
      if (tid == 0) {
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, umma_smem_desc_sw128(smem_u32(sQ) + k * 32, 0, 1024), umma_smem_desc_sw128(smem_u32(sK) + k * 32, 0, 1024),
                    idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + 64, umma_smem_desc_sw128(smem_u32(sDO) + k * 32, 0, 1024), umma_smem_desc_sw128(smem_u32(sV) + k * 32, 0, 1024),
                    idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, ph_mma, 0x612); ph_mma ^= 1;
      tc_fence_after_sync();
      {
        float p[64];
        tmem_ld64(tS, p);
        if (active) {
          float m = p[0];
#pragma unroll
          for (int k = 1; k < 64; ++k) m = fmaxf(m, p[k]);
          float l = 0.0f;
#pragma unroll
          for (int k = 0; k < 64; ++k) { p[k] = __expf(p[k] - m); l += p[k]; }
          const float inv = 1.0f / l;
#pragma unroll
          for (int k = 0; k < 64; ++k) p[k] *= inv;
        } else {
#pragma unroll
          for (int k = 0; k < 64; ++k) p[k] = 0.0f;
        }
        put_row_bf16(sP, tid, p);
        float dp[64];
        tmem_ld64(tDP, dp);
        float delta = 0.0f;
#pragma unroll
        for (int k = 0; k < 64; ++k) delta = fmaf(p[k], dp[k], delta);
#pragma unroll
        for (int k = 0; k < 64; ++k) dp[k] = g * p[k] * (dp[k] - delta);       // dS (zero for inactive rows since p = 0)
        put_row_bf16(sDS, tid, dp);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dQ = dS K_j : A K-major over keys, B = K_j viewed [N = dh][K = keys] (MN-major)
          umma_bf16(tmem + cDQ, umma_smem_desc_sw128(smem_u32(sDS) + k * 32, 0, 1024),
                    umma_smem_desc_sw128(smem_u32(sK) + k * 2048, 8192, 1024), idesc_dq, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // [dK;dV] += [dS;P]^T [Q|dO] : both operands MN-major, K = 128 tokens
          umma_bf16(tmem + cDKV, umma_smem_desc_sw128(smem_u32(sDS) + k * 2048, 16384, 1024),
                    umma_smem_desc_sw128(smem_u32(sQ) + k * 2048, 16384, 1024), idesc_kv, (!first_tile || k > 0) ? 1u : 0u);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, ph_mma, 0x613); ph_mma ^= 1;
      tc_fence_after_sync();
      {   // dq rows of this tile (active rows only); stage through sP (free now) for coalesced stores
        float dqv[64];
        tmem_ld64(tDQ, dqv);
#pragma unroll
        for (int k = 0; k < 64; ++k) dqv[k] *= a.q_scale;
        put_row_bf16(sP, tid, dqv);
      }
      __syncthreads();
      {
        const int* ttb = a.tt + b * a.S;
        const int S = a.S, jj = j + 1;
        uint8_t* dst = reinterpret_cast<uint8_t*>(a.dq + (static_cast<size_t>(b) * a.S + tb) * HD + h * 64);
        flush_rows(sP, dst, static_cast<size_t>(HD) * 2, 128, [=](int r) { return (tb + r < S) && (ttb[tb + r] == jj); });
      }
      tc_fence_before_sync();
      __syncthreads();
      first_tile = false;
    }
    // ---- flush dK_j (rows 0..63, cols 0..63 of DKV) and dV_j (rows 64..127, cols 64..127)
    {
      float v[64];
      const bool isv = tid >= 64;
      if (!first_tile) tmem_ld64(tDKV + (isv ? 64 : 0), v);
      else {
#pragma unroll
        for (int k = 0; k < 64; ++k) v[k] = 0.0f;
      }
      if (isv) {
#pragma unroll
        for (int k = 0; k < 64; ++k) v[k] = fmaf(g, v[k], usum[k]);
      }
      put_row_bf16(sDS, tid, v);      // rows 0..63: dK keys, rows 64..127: dV keys
    }
    tc_fence_before_sync();
    __syncthreads();
    {
      const size_t ld = static_cast<size_t>(2 * HD) * 2;
      uint8_t* base = reinterpret_cast<uint8_t*>(a.dkv + (static_cast<size_t>(b) * a.n_media + j) * 64 * (2 * HD) + h * 64);
      const int cc = tid & 7;
      for (int r = tid >> 3; r < 128; r += 16) {
        const int key = r & 63;
        uint8_t* dst = base + static_cast<size_t>(key) * ld + (r >= 64 ? static_cast<size_t>(HD) * 2 : 0) + cc * 16;
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(sDS + sw128(r, cc));
      }
    }
    tc_fence_after_sync();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem, ncols);
}


// ============================================================================================ perceiver resampler cores
// perceiver_resampler.py:79-95: 64 latent queries of one image attend to its nk = T*F + 64 keys, no mask.
// One CTA per (head, image).  The 64 query rows occupy rows 0..63 of a 128-row MMA tile (rows 64..127 are ignored:
// their P / dS rows are written as zero, so they contribute nothing to the key-side products).  Keys stream through in
// tiles of 64; the forward makes two passes (row max / sum, then P V) so the TMEM accumulator is never rescaled.
struct RTcArgs {
  __nv_bfloat16* o;          // [BN*64, H*64]
  float* lse;                // [BN, H, 64]
  int BN, H, nk;
};

__global__ void __launch_bounds__(128) resampler_core_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                    const __grid_constant__ CUtensorMap tmKV, const RTcArgs a) {
  pdl_launch_dependents();
  pdl_wait();      // text_time / lse are read right away: no prologue to overlap in these short kernels
  FM_DYN_SMEM(uint8_t, rs_raw);
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(rs_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = sm;
  uint8_t* sK = sQ + 16384;
  uint8_t* sV = sK + 8192;
  uint8_t* sP = sV + 8192;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sP + 16384);
  uint64_t* bar_s = bar_load + 1;          // separate barriers for S and PV: see xattn_core_fwd_tc_kernel
  uint64_t* bar_o = bar_s + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, bn = blockIdx.y;
  const int HD = a.H * 64;
  const bool rowv = tid < 64;
  if (tid == 0) {
    mbar_init(bar_load, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1); fence_mbar_init();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t tO = tS + 64;
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false, false);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
  const int ntiles = (a.nk + 63) / 64;
  uint32_t ph_load = 0, ph_s = 0, ph_o = 0;
  float m = -INFINITY, l = 0.0f;

  for (int pass = 0; pass < 2; ++pass) {
    for (int kt = 0; kt < ntiles; ++kt) {
      const int nvalid = min(64, a.nk - kt * 64);
      const bool first = (pass == 0 && kt == 0);
      if (tid == 0) {
        mbar_arrive_expect_tx(bar_load, (first ? 16384 : 0) + 8192 + (pass ? 8192 : 0));
        if (first) tma_load_2d(sQ, &tmQ, bar_load, h * 64, bn * 64);
        const int krow = bn * a.nk + kt * 64;
        tma_load_2d(sK, &tmKV, bar_load, h * 64, krow);
        if (pass) tma_load_2d(sV, &tmKV, bar_load, HD + h * 64, krow);
      }
      mbar_wait(bar_load, ph_load, 0x620); ph_load ^= 1;
      if (tid == 0) {
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, umma_smem_desc_sw128(smem_u32(sQ) + k * 32, 0, 1024), umma_smem_desc_sw128(smem_u32(sK) + k * 32, 0, 1024),
                    idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
      }
      mbar_wait(bar_s, ph_s, 0x621); ph_s ^= 1;
      tc_fence_after_sync();
      float s[64];
      tmem_ld64(tS, s);
      if (pass == 0) {
        float mt = m;
#pragma unroll
        for (int k = 0; k < 64; ++k) if (k < nvalid) mt = fmaxf(mt, s[k]);
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 64; ++k) if (k < nvalid) acc += __expf(s[k] - mt);
        l = l * __expf(m - mt) + acc;
        m = mt;
        tc_fence_before_sync();
        __syncthreads();                       // everyone has read S before the next tile's MMA overwrites it
      } else {
        const float inv = 1.0f / l;
#pragma unroll
        for (int k = 0; k < 64; ++k) s[k] = (rowv && k < nvalid) ? __expf(s[k] - m) * inv : 0.0f;
        put_row_bf16(sP, tid, s);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after_sync();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + 64, umma_smem_desc_sw128(smem_u32(sP) + k * 32, 0, 1024),
                      umma_smem_desc_sw128(smem_u32(sV) + k * 2048, 8192, 1024), idesc_o, (kt > 0 || k > 0) ? 1u : 0u);
          umma_commit(bar_o);
        }
        mbar_wait(bar_o, ph_o, 0x622); ph_o ^= 1;
        tc_fence_after_sync();
      }
    }
  }
  {
    float o[64];
    tmem_ld64(tO, o);
    put_row_bf16(sQ, tid, o);
    if (rowv && a.lse) a.lse[(static_cast<size_t>(bn) * a.H + h) * 64 + tid] = m + __logf(l);
  }
  tc_fence_before_sync();
  __syncthreads();
  {
    uint8_t* dst = reinterpret_cast<uint8_t*>(a.o + (static_cast<size_t>(bn) * 64) * HD + h * 64);
    flush_rows(sQ, dst, static_cast<size_t>(HD) * 2, 64, [](int) { return true; });
  }
  tc_fence_after_sync();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

struct RTcBwdArgs {
  const __nv_bfloat16* o;    // saved forward output [BN*64, H*64]
  const __nv_bfloat16* d_o;  // [BN*64, H*64]
  const float* lse;          // [BN, H, 64]
  __nv_bfloat16* dq;         // [BN*64, H*64] = q_scale * dS K
  __nv_bfloat16* dkv;        // [BN*nk, 2*H*64]
  float q_scale;
  int BN, H, nk;
  int tmem_compact;          // 1: 256 TMEM columns ([dK;dV] reuses the S | dP columns), two CTAs per SM side by side
};

__global__ void __launch_bounds__(128) resampler_core_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                    const __grid_constant__ CUtensorMap tmDO,
                                                                    const __grid_constant__ CUtensorMap tmKV, const RTcBwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();      // text_time / lse are read right away: no prologue to overlap in these short kernels
  FM_DYN_SMEM(uint8_t, rb_raw);
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(rb_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = sm;
  uint8_t* sDO = sQ + 16384;
  uint8_t* sDS = sDO + 16384;
  uint8_t* sP = sDS + 16384;
  uint8_t* sK = sP + 16384;
  uint8_t* sV = sK + 8192;
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(sV + 8192);
  uint64_t* bar_q = bar_kv + 1;
  uint64_t* bar_mma = bar_q + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, bn = blockIdx.y;
  const int HD = a.H * 64;
  const bool rowv = tid < 64;
  if (tid == 0) {
    mbar_init(bar_kv, 1); mbar_init(bar_q, 1); mbar_init(bar_mma, 1); fence_mbar_init();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmKV);
    mbar_arrive_expect_tx(bar_q, 16384 + 16384);
    tma_load_2d(sQ, &tmQ, bar_q, h * 64, bn * 64);
    tma_load_2d(sDO, &tmDO, bar_q, h * 64, bn * 64);
  }
  // TMEM map.  Wide: S 0 | dP 64 | dQ 128 | [dK;dV] 256 (512 columns = one CTA per SM at a time).  Compact: dQ accumulates
  // across key tiles and keeps columns 128..191; [dK;dV] is produced and read out within one key tile, after S and dP have been
  // consumed (the __syncthreads before its MMA) and before the next tile's S / dP MMAs (the __syncthreads after its
  // read-out), so it reuses columns 0..127 -> 256 columns.
  const uint32_t ncols = a.tmem_compact ? 256u : 512u;
  const uint32_t cDKV = a.tmem_compact ? 0u : 256u;
  if (warp == 0) tmem_alloc(tmem_slot, ncols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  const uint32_t tS = tmem + lane_base, tDP = tS + 64, tDQ = tS + 128, tDKV = tS + cDKV;

  // delta = rowsum(dO * O), lse
  float delta = 0.0f, lse = 0.0f;
  if (rowv) {
    const size_t roff = (static_cast<size_t>(bn) * 64 + tid) * HD + h * 64;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = *reinterpret_cast<const uint4*>(a.d_o + roff + c * 8);
      const uint4 w = *reinterpret_cast<const uint4*>(a.o + roff + c * 8);
      const uint32_t u4[4] = {u.x, u.y, u.z, u.w}, w4[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack_bf16x2(u4[e]), y = unpack_bf16x2(w4[e]);
        delta = fmaf(x.x, y.x, delta);
        delta = fmaf(x.y, y.y, delta);
      }
    }
    lse = a.lse[(static_cast<size_t>(bn) * a.H + h) * 64 + tid];
  }
  mbar_wait(bar_q, 0, 0x630);

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false, false);
  constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64, false, true);
  constexpr uint32_t idesc_kv = umma_idesc_bf16(128, 128, true, true);
  const int ntiles = (a.nk + 63) / 64;
  uint32_t ph_kv = 0, ph_mma = 0;
  for (int kt = 0; kt < ntiles; ++kt) {
    const int nvalid = min(64, a.nk - kt * 64);
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_kv, 8192 + 8192);
      const int krow = bn * a.nk + kt * 64;
      tma_load_2d(sK, &tmKV, bar_kv, h * 64, krow);
      tma_load_2d(sV, &tmKV, bar_kv, HD + h * 64, krow);
    }
    mbar_wait(bar_kv, ph_kv, 0x631); ph_kv ^= 1;
    if (tid == 0) {
      tc_fence_after_sync();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, umma_smem_desc_sw128(smem_u32(sQ) + k * 32, 0, 1024), umma_smem_desc_sw128(smem_u32(sK) + k * 32, 0, 1024),
                  idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem + 64, umma_smem_desc_sw128(smem_u32(sDO) + k * 32, 0, 1024), umma_smem_desc_sw128(smem_u32(sV) + k * 32, 0, 1024),
                  idesc_s, k > 0 ? 1u : 0u);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma, 0x632); ph_mma ^= 1;
    tc_fence_after_sync();
    {
      float p[64];
      tmem_ld64(tS, p);
#pragma unroll
      for (int k = 0; k < 64; ++k) p[k] = (rowv && k < nvalid) ? __expf(p[k] - lse) : 0.0f;
      put_row_bf16(sP, tid, p);
      float dp[64];
      tmem_ld64(tDP, dp);
#pragma unroll
      for (int k = 0; k < 64; ++k) dp[k] = p[k] * (dp[k] - delta);
      put_row_bf16(sDS, tid, dp);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem + 128, umma_smem_desc_sw128(smem_u32(sDS) + k * 32, 0, 1024),
                  umma_smem_desc_sw128(smem_u32(sK) + k * 2048, 8192, 1024), idesc_dq, (kt > 0 || k > 0) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16(tmem + cDKV, umma_smem_desc_sw128(smem_u32(sDS) + k * 2048, 16384, 1024),
                  umma_smem_desc_sw128(smem_u32(sQ) + k * 2048, 16384, 1024), idesc_kv, k > 0 ? 1u : 0u);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma, 0x633); ph_mma ^= 1;
    tc_fence_after_sync();
    {
      float v[64];
      tmem_ld64(tDKV + (tid >= 64 ? 64 : 0), v);
      put_row_bf16(sDS, tid, v);               // rows 0..63: dK of this key tile, rows 64..127: dV
    }
    tc_fence_before_sync();
    __syncthreads();
    {
      const size_t ld = static_cast<size_t>(2 * HD) * 2;
      uint8_t* base = reinterpret_cast<uint8_t*>(a.dkv + (static_cast<size_t>(bn) * a.nk + kt * 64) * (2 * HD) + h * 64);
      const int cc = tid & 7;
      for (int r = tid >> 3; r < 128; r += 16) {
        const int key = r & 63;
        if (key < nvalid)
          *reinterpret_cast<uint4*>(base + static_cast<size_t>(key) * ld + (r >= 64 ? static_cast<size_t>(HD) * 2 : 0) + cc * 16) =
              *reinterpret_cast<const uint4*>(sDS + sw128(r, cc));
      }
    }
    tc_fence_after_sync();
    __syncthreads();
  }
  {
    float dqv[64];
    tmem_ld64(tDQ, dqv);
#pragma unroll
    for (int k = 0; k < 64; ++k) dqv[k] *= a.q_scale;
    put_row_bf16(sP, tid, dqv);
  }
  tc_fence_before_sync();
  __syncthreads();
  {
    uint8_t* dst = reinterpret_cast<uint8_t*>(a.dq + (static_cast<size_t>(bn) * 64) * HD + h * 64);
    flush_rows(sP, dst, static_cast<size_t>(HD) * 2, 64, [](int) { return true; });
  }
  tc_fence_after_sync();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

}  // namespace fm
