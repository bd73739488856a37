// Attention cores (score -> softmax -> weighted sum) for the two modules, forward and backward.
//
// Round-1 implementation: fp32 CUDA-core kernels, one thread per query row, K/V slabs staged in shared memory
// (every key row is a warp-wide broadcast read).  They carry < 1.5 % of the hot-path FLOPs; the
// tcgen05 version of these cores is the next kernel on the list (DESIGN.md §"what comes next").
//
// Gated cross-attention core  (gated_cross_attention.py:95-124 of the reference):
//   token i of sample b may attend only to the 64 latents of image number text_time[b,i] (1-based);
//   text_time == 0  -> output row is exactly zero;   text_time > n_media -> uniform average over ALL keys.
//   => exactly one 64-key slab is live per token; masked keys are never computed.
// Perceiver-resampler core   (perceiver_resampler.py:79-95): 64 latent queries over F+64 keys, no mask.
//
// Layouts (bf16, row-major):  q/o/dq/do : [rows, H*64] with head h in columns [64h, 64h+64)
//                             kv/dkv    : [key rows, 2*H*64], K in columns [0, H*64), V in [H*64, 2*H*64)
// q is already multiplied by dim_head^-0.5 (fused into the to_q GEMM epilogue).
#pragma once
#include "ptx.cuh"

namespace fm {

constexpr int AC_DH = 64;      // dim_head the kernels are specialised for
constexpr int AC_PAD = 65;     // padded row stride (floats) for row-per-thread shared-memory tiles

__device__ __forceinline__ void load_row64(const __nv_bfloat16* p, float (&v)[64], float mul = 1.0f) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 t = *reinterpret_cast<const uint4*>(p + c * 8);
    const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), cc = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[c * 8 + 0] = a.x * mul;  v[c * 8 + 1] = a.y * mul;  v[c * 8 + 2] = b.x * mul; v[c * 8 + 3] = b.y * mul;
    v[c * 8 + 4] = cc.x * mul; v[c * 8 + 5] = cc.y * mul; v[c * 8 + 6] = d.x * mul; v[c * 8 + 7] = d.y * mul;
  }
}
__device__ __forceinline__ void store_row64(__nv_bfloat16* p, const float (&v)[64], float mul = 1.0f) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 u;
    u.x = pack_bf16x2(v[c * 8 + 0] * mul, v[c * 8 + 1] * mul); u.y = pack_bf16x2(v[c * 8 + 2] * mul, v[c * 8 + 3] * mul);
    u.z = pack_bf16x2(v[c * 8 + 4] * mul, v[c * 8 + 5] * mul); u.w = pack_bf16x2(v[c * 8 + 6] * mul, v[c * 8 + 7] * mul);
    *reinterpret_cast<uint4*>(p + c * 8) = u;
  }
}
// Cooperative load of a [64 keys x 64] bf16 slab (row stride ld elements) into fp32 shared memory [64][64].
// Rows >= valid_rows are zero-filled.
__device__ __forceinline__ void load_slab(float* dst, const __nv_bfloat16* src, long long ld, int valid_rows) {
  for (int idx = threadIdx.x; idx < 64 * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = idx & 7;
    float v[8];
    if (r < valid_rows) {
      const uint4 t = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(r) * ld + c * 8);
      const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), cc = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = cc.x; v[5] = cc.y; v[6] = d.x; v[7] = d.y;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.0f;
    }
    float* o = dst + r * 64 + c * 8;
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// ============================================================================================ gated xattn core
struct XCoreArgs {
  const __nv_bfloat16* q;   // [B*S, H*64]
  const __nv_bfloat16* kv;  // [B*n_media*64, 2*H*64]
  const int* tt;            // [B, S] text_time
  __nv_bfloat16* o;         // [B*S, H*64]
  int B, S, H, n_media;
};

__global__ void __launch_bounds__(128) xattn_core_fwd_kernel(const XCoreArgs a) {
  __shared__ __align__(16) float Ks[64 * 64];
  __shared__ __align__(16) float Vs[64 * 64];
  __shared__ float vmean[64];
  __shared__ int s_min, s_max, s_any_uniform;
  const int b = blockIdx.z, h = blockIdx.y;
  const int t = blockIdx.x * 128 + threadIdx.x;
  const bool valid = t < a.S;
  const int mytt = valid ? a.tt[b * a.S + t] : 0;
  const int HD = a.H * AC_DH;
  const long long ldkv = 2LL * HD;
  if (threadIdx.x == 0) { s_min = 0x7fffffff; s_max = -1; s_any_uniform = 0; }
  __syncthreads();
  if (valid && mytt >= 1 && mytt <= a.n_media) { atomicMin(&s_min, mytt - 1); atomicMax(&s_max, mytt - 1); }
  if (valid && mytt > a.n_media) s_any_uniform = 1;
  __syncthreads();
  const int jmin = s_min, jmax = s_max, any_uniform = s_any_uniform;

  float o[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) o[d] = 0.0f;

  for (int j = jmin; j <= jmax; ++j) {
    __syncthreads();
    const __nv_bfloat16* kbase = a.kv + (static_cast<size_t>(b) * a.n_media + j) * 64 * ldkv + h * AC_DH;
    load_slab(Ks, kbase, ldkv, 64);
    load_slab(Vs, kbase + HD, ldkv, 64);
    __syncthreads();
    if (valid && mytt == j + 1) {
      float s[64];
      {
        float qv[64];
        load_row64(a.q + static_cast<size_t>(b * a.S + t) * HD + h * AC_DH, qv);
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          float acc = 0.0f;
#pragma unroll
          for (int d = 0; d < 64; ++d) acc = fmaf(qv[d], Ks[k * 64 + d], acc);
          s[k] = acc;
        }
      }
      float m = s[0];
#pragma unroll
      for (int k = 1; k < 64; ++k) m = fmaxf(m, s[k]);
      float l = 0.0f;
#pragma unroll
      for (int k = 0; k < 64; ++k) { s[k] = __expf(s[k] - m); l += s[k]; }
      const float inv = 1.0f / l;
#pragma unroll
      for (int k = 0; k < 64; ++k) {
        const float p = s[k] * inv;
#pragma unroll
        for (int d = 0; d < 64; ++d) o[d] = fmaf(p, Vs[k * 64 + d], o[d]);
      }
    }
  }
  if (any_uniform) {   // rows with more <image> markers than images: uniform average of all values
    __syncthreads();
    const int nkeys = a.n_media * 64;
    if (threadIdx.x < 64) {
      float acc = 0.0f;
      const __nv_bfloat16* vb = a.kv + static_cast<size_t>(b) * nkeys * ldkv + HD + h * AC_DH + threadIdx.x;
      for (int k = 0; k < nkeys; ++k) acc += __bfloat162float(vb[static_cast<size_t>(k) * ldkv]);
      vmean[threadIdx.x] = acc / nkeys;
    }
    __syncthreads();
    if (valid && mytt > a.n_media) {
#pragma unroll
      for (int d = 0; d < 64; ++d) o[d] = vmean[d];
    }
  }
  if (valid) store_row64(a.o + static_cast<size_t>(b * a.S + t) * HD + h * AC_DH, o);
}

struct XCoreBwdArgs {
  const __nv_bfloat16* q;    // [B*S, H*64] (scaled)
  const __nv_bfloat16* kv;   // [B*n_media*64, 2*H*64]
  const int* tt;
  const __nv_bfloat16* d_o;  // [B*S, H*64] gradient w.r.t. o BEFORE the gate
  const float* gate;         // optional: d_o is multiplied by tanh(*gate)
  __nv_bfloat16* dq;         // [B*S, H*64]  = q_scale * dS K  (gradient w.r.t. the unscaled to_q output)
  __nv_bfloat16* dkv;        // [B*n_media*64, 2*H*64]
  float q_scale;
  int B, S, H, n_media;
};

constexpr int XBWD_SMEM_FLOATS = 2 * 64 * 64 + 4 * 128 * AC_PAD + 64;
constexpr int XBWD_SMEM_BYTES = XBWD_SMEM_FLOATS * 4;

// One CTA per (head, sample): owns dK/dV of every slab of that sample, so they are written once, deterministically.
__global__ void __launch_bounds__(128) xattn_core_bwd_kernel(const XCoreBwdArgs a) {
  extern __shared__ __align__(16) float xsm[];
  float* Ks = xsm;
  float* Vs = Ks + 64 * 64;
  float* Ps = Vs + 64 * 64;           // [128][65]
  float* dSs = Ps + 128 * AC_PAD;     // [128][65]
  float* Qs = dSs + 128 * AC_PAD;     // [128][65]
  float* dOs = Qs + 128 * AC_PAD;     // [128][65]
  float* usum = dOs + 128 * AC_PAD;   // [64] sum of gated d_o over "uniform" rows / n_keys
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x;
  const int HD = a.H * AC_DH;
  const long long ldkv = 2LL * HD;
  const float g = a.gate ? tanhf(__ldg(a.gate)) : 1.0f;
  const int nkeys = a.n_media * 64;

  // ---- prelude: rows that get no gradient through q (tt == 0 or tt > n_media) and the uniform-row dV term
  if (tid < 64) usum[tid] = 0.0f;
  __syncthreads();
  for (int t0 = 0; t0 < a.S; t0 += 128) {
    const int t = t0 + tid;
    if (t < a.S) {
      const int mytt = a.tt[b * a.S + t];
      if (mytt < 1 || mytt > a.n_media) {
        float z[64];
#pragma unroll
        for (int d = 0; d < 64; ++d) z[d] = 0.0f;
        store_row64(a.dq + static_cast<size_t>(b * a.S + t) * HD + h * AC_DH, z);
        if (mytt > a.n_media) {
          float dv[64];
          load_row64(a.d_o + static_cast<size_t>(b * a.S + t) * HD + h * AC_DH, dv, g / nkeys);
#pragma unroll
          for (int d = 0; d < 64; ++d) atomicAdd(&usum[d], dv[d]);
        }
      }
    }
  }
  __syncthreads();

  const int pk = tid & 63, ph = tid >> 6;   // phase-2 ownership: key pk, columns [32*ph, 32*ph+32)
  for (int j = 0; j < a.n_media; ++j) {
    __syncthreads();
    const __nv_bfloat16* kbase = a.kv + (static_cast<size_t>(b) * a.n_media + j) * 64 * ldkv + h * AC_DH;
    load_slab(Ks, kbase, ldkv, 64);
    load_slab(Vs, kbase + HD, ldkv, 64);
    float accK[32], accV[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) { accK[e] = 0.0f; accV[e] = usum[ph * 32 + e]; }
    __syncthreads();

    for (int t0 = 0; t0 < a.S; t0 += 128) {
      const int t = t0 + tid;
      const bool active = (t < a.S) && (a.tt[b * a.S + t] == j + 1);
      if (!__syncthreads_or(active)) continue;
      // ---------------- phase 1: one thread per token
      if (active) {
        const size_t roff = static_cast<size_t>(b * a.S + t) * HD + h * AC_DH;
        float* Prow = Ps + tid * AC_PAD;
        float* dSrow = dSs + tid * AC_PAD;
        float m = -INFINITY;
        {   // scores -> Prow
          float qv[64];
          load_row64(a.q + roff, qv);
#pragma unroll
          for (int d = 0; d < 64; ++d) Qs[tid * AC_PAD + d] = qv[d];
#pragma unroll 4
          for (int k = 0; k < 64; ++k) {
            float acc = 0.0f;
#pragma unroll
            for (int d = 0; d < 64; ++d) acc = fmaf(qv[d], Ks[k * 64 + d], acc);
            Prow[k] = acc;
            m = fmaxf(m, acc);
          }
        }
        float l = 0.0f;
#pragma unroll 8
        for (int k = 0; k < 64; ++k) { const float e = __expf(Prow[k] - m); Prow[k] = e; l += e; }
        const float inv = 1.0f / l;
        float delta = 0.0f;
        {   // P normalised; dP = dO V^T -> dSrow; delta = sum P*dP
          float dv[64];
          load_row64(a.d_o + roff, dv, g);
#pragma unroll
          for (int d = 0; d < 64; ++d) dOs[tid * AC_PAD + d] = dv[d];
#pragma unroll 4
          for (int k = 0; k < 64; ++k) {
            float acc = 0.0f;
#pragma unroll
            for (int d = 0; d < 64; ++d) acc = fmaf(dv[d], Vs[k * 64 + d], acc);
            const float pk_ = Prow[k] * inv;
            Prow[k] = pk_;
            dSrow[k] = acc;
            delta = fmaf(pk_, acc, delta);
          }
        }
        {   // dS = P*(dP - delta); dq = dS K
          float dqv[64];
#pragma unroll
          for (int d = 0; d < 64; ++d) dqv[d] = 0.0f;
#pragma unroll 4
          for (int k = 0; k < 64; ++k) {
            const float ds = Prow[k] * (dSrow[k] - delta);
            dSrow[k] = ds;
#pragma unroll
            for (int d = 0; d < 64; ++d) dqv[d] = fmaf(ds, Ks[k * 64 + d], dqv[d]);
          }
          store_row64(a.dq + roff, dqv, a.q_scale);
        }
      } else {
        for (int k = 0; k < 64; ++k) { Ps[tid * AC_PAD + k] = 0.0f; dSs[tid * AC_PAD + k] = 0.0f; }
        for (int d = 0; d < 64; ++d) { Qs[tid * AC_PAD + d] = 0.0f; dOs[tid * AC_PAD + d] = 0.0f; }
      }
      __syncthreads();
      // ---------------- phase 2: thread (key pk, half ph) accumulates dK / dV over the 128 tokens of the tile
      for (int i = 0; i < 128; ++i) {
        const float ds = dSs[i * AC_PAD + pk];
        const float pp = Ps[i * AC_PAD + pk];
        const float* qi = Qs + i * AC_PAD + ph * 32;
        const float* di = dOs + i * AC_PAD + ph * 32;
#pragma unroll
        for (int e = 0; e < 32; ++e) { accK[e] = fmaf(ds, qi[e], accK[e]); accV[e] = fmaf(pp, di[e], accV[e]); }
      }
      __syncthreads();
    }
    // ---------------- flush dK_j, dV_j
    __nv_bfloat16* dk = a.dkv + ((static_cast<size_t>(b) * a.n_media + j) * 64 + pk) * ldkv + h * AC_DH + ph * 32;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 u, w;
      u.x = pack_bf16x2(accK[c * 8 + 0], accK[c * 8 + 1]); u.y = pack_bf16x2(accK[c * 8 + 2], accK[c * 8 + 3]);
      u.z = pack_bf16x2(accK[c * 8 + 4], accK[c * 8 + 5]); u.w = pack_bf16x2(accK[c * 8 + 6], accK[c * 8 + 7]);
      w.x = pack_bf16x2(accV[c * 8 + 0], accV[c * 8 + 1]); w.y = pack_bf16x2(accV[c * 8 + 2], accV[c * 8 + 3]);
      w.z = pack_bf16x2(accV[c * 8 + 4], accV[c * 8 + 5]); w.w = pack_bf16x2(accV[c * 8 + 6], accV[c * 8 + 7]);
      *reinterpret_cast<uint4*>(dk + c * 8) = u;
      *reinterpret_cast<uint4*>(dk + HD + c * 8) = w;
    }
  }
}

// ============================================================================================ resampler core
struct RCoreArgs {
  const __nv_bfloat16* q;    // [BN*64, H*64] scaled
  const __nv_bfloat16* kv;   // [BN*nk, 2*H*64]
  __nv_bfloat16* o;          // [BN*64, H*64]
  float* lse;                // [BN, H, 64]  log-sum-exp of each score row (saved for backward)
  int BN, H, nk;
};

__global__ void __launch_bounds__(64) resampler_core_fwd_kernel(const RCoreArgs a) {
  __shared__ __align__(16) float Ks[64 * 64];
  __shared__ __align__(16) float Vs[64 * 64];
  const int h = blockIdx.x, bn = blockIdx.y, tid = threadIdx.x;
  const int HD = a.H * AC_DH;
  const long long ldkv = 2LL * HD;
  const size_t roff = (static_cast<size_t>(bn) * 64 + tid) * HD + h * AC_DH;
  float qv[64], o[64];
  load_row64(a.q + roff, qv);
#pragma unroll
  for (int d = 0; d < 64; ++d) o[d] = 0.0f;
  float m = -INFINITY, l = 0.0f;
  for (int k0 = 0; k0 < a.nk; k0 += 64) {
    const int nvalid = min(64, a.nk - k0);
    __syncthreads();
    const __nv_bfloat16* kbase = a.kv + (static_cast<size_t>(bn) * a.nk + k0) * ldkv + h * AC_DH;
    load_slab(Ks, kbase, ldkv, nvalid);
    load_slab(Vs, kbase + HD, ldkv, nvalid);
    __syncthreads();
#pragma unroll 1
    for (int kh = 0; kh < 64; kh += 32) {          // two half-tiles of 32 keys keep the score array at 32 registers
      if (kh >= nvalid) break;
      float s[32];
      float mt = m;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float acc = 0.0f;
#pragma unroll
        for (int d = 0; d < 64; ++d) acc = fmaf(qv[d], Ks[(kh + k) * 64 + d], acc);
        s[k] = (kh + k < nvalid) ? acc : -INFINITY;
        mt = fmaxf(mt, s[k]);
      }
      const float corr = __expf(m - mt);             // m = -inf on the first tile -> 0
      l *= corr;
#pragma unroll
      for (int d = 0; d < 64; ++d) o[d] *= corr;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float p = __expf(s[k] - mt);           // masked keys: exp(-inf) = 0
        l += p;
#pragma unroll
        for (int d = 0; d < 64; ++d) o[d] = fmaf(p, Vs[(kh + k) * 64 + d], o[d]);
      }
      m = mt;
    }
  }
  const float inv = 1.0f / l;
  store_row64(a.o + roff, o, inv);
  if (a.lse) a.lse[(static_cast<size_t>(bn) * a.H + h) * 64 + tid] = m + __logf(l);
}

struct RCoreBwdArgs {
  const __nv_bfloat16* q;    // [BN*64, H*64] scaled
  const __nv_bfloat16* kv;   // [BN*nk, 2*H*64]
  const __nv_bfloat16* o;    // saved forward output
  const float* lse;          // [BN, H, 64]
  const __nv_bfloat16* d_o;  // [BN*64, H*64]
  __nv_bfloat16* dq;         // = q_scale * dS K
  __nv_bfloat16* dkv;        // [BN*nk, 2*H*64]
  float q_scale;
  int BN, H, nk;
};

constexpr int RBWD_SMEM_FLOATS = 2 * 64 * 64 + 4 * 64 * AC_PAD;
constexpr int RBWD_SMEM_BYTES = RBWD_SMEM_FLOATS * 4;

__global__ void __launch_bounds__(128) resampler_core_bwd_kernel(const RCoreBwdArgs a) {
  extern __shared__ __align__(16) float rsm[];
  float* Ks = rsm;
  float* Vs = Ks + 64 * 64;
  float* Ps = Vs + 64 * 64;          // [64 queries][65]
  float* dSs = Ps + 64 * AC_PAD;
  float* Qs = dSs + 64 * AC_PAD;
  float* dOs = Qs + 64 * AC_PAD;
  const int h = blockIdx.x, bn = blockIdx.y, tid = threadIdx.x;
  const int HD = a.H * AC_DH;
  const long long ldkv = 2LL * HD;
  const bool isq = tid < 64;                                      // phase-1 threads: one per query
  const size_t roff = (static_cast<size_t>(bn) * 64 + (tid & 63)) * HD + h * AC_DH;
  float delta = 0.0f, lse = 0.0f;
  float dqv[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) dqv[d] = 0.0f;
  if (isq) {
    float t0[64], t1[64];
    load_row64(a.d_o + roff, t0);
    load_row64(a.o + roff, t1);
#pragma unroll
    for (int d = 0; d < 64; ++d) { delta = fmaf(t0[d], t1[d], delta); dOs[tid * AC_PAD + d] = t0[d]; }
    load_row64(a.q + roff, t1);
#pragma unroll
    for (int d = 0; d < 64; ++d) Qs[tid * AC_PAD + d] = t1[d];
    lse = a.lse[(static_cast<size_t>(bn) * a.H + h) * 64 + tid];
  }
  const int pk = tid & 63, ph = tid >> 6;
  for (int k0 = 0; k0 < a.nk; k0 += 64) {
    const int nvalid = min(64, a.nk - k0);
    __syncthreads();
    const __nv_bfloat16* kbase = a.kv + (static_cast<size_t>(bn) * a.nk + k0) * ldkv + h * AC_DH;
    load_slab(Ks, kbase, ldkv, nvalid);
    load_slab(Vs, kbase + HD, ldkv, nvalid);
    __syncthreads();
    if (isq) {
      {   // pass A: P = exp(S - lse)
        float qv[64];
#pragma unroll
        for (int d = 0; d < 64; ++d) qv[d] = Qs[tid * AC_PAD + d];
#pragma unroll 4
        for (int k = 0; k < 64; ++k) {
          float acc = 0.0f;
#pragma unroll
          for (int d = 0; d < 64; ++d) acc = fmaf(qv[d], Ks[k * 64 + d], acc);
          Ps[tid * AC_PAD + k] = (k < nvalid) ? __expf(acc - lse) : 0.0f;
        }
      }
      {   // pass B: dS = P * (dO V^T - delta)
        float dv[64];
#pragma unroll
        for (int d = 0; d < 64; ++d) dv[d] = dOs[tid * AC_PAD + d];
#pragma unroll 4
        for (int k = 0; k < 64; ++k) {
          float acc = 0.0f;
#pragma unroll
          for (int d = 0; d < 64; ++d) acc = fmaf(dv[d], Vs[k * 64 + d], acc);
          dSs[tid * AC_PAD + k] = Ps[tid * AC_PAD + k] * (acc - delta);
        }
      }
      // pass C: dq += dS K
#pragma unroll 4
      for (int k = 0; k < 64; ++k) {
        const float ds = dSs[tid * AC_PAD + k];
#pragma unroll
        for (int d = 0; d < 64; ++d) dqv[d] = fmaf(ds, Ks[k * 64 + d], dqv[d]);
      }
    }
    __syncthreads();
    // phase 2: dK[key] = sum_i dS[i][key] q_i ; dV[key] = sum_i P[i][key] dO_i
    float accK[32], accV[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) { accK[e] = 0.0f; accV[e] = 0.0f; }
    for (int i = 0; i < 64; ++i) {
      const float ds = dSs[i * AC_PAD + pk];
      const float pp = Ps[i * AC_PAD + pk];
      const float* qi = Qs + i * AC_PAD + ph * 32;
      const float* di = dOs + i * AC_PAD + ph * 32;
#pragma unroll
      for (int e = 0; e < 32; ++e) { accK[e] = fmaf(ds, qi[e], accK[e]); accV[e] = fmaf(pp, di[e], accV[e]); }
    }
    if (pk < nvalid) {
      __nv_bfloat16* dk = a.dkv + (static_cast<size_t>(bn) * a.nk + k0 + pk) * ldkv + h * AC_DH + ph * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 u, w;
        u.x = pack_bf16x2(accK[c * 8 + 0], accK[c * 8 + 1]); u.y = pack_bf16x2(accK[c * 8 + 2], accK[c * 8 + 3]);
        u.z = pack_bf16x2(accK[c * 8 + 4], accK[c * 8 + 5]); u.w = pack_bf16x2(accK[c * 8 + 6], accK[c * 8 + 7]);
        w.x = pack_bf16x2(accV[c * 8 + 0], accV[c * 8 + 1]); w.y = pack_bf16x2(accV[c * 8 + 2], accV[c * 8 + 3]);
        w.z = pack_bf16x2(accV[c * 8 + 4], accV[c * 8 + 5]); w.w = pack_bf16x2(accV[c * 8 + 6], accV[c * 8 + 7]);
        *reinterpret_cast<uint4*>(dk + c * 8) = u;
        *reinterpret_cast<uint4*>(dk + HD + c * 8) = w;
      }
    }
  }
  if (isq) store_row64(a.dq + roff, dqv, a.q_scale);
}

}  // namespace fm
