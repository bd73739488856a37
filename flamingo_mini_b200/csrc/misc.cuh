// Small HBM-bound helper kernels of the hot path.
#pragma once
#include "ptx.cuh"

namespace fm {

// text_time[b, i] = sum_{j<=i} media_locations[b, j]   (gated_cross_attention.py:97). One warp per row.
__global__ void text_time_kernel(const int* __restrict__ ml, int* __restrict__ tt, int B, int S) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  int carry = 0;
  for (int i0 = 0; i0 < S; i0 += 32) {
    const int i = i0 + lane;
    int v = (i < S) ? ml[b * S + i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    if (i < S) tt[b * S + i] = v + carry;
    carry += __shfl_sync(0xffffffffu, v, 31);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i8 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i8 + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(src + i8);
    const float4 b = *reinterpret_cast<const float4*>(src + i8 + 4);
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(dst + i8) = u;
  } else {
    for (long long i = i8; i < n; ++i) dst[i] = __float2bfloat16(src[i]);
  }
}

// *out += sum_i a[i]*b[i]  (bf16 inputs).  Used for d(alpha_attn) = (1-tanh^2) * sum(dO_ungated * O).
__global__ void dot_reduce_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, long long n,
                                  float* out) {
  pdl_launch_dependents();
  pdl_wait();
  float acc = 0.0f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i + 8 <= n; i += stride) {
    const uint4 x = *reinterpret_cast<const uint4*>(a + i);
    const uint4 y = *reinterpret_cast<const uint4*>(b + i);
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 p = unpack_bf16x2(xs[e]), q = unpack_bf16x2(ys[e]);
      acc = fmaf(p.x, q.x, acc);
      acc = fmaf(p.y, q.y, acc);
    }
  }
  acc = warp_sum(acc);
  __shared__ float sh[32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = acc;
  __syncthreads();
  if (w == 0) {
    float t = (l < (blockDim.x >> 5)) ? sh[l] : 0.0f;
    t = warp_sum(t);
    if (l == 0) atomicAdd(out, t);
  }
}

// d(alpha) = (1 - tanh(alpha)^2) * raw   for the two gates of a block
__global__ void alpha_grad_kernel(const float* alpha_attn, const float* alpha_ffw, const float* raw_ffw_attn,
                                  float* d_alpha_attn, float* d_alpha_ffw) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) {
    const float ta = tanhf(*alpha_attn), tf = tanhf(*alpha_ffw);
    *d_alpha_ffw = (1.0f - tf * tf) * raw_ffw_attn[0];
    *d_alpha_attn = (1.0f - ta * ta) * raw_ffw_attn[1];
  }
}

// dst[r, :] = src[r % period, :]   (fp32) — latents repeated over the batch (perceiver_resampler.py:179)
__global__ void bcast_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows, int D, int period) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // float4 index
  const int d4 = D >> 2;
  if (idx >= rows * d4) return;
  const long long r = idx / d4;
  const int c = static_cast<int>(idx - r * d4);
  reinterpret_cast<float4*>(dst)[idx] = reinterpret_cast<const float4*>(src)[(r % period) * d4 + c];
}

// out[(r % period) / group, d] += src[r, d]  (out fp32, pre-zeroed; src bf16 or fp32).
// Gradient of broadcast parameters: latents (period 64, group 1) and time_pos_emb (period T*F, group F).
__global__ void group_rowsum_kernel(const void* __restrict__ src, int src_f32, long long rows, int D, int period, int group,
                                    float* out, int rows_per_block) {
  pdl_launch_dependents();
  pdl_wait();
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float acc = 0.0f;
  int cur = -1;
  for (long long r = r0; r < r1; ++r) {
    const int gidx = static_cast<int>(r % period) / group;
    if (gidx != cur) {
      if (cur >= 0) atomicAdd(out + static_cast<size_t>(cur) * D + d, acc);
      cur = gidx; acc = 0.0f;
    }
    acc += src_f32 ? reinterpret_cast<const float*>(src)[r * D + d]
                   : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[r * D + d]);
  }
  if (cur >= 0) atomicAdd(out + static_cast<size_t>(cur) * D + d, acc);
}

// Fused AdamW step over one flat parameter arena (reference recipe training/train.sh:10-13 = HF Trainer's
// adamw_torch): decoupled weight decay, moment updates, bias-corrected update, and the bf16 tensor-core shadow of the new
// parameters written in the same pass (so no separate cast kernel follows an optimizer step).  HBM-bound: 16-20 B read and
// 14 B written per parameter.  Same arithmetic order as torch.optim.AdamW (single-tensor path).
struct AdamWArgs {
  float* p; const float* g; float* m; float* v;
  __nv_bfloat16* shadow;          // optional
  const float* decay_mask;        // optional [n]: 1 where weight decay applies, 0 elsewhere
  const float* grad_scale;        // optional device scalar multiplied into g (gradient clipping)
  long long n;
  float lr, b1, b2, eps, wd, bc1, bc2_sqrt;
};
__device__ __forceinline__ float adamw_one(const AdamWArgs& a, float p, float g, float& m, float& v, float mask) {
  p = p * (1.0f - a.lr * a.wd * mask);
  m = m + (g - m) * (1.0f - a.b1);                       // lerp, as torch
  v = v * a.b2 + g * g * (1.0f - a.b2);
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  return p - (a.lr / a.bc1) * (m / denom);
}
__global__ void __launch_bounds__(256) adamw_kernel(AdamWArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const float gs = a.grad_scale ? __ldg(a.grad_scale) : 1.0f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  const long long n4 = a.n & ~3LL;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n4; i += stride) {
    float4 p = *reinterpret_cast<const float4*>(a.p + i);
    float4 g = *reinterpret_cast<const float4*>(a.g + i);
    float4 m = *reinterpret_cast<const float4*>(a.m + i);
    float4 v = *reinterpret_cast<const float4*>(a.v + i);
    float4 k = a.decay_mask ? *reinterpret_cast<const float4*>(a.decay_mask + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    p.x = adamw_one(a, p.x, g.x * gs, m.x, v.x, k.x);
    p.y = adamw_one(a, p.y, g.y * gs, m.y, v.y, k.y);
    p.z = adamw_one(a, p.z, g.z * gs, m.z, v.z, k.z);
    p.w = adamw_one(a, p.w, g.w * gs, m.w, v.w, k.w);
    *reinterpret_cast<float4*>(a.p + i) = p;
    *reinterpret_cast<float4*>(a.m + i) = m;
    *reinterpret_cast<float4*>(a.v + i) = v;
    if (a.shadow) {
      uint2 u;
      u.x = pack_bf16x2(p.x, p.y); u.y = pack_bf16x2(p.z, p.w);
      *reinterpret_cast<uint2*>(a.shadow + i) = u;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n - n4)) {      // tail (arenas are padded to 8, so normally empty)
    const long long i = n4 + threadIdx.x;
    float m = a.m[i], v = a.v[i];
    const float p = adamw_one(a, a.p[i], a.g[i] * gs, m, v, a.decay_mask ? a.decay_mask[i] : 1.0f);
    a.p[i] = p; a.m[i] = m; a.v[i] = v;
    if (a.shadow) a.shadow[i] = __float2bfloat16(p);
  }
}

}  // namespace fm
