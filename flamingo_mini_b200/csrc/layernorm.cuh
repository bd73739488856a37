// Row LayerNorm forward/backward (eps 1e-5, affine, biased variance) — the nn.LayerNorm calls at
// perceiver_resampler.py:52-53,187, gated_cross_attention.py:74 and utils.py:46 of the reference.
// HBM-bound: one CTA per row (loop), 128-bit loads, row cached in registers, fp32 statistics, bf16 output that
// feeds the tcgen05 GEMM's A operand directly.  The forward can (a) add a per-frame embedding to the input
// (time_pos_emb, perceiver_resampler.py:166) and (b) scatter rows into a "concatenated" destination
// ([media ; latents], perceiver_resampler.py:65) so that no torch.cat copy is ever made.
#pragma once
#include "ptx.cuh"

namespace fm {

struct LnArgs {
  const void* x;        // [rows, D] bf16 or fp32 (x_f32)
  int x_f32;
  const float* add;     // optional [n_add, D] fp32 embedding rows; index = (row % add_period) / add_group
  int add_period, add_group;
  const float* gamma;   // [D]
  const float* beta;    // [D]
  void* out;            // bf16 [*, D] (or fp32 when out_f32) at row map(row) = (row / in_group) * out_group + out_off + row % in_group
  int out_f32;
  int in_group, out_group, out_off;
  __nv_bfloat16* out2;  // optional compact bf16 copy at row `row`
  float* mean;          // optional [rows]
  float* rstd;          // optional [rows]
  int rows, D;
};

constexpr int LN_THREADS = 256;
constexpr int LN_MAXC = 4;  // 8-element chunks per thread -> D <= 256*4*8 = 8192

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < (blockDim.x >> 5)) ? sh[l] : 0.0f;
  t = warp_sum(t);
  return t;
}

__device__ __forceinline__ void load8(const void* base, int x_f32, size_t off, float (&v)[8]) {
  if (x_f32) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 t = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
    const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
}
__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(void* base, int f32, size_t off, const float (&v)[8]) {
  if (f32) {
    float* p = reinterpret_cast<float*>(base) + off;
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = u;
  }
}

__global__ void __launch_bounds__(LN_THREADS) ln_fwd_kernel(const LnArgs a) {
  __shared__ float sh[LN_THREADS / 32];
  const int nchunk = a.D >> 3;
  for (int row = blockIdx.x; row < a.rows; row += gridDim.x) {
    float v[LN_MAXC][8];
    float s = 0.0f;
    const float* addp = a.add ? a.add + static_cast<size_t>((row % a.add_period) / a.add_group) * a.D : nullptr;
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
      const int c = threadIdx.x + i * LN_THREADS;
      if (c < nchunk) {
        load8(a.x, a.x_f32, static_cast<size_t>(row) * a.D + c * 8, v[i]);
        if (addp) {
          float e[8]; load8f(addp + c * 8, e);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] += e[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    }
    const float mean = block_sum(s, sh) / a.D;
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
      const int c = threadIdx.x + i * LN_THREADS;
      if (c < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; ss += d * d; }
      }
    }
    const float rstd = rsqrtf(block_sum(ss, sh) / a.D + 1e-5f);
    if (threadIdx.x == 0) {
      if (a.mean) a.mean[row] = mean;
      if (a.rstd) a.rstd[row] = rstd;
    }
    const size_t orow = static_cast<size_t>(row / a.in_group) * a.out_group + a.out_off + row % a.in_group;
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
      const int c = threadIdx.x + i * LN_THREADS;
      if (c < nchunk) {
        float gm[8], bt[8], o[8];
        load8f(a.gamma + c * 8, gm);
        load8f(a.beta + c * 8, bt);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * gm[j] + bt[j];
        store8(a.out, a.out_f32, orow * a.D + c * 8, o);
        if (a.out2) store8(a.out2, 0, static_cast<size_t>(row) * a.D + c * 8, o);
      }
    }
  }
}

// Backward.  dxn = dy (gradient w.r.t. the LN output, bf16, at the same mapped rows as the forward output).
//   dx = rstd * (dxn*gamma - mean_D(dxn*gamma) - xhat * mean_D(dxn*gamma*xhat))  [+ dres]
//   dgamma = sum_rows dxn * xhat,  dbeta = sum_rows dxn      (per-CTA partials, reduced by ln_bwd_reduce_kernel)
struct LnBwdArgs {
  const __nv_bfloat16* dy;  // mapped rows (see LnArgs)
  const __nv_bfloat16* dy2; // optional second gradient source added to dy, compact rows [rows, D]
  int in_group, out_group, out_off;
  const void* x; int x_f32;
  const float* add; int add_period, add_group;
  const float* gamma;
  const float* mean; const float* rstd;
  const void* dres; int dres_f32;   // optional residual-path gradient added to dx, [rows, D]
  void* dx; int dx_f32;             // optional output [rows, D] (null: parameter gradients only)
  float* part;                      // [gridDim.x, 2, D] partial sums of dgamma / dbeta
  int rows, D;
};

__global__ void __launch_bounds__(LN_THREADS) ln_bwd_kernel(const LnBwdArgs a) {
  __shared__ float sh[LN_THREADS / 32];
  const int nchunk = a.D >> 3;
  float pg[LN_MAXC][8], pb[LN_MAXC][8];
#pragma unroll
  for (int i = 0; i < LN_MAXC; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[i][j] = 0.0f; pb[i][j] = 0.0f; }

  for (int row = blockIdx.x; row < a.rows; row += gridDim.x) {
    const float mean = a.mean[row], rstd = a.rstd[row];
    const float* addp = a.add ? a.add + static_cast<size_t>((row % a.add_period) / a.add_group) * a.D : nullptr;
    const size_t orow = static_cast<size_t>(row / a.in_group) * a.out_group + a.out_off + row % a.in_group;
    float xh[LN_MAXC][8], dg[LN_MAXC][8];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
      const int c = threadIdx.x + i * LN_THREADS;
      if (c < nchunk) {
        float xv[8], dyv[8], gm[8];
        load8(a.x, a.x_f32, static_cast<size_t>(row) * a.D + c * 8, xv);
        if (addp) {
          float e[8]; load8f(addp + c * 8, e);
#pragma unroll
          for (int j = 0; j < 8; ++j) xv[j] += e[j];
        }
        load8(a.dy, 0, orow * a.D + c * 8, dyv);
        if (a.dy2) {
          float e2[8]; load8(a.dy2, 0, static_cast<size_t>(row) * a.D + c * 8, e2);
#pragma unroll
          for (int j = 0; j < 8; ++j) dyv[j] += e2[j];
        }
        load8f(a.gamma + c * 8, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (xv[j] - mean) * rstd;
          pg[i][j] += dyv[j] * xh[i][j];
          pb[i][j] += dyv[j];
          dg[i][j] = dyv[j] * gm[j];
          s1 += dg[i][j];
          s2 += dg[i][j] * xh[i][j];
        }
      }
    }
    if (a.dx != nullptr) {   // uniform across the block
      const float m1 = block_sum(s1, sh) / a.D;
      const float m2 = block_sum(s2, sh) / a.D;
#pragma unroll
      for (int i = 0; i < LN_MAXC; ++i) {
        const int c = threadIdx.x + i * LN_THREADS;
        if (c < nchunk) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rstd * (dg[i][j] - m1 - xh[i][j] * m2);
          if (a.dres) {
            float r[8]; load8(a.dres, a.dres_f32, static_cast<size_t>(row) * a.D + c * 8, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += r[j];
          }
          store8(a.dx, a.dx_f32, static_cast<size_t>(row) * a.D + c * 8, o);
        }
      }
    }
  }
  float* pgo = a.part + static_cast<size_t>(blockIdx.x) * 2 * a.D;
  float* pbo = pgo + a.D;
#pragma unroll
  for (int i = 0; i < LN_MAXC; ++i) {
    const int c = threadIdx.x + i * LN_THREADS;
    if (c < nchunk) { store8(pgo, 1, c * 8, pg[i]); store8(pbo, 1, c * 8, pb[i]); }
  }
}

// dgamma[d] (+)= sum_p part[p][0][d]; dbeta[d] (+)= sum_p part[p][1][d]
__global__ void ln_bwd_reduce_kernel(const float* part, int nparts, int D, float* dgamma, float* dbeta, int accumulate) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= 2 * D) return;
  float s = 0.0f;
  for (int p = 0; p < nparts; ++p) s += part[static_cast<size_t>(p) * 2 * D + d];
  float* dst = d < D ? dgamma + d : dbeta + (d - D);
  *dst = accumulate ? *dst + s : s;
}

}  // namespace fm
