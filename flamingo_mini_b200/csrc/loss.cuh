// Shifted next-token cross-entropy over the (padded) lm_head logits — the loss head of modeling_flamingo.py:287-298
// (reference: CrossEntropyLoss over logits[..., :-1, :] / labels[..., 1:]).  SURVEY.md §8(f)-3.
// HBM-bound row kernels: one CTA per row of `vocab` bf16 logits (100 KB at GPT-2's 50 258), 128-bit loads, four loads
// in flight per thread, fp32 online log-sum-exp in base 2 (one ex2 per element).  The forward keeps ONLY lse[row]
// (no log-softmax tensor); the backward writes d(logits) = (softmax - onehot) * scale in one pass.
#pragma once
#include "ptx.cuh"

namespace fm {

constexpr int CE_THREADS = 256;
constexpr float CE_LOG2E = 1.4426950408889634f;
constexpr float CE_LN2 = 0.6931471805599453f;

struct CeArgs {
  const __nv_bfloat16* logits;   // [rows, ld]; columns [vocab, ld) are padding and never read
  const long long* targets;      // [rows]; ignore_index rows contribute neither loss nor gradient
  long long ignore_index;
  float* lse;                    // [rows] natural-log sum-exp of the row
  float* row_loss;               // [rows] lse - logit[target] (0 for ignored rows)
  __nv_bfloat16* dlogits;        // backward: [rows, ld]; padding columns are written as zeros
  const float* scale;            // backward: device scalar d(loss) / n_valid
  int rows, vocab;
  long long ld;
};

__device__ __forceinline__ void ce_unpack8(const uint4& u, float (&v)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

// online (max, sum) in base 2: values are pre-multiplied by log2(e); s = sum 2^(x - m)
__device__ __forceinline__ void ce_accumulate(const float (&v)[8], int col0, int vocab, float& m, float& s) {
  float x[8];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    x[j] = (col0 + j < vocab) ? v[j] * CE_LOG2E : -INFINITY;
    mx = fmaxf(mx, x[j]);
  }
  if (mx == -INFINITY) return;                       // chunk entirely past the vocabulary (or all -inf logits)
  const float mn = fmaxf(m, mx);
  float acc = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc += ex2_approx(x[j] - mn);     // ex2(-inf) = 0
  s = s * ex2_approx(m - mn) + acc;                  // m = -inf on the first chunk: ex2(-inf) = 0, s was 0
  m = mn;
}

__global__ void __launch_bounds__(CE_THREADS) ce_fwd_kernel(const CeArgs a) {
  __shared__ float sm_m[CE_THREADS / 32], sm_s[CE_THREADS / 32];
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  const __nv_bfloat16* src = a.logits + static_cast<size_t>(row) * a.ld;
  const int nchunk = (a.vocab + 7) >> 3;
  float m = -INFINITY, s = 0.0f;
  int c = threadIdx.x;
  for (; c + 3 * CE_THREADS < nchunk; c += 4 * CE_THREADS) {        // four independent 16-byte loads in flight
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(c + i * CE_THREADS) * 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v[8];
      ce_unpack8(u[i], v);
      ce_accumulate(v, (c + i * CE_THREADS) * 8, a.vocab, m, s);
    }
  }
  for (; c < nchunk; c += CE_THREADS) {
    // the last chunk of a row may straddle `vocab`; it is still inside the padded row (ld is a multiple of 8)
    const uint4 u = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(c) * 8);
    float v[8];
    ce_unpack8(u, v);
    ce_accumulate(v, c * 8, a.vocab, m, s);
  }
  // CTA reduction of (m, s)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, m2);
    s = (mn == -INFINITY) ? 0.0f : s * ex2_approx(m - mn) + s2 * ex2_approx(m2 - mn);
    m = mn;
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sm_m[w] = m; sm_s[w] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = -INFINITY, S = 0.0f;
#pragma unroll
    for (int i = 0; i < CE_THREADS / 32; ++i) {
      const float mn = fmaxf(M, sm_m[i]);
      if (mn != -INFINITY) S = S * ex2_approx(M - mn) + sm_s[i] * ex2_approx(sm_m[i] - mn);
      M = mn;
    }
    const float lse = (M + log2f(S)) * CE_LN2;       // back to natural log
    a.lse[row] = lse;
    const long long t = a.targets[row];
    float loss = 0.0f;
    if (t != a.ignore_index && t >= 0 && t < a.vocab) loss = lse - __bfloat162float(src[t]);
    a.row_loss[row] = loss;
  }
}

__global__ void __launch_bounds__(CE_THREADS) ce_bwd_kernel(const CeArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  const __nv_bfloat16* src = a.logits + static_cast<size_t>(row) * a.ld;
  __nv_bfloat16* dst = a.dlogits + static_cast<size_t>(row) * a.ld;
  const long long t = a.targets[row];
  const bool live = (t != a.ignore_index && t >= 0 && t < a.vocab);
  const float scale = live ? __ldg(a.scale) : 0.0f;
  const float lse2 = a.lse[row] * CE_LOG2E;
  const int nchunk = static_cast<int>(a.ld >> 3);                   // padding columns included: they get zeros
  const int nread = live ? (a.vocab + 7) >> 3 : 0;                  // chunks that hold at least one real logit
  for (int c0 = threadIdx.x; c0 < nchunk; c0 += 4 * CE_THREADS) {   // four independent 16-byte loads in flight
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + i * CE_THREADS;
      u[i] = (c < nread) ? *reinterpret_cast<const uint4*>(src + static_cast<size_t>(c) * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + i * CE_THREADS;
      if (c >= nchunk) break;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (c < nread) {
        const int col0 = c * 8;
        float v[8];
        ce_unpack8(u[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float p = (col0 + j < a.vocab) ? ex2_approx(fmaf(v[j], CE_LOG2E, -lse2)) : 0.0f;
          if (col0 + j == t) p -= 1.0f;
          v[j] = p * scale;
        }
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      }
      *reinterpret_cast<uint4*>(dst + static_cast<size_t>(c) * 8) = o;
    }
  }
}

}  // namespace fm
