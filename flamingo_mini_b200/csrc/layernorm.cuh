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
constexpr int LN_MAXC_WIDE = 4;  // 8-element chunks per thread for D in (4096, 8192]; 2 otherwise (more threads per row,
                                 // fewer registers per thread -> more resident warps: these kernels are latency bound)

// Sum over the TPR threads that share one row (TPR = 32: pure shuffles; TPR > 32: one smem hop between the warps
// of the row group).  Must be called by every thread of the CTA.
template <int TPR>
__device__ __forceinline__ float row_sum(float v, float* sh) {
  v = warp_sum(v);
  if constexpr (TPR > 32) {
    constexpr int WPR = TPR / 32;                      // warps per row
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    const int g0 = (w / WPR) * WPR;
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < WPR; ++i) t += sh[g0 + i];
    v = t;
  }
  return v;
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


// TPR threads cooperate on one row (256/TPR rows per CTA).  Rows are software-pipelined: the loads of the CTA's next
// row are issued before the current row's reductions, so two rows per row-group are in flight (these kernels are bound
// by the load -> reduce -> store latency chain, not by bandwidth: r01_ln_bench.txt).
template <int TPR, int LN_MAXC>
__device__ __forceinline__ void ln_fwd_load(const LnArgs& a, int row, int tig, int nchunk, float (&v)[LN_MAXC][8]) {
  const bool rv = row < a.rows;
  const float* addp = (a.add && rv) ? a.add + static_cast<size_t>((row % a.add_period) / a.add_group) * a.D : nullptr;
#pragma unroll
  for (int i = 0; i < LN_MAXC; ++i) {
    const int c = tig + i * TPR;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[i][j] = 0.0f;
    if (rv && c < nchunk) {
      load8(a.x, a.x_f32, static_cast<size_t>(row) * a.D + c * 8, v[i]);
      if (addp) {
        float e[8]; load8f(addp + c * 8, e);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] += e[j];
      }
    }
  }
}

template <int TPR, int LN_MAXC>
__global__ void __launch_bounds__(LN_THREADS) ln_fwd_kernel(const LnArgs a) {
  __shared__ float sh[LN_THREADS / 32];
  constexpr int RPC = LN_THREADS / TPR;
  const int nchunk = a.D >> 3;
  const int grp = threadIdx.x / TPR, tig = threadIdx.x % TPR;
  const int stride = gridDim.x * RPC;
  const int niter = (a.rows + stride - 1) / stride;
  float v[LN_MAXC][8], nx[LN_MAXC][8];
  pdl_launch_dependents();
  pdl_wait();
  ln_fwd_load<TPR, LN_MAXC>(a, blockIdx.x * RPC + grp, tig, nchunk, v);
  for (int it = 0; it < niter; ++it) {
    const int row = (it * gridDim.x + blockIdx.x) * RPC + grp;
    const bool rv = row < a.rows;
    if (it + 1 < niter) ln_fwd_load<TPR, LN_MAXC>(a, row + stride, tig, nchunk, nx);     // prefetch the next row
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];                  // padding chunks hold zeros
    const float mean = row_sum<TPR>(s, sh) / a.D;
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
      const int c = tig + i * TPR;
      if (c < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; ss += d * d; }
      }
    }
    const float rstd = rsqrtf(row_sum<TPR>(ss, sh) / a.D + 1e-5f);
    if (rv) {
      if (tig == 0) {
        if (a.mean) a.mean[row] = mean;
        if (a.rstd) a.rstd[row] = rstd;
      }
      const size_t orow = static_cast<size_t>(row / a.in_group) * a.out_group + a.out_off + row % a.in_group;
#pragma unroll
      for (int i = 0; i < LN_MAXC; ++i) {
        const int c = tig + i * TPR;
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
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = nx[i][j];
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


template <int TPR, int LN_MAXC>
__device__ __forceinline__ void ln_bwd_load(const LnBwdArgs& a, int row, int tig, int nchunk, float (&xv)[LN_MAXC][8],
                                            float (&dyv)[LN_MAXC][8]) {
  const bool rv = row < a.rows;
  const float* addp = (a.add && rv) ? a.add + static_cast<size_t>((row % a.add_period) / a.add_group) * a.D : nullptr;
  const size_t orow = rv ? static_cast<size_t>(row / a.in_group) * a.out_group + a.out_off + row % a.in_group : 0;
#pragma unroll
  for (int i = 0; i < LN_MAXC; ++i) {
    const int c = tig + i * TPR;
#pragma unroll
    for (int j = 0; j < 8; ++j) { xv[i][j] = 0.0f; dyv[i][j] = 0.0f; }
    if (rv && c < nchunk) {
      load8(a.x, a.x_f32, static_cast<size_t>(row) * a.D + c * 8, xv[i]);
      if (addp) {
        float e[8]; load8f(addp + c * 8, e);
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[i][j] += e[j];
      }
      load8(a.dy, 0, orow * a.D + c * 8, dyv[i]);
      if (a.dy2) {
        float e2[8]; load8(a.dy2, 0, static_cast<size_t>(row) * a.D + c * 8, e2);
#pragma unroll
        for (int j = 0; j < 8; ++j) dyv[i][j] += e2[j];
      }
    }
  }
}

// Each element is loaded once and kept in registers for both passes; the next row's loads are issued before the current
// row's reductions (two rows in flight per row group).
template <int TPR, int LN_MAXC>
__global__ void __launch_bounds__(LN_THREADS, (LN_MAXC > 2 ? 1 : 2)) ln_bwd_kernel(const LnBwdArgs a) {
  __shared__ float sh[LN_THREADS / 32];
  constexpr int RPC = LN_THREADS / TPR;
  FM_DYN_SMEM(float, sacc);                // [2][D] cross-row-group accumulators, only when RPC > 1
  const int nchunk = a.D >> 3;
  const int grp = threadIdx.x / TPR, tig = threadIdx.x % TPR;
  float pg[LN_MAXC][8], pb[LN_MAXC][8];
  pdl_launch_dependents();
#pragma unroll
  for (int i = 0; i < LN_MAXC; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[i][j] = 0.0f; pb[i][j] = 0.0f; }
  if constexpr (RPC > 1) {
    for (int i = threadIdx.x; i < 2 * a.D; i += LN_THREADS) sacc[i] = 0.0f;
  }

  const int stride = gridDim.x * RPC;
  const int niter = (a.rows + stride - 1) / stride;
  float xv[LN_MAXC][8], dyv[LN_MAXC][8], nxv[LN_MAXC][8], ndy[LN_MAXC][8];
  pdl_wait();
  ln_bwd_load<TPR, LN_MAXC>(a, blockIdx.x * RPC + grp, tig, nchunk, xv, dyv);
  for (int it = 0; it < niter; ++it) {
    const int row = (it * gridDim.x + blockIdx.x) * RPC + grp;
    const bool rv = row < a.rows;
    if (it + 1 < niter) ln_bwd_load<TPR, LN_MAXC>(a, row + stride, tig, nchunk, nxv, ndy);   // prefetch the next row
    const float mean = rv ? a.mean[row] : 0.0f, rstd = rv ? a.rstd[row] : 0.0f;
    // pass 1 (registers): xhat, statistics of dy*gamma, per-column partial sums.  Padding chunks / rows hold zeros.
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
      const int c = tig + i * TPR;
      if (rv && c < nchunk) {
        float gm[8];
        load8f(a.gamma + c * 8, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[i][j] - mean) * rstd;
          const float dg = dyv[i][j] * gm[j];
          pg[i][j] += dyv[i][j] * xh;
          pb[i][j] += dyv[i][j];
          s1 += dg;
          s2 += dg * xh;
          xv[i][j] = xh;            // keep xhat and dy*gamma for pass 2
          dyv[i][j] = dg;
        }
      }
    }
    if (a.dx != nullptr) {   // uniform across the block
      const float m1 = row_sum<TPR>(s1, sh) / a.D;
      const float m2 = row_sum<TPR>(s2, sh) / a.D;
#pragma unroll
      for (int i = 0; i < LN_MAXC; ++i) {
        const int c = tig + i * TPR;
        if (rv && c < nchunk) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rstd * (dyv[i][j] - m1 - xv[i][j] * m2);
          if (a.dres) {
            float r[8]; load8(a.dres, a.dres_f32, static_cast<size_t>(row) * a.D + c * 8, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += r[j];
          }
          store8(a.dx, a.dx_f32, static_cast<size_t>(row) * a.D + c * 8, o);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) { xv[i][j] = nxv[i][j]; dyv[i][j] = ndy[i][j]; }
  }
  float* pgo = a.part + static_cast<size_t>(blockIdx.x) * 2 * a.D;
  float* pbo = pgo + a.D;
  if constexpr (RPC > 1) {   // fold the row groups of this CTA together first (element (c, j) at j*nchunk + c: consecutive
    __syncthreads();         // lanes hit consecutive banks, so the shared-memory atomics are conflict-free)
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
      const int c = tig + i * TPR;
      if (c < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(&sacc[j * nchunk + c], pg[i][j]); atomicAdd(&sacc[a.D + j * nchunk + c], pb[i][j]); }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.D; i += LN_THREADS) {
      const int c = i >> 3, j = i & 7;
      pgo[i] = sacc[j * nchunk + c];
      pbo[i] = sacc[a.D + j * nchunk + c];
    }
  } else {
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
      const int c = tig + i * TPR;
      if (c < nchunk) { store8(pgo, 1, c * 8, pg[i]); store8(pbo, 1, c * 8, pb[i]); }
    }
  }
}


// ================================================================================================ warp-per-row variants
// D <= LN_WARP_MAX_D (the widths of every BASELINE config except the 2048 / 4096-wide LMs): ONE WARP owns a row, lane l holds
// the 8-element chunks l, l+32, ... (MAXC of them), all reductions are warp shuffles.  No block barrier, no shared memory:
// every warp streams rows independently, so the kernels are bound by bytes in flight instead of by a load -> barrier ->
// reduce -> barrier chain shared by the row groups of a CTA (what the TPR kernels above measured at ~2 TB/s for 4096 x 768:
// profiles/r02_validate_next/summary.log).  The backward is split in two: `ln_bwd_dx_w` (critical path: dx only, no column
// accumulators -> few registers, high occupancy) and `ln_bwd_dgb_w` (dgamma / dbeta column sums; a LEAF of the backward graph,
// issued on the side stream) which re-reads x and dy — bytes are cheap here, latency on the critical path is not.
constexpr int LN_WARP_MAX_D = 1536;

template <int MAXC>
__device__ __forceinline__ void lnw_load_x(const void* x, int x_f32, const float* addp, size_t row_off, int lane, int nchunk,
                                           float (&v)[MAXC][8]) {
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + i * 32;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[i][j] = 0.0f;
    if (c < nchunk) {
      load8(x, x_f32, row_off + c * 8, v[i]);
      if (addp) {
        float e[8]; load8f(addp + c * 8, e);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] += e[j];
      }
    }
  }
}

template <int MAXC>
__global__ void __launch_bounds__(LN_THREADS) ln_fwd_w_kernel(const LnArgs a) {
  const int nchunk = a.D >> 3;
  const int lane = threadIdx.x & 31;
  const int wpc = LN_THREADS / 32;
  const int warp0 = blockIdx.x * wpc + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * wpc;
  pdl_launch_dependents();
  pdl_wait();
  for (int row = warp0; row < a.rows; row += nwarps) {
    const float* addp = a.add ? a.add + static_cast<size_t>((row % a.add_period) / a.add_group) * a.D : nullptr;
    float v[MAXC][8];
    lnw_load_x<MAXC>(a.x, a.x_f32, addp, static_cast<size_t>(row) * a.D, lane, nchunk, v);
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];                  // padding chunks hold zeros
    const float mean = warp_sum(s) / a.D;
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      if (lane + i * 32 < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; ss += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / a.D + 1e-5f);
    if (lane == 0) {
      if (a.mean) a.mean[row] = mean;
      if (a.rstd) a.rstd[row] = rstd;
    }
    const size_t orow = static_cast<size_t>(row / a.in_group) * a.out_group + a.out_off + row % a.in_group;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + i * 32;
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

// xhat and dy (+ dy2) of one row into registers
template <int MAXC>
__device__ __forceinline__ void lnw_load_bwd(const LnBwdArgs& a, int row, int lane, int nchunk, float mean, float rstd,
                                             float (&xh)[MAXC][8], float (&dyv)[MAXC][8]) {
  const float* addp = a.add ? a.add + static_cast<size_t>((row % a.add_period) / a.add_group) * a.D : nullptr;
  const size_t orow = static_cast<size_t>(row / a.in_group) * a.out_group + a.out_off + row % a.in_group;
  lnw_load_x<MAXC>(a.x, a.x_f32, addp, static_cast<size_t>(row) * a.D, lane, nchunk, xh);
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + i * 32;
#pragma unroll
    for (int j = 0; j < 8; ++j) dyv[i][j] = 0.0f;
    if (c < nchunk) {
      load8(a.dy, 0, orow * a.D + c * 8, dyv[i]);
      if (a.dy2) {
        float e2[8]; load8(a.dy2, 0, static_cast<size_t>(row) * a.D + c * 8, e2);
#pragma unroll
        for (int j = 0; j < 8; ++j) dyv[i][j] += e2[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) xh[i][j] = (xh[i][j] - mean) * rstd;
    }
  }
}

// dx = rstd * (dy*gamma - mean_D(dy*gamma) - xhat * mean_D(dy*gamma*xhat)) [+ dres]
template <int MAXC>
__global__ void __launch_bounds__(LN_THREADS) ln_bwd_dx_w_kernel(const LnBwdArgs a) {
  const int nchunk = a.D >> 3;
  const int lane = threadIdx.x & 31;
  const int wpc = LN_THREADS / 32;
  const int warp0 = blockIdx.x * wpc + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * wpc;
  pdl_launch_dependents();
  pdl_wait();
  for (int row = warp0; row < a.rows; row += nwarps) {
    const float mean = a.mean[row], rstd = a.rstd[row];
    float xh[MAXC][8], dg[MAXC][8];
    lnw_load_bwd<MAXC>(a, row, lane, nchunk, mean, rstd, xh, dg);
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + i * 32;
      if (c < nchunk) {
        float gm[8];
        load8f(a.gamma + c * 8, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) { dg[i][j] *= gm[j]; s1 += dg[i][j]; s2 += dg[i][j] * xh[i][j]; }
      }
    }
    const float m1 = warp_sum(s1) / a.D, m2 = warp_sum(s2) / a.D;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + i * 32;
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

// per-CTA partial column sums: part[cta][0][d] = sum_rows dy*xhat, part[cta][1][d] = sum_rows dy; folded by ln_bwd_reduce_kernel
template <int MAXC>
__global__ void __launch_bounds__(LN_THREADS) ln_bwd_dgb_w_kernel(const LnBwdArgs a) {
  FM_DYN_SMEM(float, sacc);                // [2][D]
  const int nchunk = a.D >> 3;
  const int lane = threadIdx.x & 31;
  const int wpc = LN_THREADS / 32;
  const int warp0 = blockIdx.x * wpc + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * wpc;
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 2 * a.D; i += LN_THREADS) sacc[i] = 0.0f;
  float pg[MAXC][8], pb[MAXC][8];
#pragma unroll
  for (int i = 0; i < MAXC; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[i][j] = 0.0f; pb[i][j] = 0.0f; }
  pdl_wait();
  for (int row = warp0; row < a.rows; row += nwarps) {
    const float mean = a.mean[row], rstd = a.rstd[row];
    float xh[MAXC][8], dyv[MAXC][8];
    lnw_load_bwd<MAXC>(a, row, lane, nchunk, mean, rstd, xh, dyv);
#pragma unroll
    for (int i = 0; i < MAXC; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) { pg[i][j] += dyv[i][j] * xh[i][j]; pb[i][j] += dyv[i][j]; }     // padding chunks hold zeros
  }
  __syncthreads();                           // sacc zeroed
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + i * 32;
    if (c < nchunk) {                        // element (c, j) at j*nchunk + c: consecutive lanes hit consecutive banks
#pragma unroll
      for (int j = 0; j < 8; ++j) { atomicAdd(&sacc[j * nchunk + c], pg[i][j]); atomicAdd(&sacc[a.D + j * nchunk + c], pb[i][j]); }
    }
  }
  __syncthreads();
  float* pgo = a.part + static_cast<size_t>(blockIdx.x) * 2 * a.D;
  float* pbo = pgo + a.D;
  for (int i = threadIdx.x; i < a.D; i += LN_THREADS) {
    const int c = i >> 3, j = i & 7;
    pgo[i] = sacc[j * nchunk + c];
    pbo[i] = sacc[a.D + j * nchunk + c];
  }
}

// dgamma[d] = sum_p part[p][0][d]; dbeta[d] = sum_p part[p][1][d].  Block = 32 columns x 8 partial groups.
__global__ void __launch_bounds__(256) ln_bwd_reduce_kernel(const float* part, int nparts, int D, float* dgamma, float* dbeta,
                                                            int accumulate) {
  __shared__ float sm[8][33];
  pdl_launch_dependents();
  pdl_wait();
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int g = threadIdx.x >> 5;
  float s = 0.0f;
  if (col < 2 * D)
    for (int p = g; p < nparts; p += 8) s += part[static_cast<size_t>(p) * 2 * D + col];
  sm[g][threadIdx.x & 31] = s;
  __syncthreads();
  if (g == 0 && col < 2 * D) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    float* dst = col < D ? dgamma + col : dbeta + (col - D);
    *dst = accumulate ? *dst + t : t;
  }
}

}  // namespace fm
