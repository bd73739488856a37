// Persistent, warp-specialised tcgen05 GEMM for sm_100a (bf16 inputs, fp32 accumulation in TMEM).
//
//   D[m, n] = sum_k A(m, k) * B(n, k)        m < M, n < N, k < K
//
// A and B each live in global memory either "K-major" (contraction index contiguous: X[m*ld + k], the layout of
// nn.Linear inputs and weights) or "MN-major" (X[k*ld + m]); the three combinations a Linear layer needs are
//   fwd   y  = x W^T      (A K-major,  B K-major)
//   dX    dx = dy W       (A K-major,  B MN-major)
//   dW    dw = dy^T x     (A MN-major, B MN-major)
// so no operand is ever transposed in HBM.
//
// Structure (one CTA per SM, 384 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles (128B swizzle) into a STAGES-deep smem ring
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (128 x BN x 16), commits to mbarriers
//   warp 2      TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4-11  epilogue: tcgen05.ld 32x32b -> registers -> fused epilogue -> 128-bit global stores
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Fused epilogues (reference ops they replace, flamingo_mini/…):
//   EPI_STORE  out = acc*scale*tanh(gate) + col_bias           (to_q *scale, to_kv, dX, dW)
//   EPI_ACT    out = act(acc), out2 = acc                        (utils.py:45-50 Linear -> GELU/sqrelu/relu)
//   EPI_RESID  out = resid + tanh(gate)*scale*acc                (gated_cross_attention.py:180,182; perceiver_resampler.py:182-183)
//   EPI_DACT   out = tanh(gate)*acc*act'(pre); red += acc*act(pre)   (backward of the FFW activation + d(alpha_ffw))
#pragma once
#include "ptx.cuh"

namespace fm {

enum GemmEpi : int { EPI_STORE = 0, EPI_ACT = 1, EPI_RESID = 2, EPI_DACT = 3 };

struct GemmArgs {
  int M, N, K;
  void* out;         long long ldo;     // EPI_*: primary output (bf16 unless out_f32)
  void* out2;        long long ldo2;    // EPI_ACT: pre-activation copy (bf16) or null
  const void* aux;   long long ldaux;   // EPI_RESID: residual; EPI_DACT: saved pre-activation (bf16)
  const float* col_bias;                // EPI_STORE: optional [N]
  const float* gate;                    // optional device scalar alpha, factor tanh(*gate)
  float* red_out;                       // EPI_DACT: optional, atomically += sum(acc * act(pre))
  float scale;
  int act;                              // 0 gelu, 1 sqrelu, 2 relu
  int out_f32;
  int aux_f32;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARPS = 8;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192) ? 4 : (BN == 128) ? 6 : 8;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void tile_coords(int tile, int num_mb, int num_nb, int& mb, int& nb) {
  constexpr int GROUP = 8;
  const int per_group = GROUP * num_nb;
  const int gid = tile / per_group;
  const int first_m = gid * GROUP;
  const int gsz = min(num_mb - first_m, GROUP);
  const int r = tile - gid * per_group;
  mb = first_m + r % gsz;
  nb = r / gsz;
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using Cfg = GemmCfg<BN>;
  constexpr int BM = GEMM_BM, BK = GEMM_BK, STAGES = Cfg::STAGES;
  constexpr int A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES;
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN must be a multiple of 64 in [64,256]");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_mb = (g.M + BM - 1) / BM;
  const int num_nb = (g.N + BN - 1) / BN;
  const int num_tiles = num_mb * num_nb;
  const int num_kb = (g.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], GEMM_EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mb, nb; tile_coords(tile, num_mb, num_nb, mb, nb);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1, 0x100 + stage);
          mbar_arrive_expect_tx(&full[stage], A_BYTES + B_BYTES);
          uint8_t* a = sA + stage * A_BYTES;
          uint8_t* b = sB + stage * B_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(a, &tmA, &full[stage], kb * BK, mb * BM);            // box 64(k) x 128(m)
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)                                  // box 64(m) x 64(k), 8 KB each
              tma_load_2d(a + c * 8192, &tmA, &full[stage], mb * BM + c * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(b, &tmB, &full[stage], kb * BK, nb * BN);            // box 64(k) x BN(n)
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(b + c * 8192, &tmB, &full[stage], nb * BN + c * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1, 0x200 + acc);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase, 0x300 + stage);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(sA + stage * A_BYTES);
          const uint32_t b_base = smem_u32(sB + stage * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major SW128: 8-row groups 1024 B apart (SBO); +32 B per 16-element K step.
            // MN-major SW128: 64-wide MN chunks 8192 B apart (LBO); 8-row K groups 1024 B apart (SBO); +2048 B per K step.
            const uint64_t a_desc = A_MN ? umma_smem_desc_sw128(a_base + k * 2048, 8192, 1024)
                                         : umma_smem_desc_sw128(a_base + k * 32, 0, 1024);
            const uint64_t b_desc = B_MN ? umma_smem_desc_sw128(b_base + k * 2048, 8192, 1024)
                                         : umma_smem_desc_sw128(b_base + k * 32, 0, 1024);
            umma_bf16(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);          // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);              // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue (8 warps)
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;          // column half of the tile
    constexpr int HALF_N = BN / 2;
    constexpr int CHUNKS = HALF_N / 32;
    float mul = g.scale;
    if (g.gate != nullptr) mul *= tanhf(__ldg(g.gate));
    int acc = 0; uint32_t acc_phase = 0;
    float red = 0.0f;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int mb, nb; tile_coords(tile, num_mb, num_nb, mb, nb);
      mbar_wait(&tfull[acc], acc_phase, 0x400 + acc);
      tc_fence_after_sync();
      const int row = mb * BM + q * 32 + lane;
      const bool row_ok = row < g.M;
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        const int col_in_tile = half * HALF_N + c * 32;
        const int n0 = nb * BN + col_in_tile;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + col_in_tile), r);
        tmem_ld_wait();
        if (!row_ok || n0 >= g.N) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);

        if constexpr (EPI == EPI_STORE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= mul;
          if (g.col_bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (n0 + j < g.N) v[j] += __ldg(g.col_bias + n0 + j);
          }
        } else if constexpr (EPI == EPI_ACT) {
          if (g.out2 != nullptr) {
            __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(g.out2) + static_cast<size_t>(row) * g.ldo2 + n0;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              if (n0 + j8 * 8 < g.N) {
                uint4 u;
                u.x = pack_bf16x2(v[j8 * 8 + 0], v[j8 * 8 + 1]); u.y = pack_bf16x2(v[j8 * 8 + 2], v[j8 * 8 + 3]);
                u.z = pack_bf16x2(v[j8 * 8 + 4], v[j8 * 8 + 5]); u.w = pack_bf16x2(v[j8 * 8 + 6], v[j8 * 8 + 7]);
                *reinterpret_cast<uint4*>(o2 + j8 * 8) = u;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = act_fwd(v[j], g.act);
        } else if constexpr (EPI == EPI_RESID) {
          if (g.aux_f32) {
            const float* rs = reinterpret_cast<const float*>(g.aux) + static_cast<size_t>(row) * g.ldaux + n0;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              if (n0 + j4 * 4 < g.N) {
                const float4 t = *reinterpret_cast<const float4*>(rs + j4 * 4);
                v[j4 * 4 + 0] = fmaf(mul, v[j4 * 4 + 0], t.x); v[j4 * 4 + 1] = fmaf(mul, v[j4 * 4 + 1], t.y);
                v[j4 * 4 + 2] = fmaf(mul, v[j4 * 4 + 2], t.z); v[j4 * 4 + 3] = fmaf(mul, v[j4 * 4 + 3], t.w);
              }
            }
          } else {
            const __nv_bfloat16* rs = reinterpret_cast<const __nv_bfloat16*>(g.aux) + static_cast<size_t>(row) * g.ldaux + n0;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              if (n0 + j8 * 8 < g.N) {
                const uint4 t = *reinterpret_cast<const uint4*>(rs + j8 * 8);
                const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), cc = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
                v[j8 * 8 + 0] = fmaf(mul, v[j8 * 8 + 0], a.x);  v[j8 * 8 + 1] = fmaf(mul, v[j8 * 8 + 1], a.y);
                v[j8 * 8 + 2] = fmaf(mul, v[j8 * 8 + 2], b.x);  v[j8 * 8 + 3] = fmaf(mul, v[j8 * 8 + 3], b.y);
                v[j8 * 8 + 4] = fmaf(mul, v[j8 * 8 + 4], cc.x); v[j8 * 8 + 5] = fmaf(mul, v[j8 * 8 + 5], cc.y);
                v[j8 * 8 + 6] = fmaf(mul, v[j8 * 8 + 6], d.x);  v[j8 * 8 + 7] = fmaf(mul, v[j8 * 8 + 7], d.y);
              }
            }
          }
        } else {  // EPI_DACT
          const __nv_bfloat16* pre = reinterpret_cast<const __nv_bfloat16*>(g.aux) + static_cast<size_t>(row) * g.ldaux + n0;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            if (n0 + j8 * 8 < g.N) {
              const uint4 t = *reinterpret_cast<const uint4*>(pre + j8 * 8);
              const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 p2 = unpack_bf16x2(w4[e]);
                float f0, f1;
                const float d0 = act_bwd(p2.x, g.act, &f0);
                const float d1 = act_bwd(p2.y, g.act, &f1);
                const int j = j8 * 8 + e * 2;
                red = fmaf(v[j], f0, red);
                red = fmaf(v[j + 1], f1, red);
                v[j] = mul * v[j] * d0;
                v[j + 1] = mul * v[j + 1] * d1;
              }
            }
          }
        }

        // ---- store (row-per-thread, 16 B vectors)
        if (g.out_f32 && (EPI == EPI_STORE || EPI == EPI_RESID)) {
          float* o = reinterpret_cast<float*>(g.out) + static_cast<size_t>(row) * g.ldo + n0;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            if (n0 + j4 * 4 < g.N)
              *reinterpret_cast<float4*>(o + j4 * 4) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(g.out) + static_cast<size_t>(row) * g.ldo + n0;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            if (n0 + j8 * 8 < g.N) {
              uint4 u;
              u.x = pack_bf16x2(v[j8 * 8 + 0], v[j8 * 8 + 1]); u.y = pack_bf16x2(v[j8 * 8 + 2], v[j8 * 8 + 3]);
              u.z = pack_bf16x2(v[j8 * 8 + 4], v[j8 * 8 + 5]); u.w = pack_bf16x2(v[j8 * 8 + 6], v[j8 * 8 + 7]);
              *reinterpret_cast<uint4*>(o + j8 * 8) = u;
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (EPI == EPI_DACT) {
      if (g.red_out != nullptr) {
        red = warp_sum(red);
        if (lane == 0) atomicAdd(g.red_out, red);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace fm
