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
  int splits;                           // > 1: serial (deterministic) split-K, fp32 EPI_STORE only
  int* flags;                           // split-K: zero-initialised, 8 ints per output tile, self re-arming
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
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/ +
                                    GEMM_EPI_WARPS * 4096 /*epilogue staging*/ + 768 /*keeps staging 1024-aligned*/;
};


// ----------------------------------------------------------------------------- epilogue staging tile (per warp, 4 KB)
// 32 rows x 128 B; the 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4): conflict-free both for
// "thread = row" accesses and for "8 lanes = one row" (coalesced) accesses.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stage_put_bf16(uint8_t* stg, int r, const float* v /*64*/) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 u;
    u.x = pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]); u.y = pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]);
    u.z = pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]); u.w = pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]);
    *reinterpret_cast<uint4*>(stg + r * 128 + ((c ^ (r & 7)) << 4)) = u;
  }
}
__device__ __forceinline__ void stage_put_f32(uint8_t* stg, int r, const float* v /*32*/) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(stg + r * 128 + ((c ^ (r & 7)) << 4)) = make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
}
// staging tile -> global rows [m0, m0+32) x 128 B starting at element column n0 (ES = element size). Coalesced.
template <int ES, bool ACCUM>
__device__ __forceinline__ void stage_flush(const uint8_t* stg, void* out, long long ld, int m0, int n0, int M, int N, int crow, int cchk) {
  const int col = n0 + cchk * (16 / ES);
  if (col >= N) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + crow;
    if (m0 + r < M) {
      uint4 u = *reinterpret_cast<const uint4*>(stg + r * 128 + ((cchk ^ (r & 7)) << 4));
      uint8_t* dst = reinterpret_cast<uint8_t*>(out) + (static_cast<size_t>(m0 + r) * ld + col) * ES;
      if constexpr (ACCUM) {
        float4 o;
        asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "l"(dst));
        float4 n = *reinterpret_cast<float4*>(&u);
        n.x += o.x; n.y += o.y; n.z += o.z; n.w += o.w;
        *reinterpret_cast<float4*>(dst) = n;
      } else {
        *reinterpret_cast<uint4*>(dst) = u;
      }
    }
  }
}
// global rows [m0, m0+32) x 128 B starting at element column n0 -> staging tile (zeros outside the matrix). Coalesced.
template <int ES>
__device__ __forceinline__ void stage_fill(uint8_t* stg, const void* src, long long ld, int m0, int n0, int M, int N, int crow, int cchk) {
  const int col = n0 + cchk * (16 / ES);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + crow;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (m0 + r < M && col < N)
      u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(src) + (static_cast<size_t>(m0 + r) * ld + col) * ES);
    *reinterpret_cast<uint4*>(stg + r * 128 + ((cchk ^ (r & 7)) << 4)) = u;
  }
}

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
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);   // barriers live in the 256 B before the staging tiles
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* stage_base = sB + STAGES * B_BYTES + 256;               // 8 x 4 KB epilogue staging tiles (1024-aligned)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_mb = (g.M + BM - 1) / BM;
  const int num_nb = (g.N + BN - 1) / BN;
  const int num_tiles = num_mb * num_nb;
  const int num_kb = (g.K + BK - 1) / BK;
  const int splits = g.splits > 1 ? g.splits : 1;                 // serial split-K (fp32 EPI_STORE only)
  const int kb_per_split = (num_kb + splits - 1) / splits;
  const int num_units = num_tiles * splits;                        // unit = split * num_tiles + tile (split-major)

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
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int split = unit / num_tiles;
        const int tile = unit - split * num_tiles;
        int mb, nb; tile_coords(tile, num_mb, num_nb, mb, nb);
        const int kb_end = min(num_kb, (split + 1) * kb_per_split);
        for (int kb = split * kb_per_split; kb < kb_end; ++kb) {
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
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int split = unit / num_tiles;
        mbar_wait(&tempty[acc], acc_phase ^ 1, 0x200 + acc);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        const int kb_begin = split * kb_per_split;
        const int kb_end = min(num_kb, kb_begin + kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
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
            umma_bf16(d_tmem, a_desc, b_desc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
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
    // Warp (q, cs): TMEM lane quarter q = warp % 4 (rows 32q..32q+31 of the tile), column set cs = (warp-4)/4 takes the
    // 64-column groups g = cs, cs+2, ...  Each thread owns one accumulator row; results are transposed through a
    // per-warp 4 KB XOR-swizzled staging tile so that every global access is a full 128-byte line per 8 lanes.
    const int q = warp & 3;
    const int cs = (warp - 4) >> 2;
    uint8_t* stg = stage_base + (warp - 4) * 4096;
    float mul = g.scale;
    if (g.gate != nullptr) mul *= tanhf(__ldg(g.gate));
    int acc = 0; uint32_t acc_phase = 0;
    float red = 0.0f;
    const int srow = lane;                           // row this thread owns inside the 32-row staging tile
    const int crow = lane >> 3, cchk = lane & 7;     // coalesced phase: rows crow, crow+4, ...; 16-byte chunk cchk
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int split = unit / num_tiles;
      const int tile = unit - split * num_tiles;
      int mb, nb; tile_coords(tile, num_mb, num_nb, mb, nb);
      mbar_wait(&tfull[acc], acc_phase, 0x400 + acc);
      tc_fence_after_sync();
      const int m0 = mb * BM + q * 32;
      int* myflag = nullptr;
      if (splits > 1) {
        // Serial (deterministic) split-K: warp position w of split s adds onto what the same warp position of split
        // s-1 left in `out`; one flag per (tile, epilogue warp) holds the number of splits already folded in.
        myflag = g.flags + tile * GEMM_EPI_WARPS + (warp - 4);
        if (split > 0) {
          if (lane == 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(myflag) != split) {
              if (clock64() - t0 > 4000000000LL) { atomicExch(&g_fm_device_error, 0x80000500u); __trap(); }
            }
          }
          __syncwarp();
        }
      }
#pragma unroll 1
      for (int grp = cs; grp < BN / 64; grp += 2) {
        const int col_in_tile = grp * 64;
        const int n0 = nb * BN + col_in_tile;
        if (n0 >= g.N) break;                        // warp-uniform
        float v[64];
        {
          uint32_t r0[32], r1[32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + col_in_tile);
          tmem_ld_32x32(taddr, r0);
          tmem_ld_32x32(taddr + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
        }

        if constexpr (EPI == EPI_STORE) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] *= mul;
          if (g.col_bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 64; ++j) if (n0 + j < g.N) v[j] += __ldg(g.col_bias + n0 + j);
          }
        } else if constexpr (EPI == EPI_ACT) {
          if (g.out2 != nullptr) {
            stage_put_bf16(stg, srow, v);
            __syncwarp();
            stage_flush<2, false>(stg, g.out2, g.ldo2, m0, n0, g.M, g.N, crow, cchk);
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = act_fwd_fast(v[j], g.act);
        } else if constexpr (EPI == EPI_RESID) {
          if (g.aux_f32) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              stage_fill<4>(stg, g.aux, g.ldaux, m0, n0 + h2 * 32, g.M, g.N, crow, cchk);
              __syncwarp();
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 t = *reinterpret_cast<const float4*>(stg + srow * 128 + ((c ^ (srow & 7)) << 4));
                const int j = h2 * 32 + c * 4;
                v[j] = fmaf(mul, v[j], t.x); v[j + 1] = fmaf(mul, v[j + 1], t.y);
                v[j + 2] = fmaf(mul, v[j + 2], t.z); v[j + 3] = fmaf(mul, v[j + 3], t.w);
              }
              __syncwarp();
            }
          } else {
            stage_fill<2>(stg, g.aux, g.ldaux, m0, n0, g.M, g.N, crow, cchk);
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 t = *reinterpret_cast<const uint4*>(stg + srow * 128 + ((c ^ (srow & 7)) << 4));
              const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), cc = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
              const int j = c * 8;
              v[j] = fmaf(mul, v[j], a.x);          v[j + 1] = fmaf(mul, v[j + 1], a.y);
              v[j + 2] = fmaf(mul, v[j + 2], b.x);  v[j + 3] = fmaf(mul, v[j + 3], b.y);
              v[j + 4] = fmaf(mul, v[j + 4], cc.x); v[j + 5] = fmaf(mul, v[j + 5], cc.y);
              v[j + 6] = fmaf(mul, v[j + 6], d.x);  v[j + 7] = fmaf(mul, v[j + 7], d.y);
            }
            __syncwarp();
          }
        } else {  // EPI_DACT
          stage_fill<2>(stg, g.aux, g.ldaux, m0, n0, g.M, g.N, crow, cchk);
          __syncwarp();
          const bool row_ok = (m0 + srow) < g.M;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 t = *reinterpret_cast<const uint4*>(stg + srow * 128 + ((c ^ (srow & 7)) << 4));
            const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
            const bool ok = row_ok && (n0 + c * 8) < g.N;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 p2 = unpack_bf16x2(w4[e]);
              float f0, f1;
              const float d0 = act_bwd_fast(p2.x, g.act, &f0);
              const float d1 = act_bwd_fast(p2.y, g.act, &f1);
              const int j = c * 8 + e * 2;
              if (ok) { red = fmaf(v[j], f0, red); red = fmaf(v[j + 1], f1, red); }
              v[j] = mul * v[j] * d0;
              v[j + 1] = mul * v[j + 1] * d1;
            }
          }
          __syncwarp();
        }

        // ---- store through the staging tile
        if (g.out_f32 && (EPI == EPI_STORE || EPI == EPI_RESID)) {
          const bool accum = (EPI == EPI_STORE) && splits > 1 && split > 0;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            stage_put_f32(stg, srow, v + h2 * 32);
            __syncwarp();
            if (accum) stage_flush<4, true>(stg, g.out, g.ldo, m0, n0 + h2 * 32, g.M, g.N, crow, cchk);
            else       stage_flush<4, false>(stg, g.out, g.ldo, m0, n0 + h2 * 32, g.M, g.N, crow, cchk);
            __syncwarp();
          }
        } else {
          stage_put_bf16(stg, srow, v);
          __syncwarp();
          stage_flush<2, false>(stg, g.out, g.ldo, m0, n0, g.M, g.N, crow, cchk);
          __syncwarp();
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (splits > 1) {                              // publish this warp's part of the tile (last split re-arms the flag)
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          st_release_gpu(myflag, split == splits - 1 ? 0 : split + 1);
        }
      }
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
