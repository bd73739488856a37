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
// Structure (one CTA per SM, 640 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles (128B swizzle) into a STAGES-deep smem ring
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (128 x BN x 16), commits to mbarriers
//   warp 2      TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4-19  epilogue: tcgen05.ld 32x32b -> registers -> fused epilogue -> per-warp 64B-swizzled staging tile -> TMA store;
//               epilogue INPUTS (residual / saved activation tiles) arrive the same way in reverse: a TMA load into a second
//               per-warp staging tile, issued one work item AHEAD (the first one before the accumulator is even waited for),
//               so no epilogue warp ever sits on a global-load round trip (round-2 ncu: 30 % of the DACT kernel's samples
//               were exactly that wait, profiles/r02_call2/ncu_in_situ_metrics.txt)
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Fused epilogues (reference ops they replace, flamingo_mini/…):
//   EPI_STORE  out = acc*scale*tanh(gate) + col_bias           (to_q *scale, to_kv, dX, dW)
//   EPI_ACT    out = act(acc), out2 = act'(acc)                  (utils.py:45-50 Linear -> GELU/sqrelu/relu)
//   EPI_RESID  out = resid + tanh(gate)*scale*acc                (gated_cross_attention.py:180,182; perceiver_resampler.py:182-183)
//   EPI_DACT   out = tanh(gate)*acc*aux                             (backward of the FFW activation; aux = saved act')
#pragma once
#include "ptx.cuh"

namespace fm {

enum GemmEpi : int { EPI_STORE = 0, EPI_ACT = 1, EPI_RESID = 2, EPI_DACT = 3 };

struct GemmArgs {
  int M, N, K;
  void* out;         long long ldo;     // EPI_*: primary output (bf16 unless out_f32)
  void* out2;        long long ldo2;    // EPI_ACT: act'(acc) (bf16) for the backward pass, or null
  const void* aux;   long long ldaux;   // EPI_RESID: residual; EPI_DACT: saved act'(pre) (bf16)
  const float* col_bias;                // EPI_STORE: optional [N]
  const float* gate;                    // optional device scalar alpha, factor tanh(*gate)
  float* red_out;                       // EPI_STORE: optional, atomically += sum(acc * aux) (aux bf16)
  float scale;
  int act;                              // 0 gelu, 1 sqrelu, 2 relu
  int out_f32;
  int aux_f32;
  int splits;                           // > 1: split-K, fp32 EPI_STORE only: serial (deterministic, through `flags`) or parallel
  int par_split;                        // 1: parallel split-K: every split adds its partial tile with a TMA reduce-add (output pre-zeroed)
  int* flags;                           // split-K: zero-initialised, 16 ints (one per epilogue warp) per output tile, self re-arming
  long long* trace;                     // optional debug timeline: [gridDim.x][64] clock64 stamps (see tools/gemm_trace.py)
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 16;                       // 4 per TMEM lane quarter: enough warps to hide ALU/MUFU latency
constexpr int GEMM_THREADS = 128 + GEMM_EPI_WARPS * 32;   // 640
constexpr int GEMM_STG_BYTES = 2048;                     // per-warp staging tile: 32 rows x 64 B

constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_BAR_BYTES = 512;                      // full[8] empty[8] tfull[2] tempty[2] aux[16] + TMEM slot
constexpr int GEMM_SMEM_LIMIT = 232448;                  // 227 KB: the most dynamic shared memory one sm_100 CTA may have

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  // ring depth that fits next to `tiles` staging tiles per epilogue warp (1: outputs only; 2: + one for epilogue inputs)
  static constexpr int stages_for(int tiles) {
    const int avail = GEMM_SMEM_LIMIT - 1024 /*align slack*/ - GEMM_BAR_BYTES - GEMM_EPI_WARPS * GEMM_STG_BYTES * tiles;
    const int n = avail / (A_BYTES + B_BYTES);
    return n > GEMM_MAX_STAGES ? GEMM_MAX_STAGES : n;
  }
  static constexpr int smem_bytes(int stages, int tiles) {
    return stages * (A_BYTES + B_BYTES) + GEMM_EPI_WARPS * GEMM_STG_BYTES * tiles + GEMM_BAR_BYTES + 1024;
  }
};


// ----------------------------------------------------------------------------- epilogue staging tile (per warp, 2 KB)
// 32 rows x 64 B.  The 16-byte chunk c (0..3) of row r lives at r*64 + ((c ^ ((r >> 1) & 3)) << 4), which is
// bank-conflict free both for "thread = row" accesses (each thread moves its own 64 B) and for the coalesced phase
// where 4 consecutive lanes cover one row (8 rows per instruction, full 32-byte sectors in global memory).
#ifndef FM_HOST_EMU
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_global_cg_f4(const void* p) {
  float4 o;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "l"(p));
  return o;
}
#endif
__device__ __forceinline__ uint32_t stg_off(int r, int c) { return static_cast<uint32_t>(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

// shared-address versions (st.shared / ld.shared): the tiles the TMA engine reads and writes
__device__ __forceinline__ void stage_put_s(uint32_t s, int r, const uint4 (&u)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) sts128(s + stg_off(r, c), u[c]);
}
__device__ __forceinline__ void stage_get_s(uint32_t s, int r, uint4 (&u)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) u[c] = lds128(s + stg_off(r, c));
}
__device__ __forceinline__ void stage_put(uint8_t* stg, int r, const uint4 (&u)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(stg + stg_off(r, c)) = u[c];
}
__device__ __forceinline__ void stage_get(const uint8_t* stg, int r, uint4 (&u)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) u[c] = *reinterpret_cast<const uint4*>(stg + stg_off(r, c));
}
// staging tile -> global: rows [m0, m0+32), 64 bytes per row starting at byte offset col_byte of each row.
template <bool ACCUM>
__device__ __forceinline__ void stage_flush(const uint8_t* stg, void* out, size_t ld_bytes, int m0, int M, size_t col_byte,
                                            bool col_ok, int lane) {
  const int cr = lane >> 2, cc = lane & 3;
  if (!col_ok) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + cr;
    if (m0 + r < M) {
      uint4 u = *reinterpret_cast<const uint4*>(stg + stg_off(r, cc));
      uint8_t* dst = reinterpret_cast<uint8_t*>(out) + static_cast<size_t>(m0 + r) * ld_bytes + col_byte + cc * 16;
      if constexpr (ACCUM) {
        const float4 o = ld_global_cg_f4(dst);
        float4 n = *reinterpret_cast<float4*>(&u);
        n.x += o.x; n.y += o.y; n.z += o.z; n.w += o.w;
        *reinterpret_cast<float4*>(dst) = n;
      } else {
        *reinterpret_cast<uint4*>(dst) = u;
      }
    }
  }
}
// global -> staging tile (zeros outside the matrix)
__device__ __forceinline__ void stage_fill(uint8_t* stg, const void* src, size_t ld_bytes, int m0, int M, size_t col_byte,
                                           bool col_ok, int lane) {
  const int cr = lane >> 2, cc = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + cr;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (col_ok && m0 + r < M)
      u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(src) + static_cast<size_t>(m0 + r) * ld_bytes + col_byte + cc * 16);
    *reinterpret_cast<uint4*>(stg + stg_off(r, cc)) = u;
  }
}
__device__ __forceinline__ void pack32(const float (&v)[32], uint4 (&u)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    u[c].x = pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]); u[c].y = pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]);
    u[c].z = pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]); u[c].w = pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]);
  }
}
__device__ __forceinline__ void unpack32(const uint4 (&u)[4], float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float2 a = unpack_bf16x2(u[c].x), b = unpack_bf16x2(u[c].y), cc = unpack_bf16x2(u[c].z), d = unpack_bf16x2(u[c].w);
    v[c * 8 + 0] = a.x; v[c * 8 + 1] = a.y; v[c * 8 + 2] = b.x; v[c * 8 + 3] = b.y;
    v[c * 8 + 4] = cc.x; v[c * 8 + 5] = cc.y; v[c * 8 + 6] = d.x; v[c * 8 + 7] = d.y;
  }
}

__host__ __device__ __forceinline__ void tile_coords(int tile, int num_mb, int num_nb, int& mb, int& nb) {
  constexpr int GROUP = 8;
  const int per_group = GROUP * num_nb;
  const int gid = tile / per_group;
  const int first_m = gid * GROUP;
  const int gsz = min(num_mb - first_m, GROUP);
  const int r = tile - gid * per_group;
  mb = first_m + r % gsz;
  nb = r / gsz;
}

// Up to GEMM_MAX_GROUP independent problems (same layouts / epilogue / tile width) can share one persistent launch: the
// small weight-gradient GEMMs (48-96 tiles each, long K) fill the machine only together.
constexpr int GEMM_MAX_GROUP = 4;
struct GemmGroup {
  CUtensorMap tmA[GEMM_MAX_GROUP], tmB[GEMM_MAX_GROUP];
  CUtensorMap tmOut[GEMM_MAX_GROUP];         // epilogue outputs: box = 64 bytes x 32 rows, SWIZZLE_64B (bf16: 32 columns, fp32: 16)
  CUtensorMap tmOut2;                        // EPI_ACT second output (problem 0)
  CUtensorMap tmAux;                         // epilogue input of problem 0 (residual / saved activation / dot operand), same box
  GemmArgs g[GEMM_MAX_GROUP];
  int nprob;
  int unit_start[GEMM_MAX_GROUP + 1];        // prefix sums of (tiles * splits) per problem
  int stages;                                // depth of the operand ring (host: GemmCfg<BN>::stages_for(tiles))
  int l2_ahead;                              // > 0: the producer also prefetches the operand boxes of k-block kb + l2_ahead into L2
  int tiles;                                 // staging tiles per epilogue warp: 1 (outputs) or 2 (+ epilogue inputs)
};

struct UnitInfo { int p, split, tile, mb, nb, kb_begin, kb_end, splits; };

// GROUPED = false (every epilogue except STORE is launched with one problem): problem 0 is addressed statically, so its
// arguments stay constant-bank operands instead of costing address registers in the epilogue.
template <int BN, bool GROUPED>
__host__ __device__ __forceinline__ UnitInfo locate_unit(const GemmGroup& G, int unit) {
  UnitInfo u;
  u.p = 0;
  if constexpr (GROUPED) {
#pragma unroll
    for (int i = 1; i < GEMM_MAX_GROUP; ++i) if (i < G.nprob && unit >= G.unit_start[i]) u.p = i;
  }
  const GemmArgs& g = GROUPED ? G.g[u.p] : G.g[0];
  const int local = GROUPED ? unit - G.unit_start[u.p] : unit;
  const int num_mb = (g.M + GEMM_BM - 1) / GEMM_BM;
  const int num_nb = (g.N + BN - 1) / BN;
  const int num_tiles = num_mb * num_nb;
  const int num_kb = (g.K + GEMM_BK - 1) / GEMM_BK;
  u.splits = g.splits > 1 ? g.splits : 1;                            // serial split-K (fp32 EPI_STORE only)
  const int kb_per_split = (num_kb + u.splits - 1) / u.splits;
  u.split = local / num_tiles;                                       // unit = split * num_tiles + tile (split-major)
  u.tile = local - u.split * num_tiles;
  tile_coords(u.tile, num_mb, num_nb, u.mb, u.nb);
  u.kb_begin = u.split * kb_per_split;
  u.kb_end = min(num_kb, u.kb_begin + kb_per_split);
  return u;
}

// ----------------------------------------------------------------------------- per-warp epilogue state
// Two staging tiles of 32 rows x 64 bytes per warp, both in the layout of stg_off() == the TMA SWIZZLE_64B pattern:
//   s_out  results, handed to the TMA engine with one cp.async.bulk.tensor store per tile (clipped at the matrix edge)
//   s_in   epilogue inputs, filled by TMA loads that run ONE work item ahead of the arithmetic (zero filled outside the matrix)
// An "aux load unit" is 64 bytes of one row: 32 bf16 columns or 16 fp32 columns, i.e. a bf16 input tile is one unit per item and
// an fp32 one is two (halves).  The issue cursor walks exactly the sequence of (unit, item, half) the consuming loops walk.
struct EpiWarp {
  uint32_t s_in, s_out, s_bar;     // shared addresses: input tile, output tile, the warp's input-arrival barrier
  uint32_t aux_phase;
  int store_inflight;
  int c_unit, c_item, c_half, c_mb, c_nb;      // issue cursor of the input pipeline (c_unit >= num_units: exhausted)
};
__device__ __forceinline__ int epi_lane() { return static_cast<int>(threadIdx.x & 31); }
__device__ __forceinline__ int epi_q() { return static_cast<int>((threadIdx.x >> 5) & 3); }
__device__ __forceinline__ int epi_cs() { return static_cast<int>((threadIdx.x >> 5) - 4) >> 2; }

template <int BN>
__device__ __forceinline__ void aux_seek(const GemmGroup& G, EpiWarp& w, int num_units) {     // first position >= the current one that has work
  const GemmArgs& g = G.g[0];
  while (w.c_unit < num_units) {
    const UnitInfo u = locate_unit<BN, false>(G, w.c_unit);
    w.c_mb = u.mb; w.c_nb = u.nb;
    const bool rows_ok = (u.mb * GEMM_BM + epi_q() * 32) < g.M;
    if (rows_ok && w.c_item < BN / 32 && (u.nb * BN + w.c_item * 32) < g.N) return;
    w.c_unit += static_cast<int>(gridDim.x); w.c_item = epi_cs(); w.c_half = 0;
  }
}
template <int BN>
__device__ __forceinline__ void aux_issue_and_advance(const GemmGroup& G, EpiWarp& w, int num_units, int halves) {
  if (w.c_unit >= num_units) return;
  if (epi_lane() == 0) {
    const int per = halves == 2 ? 16 : 32;                           // columns per load unit
    mbar_arrive_expect_tx_s(w.s_bar, GEMM_STG_BYTES);
    tma_load_2d_s(w.s_in, &G.tmAux, w.s_bar, w.c_nb * BN + w.c_item * 32 + w.c_half * per, w.c_mb * GEMM_BM + epi_q() * 32);
  }
  if (++w.c_half < halves) return;
  w.c_half = 0; w.c_item += 4;
  aux_seek<BN>(G, w, num_units);
}
// wait for the load unit in flight, pull it into registers, start the next one
template <int BN>
__device__ __forceinline__ void aux_take(const GemmGroup& G, EpiWarp& w, int num_units, int halves, uint4 (&u)[4]) {
  mbar_wait_s(w.s_bar, w.aux_phase, 0x600);
  w.aux_phase ^= 1;
  stage_get_s(w.s_in, epi_lane(), u);
  __syncwarp();                                                      // every lane has read the tile: it may be overwritten
  aux_issue_and_advance<BN>(G, w, num_units, halves);
}
// results: make the warp's st.shared visible to the async proxy, then one lane hands the tile to the TMA engine
__device__ __forceinline__ void out_store(EpiWarp& w, const CUtensorMap* map, int col, int row, bool reduce_add = false) {
  fence_proxy_async_smem();
  __syncwarp();
  if (epi_lane() == 0) {
    if (reduce_add) tma_reduce_add_2d_s(map, w.s_out, col, row);
    else            tma_store_2d_s(map, w.s_out, col, row);
    bulk_commit();
  }
  w.store_inflight = 1;
}
__device__ __forceinline__ void out_acquire(EpiWarp& w) {            // before s_out is written again
  if (w.store_inflight) {
    if (epi_lane() == 0) bulk_wait_read0();
    __syncwarp();
    w.store_inflight = 0;
  }
}

// One epilogue work item: the 32 x 32 accumulator block (rows m0.., columns n0..) held one row per thread in v[].
template <int BN, int EPI, int ACT>
__device__ __forceinline__ void epilogue_item(const GemmGroup& G, const GemmArgs& g, const CUtensorMap* tm_out, EpiWarp& w, int num_units,
                                              float (&v)[32], int m0, int n0, float mul, bool accum, float& red, uint8_t* p_out) {
  const int lane = epi_lane();
  const int cc = lane & 3;
  bool stored_halves = false;
  if constexpr (EPI == EPI_STORE) {
    if (g.red_out != nullptr) {          // red += sum(acc * aux): d(alpha_*) as a dot of the un-gated accumulator with a bf16 tile
      uint4 u[4];
      aux_take<BN>(G, w, num_units, 1, u);
      float f[32];
      unpack32(u, f);
      red += dot32(v, f);                  // OOB rows/columns are zero-filled by the TMA load
    }
    scale32(v, mul);
    if (g.col_bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (n0 + j < g.N) v[j] += __ldg(g.col_bias + n0 + j);
    }
  } else if constexpr (EPI == EPI_ACT) {
    if (g.out2 != nullptr) {     // training: also emit act'(acc) so the backward epilogue is two multiplies
      float d[32];
      act32<ACT, true>(v, d);              // v <- act(v), d <- act'(v)
      uint4 u[4];
      pack32(d, u);
      out_acquire(w);
      stage_put_s(w.s_out, lane, u);
      out_store(w, &G.tmOut2, n0, m0);
    } else {
      act32<ACT, false>(v, v);
    }
  } else if constexpr (EPI == EPI_RESID) {
    if (g.aux_f32) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 u[4];
        aux_take<BN>(G, w, num_units, 2, u);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 t = *reinterpret_cast<const float4*>(&u[c]);
          const int j = h * 16 + c * 4;
#if FM_EPI_F32X2
          const float2 lo = fma2(f2(mul), make_float2(v[j], v[j + 1]), make_float2(t.x, t.y));
          const float2 hi = fma2(f2(mul), make_float2(v[j + 2], v[j + 3]), make_float2(t.z, t.w));
          v[j] = lo.x; v[j + 1] = lo.y; v[j + 2] = hi.x; v[j + 3] = hi.y;
#else
          v[j] = fmaf(mul, v[j], t.x); v[j + 1] = fmaf(mul, v[j + 1], t.y);
          v[j + 2] = fmaf(mul, v[j + 2], t.z); v[j + 3] = fmaf(mul, v[j + 3], t.w);
#endif
        }
        if (g.out_f32) {                   // this half is final: 16 fp32 columns = one 64-byte tile row
          uint4 o[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 t = make_float4(v[h * 16 + c * 4], v[h * 16 + c * 4 + 1], v[h * 16 + c * 4 + 2], v[h * 16 + c * 4 + 3]);
            o[c] = *reinterpret_cast<const uint4*>(&t);
          }
          out_acquire(w);
          stage_put_s(w.s_out, lane, o);
          out_store(w, tm_out, n0 + h * 16, m0);
        }
      }
      stored_halves = g.out_f32 != 0;
    } else {
      uint4 u[4];
      aux_take<BN>(G, w, num_units, 1, u);
      float r[32];
      unpack32(u, r);
      axpy32(v, mul, r);
    }
  } else {  // EPI_DACT: out = mul * acc * act'(pre) with act'(pre) saved by the forward epilogue
    uint4 u[4];
    aux_take<BN>(G, w, num_units, 1, u);
    float d[32];
    unpack32(u, d);
    scale_mul32(v, mul, d);            // (d(alpha_ffw) comes from the dW2 GEMM's STORE epilogue: sum(W2 * dW2_ungated))
  }
  if (stored_halves) return;

  // ---- store
  if (g.out_f32 && (EPI == EPI_STORE || EPI == EPI_RESID)) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 u[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 t = make_float4(v[h * 16 + c * 4], v[h * 16 + c * 4 + 1], v[h * 16 + c * 4 + 2], v[h * 16 + c * 4 + 3]);
        u[c] = *reinterpret_cast<const uint4*>(&t);
      }
      out_acquire(w);
      if (g.splits > 1 && !g.par_split) {  // serial split-K: every split goes through ordinary accesses (the fold is a read-modify-write
        stage_put(p_out, lane, u);         // ordered by a global flag; mixing it with async-proxy stores would need cross-proxy fences)
        __syncwarp();
        const bool ok = (n0 + h * 16 + cc * 4) < g.N;
        if (accum) stage_flush<true>(p_out, g.out, static_cast<size_t>(g.ldo) * 4, m0, g.M, static_cast<size_t>(n0 + h * 16) * 4, ok, lane);
        else       stage_flush<false>(p_out, g.out, static_cast<size_t>(g.ldo) * 4, m0, g.M, static_cast<size_t>(n0 + h * 16) * 4, ok, lane);
        __syncwarp();
      } else {
        stage_put_s(w.s_out, lane, u);
        out_store(w, tm_out, n0 + h * 16, m0, g.par_split != 0);      // parallel split-K: fp32 reduce-add into the zeroed output
      }
    }
  } else {
    uint4 u[4];
    pack32(v, u);
    out_acquire(w);
    stage_put_s(w.s_out, lane, u);
    out_store(w, tm_out, n0, m0);
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmGroup G) {
  using Cfg = GemmCfg<BN>;
  constexpr int BM = GEMM_BM, BK = GEMM_BK;
  constexpr int A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES;
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN must be a multiple of 64 in [64,256]");
  constexpr bool GROUPED = (EPI == EPI_STORE);
  const int STAGES = G.stages;

  FM_DYN_SMEM(uint8_t, smem_raw);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* stage_base = sB + STAGES * B_BYTES;                       // 16 x tiles x 2 KB staging tiles (1 KB aligned)
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + GEMM_EPI_WARPS * GEMM_STG_BYTES * G.tiles);
  uint64_t* empty = full + GEMM_MAX_STAGES;
  uint64_t* tfull = empty + GEMM_MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* auxb = tempty + 2;                                       // one per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(auxb + GEMM_EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_units = G.unit_start[G.nprob];

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < G.nprob; ++i) { tma_prefetch_desc(&G.tmA[i]); tma_prefetch_desc(&G.tmB[i]); tma_prefetch_desc(&G.tmOut[i]); }
    if (G.tiles > 1) tma_prefetch_desc(&G.tmAux);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], GEMM_EPI_WARPS); }
    for (int i = 0; i < GEMM_EPI_WARPS; ++i) mbar_init(&auxb[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();            // everything above touched only shared memory / TMEM / kernel parameters
  long long* trace = G.g[0].trace ? G.g[0].trace + static_cast<size_t>(blockIdx.x) * 64 : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = clock64();

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const UnitInfo u = locate_unit<BN, GROUPED>(G, unit);
        const CUtensorMap* tmA = GROUPED ? &G.tmA[u.p] : &G.tmA[0];
        const CUtensorMap* tmB = GROUPED ? &G.tmB[u.p] : &G.tmB[0];
        for (int kb = u.kb_begin; kb < u.kb_end; ++kb) {
          // Cold operands: the ring holds ~190 KB per SM and an HBM round trip under load is ~2 us, so a CTA that waits for
          // DRAM on every stage streams ~50 B/cycle whatever the tile shape (round-2 measurement on the dW GEMMs).  Asking the
          // L2 for the boxes of a k-block well ahead of the ring turns the ring's own loads into L2 hits.
          if (G.l2_ahead > 0 && kb + G.l2_ahead < u.kb_end) {
            const int kp = (kb + G.l2_ahead) * BK;
            if constexpr (!A_MN) {
              tma_prefetch_l2_2d(tmA, kp, u.mb * BM);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) if (u.mb * BM + c * 64 < (GROUPED ? G.g[u.p].M : G.g[0].M)) tma_prefetch_l2_2d(tmA, u.mb * BM + c * 64, kp);
            }
            if constexpr (!B_MN) {
              tma_prefetch_l2_2d(tmB, kp, u.nb * BN);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c) if (u.nb * BN + c * 64 < (GROUPED ? G.g[u.p].N : G.g[0].N)) tma_prefetch_l2_2d(tmB, u.nb * BN + c * 64, kp);
            }
          }
          mbar_wait(&empty[stage], phase ^ 1, 0x100 + stage);
          mbar_arrive_expect_tx(&full[stage], A_BYTES + B_BYTES);
          uint8_t* a = sA + stage * A_BYTES;
          uint8_t* b = sB + stage * B_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(a, tmA, &full[stage], kb * BK, u.mb * BM);            // box 64(k) x 128(m)
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)                                   // box 64(m) x 64(k), 8 KB each
              tma_load_2d(a + c * 8192, tmA, &full[stage], u.mb * BM + c * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(b, tmB, &full[stage], kb * BK, u.nb * BN);            // box 64(k) x BN(n)
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(b + c * 8192, tmB, &full[stage], u.nb * BN + c * 64, kb * BK);
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
        const UnitInfo u = locate_unit<BN, GROUPED>(G, unit);
        mbar_wait(&tempty[acc], acc_phase ^ 1, 0x200 + acc);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = u.kb_begin; kb < u.kb_end; ++kb) {
          mbar_wait(&full[stage], phase, 0x300 + stage);
          tc_fence_after_sync();
          if (trace && kb == u.kb_begin) { const int ui = (unit - blockIdx.x) / gridDim.x; if (ui < 15) trace[1 + 4 * ui] = clock64(); }
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
            umma_bf16(d_tmem, a_desc, b_desc, idesc, (kb > u.kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);          // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);              // accumulator complete -> epilogue
        if (trace) { const int ui = (unit - blockIdx.x) / gridDim.x; if (ui < 15) trace[2 + 4 * ui] = clock64(); }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue (16 warps)
    // Warp (q, cs): TMEM lane quarter q = warp % 4 (rows 32q..32q+31 of the tile); column set cs = (warp-4)/4 takes the
    // 32-column items cs, cs+4, ...  Each thread owns one accumulator row of the item.
    EpiWarp w;
    uint8_t* p_out = stage_base + (warp - 4) * GEMM_STG_BYTES * G.tiles;
    w.s_out = smem_u32(p_out); w.s_in = w.s_out + GEMM_STG_BYTES;     // s_in only meaningful when G.tiles == 2
    w.s_bar = smem_u32(&auxb[warp - 4]); w.aux_phase = 0; w.store_inflight = 0;
    w.c_unit = num_units; w.c_item = epi_cs(); w.c_half = 0; w.c_mb = 0; w.c_nb = 0;
    // which epilogues consume TMA-loaded inputs, and in how many 64-byte units per item
    const bool has_aux = (EPI == EPI_RESID || EPI == EPI_DACT || (EPI == EPI_STORE && G.g[0].red_out != nullptr)) && G.tiles > 1;
    const int halves = (EPI == EPI_RESID && G.g[0].aux_f32) ? 2 : 1;
    if (has_aux) {                                         // the first input tile is in flight before any accumulator exists
      w.c_unit = blockIdx.x;
      aux_seek<BN>(G, w, num_units);
      aux_issue_and_advance<BN>(G, w, num_units, halves);
    }
    const int q = epi_q(), cs = epi_cs();
    int acc = 0; uint32_t acc_phase = 0;
    int cur_p = -1;
    float mul = 1.0f;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const UnitInfo u = locate_unit<BN, GROUPED>(G, unit);
      const GemmArgs& g = GROUPED ? G.g[u.p] : G.g[0];
      const CUtensorMap* tm_out = GROUPED ? &G.tmOut[u.p] : &G.tmOut[0];
      if (u.p != cur_p) {                            // per-problem scale / gate
        cur_p = u.p;
        mul = g.scale;
        if (g.gate != nullptr) mul *= tanhf(__ldg(g.gate));
      }
      mbar_wait(&tfull[acc], acc_phase, 0x400 + acc);
      tc_fence_after_sync();
      if (trace && warp == 4 && lane == 0) { const int ui = (unit - blockIdx.x) / gridDim.x; if (ui < 15) trace[3 + 4 * ui] = clock64(); }
      const int m0 = u.mb * BM + q * 32;
      int* myflag = nullptr;
      const bool serial_split = u.splits > 1 && !g.par_split;
      if (serial_split) {
        // Serial (deterministic) split-K: warp position w of split s adds onto what the same warp position of split
        // s-1 left in `out`; one flag per (tile, epilogue warp) holds the number of splits already folded in.
        myflag = g.flags + u.tile * GEMM_EPI_WARPS + (warp - 4);
        if (u.split > 0) {
          if (lane == 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(myflag) != u.split) {
              if (clock64() - t0 > 4000000000LL) { atomicExch(&g_fm_device_error, 0x80000500u); __trap(); }
            }
          }
          __syncwarp();
        }
      }
      const bool accum = serial_split && u.split > 0;
      float red = 0.0f;
      if (m0 < g.M) {                                // (warp-uniform) this lane quarter has rows inside the matrix
#pragma unroll 1
        for (int item = cs; item < BN / 32; item += 4) {
          const int col_in_tile = item * 32;
          const int n0 = u.nb * BN + col_in_tile;
          if (n0 >= g.N) break;                        // warp-uniform
          float v[32];
          {
            uint32_t r0[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + col_in_tile), r0);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
          }
          if constexpr (EPI == EPI_ACT || EPI == EPI_DACT) {
            if (g.act == 0)      epilogue_item<BN, EPI, 0>(G, g, tm_out, w, num_units, v, m0, n0, mul, accum, red, p_out);
            else if (g.act == 1) epilogue_item<BN, EPI, 1>(G, g, tm_out, w, num_units, v, m0, n0, mul, accum, red, p_out);
            else                 epilogue_item<BN, EPI, 2>(G, g, tm_out, w, num_units, v, m0, n0, mul, accum, red, p_out);
          } else {
            epilogue_item<BN, EPI, 0>(G, g, tm_out, w, num_units, v, m0, n0, mul, accum, red, p_out);
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (trace && warp == 4 && lane == 0) { const int ui = (unit - blockIdx.x) / gridDim.x; if (ui < 15) trace[4 + 4 * ui] = clock64(); }
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if constexpr (EPI == EPI_DACT || EPI == EPI_STORE) {       // per-unit flush: problems of a group have their own red_out
        if (g.red_out != nullptr) {
          red = warp_sum(red);
          if (lane == 0) atomicAdd(g.red_out, red);
        }
      }
      if (serial_split) {                            // publish this warp's part of the tile (last split re-arms the flag)
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          st_release_gpu(myflag, u.split == u.splits - 1 ? 0 : u.split + 1);
        }
      }
    }
    if (lane == 0) bulk_wait0();                     // every output tile of this warp has been written
  }

  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (trace && threadIdx.x == 0) trace[63] = clock64();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace fm
