// Self-test of the async-hardware model (tc_emu.h): the emulator must CATCH protocol mistakes, otherwise a green run of
// the library on it means little.  Each case is one tiny kernel built from the same fm:: wrappers the library uses.
//   ok            TMA -> wait -> tcgen05.mma -> commit -> wait -> tcgen05.ld gives A B^T            (exit 0)
//   no_tma_wait   shared memory read without waiting for the TMA barrier sees 0xCD fill, not data    (exit 0, lazy completion)
//   no_mma_wait   tcgen05.ld without waiting for the commit barrier sees NaN, not the product        (exit 0, lazy completion)
//   wrong_lanes   warp 1 reads TMEM lanes 0..31                                                      (abort, message)
//   tx_mismatch   expect_tx larger than the bytes the TMA delivers -> dead-lock report with the tag  (abort, message)
//   misaligned    swizzled TMA destination that is not 1024-byte aligned                             (abort, message)
//   unallocated   tcgen05.mma into TMEM columns that were never allocated                            (abort, message)
// Driven by tests/test_emu_selftest_cpu.py.  TEST INFRASTRUCTURE ONLY.
#define FM_HOST_EMU 1
#include "simt_emu.h"

#include "../../flamingo_mini_b200/csrc/ptx.cuh"

#include <cstdio>
#include <string>
#include <vector>

using namespace fm;
using bf16 = __nv_bfloat16;

static std::string g_case;
static int g_bad = 0;

// A: [128][64] bf16 K-major, B: [64][64] bf16 K-major; D = A B^T (128 x 64)
static void mini_kernel(CUtensorMap tmA, CUtensorMap tmB, float* out) {
  FM_DYN_SMEM(uint8_t, raw);
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = sm;
  uint8_t* sB = sm + 16384;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sB + 8192);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(bar_load, 1); mbar_init(bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  __syncthreads();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_load, 16384 + 8192 + (g_case == "tx_mismatch" ? 16 : 0));
    tma_load_2d(g_case == "misaligned" ? sA + 128 : sA, &tmA, bar_load, 0, 0);
    tma_load_2d(sB, &tmB, bar_load, 0, 0);
  }
  if (g_case == "no_tma_wait") {
    if (tid == 1) { for (int i = 0; i < 16384; ++i) if (sA[i] != 0xCD) { ++g_bad; break; } }     // nothing may have landed yet
    __syncthreads();
  }
  mbar_wait(bar_load, 0, 0x777);
  if (tid == 0) {
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem + (g_case == "unallocated" ? 64 : 0), umma_smem_desc_sw128(smem_u32(sA) + k * 32, 0, 1024),
                umma_smem_desc_sw128(smem_u32(sB) + k * 32, 0, 1024), umma_idesc_bf16(128, 64, false, false), k > 0);
    umma_commit(bar_mma);
  }
  uint32_t r[32];
  if (g_case == "no_mma_wait") {
    __syncthreads();                              // the MMAs are issued and committed, but nobody has waited for them
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), r);
    float f;
    memcpy(&f, &r[0], 4);
    if (f == f) ++g_bad;                          // must still be the NaN fill
    __syncthreads();
  }
  mbar_wait(bar_mma, 0, 0x778);
  const int q = (g_case == "wrong_lanes" && warp == 1) ? 0 : warp;
  for (int h = 0; h < 2; ++h) {
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + h * 32, r);
    tmem_ld_wait();
    memcpy(out + (warp * 32 + (tid & 31)) * 64 + h * 32, r, 128);
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main(int argc, char** argv) {
  g_case = argc > 1 ? argv[1] : "ok";
  std::vector<bf16> A(128 * 64), B(64 * 64);
  for (size_t i = 0; i < A.size(); ++i) A[i] = __float2bfloat16(static_cast<float>((i * 37 % 101)) / 50.0f - 1.0f);
  for (size_t i = 0; i < B.size(); ++i) B[i] = __float2bfloat16(static_cast<float>((i * 53 % 89)) / 40.0f - 1.0f);
  CUtensorMap tmA, tmB;
  cuuint64_t dA[2] = {64, 128}, dB[2] = {64, 64}, st[1] = {128};
  cuuint32_t bA[2] = {64, 128}, bB[2] = {64, 64}, es[2] = {1, 1};
  if (emu::encode_tiled(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A.data(), dA, st, bA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
      emu::encode_tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B.data(), dB, st, bB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("encode failed\n");
    return 2;
  }
  std::vector<float> out(128 * 64, -1.0f);
  emu::launch(1, 128, 16384 + 8192 + 1024 + 64, [&] { mini_kernel(tmA, tmB, out.data()); });
  double worst = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      double want = 0;
      for (int k = 0; k < 64; ++k) want += static_cast<double>(__bfloat162float(A[m * 64 + k])) * __bfloat162float(B[n * 64 + k]);
      worst = std::max(worst, std::fabs(out[m * 64 + n] - want));
    }
  if (worst > 1e-3 || g_bad) { printf("SELFTEST %s FAILED: max abs err %g, lazy-completion violations %d\n", g_case.c_str(), worst, g_bad); return 1; }
  printf("SELFTEST %s OK\n", g_case.c_str());
  return 0;
}
