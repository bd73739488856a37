// libflamingo_b200.so — host side + C ABI (include/flamingo_b200.h).  Single translation unit:
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC flamingo_b200.cu
// The host code only sequences kernels on the caller's stream; it owns no device memory.
#include "../../include/flamingo_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "attn_tc.cuh"
#include "gemm_tc.cuh"
#include "layernorm.cuh"
#include "loss.cuh"
#include "misc.cuh"

using namespace fm;
typedef __nv_bfloat16 bf16;

// ================================================================================================ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU_TRY(expr)                                                                                   \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess) return fail(FM_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define FM_TRY(expr)            \
  do {                          \
    int r__ = (expr);           \
    if (r__ != FM_OK) return r__; \
  } while (0)
#define KERNEL_CHECK() CU_TRY(cudaGetLastError())

extern "C" int fm_version(void) { return 1; }
extern "C" const char* fm_last_error(void) { return g_err; }
extern "C" unsigned int fm_device_error(void) {
  unsigned int v = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&v, g_fm_device_error, sizeof(v));
  return v;
}
extern "C" int fm_abi_sizes(int* out5) {
  out5[0] = (int)sizeof(fm_gemm_desc);
  out5[1] = (int)sizeof(fm_xattn_cfg);
  out5[2] = (int)sizeof(fm_xattn_layout);
  out5[3] = (int)sizeof(fm_resampler_cfg);
  out5[4] = (int)sizeof(fm_resampler_layout);
  return FM_OK;
}


// ================================================================================================ launch counter / profiler
// fm_launch_count(): number of kernels this library has launched (the bench's "gpu_launches" claim).
// fm_profile_enable(1): from now on every launch is bracketed by CUDA events recorded on the launching stream;
// fm_profile_report() synchronises and returns one line per kernel tag: "tag launches total_ms flops bytes".
#include <atomic>
#include <map>
#include <string>
#include <vector>
static std::atomic<unsigned long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
struct ProfRec { std::string tag; cudaEvent_t a, b; double flops, bytes; };     // tag starts with '@' when recorded under capture
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
// Which module-level entry point the launches belong to ("x/" gated xattn block, "r/" resampler, "" raw ops): prefix of the
// profiler tag, so bench.py can report the xattn blocks' TFLOP/s on their own (BASELINE.json metric, second half).
static thread_local const char* g_scope = "";
// Under stream capture the events become EXTERNAL event-record nodes of the graph: every replay re-stamps them, so the
// per-kernel durations are those of the replayed graph itself (no host launch latency inside the intervals, and the sum over
// kernels cannot exceed the replay's duration).  Eager launches record ordinary events.
static inline bool stream_capturing(cudaStream_t s) {
#ifndef FM_HOST_EMU
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(s, &st) == cudaSuccess && st == cudaStreamCaptureStatusActive;
#else
  (void)s;
  return false;
#endif
}
static inline void prof_record(cudaEvent_t e, cudaStream_t s, bool captured) {
#ifndef FM_HOST_EMU
  if (captured) { cudaEventRecordWithFlags(e, s, cudaEventRecordExternal); return; }
#endif
  (void)captured;
  cudaEventRecord(e, s);
}
struct ProfScope {
  cudaStream_t s; bool on; bool captured = false; ProfRec rec;
  ProfScope(const char* tag, double flops, double bytes, cudaStream_t st) : s(st), on(g_prof_on.load() != 0) {
    g_launches.fetch_add(1);
    if (on) {
      captured = stream_capturing(s);
      rec.tag = std::string(captured ? "@" : "") + g_scope + tag; rec.flops = flops; rec.bytes = bytes;
      rec.a = rec.b = nullptr;
      if (g_prof_on.load() == 1) {            // mode 2 = launch log only (tag / flops / bytes in launch order, no events)
        cudaEventCreate(&rec.a); cudaEventCreate(&rec.b);
        prof_record(rec.a, s, captured);
      }
    }
  }
  ~ProfScope() {
    if (on) {
      if (rec.a != nullptr) prof_record(rec.b, s, captured);
      std::lock_guard<std::mutex> lk(g_prof_mu);
      g_prof_recs.push_back(rec);
    }
  }
};
extern "C" unsigned long long fm_launch_count(void) { return g_launches.load(); }
extern "C" int fm_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof_recs) { if (r.a) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } }
  g_prof_recs.clear();
  g_prof_on.store(on);
  return FM_OK;
}
// launch log: one line per launch recorded since fm_profile_enable(1 or 2), in launch order: "tag flops bytes"
extern "C" int fm_profile_log(char* buf, size_t n) {
  if (!buf || n == 0) return fail(FM_EINVAL, "fm_profile_log: no buffer");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  size_t off = 0;
  buf[0] = 0;
  for (auto& r : g_prof_recs) {
    int w = snprintf(buf + off, n - off, "%s %.6e %.6e\n", r.tag.c_str(), r.flops, r.bytes);
    if (w < 0 || (size_t)w >= n - off) return fail(FM_EINVAL, "fm_profile_log: buffer too small for %zu launches", g_prof_recs.size());
    off += (size_t)w;
  }
  return FM_OK;
}
extern "C" int fm_profile_report(char* buf, size_t n) {
  if (!buf || n == 0) return fail(FM_EINVAL, "fm_profile_report: no buffer");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(FM_ECUDA, "cudaDeviceSynchronize failed: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(g_prof_mu);
  struct Agg { long n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_prof_recs) {
    float ms = 0.f;
    if (r.a == nullptr || cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    Agg& a = agg[r.tag];
    a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
  }
  size_t off = 0;
  buf[0] = 0;
  for (auto& kv : agg) {
    int w = snprintf(buf + off, n - off, "%s %ld %.6f %.6e %.6e\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes);
    if (w < 0 || (size_t)w >= n - off) break;
    off += (size_t)w;
  }
  return FM_OK;
}

// ================================================================================================ device info
static int g_num_sms = 0;
static int device_init() {
  static std::once_flag once;
  static int status = FM_OK;
  std::call_once(once, [] {
    int dev = 0, major = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      status = FM_ECUDA;
      return;
    }
    if (major != 10) { status = FM_EUNSUPPORTED; return; }
    g_num_sms = sms;
  });
  if (status == FM_ECUDA) return fail(FM_ECUDA, "no usable CUDA device (libflamingo_b200 has no CPU fallback)");
  if (status == FM_EUNSUPPORTED) return fail(FM_EUNSUPPORTED, "libflamingo_b200 requires an sm_100 (B200) device");
  return FM_OK;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per (function, device): remember which pairs are done (a process may
// drive several devices, e.g. the 2-GPU NCCL test's parent process).
static cudaError_t ensure_dyn_smem(const void* kern, int bytes) {
#ifdef FM_HOST_EMU
  (void)kern; (void)bytes;
  return cudaSuccess;
#else
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, cudaError_t> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  auto it = done.find({kern, dev});
  if (it != done.end()) return it->second;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  done[{kern, dev}] = e;
  return e;
#endif
}

// ================================================================================================ options / launch helper
// Scheduling switches (include/flamingo_b200.h).  None of them changes a result beyond floating-point summation order.
static std::atomic<int> g_opt[FM_OPT_COUNT] = {{1}, {1}, {0}, {1}, {1}, {1}, {0}, {1}, {1}, {0}, {1}};
static inline bool opt(int key) { return g_opt[key].load(std::memory_order_relaxed) != 0; }
extern "C" int fm_set_option(int key, int value) {
  if (key < 0 || key >= FM_OPT_COUNT) return fail(FM_EINVAL, "unknown option %d", key);
  if (key == FM_OPT_SM_RESERVE) {
    if (value < 0 || value > 128) return fail(FM_EINVAL, "FM_OPT_SM_RESERVE must be in [0, 128] (got %d)", value);
    g_opt[key].store(value);
  } else {
    g_opt[key].store(value ? 1 : 0);
  }
  return FM_OK;
}
// SMs a persistent GEMM grid may occupy (FM_OPT_SM_RESERVE leaves room for NCCL's CTAs)
static inline int gemm_sms() {
  const int n = g_num_sms - g_opt[FM_OPT_SM_RESERVE].load(std::memory_order_relaxed);
  return n < 1 ? 1 : n;
}

// Programmatic dependent launch is requested only when the previous operation this thread enqueued on the same stream
// was one of this library's kernels (all of which call griddepcontrol.wait before touching global memory): a kernel
// that follows a memset, an event wait or foreign work keeps the ordinary full dependency.
struct PdlTrack {
  cudaStream_t st[2] = {nullptr, nullptr};
  bool after_kernel[2] = {false, false};
  int slot(cudaStream_t s) {
    if (st[0] == s) return 0;
    if (st[1] == s) return 1;
    const int i = (st[0] == nullptr) ? 0 : 1;       // at most two streams per API call: the caller's and the side stream
    st[i] = s; after_kernel[i] = false;
    return i;
  }
  void reset() { st[0] = st[1] = nullptr; after_kernel[0] = after_kernel[1] = false; }
};
static thread_local PdlTrack g_pdl;
static inline void note_other(cudaStream_t s) { g_pdl.after_kernel[g_pdl.slot(s)] = false; }   // memset / event wait enqueued on s
struct ApiScope {          // every extern "C" entry point that launches kernels starts from "unknown predecessor"
  const char* prev_scope;
  explicit ApiScope(const char* scope = "") : prev_scope(g_scope) { g_scope = scope; g_pdl.reset(); }
  ~ApiScope() { g_scope = prev_scope; g_pdl.reset(); }
};

template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
#ifdef FM_HOST_EMU      // tests/cpu_harness: the kernel is an ordinary host function run thread-per-thread by the emulator
  g_pdl.after_kernel[g_pdl.slot(s)] = true;
  emu::launch(grid, block, smem, [&] { kern(args...); });
  return cudaSuccess;
#else
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  const int sl = g_pdl.slot(s);
  if (opt(FM_OPT_PDL) && g_pdl.after_kernel[sl] && g_prof_on.load() == 0) {     // profiler events sit between kernels
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
  }
  g_pdl.after_kernel[sl] = true;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
#endif
}

// ================================================================================================ TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// 2-D bf16 tensor [outer, inner] with row pitch ld elements; box = box_inner x box_outer, 128B swizzle, zero OOB fill.
static int make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                        uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(FM_ECUDA, "cuTensorMapEncodeTiled entry point not found");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld % 8) != 0)
    return fail(FM_EINVAL, "GEMM operand must be 16-byte aligned with a leading dimension multiple of 8 (ptr=%p ld=%llu)", ptr,
                (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FM_ECUDA, "cuTensorMapEncodeTiled failed with %d (inner=%llu outer=%llu ld=%llu)", (int)r,
                                     (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
  return FM_OK;
}

// epilogue IO map: 2-D bf16 / fp32 tensor [outer = rows, inner = columns], box = 64 bytes x 32 rows, SWIZZLE_64B (the layout of the
// per-warp staging tiles in gemm_tc.cuh); TMA clips stores and zero-fills loads outside the tensor
static int make_tmap_io(CUtensorMap* m, const void* ptr, int f32, uint64_t inner, uint64_t outer, uint64_t ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(FM_ECUDA, "cuTensorMapEncodeTiled entry point not found");
  const uint64_t es = f32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * es) % 16 != 0)
    return fail(FM_EINVAL, "GEMM epilogue tensor must be 16-byte aligned with a row pitch multiple of 16 bytes (ptr=%p ld=%llu)", ptr, (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * es};
  cuuint32_t box[2] = {(cuuint32_t)(64 / es), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FM_ECUDA, "cuTensorMapEncodeTiled(epilogue map) failed with %d (inner=%llu outer=%llu ld=%llu)", (int)r,
                                     (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
  return FM_OK;
}

// ================================================================================================ GEMM launch
// Serial split-K cuts K into ceil(num_kb / splits)-block ranges; with an unlucky (K, splits) pair the last ranges would be
// empty (K = 240 -> 4 blocks, splits = 3 -> ranges of 2: the third split has nothing to add and would fold an unwritten
// accumulator).  Clamp to the number of NON-empty ranges.  (Found by tools/emu_fuzz.py --kind gemm on the host emulator.)
static int effective_splits(int K, int splits) {
  if (splits <= 1) return 1;
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;
  const int per = (num_kb + splits - 1) / splits;
  return (num_kb + per - 1) / per;
}
template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm_inst(const fm_gemm_desc* ds, int nprob, cudaStream_t s) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, EPI>;
  const cudaError_t attr_err = ensure_dyn_smem(reinterpret_cast<const void*>(kern), GEMM_SMEM_LIMIT);
  if (attr_err != cudaSuccess) return fail(FM_ECUDA, "cudaFuncSetAttribute(smem=%d) failed: %s", GEMM_SMEM_LIMIT, cudaGetErrorString(attr_err));
  if (nprob < 1 || nprob > GEMM_MAX_GROUP || (nprob > 1 && EPI != EPI_STORE))
    return fail(FM_EINVAL, "GEMM group of %d problems (max %d, STORE epilogue only)", nprob, GEMM_MAX_GROUP);
  GemmGroup G;
  memset(&G, 0, sizeof(G));
  G.nprob = nprob;
  double flops = 0.0, bytes = 0.0;
  int units = 0;
  for (int i = 0; i < nprob; ++i) {
    const fm_gemm_desc& d = ds[i];
    if (!A_MN) FM_TRY(make_tmap_2d(&G.tmA[i], d.A, d.K, d.M, d.lda, GEMM_BK, GEMM_BM));
    else       FM_TRY(make_tmap_2d(&G.tmA[i], d.A, d.M, d.K, d.lda, 64, GEMM_BK));
    if (!B_MN) FM_TRY(make_tmap_2d(&G.tmB[i], d.B, d.K, d.N, d.ldb, GEMM_BK, BN));
    else       FM_TRY(make_tmap_2d(&G.tmB[i], d.B, d.N, d.K, d.ldb, 64, GEMM_BK));
    GemmArgs& g = G.g[i];
    g.M = d.M; g.N = d.N; g.K = d.K;
    g.out = d.out; g.ldo = d.ldo; g.out2 = d.out2; g.ldo2 = d.ldo2; g.aux = d.aux; g.ldaux = d.ldaux;
    g.col_bias = d.col_bias; g.gate = d.gate; g.red_out = d.red_out; g.scale = d.scale; g.act = d.act;
    g.out_f32 = d.out_f32; g.aux_f32 = d.aux_f32;
    // d.splits < 0: PARALLEL split-K over |splits| K ranges (the caller has zeroed `out`; fp32 STORE only; also inside groups)
    g.par_split = (d.splits < 0 && d.out_f32 && EPI == EPI_STORE) ? 1 : 0;
    g.splits = g.par_split ? effective_splits(d.K, -d.splits) : (nprob == 1 ? effective_splits(d.K, d.splits) : 1);
    if (g.splits <= 1) g.par_split = 0;
    g.flags = d.splitk_flags; g.trace = d.trace;
    FM_TRY(make_tmap_io(&G.tmOut[i], d.out, d.out_f32, d.N, d.M, d.ldo));
    G.unit_start[i] = units;
    units += ((d.M + GEMM_BM - 1) / GEMM_BM) * ((d.N + BN - 1) / BN) * g.splits;
    flops += 2.0 * d.M * d.N * d.K;
    bytes += 2.0 * ((double)d.M * d.K + (double)d.N * d.K + (double)d.M * d.N);
  }
  G.unit_start[nprob] = units;
  // epilogue inputs (problem 0 only: groups are plain STORE problems) and the second ACT output
  G.tmAux = G.tmOut[0]; G.tmOut2 = G.tmOut[0];      // placeholders (never dereferenced unless built below)
  G.tiles = 1;
  {
    const fm_gemm_desc& d = ds[0];
    if (EPI == EPI_RESID || EPI == EPI_DACT || (EPI == EPI_STORE && d.red_out && d.aux)) {
      FM_TRY(make_tmap_io(&G.tmAux, d.aux, (EPI == EPI_RESID) ? d.aux_f32 : 0, d.N, d.M, d.ldaux));
      G.tiles = 2;
    }
    if (EPI == EPI_ACT && d.out2) FM_TRY(make_tmap_io(&G.tmOut2, d.out2, 0, d.N, d.M, d.ldo2));
  }
  G.stages = Cfg::stages_for(G.tiles);
  // operand L2 prefetch distance (in 64-deep k-blocks) for long contractions; FM_OPT_EPI_PREFETCH (historic name) switches it.
  // OFF by default: measured on B200 (profiles/r02_call11_final) the extra TMA prefetch operations slow the producer down - the
  // K = 3072 main loops took 31.7k instead of 21.8k cycles on L2-resident operands, and C3/C4/C5 steps were 3-4 % longer.
  G.l2_ahead = 0;
  if (opt(FM_OPT_EPI_PREFETCH)) {
    int max_kb = 0;
    for (int i = 0; i < nprob; ++i) { const int kb = (ds[i].K + GEMM_BK - 1) / GEMM_BK; if (kb > max_kb) max_kb = kb; }
    if (max_kb >= 16) G.l2_ahead = 2 * G.stages;
  }
  if (G.stages < 2) return fail(FM_EINVAL, "GEMM tile width %d leaves no room for a 2-stage operand ring", BN);
  const int smem_bytes = Cfg::smem_bytes(G.stages, G.tiles);
  const int grid = units < gemm_sms() ? units : gemm_sms();
  {
    char tag[64];
    snprintf(tag, sizeof(tag), nprob > 1 ? "gemm_a%db%d_epi%d_bn%d_group" : "gemm_a%db%d_epi%d_bn%d", (int)A_MN, (int)B_MN, EPI, BN);
    ProfScope ps(tag, flops, bytes, s);
#ifdef FM_HOST_EMU
    emu::concurrent_next = true;       // CTAs of a split-K launch wait for each other: all of them must be resident
#endif
    (void)launch_k(kern, grid, GEMM_THREADS, (size_t)smem_bytes, s, G);
  }
  KERNEL_CHECK();
  return FM_OK;
}

static int pick_bn(int M, int N) {
  // maximise (wave efficiency) x (per-tile MMA efficiency: wider tiles move fewer smem bytes per FLOP)
  const int cands[4] = {256, 192, 128, 64};
  const double tile_eff[4] = {1.0, 0.95, 0.85, 0.6};
  const int mb = (M + GEMM_BM - 1) / GEMM_BM;
  double best = -1.0;
  int best_bn = 128;
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    const int nb = (N + bn - 1) / bn;
    const double fill = (double)N / ((double)nb * bn);               // wasted columns of the last tile
    const long tiles = (long)mb * nb;
    const long waves = (tiles + gemm_sms() - 1) / gemm_sms();
    const double wave_eff = (double)tiles / ((double)waves * gemm_sms());
    const double score = wave_eff * tile_eff[i] * fill;
    if (score > best + 1e-9) { best = score; best_bn = bn; }
  }
  return best_bn;
}

template <bool A_MN, bool B_MN, int EPI>
static int launch_gemm_bn(const fm_gemm_desc* d, int n, int bn, cudaStream_t s) {
  switch (bn) {
    case 64:  return launch_gemm_inst<64, A_MN, B_MN, EPI>(d, n, s);
    case 128: return launch_gemm_inst<128, A_MN, B_MN, EPI>(d, n, s);
    case 192: return launch_gemm_inst<192, A_MN, B_MN, EPI>(d, n, s);
    case 256: return launch_gemm_inst<256, A_MN, B_MN, EPI>(d, n, s);
  }
  return fail(FM_EINVAL, "unsupported GEMM tile width %d", bn);
}

static int run_gemm(const fm_gemm_desc& d, cudaStream_t s) {
  FM_TRY(device_init());
  if (d.M <= 0 || d.N <= 0 || d.K <= 0) return fail(FM_EINVAL, "GEMM with empty dimension M=%d N=%d K=%d", d.M, d.N, d.K);
  if (d.N % 8 != 0 || d.ldo % 8 != 0) return fail(FM_EINVAL, "GEMM N and ldo must be multiples of 8 (N=%d ldo=%lld)", d.N, d.ldo);
  if (!d.A || !d.B || !d.out) return fail(FM_EINVAL, "GEMM null operand");
  if ((d.epi == EPI_RESID || d.epi == EPI_DACT || (d.epi == EPI_STORE && d.red_out)) && (!d.aux || d.ldaux % 8 != 0)) return fail(FM_EINVAL, "GEMM epilogue %d needs aux with ld %% 8 == 0", d.epi);
  if (d.epi == EPI_ACT && d.out2 && d.ldo2 % 8 != 0) return fail(FM_EINVAL, "GEMM ldo2 must be a multiple of 8");
  if (d.epi == EPI_DACT && d.red_out) return fail(FM_EINVAL, "GEMM DACT epilogue has no reduction output (d(alpha_ffw) comes from the dW2 GEMM: STORE epilogue with aux = W2, red_out)");
  fm_gemm_desc dd = d;
  int bn = d.bn;
  const bool can_split = d.epi == EPI_STORE && d.out_f32 && d.splitk_flags != nullptr;
  const bool par_split = d.epi == EPI_STORE && d.out_f32 && d.splits < 0;
  if (!can_split && !par_split) dd.splits = 1;
  if (can_split && d.splits == 0) {
    // gradient-shaped problem: few output tiles, long K.  Use the widest tile and cut K so the units fill the SMs;
    // the serial fold keeps the chain short (<= 4).
    const int wide = d.bn ? d.bn : (d.N >= 256 ? 256 : d.N >= 192 ? 192 : d.N >= 128 ? 128 : 64);
    const int tiles = ((d.M + GEMM_BM - 1) / GEMM_BM) * ((d.N + wide - 1) / wide);
    const int num_kb = (d.K + GEMM_BK - 1) / GEMM_BK;
    int sp = 1;
    if (false && tiles * 2 <= g_num_sms) {   // measured (gemm_bench): the serial fold chain costs more than it gains at these sizes
      sp = g_num_sms / tiles;
      if (sp > 4) sp = 4;
      if (sp > num_kb / 4) sp = num_kb / 4;
      if (sp < 1) sp = 1;
    }
    dd.splits = sp;
    if (sp > 1) bn = wide;
  }
  if (bn == 0) bn = pick_bn(d.M, d.N);
  const int key = (d.a_mn ? 2 : 0) | (d.b_mn ? 1 : 0);
  if (key == 0) {
    if (d.epi == EPI_STORE) return launch_gemm_bn<false, false, EPI_STORE>(&dd, 1, bn, s);
    if (d.epi == EPI_ACT) return launch_gemm_bn<false, false, EPI_ACT>(&dd, 1, bn, s);
    if (d.epi == EPI_RESID) return launch_gemm_bn<false, false, EPI_RESID>(&dd, 1, bn, s);
  } else if (key == 1) {
    if (d.epi == EPI_STORE) return launch_gemm_bn<false, true, EPI_STORE>(&dd, 1, bn, s);
    if (d.epi == EPI_DACT) return launch_gemm_bn<false, true, EPI_DACT>(&dd, 1, bn, s);
  } else if (key == 3) {
    if (d.epi == EPI_STORE) return launch_gemm_bn<true, true, EPI_STORE>(&dd, 1, bn, s);
  }
  return fail(FM_EINVAL, "GEMM variant not built: a_mn=%d b_mn=%d epi=%d", d.a_mn, d.b_mn, d.epi);
}
// ---- grouped launches: up to GEMM_MAX_GROUP independent STORE problems with the same operand layouts in ONE persistent launch.
// Tile width: the candidate with the smallest modelled makespan under the kernel's static schedule (CTA j takes units
// j, j+grid, ...).  Per 64-deep k-block a CTA is bound by max(tensor pipe, bytes in flight per SM): r01_engineering_log.md #7.
static double group_makespan(const fm_gemm_desc* ds, int n, int bn) {
  const double kb_us = fmax(0.107 * bn / 64.0, (16.0 + 8.0 * bn / 64.0) / 192.0);
  const double epi_us = 0.5 * bn / 64.0;
  static thread_local std::vector<double> load;
  const int sms = gemm_sms();
  load.assign((size_t)sms, 0.0);
  long unit = 0;
  double worst = 0.0;
  for (int i = 0; i < n; ++i) {
    const long tiles = (long)((ds[i].M + GEMM_BM - 1) / GEMM_BM) * ((ds[i].N + bn - 1) / bn);
    const double c = kb_us * ((ds[i].K + GEMM_BK - 1) / GEMM_BK);
    for (long t = 0; t < tiles; ++t, ++unit) {
      double& l = load[(size_t)(unit % sms)];
      l += c;
      if (l + epi_us > worst) worst = l + epi_us;     // a CTA's last epilogue is never hidden
    }
  }
  return worst;
}
static int pick_bn_group(const fm_gemm_desc* ds, int n) {
  const int cands[4] = {256, 192, 128, 64};
  double best = 1e30;
  int best_bn = 64;
  for (int i = 0; i < 4; ++i) {
    const double t = group_makespan(ds, n, cands[i]);
    if (t < best - 1e-9) { best = t; best_bn = cands[i]; }
  }
  return best_bn;
}
static int run_gemm_group(const fm_gemm_desc* ds, int n, cudaStream_t s) {
  FM_TRY(device_init());
  if (!ds || n < 1 || n > GEMM_MAX_GROUP) return fail(FM_EINVAL, "GEMM group needs 1..%d problems (got %d)", GEMM_MAX_GROUP, n);
  for (int i = 0; i < n; ++i) {
    const fm_gemm_desc& d = ds[i];
    if (d.epi != EPI_STORE || d.a_mn != ds[0].a_mn || d.b_mn != ds[0].b_mn)
      return fail(FM_EINVAL, "GEMM group: problem %d must use the STORE epilogue and the layouts of problem 0", i);
    if (d.red_out) return fail(FM_EINVAL, "GEMM group: problem %d carries a reduction output (single launches only)", i);
  }
  if (n == 1 || !opt(FM_OPT_GEMM_GROUP)) {
    for (int i = 0; i < n; ++i) FM_TRY(run_gemm(ds[i], s));
    return FM_OK;
  }
  fm_gemm_desc dd[GEMM_MAX_GROUP];
  for (int i = 0; i < n; ++i) {
    const fm_gemm_desc& d = ds[i];
    if (d.M <= 0 || d.N <= 0 || d.K <= 0) return fail(FM_EINVAL, "GEMM group: problem %d has an empty dimension", i);
    if (d.N % 8 != 0 || d.ldo % 8 != 0 || !d.A || !d.B || !d.out) return fail(FM_EINVAL, "GEMM group: bad problem %d (N, ldo multiples of 8; non-null operands)", i);
    dd[i] = d;
    if (!(d.splits < 0 && d.out_f32)) dd[i].splits = 1;      // parallel split-K survives inside a group
  }
  const int bn = ds[0].bn ? ds[0].bn : pick_bn_group(dd, n);
  const int key = (ds[0].a_mn ? 2 : 0) | (ds[0].b_mn ? 1 : 0);
  if (key == 0) return launch_gemm_bn<false, false, EPI_STORE>(dd, n, bn, s);
  if (key == 1) return launch_gemm_bn<false, true, EPI_STORE>(dd, n, bn, s);
  if (key == 3) return launch_gemm_bn<true, true, EPI_STORE>(dd, n, bn, s);
  return fail(FM_EINVAL, "GEMM group variant not built: a_mn=%d b_mn=%d", ds[0].a_mn, ds[0].b_mn);
}

// Weight-gradient GEMMs (dW = dY^T X: few output tiles, K = rows of the batch) are bound by L2 -> SM operand bandwidth, and the
// bytes a tile pulls per FLOP fall with its width: plan 256-wide tiles and cut K into as many PARALLEL ranges as it takes to
// give every SM a unit (each range adds its partial tile with a TMA reduce-add into the zeroed gradient).
// Returns the (negative) splits value for fm_gemm_desc and sets *bn; 1 = no split.
static int plan_dw(int M, int N, int K, int* bn, int units_other = 0) {
  *bn = 0;
  if (!opt(FM_OPT_DW_SPLITK)) return 1;
  const int wide = N >= 256 ? 256 : N >= 192 ? 192 : N >= 128 ? 128 : 64;
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + wide - 1) / wide);
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;
  const int room = gemm_sms() - units_other;
  if (tiles * 2 > room || num_kb < 16) return 1;
  int sp = room / tiles;
  if (sp > 8) sp = 8;
  while (sp > 1 && num_kb / sp < 8) --sp;                  // at least 8 k-blocks (512 rows of the batch) per range
  if (sp <= 1) return 1;
  *bn = wide;
  return -sp;
}

// The same for a GROUP of weight-gradient problems sharing one launch: one tile width for all, and per problem as many K ranges
// as make the units of all problems about equally long while their number stays within the SMs.
static bool plan_dw_group(fm_gemm_desc* ds, int n) {
  if (!opt(FM_OPT_DW_SPLITK) || n < 1) return false;
  int wide = 256;
  for (int i = 0; i < n; ++i) {
    const int w = ds[i].N >= 256 ? 256 : ds[i].N >= 192 ? 192 : ds[i].N >= 128 ? 128 : 64;
    if (w < wide) wide = w;
  }
  int tiles[GEMM_MAX_GROUP], kbs[GEMM_MAX_GROUP], total = 0;
  for (int i = 0; i < n; ++i) {
    tiles[i] = ((ds[i].M + GEMM_BM - 1) / GEMM_BM) * ((ds[i].N + wide - 1) / wide);
    kbs[i] = (ds[i].K + GEMM_BK - 1) / GEMM_BK;
    total += tiles[i];
  }
  if (total * 2 > gemm_sms()) return false;
  for (int t = 8; t <= 4096; ++t) {                      // smallest unit length (in k-blocks) whose unit count fits the SMs
    int units = 0;
    for (int i = 0; i < n; ++i) units += tiles[i] * ((kbs[i] + t - 1) / t);
    if (units <= gemm_sms()) {
      bool any = false;
      for (int i = 0; i < n; ++i) {
        const int sp = (kbs[i] + t - 1) / t;
        ds[i].splits = sp > 1 ? -sp : 1;
        ds[i].bn = wide;
        any = any || sp > 1;
      }
      return any;
    }
  }
  return false;
}

extern "C" size_t fm_gemm_splitk_flag_ints(int M, int N) {
  return (size_t)((M + GEMM_BM - 1) / GEMM_BM) * (size_t)((N + 63) / 64) * GEMM_EPI_WARPS;
}
extern "C" int fm_gemm_bf16(const fm_gemm_desc* d, fm_stream_t stream) {
  ApiScope api_scope;
  if (!d) return fail(FM_EINVAL, "null descriptor");
  return run_gemm(*d, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int fm_gemm_bf16_group(const fm_gemm_desc* d, int n, fm_stream_t stream) {
  ApiScope api_scope;
  return run_gemm_group(d, n, reinterpret_cast<cudaStream_t>(stream));
}

// builder for the common cases
static fm_gemm_desc mk_gemm(int M, int N, int K, const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn,
                            int epi, void* out, long long ldo, int out_f32, int* splitk_flags = nullptr) {
  fm_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.splitk_flags = splitk_flags;
  d.M = M; d.N = N; d.K = K; d.A = A; d.lda = lda; d.a_mn = a_mn; d.B = B; d.ldb = ldb; d.b_mn = b_mn;
  d.epi = epi; d.out = out; d.ldo = ldo; d.out_f32 = out_f32; d.scale = 1.0f;
  return d;
}

// ================================================================================================ side stream for dW GEMMs
// Weight-gradient GEMMs are leaves of the backward graph (nothing downstream reads them until the step ends), and the
// small ones (dWq, dWout, dWkv: 48-96 CTAs) cannot fill 148 SMs.  They are therefore issued on a library-owned side
// stream that forks from / joins back into the caller's stream with events, so they overlap the dX / LayerNorm /
// attention chain.  Under CUDA-graph capture the fork/join simply become parallel branches of the graph.
struct SideStream {
  cudaStream_t main = nullptr, side = nullptr;
  cudaEvent_t ev[8];
  int nfork = 0;
  bool ok = false;
  static std::mutex& mu() { static std::mutex m; return m; }
  explicit SideStream(cudaStream_t m) : main(m) {
    if (!opt(FM_OPT_SIDE_STREAM)) return;     // ok stays false: callers fall back to the main stream
    static cudaStream_t s_side = nullptr;
    static cudaEvent_t s_ev[8];
    static bool s_ok = false;
    std::lock_guard<std::mutex> lk(mu());
    if (!s_ok) {
      if (cudaStreamCreateWithFlags(&s_side, cudaStreamNonBlocking) != cudaSuccess) return;
      for (int i = 0; i < 8; ++i)
        if (cudaEventCreateWithFlags(&s_ev[i], cudaEventDisableTiming) != cudaSuccess) return;
      s_ok = true;
    }
    side = s_side;
    for (int i = 0; i < 8; ++i) ev[i] = s_ev[i];
    ok = true;
  }
  // side stream waits for everything enqueued on the main stream so far
  int fork() {
    if (!ok) return FM_OK;
    cudaEvent_t e = ev[nfork % 7];
    ++nfork;
    CU_TRY(cudaEventRecord(e, main));
    CU_TRY(cudaStreamWaitEvent(side, e, 0));
    note_other(side);
    return FM_OK;
  }
  // main stream waits for everything enqueued on the side stream
  int join() {
    if (!ok || nfork == 0) return FM_OK;
    CU_TRY(cudaEventRecord(ev[7], side));
    CU_TRY(cudaStreamWaitEvent(main, ev[7], 0));
    note_other(main);
    return FM_OK;
  }
};

// ================================================================================================ LayerNorm / misc launchers
// threads per row / chunks per thread: the smallest TPR in {32,64,128,256} with <= 2 eight-element chunks per thread
// (D <= 4096), else 256 threads x 4 chunks (D <= 8192)
static int ln_maxc(int D) { return D <= 4096 ? 2 : LN_MAXC_WIDE; }
static int ln_tpr(int D) { const int need = (D / 8 + ln_maxc(D) - 1) / ln_maxc(D); return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : 256; }
static int ln_grid(int rows, int tpr, int ctas_per_sm) {
  const int rpc = LN_THREADS / tpr;
  const int want = (rows + rpc - 1) / rpc;
  const int cap = g_num_sms * ctas_per_sm;
  return want < cap ? want : cap;
}

// warp-per-row kernels for D <= LN_WARP_MAX_D: chunks per lane
static int lnw_maxc(int D) { return (D / 8 + 31) / 32; }
template <typename A, typename K1, typename K2, typename K3, typename K4, typename K6>
static void launch_lnw(int maxc, int grid, size_t smem, cudaStream_t s, const A& a, K1 k1, K2 k2, K3 k3, K4 k4, K6 k6) {
  switch (maxc) {
    case 1:  (void)launch_k(k1, grid, LN_THREADS, smem, s, a); break;
    case 2:  (void)launch_k(k2, grid, LN_THREADS, smem, s, a); break;
    case 3:  (void)launch_k(k3, grid, LN_THREADS, smem, s, a); break;
    case 4:  (void)launch_k(k4, grid, LN_THREADS, smem, s, a); break;
    default: (void)launch_k(k6, grid, LN_THREADS, smem, s, a); break;
  }
}
static int lnw_grid(int rows, int ctas_per_sm) {
  const int wpc = LN_THREADS / 32;
  const int want = (rows + wpc - 1) / wpc;
  const int cap = (g_num_sms > 0 ? g_num_sms : 1) * ctas_per_sm;
  return want < cap ? want : cap;
}

static int run_ln_fwd(const LnArgs& a, cudaStream_t s) {
  FM_TRY(device_init());
  if (a.D % 8 != 0 || a.D > LN_THREADS * LN_MAXC_WIDE * 8 || a.rows <= 0) return fail(FM_EINVAL, "LayerNorm: D=%d must be a multiple of 8 and <= %d", a.D, LN_THREADS * LN_MAXC_WIDE * 8);
  {
    ProfScope ps("ln_fwd", 0.0, (double)a.rows * a.D * ((a.x_f32 ? 4 : 2) + (a.out_f32 ? 4 : 2) + (a.out2 ? 2 : 0)), s);
    if (a.D <= LN_WARP_MAX_D) {
      launch_lnw(lnw_maxc(a.D), lnw_grid(a.rows, 6), 0, s, a, ln_fwd_w_kernel<1>, ln_fwd_w_kernel<2>, ln_fwd_w_kernel<3>, ln_fwd_w_kernel<4>,
                 ln_fwd_w_kernel<6>);
    } else {
      const int tpr = ln_tpr(a.D);
      const int grid = ln_grid(a.rows, tpr, 3);     // fewer, longer-lived CTAs: rows are software-pipelined inside the kernel
      if (ln_maxc(a.D) == 2) {
        switch (tpr) {
          case 32:  (void)launch_k(ln_fwd_kernel<32, 2>, grid, LN_THREADS, 0, s, a); break;
          case 64:  (void)launch_k(ln_fwd_kernel<64, 2>, grid, LN_THREADS, 0, s, a); break;
          case 128: (void)launch_k(ln_fwd_kernel<128, 2>, grid, LN_THREADS, 0, s, a); break;
          default:  (void)launch_k(ln_fwd_kernel<256, 2>, grid, LN_THREADS, 0, s, a); break;
        }
      } else {
        (void)launch_k(ln_fwd_kernel<256, LN_MAXC_WIDE>, grid, LN_THREADS, 0, s, a);
      }
    }
  }
  KERNEL_CHECK();
  return FM_OK;
}
static size_t ln_part_bytes(int D) { return (size_t)448 * 2 * (size_t)D * sizeof(float); }
// `ss` (optional): the dgamma/dbeta work is a leaf of the backward graph, so with FM_OPT_LN_REDUCE_SIDE it is issued on the
// side stream (the caller then owns a distinct `part` buffer per LayerNorm until its next ss.join()).
static int run_ln_bwd(LnBwdArgs a, float* dgamma, float* dbeta, cudaStream_t s, SideStream* ss = nullptr) {
  FM_TRY(device_init());
  if (a.D % 8 != 0 || a.D > LN_THREADS * LN_MAXC_WIDE * 8 || a.rows <= 0) return fail(FM_EINVAL, "LayerNorm bwd: bad D=%d", a.D);
  const double row_bytes = (double)a.D * ((a.x_f32 ? 4 : 2) + 2 + (a.dy2 ? 2 : 0));
  if (a.D <= LN_WARP_MAX_D) {
    // dx on the caller's stream; column sums (re-reading x and dy) + their fold on the side stream when there is one
    const int maxc = lnw_maxc(a.D);
    cudaStream_t sr = s;
    if (ss && ss->ok && opt(FM_OPT_LN_REDUCE_SIDE)) { FM_TRY(ss->fork()); sr = ss->side; }
    if (a.dx != nullptr) {
      ProfScope ps("ln_bwd_dx", 0.0, (double)a.rows * (row_bytes + a.D * ((a.dres ? (a.dres_f32 ? 4 : 2) : 0) + (a.dx_f32 ? 4 : 2))), s);
      launch_lnw(maxc, lnw_grid(a.rows, 6), 0, s, a, ln_bwd_dx_w_kernel<1>, ln_bwd_dx_w_kernel<2>, ln_bwd_dx_w_kernel<3>, ln_bwd_dx_w_kernel<4>,
                 ln_bwd_dx_w_kernel<6>);
    }
    KERNEL_CHECK();
    int grid = lnw_grid(a.rows, 2);
    if (grid > 448) grid = 448;
    {
      ProfScope ps("ln_bwd_dgb", 0.0, (double)a.rows * row_bytes + (double)grid * 2 * a.D * 4, sr);
      launch_lnw(maxc, grid, (size_t)2 * a.D * sizeof(float), sr, a, ln_bwd_dgb_w_kernel<1>, ln_bwd_dgb_w_kernel<2>, ln_bwd_dgb_w_kernel<3>,
                 ln_bwd_dgb_w_kernel<4>, ln_bwd_dgb_w_kernel<6>);
    }
    KERNEL_CHECK();
    {
      ProfScope ps("ln_bwd_reduce", 0.0, (double)grid * 2 * a.D * 4, sr);
      (void)launch_k(ln_bwd_reduce_kernel, (2 * a.D + 31) / 32, 256, 0, sr, a.part, grid, a.D, dgamma, dbeta, 0);
    }
    KERNEL_CHECK();
    return FM_OK;
  }
  const int tpr = ln_tpr(a.D);
  int grid = ln_grid(a.rows, tpr, 2);
  if (grid > 448) grid = 448;
  const size_t sm = tpr < LN_THREADS ? (size_t)2 * a.D * sizeof(float) : 0;
  {
    ProfScope ps("ln_bwd", 0.0, (double)a.rows * a.D * ((a.x_f32 ? 4 : 2) + 2 + (a.dy2 ? 2 : 0) + (a.dres ? (a.dres_f32 ? 4 : 2) : 0) + (a.dx ? (a.dx_f32 ? 4 : 2) : 0)), s);
    if (ln_maxc(a.D) == 2) {
      switch (tpr) {
        case 32:  (void)launch_k(ln_bwd_kernel<32, 2>, grid, LN_THREADS, sm, s, a); break;
        case 64:  (void)launch_k(ln_bwd_kernel<64, 2>, grid, LN_THREADS, sm, s, a); break;
        case 128: (void)launch_k(ln_bwd_kernel<128, 2>, grid, LN_THREADS, sm, s, a); break;
        default:  (void)launch_k(ln_bwd_kernel<256, 2>, grid, LN_THREADS, sm, s, a); break;
      }
    } else {
      (void)launch_k(ln_bwd_kernel<256, LN_MAXC_WIDE>, grid, LN_THREADS, sm, s, a);
    }
  }
  KERNEL_CHECK();
  cudaStream_t sr = s;
  if (ss && ss->ok && opt(FM_OPT_LN_REDUCE_SIDE)) { FM_TRY(ss->fork()); sr = ss->side; }
  {
    ProfScope ps("ln_bwd_reduce", 0.0, (double)grid * 2 * a.D * 4, sr);
    (void)launch_k(ln_bwd_reduce_kernel, (2 * a.D + 31) / 32, 256, 0, sr, a.part, grid, a.D, dgamma, dbeta, 0);
  }
  KERNEL_CHECK();
  return FM_OK;
}
static LnArgs mk_ln(const void* x, int x_f32, const float* gamma, const float* beta, void* out, int out_f32, float* mean, float* rstd,
                    int rows, int D) {
  LnArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.x_f32 = x_f32; a.gamma = gamma; a.beta = beta; a.out = out; a.out_f32 = out_f32; a.mean = mean; a.rstd = rstd;
  a.rows = rows; a.D = D; a.in_group = rows; a.out_group = rows; a.out_off = 0; a.add_period = 1; a.add_group = 1;
  return a;
}
static LnBwdArgs mk_ln_bwd(const void* dy, const void* x, int x_f32, const float* gamma, const float* mean, const float* rstd,
                           const void* dres, int dres_f32, void* dx, int dx_f32, void* part, int rows, int D) {
  LnBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.dy = (const bf16*)dy; a.x = x; a.x_f32 = x_f32; a.gamma = gamma; a.mean = mean; a.rstd = rstd; a.dres = dres; a.dres_f32 = dres_f32;
  a.dx = dx; a.dx_f32 = dx_f32; a.part = (float*)part; a.rows = rows; a.D = D;
  a.in_group = rows; a.out_group = rows; a.out_off = 0; a.add_period = 1; a.add_group = 1;
  return a;
}

extern "C" int fm_layernorm_fwd(const void* x, int x_f32, const float* gamma, const float* beta, void* out, int out_f32, float* mean,
                                float* rstd, int rows, int D, fm_stream_t stream) {
  ApiScope api_scope;
  return run_ln_fwd(mk_ln(x, x_f32, gamma, beta, out, out_f32, mean, rstd, rows, D), (cudaStream_t)stream);
}
extern "C" size_t fm_layernorm_bwd_scratch_bytes(int D) { return ln_part_bytes(D); }
extern "C" int fm_layernorm_bwd(const void* dy, const void* x, int x_f32, const float* gamma, const float* mean, const float* rstd,
                                const void* dres, int dres_f32, void* dx, int dx_f32, float* dgamma, float* dbeta, void* part, int rows,
                                int D, fm_stream_t stream) {
  ApiScope api_scope;
  return run_ln_bwd(mk_ln_bwd(dy, x, x_f32, gamma, mean, rstd, dres, dres_f32, dx, dx_f32, part, rows, D), dgamma, dbeta, (cudaStream_t)stream);
}
extern "C" int fm_text_time(const int* ml, int* tt, int B, int S, fm_stream_t stream) {
  ApiScope api_scope;
  FM_TRY(device_init());
  if (B <= 0 || S <= 0) return fail(FM_EINVAL, "text_time: empty input");
  ProfScope ps("text_time", 0.0, 8.0 * B * S, (cudaStream_t)stream);
  (void)launch_k(text_time_kernel, (B + 3) / 4, 128, 0, (cudaStream_t)stream, ml, tt, B, S);
  KERNEL_CHECK();
  return FM_OK;
}
static int run_cast(const float* src, void* dst, long long n, cudaStream_t s) {
  if (n <= 0) return FM_OK;
  const long long threads = (n + 7) / 8;
  ProfScope ps("cast_f32_bf16", 0.0, 6.0 * n, s);
  (void)launch_k(cast_f32_bf16_kernel, (unsigned)((threads + 255) / 256), 256, 0, s, src, (bf16*)dst, n);
  KERNEL_CHECK();
  return FM_OK;
}
extern "C" int fm_cast_f32_to_bf16(const float* src, void* dst, long long n, fm_stream_t stream) {
  ApiScope api_scope;
  FM_TRY(device_init());
  return run_cast(src, dst, n, (cudaStream_t)stream);
}

// ================================================================================================ optimizer
extern "C" int fm_adamw_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, const float* decay_mask,
                             const float* grad_scale, long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                             int step, fm_stream_t stream) {
  ApiScope api_scope;
  FM_TRY(device_init());
  if (!p || !g || !m || !v || n <= 0 || step < 1) return fail(FM_EINVAL, "fm_adamw_step: null pointer, empty arena or step < 1");
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
        reinterpret_cast<uintptr_t>(decay_mask)) & 15) != 0 || (reinterpret_cast<uintptr_t>(shadow_bf16) & 7) != 0)
    return fail(FM_EINVAL, "fm_adamw_step: arenas must be 16-byte aligned");
  AdamWArgs a;
  a.p = p; a.g = g; a.m = m; a.v = v; a.shadow = (bf16*)shadow_bf16; a.decay_mask = decay_mask; a.grad_scale = grad_scale; a.n = n;
  a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.wd = weight_decay;
  a.bc1 = 1.0f - powf(beta1, (float)step);
  a.bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  cudaStream_t s = (cudaStream_t)stream;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)g_num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  {
    ProfScope ps("adamw", 0.0, (double)n * (16.0 + (decay_mask ? 4.0 : 0.0) + 12.0 + (shadow_bf16 ? 2.0 : 0.0)), s);
    (void)launch_k(adamw_kernel, (unsigned)blocks, 256, 0, s, a);
  }
  KERNEL_CHECK();
  return FM_OK;
}

// ================================================================================================ loss head
static int check_ce(const void* logits, long long ld, int rows, int vocab, const long long* targets, const float* lse) {
  if (!logits || !targets || !lse) return fail(FM_EINVAL, "cross entropy: null pointer");
  if (rows <= 0 || vocab <= 0 || ld < vocab || ld % 8 != 0) return fail(FM_EINVAL, "cross entropy: need rows > 0, 0 < vocab <= ld, ld %% 8 == 0 (rows=%d vocab=%d ld=%lld)", rows, vocab, ld);
  if ((reinterpret_cast<uintptr_t>(logits) & 15) != 0) return fail(FM_EINVAL, "cross entropy: logits must be 16-byte aligned");
  return FM_OK;
}
extern "C" int fm_cross_entropy_fwd(const void* logits, long long ld, int rows, int vocab, const long long* targets,
                                    long long ignore_index, float* lse, float* row_loss, fm_stream_t stream) {
  ApiScope api_scope;
  FM_TRY(device_init());
  FM_TRY(check_ce(logits, ld, rows, vocab, targets, lse));
  if (!row_loss) return fail(FM_EINVAL, "cross entropy: null row_loss");
  CeArgs a;
  memset(&a, 0, sizeof(a));
  a.logits = (const bf16*)logits; a.targets = targets; a.ignore_index = ignore_index; a.lse = lse; a.row_loss = row_loss;
  a.rows = rows; a.vocab = vocab; a.ld = ld;
  cudaStream_t s = (cudaStream_t)stream;
  {
    ProfScope ps("ce_fwd", 0.0, 2.0 * rows * (double)vocab, s);
    (void)launch_k(ce_fwd_kernel, rows, CE_THREADS, 0, s, a);
  }
  KERNEL_CHECK();
  return FM_OK;
}
extern "C" int fm_cross_entropy_bwd(const void* logits, long long ld, int rows, int vocab, const long long* targets,
                                    long long ignore_index, const float* lse, const float* scale, void* dlogits, fm_stream_t stream) {
  ApiScope api_scope;
  FM_TRY(device_init());
  FM_TRY(check_ce(logits, ld, rows, vocab, targets, lse));
  if (!scale || !dlogits || (reinterpret_cast<uintptr_t>(dlogits) & 15) != 0) return fail(FM_EINVAL, "cross entropy bwd: scale / 16-byte aligned dlogits required");
  CeArgs a;
  memset(&a, 0, sizeof(a));
  a.logits = (const bf16*)logits; a.targets = targets; a.ignore_index = ignore_index; a.lse = const_cast<float*>(lse);
  a.dlogits = (bf16*)dlogits; a.scale = scale; a.rows = rows; a.vocab = vocab; a.ld = ld;
  cudaStream_t s = (cudaStream_t)stream;
  {
    ProfScope ps("ce_bwd", 0.0, 2.0 * rows * ((double)vocab + (double)ld), s);
    (void)launch_k(ce_bwd_kernel, rows, CE_THREADS, 0, s, a);
  }
  KERNEL_CHECK();
  return FM_OK;
}

// ================================================================================================ workspace carving
struct Carver {
  char* base; size_t off;
  explicit Carver(void* p) : base((char*)p), off(0) {}
  template <typename T> T* take(size_t n) {
    T* r = base ? (T*)(base + off) : nullptr;
    off += ((n * sizeof(T) + 255) / 256) * 256;
    return r;
  }
};
static long long align8(long long v) { return (v + 7) / 8 * 8; }


// ================================================================================================ gated xattn block
static int check_xattn_cfg(const fm_xattn_cfg* c) {
  if (!c) return fail(FM_EINVAL, "null cfg");
  if (c->heads < 1 || c->heads > 64 || c->dim_head != 64) return fail(FM_EINVAL, "attention cores are specialised for dim_head=64 with 1..64 heads (got heads=%d, dim_head=%d)", c->heads, c->dim_head);
  if (c->B <= 0 || c->S <= 0 || c->n_media <= 0) return fail(FM_EINVAL, "empty xattn problem B=%d S=%d n_media=%d", c->B, c->S, c->n_media);
  if (c->D % 64 != 0 || c->Dv % 64 != 0 || c->ff_inner % 64 != 0) return fail(FM_EINVAL, "D, Dv, ff_inner must be multiples of 64 (got %d, %d, %d)", c->D, c->Dv, c->ff_inner);
  if (c->act < 0 || c->act > 2) return fail(FM_EINVAL, "unknown activation %d", c->act);
  return FM_OK;
}
extern "C" int fm_xattn_layout_of(const fm_xattn_cfg* c, fm_xattn_layout* L) {
  FM_TRY(check_xattn_cfg(c));
  const long long I = (long long)c->heads * c->dim_head;
  long long o = 0;
  L->attn_norm_w = o; o += c->D;
  L->attn_norm_b = o; o += c->D;
  L->to_q = o; o += I * c->D;
  L->to_kv = o; o += 2 * I * c->Dv;
  L->to_out = o; o += (long long)c->D * I;
  L->ffw_norm_w = o; o += c->D;
  L->ffw_norm_b = o; o += c->D;
  L->ffw_w1 = o; o += (long long)c->ff_inner * c->D;
  L->ffw_w2 = o; o += (long long)c->D * c->ff_inner;
  L->alpha_attn = o; o += 1;
  L->alpha_ffw = o; o += 1;
  L->total = align8(o);
  return FM_OK;
}

struct XSaved {
  bf16 *yn, *q, *o, *y1n, *h_pre /* act'(pre-activation) */, *h_act;
  float *y1, *mean1, *rstd1, *mean2, *rstd2;
  size_t bytes;
};
static XSaved carve_xsaved(const fm_xattn_cfg* c, void* p) {
  const size_t M = (size_t)c->B * c->S, I = c->heads * 64;
  Carver cv(p);
  XSaved s;
  s.yn = cv.take<bf16>(M * c->D);
  s.q = cv.take<bf16>(M * I);
  s.o = cv.take<bf16>(M * I);
  s.y1 = cv.take<float>(M * c->D);
  s.y1n = cv.take<bf16>(M * c->D);
  s.h_pre = cv.take<bf16>(M * c->ff_inner);
  s.h_act = cv.take<bf16>(M * c->ff_inner);
  s.mean1 = cv.take<float>(M); s.rstd1 = cv.take<float>(M);
  s.mean2 = cv.take<float>(M); s.rstd2 = cv.take<float>(M);
  s.bytes = cv.off;
  return s;
}
struct XScratch {
  bf16 *dyo, *dh, *dy1n, *dy1, *do_u, *dq, *dkv, *dyn;
  float* red;      // [8] floats followed by the split-K flags (one memset clears both)
  int* flags;
  void* ln_part[2];   // one per LayerNorm backward of the block (their folds may still be running on the side stream)
  size_t bytes;
};
static constexpr size_t SPLITK_FLAG_INTS = 16384;
static XScratch carve_xscratch(const fm_xattn_cfg* c, void* p) {
  const size_t M = (size_t)c->B * c->S, I = c->heads * 64, V = (size_t)c->B * c->n_media * 64;
  Carver cv(p);
  XScratch s;
  s.dyo = cv.take<bf16>(c->y_f32 ? M * c->D : 0);
  s.dh = cv.take<bf16>(M * c->ff_inner);
  s.dy1n = cv.take<bf16>(M * c->D);
  s.dy1 = cv.take<bf16>(M * c->D);
  s.do_u = cv.take<bf16>(M * I);
  s.dq = cv.take<bf16>(M * I);
  s.dkv = cv.take<bf16>(V * 2 * I);
  s.dyn = cv.take<bf16>(M * c->D);
  s.red = cv.take<float>(64);
  s.flags = cv.take<int>(SPLITK_FLAG_INTS);
  for (int i = 0; i < 2; ++i) s.ln_part[i] = cv.take<char>(ln_part_bytes(c->D));
  s.bytes = cv.off;
  return s;
}
extern "C" size_t fm_xattn_saved_bytes(const fm_xattn_cfg* c) { return check_xattn_cfg(c) == FM_OK ? carve_xsaved(c, nullptr).bytes : 0; }
extern "C" size_t fm_xattn_scratch_bytes(const fm_xattn_cfg* c) { return check_xattn_cfg(c) == FM_OK ? carve_xscratch(c, nullptr).bytes : 0; }

extern "C" int fm_xattn_fwd(const fm_xattn_cfg* c, const float* wf, const void* wb_, const void* y, const void* vis, const int* tt,
                            void* kv, int kv_given, void* y_out, void* saved, fm_stream_t stream) {
  ApiScope api_scope("x/");
  FM_TRY(check_xattn_cfg(c));
  FM_TRY(device_init());
  if (!wf || !wb_ || !y || !tt || !kv || !y_out || !saved) return fail(FM_EINVAL, "fm_xattn_fwd: null pointer");
  if (!kv_given && !vis) return fail(FM_EINVAL, "fm_xattn_fwd: visual features required unless kv_given");
  cudaStream_t s = (cudaStream_t)stream;
  fm_xattn_layout L;
  FM_TRY(fm_xattn_layout_of(c, &L));
  const bf16* wb = (const bf16*)wb_;
  const int M = c->B * c->S, D = c->D, Dv = c->Dv, FF = c->ff_inner, I = c->heads * 64, V = c->B * c->n_media * 64;
  XSaved sv = carve_xsaved(c, saved);

  // 1. yn = LN(y)                                                         gated_cross_attention.py:74
  FM_TRY(run_ln_fwd(mk_ln(y, c->y_f32, wf + L.attn_norm_w, wf + L.attn_norm_b, sv.yn, 0, sv.mean1, sv.rstd1, M, D), s));
  // 2. q = (yn Wq^T) * dim_head^-0.5                                      :77-78
  // 3. [k | v] = vis Wkv^T                                                :84-86   (independent of 2: one grouped launch)
  {
    fm_gemm_desc grp[2];
    int n = 0;
    grp[n] = mk_gemm(M, I, D, sv.yn, D, 0, wb + L.to_q, D, 0, EPI_STORE, sv.q, I, 0);
    grp[n++].scale = 0.125f;
    if (!kv_given) grp[n++] = mk_gemm(V, 2 * I, Dv, vis, Dv, 0, wb + L.to_kv, Dv, 0, EPI_STORE, kv, 2 * I, 0);
    FM_TRY(run_gemm_group(grp, n, s));
  }
  // 4. masked softmax(q k^T) v                                            :95-124   (tcgen05: attn_tc.cuh)
  {
    const cudaError_t aerr = ensure_dyn_smem(reinterpret_cast<const void*>(xattn_core_fwd_tc_kernel), XTC_FWD_SMEM);
    if (aerr != cudaSuccess) return fail(FM_ECUDA, "cudaFuncSetAttribute(xattn_core_fwd_tc) failed: %s", cudaGetErrorString(aerr));
    CUtensorMap tmQ, tmKV;
    FM_TRY(make_tmap_2d(&tmQ, sv.q, I, M, I, 64, 128));
    FM_TRY(make_tmap_2d(&tmKV, kv, 2 * I, V, 2 * I, 64, 64));
    XTcArgs a;
    a.tt = tt; a.o = sv.o; a.B = c->B; a.S = c->S; a.H = c->heads; a.n_media = c->n_media;
    ProfScope ps("xattn_core_fwd", 4.0 * M * 64 * I, 2.0 * (2.0 * M * I + 2.0 * V * I), s);
    (void)launch_k(xattn_core_fwd_tc_kernel, dim3((c->S + 127) / 128, c->heads, c->B), 128, XTC_FWD_SMEM, s, tmQ, tmKV, a);
    KERNEL_CHECK();
  }
  // 5. y1 = y + tanh(alpha_attn) * (o Wout^T)                             :126, :180
  {
    fm_gemm_desc g = mk_gemm(M, D, I, sv.o, I, 0, wb + L.to_out, I, 0, EPI_RESID, sv.y1, D, 1);
    g.aux = y; g.ldaux = D; g.aux_f32 = c->y_f32; g.gate = wf + L.alpha_attn;
    FM_TRY(run_gemm(g, s));
  }
  // 6. y1n = LN(y1)                                                       utils.py:46
  FM_TRY(run_ln_fwd(mk_ln(sv.y1, 1, wf + L.ffw_norm_w, wf + L.ffw_norm_b, sv.y1n, 0, sv.mean2, sv.rstd2, M, D), s));
  // 7. h = act(y1n W1^T)                                                  utils.py:47-48
  {
    fm_gemm_desc g = mk_gemm(M, FF, D, sv.y1n, D, 0, wb + L.ffw_w1, D, 0, EPI_ACT, sv.h_act, FF, 0);
    g.out2 = c->training ? sv.h_pre : nullptr; g.ldo2 = FF; g.act = c->act;
    FM_TRY(run_gemm(g, s));
  }
  // 8. y_out = y1 + tanh(alpha_ffw) * (h W2^T)                            utils.py:49, gated_cross_attention.py:182
  {
    fm_gemm_desc g = mk_gemm(M, D, FF, sv.h_act, FF, 0, wb + L.ffw_w2, FF, 0, EPI_RESID, y_out, D, c->y_f32);
    g.aux = sv.y1; g.ldaux = D; g.aux_f32 = 1; g.gate = wf + L.alpha_ffw;
    FM_TRY(run_gemm(g, s));
  }
  return FM_OK;
}

extern "C" int fm_xattn_bwd(const fm_xattn_cfg* c, const float* wf, const void* wb_, const void* y, const void* vis, const int* tt,
                            const void* kv, const void* saved, const void* dy_out, void* dy, void* dvis, float* gf, void* scratch,
                            fm_stream_t stream) {
  ApiScope api_scope("x/");
  FM_TRY(check_xattn_cfg(c));
  FM_TRY(device_init());
  if (!wf || !wb_ || !y || !tt || !kv || !saved || !dy_out || !dy || !gf || !scratch) return fail(FM_EINVAL, "fm_xattn_bwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  fm_xattn_layout L;
  FM_TRY(fm_xattn_layout_of(c, &L));
  const bf16* wb = (const bf16*)wb_;
  const int M = c->B * c->S, D = c->D, Dv = c->Dv, FF = c->ff_inner, I = c->heads * 64, V = c->B * c->n_media * 64;
  XSaved sv = carve_xsaved(c, const_cast<void*>(saved));
  XScratch sc = carve_xscratch(c, scratch);
  CU_TRY(cudaMemsetAsync(sc.red, 0, (size_t)((char*)(sc.flags + SPLITK_FLAG_INTS) - (char*)sc.red), s)); note_other(s);

  const bf16* dyo = (const bf16*)dy_out;
  if (c->y_f32) {
    FM_TRY(run_cast((const float*)dy_out, sc.dyo, (long long)M * D, s));
    dyo = sc.dyo;
  }
  // dh = tanh(a_f) * (dyo W2) * act'(h_pre);  red[0] = sum((dyo W2) * act(h_pre))   (both saved by the forward epilogue)
  {
    fm_gemm_desc g = mk_gemm(M, FF, D, dyo, D, 0, wb + L.ffw_w2, FF, 1, EPI_DACT, sc.dh, FF, 0);
    g.aux = sv.h_pre; g.ldaux = FF; g.gate = wf + L.alpha_ffw; g.act = c->act;
    FM_TRY(run_gemm(g, s));
  }
  SideStream ss(s);
  cudaStream_t s2 = ss.ok ? ss.side : s;     // weight-gradient GEMMs run on the side stream
  FM_TRY(ss.fork());
  // dW2[d, f] = tanh(a_f) * sum_m dyo[m, d] h_act[m, f]
  int bn_dw2 = 0, bn_dw1 = 0;
  const int sp_dw2 = plan_dw(D, FF, M, &bn_dw2), sp_dw1 = plan_dw(FF, D, M, &bn_dw1);
  if (sp_dw2 < 0 || sp_dw1 < 0) {      // parallel split-K adds into zeroed gradients: ffw.1.weight and ffw.3.weight are adjacent
    CU_TRY(cudaMemsetAsync(gf + L.ffw_w1, 0, sizeof(float) * 2 * (size_t)FF * D, s2)); note_other(s2);
  }
  {
    fm_gemm_desc g = mk_gemm(D, FF, M, dyo, D, 1, sv.h_act, FF, 1, EPI_STORE, gf + L.ffw_w2, FF, 1, sc.flags);
    if (sp_dw2 < 0) { g.splits = sp_dw2; g.bn = bn_dw2; }
    g.gate = wf + L.alpha_ffw;
    // sum(dY W2 * h) == sum(W2 * (dY^T h)): the un-gated accumulator of this GEMM dotted with W2 gives d(alpha_ffw)'s raw sum
    g.aux = wb + L.ffw_w2; g.ldaux = FF; g.red_out = sc.red + 0;
    FM_TRY(run_gemm(g, s2));
  }
  // dW1[f, d] = sum_m dh[m, f] y1n[m, d]
  {
    fm_gemm_desc g = mk_gemm(FF, D, M, sc.dh, FF, 1, sv.y1n, D, 1, EPI_STORE, gf + L.ffw_w1, D, 1, sc.flags);
    if (sp_dw1 < 0) { g.splits = sp_dw1; g.bn = bn_dw1; }
    FM_TRY(run_gemm(g, s2));
  }
  // dy1n = dh W1
  FM_TRY(run_gemm(mk_gemm(M, D, FF, sc.dh, FF, 0, wb + L.ffw_w1, D, 1, EPI_STORE, sc.dy1n, D, 0), s));
  // dy1 = dy_out + LNbwd(dy1n)
  FM_TRY(run_ln_bwd(mk_ln_bwd(sc.dy1n, sv.y1, 1, wf + L.ffw_norm_w, sv.mean2, sv.rstd2, dy_out, c->y_f32, sc.dy1, 0, sc.ln_part[0], M, D),
                    gf + L.ffw_norm_w, gf + L.ffw_norm_b, s, &ss));
  // do_u = dy1 Wout   (gradient w.r.t. o before the gate)
  // red[1] = sum(do_u * o) feeds d(alpha_attn): either from this GEMM's epilogue (fp32 accumulators against the saved o) or
  // from a separate pass over the bf16 do_u on the side stream
  {
    fm_gemm_desc g = mk_gemm(M, I, D, sc.dy1, D, 0, wb + L.to_out, I, 1, EPI_STORE, sc.do_u, I, 0);
    if (opt(FM_OPT_DATTN_FROM_GEMM)) { g.aux = sv.o; g.ldaux = I; g.red_out = sc.red + 1; }
    FM_TRY(run_gemm(g, s));
  }
  FM_TRY(ss.fork());
  if (!opt(FM_OPT_DATTN_FROM_GEMM)) {
    ProfScope ps("dot_reduce", 0.0, 4.0 * M * I, s2);
    (void)launch_k(dot_reduce_kernel, g_num_sms * 2, 256, 0, s2, sc.do_u, sv.o, (long long)M * I, sc.red + 1);
  }
  KERNEL_CHECK();
  // attention core backward (tcgen05: attn_tc.cuh)
  {
    const cudaError_t aerr = ensure_dyn_smem(reinterpret_cast<const void*>(xattn_core_bwd_tc_kernel), XTC_BWD_SMEM);
    if (aerr != cudaSuccess) return fail(FM_ECUDA, "cudaFuncSetAttribute(xattn_core_bwd_tc) failed: %s", cudaGetErrorString(aerr));
    CUtensorMap tmQ, tmDO, tmKV;
    FM_TRY(make_tmap_2d(&tmQ, sv.q, I, M, I, 64, 128));
    FM_TRY(make_tmap_2d(&tmDO, sc.do_u, I, M, I, 64, 128));
    FM_TRY(make_tmap_2d(&tmKV, kv, 2 * I, V, 2 * I, 64, 64));
    XTcBwdArgs a;
    a.tt = tt; a.gate = wf + L.alpha_attn; a.d_o = sc.do_u; a.dq = sc.dq; a.dkv = sc.dkv; a.q_scale = 0.125f;
    a.B = c->B; a.S = c->S; a.H = c->heads; a.n_media = c->n_media; a.tmem_compact = opt(FM_OPT_ATTN_TMEM_COMPACT);
    ProfScope ps("xattn_core_bwd", 10.0 * M * 64 * I, 2.0 * (3.0 * M * I + 4.0 * V * I), s);
    (void)launch_k(xattn_core_bwd_tc_kernel, dim3(c->heads, c->B), 128, XTC_BWD_SMEM, s, tmQ, tmDO, tmKV, a);
    KERNEL_CHECK();
  }
  FM_TRY(ss.fork());
  // dWout[d, i] = tanh(a_a) sum_m dy1[m, d] o[m, i];  dWq[i, d] = sum_m dq[m, i] yn[m, d];  dWkv[c, e] = sum_r dkv[r, c] vis[r, e]
  // -> ONE grouped launch (48 + 48 + 96 tiles at C2)
  {
    fm_gemm_desc grp[3];
    int n = 0;
    grp[n] = mk_gemm(D, I, M, sc.dy1, D, 1, sv.o, I, 1, EPI_STORE, gf + L.to_out, I, 1); grp[n].gate = wf + L.alpha_attn; ++n;
    grp[n++] = mk_gemm(I, D, M, sc.dq, I, 1, sv.yn, D, 1, EPI_STORE, gf + L.to_q, D, 1);
    if (vis) grp[n++] = mk_gemm(2 * I, Dv, V, sc.dkv, 2 * I, 1, vis, Dv, 1, EPI_STORE, gf + L.to_kv, Dv, 1);
    if (opt(FM_OPT_GEMM_GROUP) && plan_dw_group(grp, n)) {       // to_q, to_kv, to_out are adjacent in the arena: one memset
      CU_TRY(cudaMemsetAsync(gf + L.to_q, 0, sizeof(float) * (size_t)(L.ffw_norm_w - L.to_q), s2)); note_other(s2);
    }
    FM_TRY(run_gemm_group(grp, n, s2));
  }
  // dyn = dq Wq;  dvis = dkv Wkv   (independent: one grouped launch)
  {
    fm_gemm_desc grp[2];
    int n = 0;
    grp[n++] = mk_gemm(M, D, I, sc.dq, I, 0, wb + L.to_q, D, 1, EPI_STORE, sc.dyn, D, 0);
    if (vis && dvis) grp[n++] = mk_gemm(V, Dv, 2 * I, sc.dkv, 2 * I, 0, wb + L.to_kv, Dv, 1, EPI_STORE, dvis, Dv, 0);
    FM_TRY(run_gemm_group(grp, n, s));
  }
  // dy = dy1 + LNbwd(dyn)
  FM_TRY(run_ln_bwd(mk_ln_bwd(sc.dyn, y, c->y_f32, wf + L.attn_norm_w, sv.mean1, sv.rstd1, sc.dy1, 0, dy, c->y_f32, sc.ln_part[1], M, D),
                    gf + L.attn_norm_w, gf + L.attn_norm_b, s, &ss));
  if (!vis) { CU_TRY(cudaMemsetAsync(gf + L.to_kv, 0, sizeof(float) * 2 * I * Dv, s)); note_other(s); }
  // Everything still running on the side stream is a leaf (weight gradients, LayerNorm folds).  Normally the caller's stream waits
  // for it here; with FM_OPT_DEFER_JOIN the wait is left to fm_side_join(), so those kernels overlap whatever the caller enqueues
  // next.  The gate gradients need both raw sums (red[0] may come from the side stream): they follow the side work in that case.
  const bool defer = ss.ok && opt(FM_OPT_DEFER_JOIN);
  if (defer) FM_TRY(ss.fork()); else FM_TRY(ss.join());
  {
    cudaStream_t sa = defer ? s2 : s;
    ProfScope ps("alpha_grad", 0.0, 32.0, sa);
    (void)launch_k(alpha_grad_kernel, 1, 32, 0, sa, wf + L.alpha_attn, wf + L.alpha_ffw, sc.red, gf + L.alpha_attn, gf + L.alpha_ffw);
  }
  KERNEL_CHECK();
  return FM_OK;
}
extern "C" int fm_get_option(int key) {
  return (key < 0 || key >= FM_OPT_COUNT) ? -1 : g_opt[key].load(std::memory_order_relaxed);
}
extern "C" int fm_side_join(fm_stream_t stream) {
  ApiScope api_scope;
  SideStream ss((cudaStream_t)stream);
  if (!ss.ok) return FM_OK;
  ss.nfork = 1;                 // join() is a no-op for a SideStream that has not forked: this one joins whatever is outstanding
  return ss.join();
}

// ================================================================================================ attention cores on their own
// the same kernels fm_xattn_fwd / fm_resampler_fwd launch, exported for the stand-alone module forwards
extern "C" int fm_xattn_core_fwd(const void* q, const void* kv, const int* tt, void* o, int B, int S, int n_media, int heads,
                                 fm_stream_t stream) {
  ApiScope api_scope;
  FM_TRY(device_init());
  if (!q || !kv || !tt || !o || B <= 0 || S <= 0 || n_media <= 0 || heads < 1 || heads > 64) return fail(FM_EINVAL, "fm_xattn_core_fwd: bad arguments");
  const cudaError_t aerr = ensure_dyn_smem(reinterpret_cast<const void*>(xattn_core_fwd_tc_kernel), XTC_FWD_SMEM);
  if (aerr != cudaSuccess) return fail(FM_ECUDA, "cudaFuncSetAttribute(xattn_core_fwd_tc) failed: %s", cudaGetErrorString(aerr));
  cudaStream_t s = (cudaStream_t)stream;
  const int I = heads * 64, M = B * S, V = B * n_media * 64;
  CUtensorMap tmQ, tmKV;
  FM_TRY(make_tmap_2d(&tmQ, q, I, M, I, 64, 128));
  FM_TRY(make_tmap_2d(&tmKV, kv, 2 * I, V, 2 * I, 64, 64));
  XTcArgs a;
  a.tt = tt; a.o = (bf16*)o; a.B = B; a.S = S; a.H = heads; a.n_media = n_media;
  {
    ProfScope ps("xattn_core_fwd", 4.0 * M * 64 * I, 2.0 * (2.0 * M * I + 2.0 * V * I), s);
    (void)launch_k(xattn_core_fwd_tc_kernel, dim3((S + 127) / 128, heads, B), 128, XTC_FWD_SMEM, s, tmQ, tmKV, a);
  }
  KERNEL_CHECK();
  return FM_OK;
}
extern "C" int fm_resampler_core_fwd(const void* q, const void* kv, void* o, float* lse, int BN, int nk, int heads, fm_stream_t stream) {
  ApiScope api_scope;
  FM_TRY(device_init());
  if (!q || !kv || !o || BN <= 0 || nk <= 0 || heads < 1 || heads > 64) return fail(FM_EINVAL, "fm_resampler_core_fwd: bad arguments");
  const cudaError_t aerr = ensure_dyn_smem(reinterpret_cast<const void*>(resampler_core_fwd_tc_kernel), XTC_FWD_SMEM);
  if (aerr != cudaSuccess) return fail(FM_ECUDA, "cudaFuncSetAttribute(resampler_core_fwd_tc) failed: %s", cudaGetErrorString(aerr));
  cudaStream_t s = (cudaStream_t)stream;
  const int I = heads * 64, R = BN * 64, KV = BN * nk;
  CUtensorMap tmQ, tmKV;
  FM_TRY(make_tmap_2d(&tmQ, q, I, R, I, 64, 128));
  FM_TRY(make_tmap_2d(&tmKV, kv, 2 * I, KV, 2 * I, 64, 64));
  RTcArgs a;
  a.o = (bf16*)o; a.lse = lse; a.BN = BN; a.H = heads; a.nk = nk;
  {
    ProfScope ps("resampler_core_fwd", 4.0 * R * nk * I, 2.0 * (2.0 * R * I + 2.0 * KV * I), s);
    (void)launch_k(resampler_core_fwd_tc_kernel, dim3(heads, BN), 128, XTC_FWD_SMEM, s, tmQ, tmKV, a);
  }
  KERNEL_CHECK();
  return FM_OK;
}

// backward of the two cores (same kernels fm_xattn_bwd / fm_resampler_bwd launch): d_o is the gradient w.r.t. the core output o,
// dq comes back multiplied by q_scale (i.e. w.r.t. the un-scaled query projection), dkv = [dK | dV] in the layout of kv
extern "C" int fm_xattn_core_bwd(const void* q, const void* kv, const int* tt, const void* d_o, void* dq, void* dkv, int B, int S,
                                 int n_media, int heads, float q_scale, fm_stream_t stream) {
  ApiScope api_scope;
  FM_TRY(device_init());
  if (!q || !kv || !tt || !d_o || !dq || !dkv || B <= 0 || S <= 0 || n_media <= 0 || heads < 1 || heads > 64)
    return fail(FM_EINVAL, "fm_xattn_core_bwd: bad arguments");
  const cudaError_t aerr = ensure_dyn_smem(reinterpret_cast<const void*>(xattn_core_bwd_tc_kernel), XTC_BWD_SMEM);
  if (aerr != cudaSuccess) return fail(FM_ECUDA, "cudaFuncSetAttribute(xattn_core_bwd_tc) failed: %s", cudaGetErrorString(aerr));
  cudaStream_t s = (cudaStream_t)stream;
  const int I = heads * 64, M = B * S, V = B * n_media * 64;
  CUtensorMap tmQ, tmDO, tmKV;
  FM_TRY(make_tmap_2d(&tmQ, q, I, M, I, 64, 128));
  FM_TRY(make_tmap_2d(&tmDO, d_o, I, M, I, 64, 128));
  FM_TRY(make_tmap_2d(&tmKV, kv, 2 * I, V, 2 * I, 64, 64));
  XTcBwdArgs a;
  a.tt = tt; a.gate = nullptr; a.d_o = (const bf16*)d_o; a.dq = (bf16*)dq; a.dkv = (bf16*)dkv; a.q_scale = q_scale;
  a.B = B; a.S = S; a.H = heads; a.n_media = n_media; a.tmem_compact = opt(FM_OPT_ATTN_TMEM_COMPACT);
  {
    ProfScope ps("xattn_core_bwd", 10.0 * M * 64 * I, 2.0 * (3.0 * M * I + 4.0 * V * I), s);
    (void)launch_k(xattn_core_bwd_tc_kernel, dim3(heads, B), 128, XTC_BWD_SMEM, s, tmQ, tmDO, tmKV, a);
  }
  KERNEL_CHECK();
  return FM_OK;
}
extern "C" int fm_resampler_core_bwd(const void* q, const void* kv, const void* o, const void* d_o, const float* lse, void* dq, void* dkv,
                                     int BN, int nk, int heads, float q_scale, fm_stream_t stream) {
  ApiScope api_scope;
  FM_TRY(device_init());
  if (!q || !kv || !o || !d_o || !lse || !dq || !dkv || BN <= 0 || nk <= 0 || heads < 1 || heads > 64)
    return fail(FM_EINVAL, "fm_resampler_core_bwd: bad arguments");
  const cudaError_t aerr = ensure_dyn_smem(reinterpret_cast<const void*>(resampler_core_bwd_tc_kernel), XTC_BWD_SMEM);
  if (aerr != cudaSuccess) return fail(FM_ECUDA, "cudaFuncSetAttribute(resampler_core_bwd_tc) failed: %s", cudaGetErrorString(aerr));
  cudaStream_t s = (cudaStream_t)stream;
  const int I = heads * 64, R = BN * 64, KV = BN * nk;
  CUtensorMap tmQ, tmDO, tmKV;
  FM_TRY(make_tmap_2d(&tmQ, q, I, R, I, 64, 128));
  FM_TRY(make_tmap_2d(&tmDO, d_o, I, R, I, 64, 128));
  FM_TRY(make_tmap_2d(&tmKV, kv, 2 * I, KV, 2 * I, 64, 64));
  RTcBwdArgs a;
  a.o = (const bf16*)o; a.d_o = (const bf16*)d_o; a.lse = lse; a.dq = (bf16*)dq; a.dkv = (bf16*)dkv; a.q_scale = q_scale;
  a.BN = BN; a.H = heads; a.nk = nk; a.tmem_compact = opt(FM_OPT_ATTN_TMEM_COMPACT);
  {
    ProfScope ps("resampler_core_bwd", 10.0 * R * nk * I, 2.0 * (4.0 * R * I + 4.0 * KV * I), s);
    (void)launch_k(resampler_core_bwd_tc_kernel, dim3(heads, BN), 128, XTC_BWD_SMEM, s, tmQ, tmDO, tmKV, a);
  }
  KERNEL_CHECK();
  return FM_OK;
}

// ================================================================================================ perceiver resampler
static int check_res_cfg(const fm_resampler_cfg* c) {
  if (!c) return fail(FM_EINVAL, "null cfg");
  if (c->heads < 1 || c->heads > 64 || c->dim_head != 64 || c->n_latents != 64)
    return fail(FM_EINVAL, "attention cores are specialised for dim_head=64, num_latents=64 with 1..64 heads (got heads=%d, dim_head=%d, num_latents=%d)", c->heads, c->dim_head, c->n_latents);
  if (c->BN <= 0 || c->T <= 0 || c->F <= 0 || c->depth <= 0) return fail(FM_EINVAL, "empty resampler problem");
  if (c->T > c->n_time_embeds) return fail(FM_EINVAL, "n_frames=%d exceeds num_time_embeds=%d (perceiver_resampler.py:166)", c->T, c->n_time_embeds);
  if (c->Dv % 64 != 0 || c->ff_inner % 64 != 0) return fail(FM_EINVAL, "Dv, ff_inner must be multiples of 64 (got %d, %d)", c->Dv, c->ff_inner);
  if (c->act < 0 || c->act > 2) return fail(FM_EINVAL, "unknown activation %d", c->act);
  return FM_OK;
}
extern "C" int fm_resampler_layout_of(const fm_resampler_cfg* c, fm_resampler_layout* L) {
  FM_TRY(check_res_cfg(c));
  const long long I = c->heads * 64, Dv = c->Dv, FF = c->ff_inner;
  long long o = 0;
  L->latents = o; o += (long long)c->n_latents * Dv;
  L->time_pos_emb = o; o += (long long)c->n_time_embeds * Dv;
  L->layer0 = o;
  long long r = 0;
  L->norm_media_w = r; r += Dv;
  L->norm_media_b = r; r += Dv;
  L->norm_latents_w = r; r += Dv;
  L->norm_latents_b = r; r += Dv;
  L->to_q = r; r += I * Dv;
  L->to_k = r; r += I * Dv;
  L->to_v = r; r += I * Dv;
  L->to_out = r; r += Dv * I;
  L->ffw_norm_w = r; r += Dv;
  L->ffw_norm_b = r; r += Dv;
  L->ffw_w1 = r; r += FF * Dv;
  L->ffw_w2 = r; r += Dv * FF;
  L->layer_stride = r;
  o += r * c->depth;
  L->norm_w = o; o += Dv;
  L->norm_b = o; o += Dv;
  L->total = align8(o);
  return FM_OK;
}

struct RLayerSaved {
  bf16 *kv_in, *lat_n, *q, *kv, *o, *xn2, *h_pre, *h_act;
  float *x_mid, *mean_l, *rstd_l, *mean2, *rstd2, *lse;
};
struct RSaved {
  float* x[17];        // x[l] = latent state entering layer l; x[depth] = final state
  float *mean_m, *rstd_m, *mean_f, *rstd_f;
  RLayerSaved layer[16];
  size_t bytes;
};
static RSaved carve_rsaved(const fm_resampler_cfg* c, void* p) {
  const size_t R = (size_t)c->BN * 64, Mm = (size_t)c->BN * c->T * c->F, nk = (size_t)c->T * c->F + 64, KV = (size_t)c->BN * nk;
  const size_t Dv = c->Dv, FF = c->ff_inner, I = c->heads * 64;
  Carver cv(p);
  RSaved s;
  for (int l = 0; l <= c->depth; ++l) s.x[l] = cv.take<float>(R * Dv);
  s.mean_m = cv.take<float>(Mm); s.rstd_m = cv.take<float>(Mm);
  s.mean_f = cv.take<float>(R); s.rstd_f = cv.take<float>(R);
  for (int l = 0; l < c->depth; ++l) {
    RLayerSaved& y = s.layer[l];
    y.kv_in = cv.take<bf16>(KV * Dv);
    y.lat_n = cv.take<bf16>(R * Dv);
    y.q = cv.take<bf16>(R * I);
    y.kv = cv.take<bf16>(KV * 2 * I);
    y.o = cv.take<bf16>(R * I);
    y.x_mid = cv.take<float>(R * Dv);
    y.xn2 = cv.take<bf16>(R * Dv);
    y.h_pre = cv.take<bf16>(R * FF);
    y.h_act = cv.take<bf16>(R * FF);
    y.mean_l = cv.take<float>(R); y.rstd_l = cv.take<float>(R);
    y.mean2 = cv.take<float>(R); y.rstd2 = cv.take<float>(R);
    y.lse = cv.take<float>((size_t)c->BN * c->heads * 64);
  }
  s.bytes = cv.off;
  return s;
}
struct RScratch {
  bf16 *dx_a, *dx_b, *dx_mid, *dh, *dxn2, *d_o, *dq, *dkv, *dkv_in, *dlat_q;
  float* dmedia;
  int* flags;
  void* ln_part[3];   // FFW norm / media norm / latent norm of a layer (reused after the per-layer ss.join())
  size_t bytes;
};
static RScratch carve_rscratch(const fm_resampler_cfg* c, void* p) {
  const size_t R = (size_t)c->BN * 64, Mm = (size_t)c->BN * c->T * c->F, nk = (size_t)c->T * c->F + 64, KV = (size_t)c->BN * nk;
  const size_t Dv = c->Dv, FF = c->ff_inner, I = c->heads * 64;
  Carver cv(p);
  RScratch s;
  s.dx_a = cv.take<bf16>(R * Dv); s.dx_b = cv.take<bf16>(R * Dv); s.dx_mid = cv.take<bf16>(R * Dv);
  s.dh = cv.take<bf16>(R * FF);
  s.dxn2 = cv.take<bf16>(R * Dv);
  s.d_o = cv.take<bf16>(R * I);
  s.dq = cv.take<bf16>(R * I);
  s.dkv = cv.take<bf16>(KV * 2 * I);
  s.dkv_in = cv.take<bf16>(KV * Dv);
  s.dlat_q = cv.take<bf16>(R * Dv);
  s.dmedia = cv.take<float>(Mm * Dv);
  s.flags = cv.take<int>(SPLITK_FLAG_INTS);
  for (int i = 0; i < 3; ++i) s.ln_part[i] = cv.take<char>(ln_part_bytes(c->Dv));
  s.bytes = cv.off;
  return s;
}
extern "C" size_t fm_resampler_saved_bytes(const fm_resampler_cfg* c) {
  return (check_res_cfg(c) == FM_OK && c->depth <= 16) ? carve_rsaved(c, nullptr).bytes : 0;
}
extern "C" size_t fm_resampler_scratch_bytes(const fm_resampler_cfg* c) { return check_res_cfg(c) == FM_OK ? carve_rscratch(c, nullptr).bytes : 0; }

extern "C" int fm_resampler_fwd(const fm_resampler_cfg* c, const float* wf, const void* wb_, const void* x_f, void* out, int out_f32,
                                void* saved, fm_stream_t stream) {
  ApiScope api_scope("r/");
  FM_TRY(check_res_cfg(c));
  FM_TRY(device_init());
  if (c->depth > 16) return fail(FM_EINVAL, "resampler depth %d > 16 not supported", c->depth);
  if (!wf || !wb_ || !x_f || !out || !saved) return fail(FM_EINVAL, "fm_resampler_fwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  fm_resampler_layout L;
  FM_TRY(fm_resampler_layout_of(c, &L));
  const bf16* wb = (const bf16*)wb_;
  const int R = c->BN * 64, TF = c->T * c->F, Mm = c->BN * TF, nk = TF + 64, KV = c->BN * nk, Dv = c->Dv, FF = c->ff_inner, I = c->heads * 64;
  RSaved sv = carve_rsaved(c, saved);

  // x0 = latents repeated over the batch                                      perceiver_resampler.py:179
  {
    const long long n4 = (long long)R * (Dv / 4);
    ProfScope ps("bcast_rows", 0.0, 4.0 * R * Dv, s);
    (void)launch_k(bcast_rows_kernel, (unsigned)((n4 + 255) / 256), 256, 0, s, wf + L.latents, sv.x[0], R, Dv, 64);
    KERNEL_CHECK();
  }
  for (int l = 0; l < c->depth; ++l) {
    const long long lb = L.layer0 + (long long)l * L.layer_stride;
    const float* wfl = wf + lb;
    const bf16* wbl = wb + lb;
    RLayerSaved& y = sv.layer[l];
    // media rows: LN_media(x_f + time_pos_emb) -> kv_in[bn, 0:TF]               :166, :52, :65
    {
      LnArgs a = mk_ln(x_f, c->x_f32, wfl + L.norm_media_w, wfl + L.norm_media_b, y.kv_in, 0, sv.mean_m, sv.rstd_m, Mm, Dv);
      a.add = wf + L.time_pos_emb; a.add_period = TF; a.add_group = c->F;
      a.in_group = TF; a.out_group = nk; a.out_off = 0;
      FM_TRY(run_ln_fwd(a, s));
    }
    // latent rows: LN_latents(x) -> kv_in[bn, TF:TF+64] and compact copy         :53, :65
    {
      LnArgs a = mk_ln(sv.x[l], 1, wfl + L.norm_latents_w, wfl + L.norm_latents_b, y.kv_in, 0, y.mean_l, y.rstd_l, R, Dv);
      a.in_group = 64; a.out_group = nk; a.out_off = TF; a.out2 = y.lat_n;
      FM_TRY(run_ln_fwd(a, s));
    }
    // q = (lat_n Wq^T) * dim_head^-0.5                                           :57, :79
    // [k | v] = kv_in [Wk ; Wv]^T                                                :69-70   (one grouped launch with q)
    {
      fm_gemm_desc grp[2];
      grp[0] = mk_gemm(R, I, Dv, y.lat_n, Dv, 0, wbl + L.to_q, Dv, 0, EPI_STORE, y.q, I, 0);
      grp[0].scale = 0.125f;
      grp[1] = mk_gemm(KV, 2 * I, Dv, y.kv_in, Dv, 0, wbl + L.to_k, Dv, 0, EPI_STORE, y.kv, 2 * I, 0);
      FM_TRY(run_gemm_group(grp, 2, s));
    }
    // softmax(q k^T) v                                                           :85-95   (tcgen05: attn_tc.cuh)
    {
      const cudaError_t aerr = ensure_dyn_smem(reinterpret_cast<const void*>(resampler_core_fwd_tc_kernel), XTC_FWD_SMEM);
      if (aerr != cudaSuccess) return fail(FM_ECUDA, "cudaFuncSetAttribute(resampler_core_fwd_tc) failed: %s", cudaGetErrorString(aerr));
      CUtensorMap tmQ, tmKV;
      FM_TRY(make_tmap_2d(&tmQ, y.q, I, R, I, 64, 128));
      FM_TRY(make_tmap_2d(&tmKV, y.kv, 2 * I, KV, 2 * I, 64, 64));
      RTcArgs a;
      a.o = y.o; a.lse = y.lse; a.BN = c->BN; a.H = c->heads; a.nk = nk;
      ProfScope ps("resampler_core_fwd", 4.0 * R * nk * I, 2.0 * (2.0 * R * I + 2.0 * KV * I), s);
      (void)launch_k(resampler_core_fwd_tc_kernel, dim3(c->heads, c->BN), 128, XTC_FWD_SMEM, s, tmQ, tmKV, a);
      KERNEL_CHECK();
    }
    // x_mid = x + o Wout^T                                                       :96, :182
    {
      fm_gemm_desc g = mk_gemm(R, Dv, I, y.o, I, 0, wbl + L.to_out, I, 0, EPI_RESID, y.x_mid, Dv, 1);
      g.aux = sv.x[l]; g.ldaux = Dv; g.aux_f32 = 1;
      FM_TRY(run_gemm(g, s));
    }
    // x_next = x_mid + FFW(x_mid)                                                :183, utils.py:45-50
    FM_TRY(run_ln_fwd(mk_ln(y.x_mid, 1, wfl + L.ffw_norm_w, wfl + L.ffw_norm_b, y.xn2, 0, y.mean2, y.rstd2, R, Dv), s));
    {
      fm_gemm_desc g = mk_gemm(R, FF, Dv, y.xn2, Dv, 0, wbl + L.ffw_w1, Dv, 0, EPI_ACT, y.h_act, FF, 0);
      g.out2 = c->training ? y.h_pre : nullptr; g.ldo2 = FF; g.act = c->act;
      FM_TRY(run_gemm(g, s));
    }
    {
      fm_gemm_desc g = mk_gemm(R, Dv, FF, y.h_act, FF, 0, wbl + L.ffw_w2, FF, 0, EPI_RESID, sv.x[l + 1], Dv, 1);
      g.aux = y.x_mid; g.ldaux = Dv; g.aux_f32 = 1;
      FM_TRY(run_gemm(g, s));
    }
  }
  // final norm                                                                  :187
  FM_TRY(run_ln_fwd(mk_ln(sv.x[c->depth], 1, wf + L.norm_w, wf + L.norm_b, out, out_f32, sv.mean_f, sv.rstd_f, R, Dv), s));
  return FM_OK;
}

static int resampler_bwd_impl(const fm_resampler_cfg* c, const float* wf, const void* wb_, const void* x_f, const void* saved,
                              const void* dout, float* gf, void* scratch, fm_layer_cb layer_done, void* user, fm_stream_t stream);
extern "C" int fm_resampler_bwd(const fm_resampler_cfg* c, const float* wf, const void* wb_, const void* x_f, const void* saved,
                                const void* dout, float* gf, void* scratch, fm_stream_t stream) {
  ApiScope api_scope("r/");
  return resampler_bwd_impl(c, wf, wb_, x_f, saved, dout, gf, scratch, nullptr, nullptr, stream);
}
extern "C" int fm_resampler_bwd_notify(const fm_resampler_cfg* c, const float* wf, const void* wb_, const void* x_f, const void* saved,
                                       const void* dout, float* gf, void* scratch, fm_layer_cb layer_done, void* user, fm_stream_t stream) {
  ApiScope api_scope("r/");
  return resampler_bwd_impl(c, wf, wb_, x_f, saved, dout, gf, scratch, layer_done, user, stream);
}
static int resampler_bwd_impl(const fm_resampler_cfg* c, const float* wf, const void* wb_, const void* x_f, const void* saved,
                              const void* dout, float* gf, void* scratch, fm_layer_cb layer_done, void* user, fm_stream_t stream) {
  FM_TRY(check_res_cfg(c));
  FM_TRY(device_init());
  if (c->depth > 16) return fail(FM_EINVAL, "resampler depth %d > 16 not supported", c->depth);
  if (!wf || !wb_ || !x_f || !saved || !dout || !gf || !scratch) return fail(FM_EINVAL, "fm_resampler_bwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  fm_resampler_layout L;
  FM_TRY(fm_resampler_layout_of(c, &L));
  const bf16* wb = (const bf16*)wb_;
  const int R = c->BN * 64, TF = c->T * c->F, Mm = c->BN * TF, nk = TF + 64, KV = c->BN * nk, Dv = c->Dv, FF = c->ff_inner, I = c->heads * 64;
  RSaved sv = carve_rsaved(c, const_cast<void*>(saved));
  RScratch sc = carve_rscratch(c, scratch);

  const cudaError_t aerr = ensure_dyn_smem(reinterpret_cast<const void*>(resampler_core_bwd_tc_kernel), XTC_BWD_SMEM);
  if (aerr != cudaSuccess) return fail(FM_ECUDA, "cudaFuncSetAttribute(resampler_core_bwd) failed: %s", cudaGetErrorString(aerr));

  CU_TRY(cudaMemsetAsync(sc.flags, 0, SPLITK_FLAG_INTS * sizeof(int), s)); note_other(s);
  bf16* dx_cur = sc.dx_a;
  bf16* dx_nxt = sc.dx_b;
  SideStream ss(s);
  cudaStream_t s2 = ss.ok ? ss.side : s;     // weight-gradient GEMMs run on the side stream (see SideStream)
  // final norm backward
  FM_TRY(run_ln_bwd(mk_ln_bwd(dout, sv.x[c->depth], 1, wf + L.norm_w, sv.mean_f, sv.rstd_f, nullptr, 0, dx_cur, 0, sc.ln_part[0], R, Dv),
                    gf + L.norm_w, gf + L.norm_b, s, &ss));
  for (int l = c->depth - 1; l >= 0; --l) {
    const long long lb = L.layer0 + (long long)l * L.layer_stride;
    const float* wfl = wf + lb;
    const bf16* wbl = wb + lb;
    float* gl = gf + lb;
    const RLayerSaved& y = sv.layer[l];
    FM_TRY(ss.join());          // the scratch buffers are reused per layer: last layer's dW GEMMs must have read them
    if (layer_done && l + 1 < c->depth) {          // layer l+1's gradients are now ordered before anything enqueued on s
      layer_done(user, l + 1);
      g_pdl.reset();                               // the callback may have enqueued foreign work on s
    }
    // ---- FFW backward
    {
      fm_gemm_desc g = mk_gemm(R, FF, Dv, dx_cur, Dv, 0, wbl + L.ffw_w2, FF, 1, EPI_DACT, sc.dh, FF, 0);
      g.aux = y.h_pre; g.ldaux = FF; g.act = c->act;
      FM_TRY(run_gemm(g, s));
    }
    FM_TRY(ss.fork());
    {
      int bn2 = 0, bn1 = 0;
      const int sp2 = plan_dw(Dv, FF, R, &bn2), sp1 = plan_dw(FF, Dv, R, &bn1);
      if (sp2 < 0 || sp1 < 0) { CU_TRY(cudaMemsetAsync(gl + L.ffw_w1, 0, sizeof(float) * 2 * (size_t)FF * Dv, s2)); note_other(s2); }
      fm_gemm_desc g2 = mk_gemm(Dv, FF, R, dx_cur, Dv, 1, y.h_act, FF, 1, EPI_STORE, gl + L.ffw_w2, FF, 1, sc.flags);
      if (sp2 < 0) { g2.splits = sp2; g2.bn = bn2; }
      FM_TRY(run_gemm(g2, s2));
      fm_gemm_desc g1 = mk_gemm(FF, Dv, R, sc.dh, FF, 1, y.xn2, Dv, 1, EPI_STORE, gl + L.ffw_w1, Dv, 1, sc.flags);
      if (sp1 < 0) { g1.splits = sp1; g1.bn = bn1; }
      FM_TRY(run_gemm(g1, s2));
    }
    FM_TRY(run_gemm(mk_gemm(R, Dv, FF, sc.dh, FF, 0, wbl + L.ffw_w1, Dv, 1, EPI_STORE, sc.dxn2, Dv, 0), s));
    FM_TRY(run_ln_bwd(mk_ln_bwd(sc.dxn2, y.x_mid, 1, wfl + L.ffw_norm_w, y.mean2, y.rstd2, dx_cur, 0, sc.dx_mid, 0, sc.ln_part[0], R, Dv),
                      gl + L.ffw_norm_w, gl + L.ffw_norm_b, s, &ss));
    // ---- attention backward
    FM_TRY(run_gemm(mk_gemm(R, I, Dv, sc.dx_mid, Dv, 0, wbl + L.to_out, I, 1, EPI_STORE, sc.d_o, I, 0), s));
    {
      CUtensorMap tmQ, tmDO, tmKV;
      FM_TRY(make_tmap_2d(&tmQ, y.q, I, R, I, 64, 128));
      FM_TRY(make_tmap_2d(&tmDO, sc.d_o, I, R, I, 64, 128));
      FM_TRY(make_tmap_2d(&tmKV, y.kv, 2 * I, KV, 2 * I, 64, 64));
      RTcBwdArgs a;
      a.o = y.o; a.d_o = sc.d_o; a.lse = y.lse; a.dq = sc.dq; a.dkv = sc.dkv; a.q_scale = 0.125f; a.BN = c->BN; a.H = c->heads; a.nk = nk;
      a.tmem_compact = opt(FM_OPT_ATTN_TMEM_COMPACT);
      ProfScope ps("resampler_core_bwd", 10.0 * R * nk * I, 2.0 * (4.0 * R * I + 4.0 * KV * I), s);
      (void)launch_k(resampler_core_bwd_tc_kernel, dim3(c->heads, c->BN), 128, XTC_BWD_SMEM, s, tmQ, tmDO, tmKV, a);
      KERNEL_CHECK();
    }
    FM_TRY(ss.fork());
    {   // dWout, dWq, dW[k|v] of this layer in one grouped launch
      fm_gemm_desc grp[3];
      grp[0] = mk_gemm(Dv, I, R, sc.dx_mid, Dv, 1, y.o, I, 1, EPI_STORE, gl + L.to_out, I, 1);
      grp[1] = mk_gemm(I, Dv, R, sc.dq, I, 1, y.lat_n, Dv, 1, EPI_STORE, gl + L.to_q, Dv, 1);
      grp[2] = mk_gemm(2 * I, Dv, KV, sc.dkv, 2 * I, 1, y.kv_in, Dv, 1, EPI_STORE, gl + L.to_k, Dv, 1);
      if (opt(FM_OPT_GEMM_GROUP) && plan_dw_group(grp, 3)) {     // to_q, to_k, to_v, to_out are adjacent in a layer's slice
        CU_TRY(cudaMemsetAsync(gl + L.to_q, 0, sizeof(float) * (size_t)(L.ffw_norm_w - L.to_q), s2)); note_other(s2);
      }
      FM_TRY(run_gemm_group(grp, 3, s2));
    }
    {   // dlat_q = dq Wq;  dkv_in = dkv [Wk ; Wv]   (independent: one grouped launch)
      fm_gemm_desc grp[2];
      grp[0] = mk_gemm(R, Dv, I, sc.dq, I, 0, wbl + L.to_q, Dv, 1, EPI_STORE, sc.dlat_q, Dv, 0);
      grp[1] = mk_gemm(KV, Dv, 2 * I, sc.dkv, 2 * I, 0, wbl + L.to_k, Dv, 1, EPI_STORE, sc.dkv_in, Dv, 0);
      FM_TRY(run_gemm_group(grp, 2, s));
    }
    // media rows: only parameter gradients survive, plus d(x_f + time_pos_emb) accumulated over layers for d(time_pos_emb)
    {
      LnBwdArgs a = mk_ln_bwd(sc.dkv_in, x_f, c->x_f32, wfl + L.norm_media_w, sv.mean_m, sv.rstd_m,
                              (l == c->depth - 1) ? nullptr : sc.dmedia, 1, sc.dmedia, 1, sc.ln_part[1], Mm, Dv);
      a.add = wf + L.time_pos_emb; a.add_period = TF; a.add_group = c->F;
      a.in_group = TF; a.out_group = nk; a.out_off = 0;
      FM_TRY(run_ln_bwd(a, gl + L.norm_media_w, gl + L.norm_media_b, s, &ss));
    }
    // latent rows: dx = dx_mid + LNbwd(dkv_in[latent rows] + dlat_q)
    {
      LnBwdArgs a = mk_ln_bwd(sc.dkv_in, sv.x[l], 1, wfl + L.norm_latents_w, y.mean_l, y.rstd_l, sc.dx_mid, 0, dx_nxt, 0, sc.ln_part[2], R, Dv);
      a.dy2 = sc.dlat_q;
      a.in_group = 64; a.out_group = nk; a.out_off = TF;
      FM_TRY(run_ln_bwd(a, gl + L.norm_latents_w, gl + L.norm_latents_b, s, &ss));
    }
    bf16* t = dx_cur; dx_cur = dx_nxt; dx_nxt = t;
  }
  FM_TRY(ss.join());
  if (layer_done) { layer_done(user, 0); g_pdl.reset(); }
  // d(latents)[i] = sum_bn dx0[bn, i];  d(time_pos_emb)[t] = sum_{bn, f} dmedia[bn, t, f]
  CU_TRY(cudaMemsetAsync(gf + L.latents, 0, sizeof(float) * (size_t)(L.layer0 - L.latents), s)); note_other(s);
  {
    const int rpb = 64;
    {
      ProfScope ps("group_rowsum", 0.0, 2.0 * R * Dv, s);
      (void)launch_k(group_rowsum_kernel, dim3((Dv + 127) / 128, (R + rpb - 1) / rpb), 128, 0, s, dx_cur, 0, R, Dv, 64, 1, gf + L.latents, rpb);
    }
    KERNEL_CHECK();
    {
      ProfScope ps("group_rowsum", 0.0, 4.0 * Mm * Dv, s);
      (void)launch_k(group_rowsum_kernel, dim3((Dv + 127) / 128, (Mm + rpb - 1) / rpb), 128, 0, s, sc.dmedia, 1, Mm, Dv, TF, c->F, gf + L.time_pos_emb, rpb);
    }
    KERNEL_CHECK();
  }
  return FM_OK;
}
