// The CUDA-core kernels (csrc/: LayerNorm fwd/bwd/reduce, loss head, misc) executed on the HOST,
// thread per thread, by tests/cpu_harness/simt_emu.h and compared with straightforward double-precision loops.
// These kernels were (re)written after round 1's GPU budget was spent: this runs their index arithmetic, row
// pipelining, shuffles, shared-memory folds and edge handling before they reach a B200.  Built with g++ -DFM_HOST_EMU
// (tests/test_simt_emu_cpu.py); the grid / TPR selection below restates the launchers in csrc/flamingo_b200.cu.
// TEST INFRASTRUCTURE ONLY.
#include "simt_emu.h"

#include "../../flamingo_mini_b200/csrc/layernorm.cuh"
#include "../../flamingo_mini_b200/csrc/loss.cuh"
#include "../../flamingo_mini_b200/csrc/misc.cuh"

#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

using namespace fm;
using bf16 = __nv_bfloat16;

static std::mt19937 rng(1234);
static float frand(float s = 1.0f) { return std::normal_distribution<float>(0.0f, s)(rng); }
static float bfr(float x) { return __bfloat162float(__float2bfloat16(x)); }
static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); ++g_fail; } } while (0)

struct Err { double num = 0, den = 0; void add(double got, double want) { num += (got - want) * (got - want); den += want * want; }
             double rel() const { return std::sqrt(num / (den + 1e-30)); } };

// launcher restatement (csrc/flamingo_b200.cu: ln_maxc / ln_tpr / ln_grid), with the SM count as a parameter so that
// both the "one row iteration" and the "several pipelined iterations" regimes are exercised
static int ln_maxc(int D) { return D <= 4096 ? 2 : LN_MAXC_WIDE; }
static int ln_tpr(int D) { const int need = (D / 8 + ln_maxc(D) - 1) / ln_maxc(D); return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : 256; }
static int ln_grid(int rows, int tpr, int cap) { const int rpc = LN_THREADS / tpr; const int want = (rows + rpc - 1) / rpc; return want < cap ? want : cap; }

template <typename A, typename K32, typename K64, typename K128, typename K256, typename KW>
static void dispatch(int D, int grid, size_t smem, const A& a, K32 k32, K64 k64, K128 k128, K256 k256, KW kw) {
  const int tpr = ln_tpr(D);
  if (ln_maxc(D) == 2) {
    switch (tpr) {
      case 32: emu::launch(grid, LN_THREADS, smem, [&] { k32(a); }); break;
      case 64: emu::launch(grid, LN_THREADS, smem, [&] { k64(a); }); break;
      case 128: emu::launch(grid, LN_THREADS, smem, [&] { k128(a); }); break;
      default: emu::launch(grid, LN_THREADS, smem, [&] { k256(a); }); break;
    }
  } else {
    emu::launch(grid, LN_THREADS, smem, [&] { kw(a); });
  }
}

// warp-per-row kernels (D <= LN_WARP_MAX_D): restates launch_lnw / lnw_grid of csrc/flamingo_b200.cu
static int lnw_maxc(int D) { return (D / 8 + 31) / 32; }
static int lnw_grid(int rows, int cap) { const int want = (rows + 7) / 8; return want < cap ? want : cap; }
template <typename A, typename K1, typename K2, typename K3, typename K4, typename K6>
static void dispatch_w(int D, int grid, size_t smem, const A& a, K1 k1, K2 k2, K3 k3, K4 k4, K6 k6) {
  switch (lnw_maxc(D)) {
    case 1: emu::launch(grid, LN_THREADS, smem, [&] { k1(a); }); break;
    case 2: emu::launch(grid, LN_THREADS, smem, [&] { k2(a); }); break;
    case 3: emu::launch(grid, LN_THREADS, smem, [&] { k3(a); }); break;
    case 4: emu::launch(grid, LN_THREADS, smem, [&] { k4(a); }); break;
    default: emu::launch(grid, LN_THREADS, smem, [&] { k6(a); }); break;
  }
}

struct LnCase { int rows, D, x_f32, out_f32, with_add, scatter, with_out2, cap; };

static void test_ln_fwd(const LnCase& c) {
  const int rows = c.rows, D = c.D;
  // scatter: rows come in groups of in_group and land at (g * out_group + out_off + r) — [media ; latents] layout
  const int in_group = c.scatter ? 5 : rows, out_group = c.scatter ? 9 : rows, out_off = c.scatter ? 3 : 0;
  const int out_rows = c.scatter ? ((rows + in_group - 1) / in_group) * out_group : rows;
  const int add_period = c.with_add ? 6 : 1, add_group = c.with_add ? 3 : 1;       // 2 embedding rows ("frames" of 3 rows)
  std::vector<float> xf(static_cast<size_t>(rows) * D), add(2 * D), gamma(D), beta(D);
  std::vector<bf16> xb(xf.size());
  for (size_t i = 0; i < xf.size(); ++i) { xf[i] = frand() + 0.3f; xb[i] = __float2bfloat16(xf[i]); if (!c.x_f32) xf[i] = __bfloat162float(xb[i]); }
  for (auto& v : add) v = frand(0.5f);
  for (int d = 0; d < D; ++d) { gamma[d] = 1.0f + frand(0.2f); beta[d] = frand(0.1f); }
  std::vector<float> outf(static_cast<size_t>(out_rows) * D, -777.0f), mean(rows, -1), rstd(rows, -1);
  std::vector<bf16> outb(outf.size(), __float2bfloat16(-777.0f)), out2(static_cast<size_t>(rows) * D, __float2bfloat16(-777.0f));
  LnArgs a;
  memset(&a, 0, sizeof(a));
  a.x = c.x_f32 ? static_cast<const void*>(xf.data()) : static_cast<const void*>(xb.data());
  a.x_f32 = c.x_f32;
  a.add = c.with_add ? add.data() : nullptr; a.add_period = add_period; a.add_group = add_group;
  a.gamma = gamma.data(); a.beta = beta.data();
  a.out = c.out_f32 ? static_cast<void*>(outf.data()) : static_cast<void*>(outb.data());
  a.out_f32 = c.out_f32; a.in_group = in_group; a.out_group = out_group; a.out_off = out_off;
  a.out2 = c.with_out2 ? out2.data() : nullptr;
  a.mean = mean.data(); a.rstd = rstd.data(); a.rows = rows; a.D = D;
  int grid = ln_grid(rows, ln_tpr(D), c.cap);
  if (D <= LN_WARP_MAX_D) {
    grid = lnw_grid(rows, c.cap);
    dispatch_w(D, grid, 0, a, ln_fwd_w_kernel<1>, ln_fwd_w_kernel<2>, ln_fwd_w_kernel<3>, ln_fwd_w_kernel<4>, ln_fwd_w_kernel<6>);
  } else {
    dispatch(D, grid, 0, a, ln_fwd_kernel<32, 2>, ln_fwd_kernel<64, 2>, ln_fwd_kernel<128, 2>, ln_fwd_kernel<256, 2>,
             ln_fwd_kernel<256, LN_MAXC_WIDE>);
  }
  Err e, e2;
  std::vector<char> written(out_rows, 0);
  for (int r = 0; r < rows; ++r) {
    std::vector<double> v(D);
    double m = 0, s = 0;
    for (int d = 0; d < D; ++d) { v[d] = xf[static_cast<size_t>(r) * D + d] + (c.with_add ? add[static_cast<size_t>((r % add_period) / add_group) * D + d] : 0.0f); m += v[d]; }
    m /= D;
    for (int d = 0; d < D; ++d) s += (v[d] - m) * (v[d] - m);
    const double rs = 1.0 / std::sqrt(s / D + 1e-5);
    CHECK(std::fabs(mean[r] - m) < 1e-4 && std::fabs(rstd[r] - rs) < 1e-3 * rs, "ln_fwd stats row %d: %g %g vs %g %g", r, mean[r], rstd[r], m, rs);
    const size_t orow = static_cast<size_t>(r / in_group) * out_group + out_off + r % in_group;
    written[orow] = 1;
    for (int d = 0; d < D; ++d) {
      const double want = (v[d] - m) * rs * gamma[d] + beta[d];
      e.add(c.out_f32 ? outf[orow * D + d] : __bfloat162float(outb[orow * D + d]), want);
      if (c.with_out2) e2.add(__bfloat162float(out2[static_cast<size_t>(r) * D + d]), want);
    }
  }
  for (int r = 0; r < out_rows; ++r)           // rows outside the map must stay untouched
    if (!written[r])
      for (int d = 0; d < D; d += 7) {
        const float got = c.out_f32 ? outf[static_cast<size_t>(r) * D + d] : __bfloat162float(outb[static_cast<size_t>(r) * D + d]);
        CHECK(std::fabs(got + 777.0f) < 2.0f, "ln_fwd wrote an unmapped row %d", r);
      }
  const double tol = c.out_f32 ? 2e-6 : 4e-3;
  CHECK(e.rel() < tol, "ln_fwd rows=%d D=%d grid=%d: rel err %g", rows, D, grid, e.rel());
  if (c.with_out2) CHECK(e2.rel() < 4e-3, "ln_fwd out2 rows=%d D=%d: rel err %g", rows, D, e2.rel());
}

struct LnBwdCase { int rows, D, x_f32, with_add, scatter, with_dy2, dres /*0 none, 1 bf16, 2 f32*/, dx /*0 none, 1 bf16, 2 f32*/, cap; };

static void test_ln_bwd(const LnBwdCase& c) {
  const int rows = c.rows, D = c.D;
  const int in_group = c.scatter ? 5 : rows, out_group = c.scatter ? 9 : rows, out_off = c.scatter ? 3 : 0;
  const int out_rows = c.scatter ? ((rows + in_group - 1) / in_group) * out_group : rows;
  const int add_period = c.with_add ? 6 : 1, add_group = c.with_add ? 3 : 1;
  std::vector<float> xf(static_cast<size_t>(rows) * D), add(2 * D), gamma(D), mean(rows), rstd(rows), dresf(xf.size());
  std::vector<bf16> xb(xf.size()), dy(static_cast<size_t>(out_rows) * D), dy2(xf.size()), dresb(xf.size());
  for (size_t i = 0; i < xf.size(); ++i) { xf[i] = frand() + 0.3f; xb[i] = __float2bfloat16(xf[i]); if (!c.x_f32) xf[i] = __bfloat162float(xb[i]); }
  for (auto& v : add) v = frand(0.5f);
  for (int d = 0; d < D; ++d) gamma[d] = 1.0f + frand(0.2f);
  for (auto& v : dy) v = __float2bfloat16(frand());
  for (auto& v : dy2) v = __float2bfloat16(frand());
  for (size_t i = 0; i < dresf.size(); ++i) { dresf[i] = frand(); dresb[i] = __float2bfloat16(dresf[i]); if (c.dres == 1) dresf[i] = __bfloat162float(dresb[i]); }
  std::vector<std::vector<double>> v(rows, std::vector<double>(D));
  for (int r = 0; r < rows; ++r) {
    double m = 0, s = 0;
    for (int d = 0; d < D; ++d) { v[r][d] = xf[static_cast<size_t>(r) * D + d] + (c.with_add ? add[static_cast<size_t>((r % add_period) / add_group) * D + d] : 0.0f); m += v[r][d]; }
    m /= D;
    for (int d = 0; d < D; ++d) s += (v[r][d] - m) * (v[r][d] - m);
    mean[r] = static_cast<float>(m); rstd[r] = static_cast<float>(1.0 / std::sqrt(s / D + 1e-5));
  }
  const int tpr = ln_tpr(D);
  int grid = D <= LN_WARP_MAX_D ? lnw_grid(rows, c.cap) : ln_grid(rows, tpr, c.cap);
  if (grid > 448) grid = 448;
  std::vector<float> part(static_cast<size_t>(grid) * 2 * D, 1e30f), dxf(xf.size(), -777.0f), dgamma(D, 5.0f), dbeta(D, 5.0f);
  std::vector<bf16> dxb(xf.size());
  LnBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.dy = dy.data(); a.dy2 = c.with_dy2 ? dy2.data() : nullptr;
  a.in_group = in_group; a.out_group = out_group; a.out_off = out_off;
  a.x = c.x_f32 ? static_cast<const void*>(xf.data()) : static_cast<const void*>(xb.data()); a.x_f32 = c.x_f32;
  a.add = c.with_add ? add.data() : nullptr; a.add_period = add_period; a.add_group = add_group;
  a.gamma = gamma.data(); a.mean = mean.data(); a.rstd = rstd.data();
  a.dres = c.dres == 0 ? nullptr : c.dres == 1 ? static_cast<const void*>(dresb.data()) : static_cast<const void*>(dresf.data());
  a.dres_f32 = c.dres == 2;
  a.dx = c.dx == 0 ? nullptr : c.dx == 1 ? static_cast<void*>(dxb.data()) : static_cast<void*>(dxf.data());
  a.dx_f32 = c.dx == 2;
  a.part = part.data(); a.rows = rows; a.D = D;
  const size_t smem = tpr < LN_THREADS ? static_cast<size_t>(2) * D * sizeof(float) : 0;
  if (D <= LN_WARP_MAX_D) {
    if (c.dx) dispatch_w(D, lnw_grid(rows, 3 * c.cap), 0, a, ln_bwd_dx_w_kernel<1>, ln_bwd_dx_w_kernel<2>, ln_bwd_dx_w_kernel<3>, ln_bwd_dx_w_kernel<4>,
                         ln_bwd_dx_w_kernel<6>);
    dispatch_w(D, grid, static_cast<size_t>(2) * D * sizeof(float), a, ln_bwd_dgb_w_kernel<1>, ln_bwd_dgb_w_kernel<2>, ln_bwd_dgb_w_kernel<3>,
               ln_bwd_dgb_w_kernel<4>, ln_bwd_dgb_w_kernel<6>);
  } else {
    dispatch(D, grid, smem, a, ln_bwd_kernel<32, 2>, ln_bwd_kernel<64, 2>, ln_bwd_kernel<128, 2>, ln_bwd_kernel<256, 2>,
             ln_bwd_kernel<256, LN_MAXC_WIDE>);
  }
  emu::launch((2 * D + 31) / 32, 256, 0, [&] { ln_bwd_reduce_kernel(part.data(), grid, D, dgamma.data(), dbeta.data(), 0); });
  Err ex, eg, eb;
  std::vector<double> wg(D, 0.0), wb(D, 0.0);
  for (int r = 0; r < rows; ++r) {
    const size_t orow = static_cast<size_t>(r / in_group) * out_group + out_off + r % in_group;
    std::vector<double> g(D), xh(D);
    double m1 = 0, m2 = 0;
    for (int d = 0; d < D; ++d) {
      const double dyv = __bfloat162float(dy[orow * D + d]) + (c.with_dy2 ? __bfloat162float(dy2[static_cast<size_t>(r) * D + d]) : 0.0f);
      xh[d] = (v[r][d] - mean[r]) * rstd[r];
      g[d] = dyv * gamma[d];
      m1 += g[d]; m2 += g[d] * xh[d];
      wg[d] += dyv * xh[d]; wb[d] += dyv;
    }
    m1 /= D; m2 /= D;
    if (c.dx)
      for (int d = 0; d < D; ++d) {
        const double want = rstd[r] * (g[d] - m1 - xh[d] * m2) + (c.dres ? dresf[static_cast<size_t>(r) * D + d] : 0.0f);
        ex.add(c.dx == 2 ? dxf[static_cast<size_t>(r) * D + d] : __bfloat162float(dxb[static_cast<size_t>(r) * D + d]), want);
      }
  }
  for (int d = 0; d < D; ++d) { eg.add(dgamma[d], wg[d]); eb.add(dbeta[d], wb[d]); }
  if (c.dx) CHECK(ex.rel() < (c.dx == 2 ? 1e-5 : 4e-3), "ln_bwd dx rows=%d D=%d grid=%d: rel err %g", rows, D, grid, ex.rel());
  CHECK(eg.rel() < 1e-5 && eb.rel() < 1e-5, "ln_bwd dgamma/dbeta rows=%d D=%d grid=%d: rel err %g %g", rows, D, grid, eg.rel(), eb.rel());
}

static void test_ln_reduce_accumulate() {
  const int D = 40, nparts = 19;
  std::vector<float> part(static_cast<size_t>(nparts) * 2 * D), dg(D, 2.0f), db(D, -1.0f);
  for (auto& v : part) v = frand();
  emu::launch((2 * D + 31) / 32, 256, 0, [&] { ln_bwd_reduce_kernel(part.data(), nparts, D, dg.data(), db.data(), 1); });
  for (int d = 0; d < D; ++d) {
    double g = 2.0, b = -1.0;
    for (int p = 0; p < nparts; ++p) { g += part[static_cast<size_t>(p) * 2 * D + d]; b += part[static_cast<size_t>(p) * 2 * D + D + d]; }
    CHECK(std::fabs(dg[d] - g) < 1e-4 && std::fabs(db[d] - b) < 1e-4, "ln_bwd_reduce accumulate col %d", d);
  }
}

// ---------------------------------------------------------------------------------------------------- loss head
static void test_cross_entropy(int rows, int vocab, int ld, bool neg_inf_padding_inside) {
  std::vector<bf16> logits(static_cast<size_t>(rows) * ld), dl(logits.size(), __float2bfloat16(9.0f));
  std::vector<long long> tgt(rows);
  for (int r = 0; r < rows; ++r) {
    for (int c = 0; c < ld; ++c) logits[static_cast<size_t>(r) * ld + c] = __float2bfloat16(c < vocab ? frand(3.0f) + (r % 3) * 20.0f : 1e30f);   // padding is garbage
    tgt[r] = static_cast<long long>(rng() % vocab);
  }
  if (rows > 2) tgt[1] = -100;                                     // ignored row
  if (rows > 3) tgt[3] = vocab - 1;                                // last real column
  if (neg_inf_padding_inside && vocab > 4)                         // the aligned lm_head's -inf bias columns are INSIDE vocab
    for (int r = 0; r < rows; ++r) { logits[static_cast<size_t>(r) * ld + vocab - 2] = __float2bfloat16(-INFINITY); if (tgt[r] == vocab - 2) tgt[r] = 0; }
  std::vector<float> lse(rows, -1), row_loss(rows, -1), scale(1, 0.37f);
  CeArgs a;
  memset(&a, 0, sizeof(a));
  a.logits = logits.data(); a.targets = tgt.data(); a.ignore_index = -100; a.lse = lse.data(); a.row_loss = row_loss.data();
  a.dlogits = dl.data(); a.scale = scale.data(); a.rows = rows; a.vocab = vocab; a.ld = ld;
  emu::launch(rows, CE_THREADS, 0, [&] { ce_fwd_kernel(a); });
  emu::launch(rows, CE_THREADS, 0, [&] { ce_bwd_kernel(a); });
  for (int r = 0; r < rows; ++r) {
    const bf16* src = logits.data() + static_cast<size_t>(r) * ld;
    double mx = -INFINITY, s = 0;
    for (int c = 0; c < vocab; ++c) mx = std::max(mx, static_cast<double>(__bfloat162float(src[c])));
    for (int c = 0; c < vocab; ++c) s += std::exp(__bfloat162float(src[c]) - mx);
    const double want_lse = mx + std::log(s);
    CHECK(std::fabs(lse[r] - want_lse) < 2e-4 * std::max(1.0, std::fabs(want_lse)), "ce lse row %d: %g vs %g", r, lse[r], want_lse);
    const bool live = tgt[r] != -100;
    const double want_loss = live ? want_lse - __bfloat162float(src[tgt[r]]) : 0.0;
    CHECK(std::fabs(row_loss[r] - want_loss) < 2e-4 * std::max(1.0, std::fabs(want_lse)), "ce loss row %d: %g vs %g", r, row_loss[r], want_loss);
    Err e;
    for (int c = 0; c < ld; ++c) {
      double want = 0.0;
      if (live && c < vocab) want = (std::exp(__bfloat162float(src[c]) - want_lse) - (c == tgt[r] ? 1.0 : 0.0)) * 0.37;
      const float got = __bfloat162float(dl[static_cast<size_t>(r) * ld + c]);
      if (!live || c >= vocab) CHECK(got == 0.0f, "ce bwd row %d col %d must be exactly zero, got %g", r, c, got);
      e.add(got, want);
    }
    if (live) CHECK(e.rel() < 5e-3, "ce bwd row %d: rel err %g", r, e.rel());
  }
}

// ---------------------------------------------------------------------------------------------------- misc kernels
static void test_text_time(int B, int S) {
  std::vector<int> ml(static_cast<size_t>(B) * S), tt(ml.size(), -5);
  for (auto& v : ml) v = (rng() % 7 == 0);
  emu::launch((B + 3) / 4, 128, 0, [&] { text_time_kernel(ml.data(), tt.data(), B, S); });
  for (int b = 0; b < B; ++b) {
    int acc = 0;
    for (int i = 0; i < S; ++i) { acc += ml[static_cast<size_t>(b) * S + i]; CHECK(tt[static_cast<size_t>(b) * S + i] == acc, "text_time b=%d i=%d: %d vs %d", b, i, tt[static_cast<size_t>(b) * S + i], acc); }
  }
}

static void test_cast(long long n) {
  std::vector<float> src(n);
  std::vector<bf16> dst(n + 8, __float2bfloat16(-3.0f));
  for (auto& v : src) v = frand(4.0f);
  const long long threads = (n + 7) / 8;
  emu::launch(static_cast<unsigned>((threads + 255) / 256), 256, 0, [&] { cast_f32_bf16_kernel(src.data(), dst.data(), n); });
  for (long long i = 0; i < n; ++i) CHECK(__bfloat162float(dst[i]) == bfr(src[i]), "cast element %lld", i);
  for (long long i = n; i < n + 8; ++i) CHECK(__bfloat162float(dst[i]) == -3.0f, "cast wrote past the end (%lld)", i);
}

static void test_dot_reduce(long long n, int grid) {
  std::vector<bf16> a(n), b(n);
  double want = 0;
  for (long long i = 0; i < n; ++i) { a[i] = __float2bfloat16(frand()); b[i] = __float2bfloat16(frand()); want += static_cast<double>(__bfloat162float(a[i])) * __bfloat162float(b[i]); }
  float out = 1.5f;
  emu::launch(grid, 256, 0, [&] { dot_reduce_kernel(a.data(), b.data(), n, &out); });
  CHECK(std::fabs(out - (want + 1.5)) < 1e-3 * std::sqrt(static_cast<double>(n)), "dot_reduce n=%lld: %g vs %g", n, out, want + 1.5);
}

static void test_bcast_and_rowsum() {
  const int D = 24, period = 6, group = 3;
  const long long rows = 41;
  std::vector<float> src(static_cast<size_t>(period) * D), dst(static_cast<size_t>(rows) * D, -1.0f);
  for (auto& v : src) v = frand();
  const long long n4 = rows * (D / 4);
  emu::launch(static_cast<unsigned>((n4 + 127) / 128), 128, 0, [&] { bcast_rows_kernel(src.data(), dst.data(), rows, D, period); });
  for (long long r = 0; r < rows; ++r)
    for (int d = 0; d < D; ++d) CHECK(dst[r * D + d] == src[(r % period) * D + d], "bcast_rows r=%lld d=%d", r, d);
  for (int f32 = 0; f32 < 2; ++f32) {
    std::vector<float> xf(static_cast<size_t>(rows) * D), out(static_cast<size_t>(period / group) * D, 0.0f);
    std::vector<bf16> xb(xf.size());
    for (size_t i = 0; i < xf.size(); ++i) { xb[i] = __float2bfloat16(frand()); xf[i] = __bfloat162float(xb[i]); }
    const int rpb = 7;
    emu::launch(dim3((D + 31) / 32, static_cast<unsigned>((rows + rpb - 1) / rpb)), 32, 0, [&] {
      group_rowsum_kernel(f32 ? static_cast<const void*>(xf.data()) : static_cast<const void*>(xb.data()), f32, rows, D, period, group, out.data(), rpb);
    });
    for (int g = 0; g < period / group; ++g)
      for (int d = 0; d < D; ++d) {
        double want = 0;
        for (long long r = 0; r < rows; ++r) if ((r % period) / group == g) want += xf[r * D + d];
        CHECK(std::fabs(out[static_cast<size_t>(g) * D + d] - want) < 1e-4, "group_rowsum f32=%d g=%d d=%d: %g vs %g", f32, g, d, out[static_cast<size_t>(g) * D + d], want);
      }
  }
}

static void test_activations() {
  // the fast GELU / GELU' of the GEMM epilogues against erf, and the two rectifiers (utils.py:36-40)
  double worst_f = 0, worst_d = 0;
  for (float x = -9.0f; x <= 9.0f; x += 0.00317f) {
    float f;
    const float d = act_bwd_fast(x, 0, &f);
    const double cdf = 0.5 * (1.0 + std::erf(x * 0.7071067811865476)), pdf = 0.3989422804014327 * std::exp(-0.5 * x * x);
    worst_f = std::max(worst_f, std::fabs(f - x * cdf));
    worst_d = std::max(worst_d, std::fabs(d - (cdf + x * pdf)));
    CHECK(act_fwd_fast(x, 0) == f, "act_fwd_fast and act_bwd_fast disagree at %g", x);
    float f1, f2;
    CHECK(act_bwd_fast(x, 1, &f1) == 2.0f * std::max(x, 0.0f) && f1 == std::max(x, 0.0f) * std::max(x, 0.0f), "sqrelu at %g", x);
    CHECK(act_bwd_fast(x, 2, &f2) == (x > 0 ? 1.0f : 0.0f) && f2 == std::max(x, 0.0f), "relu at %g", x);
  }
  CHECK(worst_f < 2e-6 && worst_d < 2e-6, "fast GELU: max abs error %g (value) %g (derivative)", worst_f, worst_d);
  // packed (FFMA2) epilogue helpers: pairs of columns, same accuracy as the scalar evaluation, both lanes independent
  double worst2_f = 0, worst2_d = 0;
  for (float x = -9.0f; x <= 9.0f; x += 0.00731f) {
    const float y = 1.7f - 0.9f * x;                                         // an unrelated value in the other lane
    float2 f, d, f_only, unused;
    gelu2<true>(make_float2(x, y), f, d);
    gelu2<false>(make_float2(y, x), f_only, unused);
    CHECK(f_only.x == f.y && f_only.y == f.x, "gelu2 lanes are not independent at %g", x);
    for (int l = 0; l < 2; ++l) {
      const double v = l ? y : x, cdf = 0.5 * (1.0 + std::erf(v * 0.7071067811865476)), pdf = 0.3989422804014327 * std::exp(-0.5 * v * v);
      worst2_f = std::max(worst2_f, std::fabs((l ? f.y : f.x) - v * cdf));
      worst2_d = std::max(worst2_d, std::fabs((l ? d.y : d.x) - (cdf + v * pdf)));
    }
  }
  CHECK(worst2_f < 2e-6 && worst2_d < 2e-6, "packed GELU: max abs error %g (value) %g (derivative)", worst2_f, worst2_d);
  for (int act = 0; act < 3; ++act) {
    float v[32], w[32], d[32], c[32], ref_f[32], ref_d[32];
    for (int j = 0; j < 32; ++j) { v[j] = w[j] = frand(3.0f); c[j] = frand(); ref_d[j] = act_bwd_fast(v[j], act, &ref_f[j]); }
    if (act == 0) { act32<0, true>(v, d); act32<0, false>(w, w); }
    if (act == 1) { act32<1, true>(v, d); act32<1, false>(w, w); }
    if (act == 2) { act32<2, true>(v, d); act32<2, false>(w, w); }
    for (int j = 0; j < 32; ++j) {
      CHECK(std::fabs(v[j] - ref_f[j]) < 1e-6 && std::fabs(d[j] - ref_d[j]) < 1e-6 && w[j] == v[j], "act32 act=%d col %d: %g/%g vs %g/%g", act, j, v[j], d[j], ref_f[j], ref_d[j]);
    }
    float a[32], r[32];
    double dot = 0;
    for (int j = 0; j < 32; ++j) { a[j] = frand(); r[j] = a[j]; dot += static_cast<double>(a[j]) * c[j]; }
    CHECK(std::fabs(dot32(a, c) - dot) < 1e-5, "dot32: %g vs %g", dot32(a, c), dot);
    scale32(r, 0.37f);
    for (int j = 0; j < 32; ++j) CHECK(r[j] == a[j] * 0.37f, "scale32 col %d", j);
    for (int j = 0; j < 32; ++j) r[j] = a[j];
    axpy32(r, -1.3f, c);
    for (int j = 0; j < 32; ++j) CHECK(r[j] == fmaf(-1.3f, a[j], c[j]), "axpy32 col %d", j);
    for (int j = 0; j < 32; ++j) r[j] = a[j];
    scale_mul32(r, 0.6f, c);
    for (int j = 0; j < 32; ++j) CHECK(r[j] == 0.6f * a[j] * c[j], "scale_mul32 col %d", j);
  }
}

int main(int argc, char** argv) {
  const std::string only = argc > 1 ? argv[1] : "";
  auto want = [&](const char* n) { return only.empty() || only == n; };
  if (want("act")) test_activations();
  if (want("ln_fwd")) {
    // rows, D, x_f32, out_f32, add, scatter, out2, CTA cap (small cap => several pipelined iterations per CTA)
    const LnCase cases[] = {
        {37, 64, 0, 0, 0, 0, 0, 2},   {37, 64, 1, 1, 1, 1, 1, 100}, {23, 256, 1, 0, 1, 1, 0, 1},  {50, 768, 0, 0, 0, 0, 1, 3},
        {11, 768, 1, 1, 1, 1, 0, 2},  {9, 1024, 0, 0, 1, 0, 0, 1},  {7, 1280, 1, 0, 0, 1, 1, 2},  {5, 2048, 0, 1, 0, 0, 0, 1},
        {4, 4096, 1, 0, 0, 0, 0, 1},  {3, 8192, 0, 0, 1, 0, 0, 1},  {3, 4104, 1, 1, 0, 0, 0, 2},  {1, 8, 1, 1, 0, 0, 0, 4},
        {300, 128, 0, 0, 1, 1, 1, 5}, {12, 1536, 0, 0, 0, 0, 0, 2}, {30, 768, 0, 0, 0, 0, 1, 1}, {6, 1544, 1, 0, 0, 0, 0, 1},
    };
    for (const auto& c : cases) test_ln_fwd(c);
  }
  if (want("ln_bwd")) {
    // rows, D, x_f32, add, scatter, dy2, dres, dx, CTA cap
    const LnBwdCase cases[] = {
        {37, 64, 0, 0, 0, 0, 0, 1, 2},  {37, 64, 1, 1, 1, 1, 2, 2, 100}, {23, 256, 1, 1, 1, 0, 1, 1, 1}, {50, 768, 0, 0, 0, 1, 2, 2, 3},
        {11, 768, 1, 1, 1, 1, 0, 0, 2}, {9, 1024, 0, 0, 0, 0, 1, 2, 1},  {7, 1280, 1, 0, 1, 1, 2, 1, 2}, {5, 2048, 0, 0, 0, 0, 0, 2, 1},
        {4, 4096, 1, 0, 0, 0, 2, 2, 1}, {3, 8192, 0, 1, 0, 0, 0, 2, 1},  {3, 4104, 1, 0, 0, 1, 1, 2, 2}, {1, 8, 1, 0, 0, 0, 0, 2, 4},
        {130, 128, 0, 1, 1, 1, 2, 1, 3}, {12, 1536, 0, 0, 0, 0, 1, 1, 2}, {30, 768, 0, 0, 0, 0, 1, 1, 1}, {6, 1544, 1, 0, 0, 0, 0, 2, 1},
    };
    for (const auto& c : cases) test_ln_bwd(c);
    test_ln_reduce_accumulate();
  }
  if (want("ce")) {
    test_cross_entropy(6, 50258, 50304, true);
    test_cross_entropy(5, 1000, 1000, false);
    test_cross_entropy(4, 515, 520, false);
    test_cross_entropy(5, 8, 8, false);
    test_cross_entropy(3, 8200, 8208, true);      // exactly one trip of the 4-deep loop plus a ragged tail
  }
  if (want("misc")) {
    test_text_time(5, 70); test_text_time(1, 32); test_text_time(9, 1);
    test_cast(8 * 300); test_cast(8 * 300 + 5); test_cast(3);
    test_dot_reduce(8 * 1000, 3); test_dot_reduce(8 * 17, 1);
    test_bcast_and_rowsum();
  }
  if (g_fail) { printf("SIMT EMU FAILED: %d checks\n", g_fail); return 1; }
  printf("SIMT EMU OK\n");
  return 0;
}
