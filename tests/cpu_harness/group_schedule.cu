// Host-side check of the persistent GEMM's unit schedule (csrc/gemm_tc.cuh: locate_unit / tile_coords): for a set of
// problem groups, every (problem, split, output tile) must be visited exactly once by the units 0..num_units-1, K ranges of
// the splits of a tile must partition [0, num_kb), and the three warp roles (which call locate_unit independently) see
// the same sequence by construction.  Built and run by tests/test_group_schedule_cpu.py (no GPU needed).
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "../../flamingo_mini_b200/csrc/gemm_tc.cuh"

using namespace fm;

template <int BN>
static int check_group(const std::vector<std::tuple<int, int, int, int>>& probs /* M, N, K, splits */) {
  GemmGroup G;
  memset(&G, 0, sizeof(G));
  G.nprob = (int)probs.size();
  int units = 0;
  for (int i = 0; i < G.nprob; ++i) {
    G.g[i].M = std::get<0>(probs[i]); G.g[i].N = std::get<1>(probs[i]); G.g[i].K = std::get<2>(probs[i]);
    G.g[i].splits = std::get<3>(probs[i]);
    G.unit_start[i] = units;
    const int sp = G.g[i].splits > 1 ? G.g[i].splits : 1;
    units += ((G.g[i].M + GEMM_BM - 1) / GEMM_BM) * ((G.g[i].N + BN - 1) / BN) * sp;
  }
  G.unit_start[G.nprob] = units;
  std::map<std::tuple<int, int, int>, std::vector<std::pair<int, int>>> seen;   // (p, mb, nb) -> K ranges
  for (int u = 0; u < units; ++u) {
    const UnitInfo a = G.nprob > 1 ? locate_unit<BN, true>(G, u) : locate_unit<BN, true>(G, u);
    if (G.nprob == 1) {      // the statically addressed variant must agree with the grouped one
      const UnitInfo b = locate_unit<BN, false>(G, u);
      if (a.p != b.p || a.mb != b.mb || a.nb != b.nb || a.kb_begin != b.kb_begin || a.kb_end != b.kb_end || a.split != b.split || a.tile != b.tile) {
        printf("FAIL static/grouped mismatch at unit %d\n", u);
        return 1;
      }
    }
    if (a.p < 0 || a.p >= G.nprob) { printf("FAIL unit %d: problem %d\n", u, a.p); return 1; }
    const int num_mb = (G.g[a.p].M + GEMM_BM - 1) / GEMM_BM, num_nb = (G.g[a.p].N + BN - 1) / BN;
    if (a.mb < 0 || a.mb >= num_mb || a.nb < 0 || a.nb >= num_nb) { printf("FAIL unit %d: tile (%d,%d) outside %dx%d\n", u, a.mb, a.nb, num_mb, num_nb); return 1; }
    if (a.tile < 0 || a.tile >= num_mb * num_nb) { printf("FAIL unit %d: tile index %d\n", u, a.tile); return 1; }
    seen[{a.p, a.mb, a.nb}].push_back({a.kb_begin, a.kb_end});
  }
  size_t want = 0;
  for (int i = 0; i < G.nprob; ++i) {
    const int num_mb = (G.g[i].M + GEMM_BM - 1) / GEMM_BM, num_nb = (G.g[i].N + BN - 1) / BN;
    const int num_kb = (G.g[i].K + GEMM_BK - 1) / GEMM_BK;
    const int sp = G.g[i].splits > 1 ? G.g[i].splits : 1;
    want += (size_t)num_mb * num_nb;
    for (int mb = 0; mb < num_mb; ++mb)
      for (int nb = 0; nb < num_nb; ++nb) {
        auto it = seen.find({i, mb, nb});
        if (it == seen.end() || (int)it->second.size() != sp) { printf("FAIL problem %d tile (%d,%d): visited %d times, want %d\n", i, mb, nb, it == seen.end() ? 0 : (int)it->second.size(), sp); return 1; }
        std::vector<char> cover(num_kb, 0);
        for (auto& r : it->second)
          for (int kb = r.first; kb < r.second; ++kb) {
            if (kb < 0 || kb >= num_kb || cover[kb]) { printf("FAIL problem %d tile (%d,%d): k-block %d out of range or covered twice\n", i, mb, nb, kb); return 1; }
            cover[kb] = 1;
          }
        for (int kb = 0; kb < num_kb; ++kb)
          if (!cover[kb]) { printf("FAIL problem %d tile (%d,%d): k-block %d never covered\n", i, mb, nb, kb); return 1; }
      }
  }
  if (seen.size() != want) { printf("FAIL: %zu distinct tiles visited, want %zu\n", seen.size(), want); return 1; }
  return 0;
}

template <int BN>
static int run_all() {
  int bad = 0;
  // the groups the modules launch at C2 / C4 / C5 (dW trio, q+kv, dyn+dvis) and ragged / degenerate cases
  bad += check_group<BN>({{768, 512, 4096, 1}, {512, 768, 4096, 1}, {1024, 768, 2048, 1}});
  bad += check_group<BN>({{4096, 512, 768, 1}, {2048, 1024, 768, 1}});
  bad += check_group<BN>({{4096, 768, 512, 1}, {2048, 768, 1024, 1}});
  bad += check_group<BN>({{2048, 512, 4096, 1}, {512, 2048, 4096, 1}, {1024, 1024, 2048, 1}});
  bad += check_group<BN>({{4096, 512, 4096, 1}, {256, 1024, 1024, 1}});
  bad += check_group<BN>({{136, 8, 72, 1}, {1000, 520, 200, 1}, {128, 64, 64, 1}, {300, 200, 136, 1}});
  bad += check_group<BN>({{1, 8, 1, 1}});
  bad += check_group<BN>({{130 * 8, 200, 1000, 1}});
  // single problems with serial split-K (uneven K ranges included)
  bad += check_group<BN>({{512, 768, 4096, 4}});
  bad += check_group<BN>({{1024, 768, 3648, 3}});
  bad += check_group<BN>({{130 * 8, 200, 1000, 2}});
  return bad;
}

int main() {
  const int bad = run_all<64>() + run_all<128>() + run_all<192>() + run_all<256>();
  if (bad) { printf("GROUP SCHEDULE FAIL (%d)\n", bad); return 1; }
  printf("GROUP SCHEDULE OK\n");
  return 0;
}
