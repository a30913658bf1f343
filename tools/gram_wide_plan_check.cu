// Host-side check of gram_wide_plan: every packed tile (i <= j) is covered exactly once, staged offsets are
// consistent, shared memory fits.   nvcc -o tools/build/gwplan tools/gram_wide_plan_check.cu && tools/build/gwplan
#include <cstdio>
#include <map>
#include "../linearresponsevariationalbayes.py_b200/csrc/gram_wide.cuh"
namespace lrvb { void set_error(const char*, ...) {} long long g_launches = 0; }
using namespace lrvb;
int main() {
  int bad = 0;
  for (int K = 96; K <= 256; K += 8) {
    GwPlan pl = gram_wide_plan(K);
    const int T0 = K / 8, T2 = 2 * T0;
    std::map<std::pair<int,int>, int> cover;
    long sumload = 0;
    for (size_t g = 0; g < pl.groups.size(); ++g) {
      const GwGroup& G = pl.groups[g];
      if (G.nblk > kGwMaxBlk || gram_wide_smem(G.ZS, G.TN) > kGwSmemCap || G.TN % 4 || G.ZS % 8 != 4) { printf("K=%d group %zu bad geometry\n", K, g); ++bad; }
      for (int w = 0; w < kGwWarps; ++w) {
        const GwJob& j = G.job[w];
        if (!j.ni) continue;
        // staged offsets must point at blocks whose X column / class match the packed tile coordinates
        auto chk = [&](int off, int t0, int nt) {
          for (int b = 0; b < G.nblk; ++b) if (G.blk_off[b] == off) {
            const int sq = t0 >= T0, xcol = 8 * (t0 - sq * T0);
            if (G.blk_sq[b] != sq || G.blk_x[b] != xcol || G.blk_cols[b] != 8 * nt) { printf("K=%d block mismatch\n", K); ++bad; }
            return;
          }
          printf("K=%d offset not found\n", K); ++bad;
        };
        chk(j.offA, j.ti0, j.ni); chk(j.offB, j.tj0, j.nj);
        const int wexp = j.tj0 < T0 ? 0 : (j.ti0 < T0 ? 1 : 2);
        if (wexp != j.wrow) { printf("K=%d weight row\n", K); ++bad; }
        for (int a = 0; a < j.ni; ++a) for (int b = 0; b < j.nj; ++b) if (!j.stair || a <= b) cover[{j.ti0 + a, j.tj0 + b}]++;
      }
      sumload += pl.load[g];
    }
    for (int i = 0; i < T2; ++i) for (int j = i; j < T2; ++j) if (cover[{i, j}] != 1) { printf("K=%d tile (%d,%d) covered %d times\n", K, i, j, cover[{i,j}]); ++bad; }
    if ((int)cover.size() != T2 * (T2 + 1) / 2) { printf("K=%d extra tiles %zu\n", K, cover.size()); ++bad; }
    printf("K=%3d T2=%2d groups=%2zu ctas=%3zu live=%4d  balance=%.3f  smem=%zu  TN:", K, T2, pl.groups.size(), pl.ctas.size(), pl.tiles_live, pl.tiles_live / (4.0 * sumload), pl.smem);
    std::vector<int> nc(pl.groups.size(), 0);
    for (auto& c : pl.ctas) nc[c.group]++;
    for (size_t g = 0; g < pl.groups.size(); ++g) printf(" %d/%d/%d/%d", pl.groups[g].TN, pl.groups[g].nblk, pl.load[g], nc[g]);
    printf("\n");
  }
  printf(bad ? "FAILED %d\n" : "plan ok\n", bad);
  return bad != 0;
}
