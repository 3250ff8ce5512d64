// Host-side model of the weight-gradient kernel's accumulation-chain protocol (csrc/igemm.cuh wgrad_plan_chains +
// the per-CTA chunk arithmetic of csrc/wgrad_igemm.cu): for a sweep of (pixel tiles, splits, passes, kpix, chain bound)
//   * the chunks of every split cover its tile range exactly once, in order, within n_flush partial slots,
//   * no chunk accumulates more than the bound (in MMAs), and chunks are balanced (sizes differ by < chain_tiles),
//   * the bias-column predicate of the epilogue ("this chunk issued a column-sum MMA") equals a direct search,
//   * the partial-slot index (split * n_flush + chunk) enumerates [0, splits * n_flush) once.
// Built and run by tests/test_cpu.py.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <vector>

#include "../faceoff_b200/csrc/igemm.cuh"

int main() {
  long bad = 0, checked = 0;
  const int tiles_list[] = {1, 2, 7, 148, 149, 960, 1920, 3840, 15360, 30720, 30721, 122880, 245760, 1000003};
  const int passes_list[] = {1, 3, 4, 9};
  const int chain_list[] = {0, 1, 64, 2048, 4096, 1 << 20};
  for (int total : tiles_list)
    for (int passes : passes_list)
      for (int kpix : {64, 128})
        for (int chain : chain_list) {
          int splits = 148 / passes;
          if (splits > total) splits = total;
          const int per = (total + splits - 1) / splits;
          int n_flush, chain_tiles;
          fo::wgrad_plan_chains(per, kpix, chain, &n_flush, &chain_tiles);
          if (n_flush < 1 || chain_tiles < 1) { if (bad++ < 5) printf("degenerate plan\n"); continue; }
          if (chain > 0 && (long long)chain_tiles * (kpix / 16) > (long long)chain + (kpix / 16) - 1 && n_flush > 1 && bad++ < 5)
            printf("chain too long: total %d passes %d kpix %d chain %d -> tiles %d\n", total, passes, kpix, chain, chain_tiles);
          if ((long long)n_flush * chain_tiles < per && bad++ < 5) printf("chunks do not cover a split: per %d nf %d ct %d\n", per, n_flush, chain_tiles);
          std::vector<char> slot_seen((size_t)splits * n_flush, 0);
          for (int split = 0; split < splits; ++split) {
            const int t_begin = split * per, t_end = std::min(total, t_begin + per);
            const int n_my = std::max(0, t_end - t_begin);
            int covered = 0;
            for (int chunk = 0; chunk < n_flush; ++chunk) {
              const int c0 = std::min(n_my, chunk * chain_tiles), c1 = std::min(n_my, c0 + chain_tiles);
              if (c0 != covered && bad++ < 5) printf("gap: split %d chunk %d c0 %d covered %d\n", split, chunk, c0, covered);
              covered = c1;
              const size_t slot = (size_t)split * n_flush + chunk;
              if (slot_seen[slot]++ && bad++ < 5) printf("slot reused\n");
              for (int pass = 0; pass < passes; ++pass) {
                const int first = c0 + ((pass - c0 % passes) + passes) % passes;   // the epilogue's predicate
                bool any = false;
                for (int i = c0; i < c1 && !any; ++i) any = (i % passes) == pass;
                if (((c1 > c0) && first < c1) != any && bad++ < 5) printf("bias predicate: c0 %d c1 %d pass %d\n", c0, c1, pass);
                ++checked;
              }
            }
            if (covered != n_my && bad++ < 5) printf("split %d: covered %d of %d\n", split, covered, n_my);
          }
        }
  printf("checked=%ld bad=%ld\n", checked, bad);
  return bad != 0;
}
