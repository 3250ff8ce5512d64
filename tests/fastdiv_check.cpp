// Host-side check of the multiply-high divider used by the conv kernels' tile decode (csrc/igemm.cuh FastDiv): every
// divisor 1..4096 plus a spread of larger ones, against dividends covering [0, 2^31).  Built and run by tests/test_cpu.py.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../faceoff_b200/csrc/igemm.cuh"

int main() {
  long bad = 0, checked = 0;
  for (int d = 1; d <= 2000000; d += (d <= 4096 ? 1 : 997)) {
    const fo::FastDiv f = fo::make_fastdiv(d);
    for (long n = 0; n < (1L << 31); n += (n < 70000 ? 1 : 1000003)) {
      const uint32_t q = f.div == 1 ? (uint32_t)n : (uint32_t)((((uint64_t)(uint32_t)n * f.mul) >> 32) >> f.shr);
      ++checked;
      if (q != (uint32_t)(n / d) && bad++ < 5) printf("bad d=%d n=%ld q=%u\n", d, n, q);
    }
    // the largest dividends
    for (long n = (1L << 31) - 70000; n < (1L << 31); ++n) {
      const uint32_t q = f.div == 1 ? (uint32_t)n : (uint32_t)((((uint64_t)(uint32_t)n * f.mul) >> 32) >> f.shr);
      if (q != (uint32_t)(n / d) && bad++ < 5) printf("bad d=%d n=%ld q=%u\n", d, n, q);
    }
  }
  printf("checked=%ld bad=%ld\n", checked, bad);
  return bad != 0;
}
