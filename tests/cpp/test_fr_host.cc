// Host field helpers (csrc/fr_host.h): the CIOS Montgomery product against the double-and-add product, on edge values
// and pseudo-random ones.  Built and run by tests/test_plan.py (CPU only).
#include <stdio.h>

#include <initializer_list>

#include "../../halo2-dynamic-sha256_b200/csrc/fr_host.h"

using namespace h2sha;

static uint64_t rng_state = 0x9E3779B97F4A7C15ULL;
static uint64_t next64() {
  uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static U256 random_fr() {
  U256 a = {{next64(), next64(), next64(), next64() >> 3}};
  while (fr::geq_p(a.l)) fr::sub_p(a.l);
  return a;
}

int main() {
  const U256 r = fr::mont_r();
  const U256 r_inv = fr::inv(r);
  U256 pm1;
  for (int i = 0; i < 4; i++) pm1.l[i] = fr::P[i];
  pm1.l[0] -= 1;
  U256 edge[] = {fr::from_u64(0), fr::from_u64(1), fr::from_u64(2), r, pm1, fr::from_u64(0xffffffffffffffffULL)};
  int n_checked = 0;
  for (int k = 0; k < 200; k++) {
    const U256 a = k < 6 ? edge[k] : random_fr();
    const U256 b = (k % 7 == 0) ? edge[k % 6] : random_fr();
    const U256 want = fr::mul(fr::mul(a, b), r_inv);   // a * b * R^-1 mod p
    const U256 got = fr::mont_mul(a, b);
    if (!(want == got)) { printf("mont_mul mismatch at case %d\n", k); return 1; }
    n_checked++;
  }
  // to_mont through R^2 and back
  const U256 r2 = fr::to_mont(r);
  for (uint64_t v : {0ULL, 1ULL, 255ULL, 65535ULL, 0x5555ULL}) {
    const U256 m = fr::mont_mul(fr::from_u64(v), r2);
    if (!(m == fr::to_mont(fr::from_u64(v)))) { printf("to_mont mismatch for %llu\n", (unsigned long long)v); return 1; }
    if (!(fr::mont_mul(m, fr::from_u64(1)) == fr::from_u64(v))) { printf("from_mont mismatch for %llu\n", (unsigned long long)v); return 1; }
  }
  printf("ok %d\n", n_checked);
  return 0;
}
