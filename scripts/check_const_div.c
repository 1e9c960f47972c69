// check_const_div.c -- exhaustive check of the constant-divisor division used by the kernels
// (cuda_flow3d_b200/csrc/common.cuh: div_const): for NC divisors c (level spacings 2h / 4h with
// h = W/(float)w_level, plus random mantissas) and EVERY float mantissa of x, the five-FMA sequence
// must equal the IEEE quotient x / c.
//   gcc -O2 -ffp-contract=off -o /tmp/check_const_div scripts/check_const_div.c -lm && /tmp/check_const_div 200 1
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float div5(float x, float c, float r) {
  float q = x * r;
  float e = fmaf(-c, q, x);
  q = fmaf(r, e, q);
  e = fmaf(-c, q, x);
  q = fmaf(r, e, q);
  return q;
}
static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(int argc, char** argv) {
  const int nc = argc > 1 ? atoi(argv[1]) : 100;
  s += argc > 2 ? atoi(argv[2]) : 0;
  long bad = 0, n = 0;
  for (int ic = 0; ic < nc; ic++) {
    float c;
    if (ic % 3 == 0) {
      uint32_t b = (uint32_t)rnd();
      b = (b & 0x007fffffu) | ((uint32_t)(127 + (rnd() % 4)) << 23);
      memcpy(&c, &b, 4);
    } else {
      const int W = 4 + (int)(rnd() % 3000), cw = 4 + (int)(rnd() % W);
      const float h = W / (float)cw;
      c = (ic % 3 == 2) ? 4.f * h : h + h;
    }
    const float r = 1.0f / c;
    for (uint32_t m = 0; m < (1u << 23); m++) {
      const uint32_t b = m | (130u << 23);
      float x;
      memcpy(&x, &b, 4);
      n++;
      if (div5(x, c, r) != x / c) bad++;
    }
  }
  printf("divisors=%d quotients=%ld mismatches=%ld\n", nc, n, bad);
  return bad != 0;
}
