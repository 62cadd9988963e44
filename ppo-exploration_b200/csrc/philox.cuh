// Philox4x32-10 counter-based generator (shared by es.cu: noise table / population offsets, sample.cu: policy sampling).
#pragma once
#include <stdint.h>

namespace ppx {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
}
// 4 x 32 random bits for counter (c0, c1, c2, c3) under the 64-bit key
__device__ __forceinline__ void philox4x32(uint32_t (&c)[4], uint64_t key) {
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
// two independent N(0,1) from two 32-bit words (Box-Muller)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;       // (0,1]
  const float u2 = (float)b * 2.3283064365386963e-10f;
  const float rad = sqrtf(-2.0f * logf(fminf(u1, 1.0f)));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  z0 = rad * cs; z1 = rad * sn;
}

}  // namespace ppx
