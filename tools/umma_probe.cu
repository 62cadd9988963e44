// Probe for the operand conventions the fused tensor-core MLP kernels (csrc/mlp_tc.cu) rely on:
//   (1) kind::tf32 TRUNCATES its 32-bit operands (so the raw fp32 tile can serve as the "hi" operand of 3xTF32);
//   (2) MN-major SWIZZLE_128B descriptors over a [K rows][32 MN elements] block layout (LBO = block pitch);
//   (3) M = 128 instructions whose upper A rows are unrelated shared memory leave the lower D rows intact.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ppo-exploration_b200/csrc -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tc_common.cuh"

namespace ppx {
int fail(int code, const char*, ...) { return code; }
void count_launch(int) {}
int sm_count() { return 148; }
}  // namespace ppx
using namespace ppx::tc;

__global__ void __launch_bounds__(128, 1) probe(const float* P, const float* Q, float* D1, float* D2, float* D3, int mask_hi, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* Praw = smem; uint8_t* Plo = smem + 32768; uint8_t* Qraw = smem + 65536; uint8_t* Qlo = smem + 98304;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  // thread = row s: write its 64 P and 64 Q values (raw | lo) into the swizzled [block][s][32] layout
  for (int c = 0; c < 64; ++c) {
    const float p = P[tid * 64 + c], q = Q[tid * 64 + c];
    const float ph = __uint_as_float(__float_as_uint(p) & 0xFFFFE000u), qh = __uint_as_float(__float_as_uint(q) & 0xFFFFE000u);
    const uint32_t off = (c >> 5) * 16384 + sw128_off(tid, c & 31);
    *reinterpret_cast<float*>(Praw + off) = mask_hi ? ph : p;
    *reinterpret_cast<float*>(Plo + off) = p - ph;
    *reinterpret_cast<float*>(Qraw + off) = mask_hi ? qh : q;
    *reinterpret_cast<float*>(Qlo + off) = q - qh;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && elect_one()) {
    const uint32_t pr = smem_u32(Praw), pl = smem_u32(Plo), qr = smem_u32(Qraw), ql = smem_u32(Qlo);
    // test 1: K-major.  D1[s][n] = sum_i P[s][i] Q[n][i]
    constexpr uint32_t id1 = make_idesc_major(64, 0, 0);
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t o = (ks >> 2) * 16384 + (ks & 3) * 32;
      umma_tf32(tmem, make_desc(pl + o), make_desc(qr + o), id1, ks ? 1u : 0u);
      umma_tf32(tmem, make_desc(pr + o), make_desc(ql + o), id1, 1u);
    }
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t o = (ks >> 2) * 16384 + (ks & 3) * 32;
      umma_tf32(tmem, make_desc(pr + o), make_desc(qr + o), id1, 1u);
    }
    // test 2: MN-major.  D2[i][j] = sum_s P[s][i] Q[s][j]
    constexpr uint32_t id2 = make_idesc_major(64, 1, 1);
    for (int ks = 0; ks < 16; ++ks) {
      const uint32_t o = ks * 1024;
      umma_tf32(tmem + 64, make_desc_mn(pl + o, 16384, 1024), make_desc_mn(qr + o, 16384, 1024), id2, ks ? 1u : 0u);
      umma_tf32(tmem + 64, make_desc_mn(pr + o, 16384, 1024), make_desc_mn(ql + o, 16384, 1024), id2, 1u);
    }
    for (int ks = 0; ks < 16; ++ks) {
      const uint32_t o = ks * 1024;
      umma_tf32(tmem + 64, make_desc_mn(pr + o, 16384, 1024), make_desc_mn(qr + o, 16384, 1024), id2, 1u);
    }
    // test 3: A K-major, B MN-major.  D3[s][j] = sum_{i<64} P[s][i] Q[i][j]
    constexpr uint32_t id3 = make_idesc_major(64, 0, 1);
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t oa = (ks >> 2) * 16384 + (ks & 3) * 32, ob = ks * 1024;
      uint32_t lbo = 16384, sbo = 1024;
      if (variant == 1) { lbo = 1024; sbo = 16384; }
      umma_tf32(tmem + 128, make_desc(pl + oa), make_desc_mn(qr + ob, lbo, sbo), id3, ks ? 1u : 0u);
      umma_tf32(tmem + 128, make_desc(pr + oa), make_desc_mn(ql + ob, lbo, sbo), id3, 1u);
      umma_tf32(tmem + 128, make_desc(pr + oa), make_desc_mn(qr + ob, lbo, sbo), id3, 1u);
    }
    umma_commit(smem_u32(&bar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int t = 0; t < 3; ++t)
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * 64 + c0), v);
      float* D = t == 2 ? D3 : (t ? D2 : D1);
      for (int j = 0; j < 32; ++j) D[tid * 64 + c0 + j] = __uint_as_float(v[j]);
    }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

int main() {
  std::vector<float> P(128 * 64), Q(128 * 64);
  srand(1);
  for (auto& x : P) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  for (auto& x : Q) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  float *dP, *dQ, *d1, *d2, *d3;
  cudaMalloc(&dP, P.size() * 4); cudaMalloc(&dQ, Q.size() * 4); cudaMalloc(&d1, 128 * 64 * 4); cudaMalloc(&d2, 128 * 64 * 4); cudaMalloc(&d3, 128 * 64 * 4);
  cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 131072 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int variant = 0; variant < 2; ++variant) { int mask = 0;
    probe<<<1, 128, smem>>>(dP, dQ, d1, d2, d3, mask, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> D1(128 * 64), D2(128 * 64);
    cudaMemcpy(D1.data(), d1, D1.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D2.data(), d2, D2.size() * 4, cudaMemcpyDeviceToHost);
    std::vector<float> D3(128 * 64);
    cudaMemcpy(D3.data(), d3, D3.size() * 4, cudaMemcpyDeviceToHost);
    double e3 = 0, s3 = 0;
    for (int s = 0; s < 128; ++s)
      for (int j = 0; j < 64; ++j) {
        double r = 0;
        for (int i = 0; i < 64; ++i) r += (double)P[s * 64 + i] * Q[i * 64 + j];
        e3 = fmax(e3, fabs(r - D3[s * 64 + j])); s3 = fmax(s3, fabs(r));
      }
    printf("variant %d: A K-major x B MN-major: max err %.3e (scale %.3f)\n", variant, e3, s3);
    printf("D2[0][0..3] = %g %g %g %g   D3[0][0..3] = %g %g %g %g\n", D2[0], D2[1], D2[2], D2[3], D3[0], D3[1], D3[2], D3[3]);
    double e1 = 0, e2 = 0, s1 = 0, s2 = 0;
    for (int s = 0; s < 128; ++s)
      for (int n = 0; n < 64; ++n) {
        double r = 0;
        for (int i = 0; i < 64; ++i) r += (double)P[s * 64 + i] * Q[n * 64 + i];
        e1 = fmax(e1, fabs(r - D1[s * 64 + n])); s1 = fmax(s1, fabs(r));
      }
    for (int i = 0; i < 64; ++i)
      for (int j = 0; j < 64; ++j) {
        double r = 0;
        for (int s = 0; s < 128; ++s) r += (double)P[s * 64 + i] * Q[s * 64 + j];
        e2 = fmax(e2, fabs(r - D2[i * 64 + j])); s2 = fmax(s2, fabs(r));
      }
    printf("mask_hi=%d  K-major: max err %.3e (scale %.3f, rel %.2e)   MN-major: max err %.3e (scale %.3f, rel %.2e)\n", mask, e1, s1,
           e1 / s1, e2, s2, e2 / s2);
  }
  return 0;
}
