// Probe 3 (round 2): operand forms the restructured policy-MLP backward (csrc/mlp_tc.cu v2) wants to rely on.
//   T1  MN-major tf32 operands with the SWIZZLE_128B_BASE32B layout (descriptor layout type 1, 32-byte swizzle atoms:
//       byte bits [5,7) ^= bits [7,9)), A (M = 64) and B (N = 64), reduction over the 128 rows (samples):
//       D2[i][j] = sum_s P[s][i] Q[s][j].  Image = [32-column block][row s][128 B]; two (LBO, SBO) conventions tried.
//   T2  A operand read from TENSOR MEMORY (written by the owning threads with tcgen05.st), B K-major SWIZZLE_128B:
//       D1[s][n] = sum_i P[s][i] Q[n][i], 3xTF32 with the lo part in TMEM as well.
//   T3  thin product: A = MN-major image of P (M = 64), B = small K-major image of X^T [16][128]:
//       D3[i][n] = sum_s P[s][i] X[s][n]  (N = 16).
//   T4  cycles per MMA for the shapes above (256 dependent MMAs each).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ppo-exploration_b200/csrc -o tools/umma_probe3 tools/umma_probe3.cu
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

__host__ __device__ constexpr uint32_t idesc_full(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
// MN-major descriptor, layout type selectable (1 = SWIZZLE_128B_BASE32B)
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)type << 61;
  return d;
}
__device__ __forceinline__ float lo_of(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// byte offset of element (row s, column c < 64) in the MN-major BASE32B image [2 blocks][128 rows][128 B]
__device__ __forceinline__ uint32_t mn_off(int s, int c) {
  const int cc = c & 31;
  return (uint32_t)((c >> 5) * 16384 + s * 128 + ((((cc >> 3) ^ (s & 3)) << 5) | ((cc & 7) << 2)));
}

// out: D1 [128][64], D2a/D2b [64][64] (two LBO/SBO conventions), D3 [64][16]; cyc[2*k], cyc[2*k+1]
__global__ void __launch_bounds__(128, 1) probe(const float* P, const float* Q, const float* X, float* D1, float* D2a, float* D2b,
                                                 float* D3, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t Pmn = smem, Pmn_lo = smem + 32768, Qmn = smem + 65536, Qmn_lo = smem + 98304;   // MN-major images
  const uint32_t Qk = smem + 131072, Qk_lo = smem + 163840;                                       // K-major SW128 [2 kb][128][128B]
  const uint32_t Xk = smem + 196608, Xk_lo = smem + 196608 + 8192;                                // K-major SW128 [4 kb][16][128B]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  for (uint32_t a = tid * 4; a < 196608 + 16384; a += 128 * 4) asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem + a), "f"(0.f));
  __syncthreads();
  const uint32_t tmem_pre = tmem_slot;
  // thread = row s
  uint32_t praw[64], plo[64];
  for (int c = 0; c < 64; ++c) {
    const float p = P[tid * 64 + c], q = Q[tid * 64 + c];
    praw[c] = __float_as_uint(p); plo[c] = __float_as_uint(lo_of(p));
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(Pmn + mn_off(tid, c)), "f"(p));
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(Pmn_lo + mn_off(tid, c)), "f"(lo_of(p)));
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(Qmn + mn_off(tid, c)), "f"(q));
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(Qmn_lo + mn_off(tid, c)), "f"(lo_of(q)));
    const uint32_t ko = (uint32_t)((c >> 5) * 16384) + sw128_off(tid, c & 31);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(Qk + ko), "f"(q));
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(Qk_lo + ko), "f"(lo_of(q)));
  }
  for (int n = 0; n < 16; ++n) {       // X^T small image: B[n][s]
    const float x = X[tid * 16 + n];
    const uint32_t xo = (uint32_t)((tid >> 5) * 2048) + sw128_off(n, tid & 31);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(Xk + xo), "f"(x));
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(Xk_lo + xo), "f"(lo_of(x)));
  }
  // A operand of T2 in tensor memory: lane = row s, columns 256.. (raw) and 320.. (lo)
  {
    const uint32_t ta = tmem_pre + ((uint32_t)(warp * 32) << 16);
    uint32_t v[32];
    for (int h = 0; h < 2; ++h) {
      for (int j = 0; j < 32; ++j) v[j] = praw[h * 32 + j];
      tmem_st32(ta + 256 + h * 32, v);
      for (int j = 0; j < 32; ++j) v[j] = plo[h * 32 + j];
      tmem_st32(ta + 320 + h * 32, v);
    }
    tmem_st_wait();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && elect_one()) {
    // ---- T2: A from TMEM (M = 128), B K-major; D1 -> columns 0..63
    {
      constexpr uint32_t id = idesc_full(128, 64, 0, 0);
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t ob = (ks >> 2) * 16384 + (ks & 3) * 32;
        umma_tf32_ts(tmem, tmem + 320 + ks * 8, make_desc(Qk + ob), id, ks ? 1u : 0u);
        umma_tf32_ts(tmem, tmem + 256 + ks * 8, make_desc(Qk_lo + ob), id, 1u);
      }
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t ob = (ks >> 2) * 16384 + (ks & 3) * 32;
        umma_tf32_ts(tmem, tmem + 256 + ks * 8, make_desc(Qk + ob), id, 1u);
      }
    }
    // ---- T1: MN-major BASE32B both operands, M = 64, N = 64, K = 128; two conventions -> columns 64.. and 128..
    for (int var = 0; var < 2; ++var) {
      constexpr uint32_t id = idesc_full(64, 64, 1, 1);
      const uint32_t lbo = var == 0 ? 16384u : 512u, sbo = var == 0 ? 512u : 16384u;
      const uint32_t d = tmem + 64 + var * 64;
      for (int ks = 0; ks < 16; ++ks) {
        const uint32_t o = ks * 1024;
        umma_tf32(d, desc_mn(Pmn_lo + o, lbo, sbo, 1), desc_mn(Qmn + o, lbo, sbo, 1), id, ks ? 1u : 0u);
        umma_tf32(d, desc_mn(Pmn + o, lbo, sbo, 1), desc_mn(Qmn_lo + o, lbo, sbo, 1), id, 1u);
      }
      for (int ks = 0; ks < 16; ++ks) {
        const uint32_t o = ks * 1024;
        umma_tf32(d, desc_mn(Pmn + o, lbo, sbo, 1), desc_mn(Qmn + o, lbo, sbo, 1), id, 1u);
      }
    }
    // ---- T3: A MN-major (M = 64), B small K-major (N = 16), K = 128 -> columns 192..207
    {
      constexpr uint32_t id = idesc_full(64, 16, 1, 0);
      const uint32_t d = tmem + 192;
      for (int ks = 0; ks < 16; ++ks) {
        const uint32_t oa = ks * 1024, ob = (ks >> 2) * 2048 + (ks & 3) * 32;
        umma_tf32(d, desc_mn(Pmn_lo + oa, 16384, 512, 1), make_desc(Xk + ob), id, ks ? 1u : 0u);
        umma_tf32(d, desc_mn(Pmn + oa, 16384, 512, 1), make_desc(Xk_lo + ob), id, 1u);
      }
      for (int ks = 0; ks < 16; ++ks) {
        const uint32_t oa = ks * 1024, ob = (ks >> 2) * 2048 + (ks & 3) * 32;
        umma_tf32(d, desc_mn(Pmn + oa, 16384, 512, 1), make_desc(Xk + ob), id, 1u);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  {
    uint32_t v[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      for (int j = 0; j < 32; ++j) D1[tid * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    for (int var = 0; var < 2; ++var)
      for (int c0 = 0; c0 < 64; c0 += 32) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(64 + var * 64 + c0), v);
        if (lane < 16)
          for (int j = 0; j < 32; ++j) (var ? D2b : D2a)[(warp * 16 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
      }
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 192u, v);
    if (lane < 16)
      for (int j = 0; j < 16; ++j) D3[(warp * 16 + lane) * 16 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  // ---- T4: timings (results are garbage accumulations into columns 384..)
  uint32_t phase = 1;
  for (int cfg = 0; cfg < 8; ++cfg) {
    long long t0 = 0;
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem + 384;
        t0 = clock64();
        for (int rep = 0; rep < 16; ++rep)
          for (int ks = 0; ks < 16; ++ks) {
            const uint32_t acc = (rep | ks) ? 1u : 0u;
            const uint32_t oa = ks * 1024, okb = ((ks & 7) >> 2) * 16384 + (ks & 3) * 32, oxb = (ks >> 2) * 2048 + (ks & 3) * 32;
            switch (cfg) {
              case 0: umma_tf32(d, desc_mn(Pmn + oa, 16384, 512, 1), desc_mn(Qmn + oa, 16384, 512, 1), idesc_full(64, 64, 1, 1), acc); break;
              case 1: umma_tf32(d, desc_mn(Pmn + oa, 16384, 512, 1), make_desc(Xk + oxb), idesc_full(64, 16, 1, 0), acc); break;
              case 2: umma_tf32(d, desc_mn(Pmn + oa, 16384, 512, 1), make_desc(Xk + oxb), idesc_full(128, 16, 1, 0), acc); break;
              case 3: umma_tf32(d, desc_mn(Pmn + oa, 16384, 512, 1), make_desc(Xk + oxb), idesc_full(64, 8, 1, 0), acc); break;
              case 4: umma_tf32_ts(d, tmem + 256 + (ks & 7) * 8, make_desc(Qk + okb), idesc_full(128, 64, 0, 0), acc); break;
              case 5: umma_tf32(d, make_desc(Qk + okb), make_desc(Qk + okb), idesc_full(128, 64, 0, 0), acc); break;
              case 6: umma_tf32(d, desc_mn(Pmn + oa, 16384, 512, 1), desc_mn(Qmn + oa, 16384, 512, 1), idesc_full(128, 64, 1, 1), acc); break;
              default: umma_tf32(d, desc_mn(Pmn + oa, 16384, 512, 1), desc_mn(Qmn + oa, 16384, 512, 1), idesc_full(128, 96, 1, 1), acc); break;
            }
          }
        const long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        cyc[cfg * 2] = t1 - t0;
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    tc_fence_after();
    if (tid == 0) cyc[cfg * 2 + 1] = clock64() - t0;
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
  std::vector<float> P(128 * 64), Q(128 * 64), X(128 * 16);
  srand(1);
  for (auto& x : P) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  for (auto& x : Q) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  for (auto& x : X) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  float *dP, *dQ, *dX, *d1, *d2a, *d2b, *d3; long long* dC;
  cudaMalloc(&dP, P.size() * 4); cudaMalloc(&dQ, Q.size() * 4); cudaMalloc(&dX, X.size() * 4);
  cudaMalloc(&d1, 128 * 64 * 4); cudaMalloc(&d2a, 64 * 64 * 4); cudaMalloc(&d2b, 64 * 64 * 4); cudaMalloc(&d3, 64 * 16 * 4);
  cudaMalloc(&dC, 16 * 8);
  cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(d1, 0, 128 * 64 * 4); cudaMemset(d2a, 0, 64 * 64 * 4); cudaMemset(d2b, 0, 64 * 64 * 4); cudaMemset(d3, 0, 64 * 16 * 4);
  const int smem = 196608 + 16384 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, smem>>>(dP, dQ, dX, d1, d2a, d2b, d3, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> D1(128 * 64), D2a(64 * 64), D2b(64 * 64), D3(64 * 16);
  long long C[16];
  cudaMemcpy(D1.data(), d1, D1.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D2a.data(), d2a, D2a.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D2b.data(), d2b, D2b.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D3.data(), d3, D3.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(C, dC, sizeof(C), cudaMemcpyDeviceToHost);
  double e1 = 0, s1 = 0, e2a = 0, e2b = 0, s2 = 0, e3 = 0, s3 = 0;
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
      e2a = fmax(e2a, fabs(r - D2a[i * 64 + j])); e2b = fmax(e2b, fabs(r - D2b[i * 64 + j])); s2 = fmax(s2, fabs(r));
    }
  for (int i = 0; i < 64; ++i)
    for (int n = 0; n < 16; ++n) {
      double r = 0;
      for (int s = 0; s < 128; ++s) r += (double)P[s * 64 + i] * X[s * 16 + n];
      e3 = fmax(e3, fabs(r - D3[i * 16 + n])); s3 = fmax(s3, fabs(r));
    }
  printf("T2 A-from-TMEM x B K-major (3xTF32):      max err %.3e (scale %.3f, rel %.2e)\n", e1, s1, e1 / s1);
  printf("T1 MN-major BASE32B (LBO 16384, SBO 512):  max err %.3e (scale %.3f, rel %.2e)\n", e2a, s2, e2a / s2);
  printf("T1 MN-major BASE32B (LBO 512, SBO 16384):  max err %.3e (scale %.3f, rel %.2e)\n", e2b, s2, e2b / s2);
  printf("T3 A MN-major x small K-major B (N = 16):  max err %.3e (scale %.3f, rel %.2e)\n", e3, s3, e3 / s3);
  printf("D2a[0][0..3] = %g %g %g %g   D3[0][0..3] = %g %g %g %g\n", D2a[0], D2a[1], D2a[2], D2a[3], D3[0], D3[1], D3[2], D3[3]);
  const char* names[8] = {"M64 N64 MNxMN", "M64 N16 MNxK", "M128 N16 MNxK", "M64 N8 MNxK", "M128 N64 TMEMxK", "M128 N64 KxK",
                          "M128 N64 MNxMN", "M128 N96 MNxMN"};
  for (int c = 0; c < 8; ++c)
    printf("%-16s: issue %.1f cycles/MMA, issue+complete %.1f cycles/MMA\n", names[c], C[2 * c] / 256.0, C[2 * c + 1] / 256.0);
  return 0;
}
