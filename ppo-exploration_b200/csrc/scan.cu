// GAE / discounted-return reverse scans over [T,N] rollouts.
//
// Replaces RolloutStorage.compute_returns_and_advantages (buffer.py:203-230), the dual-head
// IntrinsicStorage version (buffer.py:321-362) and SilModule.discount_with_dones (sil_module.py:99-105).
//
// Design (B200): the recurrence X_t = delta_t + c_t * X_{t+1} is a composition of affine maps, so it
// is scanned in parallel along T.  A CTA owns COLS adjacent env columns; rows of the [T,N] arrays are
// streamed time-tile by time-tile (last tile first) into shared memory with fully coalesced row loads,
// then ONE WARP PER ENV COLUMN scans its column: lane l holds the affine map of row (32*g + l), a
// 5-step Kogge-Stone suffix scan over the warp composes them, and the warp-uniform carry links
// 32-row groups and tiles.  Terminal masking is just c_t = 0 ("segmented" scan).  Results go back
// through shared memory so stores are coalesced too.  Algorithmic traffic: 17 B/transition single
// head, 33 B dual; nothing is read twice from HBM except one boundary row per tile.
//
// Precision follows the reference's numpy promotion (see oracle/rollout.py:gae): gamma*V_{t+1} is an
// f32 product, delta and the carry are f64, advantages are rounded to f32 on store and
// returns = f32(adv) + V in f32.  The intrinsic head forms its delta in f32 like the reference; its
// carry is held in f64 here (the reference carries it in f32), which only removes rounding noise.
#include "common.cuh"

namespace ppx {
namespace {

template <int COLS>
struct ScanCfg {
  static constexpr int kRows = 2048 / COLS;       // time-tile height: 64 / 128 / 256 rows
  static constexpr int kLd = COLS + 1;            // padded leading dim -> conflict-free column reads
  static constexpr int kThreads = COLS * 32;
};

template <int COLS, bool DUAL>
__global__ void __launch_bounds__(COLS * 32)
gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
           const uint8_t* __restrict__ masks, const float* __restrict__ last_value,
           const uint8_t* __restrict__ last_done, const float* __restrict__ int_rewards,
           const float* __restrict__ int_values, const float* __restrict__ last_int_value,
           float g32, double gl, float gi32, double gil, int T, int N,
           float* __restrict__ adv, float* __restrict__ ret, float* __restrict__ iadv,
           float* __restrict__ iret) {
  using C = ScanCfg<COLS>;
  constexpr int LD = C::kLd, ROWS = C::kRows;
  __shared__ float s_r[ROWS * LD];
  __shared__ float s_v[(ROWS + 1) * LD];
  __shared__ uint8_t s_m[(ROWS + 1) * LD];
  __shared__ float s_ir[DUAL ? ROWS * LD : 1];
  __shared__ float s_iv[DUAL ? (ROWS + 1) * LD : 1];

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int c0 = blockIdx.x * COLS;
  double carry = 0.0, icarry = 0.0;               // warp-uniform: X_{t+1} of this warp's column

  const int n_tiles = (T + ROWS - 1) / ROWS;
  for (int tile = n_tiles - 1; tile >= 0; --tile) {
    const int t0 = tile * ROWS;
    const int rows = min(ROWS, T - t0);
    const int t1 = t0 + rows;
    // ---- coalesced tile load (rows t0..t1-1, plus the boundary row t1) ----
    for (int i = tid; i < (rows + 1) * COLS; i += C::kThreads) {
      const int row = i / COLS, c = i - row * COLS, col = c0 + c;
      if (col >= N) continue;
      const int s = row * LD + c;
      if (row < rows) {
        const size_t g = (size_t)(t0 + row) * N + col;
        s_r[s] = ld_stream(rewards + g);
        s_v[s] = ld_stream(values + g);
        s_m[s] = masks[g];
        if (DUAL) {
          s_ir[s] = ld_stream(int_rewards + g);
          s_iv[s] = ld_stream(int_values + g);
        }
      } else if (t1 < T) {                         // boundary row inside the rollout
        const size_t g = (size_t)t1 * N + col;
        s_v[s] = values[g];
        s_m[s] = masks[g];
        if (DUAL) s_iv[s] = int_values[g];
      } else {                                     // bootstrap row: last_value / dones (buffer.py:221-223)
        s_v[s] = last_value[col];
        s_m[s] = last_done[col];
        if (DUAL) s_iv[s] = last_int_value[col];
      }
    }
    __syncthreads();
    // ---- one warp per env column: suffix scan of affine maps, 32 rows at a time ----
    if (c0 + w < N) {
      for (int g = (rows - 1) >> 5; g >= 0; --g) {
        const int row = (g << 5) + lane;
        const bool valid = row < rows;
        double a = 0.0, b = 1.0, ia = 0.0, ib = 1.0;
        if (valid) {
          const int s = row * LD + w, sn = s + LD;
          const double nnt = 1.0 - (double)s_m[sn];
          const float gv = __fmul_rn(g32, s_v[sn]);
          a = __dsub_rn(__dadd_rn((double)s_r[s], __dmul_rn((double)gv, nnt)), (double)s_v[s]);
          b = gl * nnt;
          if (DUAL) {
            ia = (double)__fsub_rn(__fadd_rn(s_ir[s], __fmul_rn(gi32, s_iv[sn])), s_iv[s]);
            ib = gil;
          }
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const double ao = __shfl_down_sync(0xffffffffu, a, d);
          const double bo = __shfl_down_sync(0xffffffffu, b, d);
          double iao = 0.0, ibo = 1.0;
          if (DUAL) {
            iao = __shfl_down_sync(0xffffffffu, ia, d);
            ibo = __shfl_down_sync(0xffffffffu, ib, d);
          }
          if (lane + d < 32) {
            a = a + b * ao;
            b = b * bo;
            if (DUAL) {
              ia = ia + ib * iao;
              ib = ib * ibo;
            }
          }
        }
        const double x = a + b * carry;
        carry = __shfl_sync(0xffffffffu, x, 0);
        if (valid) s_r[row * LD + w] = (float)x;           // advantage overwrites its reward slot
        if (DUAL) {
          const double ix = ia + ib * icarry;
          icarry = __shfl_sync(0xffffffffu, ix, 0);
          if (valid) s_ir[row * LD + w] = (float)ix;
        }
      }
    }
    __syncthreads();
    // ---- coalesced store: advantages, returns = f32(adv) + V (buffer.py:230, :361-362) ----
    for (int i = tid; i < rows * COLS; i += C::kThreads) {
      const int row = i / COLS, c = i - row * COLS, col = c0 + c;
      if (col >= N) continue;
      const int s = row * LD + c;
      const size_t g = (size_t)(t0 + row) * N + col;
      const float av = s_r[s];
      adv[g] = av;
      ret[g] = __fadd_rn(av, s_v[s]);
      if (DUAL) {
        const float iav = s_ir[s];
        iadv[g] = iav;
        iret[g] = __fadd_rn(iav, s_iv[s]);
      }
    }
    __syncthreads();
  }
}

// Wide rollouts (N >= 64k env columns): one THREAD per env column, sequential in t in exactly the reference's
// operation order (bit-exact, including the f32 intrinsic carry), rows read fully coalesced across the warp.
// The loads of the unrolled steps do not depend on the carry chain, so ~8 rows x 3 arrays are in flight per
// thread: HBM-bound.  (Narrow rollouts keep the warp-per-column parallel scan above: not enough columns.)
template <bool DUAL>
__global__ void __launch_bounds__(128)
gae_seq_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ masks,
               const float* __restrict__ last_value, const uint8_t* __restrict__ last_done,
               const float* __restrict__ int_rewards, const float* __restrict__ int_values,
               const float* __restrict__ last_int_value, float g32, double gl, float gi32, float gil32, int T, int N,
               float* __restrict__ adv, float* __restrict__ ret, float* __restrict__ iadv, float* __restrict__ iret) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double carry = 0.0;
  float icarry = 0.f;
  float nv = last_value[n];
  double nnt = 1.0 - (double)last_done[n];
  float niv = DUAL ? last_int_value[n] : 0.f;
  constexpr int U = 8;                                        // rows fetched ahead of the dependent chain
  for (int t1 = T; t1 > 0; t1 -= U) {
    float r[U], v[U], ir[U], iv[U];
    uint8_t m[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {                             // independent loads: U x (3 | 5) in flight per thread
      const int t = t1 - 1 - u;
      if (t >= 0) {
        const size_t g = (size_t)t * N + n;
        r[u] = __ldcs(rewards + g); v[u] = __ldcs(values + g); m[u] = __ldcs(masks + g);
        if (DUAL) { ir[u] = __ldcs(int_rewards + g); iv[u] = __ldcs(int_values + g); }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t1 - 1 - u;
      if (t < 0) break;
      const size_t g = (size_t)t * N + n;
      const float gv = __fmul_rn(g32, nv);
      const double delta = __dsub_rn(__dadd_rn((double)r[u], __dmul_rn((double)gv, nnt)), (double)v[u]);
      carry = __dadd_rn(delta, __dmul_rn(__dmul_rn(gl, nnt), carry));
      const float a = (float)carry;
      __stcs(adv + g, a);
      __stcs(ret + g, __fadd_rn(a, v[u]));
      nv = v[u];
      nnt = 1.0 - (double)m[u];
      if (DUAL) {
        const float idelta = __fsub_rn(__fadd_rn(ir[u], __fmul_rn(gi32, niv)), iv[u]);
        icarry = __fadd_rn(idelta, __fmul_rn(gil32, icarry));
        __stcs(iadv + g, icarry);
        __stcs(iret + g, __fadd_rn(icarry, iv[u]));
        niv = iv[u];
      }
    }
  }
}

// thread-per-column, sequential in t: same operation order as the reference loop -> bit-exact.
__global__ void discount_kernel(const double* __restrict__ r, const uint8_t* __restrict__ d, double gamma,
                                int T, int N, double* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double carry = 0.0;
  for (int t = T - 1; t >= 0; --t) {
    const size_t g = (size_t)t * N + n;
    carry = __dadd_rn(r[g], __dmul_rn(__dmul_rn(gamma, carry), 1.0 - (double)d[g]));
    out[g] = carry;
  }
}

template <bool DUAL>
int launch_gae(const float* rewards, const float* values, const uint8_t* masks, const float* last_value,
               const uint8_t* last_done, double gamma, double lam, const float* int_rewards,
               const float* int_values, const float* last_int_value, double int_gamma, int T, int N, float* adv,
               float* ret, float* iadv, float* iret, cudaStream_t st) {
  PPX_REQUIRE(T > 0 && N > 0, "gae: T=%d N=%d must be positive", T, N);
  PPX_REQUIRE(rewards && values && masks && last_value && last_done && adv && ret, "gae: null pointer");
  if (DUAL) PPX_REQUIRE(int_rewards && int_values && last_int_value && iadv && iret, "gae_dual: null pointer");
  const float g32 = (float)gamma, gi32 = (float)int_gamma;
  const double gl = gamma * lam;
  const double gil = (double)(float)(int_gamma * lam);
  const int sms = sm_count();
  if (N >= 65536) {
    gae_seq_kernel<DUAL><<<(unsigned)ceil_div(N, 128), 128, 0, st>>>(rewards, values, masks, last_value, last_done,
        int_rewards, int_values, last_int_value, g32, gl, gi32, (float)(int_gamma * lam), T, N, adv, ret, iadv, iret);
    return after_launch(DUAL ? "gae_dual(seq)" : "gae(seq)");
  }
  // widest column tile that still gives every SM a CTA
  int cols = 32;
  if (ceil_div(N, 32) < sms) cols = 16;
  if (ceil_div(N, 16) < sms) cols = 8;
#define PPX_GAE_LAUNCH(C)                                                                                   \
  gae_kernel<C, DUAL><<<(unsigned)ceil_div(N, C), C * 32, 0, st>>>(rewards, values, masks, last_value,       \
      last_done, int_rewards, int_values, last_int_value, g32, gl, gi32, gil, T, N, adv, ret, iadv, iret)
  if (cols == 32) PPX_GAE_LAUNCH(32);
  else if (cols == 16) PPX_GAE_LAUNCH(16);
  else PPX_GAE_LAUNCH(8);
#undef PPX_GAE_LAUNCH
  return after_launch(DUAL ? "gae_dual" : "gae");
}

}  // namespace
}  // namespace ppx

extern "C" int ppx_gae(const float* rewards, const float* values, const uint8_t* masks, const float* last_value,
                       const uint8_t* last_done, double gamma, double lam, int T, int N, float* advantages,
                       float* returns, void* stream) {
  return ppx::launch_gae<false>(rewards, values, masks, last_value, last_done, gamma, lam, nullptr, nullptr, nullptr,
                                0.0, T, N, advantages, returns, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int ppx_gae_dual(const float* rewards, const float* values, const uint8_t* masks, const float* last_value,
                            const uint8_t* last_done, double gamma, double lam, const float* int_rewards,
                            const float* int_values, const float* last_int_value, double int_gamma, int T, int N,
                            float* advantages, float* returns, float* int_advantages, float* int_returns,
                            void* stream) {
  return ppx::launch_gae<true>(rewards, values, masks, last_value, last_done, gamma, lam, int_rewards, int_values,
                               last_int_value, int_gamma, T, N, advantages, returns, int_advantages, int_returns,
                               (cudaStream_t)stream);
}

extern "C" int ppx_discount(const double* rewards, const uint8_t* dones, double gamma, int T, int N, double* out,
                            void* stream) {
  PPX_REQUIRE(T > 0 && N > 0 && rewards && dones && out, "discount: bad arguments");
  ppx::discount_kernel<<<(unsigned)ppx::ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(rewards, dones, gamma, T, N, out);
  return ppx::after_launch("discount");
}
