// fp32 dense layers: forward, data-gradient and weight-gradient, with fused bias / activation /
// activation-derivative epilogues and strided batching (actor | critic | int_critic side by side).
//
// Replaces the nn.Linear + Tanh/LeakyReLU/ELU stacks of models.py:141-150 (policy), :220-234 (RND),
// :281-291 (ICM) and their autograd backward inside train() (algorithms.py:242, :464, :500, :696).
//
// One tiled SIMT GEMM core, C[i,j] = sum_r A(i,r) * B(r,j), instantiated for the three layouts:
//   forward   Y  = X  @ W      A = X  (r contiguous), B = W  (j contiguous)
//   dgrad     dX = dY @ W^T    A = dY (r contiguous), B = W  (r contiguous)
//   wgrad     dW = X^T @ dY    A = X  (i contiguous), B = dY (j contiguous), split over r = batch rows
// 128x64 CTA tile, 16-deep k-slab, 256 threads with an 8x4 register tile each, register-staged
// double buffering.  fp32 accumulate in a fixed order -> deterministic; the wgrad split-M partials are
// reduced by a second kernel in a fixed order (no float atomics).  This is the exact-fp32 path the
// 1e-5 parity bound needs; layers whose output width is <= 4 (value heads, action means) use
// dedicated row-dot kernels instead of wasting a 64-wide tile.
#include "common.cuh"

namespace ppx {
namespace {

constexpr int BI = 128, BJ = 64, BR = 16, NT = 256, TI = 8, TJ = 4;
constexpr int LDA_S = BI + 4, LDB_S = BJ + 4;

__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case PPX_ACT_TANH: return tanhf(x);
    case PPX_ACT_LEAKY_RELU: return x > 0.f ? x : 0.01f * x;
    case PPX_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case PPX_ACT_RELU: return fmaxf(x, 0.f);
    default: return x;
  }
}
// derivative expressed through the post-activation value h
__device__ __forceinline__ float act_bwd(float h, int act) {
  switch (act) {
    case PPX_ACT_TANH: return 1.f - h * h;
    case PPX_ACT_LEAKY_RELU: return h > 0.f ? 1.f : 0.01f;
    case PPX_ACT_ELU: return h > 0.f ? 1.f : h + 1.f;
    case PPX_ACT_RELU: return h > 0.f ? 1.f : 0.f;
    default: return 1.f;
  }
}

// up to 4 consecutive floats starting at p (first `valid` in range), zero filled
__device__ __forceinline__ float4 ld4_guard(const float* p, int valid) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid >= 4 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) return __ldg(reinterpret_cast<const float4*>(p));
  if (valid > 0) v.x = __ldg(p);
  if (valid > 1) v.y = __ldg(p + 1);
  if (valid > 2) v.z = __ldg(p + 2);
  if (valid > 3) v.w = __ldg(p + 3);
  return v;
}

struct GemmP {
  const float* A; const float* B; float* C;
  int I, J, R;                 // output rows, output cols, reduction length
  int lda, ldb, ldc;
  int64_t sA, sB, sC;          // batch strides (elements)
  int splits, r_chunk;         // wgrad: reduction split
  int64_t sSplit;              // wgrad: stride between split partials in C
  // epilogue operands
  const float* bias; int64_t sBias;
  const float* H; int ldh; int64_t sH;
  int act;
  float* bsum; int64_t sBsum;  // wgrad: column sums of B (dbias partials), stride per (batch*split)
};

enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_WGRAD = 2 };

template <bool A_R, bool B_J, int EPI>
__global__ void __launch_bounds__(NT) gemm_kernel(GemmP p) {
  __shared__ __align__(16) float As[BR * LDA_S];
  __shared__ __align__(16) float Bs[BR * LDB_S];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int i0 = blockIdx.x * BI, j0 = blockIdx.y * BJ;
  int z = blockIdx.z, split = 0;
  if (EPI == EPI_WGRAD) { split = z % p.splits; z /= p.splits; }
  const float* A = p.A + z * p.sA;
  const float* B = p.B + z * p.sB;
  int r_begin = 0, r_end = p.R;
  if (EPI == EPI_WGRAD) { r_begin = split * p.r_chunk; r_end = min(p.R, r_begin + p.r_chunk); }

  float acc[TI][TJ];
#pragma unroll
  for (int a = 0; a < TI; ++a)
#pragma unroll
    for (int b = 0; b < TJ; ++b) acc[a][b] = 0.f;
  float bsum[TJ] = {0.f, 0.f, 0.f, 0.f};
  const bool do_bsum = (EPI == EPI_WGRAD) && p.bsum != nullptr && blockIdx.x == 0 && ty == 0;

  float4 ra[2], rb;
  auto fetch = [&](int r0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int f = tid + q * NT;
      if (A_R) {                       // 4 consecutive r of one row i
        const int i = f >> 2, r4 = (f & 3) << 2;
        const int gi = i0 + i, gr = r0 + r4;
        ra[q] = (gi < p.I) ? ld4_guard(A + (int64_t)gi * p.lda + gr, r_end - gr) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {                         // 4 consecutive i of one reduction row r
        const int r = f >> 5, i4 = (f & 31) << 2;
        const int gr = r0 + r, gi = i0 + i4;
        ra[q] = (gr < r_end) ? ld4_guard(A + (int64_t)gr * p.lda + gi, p.I - gi) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (B_J) {                         // 4 consecutive j of one reduction row r
      const int r = tid >> 4, j4 = (tid & 15) << 2;
      const int gr = r0 + r, gj = j0 + j4;
      rb = (gr < r_end) ? ld4_guard(B + (int64_t)gr * p.ldb + gj, p.J - gj) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {                           // 4 consecutive r of one column j
      const int j = tid >> 2, r4 = (tid & 3) << 2;
      const int gj = j0 + j, gr = r0 + r4;
      rb = (gj < p.J) ? ld4_guard(B + (int64_t)gj * p.ldb + gr, r_end - gr) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int f = tid + q * NT;
      if (A_R) {
        const int i = f >> 2, r4 = (f & 3) << 2;
        As[(r4 + 0) * LDA_S + i] = ra[q].x; As[(r4 + 1) * LDA_S + i] = ra[q].y;
        As[(r4 + 2) * LDA_S + i] = ra[q].z; As[(r4 + 3) * LDA_S + i] = ra[q].w;
      } else {
        const int r = f >> 5, i4 = (f & 31) << 2;
        *reinterpret_cast<float4*>(&As[r * LDA_S + i4]) = ra[q];
      }
    }
    if (B_J) {
      const int r = tid >> 4, j4 = (tid & 15) << 2;
      *reinterpret_cast<float4*>(&Bs[r * LDB_S + j4]) = rb;
    } else {
      const int j = tid >> 2, r4 = (tid & 3) << 2;
      Bs[(r4 + 0) * LDB_S + j] = rb.x; Bs[(r4 + 1) * LDB_S + j] = rb.y;
      Bs[(r4 + 2) * LDB_S + j] = rb.z; Bs[(r4 + 3) * LDB_S + j] = rb.w;
    }
  };

  if (r_begin < r_end) fetch(r_begin);
  for (int r0 = r_begin; r0 < r_end; r0 += BR) {
    stash();
    __syncthreads();
    if (r0 + BR < r_end) fetch(r0 + BR);          // prefetch next slab into registers
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[r * LDA_S + ty * TI]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[r * LDA_S + ty * TI + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[r * LDB_S + tx * TJ]);
      const float av[TI] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[TJ] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int a = 0; a < TI; ++a)
#pragma unroll
        for (int c = 0; c < TJ; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
      if (do_bsum) {
#pragma unroll
        for (int c = 0; c < TJ; ++c) bsum[c] += bv[c];
      }
    }
    __syncthreads();
  }

  // ---- epilogue ----
  const int zz = blockIdx.z;
  float* C = p.C + (EPI == EPI_WGRAD ? (int64_t)zz * p.sSplit : (int64_t)z * p.sC);
  const int gj = j0 + tx * TJ;
#pragma unroll
  for (int a = 0; a < TI; ++a) {
    const int gi = i0 + ty * TI + a;
    if (gi >= p.I) continue;
    float v[TJ];
#pragma unroll
    for (int c = 0; c < TJ; ++c) {
      float x = acc[a][c];
      if (gj + c < p.J) {
        if (EPI == EPI_FWD) {
          if (p.bias) x += __ldg(p.bias + z * p.sBias + gj + c);
          x = act_fwd(x, p.act);
        } else if (EPI == EPI_DGRAD) {
          if (p.H) x *= act_bwd(__ldg(p.H + z * p.sH + (int64_t)gi * p.ldh + gj + c), p.act);
        }
      }
      v[c] = x;
    }
    float* dst = C + (int64_t)gi * p.ldc + gj;
    if (gj + 3 < p.J && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int c = 0; c < TJ; ++c)
        if (gj + c < p.J) dst[c] = v[c];
    }
  }
  if (do_bsum) {
    float* bs = p.bsum + (int64_t)zz * p.sBsum;
#pragma unroll
    for (int c = 0; c < TJ; ++c)
      if (gj + c < p.J) bs[gj + c] = bsum[c];
  }
}

// out[e] = sum_s part[s*stride + e].  32 outputs x 8 split-phases per CTA: every thread keeps its loads
// independent (in flight together), the 8 phase sums are combined through shared memory in a fixed order
// -> deterministic.  One launch reduces the weight block [0,n_w) into out_w and the bias block into out_b.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ part, int splits, int64_t stride, int64_t n_w, int64_t n_b,
                       float* __restrict__ out_w, float* __restrict__ out_b, int batch, int64_t part_bstride,
                       int64_t outw_bstride, int64_t outb_bstride) {
  __shared__ float s_p[8][33];
  const int lane = threadIdx.x & 31, ph = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * 32 + lane;
  const int z = blockIdx.y;
  const int64_t n = n_w + (out_b ? n_b : 0);
  float s = 0.f;
  if (e < n) {
    const float* p = part + z * part_bstride + e;
#pragma unroll 4
    for (int k = ph; k < splits; k += 8) s += p[(int64_t)k * stride];
  }
  s_p[ph][lane] = s;
  __syncthreads();
  if (ph == 0 && e < n) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += s_p[q][lane];
    if (e < n_w) out_w[z * outw_bstride + e] = t;
    else out_b[z * outb_bstride + (e - n_w)] = t;
  }
}

// ---------------- narrow outputs (N <= 4): value heads, action means ----------------
constexpr int SN_MAXK = 1024;

template <int NN>
__global__ void __launch_bounds__(256)
fwd_small_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, const float* __restrict__ bias, int M,
                 int K, int act, float* __restrict__ Y, int ldy, int64_t sX, int64_t sW, int64_t sB, int64_t sY) {
  __shared__ float Ws[SN_MAXK * NN];
  const int z = blockIdx.y;
  X += z * sX; W += z * sW; Y += z * sY;
  for (int e = threadIdx.x; e < K * NN; e += blockDim.x) Ws[e] = W[e];
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* x = X + (int64_t)m * ldx;
  float acc[NN];
#pragma unroll
  for (int n = 0; n < NN; ++n) acc[n] = 0.f;
  int k = 0;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    for (; k + 4 <= K; k += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + k));
      const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int n = 0; n < NN; ++n) acc[n] = fmaf(xv[e], Ws[(k + e) * NN + n], acc[n]);
    }
  }
  for (; k < K; ++k) {
    const float xv = __ldg(x + k);
#pragma unroll
    for (int n = 0; n < NN; ++n) acc[n] = fmaf(xv, Ws[k * NN + n], acc[n]);
  }
#pragma unroll
  for (int n = 0; n < NN; ++n) {
    float v = acc[n];
    if (bias) v += __ldg(bias + z * sB + n);
    Y[(int64_t)m * ldy + n] = act_fwd(v, act);
  }
}

// dW[k,n] = sum_m X[m,k] dY[m,n]; CTA owns a slab of rows, thread owns column k for a row phase
template <int NN>
__global__ void __launch_bounds__(256)
wgrad_small_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ dY, int lddy, int M, int K,
                   int rows_per_cta, float* __restrict__ part /*[z][cta][K*NN + NN]*/, int64_t sX, int64_t sDY) {
  extern __shared__ float sm[];                 // [phases][K*NN]
  const int z = blockIdx.y;
  X += z * sX; dY += z * sDY;
  const int phases = max(1, (int)blockDim.x / K);
  const int k = threadIdx.x % K, ph = threadIdx.x / K;
  const bool active = ph < phases;
  const int m0 = blockIdx.x * rows_per_cta, m1 = min(M, m0 + rows_per_cta);
  float acc[NN], bs[NN];
#pragma unroll
  for (int n = 0; n < NN; ++n) { acc[n] = 0.f; bs[n] = 0.f; }
  if (active) {
    for (int m = m0 + ph; m < m1; m += phases) {
      const float xv = __ldg(X + (int64_t)m * ldx + k);
#pragma unroll
      for (int n = 0; n < NN; ++n) {
        const float g = __ldg(dY + (int64_t)m * lddy + n);
        acc[n] = fmaf(xv, g, acc[n]);
        if (k == 0) bs[n] += g;
      }
    }
#pragma unroll
    for (int n = 0; n < NN; ++n) sm[(ph * K + k) * NN + n] = acc[n];
    if (k == 0)
#pragma unroll
      for (int n = 0; n < NN; ++n) sm[phases * K * NN + ph * NN + n] = bs[n];
  }
  __syncthreads();
  float* out = part + ((int64_t)z * gridDim.x + blockIdx.x) * (K * NN + NN);
  for (int e = threadIdx.x; e < K * NN + NN; e += blockDim.x) {
    float s = 0.f;
    if (e < K * NN) for (int q = 0; q < phases; ++q) s += sm[q * K * NN + e];
    else for (int q = 0; q < phases; ++q) s += sm[phases * K * NN + q * NN + (e - K * NN)];
    out[e] = s;
  }
}

int wgrad_splits(int M, int K, int N, int batch) {
  const int64_t tiles = ceil_div(K, BI) * ceil_div(N, BJ) * batch;
  int64_t want = ceil_div(2 * (int64_t)sm_count(), tiles);
  const int64_t max_splits = std::max<int64_t>(1, M / 256);
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  return (int)want;
}
int wgrad_small_ctas(int M) {
  int64_t c = ceil_div(M, 512);
  const int64_t cap = 2 * (int64_t)sm_count();
  if (c > cap) c = cap;
  return (int)std::max<int64_t>(1, c);
}
bool use_small(int K, int N) { return N <= 4 && K <= SN_MAXK; }

}  // namespace
}  // namespace ppx

using namespace ppx;

extern "C" int ppx_linear_fwd(const float* X, int ldx, const float* W, const float* bias, int M, int K, int N, int act,
                              float* Y, int ldy, int batch, int64_t strideX, int64_t strideW, int64_t strideB,
                              int64_t strideY, void* stream) {
  PPX_REQUIRE(X && W && Y, "linear_fwd: null pointer");
  PPX_REQUIRE(M >= 0 && K > 0 && N > 0 && batch >= 1 && ldx >= K && ldy >= N, "linear_fwd: bad shape M=%d K=%d N=%d ldx=%d ldy=%d", M, K, N, ldx, ldy);
  if (M == 0) return PPX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (use_small(K, N)) {
    dim3 grid((unsigned)ceil_div(M, 256), (unsigned)batch);
#define PPX_FS(NN) fwd_small_kernel<NN><<<grid, 256, 0, st>>>(X, ldx, W, bias, M, K, act, Y, ldy, strideX, strideW, strideB, strideY)
    switch (N) { case 1: PPX_FS(1); break; case 2: PPX_FS(2); break; case 3: PPX_FS(3); break; default: PPX_FS(4); }
#undef PPX_FS
    return after_launch("linear_fwd(small)");
  }
  GemmP p{};
  p.A = X; p.B = W; p.C = Y; p.I = M; p.J = N; p.R = K; p.lda = ldx; p.ldb = N; p.ldc = ldy;
  p.sA = strideX; p.sB = strideW; p.sC = strideY; p.bias = bias; p.sBias = strideB; p.act = act; p.splits = 1;
  dim3 grid((unsigned)ceil_div(M, BI), (unsigned)ceil_div(N, BJ), (unsigned)batch);
  gemm_kernel<true, true, EPI_FWD><<<grid, NT, 0, st>>>(p);
  return after_launch("linear_fwd");
}

extern "C" int ppx_linear_bwd_data(const float* dY, int lddy, const float* W, int M, int K, int N, const float* H, int ldh,
                                   int act, float* dX, int lddx, int batch, int64_t strideDY, int64_t strideW,
                                   int64_t strideH, int64_t strideDX, void* stream) {
  PPX_REQUIRE(dY && W && dX, "linear_bwd_data: null pointer");
  PPX_REQUIRE(M >= 0 && K > 0 && N > 0 && batch >= 1 && lddy >= N && lddx >= K, "linear_bwd_data: bad shape");
  if (M == 0) return PPX_OK;
  GemmP p{};
  p.A = dY; p.B = W; p.C = dX; p.I = M; p.J = K; p.R = N; p.lda = lddy; p.ldb = N; p.ldc = lddx;
  p.sA = strideDY; p.sB = strideW; p.sC = strideDX; p.H = (act == PPX_ACT_NONE) ? nullptr : H; p.ldh = ldh; p.sH = strideH;
  p.act = act; p.splits = 1;
  dim3 grid((unsigned)ceil_div(M, BI), (unsigned)ceil_div(K, BJ), (unsigned)batch);
  gemm_kernel<true, false, EPI_DGRAD><<<grid, NT, 0, (cudaStream_t)stream>>>(p);
  return after_launch("linear_bwd_data");
}

extern "C" int64_t ppx_linear_bwd_weight_workspace(int M, int K, int N, int batch) {
  if (use_small(K, N)) return (int64_t)batch * wgrad_small_ctas(M) * ((int64_t)K * N + N);
  return (int64_t)batch * wgrad_splits(M, K, N, batch) * ((int64_t)K * N + N);
}

extern "C" int ppx_linear_bwd_weight(const float* X, int ldx, const float* dY, int lddy, int M, int K, int N, float* dW,
                                     float* dbias, float* workspace, int batch, int64_t strideX, int64_t strideDY,
                                     int64_t strideDW, int64_t strideDB, void* stream) {
  PPX_REQUIRE(X && dY && dW && workspace, "linear_bwd_weight: null pointer");
  PPX_REQUIRE(M > 0 && K > 0 && N > 0 && batch >= 1 && ldx >= K && lddy >= N, "linear_bwd_weight: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t per = (int64_t)K * N + N;
  if (use_small(K, N)) {
    const int ctas = wgrad_small_ctas(M);
    const int rows = (int)ceil_div(M, ctas);
    const int threads = K >= 256 ? ((K + 31) / 32) * 32 : 256;
    PPX_REQUIRE(threads <= 1024, "linear_bwd_weight(small): K=%d too wide", K);
    const int phases = std::max(1, threads / K);
    const size_t smem = (size_t)(phases * K * N + phases * N) * sizeof(float);
    dim3 grid((unsigned)ctas, (unsigned)batch);
#define PPX_WS(NN) wgrad_small_kernel<NN><<<grid, threads, smem, st>>>(X, ldx, dY, lddy, M, K, rows, workspace, strideX, strideDY)
    switch (N) { case 1: PPX_WS(1); break; case 2: PPX_WS(2); break; case 3: PPX_WS(3); break; default: PPX_WS(4); }
#undef PPX_WS
    int rc = after_launch("linear_bwd_weight(small)");
    if (rc) return rc;
    dim3 rg((unsigned)ceil_div(per, 32), (unsigned)batch);
    reduce_partials_kernel<<<rg, 256, 0, st>>>(workspace, ctas, per, (int64_t)K * N, N, dW, dbias, batch, (int64_t)ctas * per,
                                               strideDW, strideDB);
    return after_launch("linear_bwd_weight(reduce)");
  }
  const int splits = wgrad_splits(M, K, N, batch);
  GemmP p{};
  p.A = X; p.B = dY; p.I = K; p.J = N; p.R = M; p.lda = ldx; p.ldb = lddy; p.ldc = N;
  p.sA = strideX; p.sB = strideDY; p.splits = splits; p.r_chunk = (int)(ceil_div(ceil_div(M, splits), BR) * BR);
  p.C = workspace; p.sSplit = per; p.bsum = dbias ? workspace + (int64_t)K * N : nullptr; p.sBsum = per;
  dim3 grid((unsigned)ceil_div(K, BI), (unsigned)ceil_div(N, BJ), (unsigned)(batch * splits));
  gemm_kernel<false, true, EPI_WGRAD><<<grid, NT, 0, st>>>(p);
  int rc = after_launch("linear_bwd_weight");
  if (rc) return rc;
  dim3 rg((unsigned)ceil_div(per, 32), (unsigned)batch);
  reduce_partials_kernel<<<rg, 256, 0, st>>>(workspace, splits, per, (int64_t)K * N, N, dW, dbias, batch, (int64_t)splits * per,
                                             strideDW, strideDB);
  return after_launch("linear_bwd_weight(reduce)");
}
