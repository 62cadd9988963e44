// Blackwell tensor-core dense layer:  C[M,N] = epi( A[M,R] . B[N,R]^T )  in fp32-equivalent precision.
//
// tcgen05.mma kind::tf32 with a 3-pass hi/lo split ("3xTF32"): x = hi + lo with hi = x truncated to
// tf32 (exact), lo = x - hi (exact in fp32, tf32-truncated by the tensor core):
//      A.B  ~=  A_lo.B_hi + A_hi.B_lo + A_hi.B_hi          (dropped term lo.lo ~ 2^-22 relative)
// accumulated in fp32 in TMEM.  Error ~2^-21 per product, inside the 1e-5 parity bound that rules out
// single-pass tf32/bf16 (SURVEY "Hard parts").  The tensor core's fp32 accumulator TRUNCATES, so a long
// reduction drifts (measured 2.4e-5 relative at K = 3136): the reduction is therefore cut into chunks of
// 128 products that ping-pong between two TMEM accumulators; the epilogue warps drain each finished chunk
// (tcgen05.ld) and add it into fp32 registers with round-to-nearest while the next chunk's MMAs run.  Used for the forward (A = activations, B = W^T) and the
// data-gradient (A = dY, B = W) of dense layers; B is pre-split once per optimiser step
// (ppx_tc_split), A is split on the fly in shared memory so activations are read from HBM once.
//
// Pipeline (one 128 x BN output tile per CTA, 320 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads of A (raw), B_hi, B_lo, 128B-swizzled,
//               K-major, 32 fp32 (=128 B) of K per stage, mbarrier complete_tx
//   warps 2-5   splitters: in-place hi = x & ~0x1fff, lo -> second buffer (element-wise, so the TMA
//               swizzle pattern is preserved), fence.proxy.async, arrive
//   warp 1      MMA issuer: one elected lane issues 4 k-steps x 3 tcgen05.mma (M=128, N=BN, K=8) per
//               stage; tcgen05.commit releases the stage / signals the epilogue; owns TMEM alloc/dealloc
//   warps 6-9   epilogue: per 128-product chunk tcgen05.ld 32x32b.x32 (warp q reads TMEM lanes 32q..32q+31 =
//               tile rows) + register accumulation; at the end bias+activation (forward) or
//               activation-derivative (dgrad), 128-byte row stores
#include "tc_common.cuh"

namespace ppx {
namespace tc {

constexpr int BK = 32;           // fp32 elements of K per stage = one 128-byte swizzle row
constexpr int UMMA_K = 8;        // tf32
constexpr int kThreads = 320;
constexpr int kSplitThreads = 128;
constexpr int KBC = 4;           // k-blocks (of BK) per TMEM accumulation chunk = 128 products

__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case PPX_ACT_TANH: return tanhf(x);
    case PPX_ACT_LEAKY_RELU: return x > 0.f ? x : 0.01f * x;
    case PPX_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case PPX_ACT_RELU: return fmaxf(x, 0.f);
    default: return x;
  }
}
__device__ __forceinline__ float act_bwd(float h, int act) {
  switch (act) {
    case PPX_ACT_TANH: return 1.f - h * h;
    case PPX_ACT_LEAKY_RELU: return h > 0.f ? 1.f : 0.01f;
    case PPX_ACT_ELU: return h > 0.f ? 1.f : h + 1.f;
    case PPX_ACT_RELU: return h > 0.f ? 1.f : 0.f;
    default: return 1.f;
  }
}

struct Params {
  int M, N, R;                  // C is [M,N]; reduction length R
  int ldc;
  float* C;
  const float* bias;            // forward epilogue (may be null)
  const float* H; int ldh;      // dgrad epilogue: post-activation of the producer layer (may be null)
  int act;
  int dgrad;                    // 0: C = act(acc + bias)   1: C = acc * act'(H)
  const double* a_mean;         // optional per-column transform of A before the hi/lo split (RND observation
  const double* a_istd;         //   normalisation, algorithms.py:111-118): A' = clip((A - mean) * istd, +-a_clip) in f64
  float a_clip;
  float* part; int ldp;         // split-K (gridDim.z > 1): raw partial accumulators [gridDim.z][M][ldp], finished by splitk_finish_kernel
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBhi,
               const __grid_constant__ CUtensorMap mapBlo, Params p) {
  constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128;
  constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));   // two accumulators
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[3 * STAGES + 4];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;     // N tiles of one row block run side by side (A tile shared through L2)
  // split-K: CTA z of gridDim.z works on k-blocks [kb0, kb0 + num_kb) (few M x N tiles and a long reduction -- the wide
  // first layers at minibatch size 4096 have 32 tiles for 148 SMs and 882 k-blocks)
  const int num_kb_all = (p.R + BK - 1) / BK;
  const int kb_per = (num_kb_all + (int)gridDim.z - 1) / (int)gridDim.z;
  const int kb0 = (int)blockIdx.z * kb_per;
  const int num_kb = max(0, min(num_kb_all, kb0 + kb_per) - kb0);

  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto ready_bar = [&](int s) { return smem_u32(&bars[STAGES + s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[2 * STAGES + s]); };
  auto tmem_full_bar = [&](int b) { return smem_u32(&bars[3 * STAGES + b]); };
  auto tmem_empty_bar = [&](int b) { return smem_u32(&bars[3 * STAGES + 2 + b]); };
  const int num_chunks = (num_kb + KBC - 1) / KBC;
  auto a_hi = [&](int s) { return smem + (size_t)s * STAGE_BYTES; };
  auto a_lo = [&](int s) { return smem + (size_t)s * STAGE_BYTES + A_BYTES; };
  auto b_hi = [&](int s) { return smem + (size_t)s * STAGE_BYTES + 2 * A_BYTES; };
  auto b_lo = [&](int s) { return smem + (size_t)s * STAGE_BYTES + 2 * A_BYTES + B_BYTES; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(ready_bar(s), kSplitThreads);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full_bar(b), 1); mbar_init(tmem_empty_bar(b), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation by one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_expect_tx(full_bar(s), A_BYTES + 2 * B_BYTES);
        tma_load_2d(smem_u32(a_hi(s)), &mapA, full_bar(s), (kb0 + kb) * BK, m0);
        tma_load_2d(smem_u32(b_hi(s)), &mapBhi, full_bar(s), (kb0 + kb) * BK, n0);
        tma_load_2d(smem_u32(b_lo(s)), &mapBlo, full_bar(s), (kb0 + kb) * BK, n0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    constexpr uint32_t idesc = make_idesc(BN);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      const int chunk = kb / KBC, buf = chunk & 1;
      const bool chunk_first = (kb % KBC) == 0, chunk_last = (kb % KBC) == KBC - 1 || kb == num_kb - 1;
      if (chunk_first) {                                      // the epilogue must have drained this accumulator
        mbar_wait(tmem_empty_bar(buf), ((chunk >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      mbar_wait(ready_bar(s), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t ah = smem_u32(a_hi(s)), al = smem_u32(a_lo(s)), bh = smem_u32(b_hi(s)), bl = smem_u32(b_lo(s));
        const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          const uint32_t off = kk * UMMA_K * 4;            // 32 bytes along K inside the swizzle atom
          const uint64_t dah = make_desc(ah + off), dal = make_desc(al + off);
          const uint64_t dbh = make_desc(bh + off), dbl = make_desc(bl + off);
          umma_tf32(tacc, dal, dbh, idesc, (chunk_first && kk == 0) ? 0u : 1u);   // small terms first
          umma_tf32(tacc, dah, dbl, idesc, 1u);
          umma_tf32(tacc, dah, dbh, idesc, 1u);
        }
        umma_commit(empty_bar(s));                          // stage free once these MMAs have read it
        if (chunk_last) umma_commit(tmem_full_bar(buf));    // this chunk's accumulator is complete
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ------------------------------ splitters (warps 2..5) ------------------------------
    const int t = threadIdx.x - 64;
    __shared__ double s_norm[STAGES][2][BK];                 // mean | istd of the stage's 32 columns
    const bool norm = p.a_mean != nullptr;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      if (norm) {                                            // fetch the column constants while the TMA is in flight
        if (t < 2 * BK) {
          const int k = (kb0 + kb) * BK + (t & (BK - 1));
          const double* src = (t < BK) ? p.a_mean : p.a_istd;
          s_norm[s][t >> 5][t & (BK - 1)] = k < p.R ? __ldg(src + k) : 0.0;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");       // splitter warps only
      }
      mbar_wait(full_bar(s), ph);
      uint4* hi = reinterpret_cast<uint4*>(a_hi(s));
      uint4* lo = reinterpret_cast<uint4*>(a_lo(s));
#pragma unroll
      for (int i = 0; i < (int)(A_BYTES / 16) / kSplitThreads; ++i) {
        const int e = t + i * kSplitThreads;
        uint4 v = hi[e];
        if (norm) {
          // physical 16-byte chunk e of the 128B-swizzled tile: row = e / 8, logical chunk = (e % 8) ^ (row % 8)
          const int row = e >> 3, c4 = (((e & 7) ^ (row & 7)) << 2);
          const double* mu = &s_norm[s][0][c4];
          const double* is = &s_norm[s][1][c4];
          const float cl = p.a_clip;
          v.x = __float_as_uint(fminf(fmaxf((float)(((double)__uint_as_float(v.x) - mu[0]) * is[0]), -cl), cl));
          v.y = __float_as_uint(fminf(fmaxf((float)(((double)__uint_as_float(v.y) - mu[1]) * is[1]), -cl), cl));
          v.z = __float_as_uint(fminf(fmaxf((float)(((double)__uint_as_float(v.z) - mu[2]) * is[2]), -cl), cl));
          v.w = __float_as_uint(fminf(fmaxf((float)(((double)__uint_as_float(v.w) - mu[3]) * is[3]), -cl), cl));
        }
        uint4 h, l;
        h.x = v.x & 0xFFFFE000u; h.y = v.y & 0xFFFFE000u; h.z = v.z & 0xFFFFE000u; h.w = v.w & 0xFFFFE000u;
        l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
        l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
        l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
        l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
        hi[e] = h;
        lo[e] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      mbar_arrive(ready_bar(s));
    }
  } else {
    // ------------------------------ epilogue (warps 6..9) ------------------------------
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;
    for (int chunk = 0; chunk < num_chunks; ++chunk) {
      const int buf = chunk & 1;
      mbar_wait(tmem_full_bar(buf), (chunk >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);     // round-to-nearest fp32 adds
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(buf));
    }
    if (row < p.M && p.part) {                               // split-K: raw partial sums; bias / activation in the finish kernel
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float* dst = p.part + ((size_t)blockIdx.z * p.M + row) * p.ldp + n0 + c0;
        if (n0 + c0 + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(acc[c0 + j], acc[c0 + j + 1], acc[c0 + j + 2], acc[c0 + j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < p.N) dst[j] = acc[c0 + j];
        }
      }
    } else if (row < p.M) {
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float* dst = p.C + (size_t)row * p.ldc + n0 + c0;
        const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (n0 + c0 + 32 <= p.N);
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = n0 + c0 + j;
          float x = acc[c0 + j];
          if (col < p.N) {
            if (!p.dgrad) {
              if (p.bias) x += __ldg(p.bias + col);
              x = act_fwd(x, p.act);
            } else if (p.H) {
              x *= act_bwd(__ldg(p.H + (size_t)row * p.ldh + col), p.act);
            }
          }
          o[j] = x;
        }
        if (vec) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < p.N) dst[j] = o[j];
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// hi/lo split of a matrix, optionally transposed: src [rows, cols] (row pitch ld) -> hi/lo [rows, cols] and/or
// hiT/loT [cols, rows]; rawT [cols, rows] = plain transpose (the on-the-fly-split A operand of the wgrad GEMM)
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ src, int ld, int rows, int cols, float* __restrict__ hi, float* __restrict__ lo,
             float* __restrict__ hiT, float* __restrict__ loT, float* __restrict__ rawT) {
  __shared__ float th[32][33], tl[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i;
    float h = 0.f, l = 0.f;
    if (r < rows && c < cols) {
      const float v = src[(size_t)r * ld + c];
      if (rawT) { h = v; }
      else {
        h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        l = v - h;
        if (hi) { hi[(size_t)r * cols + c] = h; lo[(size_t)r * cols + c] = l; }
      }
    }
    th[i][threadIdx.x] = h;
    tl[i][threadIdx.x] = l;
  }
  __syncthreads();
  if (hiT || rawT) {
    const int rr = r0 + threadIdx.x;                         // transposed: output row = source column
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int cc = blockIdx.x * 32 + i;
      if (rr < rows && cc < cols) {
        if (rawT) rawT[(size_t)cc * rows + rr] = th[threadIdx.x][i];
        else {
          hiT[(size_t)cc * rows + rr] = th[threadIdx.x][i];
          loT[(size_t)cc * rows + rr] = tl[threadIdx.x][i];
        }
      }
    }
  }
}

// split-K tail: C = epi(sum over the S partials, in split order) -- same epilogue as the single-pass kernel
__global__ void __launch_bounds__(256) splitk_finish_kernel(Params p, int S) {
  const int64_t total4 = (int64_t)p.M * (p.ldp / 4);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(e / (p.ldp / 4)), col = (int)(e % (p.ldp / 4)) * 4;
    if (col >= p.N) continue;
    float4 a = *reinterpret_cast<const float4*>(p.part + (size_t)row * p.ldp + col);
    for (int z = 1; z < S; ++z) {
      const float4 b = *reinterpret_cast<const float4*>(p.part + ((size_t)z * p.M + row) * p.ldp + col);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    float o[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (col + j >= p.N) continue;
      float x = o[j];
      if (!p.dgrad) {
        if (p.bias) x += __ldg(p.bias + col + j);
        x = act_fwd(x, p.act);
      } else if (p.H) {
        x *= act_bwd(__ldg(p.H + (size_t)row * p.ldh + col + j), p.act);
      }
      p.C[(size_t)row * p.ldc + col + j] = x;
    }
  }
}

// out[c] = sum over rows of src[r, c], fixed order: block = 32 columns x 8 row slices, slices combined 0..7
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ src, int ld, int rows, int cols, float* __restrict__ out) {
  __shared__ float sl[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
  float s = 0.f;
  if (c < cols)
    for (int r = slice; r < rows; r += 8) s += src[(size_t)r * ld + c];
  sl[slice][threadIdx.x & 31] = s;
  __syncthreads();
  if (slice == 0 && c < cols) {
#pragma unroll
    for (int q = 1; q < 8; ++q) s += sl[q][threadIdx.x & 31];
    out[c] = s;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)p;
  }
  return fn;
}

// 2D fp32 tensor [rows, cols] with row pitch ld (elements); box = 32 cols x box_rows, 128B swizzle, zero OOB fill
static int make_map(CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_rows) {
  EncodeFn enc = get_encode();
  if (!enc) return fail(PPX_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PPX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld);
  return PPX_OK;
}

// How many ways to cut the reduction: minimise waves(S) / S (+ a little per split for the finish pass) over S <= 8 with at
// least 16 k-blocks per split; 1 when the tiles already fill the machine.
static int pick_splits(int M, int R, int N, int BN) {
  const int tiles = ceil_div(M, BM) * ceil_div(N, BN), sms = sm_count(), nkb = ceil_div(R, BK);
  if (tiles >= 4 * sms) return 1;                             // many waves: the quantisation loss is small
  int best = 1;
  double best_cost = (double)ceil_div(tiles, sms);
  for (int S = 2; S <= 8 && nkb / S >= 16; ++S) {
    const double cost = (double)ceil_div(tiles * S, sms) / S * (1.0 + 0.03 * (S - 1));
    if (cost < best_cost * 0.97) { best = S; best_cost = cost; }
  }
  return best;
}

template <int BN, int STAGES>
static int launch(const CUtensorMap& ma, const CUtensorMap& mbh, const CUtensorMap& mbl, const Params& p, cudaStream_t st, int S = 1) {
  constexpr size_t smem = (size_t)STAGES * (2 * BM * 128 + 2 * BN * 128) + 1024;
  static bool configured = false;
  if (!configured) {
    PPX_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(p.N, BN), (unsigned)ceil_div(p.M, BM), (unsigned)S);
  tc_gemm_kernel<BN, STAGES><<<grid, kThreads, smem, st>>>(ma, mbh, mbl, p);
  int rc = after_launch("tc_gemm");
  if (rc || S == 1) return rc;
  const int64_t total4 = (int64_t)p.M * (p.ldp / 4);
  splitk_finish_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total4, 256), (int64_t)sm_count() * 8), 256, 0, st>>>(p, S);
  return after_launch("tc_gemm(split-K finish)");
}

}  // namespace tc
}  // namespace ppx

using namespace ppx;

extern "C" int ppx_tc_supported(int M, int R, int N, int lda, int ldb, const void* A, const void* B) {
  if (R < 4 || N < 16 || M < 1) return 0;
  if (R % 4 || lda % 4 || ldb % 4) return 0;                 // TMA: 16-byte pitches
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return 0;
  return 1;
}

extern "C" int ppx_tc_split(const float* src, int rows, int cols, float* hi, float* lo, float* hiT, float* loT, void* stream) {
  PPX_REQUIRE(src && rows > 0 && cols > 0 && ((hi && lo) || (hiT && loT)), "tc_split: bad arguments");
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32)), block(32, 8);
  tc::split_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, cols, rows, cols, hi, lo, hiT, loT, nullptr);
  return after_launch("tc_split");
}

static inline int splitk_ldp(int N) { return (N + 3) & ~3; }

extern "C" int64_t ppx_tc_linear_workspace(int M, int R, int N) {
  if (M < 1 || R < 4 || N < 16) return 0;
  const int S = tc::pick_splits(M, R, N, N <= 64 ? 64 : 128);
  return S == 1 ? 0 : (int64_t)S * M * splitk_ldp(N);
}

extern "C" int ppx_tc_linear(const float* A, int lda, const float* Bhi, const float* Blo, int ldb, int M, int R, int N,
                             const float* bias, const float* H, int ldh, int act, int dgrad, const double* a_mean,
                             const double* a_istd, float a_clip, float* C, int ldc, void* stream) {
  return ppx_tc_linear_ws(A, lda, Bhi, Blo, ldb, M, R, N, bias, H, ldh, act, dgrad, a_mean, a_istd, a_clip, C, ldc, nullptr, 0, stream);
}

extern "C" int ppx_tc_linear_ws(const float* A, int lda, const float* Bhi, const float* Blo, int ldb, int M, int R, int N,
                                const float* bias, const float* H, int ldh, int act, int dgrad, const double* a_mean,
                                const double* a_istd, float a_clip, float* C, int ldc, float* workspace, int64_t workspace_floats,
                                void* stream) {
  PPX_REQUIRE(A && Bhi && Blo && C, "tc_linear: null pointer");
  PPX_REQUIRE(ppx_tc_supported(M, R, N, lda, ldb, A, Bhi) && !((uintptr_t)Blo & 15), "tc_linear: shape/alignment not supported (M=%d R=%d N=%d lda=%d ldb=%d)", M, R, N, lda, ldb);
  const int BN = N <= 64 ? 64 : 128;
  CUtensorMap ma, mbh, mbl;
  int rc = tc::make_map(&ma, A, M, R, lda, tc::BM);
  if (rc) return rc;
  rc = tc::make_map(&mbh, Bhi, N, R, ldb, BN);
  if (rc) return rc;
  rc = tc::make_map(&mbl, Blo, N, R, ldb, BN);
  if (rc) return rc;
  PPX_REQUIRE((a_mean == nullptr) == (a_istd == nullptr), "tc_linear: a_mean / a_istd must be given together");
  tc::Params p{M, N, R, ldc, C, bias, H, ldh, act, dgrad, a_mean, a_istd, a_clip, nullptr, 0};
  int S = 1;
  if (workspace && !((uintptr_t)workspace & 15)) {            // split-K only with a workspace of ppx_tc_linear_workspace() floats
    S = tc::pick_splits(M, R, N, BN);
    if (S > 1 && workspace_floats >= (int64_t)S * M * splitk_ldp(N)) { p.part = workspace; p.ldp = splitk_ldp(N); }
    else S = 1;
  }
  if (BN == 64) return tc::launch<64, 4>(ma, mbh, mbl, p, (cudaStream_t)stream, S);
  return tc::launch<128, 3>(ma, mbh, mbl, p, (cudaStream_t)stream, S);
}

// Weight gradient of a wide layer on the tensor cores:  dW [K,N] = X^T [K,M] . dY [M,N]  (+ dbias = colsum dY).
// The reduction runs over the M samples, so both operands are needed K-major in M: X is transposed once
// (Xt [K,M], split hi/lo on the fly as the A operand), dY is split and transposed (hiT/loT [N,M], the B operand), and
// the same 3xTF32 kernel as the forward produces dW row-major [K,N] -- the in-major weight layout.
extern "C" int64_t ppx_tc_wgrad_workspace(int M, int K, int N) {
  return (int64_t)M * K + 2 * (int64_t)M * N + ppx_tc_linear_workspace(K, M, N);       // + the split-K partials of the GEMM
}

extern "C" int ppx_tc_wgrad_supported(int M, int K, int N, const void* X, const void* dY) {
  return (M >= 256 && M % 4 == 0 && K >= 128 && N >= 16 && X && dY) ? 1 : 0;
}

extern "C" int ppx_tc_wgrad(const float* X, int ldx, const float* dY, int lddy, int M, int K, int N, float* dW, float* dbias,
                            float* workspace, void* stream) {
  PPX_REQUIRE(X && dY && dW && workspace && ldx >= K && lddy >= N, "tc_wgrad: bad arguments");
  PPX_REQUIRE(ppx_tc_wgrad_supported(M, K, N, X, dY) && (((uintptr_t)workspace | (uintptr_t)dW) & 15) == 0, "tc_wgrad: shape/alignment not supported (M=%d K=%d N=%d)", M, K, N);
  cudaStream_t st = (cudaStream_t)stream;
  float* Xt = workspace;
  float* dYhiT = Xt + (size_t)M * K;
  float* dYloT = dYhiT + (size_t)M * N;
  dim3 block(32, 8);
  tc::split_kernel<<<dim3((unsigned)ceil_div(K, 32), (unsigned)ceil_div(M, 32)), block, 0, st>>>(X, ldx, M, K, nullptr, nullptr, nullptr, nullptr, Xt);
  tc::split_kernel<<<dim3((unsigned)ceil_div(N, 32), (unsigned)ceil_div(M, 32)), block, 0, st>>>(dY, lddy, M, N, nullptr, nullptr, dYhiT, dYloT, nullptr);
  int rc = after_launch("tc_wgrad(transpose/split)", 2);
  if (rc) return rc;
  if (dbias) {
    tc::colsum_kernel<<<(unsigned)ceil_div(N, 32), 256, 0, st>>>(dY, lddy, M, N, dbias);
    rc = after_launch("tc_wgrad(colsum)");
    if (rc) return rc;
  }
  float* part = dYloT + (size_t)M * N;                        // M % 4 == 0: 16-byte aligned like the workspace
  return ppx_tc_linear_ws(Xt, M, dYhiT, dYloT, M, K, M, N, nullptr, nullptr, 0, PPX_ACT_NONE, 0, nullptr, nullptr, 0.f, dW, N,
                          part, ppx_tc_linear_workspace(K, M, N), stream);
}
