// Minibatch shuffle-gather and advantage statistics.
//
// Replaces RolloutStorage.get/_get_samples (buffer.py:233-267, :365-394).  The reference first
// rewrites every array into the env-major flat layout (swap_and_flatten, buffer.py:40-52) and then
// fancy-indexes it; here the [T,N,...] arrays stay where they are and the flat index is decoded on
// the fly: flat i -> (t, n) = (i % T, i / T).  One launch gathers every field of the RolloutSample.
// Algorithmic traffic: 2 x row bytes per sample + the 8-byte index.
#include "common.cuh"

namespace ppx {
namespace {

struct GatherArgs {
  const char* src[PPX_MAX_GATHER];
  char* dst[PPX_MAX_GATHER];
  int row_bytes[PPX_MAX_GATHER];
  int vec16[PPX_MAX_GATHER];
};

template <typename V>
__device__ __forceinline__ void gather_rows(const char* __restrict__ src, char* __restrict__ dst, int row_bytes,
                                            const int64_t* __restrict__ idx, int64_t B, int T, int N) {
  const int words = row_bytes / (int)sizeof(V);
  const int64_t total = B * words;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = g / words;
    const int w = (int)(g - b * words);
    const int64_t i = __ldg(idx + b);
    const int64_t row = (i % T) * N + i / T;
    reinterpret_cast<V*>(dst)[g] = __ldg(reinterpret_cast<const V*>(src + row * row_bytes) + w);
  }
}

__global__ void __launch_bounds__(256)
gather_kernel(GatherArgs args, const int64_t* __restrict__ idx, int64_t B, int T, int N) {
  const int a = blockIdx.y;
  if (args.vec16[a]) gather_rows<int4>(args.src[a], args.dst[a], args.row_bytes[a], idx, B, T, N);
  else gather_rows<int>(args.src[a], args.dst[a], args.row_bytes[a], idx, B, T, N);
}

// one CTA, two passes (mean, then centred sum of squares), f64 accumulation, fixed order -> deterministic
__global__ void __launch_bounds__(1024) mean_std_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ out) {
  __shared__ double s_red[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i];
  const double mean = block_sum(s, s_red) / (double)n;
  double q = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = (double)x[i] - mean;
    q += d * d;
  }
  q = block_sum(q, s_red);
  if (threadIdx.x == 0) {
    out[0] = mean;
    out[1] = sqrt(q / (double)(n - 1));          // unbiased, like torch.Tensor.std()
  }
}

}  // namespace
}  // namespace ppx

extern "C" int ppx_gather_minibatch(const void* const* srcs_host, void* const* dsts_host, const int* row_bytes_host,
                                    int n_arrays, const int64_t* idx, int64_t B, int T, int N, void* stream) {
  PPX_REQUIRE(srcs_host && dsts_host && row_bytes_host && idx, "gather_minibatch: null pointer");
  PPX_REQUIRE(n_arrays >= 1 && n_arrays <= PPX_MAX_GATHER, "gather_minibatch: n_arrays=%d (1..%d)", n_arrays, PPX_MAX_GATHER);
  PPX_REQUIRE(B >= 0 && T > 0 && N > 0, "gather_minibatch: B=%lld T=%d N=%d", (long long)B, T, N);
  if (B == 0) return PPX_OK;
  ppx::GatherArgs args;
  int64_t max_words = 0;
  for (int a = 0; a < n_arrays; ++a) {
    PPX_REQUIRE(srcs_host[a] && dsts_host[a], "gather_minibatch: array %d null", a);
    PPX_REQUIRE(row_bytes_host[a] > 0 && row_bytes_host[a] % 4 == 0, "gather_minibatch: row_bytes[%d]=%d must be a positive multiple of 4", a, row_bytes_host[a]);
    args.src[a] = (const char*)srcs_host[a];
    args.dst[a] = (char*)dsts_host[a];
    args.row_bytes[a] = row_bytes_host[a];
    args.vec16[a] = (row_bytes_host[a] % 16 == 0) && ((uintptr_t)srcs_host[a] % 16 == 0) && ((uintptr_t)dsts_host[a] % 16 == 0);
    const int64_t words = B * (row_bytes_host[a] / (args.vec16[a] ? 16 : 4));
    if (words > max_words) max_words = words;
  }
  int64_t gx = ppx::ceil_div(max_words, 256);
  const int64_t cap = (int64_t)ppx::sm_count() * 16;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)n_arrays);
  ppx::gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(args, idx, B, T, N);
  return ppx::after_launch("gather_minibatch");
}

extern "C" int ppx_mean_std(const float* x, int64_t n, double* out2, void* stream) {
  PPX_REQUIRE(x && out2 && n >= 1, "mean_std: bad arguments");
  ppx::mean_std_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, out2);
  return ppx::after_launch("mean_std");
}
