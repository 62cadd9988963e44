// Global-norm gradient clipping fused with the Adam step on a flat parameter vector.
//
// Replaces torch.nn.utils.clip_grad_norm_ + torch.optim.Adam.step (algorithms.py:243-244, :465-466,
// :501-502, :697-699).  Two launches: per-CTA f64 sum of squares over the clipped prefix of the
// gradient, then every CTA re-reduces those partials in the same fixed order (so all agree on the
// clip coefficient without a grid sync) and applies the update.  32 B/param/step.
#include "common.cuh"

namespace ppx {
namespace {

constexpr int kNormBlocks = 512;

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ partials, int64_t* step_dev) {
  __shared__ double s_red[32];
  if (step_dev && blockIdx.x == 0 && threadIdx.x == 0) *step_dev += 1;      // read by adam_kernel (next launch)
  double s = 0.0;
  const int64_t tid0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = (((uintptr_t)g & 15) == 0) ? n / 4 : 0;
  for (int64_t q = tid0; q < n4; q += nth) {
    const float4 v = ld_stream4(reinterpret_cast<const float4*>(g) + q);
    s += (double)v.x * (double)v.x + (double)v.y * (double)v.y + (double)v.z * (double)v.z + (double)v.w * (double)v.w;
  }
  for (int64_t i = n4 * 4 + tid0; i < n; i += nth) {
    const double v = (double)g[i];
    s += v * v;
  }
  s = block_sum(s, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            const double* __restrict__ partials, int n_partials, float max_norm, int64_t n_clip, float w1 /*1-beta1*/,
            double beta1, double beta2d, double lr, float beta2, float w2 /*1-beta2*/, float step_size, float bc2_sqrt,
            float eps, const int64_t* __restrict__ step_dev, double* norm_out, const float* __restrict__ extra, int n_extra) {
  __shared__ float s_coef, s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    if (step_dev) {                                          // device-resident step counter (CUDA-graph replay)
      const double t = (double)(*step_dev);
      s_step_size = (float)(lr / (1.0 - pow(beta1, t)));
      s_bc2_sqrt = (float)sqrt(1.0 - pow(beta2d, t));
    } else {
      s_step_size = step_size;
      s_bc2_sqrt = bc2_sqrt;
    }
    float coef = 1.f;
    if (n_partials > 0) {
      double s = 0.0;
      for (int k = 0; k < n_partials; ++k) s += partials[k];
      for (int k = 0; k < n_extra; ++k) s += (double)extra[k] * (double)extra[k];   // gradients not covered by the partials
      const float norm = (float)sqrt(s);
      coef = fminf(max_norm / (norm + 1e-6f), 1.f);          // clip_grad_norm_: clamp(max_norm/(norm+1e-6), max=1)
      if (blockIdx.x == 0 && norm_out) *norm_out = sqrt(s);
    }
    s_coef = coef;
  }
  __syncthreads();
  const float coef = s_coef;
  step_size = s_step_size;
  bc2_sqrt = s_bc2_sqrt;
  auto upd = [&](float gi, float& pi, float& mi, float& vi) {
    mi = mi + w1 * (gi - mi);                                 // exp_avg.lerp_(grad, 1-beta1)
    vi = vi * beta2 + w2 * gi * gi;                           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);                       // param.addcdiv_(exp_avg, denom, -step_size)
  };
  const int64_t tid0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  for (int64_t q = tid0; q < n4; q += nth) {                  // 16-byte vectors over the aligned body
    float4 g4 = ld_stream4(reinterpret_cast<const float4*>(g) + q);
    float4 p4 = reinterpret_cast<float4*>(p)[q], m4 = reinterpret_cast<float4*>(m)[q], v4 = reinterpret_cast<float4*>(v)[q];
    const int64_t i = q * 4;
    if (i + 3 < n_clip) { g4.x *= coef; g4.y *= coef; g4.z *= coef; g4.w *= coef; }
    else {
      if (i < n_clip) g4.x *= coef;
      if (i + 1 < n_clip) g4.y *= coef;
      if (i + 2 < n_clip) g4.z *= coef;
    }
    upd(g4.x, p4.x, m4.x, v4.x); upd(g4.y, p4.y, m4.y, v4.y); upd(g4.z, p4.z, m4.z, v4.z); upd(g4.w, p4.w, m4.w, v4.w);
    reinterpret_cast<float4*>(p)[q] = p4; reinterpret_cast<float4*>(m)[q] = m4; reinterpret_cast<float4*>(v)[q] = v4;
  }
  for (int64_t i = n4 * 4 + tid0; i < n; i += nth) {
    float gi = g[i];
    if (i < n_clip) gi *= coef;
    float pi = p[i], mi = m[i], vi = v[i];
    upd(gi, pi, mi, vi);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

__global__ void bump_step_kernel(int64_t* step) { *step += 1; }

}  // namespace
}  // namespace ppx

extern "C" int ppx_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double max_norm,
                             int64_t n_clip, double lr, double beta1, double beta2, double eps, int64_t step, int64_t* step_dev, double* norm_out,
                             void* workspace, void* stream) {
  using namespace ppx;
  PPX_REQUIRE(params && grads && exp_avg && exp_avg_sq && workspace, "clip_adam: null pointer");
  PPX_REQUIRE(n >= 1 && (step >= 1 || step_dev) && n_clip >= 0 && n_clip <= n, "clip_adam: n=%lld step=%lld n_clip=%lld", (long long)n, (long long)step, (long long)n_clip);
  cudaStream_t st = (cudaStream_t)stream;
  double* partials = (double*)workspace;
  int n_partials = 0;
  const bool clip = max_norm > 0.0 && n_clip > 0;
  if (step_dev) {                                            // counter holds steps done; bump first, then use
    if (!clip) {
      bump_step_kernel<<<1, 1, 0, st>>>(step_dev);
      int rc = after_launch("clip_adam step");
      if (rc) return rc;
    }
    step = 1;
  }
  if (clip) {
    n_partials = (int)std::min<int64_t>(kNormBlocks, ceil_div(n_clip, 4096));
    sumsq_kernel<<<n_partials, 256, 0, st>>>(grads, n_clip, partials, step_dev);
    int rc = after_launch("clip_adam sumsq");
    if (rc) return rc;
  }
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const int grid = (int)std::min<int64_t>(ceil_div(n, 2048), (int64_t)sm_count() * 8);
  adam_kernel<<<grid, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, partials, n_partials, (float)max_norm, n_clip,
                                    (float)(1.0 - beta1), beta1, beta2, lr, (float)beta2, (float)(1.0 - beta2),
                                    (float)(lr / bc1), (float)sqrt(bc2), (float)eps, step_dev, norm_out, nullptr, 0);
  return after_launch("clip_adam");
}

extern "C" int ppx_clip_adam_pre(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double max_norm,
                                 double lr, double beta1, double beta2, double eps, const int64_t* step_dev, double* norm_out,
                                 const double* sumsq_partials, int n_partials, const float* extra_grads, int n_extra, void* stream) {
  using namespace ppx;
  PPX_REQUIRE(params && grads && exp_avg && exp_avg_sq && step_dev && sumsq_partials, "clip_adam_pre: null pointer");
  PPX_REQUIRE(n >= 1 && max_norm > 0.0 && n_partials >= 1 && n_extra >= 0 && (n_extra == 0 || extra_grads), "clip_adam_pre: bad arguments");
  const int grid = (int)std::min<int64_t>(ceil_div(n, 2048), (int64_t)sm_count() * 8);
  adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, sumsq_partials, n_partials, (float)max_norm, n,
                                                     (float)(1.0 - beta1), beta1, beta2, lr, (float)beta2, (float)(1.0 - beta2), 0.f, 0.f,
                                                     (float)eps, step_dev, norm_out, extra_grads, n_extra);
  return after_launch("clip_adam(pre)");
}
