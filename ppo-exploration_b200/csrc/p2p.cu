// Shard-boundary exchanges of the sharded PPO update as kernels over NVLink peer memory (no NCCL on the path).
//
// The sharded update has three global points per minibatch, each a few hundred bytes to
// tens of kilobytes -- latency, not bandwidth: the advantage moments (48 B/rank), the loss partial sums
// (256 B/rank) and the flat policy gradient (39 KB/rank at h=64).  Each is fused with its consumer into ONE
// kernel that every rank launches on its own GPU:
//   p2p_moments_merge   barrier -> read the W {n, mean, M2} records from peer memory -> merged {mean, std}
//   p2p_finalize        barrier -> sum the W x 32 loss partial sums -> loss scalars, max-of-means branch, d_log_std
//   p2p_clip_adam       barrier -> g = sum over ranks of the peers' gradient vectors (rank order, so every replica
//                       computes the identical sum) -> global-norm clip -> Adam, weights replicated
// Peer pointers come from a symmetric allocation (torch.distributed._symmetric_memory is used for the allocation /
// handle exchange only); loads from peers are ld.relaxed.sys (never cached in L1), the barrier is one release store
// per peer + acquire polls on the local flag line.  Flags carry a monotonically increasing sequence number kept on
// the device, so the launches replay from a CUDA graph.  Safety of slot reuse: between two uses of the same slot a
// rank passes two other barriers, which a slower peer only signals after it has finished reading (stream order).
// A poll that sees no progress for ~4 s sets the status word and falls through instead of hanging the GPU.
#include "p2p.cuh"

namespace ppx {
namespace p2p {

__global__ void __launch_bounds__(32) moments_merge_kernel(Peers P, int nstreams, double* __restrict__ out) {
  barrier_all(P);
  const int s = threadIdx.x;                                  // stream 0: extrinsic advantages, 1: intrinsic
  if (s >= nstreams) return;
  double tot = 0.0, wsum = 0.0;
  double n[MAXW], mean[MAXW], m2[MAXW];
  for (int r = 0; r < P.W; ++r) {
    const double* rec = (const double*)P.data[r] + 3 * s;
    n[r] = ld_peer_f64(rec); mean[r] = ld_peer_f64(rec + 1); m2[r] = ld_peer_f64(rec + 2);
    tot += n[r];
    wsum += n[r] * mean[r];
  }
  const double mu = wsum / tot;
  double M2 = 0.0;
  for (int r = 0; r < P.W; ++r) { const double d = mean[r] - mu; M2 += m2[r] + n[r] * d * d; }
  out[2 * s] = mu;
  out[2 * s + 1] = sqrt(M2 / (tot - 1.0));
}

// sums[l] = sum over ranks (rank order) of the peers' 32 partial sums; written to sums_out for the finalize step
__global__ void __launch_bounds__(32) sums_allreduce_kernel(Peers P, double* __restrict__ sums_out) {
  barrier_all(P);
  const int l = threadIdx.x;
  double s = 0.0;
  for (int r = 0; r < P.W; ++r) s += ld_peer_f64((const double*)P.data[r] + l);
  sums_out[l] = s;
}

struct AdamP {
  float* p; float* m; float* v; int n; int n_clip;
  float max_norm, w1, beta2, w2, eps;
  double beta1, beta2d, lr;
  int64_t* step_dev; double* norm_out;
  float* g_out;              // optional: the summed gradient (local, NOT the symmetric buffer)
};

// one CTA: the banks this path serves are a few thousand parameters (policy MLPs); larger banks take NCCL + optim.cu
__global__ void __launch_bounds__(1024) clip_adam_kernel(Peers P, AdamP a) {
  extern __shared__ __align__(16) float s_g[];
  __shared__ double s_red[32];
  __shared__ float s_coef, s_step, s_bc2;
  if (threadIdx.x < 32) barrier_all(P);
  __syncthreads();
  double ss = 0.0;
  const int n4 = a.n / 4;
  for (int q = threadIdx.x; q < n4; q += blockDim.x) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < P.W; ++r) {
      const float4 x = ld_peer_f32x4((const float4*)P.data[r] + q);
      g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
    }
    reinterpret_cast<float4*>(s_g)[q] = g;
    const int i = q * 4;
    if (i < a.n_clip) ss += (double)g.x * g.x;
    if (i + 1 < a.n_clip) ss += (double)g.y * g.y;
    if (i + 2 < a.n_clip) ss += (double)g.z * g.z;
    if (i + 3 < a.n_clip) ss += (double)g.w * g.w;
  }
  for (int i = n4 * 4 + threadIdx.x; i < a.n; i += blockDim.x) {
    float g = 0.f;
    for (int r = 0; r < P.W; ++r) g += ld_peer_f32((const float*)P.data[r] + i);
    s_g[i] = g;
    if (i < a.n_clip) ss += (double)g * g;
  }
  ss = block_sum(ss, s_red);
  if (threadIdx.x == 0) {
    const int64_t t_ = *a.step_dev + 1;
    *a.step_dev = t_;
    const double t = (double)t_;
    s_step = (float)(a.lr / (1.0 - pow(a.beta1, t)));
    s_bc2 = (float)sqrt(1.0 - pow(a.beta2d, t));
    float coef = 1.f;
    if (a.max_norm > 0.f && a.n_clip > 0) {
      const float norm = (float)sqrt(ss);
      coef = fminf(a.max_norm / (norm + 1e-6f), 1.f);
      if (a.norm_out) *a.norm_out = sqrt(ss);
    }
    s_coef = coef;
  }
  __syncthreads();
  const float coef = s_coef, step_size = s_step, bc2_sqrt = s_bc2;
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    float gi = s_g[i];
    if (a.g_out) a.g_out[i] = gi;
    if (i < a.n_clip) gi *= coef;
    float mi = a.m[i], vi = a.v[i];
    mi = mi + a.w1 * (gi - mi);
    vi = vi * a.beta2 + a.w2 * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + a.eps;
    a.p[i] = a.p[i] - step_size * (mi / denom);
    a.m[i] = mi;
    a.v[i] = vi;
  }
}

int fill(Peers* P, const void* const* data, void* const* flags, int W, int rank, uint32_t* seq, uint32_t* status) {
  PPX_REQUIRE(data && flags && seq && status && W >= 2 && W <= MAXW && rank >= 0 && rank < W, "p2p: W=%d rank=%d", W, rank);
  for (int r = 0; r < W; ++r) {
    PPX_REQUIRE(data[r] && flags[r], "p2p: null peer pointer for rank %d", r);
    P->data[r] = data[r];
    P->flags[r] = (uint32_t*)flags[r];
  }
  P->W = W; P->rank = rank; P->seq = seq; P->status = status;
  return PPX_OK;
}

}  // namespace p2p
}  // namespace ppx

using namespace ppx;

extern "C" int64_t ppx_p2p_max_params(void) { return (200 * 1024) / 4; }

extern "C" int ppx_p2p_moments_merge(const void* const* peer_recs, void* const* peer_flags, int W, int rank, uint32_t* seq_dev,
                                     uint32_t* status_dev, int nstreams, double* stats_out, void* stream) {
  p2p::Peers P;
  int rc = p2p::fill(&P, peer_recs, peer_flags, W, rank, seq_dev, status_dev);
  if (rc) return rc;
  PPX_REQUIRE(stats_out && nstreams >= 1 && nstreams <= 2, "p2p_moments_merge: bad arguments");
  p2p::moments_merge_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(P, nstreams, stats_out);
  return after_launch("p2p_moments_merge");
}

extern "C" int ppx_p2p_sums_allreduce(const void* const* peer_sums, void* const* peer_flags, int W, int rank, uint32_t* seq_dev,
                                      uint32_t* status_dev, double* sums_out, void* stream) {
  p2p::Peers P;
  int rc = p2p::fill(&P, peer_sums, peer_flags, W, rank, seq_dev, status_dev);
  if (rc) return rc;
  PPX_REQUIRE(sums_out, "p2p_sums_allreduce: null output");
  p2p::sums_allreduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(P, sums_out);
  return after_launch("p2p_sums_allreduce");
}

extern "C" int ppx_p2p_clip_adam(float* params, const void* const* peer_grads, void* const* peer_flags, int W, int rank,
                                 uint32_t* seq_dev, uint32_t* status_dev, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 double max_norm, int64_t n_clip, double lr, double beta1, double beta2, double eps,
                                 int64_t* step_dev, double* norm_out, float* grad_sum_out, void* stream) {
  p2p::Peers P;
  int rc = p2p::fill(&P, peer_grads, peer_flags, W, rank, seq_dev, status_dev);
  if (rc) return rc;
  PPX_REQUIRE(params && exp_avg && exp_avg_sq && step_dev && n >= 1 && n <= ppx_p2p_max_params() && n_clip >= 0 && n_clip <= n,
              "p2p_clip_adam: n=%lld (max %lld)", (long long)n, (long long)ppx_p2p_max_params());
  for (int r = 0; r < W; ++r) PPX_REQUIRE(((uintptr_t)peer_grads[r] & 15) == 0, "p2p_clip_adam: gradient vectors must be 16-byte aligned");
  p2p::AdamP a{params, exp_avg, exp_avg_sq, (int)n, (int)n_clip, (float)max_norm, (float)(1.0 - beta1), (float)beta2,
               (float)(1.0 - beta2), (float)eps, beta1, beta2, lr, step_dev, norm_out, grad_sum_out};
  const size_t smem = (size_t)((n + 3) / 4 * 4) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    PPX_CUDA(cudaFuncSetAttribute(p2p::clip_adam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  p2p::clip_adam_kernel<<<1, 1024, smem, (cudaStream_t)stream>>>(P, a);
  return after_launch("p2p_clip_adam");
}
