// Fused PPO clipped-surrogate / clipped-value / entropy loss: forward + backward in three launches.
//
// Replaces the ~25 elementwise torch ops + autograd of PPO.train (algorithms.py:219-238),
// PPO_RND.train (:431-460) and the policy part of PPO_ICM.train (:670-692), plus the
// torch.distributions math of Policy.evaluate (models.py:58-71, 106-122).
//
//   head kernel     per sample: log-prob, ratio, clipped surrogate and its gradient w.r.t. the actor
//                   output; squared value errors; per-CTA f64 partial sums
//   finalise kernel one CTA: fixed-order reduction of the partials, the five loss scalars, the
//                   max-of-means branch weights (algorithms.py:232: torch.max of two scalar means ->
//                   gradient goes to the larger, split 0.5/0.5 on a tie), d_log_std
//   dvalue kernel   per sample: gradient of the SELECTED value-loss branch
// HBM traffic: 20A + 20 B per sample-visit (+20 dual), SURVEY §8d.
//
// Precision: as the reference.  Box actions are stored f64 (buffer.py:154), so Normal.log_prob, the
// ratio and the surrogate are f64; everything else f32 with f64 reductions.
#include "common.cuh"
#include "p2p.cuh"

namespace ppx {
namespace {

constexpr int kMaxBlocks = 1024;
constexpr int kPart = 32;          // doubles per CTA partial record
constexpr int kMaxBoxA = 16;
constexpr double kHalfLog2Pi = 0.91893853320467274178;   // log(sqrt(2*pi))
constexpr float kProbEps = 1.1920928955078125e-07f;      // torch.finfo(float32).eps (clamp_probs)

struct FinalArgs {
  const double* sums;
  const float* log_std; float* d_log_std;
  double* losses; double* branch;       // branch[0..3] = w1, w2, iw1, iw2
  int64_t B; int A; int discrete; int dual;
  float ent_coef, vf_coef, int_vf_coef, pw;
  int64_t* row_dev; int row_hold;       // optional step cursor (ppx_ppo_cfg)
};

// one thread: the five loss scalars, the max-of-means branch weights, d_log_std, from the 32 global sums s[]
__device__ void finalize_body(const FinalArgs& p, const double* s) {
  const int64_t row = p.row_dev ? *p.row_dev : 0;
  double* const L = p.losses + 8 * row;
  if (p.row_dev && !p.row_hold) *p.row_dev = row + 1;
  const double Bd = (double)p.B;
  const double n_terms = p.discrete ? Bd : Bd * p.A;
  const double pl = -s[0] / n_terms;
  const double m1 = s[1] / Bd, m2 = s[2] / Bd;
  // compare as the reference does, on the float32 means
  const float m1f = (float)m1, m2f = (float)m2;
  double w1 = m1f > m2f ? 1.0 : (m1f < m2f ? 0.0 : 0.5);
  const double vl = fmax(m1, m2);
  double ivl = 0.0, iw1 = 0.0;
  if (p.dual) {
    const double i1 = s[3] / Bd, i2 = s[4] / Bd;
    const float i1f = (float)i1, i2f = (float)i2;
    iw1 = i1f > i2f ? 1.0 : (i1f < i2f ? 0.0 : 0.5);
    ivl = fmax(i1, i2);
  }
  double el;
  if (p.discrete) {
    el = -s[5] / Bd;
  } else {
    double e = 0.0;
    for (int a = 0; a < p.A; ++a) {
      const float ent = 0.5f + 0.5f * 1.8378770664093453f + logf(expf(p.log_std[a]));   // Normal.entropy
      e += (double)ent;
      // d total / d log_std[a]: surrogate part + entropy part (d(-mean ent)/ds_a = -1/A)
      p.d_log_std[a] = (float)(s[8 + a] - (double)p.pw * (double)p.ent_coef / (double)p.A);
    }
    el = -e / (double)p.A;
  }
  const double total = (double)p.pw * (pl + (double)p.ent_coef * el + (double)p.vf_coef * vl) +
                       (p.dual ? (double)p.int_vf_coef * ivl : 0.0);
  L[0] = total; L[1] = pl; L[2] = vl; L[3] = el; L[4] = ivl;
  L[5] = 0.0; L[6] = 0.0; L[7] = 0.0;
  p.branch[0] = w1; p.branch[1] = 1.0 - w1; p.branch[2] = iw1; p.branch[3] = 1.0 - iw1;
}

__global__ void __launch_bounds__(32) finalize_kernel(FinalArgs p) {
  const int lane = threadIdx.x;
  __shared__ double s[kPart];
  s[lane] = p.sums[lane];
  __syncwarp();
  if (lane == 0) finalize_body(p, s);
}

__device__ unsigned int g_head_ticket = 0;

struct HeadArgs {
  const float* actor_out; const float* log_std; const double* actions; const float* old_lp;
  const float* adv; const double* adv_stats; const float* values; const float* old_values; const float* returns;
  const float* int_adv; const double* int_adv_stats; const float* int_values; const float* old_int_values;
  const float* int_returns;
  float* d_actor_out;
  double* partials;
  int64_t B; int64_t Bt; int A; int dual;       // B = rows on this rank, Bt = rows of the global minibatch
  float clip, ent_coef, pw;
  double* sums_out;                     // [32] sum of the CTA partials (written by the last CTA to finish)
  int do_final; FinalArgs fin;          // do_final: the last CTA also runs finalize_body
  // sharded do_final (W >= 2): the last CTA exchanges the 32 sums with its peers first (ppx_ppo_cfg.peer_sums_host)
  int W, rank;
  uint64_t* xs[p2p::MAXW];              // rank r's staging [2 parities][W source ranks][64 words of {u32 half, u32 seq}]
  uint32_t* seq_dev; uint32_t* status_dev;
};

__device__ __forceinline__ void st_ll_u32(uint64_t* p, uint32_t v, uint32_t tag) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(tag) : "memory");
}
__device__ __forceinline__ uint32_t ld_ll_u32_wait(const uint64_t* p, uint32_t tag, uint32_t* status) {
  uint32_t v, t;
  const uint64_t t0 = p2p::now_ns();
  for (;;) {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(t) : "l"(p) : "memory");
    if (t == tag) break;
    if (p2p::now_ns() - t0 > 4000000000ull) { atomicExch(status, 1u); break; }
  }
  return v;
}

__device__ __forceinline__ float norm_adv(float a, const double* st) {
  return (a - (float)st[0]) / ((float)st[1] + 1e-8f);       // algorithms.py:219 in f32
}

template <bool DISCRETE>
__global__ void __launch_bounds__(256) head_kernel(HeadArgs p) {
  __shared__ double s_red[32];
  __shared__ double s_inv2var[kMaxBoxA], s_invvar[kMaxBoxA], s_logscale[kMaxBoxA];
  if (!DISCRETE) {
    if (threadIdx.x < p.A) {
      const float sigma = expf(p.log_std[threadIdx.x]);                      // models.py:69
      const float var = sigma * sigma;
      s_inv2var[threadIdx.x] = 1.0 / (double)(2.f * var);
      s_invvar[threadIdx.x] = 1.0 / (double)var;
      s_logscale[threadIdx.x] = (double)logf(sigma);
    }
    __syncthreads();
  }
  double pl = 0.0, s1 = 0.0, s2 = 0.0, is1 = 0.0, is2 = 0.0, ent_sum = 0.0;
  double dls[kMaxBoxA];
#pragma unroll
  for (int a = 0; a < kMaxBoxA; ++a) dls[a] = 0.0;
  const int A = p.A;
  const float lo = 1.f - p.clip, hi = 1.f + p.clip;

  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < p.B; b += (int64_t)gridDim.x * blockDim.x) {
    float adv = norm_adv(p.adv[b], p.adv_stats);
    if (p.dual) adv += norm_adv(p.int_adv[b], p.int_adv_stats);              // algorithms.py:431-434
    if (!DISCRETE) {
      const double gscale = -(double)p.pw / ((double)p.Bt * (double)A);       // d(-mean)/d(term)
#pragma unroll
      for (int a = 0; a < kMaxBoxA; ++a) {
        if (a >= A) break;
        const float mu = tanhf(p.actor_out[b * A + a]);                      // models.py:163
        // Normal.log_prob in f64 (the actions are f64, buffer.py:154); the per-dimension constants sigma, var,
        // log(sigma), 1/(2 var), 1/var are hoisted (s_c*), divisions become multiplications by those reciprocals
        const double diff = p.actions[b * A + a] - (double)mu;
        const double d2 = diff * diff;
        const double lp = -d2 * s_inv2var[a] - s_logscale[a] - kHalfLog2Pi;
        // ratio = exp(lp - old_lp): the difference is formed in f64, exponentiated in f32 (relative error ~1e-7,
        // two orders inside the 1e-5 bound; the f64 exp was the kernel's bottleneck: 18 % -> HBM-side)
        const double ratio = (double)expf((float)(lp - (double)p.old_lp[b * A + a]));
        const double cr = fmin(fmax(ratio, (double)lo), (double)hi);
        const double x = (double)adv * ratio, y = (double)adv * cr;
        pl += fmin(x, y);
        const bool inside = ratio >= (double)lo && ratio <= (double)hi;      // clamp passes gradient (inclusive)
        double g = 0.0;                                                      // d min / d ratio
        if (inside) g = (double)adv;
        else if (x < y) g = (double)adv;
        else if (x == y) g = 0.5 * (double)adv;
        const double dlp = gscale * g * ratio;
        const double dmu = dlp * diff * s_invvar[a];
        p.d_actor_out[b * A + a] = (float)(dmu * (1.0 - (double)mu * (double)mu));
        dls[a] += dlp * (d2 * s_invvar[a] - 1.0);
      }
    } else {
      const float* l = p.actor_out + b * A;
      float mx = l[0];
      for (int j = 1; j < A; ++j) mx = fmaxf(mx, l[j]);
      float se = 0.f;
      for (int j = 0; j < A; ++j) se += expf(l[j] - mx);
      // Categorical(probs=softmax): probs renormalised, logits = log(clamp(probs, eps, 1-eps))   models.py:62
      float sq = 0.f;
      for (int j = 0; j < A; ++j) sq += expf(l[j] - mx) / se;
      const int act = (int)p.actions[b];
      float ent = 0.f, q_act = 0.f, lq_act = 0.f;
      for (int j = 0; j < A; ++j) {
        const float q = (expf(l[j] - mx) / se) / sq;
        const float lq = logf(fminf(fmaxf(q, kProbEps), 1.f - kProbEps));
        ent -= q * lq;
        if (j == act) { q_act = q; lq_act = lq; }
      }
      ent_sum += (double)ent;
      const float ratio = expf(lq_act - p.old_lp[b]);
      const float cr = fminf(fmaxf(ratio, lo), hi);
      const float x = adv * ratio, y = adv * cr;
      pl += (double)fminf(x, y);
      const bool inside = ratio >= lo && ratio <= hi;
      float g = 0.f;
      if (inside) g = adv;
      else if (x < y) g = adv;
      else if (x == y) g = 0.5f * adv;
      const float dlp = -p.pw / (float)p.Bt * g * ratio;
      const float dent = -p.pw * p.ent_coef / (float)p.Bt;                   // d(ent_coef * -mean(ent)) / d ent_b
      // gradient w.r.t. q_j, then through renormalise+softmax: dl_k = (g_k - sum_j g_j q_j) q_k
      const bool act_in = q_act >= kProbEps && q_act <= 1.f - kProbEps;
      float dot = 0.f;
      for (int j = 0; j < A; ++j) {
        const float q = (expf(l[j] - mx) / se) / sq;
        const bool in = q >= kProbEps && q <= 1.f - kProbEps;
        const float lq = logf(fminf(fmaxf(q, kProbEps), 1.f - kProbEps));
        float gq = -dent * (lq + (in ? 1.f : 0.f));
        if (j == act && act_in) gq += dlp / q;
        dot += gq * q;
      }
      for (int j = 0; j < A; ++j) {
        const float q = (expf(l[j] - mx) / se) / sq;
        const bool in = q >= kProbEps && q <= 1.f - kProbEps;
        const float lq = logf(fminf(fmaxf(q, kProbEps), 1.f - kProbEps));
        float gq = -dent * (lq + (in ? 1.f : 0.f));
        if (j == act && act_in) gq += dlp / q;
        p.d_actor_out[b * A + j] = (gq - dot) * q;
      }
    }
    {
      const float v = p.values[b], ov = p.old_values[b], R = p.returns[b];
      const float vc = ov + fminf(fmaxf(v - ov, -p.clip), p.clip);           // algorithms.py:229
      const float e1 = R - v, e2 = R - vc;
      s1 += (double)(e1 * e1);
      s2 += (double)(e2 * e2);
    }
    if (p.dual) {
      const float v = p.int_values[b], ov = p.old_int_values[b], R = p.int_returns[b];
      const float vc = ov + fminf(fmaxf(v - ov, -p.clip), p.clip);           // algorithms.py:451
      const float e1 = R - v, e2 = R - vc;
      is1 += (double)(e1 * e1);
      is2 += (double)(e2 * e2);
    }
  }
  // per-CTA partial record: ONE combined reduction (warp shuffles of every quantity, one shared-memory exchange, warp 0
  // finishes) instead of 6 + A sequential block_sum calls with two CTA barriers each
  double* out = p.partials + (int64_t)blockIdx.x * kPart;
  {
    __shared__ double s_w[8][8 + kMaxBoxA];
    double q[8 + kMaxBoxA];
    q[0] = pl; q[1] = s1; q[2] = s2; q[3] = is1; q[4] = is2; q[5] = ent_sum; q[6] = 0.0; q[7] = 0.0;
#pragma unroll
    for (int a = 0; a < kMaxBoxA; ++a) q[8 + a] = DISCRETE ? 0.0 : dls[a];
    const int nq = DISCRETE ? 6 : 8 + A;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 8 + kMaxBoxA; ++k) {
      if (k >= nq) break;
      const double v = warp_sum(q[k]);
      if (lane == 0) s_w[wid][k] = v;
    }
    __syncthreads();
    if (wid == 0 && lane < nq) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_w[w][lane];
      out[lane] = t;
    }
  }
  // ---- tail: the last CTA sums the partials in a fixed order (8 warps take interleaved CTAs, then 0..7) ----
  if (!last_block_done(&g_head_ticket)) return;
  __shared__ double s_p[8][kPart];
  __shared__ double s_sum[kPart];
  const int lane = threadIdx.x & 31, ph = threadIdx.x >> 5, nblocks = gridDim.x;
  double acc = 0.0;
  {
    int k = ph;
    for (; k + 24 < nblocks; k += 32) {                      // 4 independent loads in flight, summed in a fixed order
      const double v0 = __ldcg(p.partials + (int64_t)k * kPart + lane), v1 = __ldcg(p.partials + (int64_t)(k + 8) * kPart + lane);
      const double v2 = __ldcg(p.partials + (int64_t)(k + 16) * kPart + lane), v3 = __ldcg(p.partials + (int64_t)(k + 24) * kPart + lane);
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; k < nblocks; k += 8) acc += __ldcg(p.partials + (int64_t)k * kPart + lane);
  }
  s_p[ph][lane] = acc;
  __syncthreads();
  if (ph == 0) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += s_p[q][lane];
    if (p.W >= 2) {
      // all-reduce of the 32 sums over peer memory, value + sequence number in one 8-byte store (two per double): push
      // to every rank, poll the own slots, add in rank order -- identical sums, hence the identical max-of-means branch,
      // on every rank; no exchange kernel, no fence (same protocol as the optimiser tail, mlp_fused.cu)
      const uint32_t seq = *p.seq_dev + 1u;
      const size_t slot = (size_t)(seq & 1u) * p.W * 64;
      const uint64_t bits = (uint64_t)__double_as_longlong(t);
      for (int r = 0; r < p.W; ++r) {
        uint64_t* dst = p.xs[r] + slot + (size_t)p.rank * 64 + 2 * lane;
        st_ll_u32(dst, (uint32_t)bits, seq);
        st_ll_u32(dst + 1, (uint32_t)(bits >> 32), seq);
      }
      double g = 0.0;
      for (int r = 0; r < p.W; ++r) {
        const uint64_t* src = p.xs[p.rank] + slot + (size_t)r * 64 + 2 * lane;
        const uint32_t lo = ld_ll_u32_wait(src, seq, p.status_dev), hi = ld_ll_u32_wait(src + 1, seq, p.status_dev);
        g += __longlong_as_double((long long)(((uint64_t)hi << 32) | lo));
      }
      t = g;
      __syncwarp();
      if (lane == 0) *p.seq_dev = seq;
    }
    s_sum[lane] = t;
    p.sums_out[lane] = t;
  }
  __syncthreads();
  if (p.do_final && threadIdx.x == 0) finalize_body(p.fin, s_sum);
}

__global__ void __launch_bounds__(256)
dvalue_kernel(const float* __restrict__ values, const float* __restrict__ old_values, const float* __restrict__ returns,
              const double* __restrict__ branch, int64_t B, int64_t Bt, float clip, float scale, float* __restrict__ d_values) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float w1 = (float)branch[0], w2 = (float)branch[1];
  const float v = values[b], ov = old_values[b], R = returns[b];
  const float d = v - ov;
  const float vc = ov + fminf(fmaxf(d, -clip), clip);
  const float pass = (d >= -clip && d <= clip) ? 1.f : 0.f;                  // clamp backward mask (inclusive)
  const float g = w1 * (-2.f * (R - v)) + w2 * (-2.f * (R - vc)) * pass;
  d_values[b] = scale * g / (float)Bt;
}

// ---- stand-alone MSE and cross-entropy pieces (ICM / RND predictor losses) ----
__global__ void __launch_bounds__(256)
mse_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, double scale, float* d_a, float* d_b,
           double* __restrict__ partials) {
  __shared__ double s_red[32];
  double s = 0.0;
  const float gs = (float)(2.0 * scale / (double)n);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    s += (double)(d * d);
    if (d_a) d_a[i] = gs * d;
    if (d_b) d_b[i] = -gs * d;
  }
  s = block_sum(s, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256)
xent_kernel(const float* __restrict__ logits, const double* __restrict__ targets, int tstride, int64_t B, int C,
            double scale, float* __restrict__ d_logits, double* __restrict__ partials) {
  __shared__ double s_red[32];
  double s = 0.0;
  const float gs = (float)(scale / (double)B);
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const float* l = logits + b * C;
    float mx = l[0];
    for (int j = 1; j < C; ++j) mx = fmaxf(mx, l[j]);
    float se = 0.f;
    for (int j = 0; j < C; ++j) se += expf(l[j] - mx);
    const float lse = logf(se);
    const int t = (int)targets[b * tstride];
    s += (double)-(l[t] - mx - lse);
    for (int j = 0; j < C; ++j) d_logits[b * C + j] = gs * (expf(l[j] - mx - lse) - (j == t ? 1.f : 0.f));
  }
  s = block_sum(s, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void accum_loss_kernel(const double* __restrict__ partials, int n, double scale, double* loss_accum) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < n; ++k) s += partials[k];
    *loss_accum += s * scale;
  }
}

__global__ void row_commit_kernel(double* losses, int64_t* row_dev, const double* value, int col, int add_to_total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int64_t row = *row_dev;
  const double v = *value;
  losses[8 * row + col] = v;
  if (add_to_total) losses[8 * row] += v;
  *row_dev = row + 1;
}

int grid_for(int64_t n) {
  int64_t g = ceil_div(n, 256);
  const int64_t cap = std::min<int64_t>(kMaxBlocks, (int64_t)sm_count() * 2);   // few partials: the last-CTA tail reads them all
  return (int)std::max<int64_t>(1, std::min(g, cap));
}

double* scratch_partials(cudaStream_t) {
  // small persistent scratch for the stand-alone loss kernels (one per process; calls are stream-ordered
  // by the single learner thread, see SURVEY §8b "Threading")
  static double* buf = nullptr;
  if (!buf) cudaMalloc((void**)&buf, kMaxBlocks * sizeof(double));
  return buf;
}

}  // namespace
}  // namespace ppx

using namespace ppx;

extern "C" int64_t ppx_ppo_loss_workspace(int64_t, int) { return (int64_t)(kMaxBlocks * kPart + kPart + 8) * sizeof(double); }

namespace {
int check_cfg(const ppx_ppo_cfg* c) {
  PPX_REQUIRE(c, "ppo_loss: null cfg");
  PPX_REQUIRE(c->B >= 1 && c->A >= 1 && (c->B_total == 0 || c->B_total >= c->B), "ppo_loss: B=%lld B_total=%lld A=%d",
              (long long)c->B, (long long)c->B_total, c->A);
  if (!c->discrete) PPX_REQUIRE(c->A <= kMaxBoxA, "ppo_loss: Box head supports A <= %d", kMaxBoxA);
  return PPX_OK;
}
inline int64_t total_rows(const ppx_ppo_cfg* c) { return c->B_total > 0 ? c->B_total : c->B; }
}  // namespace

namespace {
int launch_head(const ppx_ppo_cfg* c, const float* actor_out, const float* log_std, const double* actions,
                const float* old_log_probs, const float* advantages, const double* adv_stats, const float* values,
                const float* old_values, const float* returns, const float* int_advantages, const double* int_adv_stats,
                const float* int_values, const float* old_int_values, const float* int_returns, float* d_actor_out,
                double* sums_out, void* workspace, float* d_log_std, double* losses_out, double* branch_out, void* stream) {
  int rc = check_cfg(c);
  if (rc) return rc;
  PPX_REQUIRE(actor_out && actions && old_log_probs && advantages && adv_stats && values && old_values && returns,
              "ppo_loss: null input");
  PPX_REQUIRE(d_actor_out && sums_out && workspace, "ppo_loss: null output");
  if (!c->discrete) PPX_REQUIRE(log_std, "ppo_loss: Box head needs log_std");
  if (c->dual) PPX_REQUIRE(int_advantages && int_adv_stats && int_values && old_int_values && int_returns,
                           "ppo_loss: dual head needs the intrinsic arrays");
  cudaStream_t st = (cudaStream_t)stream;
  double* partials = (double*)workspace;
  const int g = grid_for(c->B);
  HeadArgs h{actor_out, log_std, actions, old_log_probs, advantages, adv_stats, values, old_values, returns,
             int_advantages, int_adv_stats, int_values, old_int_values, int_returns, d_actor_out, partials,
             c->B, total_rows(c), c->A, c->dual, c->clip_range, c->ent_coef, c->policy_weight, sums_out, 0, FinalArgs{}};
  h.W = 0;
  if (losses_out && c->W >= 2) {
    PPX_REQUIRE(c->W <= p2p::MAXW && c->rank >= 0 && c->rank < c->W && c->peer_sums_host && c->seq_dev && c->status_dev,
                "ppo_loss: bad peer arguments (W=%d rank=%d)", c->W, c->rank);
    h.W = c->W; h.rank = c->rank; h.seq_dev = c->seq_dev; h.status_dev = c->status_dev;
    for (int q = 0; q < c->W; ++q) {
      PPX_REQUIRE(c->peer_sums_host[q], "ppo_loss: null peer pointer for rank %d", q);
      h.xs[q] = (uint64_t*)c->peer_sums_host[q];
    }
  }
  if (losses_out) {
    PPX_REQUIRE(branch_out && (c->discrete || d_log_std), "ppo_loss: fused finalize needs branch_out / d_log_std");
    h.do_final = 1;
    h.fin = FinalArgs{nullptr, log_std, d_log_std, losses_out, branch_out, total_rows(c), c->A, c->discrete, c->dual,
                      c->ent_coef, c->vf_coef, c->int_vf_coef, c->policy_weight, c->row_dev, c->row_hold};
  }
  if (c->discrete) head_kernel<true><<<g, 256, 0, st>>>(h);
  else head_kernel<false><<<g, 256, 0, st>>>(h);
  return after_launch("ppo_loss head");
}
}  // namespace

extern "C" int ppx_ppo_loss_head(const ppx_ppo_cfg* c, const float* actor_out, const float* log_std, const double* actions,
                                 const float* old_log_probs, const float* advantages, const double* adv_stats,
                                 const float* values, const float* old_values, const float* returns,
                                 const float* int_advantages, const double* int_adv_stats, const float* int_values,
                                 const float* old_int_values, const float* int_returns, float* d_actor_out,
                                 double* sums_out, void* workspace, void* stream) {
  return launch_head(c, actor_out, log_std, actions, old_log_probs, advantages, adv_stats, values, old_values, returns,
                     int_advantages, int_adv_stats, int_values, old_int_values, int_returns, d_actor_out, sums_out,
                     workspace, nullptr, nullptr, nullptr, stream);
}

extern "C" int ppx_ppo_loss_head_final(const ppx_ppo_cfg* c, const float* actor_out, const float* log_std,
                                       const double* actions, const float* old_log_probs, const float* advantages,
                                       const double* adv_stats, const float* values, const float* old_values,
                                       const float* returns, const float* int_advantages, const double* int_adv_stats,
                                       const float* int_values, const float* old_int_values, const float* int_returns,
                                       float* d_actor_out, float* d_log_std, double* losses_out, double* branch_out,
                                       void* workspace, void* stream) {
  PPX_REQUIRE(workspace && losses_out && branch_out, "ppo_loss_head_final: null pointer");
  double* sums = (double*)workspace + (int64_t)kMaxBlocks * kPart;
  return launch_head(c, actor_out, log_std, actions, old_log_probs, advantages, adv_stats, values, old_values, returns,
                     int_advantages, int_adv_stats, int_values, old_int_values, int_returns, d_actor_out, sums,
                     workspace, d_log_std, losses_out, branch_out, stream);
}

extern "C" int ppx_ppo_loss_finalize(const ppx_ppo_cfg* c, const double* sums, const float* log_std, float* d_log_std,
                                     double* losses_out, double* branch_out, void* stream) {
  int rc = check_cfg(c);
  if (rc) return rc;
  PPX_REQUIRE(sums && losses_out && branch_out, "ppo_loss_finalize: null pointer");
  if (!c->discrete) PPX_REQUIRE(log_std && d_log_std, "ppo_loss_finalize: Box head needs log_std / d_log_std");
  FinalArgs f{sums, log_std, d_log_std, losses_out, branch_out, total_rows(c), c->A, c->discrete, c->dual,
              c->ent_coef, c->vf_coef, c->int_vf_coef, c->policy_weight, c->row_dev, c->row_hold};
  finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f);
  return after_launch("ppo_loss finalize");
}

extern "C" int ppx_ppo_loss_finish(const ppx_ppo_cfg* c, const double* sums, const float* log_std, const float* values,
                                   const float* old_values, const float* returns, const float* int_values,
                                   const float* old_int_values, const float* int_returns, float* d_log_std,
                                   float* d_values, float* d_int_values, double* losses_out, void* workspace,
                                   void* stream) {
  int rc = check_cfg(c);
  if (rc) return rc;
  PPX_REQUIRE(sums && values && old_values && returns && d_values && losses_out && workspace, "ppo_loss_finish: null pointer");
  if (!c->discrete) PPX_REQUIRE(log_std && d_log_std, "ppo_loss_finish: Box head needs log_std / d_log_std");
  if (c->dual) PPX_REQUIRE(int_values && old_int_values && int_returns && d_int_values, "ppo_loss_finish: dual head arrays");
  cudaStream_t st = (cudaStream_t)stream;
  double* branch = (double*)workspace + (int64_t)kMaxBlocks * kPart + kPart;
  const int64_t Bt = total_rows(c);
  FinalArgs f{sums, log_std, d_log_std, losses_out, branch, Bt, c->A, c->discrete, c->dual,
              c->ent_coef, c->vf_coef, c->int_vf_coef, c->policy_weight, c->row_dev, c->row_hold};
  finalize_kernel<<<1, 32, 0, st>>>(f);
  rc = after_launch("ppo_loss finalize");
  if (rc) return rc;
  const unsigned gb = (unsigned)ceil_div(c->B, 256);
  dvalue_kernel<<<gb, 256, 0, st>>>(values, old_values, returns, branch, c->B, Bt, c->clip_range, c->policy_weight * c->vf_coef, d_values);
  rc = after_launch("ppo_loss dvalue");
  if (rc || !c->dual) return rc;
  dvalue_kernel<<<gb, 256, 0, st>>>(int_values, old_int_values, int_returns, branch + 2, c->B, Bt, c->clip_range, c->int_vf_coef, d_int_values);
  return after_launch("ppo_loss dvalue(int)");
}

extern "C" int ppx_ppo_loss_fwd_bwd(const ppx_ppo_cfg* c, const float* actor_out, const float* log_std, const double* actions,
                                    const float* old_log_probs, const float* advantages, const double* adv_stats,
                                    const float* values, const float* old_values, const float* returns,
                                    const float* int_advantages, const double* int_adv_stats, const float* int_values,
                                    const float* old_int_values, const float* int_returns, float* d_actor_out,
                                    float* d_log_std, float* d_values, float* d_int_values, double* losses_out,
                                    void* workspace, void* stream) {
  PPX_REQUIRE(workspace, "ppo_loss: null workspace");
  double* sums = (double*)workspace + (int64_t)kMaxBlocks * kPart;
  int rc = ppx_ppo_loss_head(c, actor_out, log_std, actions, old_log_probs, advantages, adv_stats, values, old_values,
                             returns, int_advantages, int_adv_stats, int_values, old_int_values, int_returns,
                             d_actor_out, sums, workspace, stream);
  if (rc) return rc;
  return ppx_ppo_loss_finish(c, sums, log_std, values, old_values, returns, int_values, old_int_values, int_returns,
                             d_log_std, d_values, d_int_values, losses_out, workspace, stream);
}

extern "C" int ppx_mse_fwd_bwd(const float* a, const float* b, int64_t n, double scale, float* d_a, float* d_b,
                               double* loss_accum, void* stream) {
  PPX_REQUIRE(a && b && n >= 1 && loss_accum, "mse: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  double* part = scratch_partials(st);
  const int g = grid_for(n);
  mse_kernel<<<g, 256, 0, st>>>(a, b, n, scale, d_a, d_b, part);
  int rc = after_launch("mse");
  if (rc) return rc;
  accum_loss_kernel<<<1, 32, 0, st>>>(part, g, scale / (double)n, loss_accum);
  return after_launch("mse accum");
}

extern "C" int ppx_xent_fwd_bwd(const float* logits, const double* targets, int target_stride, int64_t B, int C, double scale,
                                float* d_logits, double* loss_accum, void* stream) {
  PPX_REQUIRE(logits && targets && d_logits && loss_accum && B >= 1 && C >= 1, "xent: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  double* part = scratch_partials(st);
  const int g = grid_for(B);
  xent_kernel<<<g, 256, 0, st>>>(logits, targets, target_stride, B, C, scale, d_logits, part);
  int rc = after_launch("xent");
  if (rc) return rc;
  accum_loss_kernel<<<1, 32, 0, st>>>(part, g, scale / (double)B, loss_accum);
  return after_launch("xent accum");
}

extern "C" int ppx_loss_row_commit(double* losses, int64_t* row_dev, const double* value, int col, int add_to_total, void* stream) {
  PPX_REQUIRE(losses && row_dev && value && col >= 1 && col < 8, "loss_row_commit: bad arguments");
  row_commit_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(losses, row_dev, value, col, add_to_total);
  return after_launch("loss_row_commit");
}
