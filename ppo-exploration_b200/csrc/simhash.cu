// SimHash codes + persistent open-addressing count table with the reference's sequential semantics.
//
// Replaces RolloutStorage.sim_hash (buffer.py:188-200) and the defaultdict count_table (buffer.py:136).
//
// Semantics that must survive parallelisation: the reference walks the batch in env order,
// `count[key] += 1; reward[i] += beta/sqrt(count[key])`, so the j-th occurrence of a key inside one
// batch sees count c+j.  A bare atomicAdd gives the right final table but hands the bonuses to the
// wrong envs.  Here the batch is cut into chunks of CH consecutive elements, one CTA per chunk:
//   1. (parallel)  codes: bit b = (A[b,:] . obs[i,:] > 0), f64 dot            (buffer.py:194)
//   2. (parallel)  bitonic sort of (code, local index) in shared memory -> for every element its
//                  rank among equal codes of the chunk in index order, and one representative
//                  (the last of each run) that knows the run length m
//   3. (parallel)  representatives find / claim their slot: linear probing, atomicCAS on the 64-bit key
//   4. (ordered)   chunks pass a baton in chunk order (tickets are taken at CTA start, so a CTA only
//                  ever waits for CTAs that are already running): representative does
//                  base = atomicAdd(count[slot], m); the critical section is one L2 atomic round trip
//   5. (parallel)  count_i = base + rank_i + 1; reward_i += beta / sqrt(count_i)  (f64, buffer.py:199)
// Steps 1-3 of later chunks overlap the baton of earlier ones.  All 2^64 codes are valid keys: the
// all-ones code (the EMPTY sentinel) lives in a dedicated extra slot.
// Algorithmic traffic: 4D + 8 + 8 + 16 B/obs.
#include "common.cuh"

struct ppx_count_table {
  uint64_t* keys;        // [capacity + 1]
  uint32_t* counts;      // [capacity + 1]
  uint64_t capacity;     // power of two
  uint32_t* ctrl;        // [0] ticket, [1] chunks done, [2] keys in use, [3] overflow flag
  uint64_t used_bound;   // host-side upper bound of keys in use (avoids a sync per call)
};

namespace ppx {
namespace {

constexpr uint64_t kEmpty = ~0ull;
constexpr int CH = 2048;          // elements per chunk
constexpr int CT = 1024;          // threads per CTA (2 elements each)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {     // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

__device__ __forceinline__ uint64_t find_or_claim(uint64_t* keys, uint64_t cap, uint32_t* ctrl, uint64_t code) {
  if (code == kEmpty) return cap;                            // dedicated slot
  const uint64_t mask = cap - 1;
  uint64_t h = mix64(code) & mask;
  for (uint64_t probe = 0; probe < cap; ++probe) {
    unsigned long long prev = keys[h];
    if (prev == code) return h;
    if (prev == kEmpty) {
      prev = atomicCAS((unsigned long long*)&keys[h], (unsigned long long)kEmpty, (unsigned long long)code);
      if (prev == kEmpty) { atomicAdd(&ctrl[2], 1u); return h; }
      if (prev == code) return h;
    }
    h = (h + 1) & mask;
  }
  atomicExch(&ctrl[3], 1u);                                  // table full: flagged, reported by the host
  return cap;
}

__device__ __forceinline__ uint64_t code_of(const double* __restrict__ A, const float* __restrict__ x, int k, int D) {
  uint64_t code = 0;
  for (int b = 0; b < k; ++b) {
    const double* a = A + (size_t)b * D;
    double acc = 0.0;
    for (int d = 0; d < D; ++d) acc = fma(__ldg(a + d), (double)__ldg(x + d), acc);
    code |= (uint64_t)(acc > 0.0) << b;
  }
  return code;
}

__global__ void codes_kernel(const double* __restrict__ A, const float* __restrict__ obs, int k, int D, int64_t n,
                             uint64_t* __restrict__ codes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) codes[i] = code_of(A, obs + i * D, k, D);
}

__device__ __forceinline__ bool key_gt(uint64_t ca, uint16_t ia, uint64_t cb, uint16_t ib) {
  return ca > cb || (ca == cb && ia > ib);
}

template <bool FROM_OBS>
__global__ void __launch_bounds__(CT)
update_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ counts, uint64_t cap, uint32_t* ctrl,
              const double* __restrict__ A, const float* __restrict__ obs, int k, int D,
              const uint64_t* __restrict__ codes_in, int64_t n, double beta, void* rewards, int rewards_f64,
              uint64_t* __restrict__ codes_out, uint32_t* __restrict__ counts_out) {
  __shared__ uint64_t s_code[CH];
  __shared__ uint32_t s_base[CH];
  __shared__ uint16_t s_idx[CH];
  __shared__ int s_warp[32];
  __shared__ uint32_t s_chunk;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_chunk = atomicAdd(&ctrl[0], 1u);            // ticket = chunk id, in CTA start order
  __syncthreads();
  const uint32_t chunk = s_chunk;
  const int64_t e0 = (int64_t)chunk * CH;
  const int nv = (int)min((int64_t)CH, n - e0);               // valid elements in this chunk

  // 1. codes
  for (int j = tid; j < CH; j += CT) {
    uint64_t c = kEmpty;
    uint16_t id = 0xFFFF;
    if (j < nv) {
      c = FROM_OBS ? code_of(A, obs + (e0 + j) * D, k, D) : codes_in[e0 + j];
      id = (uint16_t)j;
      if (codes_out) codes_out[e0 + j] = c;
    }
    s_code[j] = c;
    s_idx[j] = id;
  }
  __syncthreads();

  // 2. bitonic sort by (code, idx); padding (kEmpty, 0xFFFF) sorts last
  for (int kk = 2; kk <= CH; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      const int i = ((tid & ~(j - 1)) << 1) | (tid & (j - 1));
      const int l = i | j;
      const bool up = (i & kk) == 0;
      const uint64_t ci = s_code[i], cl = s_code[l];
      const uint16_t ii = s_idx[i], il = s_idx[l];
      if (key_gt(ci, ii, cl, il) == up) {
        s_code[i] = cl; s_code[l] = ci;
        s_idx[i] = il; s_idx[l] = ii;
      }
      __syncthreads();
    }
  }

  // run starts: inclusive max-scan of (is_head ? p : 0); thread owns positions 2*tid, 2*tid+1
  const int p0 = 2 * tid, p1 = p0 + 1;
  const uint64_t c0 = s_code[p0], c1 = s_code[p1];
  const bool h0 = p0 < nv && (p0 == 0 || s_code[p0 - 1] != c0);
  const bool h1 = p1 < nv && (c1 != c0);
  int v = h1 ? p1 : (h0 ? p0 : 0);
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl = max(incl, o);
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int x = s_warp[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x = max(x, o);
    }
    s_warp[lane] = x;
  }
  __syncthreads();
  int before = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) before = 0;
  if (wid > 0) before = max(before, s_warp[wid - 1]);
  const int start0 = h0 ? p0 : before;
  const int start1 = h1 ? p1 : start0;
  const bool t0 = p0 < nv && (p0 == nv - 1 || c1 != c0);
  const bool t1 = p1 < nv && (p1 == nv - 1 || s_code[min(p1 + 1, CH - 1)] != c1);

  // 3. representatives (run tails) locate their slot
  uint64_t slot0 = 0, slot1 = 0;
  if (t0) slot0 = find_or_claim(keys, cap, ctrl, c0);
  if (t1) slot1 = find_or_claim(keys, cap, ctrl, c1);

  // 4. ordered section
  if (tid == 0) {
    unsigned done;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(done) : "l"(ctrl + 1) : "memory");
      if (done < chunk) __nanosleep(32);
    } while (done < chunk);
  }
  __syncthreads();
  if (t0) s_base[start0] = atomicAdd(&counts[slot0], (uint32_t)(p0 - start0 + 1));
  if (t1) s_base[start1] = atomicAdd(&counts[slot1], (uint32_t)(p1 - start1 + 1));
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ctrl + 1), "r"(chunk + 1) : "memory");
  }

  // 5. counts and bonus, scattered back to element order
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int p = q ? p1 : p0;
    if (p >= nv) continue;
    const int st = q ? start1 : start0;
    const uint32_t cnt = s_base[st] + (uint32_t)(p - st) + 1u;
    const int64_t e = e0 + s_idx[p];
    if (counts_out) counts_out[e] = cnt;
    if (rewards) {
      const double bonus = beta / sqrt((double)cnt);
      if (rewards_f64) ((double*)rewards)[e] += bonus;
      else ((float*)rewards)[e] = (float)((double)((float*)rewards)[e] + bonus);
    }
  }
}

__global__ void bonus_kernel(const uint32_t* __restrict__ counts, int64_t n, double beta, void* rewards, int f64) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double bonus = beta / sqrt((double)counts[i]);
  if (f64) ((double*)rewards)[i] += bonus;
  else ((float*)rewards)[i] = (float)((double)((float*)rewards)[i] + bonus);
}

__global__ void dump_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t cap,
                            uint64_t* keys_out, uint32_t* counts_out, uint64_t max_out, unsigned long long* n_out) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > cap) return;
  const bool used = (s < cap) ? (keys[s] != kEmpty) : (counts[cap] > 0);
  if (!used) return;
  const unsigned long long pos = atomicAdd(n_out, 1ull);
  if (pos < max_out) {
    if (keys_out) keys_out[pos] = (s < cap) ? keys[s] : kEmpty;
    if (counts_out) counts_out[pos] = counts[s];
  }
}

__global__ void rehash_kernel(const uint64_t* __restrict__ okeys, const uint32_t* __restrict__ ocounts, uint64_t ocap,
                              uint64_t* nkeys, uint32_t* ncounts, uint64_t ncap, uint32_t* ctrl) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > ocap) return;
  if (s == ocap) { ncounts[ncap] = ocounts[ocap]; return; }
  const uint64_t key = okeys[s];
  if (key == kEmpty) return;
  const uint64_t slot = find_or_claim(nkeys, ncap, ctrl, key);
  ncounts[slot] = ocounts[s];
}

int alloc_arrays(uint64_t cap, uint64_t** keys, uint32_t** counts) {
  PPX_CUDA(cudaMalloc((void**)keys, (cap + 1) * sizeof(uint64_t)));
  PPX_CUDA(cudaMalloc((void**)counts, (cap + 1) * sizeof(uint32_t)));
  return PPX_OK;
}

int clear_arrays(ppx_count_table* t, cudaStream_t st) {
  PPX_CUDA(cudaMemsetAsync(t->keys, 0xFF, (t->capacity + 1) * sizeof(uint64_t), st));
  PPX_CUDA(cudaMemsetAsync(t->counts, 0, (t->capacity + 1) * sizeof(uint32_t), st));
  PPX_CUDA(cudaMemsetAsync(t->ctrl, 0, 4 * sizeof(uint32_t), st));
  t->used_bound = 0;
  return PPX_OK;
}

// Make room for `incoming` new keys: sync + read the real fill only when the host bound says we might
// pass 50 % load; grow by rehashing into a table 2x (or more) the size.
int reserve(ppx_count_table* t, int64_t incoming, cudaStream_t st) {
  if (t->used_bound + (uint64_t)incoming <= t->capacity / 2) { t->used_bound += incoming; return PPX_OK; }
  uint32_t ctrl_h[4];
  PPX_CUDA(cudaMemcpyAsync(ctrl_h, t->ctrl, sizeof(ctrl_h), cudaMemcpyDeviceToHost, st));
  PPX_CUDA(cudaStreamSynchronize(st));
  if (ctrl_h[3]) return fail(PPX_ERR_CAPACITY, "count table overflowed (capacity %llu)", (unsigned long long)t->capacity);
  t->used_bound = ctrl_h[2];
  if (t->used_bound + (uint64_t)incoming > t->capacity / 2) {
    uint64_t ncap = t->capacity;
    while (t->used_bound + (uint64_t)incoming > ncap / 2) ncap <<= 1;
    uint64_t* nkeys; uint32_t* ncounts;
    int rc = alloc_arrays(ncap, &nkeys, &ncounts);
    if (rc) return rc;
    PPX_CUDA(cudaMemsetAsync(nkeys, 0xFF, (ncap + 1) * sizeof(uint64_t), st));
    PPX_CUDA(cudaMemsetAsync(ncounts, 0, (ncap + 1) * sizeof(uint32_t), st));
    PPX_CUDA(cudaMemsetAsync(t->ctrl + 2, 0, sizeof(uint32_t), st));
    rehash_kernel<<<(unsigned)ceil_div(t->capacity + 1, 256), 256, 0, st>>>(t->keys, t->counts, t->capacity, nkeys, ncounts,
                                                                         ncap, t->ctrl);
    rc = after_launch("count_table rehash");
    if (rc) return rc;
    PPX_CUDA(cudaStreamSynchronize(st));
    cudaFree(t->keys); cudaFree(t->counts);
    t->keys = nkeys; t->counts = ncounts; t->capacity = ncap;
  }
  t->used_bound += incoming;
  return PPX_OK;
}

int run_update(ppx_count_table* t, const double* A, const float* obs, int k, int D, const uint64_t* codes_in, int64_t n,
               double beta, void* rewards, int rewards_f64, uint64_t* codes_out, uint32_t* counts_out, cudaStream_t st) {
  if (n == 0) return PPX_OK;
  int rc = reserve(t, n, st);
  if (rc) return rc;
  PPX_CUDA(cudaMemsetAsync(t->ctrl, 0, 2 * sizeof(uint32_t), st));       // ticket + baton
  const unsigned grid = (unsigned)ceil_div(n, CH);
  if (obs)
    update_kernel<true><<<grid, CT, 0, st>>>(t->keys, t->counts, t->capacity, t->ctrl, A, obs, k, D, nullptr, n, beta,
                                              rewards, rewards_f64, codes_out, counts_out);
  else
    update_kernel<false><<<grid, CT, 0, st>>>(t->keys, t->counts, t->capacity, t->ctrl, nullptr, nullptr, 0, 0, codes_in, n,
                                               beta, rewards, rewards_f64, codes_out, counts_out);
  return after_launch("simhash update");
}

}  // namespace
}  // namespace ppx

extern "C" int ppx_simhash_codes(const double* A, const float* obs, int k, int D, int64_t n, uint64_t* codes, void* stream) {
  PPX_REQUIRE(A && obs && codes, "simhash_codes: null pointer");
  PPX_REQUIRE(k >= 1 && k <= 64 && D >= 1 && n >= 0, "simhash_codes: k=%d (1..64) D=%d n=%lld", k, D, (long long)n);
  if (n == 0) return PPX_OK;
  ppx::codes_kernel<<<(unsigned)ppx::ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(A, obs, k, D, n, codes);
  return ppx::after_launch("simhash_codes");
}

extern "C" int ppx_count_table_create(uint64_t capacity, ppx_count_table** out) {
  PPX_REQUIRE(out, "count_table_create: null out");
  uint64_t cap = 1024;
  while (cap < capacity) cap <<= 1;
  ppx_count_table* t = new ppx_count_table();
  t->capacity = cap;
  int rc = ppx::alloc_arrays(cap, &t->keys, &t->counts);
  if (rc) { delete t; return rc; }
  if (cudaMalloc((void**)&t->ctrl, 4 * sizeof(uint32_t)) != cudaSuccess) { delete t; return ppx::fail(PPX_ERR_CUDA, "cudaMalloc ctrl"); }
  rc = ppx::clear_arrays(t, 0);
  if (rc) { delete t; return rc; }
  PPX_CUDA(cudaStreamSynchronize(0));
  *out = t;
  return PPX_OK;
}

extern "C" int ppx_count_table_destroy(ppx_count_table* t) {
  if (!t) return PPX_OK;
  cudaFree(t->keys); cudaFree(t->counts); cudaFree(t->ctrl);
  delete t;
  return PPX_OK;
}

extern "C" int ppx_count_table_clear(ppx_count_table* t, void* stream) {
  PPX_REQUIRE(t, "count_table_clear: null table");
  return ppx::clear_arrays(t, (cudaStream_t)stream);
}

extern "C" int ppx_count_table_update(ppx_count_table* t, const uint64_t* codes, int64_t n, uint32_t* counts_out, void* stream) {
  PPX_REQUIRE(t && codes && counts_out && n >= 0, "count_table_update: bad arguments");
  return ppx::run_update(t, nullptr, nullptr, 0, 0, codes, n, 0.0, nullptr, 0, nullptr, counts_out, (cudaStream_t)stream);
}

extern "C" int ppx_simhash_update(ppx_count_table* t, const double* A, const float* obs, int k, int D, int64_t n, double beta,
                                  void* rewards_inout, int rewards_are_f64, uint64_t* codes_out, uint32_t* counts_out,
                                  void* stream) {
  PPX_REQUIRE(t && A && obs && n >= 0, "simhash_update: bad arguments");
  PPX_REQUIRE(k >= 1 && k <= 64 && D >= 1, "simhash_update: k=%d (1..64) D=%d", k, D);
  return ppx::run_update(t, A, obs, k, D, nullptr, n, beta, rewards_inout, rewards_are_f64, codes_out, counts_out,
                         (cudaStream_t)stream);
}

extern "C" int ppx_simhash_bonus(const uint32_t* counts, int64_t n, double beta, void* rewards_inout, int rewards_are_f64,
                                 void* stream) {
  PPX_REQUIRE(counts && rewards_inout && n >= 0, "simhash_bonus: bad arguments");
  if (n == 0) return PPX_OK;
  ppx::bonus_kernel<<<(unsigned)ppx::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(counts, n, beta, rewards_inout,
                                                                                       rewards_are_f64);
  return ppx::after_launch("simhash_bonus");
}

extern "C" int ppx_count_table_size(ppx_count_table* t, uint64_t* n_keys_host) {
  PPX_REQUIRE(t && n_keys_host, "count_table_size: bad arguments");
  uint32_t ctrl_h[4];
  uint32_t special = 0;
  PPX_CUDA(cudaDeviceSynchronize());
  PPX_CUDA(cudaMemcpy(ctrl_h, t->ctrl, sizeof(ctrl_h), cudaMemcpyDeviceToHost));
  PPX_CUDA(cudaMemcpy(&special, t->counts + t->capacity, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (ctrl_h[3]) return ppx::fail(PPX_ERR_CAPACITY, "count table overflowed");
  *n_keys_host = (uint64_t)ctrl_h[2] + (special > 0 ? 1 : 0);
  return PPX_OK;
}

extern "C" int ppx_count_table_dump(ppx_count_table* t, uint64_t* keys_dev, uint32_t* counts_dev, uint64_t max_out,
                                    uint64_t* n_out_host) {
  PPX_REQUIRE(t && n_out_host, "count_table_dump: bad arguments");
  unsigned long long* n_dev;
  PPX_CUDA(cudaDeviceSynchronize());
  PPX_CUDA(cudaMalloc((void**)&n_dev, sizeof(unsigned long long)));
  PPX_CUDA(cudaMemset(n_dev, 0, sizeof(unsigned long long)));
  ppx::dump_kernel<<<(unsigned)ppx::ceil_div(t->capacity + 1, 256), 256>>>(t->keys, t->counts, t->capacity, keys_dev, counts_dev,
                                                                         max_out, n_dev);
  int rc = ppx::after_launch("count_table_dump");
  unsigned long long n_h = 0;
  if (!rc && cudaMemcpy(&n_h, n_dev, sizeof(n_h), cudaMemcpyDeviceToHost) != cudaSuccess) rc = ppx::fail(PPX_ERR_CUDA, "dump memcpy");
  cudaFree(n_dev);
  *n_out_host = n_h;
  return rc;
}
