// Peer-memory primitives shared by the fused compute+exchange kernels (p2p.cu: sharded PPO update; es.cu: sharded ES
// update): release/acquire flag barrier over NVLink-mapped symmetric memory, uncached peer loads.
#pragma once
#include "common.cuh"

namespace ppx {
namespace p2p {

constexpr int MAXW = 16;

struct Peers {
  const void* data[MAXW];      // per-rank payload of this exchange (record / sums / gradient vector)
  uint32_t* flags[MAXW];       // per-rank flag array of this channel: flags[p][r] = last sequence number rank r signalled to p
  int W, rank;
  uint32_t* seq;               // local, device: sequence number of this channel
  uint32_t* status;            // local, device: != 0 after a barrier timed out
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f32x4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_peer_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t now_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// All ranks call this once per exchange (first warp of the CTA; the caller syncs the CTA afterwards).
// Everything this rank wrote before the call (earlier kernels on the stream) is visible to a peer that has seen the flag.
__device__ __forceinline__ void barrier_all(const Peers& P) {
  const int lane = threadIdx.x & 31;
  uint32_t seq = 0;
  if (lane == 0) { seq = *P.seq + 1u; *P.seq = seq; }
  seq = __shfl_sync(0xffffffffu, seq, 0);
  __threadfence_system();
  if (lane < P.W) st_release_sys(P.flags[lane] + P.rank, seq);
  if (lane < P.W) {
    const uint32_t* mine = P.flags[P.rank] + lane;
    const uint64_t t0 = now_ns();
    while ((int32_t)(ld_acquire_sys(mine) - seq) < 0) {
      if (now_ns() - t0 > 4000000000ull) { atomicExch(P.status, 1u); break; }
    }
  }
  __syncwarp();
}

// host: validate and pack the per-rank pointer tables of one exchange (defined in p2p.cu)
int fill(Peers* P, const void* const* data, void* const* flags, int W, int rank, uint32_t* seq, uint32_t* status);

}  // namespace p2p
}  // namespace ppx
