// Device side of the NVLink peer-memory exchange (lrx_p2p.cu, lrx_update.cu): buffer layout, system-scope
// flag accesses, publish + wait.
#pragma once

#include "lrx_common.cuh"

namespace p2p {

// Buffer = [global flag (uint32 at offset 0), padded to a line | per-block flags of slot 0 | of slot 1 | slot 0 | slot 1].
// The global flag serves lrx_p2p_allreduce / lrx_ppo_apply_p2p (whole-slot publish); the per-block flags serve the fused
// reduce + exchange + apply kernel of lrx_ppo_grad_apply_p2p, where block b of a rank only depends on block b of its peers.
// Inside the global flag's line: +64 the ticket of the reduction kernel (which block finishes last), +72 the relay flag
// ("every peer has arrived", written by block 0, polled locally by the other blocks); both zeroed by lrx_p2p_alloc.
constexpr int MAX_BLOCK_FLAGS = 1024;
constexpr size_t GLOBAL_FLAG_BYTES = 128;
// PUSH exchange (the default of lrx_ppo_update_p2p for up to PUSH_MAX_RANKS ranks): a rank WRITES its reduced gradient
// into every peer's receive area recv[slot][source rank] and then the source's arrival flag, so the receiver polls and
// sums LOCAL memory only: one one-way NVLink trip per minibatch instead of the pull protocol's flag round trips plus a
// remote-load round trip.
constexpr int PUSH_MAX_RANKS = 8;
constexpr size_t PUSH_FLAG_BYTES = 256;  // arrival flag of source rank r: uint32 at GLOBAL_FLAG_BYTES + 2 * MAX_BLOCK_FLAGS * 4 + 4 r
constexpr size_t HEADER = GLOBAL_FLAG_BYTES + 2 * (size_t)MAX_BLOCK_FLAGS * 4 + PUSH_FLAG_BYTES;  // 8576 = 67 lines
__host__ __device__ inline uint32_t* push_flag(char* buffer, int source_rank) {
  return reinterpret_cast<uint32_t*>(buffer + GLOBAL_FLAG_BYTES + 2 * (size_t)MAX_BLOCK_FLAGS * 4) + source_rank;
}
// receive area of `source_rank` for slot `seq & 1`: behind the two pull slots
__host__ __device__ inline float* push_recv(char* buffer, size_t slot_b, uint32_t seq, int source_rank) {
  return reinterpret_cast<float*>(buffer + HEADER + 2 * slot_b + ((size_t)(seq & 1u) * PUSH_MAX_RANKS + source_rank) * slot_b);
}
__host__ __device__ inline uint32_t* block_flag(char* buffer, uint32_t seq, int block) {
  return reinterpret_cast<uint32_t*>(buffer + GLOBAL_FLAG_BYTES) + (size_t)(seq & 1u) * MAX_BLOCK_FLAGS + block;
}

__host__ __device__ inline size_t slot_bytes(int n_floats) { return ((size_t)n_floats * 4 + 127) / 128 * 128; }

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_volatile_f32(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// Block-wide: (one thread of ONE block, `publisher`) publish seq in the own flag, then wait until every peer's flag
// has reached seq.  Ends with a block barrier; returns false (for every thread of the block) when a peer did not arrive
// within ~10 s: *error_flag is set and the caller must NOT consume the slots (they may be stale or half written).
__device__ __forceinline__ bool publish_and_wait(char* const* peers, int rank, int world, uint32_t seq, bool publisher,
                                                 int* error_flag) {
  int timed_out = 0;
  if (publisher && threadIdx.x == 0) {
    __threadfence_system();  // the gradient kernel's stores to the own slot (earlier launch) before the flag
    st_release_sys(reinterpret_cast<uint32_t*>(peers[rank]), seq);
  }
  if ((int)threadIdx.x < world) {
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(peers[threadIdx.x]);
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flag) - seq) < 0) {
      if (clock64() - t0 > 20000000000LL) {  // a peer is gone -- report instead of hanging the GPU
        *error_flag = 1;
        timed_out = 1;
        break;
      }
      __nanosleep(100);
    }
  }
  return __syncthreads_or(timed_out) == 0;
}

// Sum of element i of slot (seq & 1) over the ranks, in rank order (identical on every rank).  The loads of up to eight
// ranks are ISSUED together and only then added: a loop of load-and-add serialises the NVLink round trips (~2 us each:
// it was most of the 18 us a minibatch's exchange cost at 8 GPUs in round 1).
__device__ __forceinline__ float peer_sum(char* const* peers, int world, size_t slot_b, uint32_t seq, int i) {
  const size_t off = HEADER + (size_t)(seq & 1u) * slot_b;
  float s = 0.f;
  for (int r0 = 0; r0 < world; r0 += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = (r0 + k < world) ? ld_volatile_f32(reinterpret_cast<const float*>(peers[r0 + k] + off) + i) : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (r0 + k < world) s += v[k];
  }
  return s;
}

}  // namespace p2p
