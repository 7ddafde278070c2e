// Device side of the NVLink peer-memory exchange (lrx_p2p.cu, lrx_update.cu): buffer layout, system-scope
// flag accesses, publish + wait.
#pragma once

#include "lrx_common.cuh"

namespace p2p {

constexpr size_t HEADER = 128;  // flag (uint32 at offset 0), padded to a line

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
// has reached seq.  Ends with __syncthreads().  A peer that does not arrive within ~10 s sets *error_flag.
__device__ __forceinline__ void publish_and_wait(char* const* peers, int rank, int world, uint32_t seq, bool publisher,
                                                 int* error_flag) {
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
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
}

// Sum of element i of slot (seq & 1) over the ranks, in rank order (identical on every rank).
__device__ __forceinline__ float peer_sum(char* const* peers, int world, size_t slot_b, uint32_t seq, int i) {
  const size_t off = HEADER + (size_t)(seq & 1u) * slot_b;
  float s = 0.f;
  for (int r = 0; r < world; ++r) s += ld_volatile_f32(reinterpret_cast<const float*>(peers[r] + off) + i);
  return s;
}

}  // namespace p2p
