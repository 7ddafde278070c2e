// One-shot all-reduce of the gradient buffer over NVLink peer memory (the one exchange
// step of the path -- SUM of gradbuf[P + 8] over the ranks between backward and clip + Adam).
//
// The message is 52 KB: an NCCL all-reduce of that size is pure latency (launch + protocol, ~30-45 us at 8
// GPUs, 320 times per PPO iteration).  Here every rank owns an exchange buffer that all peers map through
// CUDA IPC: [flag | slot 0 | slot 1].  A rank's gradient kernel writes its partial sum into slot (seq & 1)
// of its OWN buffer; this kernel publishes `seq` in the own flag (system-scope release), waits until every
// peer's flag has reached `seq` (acquire), then every rank sums the G slots in rank order -- the same
// order everywhere, so all ranks get bit-identical gradients -- into a private output.  Two slots are
// enough: a rank can only publish seq + 2 after every peer has published seq + 1, i.e. finished reading
// slot seq.  No collective library call, no host synchronisation.
#include "lrx_common.cuh"
#include "lrx_p2p.cuh"

namespace {

__global__ void p2p_allreduce_kernel(char* const* __restrict__ peers, int rank, int world, size_t slot_b,
                                     float* __restrict__ out, int n, uint32_t seq, int* __restrict__ error_flag) {
  if (!p2p::publish_and_wait(peers, rank, world, seq, blockIdx.x == 0, error_flag)) return;  // `out` is left untouched
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = p2p::peer_sum(peers, world, slot_b, seq, i);
}

}  // namespace

extern "C" size_t lrx_p2p_buffer_bytes(int32_t n_floats) {
  const size_t slot = p2p::slot_bytes(n_floats);
  return p2p::HEADER + 2 * slot + 2 * (size_t)p2p::PUSH_MAX_RANKS * slot;  // pull slots, then the push receive areas
}

extern "C" int lrx_p2p_alloc(int32_t n_floats, void** buffer, unsigned char* ipc_handle) {
  LRX_REQUIRE(n_floats > 0 && buffer && ipc_handle, LRX_ERR_INVALID_ARG, "lrx_p2p_alloc: bad argument");
  void* p = nullptr;
  const size_t bytes = lrx_p2p_buffer_bytes(n_floats);
  LRX_CUDA_CHECK(cudaMalloc(&p, bytes));
  LRX_CUDA_CHECK(cudaMemset(p, 0, bytes));
  LRX_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  LRX_CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(h) == LRX_P2P_HANDLE_BYTES, "IPC handle size");
  memcpy(ipc_handle, &h, sizeof(h));
  *buffer = p;
  return LRX_OK;
}

extern "C" int lrx_p2p_open(const unsigned char* ipc_handle, void** buffer) {
  LRX_REQUIRE(ipc_handle && buffer, LRX_ERR_INVALID_ARG, "lrx_p2p_open: NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  LRX_CUDA_CHECK(cudaIpcOpenMemHandle(buffer, h, cudaIpcMemLazyEnablePeerAccess));
  return LRX_OK;
}

extern "C" int lrx_p2p_close(void* peer_buffer) {
  if (peer_buffer) LRX_CUDA_CHECK(cudaIpcCloseMemHandle(peer_buffer));
  return LRX_OK;
}

extern "C" int lrx_p2p_free(void* own_buffer) {
  if (own_buffer) LRX_CUDA_CHECK(cudaFree(own_buffer));
  return LRX_OK;
}

extern "C" float* lrx_p2p_slot(void* own_buffer, int32_t n_floats, uint32_t seq) {
  const size_t slot = p2p::slot_bytes(n_floats);
  return reinterpret_cast<float*>(static_cast<char*>(own_buffer) + p2p::HEADER + (size_t)(seq & 1u) * slot);
}

extern "C" int lrx_p2p_allreduce(void* const* peer_buffers, int32_t rank, int32_t world, int32_t n_floats, float* out,
                                 uint32_t seq, int32_t* error_flag, lrx_stream_t stream) {
  LRX_REQUIRE(peer_buffers && out && error_flag && world >= 1 && world <= 64 && rank >= 0 && rank < world && n_floats > 0 &&
                  seq != 0,
              LRX_ERR_INVALID_ARG, "lrx_p2p_allreduce: bad argument");
  const size_t slot = p2p::slot_bytes(n_floats);
  const int threads = 256, blocks = (n_floats + threads - 1) / threads < 64 ? (n_floats + threads - 1) / threads : 64;
  p2p_allreduce_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<char* const*>(peer_buffers), rank,
                                                                       world, slot, out, n_floats, seq, error_flag);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

// test helper: device-to-device copy into a raw address (a slot of the exchange buffer is not a torch tensor)
extern "C" int lrx_debug_copy_f32(float* dst, const float* src, int32_t n, lrx_stream_t stream) {
  LRX_REQUIRE(dst && src && n >= 0, LRX_ERR_INVALID_ARG, "lrx_debug_copy_f32: bad argument");
  LRX_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return LRX_OK;
}
