// Random permutation of [0, n) for the minibatch shuffle of buffer/base_buffer.py:85-88
// (jr.permutation(key, N*T) in the reference; the RNG stream itself is not part of the parity
// contract).  A keyed bijection instead of a sort: 8 half-rounds of a (possibly unbalanced) Feistel
// network on the smallest power-of-two domain >= n with cycle walking (fewer than two walks per
// element on average, none when n is a power of two -- a balanced network on an even number of
// bits made whole warps wait for their slowest lane), so out[i] is computed independently per
// element in ONE pass over 4n bytes (torch.randperm's radix sort costs ~0.7 ms per epoch at
// 8.4 M elements, this ~10 us).
#include "lrx_common.cuh"

__device__ __forceinline__ uint32_t perm_mix(uint32_t x, uint32_t k) {
  x ^= k;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

__global__ void __launch_bounds__(256)
permutation_kernel(int32_t* __restrict__ out, long long n, uint32_t lbits, uint32_t rbits, uint32_t k0, uint32_t k1) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t lmask = (1u << lbits) - 1u, rmask = (1u << rbits) - 1u;
  uint32_t x = (uint32_t)i;
  do {  // cycle walking: the network permutes [0, 2^(lbits+rbits)) with 2^(lbits+rbits) < 2n; re-apply until inside [0, n)
    uint32_t l = x >> rbits, r = x & rmask;
#pragma unroll
    for (uint32_t round = 0; round < 4; ++round) {  // 8 half-rounds, alternating which side is rewritten
      l ^= perm_mix(r, k0 + (2 * round) * 0x9E3779B9u) & lmask;
      r ^= perm_mix(l, k1 + (2 * round + 1) * 0x9E3779B9u) & rmask;
    }
    x = (l << rbits) | r;
  } while ((long long)x >= n);
  out[i] = (int32_t)x;
}

extern "C" int lrx_permutation(int32_t* out, int64_t n, uint64_t seed, lrx_stream_t stream) {
  LRX_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), LRX_ERR_INVALID_ARG, "lrx_permutation: n out of range");
  if (n == 0) return LRX_OK;
  LRX_REQUIRE(out, LRX_ERR_INVALID_ARG, "lrx_permutation: NULL pointer");
  uint32_t bits = 2;
  while (((int64_t)1 << bits) < n) ++bits;  // smallest domain >= n: no walk at all when n is a power of two
  const uint32_t k0 = (uint32_t)seed ^ 0xA511E9B3u, k1 = (uint32_t)(seed >> 32) ^ 0x63D83595u;
  permutation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n, bits - bits / 2, bits / 2, k0, k1);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

// ---------------------------------------------------------------------------------------------------- packed rows
// The flatten of buffer/base_buffer.py:38-62 made physical: one 64-byte record per row in reference order i = n*T + t, so
// that the minibatch gather of the update (base_buffer.py:90-103) touches ONE line per row.  Threads walk the planes in
// their native [T][N] order (coalesced reads); every thread writes its own whole 64-byte record (four 16-byte stores: full
// sectors even though the records of neighbouring lanes are T * 64 bytes apart).
__global__ void __launch_bounds__(256)
pack_rows_kernel(LrxBatchView b, int d, float* __restrict__ packed) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // = t * N + n
  const long long total = (long long)b.N * b.T;
  if (e >= total) return;
  const int t = (int)(e / b.N), n = (int)(e - (long long)t * b.N);
  float o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = k < d ? b.obs[(size_t)e * d + k] : 0.f;
  const float act = __int_as_float(reinterpret_cast<const int32_t*>(b.act)[e]);
  float4* dst = reinterpret_cast<float4*>(packed + ((size_t)n * b.T + t) * LRX_PACKED_FLOATS);
  dst[0] = make_float4(o[0], o[1], o[2], o[3]);
  dst[1] = make_float4(o[4], o[5], o[6], o[7]);
  dst[2] = make_float4(act, b.logp[e], b.val[e], b.ret[e]);
  dst[3] = make_float4(b.adv[e], 0.f, 0.f, 0.f);
}

extern "C" int lrx_ppo_pack_rows(const LrxBatchView* buf, int32_t obs_dim, int32_t is_box, float* packed,
                                 lrx_stream_t stream) {
  LRX_REQUIRE(buf && packed, LRX_ERR_INVALID_ARG, "lrx_ppo_pack_rows: NULL pointer");
  LRX_REQUIRE(buf->obs && buf->act && buf->logp && buf->val && buf->ret && buf->adv && buf->N > 0 && buf->T > 0,
              LRX_ERR_INVALID_ARG, "lrx_ppo_pack_rows: LrxBatchView has NULL planes or empty shape");
  LRX_REQUIRE(obs_dim >= 1 && obs_dim <= 8, LRX_ERR_UNSUPPORTED, "lrx_ppo_pack_rows: obs_dim %d does not fit the record", obs_dim);
  LRX_REQUIRE(((uintptr_t)packed & 63) == 0, LRX_ERR_INVALID_ARG, "lrx_ppo_pack_rows: packed must be 64-byte aligned");
  (void)is_box;  // the action travels as its 32 bits either way
  const long long total = (long long)buf->N * buf->T;
  pack_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*buf, obs_dim, packed);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}
