// Removal-mask utilities: bit-packing of dense 0/1 masks, union of packed masks, and the masked
// copy of the down-projection weight.  Pure streaming kernels (HBM-bound), 16-byte accesses.
#include "common.cuh"

namespace moe {

// one thread -> one 32-bit word from 32 mask bytes
__global__ void __launch_bounds__(256) mask_pack_kernel(const uint8_t* __restrict__ dense, long long n,
                                                        uint32_t* __restrict__ bits, long long n_words) {
  pdl_wait();
  pdl_launch_dependents();
  const long long w = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  const long long base = w * 32;
  uint32_t out = 0u;
  if (base + 32 <= n && (reinterpret_cast<uintptr_t>(dense) & 15) == 0) {
    const uint4* p = reinterpret_cast<const uint4*>(dense + base);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    const uint32_t q[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((q[i] >> (8 * j)) & 0xffu) out |= 1u << (4 * i + j);
  } else {
    for (int i = 0; i < 32 && base + i < n; ++i)
      if (dense[base + i]) out |= 1u << i;
  }
  bits[w] = out;
}

__global__ void __launch_bounds__(256) mask_union_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                         uint32_t* __restrict__ out, long long n_words, int vec_ok) {
  pdl_wait();
  pdl_launch_dependents();
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nvec = vec_ok ? (n_words >> 2) : 0;
  for (long long i = tid; i < nvec; i += nthreads) {
    const uint4 x = reinterpret_cast<const uint4*>(a)[i];
    const uint4 y = reinterpret_cast<const uint4*>(b)[i];
    reinterpret_cast<uint4*>(out)[i] = make_uint4(x.x | y.x, x.y | y.y, x.z | y.z, x.w | y.w);
  }
  for (long long i = (nvec << 2) + tid; i < n_words; i += nthreads) out[i] = a[i] | b[i];
}

// 8 bf16 (16 bytes) + one mask byte per thread
__global__ void __launch_bounds__(256) mask_weights_kernel(const uint4* __restrict__ w, const uint8_t* __restrict__ bits,
                                                           uint4* __restrict__ out, long long n_vec) {
  pdl_wait();
  pdl_launch_dependents();
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = tid; i < n_vec; i += nthreads) {
    uint4 v = __ldg(w + i);
    const uint32_t m = __ldg(bits + i);  // little-endian: byte i of the bit array covers elements 8i..8i+7
    uint32_t* q = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if ((m >> (2 * j)) & 1u) q[j] &= 0xffff0000u;
      if ((m >> (2 * j + 1)) & 1u) q[j] &= 0x0000ffffu;
    }
    out[i] = v;
  }
}

static int stream_grid(long long items, int per_cta) {
  long long g = (items + per_cta - 1) / per_cta;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (g > cap) g = cap;
  return g < 1 ? 1 : static_cast<int>(g);
}

}  // namespace moe

extern "C" {

int moe_mask_pack(const uint8_t* dense, long long n, uint32_t* bits, void* stream) {
  using namespace moe;
  MOE_REQUIRE(dense != nullptr && bits != nullptr && n >= 0, MOE_ERR_INVALID_ARGUMENT, "moe_mask_pack: bad args");
  if (n == 0) return MOE_OK;
  const long long n_words = (n + 31) / 32;
  const long long grid = (n_words + 255) / 256;
  MOE_REQUIRE(grid < (1LL << 31), MOE_ERR_UNSUPPORTED_SHAPE, "moe_mask_pack: mask too large");
  cudaError_t le = launch_pdl(mask_pack_kernel, dim3(static_cast<unsigned>(grid)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                              dense, n, bits, n_words);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_mask_pack launch: %s", cudaGetErrorString(le));
  return check_launch("moe_mask_pack");
}

int moe_mask_union(const uint32_t* a, const uint32_t* b, uint32_t* out, long long n_words, void* stream) {
  using namespace moe;
  MOE_REQUIRE(a != nullptr && b != nullptr && out != nullptr && n_words >= 0, MOE_ERR_INVALID_ARGUMENT,
              "moe_mask_union: bad args");
  if (n_words == 0) return MOE_OK;
  const int vec_ok = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                       reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  cudaError_t le = launch_pdl(mask_union_kernel, dim3(stream_grid(n_words / 4 + 1, 256 * 4)), dim3(256), 0,
                              static_cast<cudaStream_t>(stream), a, b, out, n_words, vec_ok);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_mask_union launch: %s", cudaGetErrorString(le));
  return check_launch("moe_mask_union");
}

int moe_mask_weights(const void* w2, const uint32_t* bits, void* w2m, int d, int h, void* stream) {
  using namespace moe;
  MOE_REQUIRE(w2 != nullptr && bits != nullptr && w2m != nullptr && d >= 1 && h >= 1, MOE_ERR_INVALID_ARGUMENT,
              "moe_mask_weights: bad args");
  MOE_REQUIRE(h % 32 == 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_mask_weights: h=%d must be a multiple of 32", h);
  MOE_REQUIRE(((reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(w2m)) & 15) == 0,
              MOE_ERR_INVALID_ARGUMENT, "moe_mask_weights: weights must be 16-byte aligned");
  const long long n_vec = static_cast<long long>(d) * h / 8;
  cudaError_t le = launch_pdl(mask_weights_kernel, dim3(stream_grid(n_vec, 256 * 4)), dim3(256), 0,
                              static_cast<cudaStream_t>(stream), static_cast<const uint4*>(w2),
                              reinterpret_cast<const uint8_t*>(bits), static_cast<uint4*>(w2m), n_vec);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_mask_weights launch: %s", cudaGetErrorString(le));
  return check_launch("moe_mask_weights");
}

}  // extern "C"
