// Streaming statistics kernels for the receivers beside the hot path:
//   moe_colsum_f32            column sums of the expert scores over (a subset of) the tokens -- GetExperts' mean score
//   moe_rownorm_colsumsq_bf16 column sums of squares of the row-normalised hidden state -- the Wanda receiver's norms
// HBM-bound: every input element is read once with 16-byte loads; outputs are accumulated with float atomics
// (positive terms; the summation order across CTAs is the only non-determinism, ~1e-7 relative).
#include "common.cuh"

namespace moe {

constexpr int kStatThreads = 256;

// out[c] += sum over rows t (with row_mask[t % period] != 0 if a mask is given) of m[t, c]
__global__ void __launch_bounds__(kStatThreads) colsum_kernel(const float* __restrict__ m, int rows, int cols,
                                                              const uint8_t* __restrict__ row_mask, int period,
                                                              float* __restrict__ out, int rows_per_cta) {
  pdl_wait();
  pdl_launch_dependents();
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  for (int c = blockIdx.x * kStatThreads + threadIdx.x; c < cols; c += gridDim.x * kStatThreads) {
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) {
      if (row_mask != nullptr && !row_mask[r % period]) continue;
      acc += __ldg(m + static_cast<size_t>(r) * cols + c);
    }
    if (r1 > r0) atomicAdd(out + c, acc);
  }
}

// out[c] += sum over rows t of (H[t, c] / max(||H[t, :]||, 1e-12))^2      (F.normalize(p=2, dim=1), then column norms^2)
constexpr int kNormRows = 16;   // rows per CTA: one warp computes two row norms, then all threads sweep the columns
__global__ void __launch_bounds__(kStatThreads) rownorm_colsumsq_kernel(const __nv_bfloat16* __restrict__ H, int rows,
                                                                        int cols, float* __restrict__ out) {
  __shared__ float s_inv2[kNormRows];
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * kNormRows;
  for (int rr = warp; rr < kNormRows; rr += kStatThreads / 32) {
    const int r = r0 + rr;
    float ss = 0.f;
    if (r < rows) {
      const uint4* row = reinterpret_cast<const uint4*>(H + static_cast<size_t>(r) * cols);   // cols % 8 == 0
      for (int i = lane; i < cols / 8; i += 32) {
        const uint4 q = __ldg(row + i);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xffff0000u);
          ss = fmaf(lo, lo, fmaf(hi, hi, ss));
        }
      }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) {
      const float nrm = fmaxf(sqrtf(ss), 1e-12f);
      s_inv2[rr] = r < rows ? 1.f / (nrm * nrm) : 0.f;
    }
  }
  __syncthreads();
  const int nrows = min(kNormRows, rows - r0);
  for (int c2 = threadIdx.x; c2 < cols / 2; c2 += kStatThreads) {   // two columns (one 32-bit word) per thread
    float a0 = 0.f, a1 = 0.f;
    for (int rr = 0; rr < nrows; ++rr) {
      const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(H + static_cast<size_t>(r0 + rr) * cols) + c2);
      const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
      a0 = fmaf(lo * lo, s_inv2[rr], a0);
      a1 = fmaf(hi * hi, s_inv2[rr], a1);
    }
    atomicAdd(out + 2 * c2, a0);
    atomicAdd(out + 2 * c2 + 1, a1);
  }
}

}  // namespace moe

extern "C" {

int moe_colsum_f32(const float* m, int T, int C, const uint8_t* row_mask, int period, float* out, void* stream) {
  using namespace moe;
  MOE_REQUIRE(m != nullptr && out != nullptr && T >= 0 && C >= 1, MOE_ERR_INVALID_ARGUMENT, "moe_colsum_f32: bad args");
  MOE_REQUIRE(row_mask == nullptr || period >= 1, MOE_ERR_INVALID_ARGUMENT, "moe_colsum_f32: row_mask needs period >= 1");
  if (T == 0) return MOE_OK;
  const int rows_per_cta = 64;
  dim3 grid((C + kStatThreads - 1) / kStatThreads, (T + rows_per_cta - 1) / rows_per_cta);
  cudaError_t le = launch_pdl(colsum_kernel, grid, dim3(kStatThreads), 0, static_cast<cudaStream_t>(stream), m, T, C, row_mask,
                              period, out, rows_per_cta);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_colsum_f32 launch: %s", cudaGetErrorString(le));
  return check_launch("moe_colsum_f32");
}

int moe_rownorm_colsumsq_bf16(const void* H, int T, int h, float* out, void* stream) {
  using namespace moe;
  MOE_REQUIRE(H != nullptr && out != nullptr && T >= 0 && h >= 8, MOE_ERR_INVALID_ARGUMENT, "moe_rownorm_colsumsq_bf16: bad args");
  MOE_REQUIRE(h % 8 == 0 && (reinterpret_cast<uintptr_t>(H) & 15) == 0, MOE_ERR_UNSUPPORTED_SHAPE,
              "moe_rownorm_colsumsq_bf16: h=%d must be a multiple of 8 and H 16-byte aligned", h);
  if (T == 0) return MOE_OK;
  cudaError_t le = launch_pdl(rownorm_colsumsq_kernel, dim3((T + kNormRows - 1) / kNormRows), dim3(kStatThreads), 0,
                              static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(H), T, h, out);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_rownorm_colsumsq_bf16 launch: %s", cudaGetErrorString(le));
  return check_launch("moe_rownorm_colsumsq_bf16");
}

}  // extern "C"
