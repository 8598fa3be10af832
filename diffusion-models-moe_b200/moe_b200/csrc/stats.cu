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
  MOE_REQUIRE((m != nullptr || T == 0) && out != nullptr && T >= 0 && C >= 1, MOE_ERR_INVALID_ARGUMENT, "moe_colsum_f32: bad args");
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
  MOE_REQUIRE((H != nullptr || T == 0) && out != nullptr && T >= 0 && h >= 8, MOE_ERR_INVALID_ARGUMENT, "moe_rownorm_colsumsq_bf16: bad args");
  MOE_REQUIRE(h % 8 == 0 && (reinterpret_cast<uintptr_t>(H) & 15) == 0, MOE_ERR_UNSUPPORTED_SHAPE,
              "moe_rownorm_colsumsq_bf16: h=%d must be a multiple of 8 and H 16-byte aligned", h);
  if (T == 0) return MOE_OK;
  cudaError_t le = launch_pdl(rownorm_colsumsq_kernel, dim3((T + kNormRows - 1) / kNormRows), dim3(kStatThreads), 0,
                              static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(H), T, h, out);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_rownorm_colsumsq_bf16 launch: %s", cudaGetErrorString(le));
  return check_launch("moe_rownorm_colsumsq_bf16");
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// Wanda scoring and union-over-time voting
// ------------------------------------------------------------------------------------------------------------
namespace moe {

constexpr int kWandaThreads = 256;
constexpr int kWandaMaxCols = 32;   // columns per thread: h <= 8192

// One CTA per weight row r.  metric[c] = |W2[r, c]| * norm[c] (>= 0, so the float bit pattern orders like the value).
// bit (r * h + c) of `bits` := (metric_adj[c] > metric_base[c]) && c is among the k largest metric_adj of the row
// (ties on the k-th value: lowest column first).  The k-th largest comes from a 4-pass, 8-bit radix select over a
// shared-memory histogram; a warp owns 32 consecutive columns, so a ballot is one output word.
__global__ void __launch_bounds__(kWandaThreads) wanda_score_mask_kernel(const __nv_bfloat16* __restrict__ w2,
                                                                         const float* __restrict__ norm_base,
                                                                         const float* __restrict__ norm_adj, int h, int k,
                                                                         uint32_t* __restrict__ bits) {
  __shared__ unsigned int s_hist[256];
  __shared__ unsigned int s_sel[2];      // [0] digit chosen in this pass, [1] keys above it (cumulative)
  __shared__ unsigned int s_warp[kWandaThreads / 32];
  pdl_wait();
  pdl_launch_dependents();
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (h + kWandaThreads - 1) / kWandaThreads;
  uint32_t key[kWandaMaxCols];
  uint32_t gt_base = 0u;                 // bit j: metric_adj > metric_base for my j-th column
#pragma unroll
  for (int j = 0; j < kWandaMaxCols; ++j) {
    key[j] = 0u;
    const int c = j * kWandaThreads + tid;
    if (j < per && c < h) {
      const float w = fabsf(__bfloat162float(w2[static_cast<size_t>(r) * h + c]));
      const float ma = w * __ldg(norm_adj + c), mb = w * __ldg(norm_base + c);
      key[j] = __float_as_uint(ma);
      if (ma > mb) gt_base |= 1u << j;
    }
  }
  // ---- k-th largest key of the row
  uint32_t prefix = 0u, prefix_mask = 0u;
  unsigned int need = static_cast<unsigned int>(k);   // rank still to be found inside the current prefix class
  for (int shift = 24; shift >= 0 && k > 0; shift -= 8) {
    s_hist[tid] = 0u;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kWandaMaxCols; ++j) {
      const int c = j * kWandaThreads + tid;
      if (j < per && c < h && (key[j] & prefix_mask) == prefix) atomicAdd(&s_hist[(key[j] >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int above = 0u;
      int dgt = 255;
      for (; dgt > 0; --dgt) {
        if (above + s_hist[dgt] >= need) break;
        above += s_hist[dgt];
      }
      s_sel[0] = static_cast<unsigned int>(dgt);
      s_sel[1] = above;
    }
    __syncthreads();
    prefix |= s_sel[0] << shift;
    prefix_mask |= 255u << shift;
    need -= s_sel[1];
    __syncthreads();
  }
  const uint32_t kth = prefix;           // value of the k-th largest key; `need` of the keys equal to it are selected
  // ---- emit: greater than kth, plus the first `need` ties in column order
  unsigned int tie_base = 0u;
  for (int j = 0; j < per; ++j) {
    const int c = j * kWandaThreads + tid;
    const bool in = c < h;
    uint32_t kj = 0u;
#pragma unroll
    for (int jj = 0; jj < kWandaMaxCols; ++jj)
      if (jj == j) kj = key[jj];
    const bool gt = in && k > 0 && kj > kth;
    const bool eq = in && k > 0 && kj == kth;
    const unsigned int eq_ballot = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) s_warp[warp] = __popc(eq_ballot);
    __syncthreads();
    unsigned int before = tie_base, total = 0u;
    for (int w = 0; w < kWandaThreads / 32; ++w) {
      if (w < warp) before += s_warp[w];
      total += s_warp[w];
    }
    before += __popc(eq_ballot & ((1u << lane) - 1u));
    const bool take = gt || (eq && before < need);
    const unsigned int word = __ballot_sync(0xffffffffu, take && ((gt_base >> j) & 1u));
    if (lane == 0 && (j * kWandaThreads + warp * 32) < h)
      bits[(static_cast<size_t>(r) * h + j * kWandaThreads + warp * 32) >> 5] = word;
    tie_base += total;
    __syncthreads();
  }
}

// out bit i := (number of the T masks with bit i set) > threshold      (masks: [T][n_words] bit words)
__global__ void __launch_bounds__(256) mask_vote_kernel(const uint32_t* __restrict__ masks, int T, long long n_words,
                                                        float threshold, uint32_t* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long w = warp0; w < n_words; w += nwarps) {
    int cnt = 0;
    for (int t = 0; t < T; ++t) cnt += (__ldg(masks + static_cast<size_t>(t) * n_words + w) >> lane) & 1u;
    const unsigned int word = __ballot_sync(0xffffffffu, static_cast<float>(cnt) > threshold);
    if (lane == 0) out[w] = word;
  }
}

}  // namespace moe

extern "C" {

int moe_wanda_score_mask(const void* w2, const float* norm_base, const float* norm_adj, int d, int h, int k,
                         uint32_t* bits, void* stream) {
  using namespace moe;
  MOE_REQUIRE(w2 && norm_base && norm_adj && bits, MOE_ERR_INVALID_ARGUMENT, "moe_wanda_score_mask: NULL pointer");
  MOE_REQUIRE(d >= 1 && h >= 32 && k >= 0 && k <= h, MOE_ERR_INVALID_ARGUMENT, "moe_wanda_score_mask: d=%d h=%d k=%d", d, h, k);
  MOE_REQUIRE(h % 32 == 0 && h <= kWandaThreads * kWandaMaxCols, MOE_ERR_UNSUPPORTED_SHAPE,
              "moe_wanda_score_mask: h=%d must be a multiple of 32 and <= %d", h, kWandaThreads * kWandaMaxCols);
  cudaError_t le = launch_pdl(wanda_score_mask_kernel, dim3(d), dim3(kWandaThreads), 0, static_cast<cudaStream_t>(stream),
                              static_cast<const __nv_bfloat16*>(w2), norm_base, norm_adj, h, k, bits);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_wanda_score_mask launch: %s", cudaGetErrorString(le));
  return check_launch("moe_wanda_score_mask");
}

int moe_mask_vote(const uint32_t* masks, int T, long long n_words, float threshold, uint32_t* out, void* stream) {
  using namespace moe;
  MOE_REQUIRE(masks && out && T >= 1 && n_words >= 0, MOE_ERR_INVALID_ARGUMENT, "moe_mask_vote: bad args");
  if (n_words == 0) return MOE_OK;
  long long ctas = (n_words * 32 + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (ctas > cap) ctas = cap;
  cudaError_t le = launch_pdl(mask_vote_kernel, dim3(static_cast<unsigned>(ctas)), dim3(256), 0,
                              static_cast<cudaStream_t>(stream), masks, T, n_words, threshold, out);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_mask_vote launch: %s", cudaGetErrorString(le));
  return check_launch("moe_mask_vote");
}

}  // extern "C"
