// Grouped / gathered down-projection for sm_100a: Y[t] = b2 + sum over the token's ACTIVE experts e of
// H[t, e-th neuron segment] * W2[:, e-th segment]^T, reading only the active experts' slices of H and W2.
//
// Driven by the compacted token -> expert permutation of the router (router.cu: moe_expert_permutation):
//   perm_tokens[perm_offsets[e] + i]   i-th token (ascending) that selected expert e; segments padded to 128 rows (-1)
//   slot_pos[t * k + j]                row of token t's j-th active expert in that list (-1: fewer than k active)
// Kernel 1 (grouped_down_kernel): one CTA per (128-row tile of one expert's list, bn output columns).  The 128 gathered
//   rows of H[:, e * es ... + es) -- es = 64 bf16 = one 128-byte swizzled k-block -- are copied into shared memory by
//   cp.async (a row per thread, zero-filled for pad rows), the expert's W2 slice [bn x 64] arrives by TMA, one elected
//   thread issues four tcgen05.mma (K = 16 each) into TMEM, the epilogue writes the fp32 partial rows.
// Kernel 2 (grouped_combine_kernel): per token, the partial rows of its active experts are added in ascending expert
//   order in fp32 (deterministic -- no atomics), + b2, -> bf16.
// Experts removed for this (timestep, layer) are absent from the lists (active = selected AND NOT removed), so the
// removal mask costs nothing here: their tiles do not exist.
//
// This is the form BASELINE.json's north star names.  Measured against the dense-masked down-projection (DESIGN.md
// section 4, profiles/r02_grouped_vs_dense.log): with K = 64 per expert there is no accumulation loop to amortise the
// gather, the TMEM read-out and k x (fp32 partial write + read) per token, so the dense GEMM over the zeroed H wins
// on every SD-1.5 shape; the grouped path is kept as a tested alternative (MOE_ERR_UNSUPPORTED_SHAPE unless es == 64).
#include "common.cuh"
#include "tcgen05.cuh"

namespace moe {
namespace grouped {

constexpr int kRows = 128;             // tokens per tile = UMMA M
constexpr int kEs = 64;                // neurons per expert = one 64-wide k-block
constexpr int kABytes = kRows * kEs * 2;

struct Args {
  const __nv_bfloat16* H;     // [T, h]
  const int* offsets;         // [E + 1], multiples of kRows
  const int* tokens;          // [capacity]
  float* part;                // [capacity, d]
  int h, d, E, bn, n_tiles, max_row_tiles;
};

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, bool valid) {
  const int bytes = valid ? 16 : 0;      // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

template <uint32_t kTmemCols>
__global__ void __launch_bounds__(kRows) grouped_down_kernel(const __grid_constant__ CUtensorMap tmap_w2, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_expert;
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kABytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rt = blockIdx.x / a.n_tiles, nt = blockIdx.x - rt * a.n_tiles;

  if (threadIdx.x == 0) {
    tc::prefetch_tensormap(&tmap_w2);
    tc::mbar_init(&bar_b, 1);
    tc::mbar_init(&bar_mma, 1);
    tc::fence_mbar_init();
  }
  pdl_wait();                       // offsets / tokens / H come from the previous kernels in the stream
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    // the expert whose (padded) list holds row rt * 128: last e with offsets[e] <= row (empty lists share an offset)
    const int row = rt * kRows;
    int e = -1;
    if (row < __ldg(a.offsets + a.E)) {
      int lo = 0, hi = a.E;         // invariant: offsets[lo] <= row < offsets[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a.offsets + mid) <= row) lo = mid; else hi = mid;
      }
      e = lo;
    }
    s_expert = e;
  }
  __syncthreads();
  const int e = s_expert;
  if (e < 0) return;                // beyond the last list: nothing to do (uniform for the CTA)

  if (warp == 0) tc::tmem_alloc<kTmemCols>(&tmem_base_s);
  // ---- B: this expert's W2 slice, rows [nt * bn, + bn), columns [e * 64, + 64)
  if (threadIdx.x == 0) {
    tc::mbar_arrive_expect_tx(&bar_b, static_cast<uint32_t>(a.bn) * 128u);
    tc::tma_load_2d(sB, &tmap_w2, &bar_b, e * kEs, nt * a.bn);
  }
  // ---- A: gathered rows, one per thread, into the 128-byte-swizzled K-major layout the UMMA descriptor expects
  {
    const int r = threadIdx.x;
    const int tok = __ldg(a.tokens + rt * kRows + r);
    const bool valid = tok >= 0;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(a.H + static_cast<size_t>(valid ? tok : 0) * a.h + e * kEs);
    const uint32_t dst = tc::smem_u32(sA) + r * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) cp_async_16(dst + ((c ^ (r & 7)) << 4), src + 16 * c, valid);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    tc::fence_proxy_async_smem();   // generic-proxy writes (cp.async) -> the tensor core's async-proxy reads
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    tc::mbar_wait(&bar_b, 0);
    tc::fence_after_thread_sync();
    if (tc::elect_one()) {
      const uint32_t idesc = tc::umma_idesc_bf16_f32(kRows, static_cast<uint32_t>(a.bn));
      const uint32_t a_addr = tc::smem_u32(sA), b_addr = tc::smem_u32(sB);
#pragma unroll
      for (int k = 0; k < kEs / 16; ++k)
        tc::umma_bf16_ss(tmem_base, tc::umma_desc_kmajor_sw128(a_addr + k * 32), tc::umma_desc_kmajor_sw128(b_addr + k * 32),
                         idesc, k != 0 ? 1u : 0u);
      tc::umma_commit(&bar_mma);
    }
    __syncwarp();
  }
  tc::mbar_wait(&bar_mma, 0);
  tc::fence_after_thread_sync();

  // ---- epilogue: TMEM lane = tile row = thread; fp32 partial row -> part[rt * 128 + r, nt * bn ...]
  {
    const int r = threadIdx.x;
    float* dst = a.part + static_cast<size_t>(rt * kRows + r) * a.d + nt * a.bn;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * warp) << 16);
    for (int c = 0; c < a.bn; c += 16) {
      uint32_t acc[16];
      tc::tmem_ld_cols<16>(taddr + c, acc);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        __stcg(reinterpret_cast<float4*>(dst + c + i),
               make_float4(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1]), __uint_as_float(acc[i + 2]),
                           __uint_as_float(acc[i + 3])));
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<kTmemCols>(tmem_base);
  (void)lane;
}

// Y[t, c4 ... c4 + 4) = b2 + sum_j part[slot_pos[t, j], c4 ...]  (fp32, ascending expert order) -> bf16
__global__ void __launch_bounds__(256) grouped_combine_kernel(const float* __restrict__ part, const int* __restrict__ slot_pos,
                                                              const float* __restrict__ b2, __nv_bfloat16* __restrict__ Y, int T,
                                                              int k, int d) {
  pdl_wait();
  pdl_launch_dependents();
  const int d4 = d >> 2;
  const long long n = static_cast<long long>(T) * d4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / d4), c4 = static_cast<int>(i - static_cast<long long>(t) * d4);
    float4 acc = b2 != nullptr ? __ldg(reinterpret_cast<const float4*>(b2) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < k; ++j) {
      const int p = __ldg(slot_pos + static_cast<size_t>(t) * k + j);
      if (p < 0) break;             // slots are filled from j = 0; the rest of the row is -1
      const float4 v = __ldcg(reinterpret_cast<const float4*>(part + static_cast<size_t>(p) * d) + c4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
    uint2 out;
    out.x = *reinterpret_cast<uint32_t*>(&lo);
    out.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(Y + static_cast<size_t>(t) * d + 4 * c4) = out;
  }
}

}  // namespace grouped
}  // namespace moe

extern "C" {

size_t moe_down_grouped_rows(int T, int k, int E) {
  // capacity (rows) of perm_tokens / the partial workspace: every list padded to a multiple of 128 rows
  return static_cast<size_t>(T > 0 ? T : 0) * static_cast<size_t>(k > 0 ? k : 0) + static_cast<size_t>(E > 0 ? E : 0) * 128;
}

size_t moe_down_grouped_workspace_bytes(int T, int k, int E, int d) {
  return moe_down_grouped_rows(T, k, E) * static_cast<size_t>(d > 0 ? d : 0) * sizeof(float);
}

int moe_down_grouped(const void* H, const int* perm_offsets, const int* perm_tokens, const int* slot_pos, const void* w2p,
                     const float* b2, void* Y, int T, int h, int d, int E, int es, int k, void* workspace,
                     size_t workspace_bytes, void* stream) {
  using namespace moe;
  using namespace moe::grouped;
  MOE_REQUIRE(T >= 0 && h >= 8 && d >= 8 && E >= 1 && k >= 0 && k <= E, MOE_ERR_INVALID_ARGUMENT,
              "moe_down_grouped: bad sizes T=%d h=%d d=%d E=%d k=%d", T, h, d, E, k);
  if (T == 0) return MOE_OK;
  MOE_REQUIRE(static_cast<long long>(E) * es == h, MOE_ERR_INVALID_ARGUMENT, "moe_down_grouped: E*es=%d*%d != h=%d", E, es, h);
  MOE_REQUIRE(es == kEs, MOE_ERR_UNSUPPORTED_SHAPE,
              "moe_down_grouped: expert size %d unsupported (one expert = one 64-wide bf16 k-block; use moe_down_proj)", es);
  MOE_REQUIRE(d % 16 == 0 && E <= 1024, MOE_ERR_UNSUPPORTED_SHAPE, "moe_down_grouped: needs d %% 16 == 0 and E <= 1024 (d=%d E=%d)", d, E);
  MOE_REQUIRE(H && perm_offsets && perm_tokens && slot_pos && w2p && Y && workspace, MOE_ERR_INVALID_ARGUMENT,
              "moe_down_grouped: NULL pointer");
  MOE_REQUIRE(workspace_bytes >= moe_down_grouped_workspace_bytes(T, k, E, d), MOE_ERR_INVALID_ARGUMENT,
              "moe_down_grouped: workspace of %zu bytes is smaller than moe_down_grouped_workspace_bytes()", workspace_bytes);
  MOE_REQUIRE(((reinterpret_cast<uintptr_t>(H) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(workspace) |
                reinterpret_cast<uintptr_t>(w2p)) & 15) == 0 && (b2 == nullptr || (reinterpret_cast<uintptr_t>(b2) & 15) == 0),
              MOE_ERR_INVALID_ARGUMENT, "moe_down_grouped: H / Y / w2p / b2 / workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  int bn = 0;
  for (int cand = 256; cand >= 16; cand -= 16)
    if (d % cand == 0) {
      bn = cand;
      break;
    }
  Args a = {};
  a.H = static_cast<const __nv_bfloat16*>(H);
  a.offsets = perm_offsets;
  a.tokens = perm_tokens;
  a.part = static_cast<float*>(workspace);
  a.h = h;
  a.d = d;
  a.E = E;
  a.bn = bn;
  a.n_tiles = d / bn;
  a.max_row_tiles = static_cast<int>(moe_down_grouped_rows(T, k, E) / kRows);
  if (k > 0) {
    CUtensorMap tw2;
    int rc = make_tmap_bf16_2d(&tw2, w2p, static_cast<uint64_t>(d), static_cast<uint64_t>(h), static_cast<uint32_t>(bn), kEs, true);
    if (rc) return rc;
    const size_t smem = 1024 + kABytes + static_cast<size_t>(bn) * 128;
    const dim3 grid(static_cast<unsigned>(a.max_row_tiles) * static_cast<unsigned>(a.n_tiles));
    cudaError_t le;
    if (bn <= 64) {
      if ((rc = ensure_dynamic_smem(reinterpret_cast<const void*>(grouped_down_kernel<64>), smem))) return rc;
      le = launch_pdl(grouped_down_kernel<64>, grid, dim3(kRows), smem, st, tw2, a);
    } else if (bn <= 128) {
      if ((rc = ensure_dynamic_smem(reinterpret_cast<const void*>(grouped_down_kernel<128>), smem))) return rc;
      le = launch_pdl(grouped_down_kernel<128>, grid, dim3(kRows), smem, st, tw2, a);
    } else {
      if ((rc = ensure_dynamic_smem(reinterpret_cast<const void*>(grouped_down_kernel<256>), smem))) return rc;
      le = launch_pdl(grouped_down_kernel<256>, grid, dim3(kRows), smem, st, tw2, a);
    }
    if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_down_grouped launch: %s", cudaGetErrorString(le));
    rc = check_launch("moe_down_grouped");
    if (rc) return rc;
  }
  const long long items = static_cast<long long>(T) * (d / 4);
  long long ctas = (items + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (ctas > cap) ctas = cap;
  cudaError_t le = launch_pdl(grouped_combine_kernel, dim3(static_cast<unsigned>(ctas)), dim3(256), 0, st,
                              static_cast<const float*>(workspace), slot_pos, b2, static_cast<__nv_bfloat16*>(Y), T, k, d);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_down_grouped (combine) launch: %s", cudaGetErrorString(le));
  return check_launch("moe_down_grouped (combine)");
}

}  // extern "C"
