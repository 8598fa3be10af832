// K2 router (per-token top-k select + histogram + column max + in-place zeroing of H) and the
// standalone K4 histogram / column-max kernels.  HBM/L2-bound integer & compare work: one warp
// per token, everything warp-uniform after the ballots, no shared-memory traffic in the select.
#include "common.cuh"

namespace moe {

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  // valid for a destination initialised to -inf (or any float) and non-NaN v
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// order-preserving float -> uint32 key (ascending)
__device__ __forceinline__ uint32_t float_key(float s) {
  uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct RouterArgs {
  const float* scores;
  const float* score_bias;   // [E] added to the scores before the selection, or null (AddExperts)
  const uint32_t* removed_bits;
  uint32_t* active_bits;
  int16_t* idx;
  unsigned long long* hist;
  float* colmax;
  __nv_bfloat16* H;
  int k, h, es, T, E, count_begin, count_end;
  uint32_t es_magic;  // floor(2^32 / es) + 1
};

constexpr int kRouterWarps = 8;

template <int SLOTS>
__global__ void __launch_bounds__(kRouterWarps * 32) router_topk_kernel(const RouterArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * kRouterWarps + warp;
  const int nwarps = gridDim.x * kRouterWarps;
  const unsigned full = 0xffffffffu;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int E = a.E, k = a.k;

  __shared__ unsigned int s_hist[SLOTS * 32];
  __shared__ float s_max[kRouterWarps][SLOTS * 32];

  uint32_t removed[SLOTS], valid[SLOTS];
  float bias[SLOTS];
#pragma unroll
  for (int j = 0; j < SLOTS; ++j) {
    const int lo = 32 * j;
    valid[j] = (E >= lo + 32) ? full : (E > lo ? ((1u << (E - lo)) - 1u) : 0u);
    removed[j] = (a.removed_bits != nullptr && lo < E) ? (__ldg(a.removed_bits + j) & valid[j]) : 0u;
    bias[j] = (a.score_bias != nullptr && lo + lane < E) ? __ldg(a.score_bias + lo + lane) : 0.f;
  }

  pdl_wait();                 // scores / H come from the previous kernel in the stream
  pdl_launch_dependents();
  uint32_t cnt[SLOTS];
  float mx[SLOTS];
#pragma unroll
  for (int j = 0; j < SLOTS; ++j) {
    cnt[j] = 0;
    mx[j] = -INFINITY;
  }

  for (int t = gwarp; t < a.T; t += nwarps) {
    uint32_t key[SLOTS];
    const float* row = a.scores + static_cast<size_t>(t) * E;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
      const int e = 32 * j + lane;
      float s = 0.f;
      if (e < E) {
        s = __ldg(row + e);
        mx[j] = fmaxf(mx[j], s);
        s += bias[j];
      }
      if ((removed[j] >> lane) & 1u) s = 0.f;  // zeroed pattern row => score exactly 0
      key[j] = float_key(s);
    }

    uint32_t sel[SLOTS];
    if (k >= E) {
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) sel[j] = valid[j];
    } else {
      uint32_t cand[SLOTS];
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) {
        cand[j] = valid[j];
        sel[j] = 0u;
      }
      int need = k, ccount = E;
      if (need > 0) {
        // skip the key prefix every candidate shares (sign/exponent bits, typically 8-10 rounds)
        uint32_t k_or = 0u, k_and = full;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j)
          if ((valid[j] >> lane) & 1u) {
            k_or |= key[j];
            k_and &= key[j];
          }
        k_or = __reduce_or_sync(full, k_or);
        k_and = __reduce_and_sync(full, k_and);
        const uint32_t diff = k_or ^ k_and;
        int bit = diff ? (31 - __clz(diff)) : -1;
        for (; bit >= 0; --bit) {
          uint32_t b[SLOTS];
          int c1 = 0;
#pragma unroll
          for (int j = 0; j < SLOTS; ++j) {
            b[j] = __ballot_sync(full, (key[j] >> bit) & 1u) & cand[j];
            c1 += __popc(b[j]);
          }
          if (c1 >= need) {
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) cand[j] = b[j];
            ccount = c1;
          } else {
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) {
              sel[j] |= b[j];
              cand[j] &= ~b[j];
            }
            need -= c1;
            ccount -= c1;
          }
          if (ccount == need) {
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) sel[j] |= cand[j];
            need = 0;
            break;
          }
        }
        if (need > 0) {
          // exact ties on the k-th key: lowest expert ids win
#pragma unroll
          for (int j = 0; j < SLOTS; ++j) {
            uint32_t w = cand[j];
            while (need > 0 && w) {
              const uint32_t low = w & (0u - w);
              sel[j] |= low;
              w ^= low;
              --need;
            }
          }
        }
      }
    }

    // this lane's copy of the word it is responsible for (lane j <-> word j)
    uint32_t my_active = 0u;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j)
      if (lane == j) my_active = sel[j] & ~removed[j];

    if (a.active_bits != nullptr && lane < SLOTS && 32 * lane < E)
      a.active_bits[static_cast<size_t>(t) * ((E + 31) >> 5) + lane] = my_active;

    if (a.idx != nullptr) {
      int16_t* out = a.idx + static_cast<size_t>(t) * k;
      int base = 0;
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) {
        if ((sel[j] >> lane) & 1u) out[base + __popc(sel[j] & lt_mask)] = static_cast<int16_t>(32 * j + lane);
        base += __popc(sel[j]);
      }
    }

    if (a.hist != nullptr && t >= a.count_begin && t < a.count_end) {
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) cnt[j] += (sel[j] >> lane) & 1u;
    }

    if (a.H != nullptr) {
      // write-only masking: zero the segments of experts that are not active (no read of H)
      __nv_bfloat16* hrow = a.H + static_cast<size_t>(t) * a.h;
      if ((a.es & 3) == 0) {
        uint2* hw = reinterpret_cast<uint2*>(hrow);
        const int nwords = a.h >> 2;
        for (int w0 = 0; w0 < nwords; w0 += 32) {
          const int w = w0 + lane;
          const bool in = w < nwords;
          const uint32_t e = in ? __umulhi(static_cast<uint32_t>(w) << 2, a.es_magic) : 0u;
          const uint32_t word = __shfl_sync(full, my_active, e >> 5);
          if (in && !((word >> (e & 31u)) & 1u)) hw[w] = make_uint2(0u, 0u);
        }
      } else {
        for (int n0 = 0; n0 < a.h; n0 += 32) {
          const int n = n0 + lane;
          const bool in = n < a.h;
          const uint32_t e = in ? (a.es == 1 ? static_cast<uint32_t>(n) : __umulhi(static_cast<uint32_t>(n), a.es_magic)) : 0u;
          const uint32_t word = __shfl_sync(full, my_active, e >> 5);
          if (in && !((word >> (e & 31u)) & 1u)) hrow[n] = __float2bfloat16(0.f);
        }
      }
    }
  }

  // ---- flush per-CTA aggregates: one global atomic per expert per CTA
  if (a.hist != nullptr) {
    for (int i = threadIdx.x; i < SLOTS * 32; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SLOTS; ++j)
      if (cnt[j]) atomicAdd(&s_hist[32 * j + lane], cnt[j]);
    __syncthreads();
    for (int i = threadIdx.x; i < E; i += blockDim.x)
      if (s_hist[i]) atomicAdd(a.hist + i, static_cast<unsigned long long>(s_hist[i]));
  }
  if (a.colmax != nullptr) {
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) s_max[warp][32 * j + lane] = mx[j];
    __syncthreads();
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
      float m = s_max[0][i];
#pragma unroll
      for (int w = 1; w < kRouterWarps; ++w) m = fmaxf(m, s_max[w][i]);
      if (m > -INFINITY) atomic_max_float(a.colmax + i, m);
    }
  }
}

// ------------------------------------------------------------------ K4 standalone histogram
__device__ __forceinline__ void hist_add_aggregated(unsigned int* bins, int v, int E, int lane) {
  const bool ok = v >= 0 && v < E;
  const unsigned peers = __match_any_sync(0xffffffffu, ok ? v : -1);
  if (ok && lane == (__ffs(peers) - 1)) atomicAdd(&bins[v], static_cast<unsigned int>(__popc(peers)));
}

__global__ void __launch_bounds__(256) hist_accumulate_kernel(const int16_t* __restrict__ idx, long long n, int E,
                                                              unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int bins[];
  for (int i = threadIdx.x; i < E; i += blockDim.x) bins[i] = 0u;
  __syncthreads();
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const bool aligned = (reinterpret_cast<uintptr_t>(idx) & 15) == 0;
  const long long nvec = aligned ? (n >> 3) : 0;  // 8 labels per 16-byte load
  const int4* v4 = reinterpret_cast<const int4*>(idx);
  // full-warp trips so that __match_any_sync always sees 32 lanes
  const long long vec_trips = (nvec + nthreads - 1) / nthreads;
  for (long long it = 0; it < vec_trips; ++it) {
    const long long i = it * nthreads + tid;
    int4 q = make_int4(-1, -1, -1, -1);
    if (i < nvec) q = __ldg(v4 + i);
    const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      hist_add_aggregated(bins, static_cast<int16_t>(w[c] & 0xffff), E, lane);
      hist_add_aggregated(bins, static_cast<int16_t>((w[c] >> 16) & 0xffff), E, lane);
    }
  }
  const long long tail0 = nvec << 3;
  const long long tail_trips = (n - tail0 + nthreads - 1) / nthreads;
  for (long long it = 0; it < tail_trips; ++it) {
    const long long i = tail0 + it * nthreads + tid;
    hist_add_aggregated(bins, i < n ? static_cast<int>(idx[i]) : -1, E, lane);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < E; i += blockDim.x)
    if (bins[i]) atomicAdd(hist + i, static_cast<unsigned long long>(bins[i]));
}

// ------------------------------------------------------------------ column max over tokens
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(256) colmax_kernel(const T* __restrict__ m, int rows, int cols,
                                                     float* __restrict__ out) {
  __shared__ float red[8][33];
  pdl_wait();
  pdl_launch_dependents();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float mx = -INFINITY;
  if (c < cols)
    for (int r = blockIdx.y * 8 + ty; r < rows; r += gridDim.y * 8)
      mx = fmaxf(mx, to_f32<T>(m[static_cast<size_t>(r) * cols + c]));
  red[ty][tx] = mx;
  __syncthreads();
  if (ty == 0 && c < cols) {
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i][tx]);
    if (mx > -INFINITY) atomic_max_float(out + c, mx);
  }
}

}  // namespace moe

extern "C" {

int moe_router_topk(const float* scores, const uint32_t* removed_bits, int k, uint32_t* active_bits,
                    int16_t* idx, unsigned long long* hist, float* score_colmax, void* H, int h, int es,
                    int T, int E, int count_begin, int count_end, void* stream) {
  return moe_router_topk_biased(scores, nullptr, removed_bits, k, active_bits, idx, hist, score_colmax, H, h, es, T, E,
                                count_begin, count_end, stream);
}

int moe_router_topk_biased(const float* scores, const float* score_bias, const uint32_t* removed_bits, int k,
                           uint32_t* active_bits, int16_t* idx, unsigned long long* hist, float* score_colmax, void* H,
                           int h, int es, int T, int E, int count_begin, int count_end, void* stream) {
  using namespace moe;
  MOE_REQUIRE(T >= 0 && E >= 1 && k >= 0 && k <= E, MOE_ERR_INVALID_ARGUMENT,
              "moe_router_topk: need T>=0, 1<=E, 0<=k<=E (T=%d E=%d k=%d)", T, E, k);
  if (T == 0) return MOE_OK;   // empty tensors have no storage to point at
  MOE_REQUIRE(scores != nullptr, MOE_ERR_INVALID_ARGUMENT, "moe_router_topk: scores is NULL");
  MOE_REQUIRE(E <= 1024, MOE_ERR_UNSUPPORTED_SHAPE, "moe_router_topk: E=%d > 1024 experts", E);
  if (H != nullptr) {
    MOE_REQUIRE(es >= 1 && h == E * es && h < 65536, MOE_ERR_INVALID_ARGUMENT,
                "moe_router_topk: H given but h=%d != E*es=%d*%d (or h >= 65536)", h, E, es);
    MOE_REQUIRE((reinterpret_cast<uintptr_t>(H) & 7) == 0, MOE_ERR_INVALID_ARGUMENT,
                "moe_router_topk: H must be 8-byte aligned");
  }
  if (T == 0) return MOE_OK;
  RouterArgs a;
  a.scores = scores;
  a.score_bias = score_bias;
  a.removed_bits = removed_bits;
  a.active_bits = active_bits;
  a.idx = idx;
  a.hist = hist;
  a.colmax = score_colmax;
  a.H = static_cast<__nv_bfloat16*>(H);
  a.k = k;
  a.h = h;
  a.es = es;
  a.T = T;
  a.E = E;
  a.count_begin = count_begin;
  a.count_end = count_end;
  a.es_magic = es >= 1 ? static_cast<uint32_t>((0x100000000ull / static_cast<unsigned>(es)) + 1ull) : 0u;
  const int ctas_needed = (T + kRouterWarps - 1) / kRouterWarps;
  const int max_ctas = sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
  const int grid = ctas_needed < max_ctas ? ctas_needed : max_ctas;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int slots = (E + 31) / 32;
  cudaError_t le = cudaSuccess;
  if (slots <= 1)
    le = launch_pdl(router_topk_kernel<1>, dim3(grid), dim3(kRouterWarps * 32), 0, st, a);
  else if (slots <= 2)
    le = launch_pdl(router_topk_kernel<2>, dim3(grid), dim3(kRouterWarps * 32), 0, st, a);
  else if (slots <= 4)
    le = launch_pdl(router_topk_kernel<4>, dim3(grid), dim3(kRouterWarps * 32), 0, st, a);
  else if (slots <= 8)
    le = launch_pdl(router_topk_kernel<8>, dim3(grid), dim3(kRouterWarps * 32), 0, st, a);
  else if (slots <= 16)
    le = launch_pdl(router_topk_kernel<16>, dim3(grid), dim3(kRouterWarps * 32), 0, st, a);
  else
    le = launch_pdl(router_topk_kernel<32>, dim3(grid), dim3(kRouterWarps * 32), 0, st, a);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_router_topk launch: %s", cudaGetErrorString(le));
  return check_launch("moe_router_topk");
}

int moe_hist_accumulate(const int16_t* idx, long long n, int E, unsigned long long* hist, void* stream) {
  using namespace moe;
  MOE_REQUIRE(n >= 0 && E >= 1 && E <= 8192, MOE_ERR_INVALID_ARGUMENT, "moe_hist_accumulate: n=%lld E=%d", n, E);
  if (n == 0) return MOE_OK;
  MOE_REQUIRE(idx != nullptr && hist != nullptr, MOE_ERR_INVALID_ARGUMENT, "moe_hist_accumulate: NULL pointer");
  const long long per_cta = 256LL * 8 * 8;  // 8 vector loads of 8 labels per thread
  long long ctas = (n + per_cta - 1) / per_cta;
  const long long max_ctas = static_cast<long long>(sm_count()) * 8;
  if (ctas > max_ctas) ctas = max_ctas;
  cudaError_t le = launch_pdl(hist_accumulate_kernel, dim3(static_cast<unsigned>(ctas)), dim3(256), E * sizeof(unsigned int),
                              static_cast<cudaStream_t>(stream), idx, n, E, hist);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_hist_accumulate launch: %s", cudaGetErrorString(le));
  return check_launch("moe_hist_accumulate");
}

static int colmax_grid_y(int rows) {
  int gy = (rows + 63) / 64;
  const int cap = moe::sm_count() * 2;
  return gy < 1 ? 1 : (gy > cap ? cap : gy);
}

int moe_colmax_f32(const float* m, int T, int C, float* out, void* stream) {
  using namespace moe;
  MOE_REQUIRE((m != nullptr || T == 0) && out != nullptr && T >= 0 && C >= 1, MOE_ERR_INVALID_ARGUMENT, "moe_colmax_f32: bad args");
  if (T == 0) return MOE_OK;
  dim3 grid((C + 31) / 32, colmax_grid_y(T));
  cudaError_t le = launch_pdl(colmax_kernel<float>, grid, dim3(256), 0, static_cast<cudaStream_t>(stream), m, T, C, out);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_colmax_f32 launch: %s", cudaGetErrorString(le));
  return check_launch("moe_colmax_f32");
}

int moe_colmax_bf16(const void* m, int T, int C, float* out, void* stream) {
  using namespace moe;
  MOE_REQUIRE((m != nullptr || T == 0) && out != nullptr && T >= 0 && C >= 1, MOE_ERR_INVALID_ARGUMENT, "moe_colmax_bf16: bad args");
  if (T == 0) return MOE_OK;
  dim3 grid((C + 31) / 32, colmax_grid_y(T));
  cudaError_t le = launch_pdl(colmax_kernel<__nv_bfloat16>, grid, dim3(256), 0, static_cast<cudaStream_t>(stream),
                              static_cast<const __nv_bfloat16*>(m), T, C, out);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_colmax_bf16 launch: %s", cudaGetErrorString(le));
  return check_launch("moe_colmax_bf16");
}

}  // extern "C"
