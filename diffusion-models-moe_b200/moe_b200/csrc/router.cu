// K2 router (per-token top-k select + histogram + column max + in-place zeroing of H) and the
// standalone K4 histogram / column-max kernels.  HBM/L2-bound integer & compare work: one warp
// per token, everything warp-uniform after the ballots, no shared-memory traffic in the select.
#include "common.cuh"
#include "route.cuh"

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

// ------------------------------------------------------------------ K2, several tokens per warp (E <= 256)
// The warp-per-token kernel above spends ~700 warp instructions per token (about 20 dependent ballot rounds): at the
// UNet-batch-16 shapes it runs at 1.3 G tokens/s = 8 % of the HBM rate of its 4E-byte score rows.  This kernel routes
// with the code of the fused layer kernel's routing stage (route.cuh): L = 4 / 8 / 16 lanes per token (8 / 4 / 2
// tokens per warp, <= 16 experts per lane in registers), the k-th largest key from a bitonic sort -- ~60-90 warp
// instructions per token -- 16-byte score loads, shared-memory histogram bins, whole-warp 16-byte zero stores.
struct RouterGeom {
  int lanes, lanes_log2, E, k, words, mask_h, h, count_begin, count_end, T, chunk_tokens;
  uint32_t es_magic;
};
struct RouterPtrs {
  const float* scores;
  const float* score_bias;
  const uint32_t* removed_bits;
  uint32_t* active_bits;
  int16_t* idx;
  unsigned long long* hist;
  float* colmax;
  __nv_bfloat16* H;
};

constexpr int kMultiWarps = 8;

template <int KPT, bool kBias, bool kColmax>
__global__ void __launch_bounds__(kMultiWarps * 32) router_multi_kernel(const RouterGeom g, const RouterPtrs a) {
  __shared__ uint32_t s_words[kMultiWarps][16];
  __shared__ unsigned int s_hist[256];
  __shared__ float s_max[kColmax ? kMultiWarps * 32 * KPT : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  pdl_wait();                 // scores / H come from the previous kernel in the stream
  pdl_launch_dependents();
  float mx[KPT];
#pragma unroll
  for (int i = 0; i < KPT; ++i) mx[i] = -INFINITY;
  const long long n_chunks = (static_cast<long long>(g.T) + g.chunk_tokens - 1) / g.chunk_tokens;
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x)
    route::route_chunk<KPT, kBias, kColmax>(g, a, static_cast<int>(c * g.chunk_tokens), g.T, warp, lane, s_words[warp],
                                            s_hist, mx);
  __syncthreads();
  if (a.hist != nullptr)
    for (int i = threadIdx.x; i < g.E; i += blockDim.x)
      if (s_hist[i]) atomicAdd(a.hist + i, static_cast<unsigned long long>(s_hist[i]));
  if constexpr (kColmax) {
    // lane `part` of every token group holds experts [part * KPT, part * KPT + KPT)
#pragma unroll
    for (int i = 0; i < KPT; ++i) s_max[(warp * 32 + lane) * KPT + i] = mx[i];
    __syncthreads();
    const int L = g.lanes;
    for (int e = threadIdx.x; e < g.E; e += blockDim.x) {
      const int part = e / KPT, i = e - part * KPT;
      float m = -INFINITY;
      for (int w = 0; w < kMultiWarps; ++w)
        for (int tl = 0; tl < 32 / L; ++tl) m = fmaxf(m, s_max[(w * 32 + tl * L + part) * KPT + i]);
      if (m > -INFINITY) atomic_max_float(a.colmax + e, m);
    }
  }
}

template <int KPT>
static cudaError_t launch_multi(const RouterGeom& g, const RouterPtrs& a, int grid, cudaStream_t st) {
  const bool bias = a.score_bias != nullptr, cmax = a.colmax != nullptr;
  if (bias && cmax) return launch_pdl(router_multi_kernel<KPT, true, true>, dim3(grid), dim3(kMultiWarps * 32), 0, st, g, a);
  if (bias) return launch_pdl(router_multi_kernel<KPT, true, false>, dim3(grid), dim3(kMultiWarps * 32), 0, st, g, a);
  if (cmax) return launch_pdl(router_multi_kernel<KPT, false, true>, dim3(grid), dim3(kMultiWarps * 32), 0, st, g, a);
  return launch_pdl(router_multi_kernel<KPT, false, false>, dim3(grid), dim3(kMultiWarps * 32), 0, st, g, a);
}

// ------------------------------------------------------------------ compacted token -> expert permutation
// From the expert-set words of the router: for every expert e the ascending list of the tokens that selected it,
// perm_tokens[perm_offsets[e] ... + perm_counts[e]), segments padded to a multiple of `row_pad` rows (pad entries -1),
// and the inverse map slot_pos[t * k + j] = position of token t's j-th active expert (ascending expert id; -1 beyond
// the token's active count).  Deterministic: a stable counting sort -- per (word, token segment, warp) counts, then an
// ordered scatter -- no atomics.  Grid (words, segments); warp q of a CTA owns a contiguous token sub-range.
constexpr int kPermWarps = 8;
constexpr int kPermMaxSegments = 64;

struct PermArgs {
  const uint32_t* bits;   // [T, W]
  int T, E, W, k, row_pad, seg_tokens, n_segments;
  int* seg_counts;        // [W * 32][n_segments]
  int* offsets;           // [E + 1]
  int* counts;            // [E]
  int* tokens;            // [capacity]
  int* slot_pos;          // [T, k]
};

// per-warp counts of one token sub-range: lane b ends up with the number of tokens whose word has bit b set
__device__ __forceinline__ int perm_count_range(const PermArgs& p, int w, int t0, int t1, int lane) {
  int cnt = 0;
  for (int base = t0; base < t1; base += 32) {
    const int t = base + lane;
    const uint32_t word = (t < t1) ? __ldg(p.bits + static_cast<size_t>(t) * p.W + w) : 0u;
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      const unsigned m = __ballot_sync(0xffffffffu, (word >> b) & 1u);
      if (lane == b) cnt += __popc(m);
    }
  }
  return cnt;
}

__device__ __forceinline__ void perm_warp_range(const PermArgs& p, int seg, int warp, int& t0, int& t1) {
  const int s0 = seg * p.seg_tokens;
  const int s1 = min(p.T, s0 + p.seg_tokens);
  const int per = ((p.seg_tokens / kPermWarps) + 31) / 32 * 32;
  t0 = min(s1, s0 + warp * per);
  t1 = (warp == kPermWarps - 1) ? s1 : min(s1, t0 + per);
}

__global__ void __launch_bounds__(kPermWarps * 32) perm_count_kernel(const PermArgs p) {
  __shared__ int s_cnt[kPermWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w = blockIdx.x, seg = blockIdx.y;
  pdl_wait();
  pdl_launch_dependents();
  int t0, t1;
  perm_warp_range(p, seg, warp, t0, t1);
  s_cnt[warp][lane] = perm_count_range(p, w, t0, t1, lane);
  __syncthreads();
  if (warp == 0) {
    int tot = 0;
#pragma unroll
    for (int q = 0; q < kPermWarps; ++q) tot += s_cnt[q][lane];
    p.seg_counts[(w * 32 + lane) * p.n_segments + seg] = tot;
  }
}

__global__ void __launch_bounds__(kPermWarps * 32) perm_scatter_kernel(const PermArgs p) {
  __shared__ int s_cnt[kPermWarps][32];
  __shared__ int s_base[32];       // first position of this CTA's segment for each of the word's experts
  __shared__ int s_total[32], s_off[33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w = blockIdx.x, seg = blockIdx.y;
  pdl_wait();
  pdl_launch_dependents();
  int t0, t1;
  perm_warp_range(p, seg, warp, t0, t1);
  s_cnt[warp][lane] = perm_count_range(p, w, t0, t1, lane);
  // padded offsets of this word's experts: sum over all experts below (every CTA redoes the small scan)
  if (warp == 0) {
    int before = 0;      // padded rows of all experts of lower words
    for (int e = lane; e < w * 32; e += 32) {
      int c = 0;
      for (int s2 = 0; s2 < p.n_segments; ++s2) c += __ldg(p.seg_counts + e * p.n_segments + s2);
      before += (c + p.row_pad - 1) / p.row_pad * p.row_pad;
    }
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    int mine = 0, below_seg = 0;
    const int e = w * 32 + lane;
    if (e < p.E)
      for (int s2 = 0; s2 < p.n_segments; ++s2) {
        const int c = __ldg(p.seg_counts + e * p.n_segments + s2);
        mine += c;
        if (s2 < seg) below_seg += c;
      }
    const int padded = (mine + p.row_pad - 1) / p.row_pad * p.row_pad;
    int incl = padded;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int off = before + incl - padded;
    s_off[lane] = off;
    if (lane == 31) s_off[32] = before + incl;
    s_total[lane] = mine;
    s_base[lane] = off + below_seg;
    if (seg == 0 && e < p.E) {
      p.offsets[e] = off;
      p.counts[e] = mine;
      if (e == p.E - 1) p.offsets[p.E] = off + padded;
    }
  }
  __syncthreads();
  // this warp's first position per expert: segment base + the counts of the warps before it
  int run = s_base[lane];
  for (int q = 0; q < warp; ++q) run += s_cnt[q][lane];
  const uint32_t lt = (1u << lane) - 1u;
  for (int base = t0; base < t1; base += 32) {
    const int t = base + lane;
    const bool ok = t < t1;
    const uint32_t word = ok ? __ldg(p.bits + static_cast<size_t>(t) * p.W + w) : 0u;
    int j0 = 0;          // active experts of this token in lower words
    if (ok)
      for (int w2 = 0; w2 < w; ++w2) j0 += __popc(__ldg(p.bits + static_cast<size_t>(t) * p.W + w2));
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      const unsigned m = __ballot_sync(0xffffffffu, (word >> b) & 1u);
      const int start = __shfl_sync(0xffffffffu, run, b);
      if ((word >> b) & 1u) {
        const int pos = start + __popc(m & lt);
        p.tokens[pos] = t;
        const int j = j0 + __popc(word & ((1u << b) - 1u));
        if (j < p.k) p.slot_pos[static_cast<size_t>(t) * p.k + j] = pos;
      }
      if (lane == b) run += __popc(m);
    }
    if (ok && w == p.W - 1)
      for (int j = j0 + __popc(word); j < p.k; ++j) p.slot_pos[static_cast<size_t>(t) * p.k + j] = -1;
  }
  // the last segment's CTA pads the tail of each of its experts' lists
  if (seg == p.n_segments - 1) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * p.row_pad; i += blockDim.x) {
      const int e_l = i / p.row_pad, r = i - e_l * p.row_pad;
      if (w * 32 + e_l >= p.E) break;
      const int begin = s_off[e_l] + s_total[e_l] + r;
      if (begin < s_off[e_l + 1]) p.tokens[begin] = -1;
    }
  }
}

// ------------------------------------------------------------------ K4 standalone histogram
// 16-byte loads (8 labels per thread and load, four loads in flight), THREAD-PRIVATE counters in shared memory: a thread
// only ever touches its own counters, so counting is a plain load / add / store -- no atomics, no match / ballot
// aggregation -- folded once per CTA into one int64 global atomic per expert.  Thread t's counters sit in the 32-bit
// words row * kHistThreads + t, so the 32 lanes of a warp always hit 32 different banks, whatever their labels:
//   kWide   (E <= kHistWideMaxE)  one 32-bit counter per expert (row = expert): 6 instructions per label
//   packed                        two 16-bit counters per word (row = expert / 2, experts 2i / 2i + 1 in the low / high
//                                 half): 9 instructions per label, half the shared memory; the grid is sized so that a
//                                 thread counts at most kHistMaxPerThread labels and a half cannot carry into the other
// Labels outside [0, E) (padding, -1) are clamped to one extra dump row instead of being branched around: the loop is
// straight-line code with 32-bit shared addresses.  (The earlier versions were instruction-bound, not memory-bound: a
// compare + branch + reconvergence pair + generic-address arithmetic per label, ~16 instructions, 2.5-2.7 TB/s whether
// the 16-bit cells conflicted 2-way or not; round 1's __match_any_sync aggregation: 159 GB/s; per-warp bins with shared
// atomics: 921 GB/s.)
constexpr int kHistThreads = 128;
constexpr int kHistMaxPerThread = 60000;
constexpr int kHistWideMaxE = 64;

template <bool kWide>
__global__ void __launch_bounds__(kHistThreads) hist_accumulate_kernel(const int16_t* __restrict__ idx, long long n, int E,
                                                                       unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int cnt[];        // [rows + 1][kHistThreads], rows = E (wide) or (E + 1) / 2 (packed)
  const int rows = kWide ? E : (E + 1) >> 1;
  for (int i = threadIdx.x; i < (rows + 1) * kHistThreads; i += blockDim.x) cnt[i] = 0u;
  __syncthreads();
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t mine = static_cast<uint32_t>(__cvta_generic_to_shared(cnt + threadIdx.x));
  const unsigned int dump = kWide ? static_cast<unsigned int>(rows) : 2u * static_cast<unsigned int>(rows);
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const bool aligned = (reinterpret_cast<uintptr_t>(idx) & 15) == 0;
  const long long nvec = aligned ? (n >> 3) : 0;  // 8 labels per 16-byte load
  const int4* v4 = reinterpret_cast<const int4*>(idx);
  // a label as an unsigned 16-bit value: negative labels (padding) are >= 32768 > E and land in the dump row
  auto add = [&](unsigned int v) {
    const unsigned int e = min(v, dump);
    uint32_t addr, inc;
    if constexpr (kWide) {
      addr = mine + e * (kHistThreads * 4u);
      inc = 1u;
    } else {
      addr = mine + (e & ~1u) * (kHistThreads * 2u);
      inc = (e & 1u) * 0xffffu + 1u;
    }
    uint32_t c;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c) : "r"(addr) : "memory");
    c += inc;
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(c) : "memory");
  };
  auto add8 = [&](const int4& q) {
    const unsigned int wv[4] = {static_cast<unsigned int>(q.x), static_cast<unsigned int>(q.y), static_cast<unsigned int>(q.z),
                                static_cast<unsigned int>(q.w)};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      add(wv[c] & 0xffffu);
      add(wv[c] >> 16);
    }
  };
  // software-pipelined: the next four 16-byte loads of a thread are in flight while it counts the current 32 labels
  // (counting is a serial load / add / store chain of ~40 cycles per label: without the prefetch a thread had no load
  // outstanding for 40 % of its time)
  long long i = tid;
  int4 q[4];
  bool have = i + 3 * nthreads < nvec;
  if (have) {
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = __ldg(v4 + i + u * nthreads);
  }
  while (have) {
    const long long i_next = i + 4 * nthreads;
    const bool have_next = i_next + 3 * nthreads < nvec;
    int4 qn[4];
    if (have_next) {
#pragma unroll
      for (int u = 0; u < 4; ++u) qn[u] = __ldg(v4 + i_next + u * nthreads);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) add8(q[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = qn[u];
    i = i_next;
    have = have_next;
  }
  for (; i < nvec; i += nthreads) add8(__ldg(v4 + i));
  for (long long j = (nvec << 3) + tid; j < n; j += nthreads) add(static_cast<unsigned int>(static_cast<unsigned short>(idx[j])));
  // fold: one thread per expert sums the 128 thread-private counters of its row
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    unsigned int tot = 0;
    const unsigned int* row = cnt + (kWide ? e : (e >> 1)) * kHistThreads;
    const int sh = kWide ? 0 : ((e & 1) << 4);
    const unsigned int msk = kWide ? 0xffffffffu : 0xffffu;
    // rotate the start so that the threads of a warp (consecutive experts) spread over the banks
    for (int j = 0; j < kHistThreads; ++j) tot += (row[(j + e) & (kHistThreads - 1)] >> sh) & msk;
    if (tot) atomicAdd(hist + e, static_cast<unsigned long long>(tot));
  }
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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t le = cudaSuccess;
  // Which kernel: the several-tokens-per-warp kernel needs enough tokens to fill the machine (a CTA routes 16-64 tokens
  // per pass with long in-register sorting networks: at the UNet-batch-2 shapes of the d >= 640 layers it launches 32-64
  // CTAs and takes 13 us where the warp-per-token kernel takes 6; at T = 65 536 it is 2.5x faster; tie at T = 8192).
  // MOE_ROUTER_LEGACY=1 / =0 force the warp-per-token / the several-tokens-per-warp kernel (tests, A-B timing).
  const char* legacy_env = getenv("MOE_ROUTER_LEGACY");
  const bool legacy = legacy_env != nullptr ? atoi(legacy_env) != 0 : T < 12288;
  const bool vec_ok = H == nullptr || (es % 4 == 0 && h % 8 == 0 && (reinterpret_cast<uintptr_t>(H) & 15) == 0);
  if (E <= 256 && vec_ok && !legacy) {
    // several tokens per warp (route.cuh): the smallest lane count per token that keeps <= 16 experts per lane
    RouterGeom g = {};
    int L = 4;
    while (L < 32 && (E + L - 1) / L > 16) L <<= 1;
    if (const char* le_env = getenv("MOE_ROUTER_LANES")) {      // experiments: more lanes per token, fewer keys per lane
      const int v = atoi(le_env);
      if ((v == 8 || v == 16 || v == 32) && v > L) L = v;
    }
    int kpt = 1;
    while (kpt * L < E) kpt <<= 1;
    g.lanes = L;
    g.lanes_log2 = L == 4 ? 2 : (L == 8 ? 3 : (L == 16 ? 4 : 5));
    g.E = E;
    g.k = k;
    g.words = (E + 31) / 32;
    g.mask_h = H != nullptr ? 1 : 0;
    g.h = h;
    g.count_begin = count_begin;
    g.count_end = count_end;
    g.T = T;
    g.chunk_tokens = kMultiWarps * (32 / L);
    g.es_magic = a.es_magic;
    RouterPtrs rp = {scores, score_bias, removed_bits, active_bits, idx, hist, score_colmax, static_cast<__nv_bfloat16*>(H)};
    const long long chunks = (static_cast<long long>(T) + g.chunk_tokens - 1) / g.chunk_tokens;
    const long long cap = static_cast<long long>(sm_count()) * 6;
    const int grid = static_cast<int>(chunks < cap ? chunks : cap);
    switch (kpt) {
      case 16: le = launch_multi<16>(g, rp, grid, st); break;
      case 8: le = launch_multi<8>(g, rp, grid, st); break;
      case 4: le = launch_multi<4>(g, rp, grid, st); break;
      case 2: le = launch_multi<2>(g, rp, grid, st); break;
      default: le = launch_multi<1>(g, rp, grid, st); break;
    }
    if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_router_topk launch: %s", cudaGetErrorString(le));
    return check_launch("moe_router_topk");
  }
  const int ctas_needed = (T + kRouterWarps - 1) / kRouterWarps;
  const int max_ctas = sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
  const int grid = ctas_needed < max_ctas ? ctas_needed : max_ctas;
  const int slots = (E + 31) / 32;
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
  MOE_REQUIRE(n >= 0 && E >= 1, MOE_ERR_INVALID_ARGUMENT, "moe_hist_accumulate: n=%lld E=%d", n, E);
  if (n == 0) return MOE_OK;
  MOE_REQUIRE(idx != nullptr && hist != nullptr, MOE_ERR_INVALID_ARGUMENT, "moe_hist_accumulate: NULL pointer");
  MOE_REQUIRE(E <= 768, MOE_ERR_UNSUPPORTED_SHAPE, "moe_hist_accumulate: E=%d > 768 bins (thread-private counters exceed shared memory)", E);
  bool wide = E <= kHistWideMaxE;
  if (const char* e = getenv("MOE_HIST_WIDE")) wide = atoi(e) != 0 && E <= 96;      // A-B timing
  const size_t smem = static_cast<size_t>((wide ? E : (E + 1) / 2) + 1) * kHistThreads * sizeof(unsigned int);
  const void* kfn = wide ? reinterpret_cast<const void*>(hist_accumulate_kernel<true>)
                         : reinterpret_cast<const void*>(hist_accumulate_kernel<false>);
  int rc = ensure_dynamic_smem(kfn, smem);
  if (rc) return rc;
  const long long per_cta = static_cast<long long>(kHistThreads) * 8 * 8;  // 8 vector loads of 8 labels per thread
  long long ctas = (n + per_cta - 1) / per_cta;
  const int per_sm = smem > 0 ? static_cast<int>((200 * 1024) / smem) : 16;
  const long long max_ctas = static_cast<long long>(sm_count()) * (per_sm < 1 ? 1 : (per_sm > 16 ? 16 : per_sm));
  if (ctas > max_ctas) ctas = max_ctas;
  // 16-bit thread-private counters (packed layout): bound the labels per thread; 32-bit ones hold any count a launch can reach
  const long long min_ctas = wide ? 1 : (n + static_cast<long long>(kHistThreads) * kHistMaxPerThread - 1) / (static_cast<long long>(kHistThreads) * kHistMaxPerThread);
  if (ctas < min_ctas) ctas = min_ctas;
  MOE_REQUIRE(ctas <= 0x7fffffffLL, MOE_ERR_UNSUPPORTED_SHAPE, "moe_hist_accumulate: n=%lld too large", n);
  cudaError_t le = wide ? launch_pdl(hist_accumulate_kernel<true>, dim3(static_cast<unsigned>(ctas)), dim3(kHistThreads), smem,
                                     static_cast<cudaStream_t>(stream), idx, n, E, hist)
                        : launch_pdl(hist_accumulate_kernel<false>, dim3(static_cast<unsigned>(ctas)), dim3(kHistThreads), smem,
                                     static_cast<cudaStream_t>(stream), idx, n, E, hist);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_hist_accumulate launch: %s", cudaGetErrorString(le));
  return check_launch("moe_hist_accumulate");
}

size_t moe_expert_permutation_workspace_bytes(int T, int E) {
  (void)T;
  return static_cast<size_t>((E + 31) / 32 * 32) * moe::kPermMaxSegments * sizeof(int);
}

int moe_expert_permutation(const uint32_t* active_bits, int T, int E, int k, int row_pad, int* perm_offsets, int* perm_counts,
                           int* perm_tokens, int* slot_pos, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace moe;
  MOE_REQUIRE(T >= 0 && E >= 1 && E <= 1024 && k >= 0 && k <= E && row_pad >= 1 && row_pad <= 256, MOE_ERR_INVALID_ARGUMENT,
              "moe_expert_permutation: T=%d E=%d k=%d row_pad=%d", T, E, k, row_pad);
  MOE_REQUIRE(perm_offsets && perm_counts && workspace, MOE_ERR_INVALID_ARGUMENT, "moe_expert_permutation: NULL offsets / counts / workspace");
  MOE_REQUIRE(workspace_bytes >= moe_expert_permutation_workspace_bytes(T, E), MOE_ERR_INVALID_ARGUMENT,
              "moe_expert_permutation: workspace of %zu bytes is too small", workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (T == 0 || k == 0) {     // no token selects anything: every list is empty
    cudaError_t e = cudaMemsetAsync(perm_offsets, 0, sizeof(int) * (E + 1), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(perm_counts, 0, sizeof(int) * E, st);
    if (e != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_expert_permutation: %s", cudaGetErrorString(e));
    return MOE_OK;
  }
  MOE_REQUIRE(active_bits && perm_tokens && slot_pos, MOE_ERR_INVALID_ARGUMENT, "moe_expert_permutation: NULL bits / tokens / slot_pos");
  PermArgs p = {};
  p.bits = active_bits;
  p.T = T;
  p.E = E;
  p.W = (E + 31) / 32;
  p.k = k;
  p.row_pad = row_pad;
  int seg = (T + kPermMaxSegments - 1) / kPermMaxSegments;
  seg = (seg + 255) / 256 * 256;            // whole 32-token groups per warp
  if (seg < 256) seg = 256;
  p.seg_tokens = seg;
  p.n_segments = (T + seg - 1) / seg;
  p.seg_counts = static_cast<int*>(workspace);
  p.offsets = perm_offsets;
  p.counts = perm_counts;
  p.tokens = perm_tokens;
  p.slot_pos = slot_pos;
  dim3 grid(static_cast<unsigned>(p.W), static_cast<unsigned>(p.n_segments));
  cudaError_t le = launch_pdl(perm_count_kernel, grid, dim3(kPermWarps * 32), 0, st, p);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_expert_permutation (count) launch: %s", cudaGetErrorString(le));
  int rc = check_launch("moe_expert_permutation (count)");
  if (rc) return rc;
  le = launch_pdl(perm_scatter_kernel, grid, dim3(kPermWarps * 32), 0, st, p);
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_expert_permutation (scatter) launch: %s", cudaGetErrorString(le));
  return check_launch("moe_expert_permutation (scatter)");
}

int moe_router_topk_perm(const float* scores, const uint32_t* removed_bits, int k, uint32_t* active_bits, int16_t* idx,
                         unsigned long long* hist, int* perm_offsets, int* perm_counts, int* perm_tokens, int* slot_pos,
                         int row_pad, int T, int E, int count_begin, int count_end, void* workspace, size_t workspace_bytes,
                         void* stream) {
  using namespace moe;
  MOE_REQUIRE(active_bits != nullptr || T == 0, MOE_ERR_INVALID_ARGUMENT, "moe_router_topk_perm: active_bits is required");
  int rc = moe_router_topk_biased(scores, nullptr, removed_bits, k, active_bits, idx, hist, nullptr, nullptr, 0, 0, T, E,
                                  count_begin, count_end, stream);
  if (rc) return rc;
  return moe_expert_permutation(active_bits, T, E, k, row_pad, perm_offsets, perm_counts, perm_tokens, slot_pos, workspace,
                                workspace_bytes, stream);
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
