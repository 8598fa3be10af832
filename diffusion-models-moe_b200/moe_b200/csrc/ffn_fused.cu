// Fused MoEfied GEGLU feed-forward for sm_100a: ONE persistent dataflow kernel per FFN layer call
//
//     phase 1   GEGLU up-projection  H = v * act(g), expert scores          (tcgen05, cta_group::2)
//     routing   per-token top-k over the expert scores, histogram, write-only masking of H
//     phase 3   down-projection      Y = H W2^T + b2                        (tcgen05, cta_group::2)
//
// replacing the K1 -> K2 -> K3 launch triple (gemm_tc.cu / router.cu) on the hot path: the in-kernel timeline
// (tools/trace_timeline.py) showed ~2.6 us from CTA entry to the first MMA plus ~1 us of tail for every GEMM
// launch and a router launch in between, i.e. more fixed cost than tensor time for the SD-1.5 layer shapes.
//
// CTA pairs (clusters of 2, one CTA per SM, grid = all SMs) walk a static work list: first their phase-1
// tiles (256 token rows x nv neuron pairs), then their phase-3 items (256 rows x bn outputs x one K slice).
// Cross-CTA dependencies go through three int arrays in a small global workspace, per 128-row block m:
//     done[m]    += 1 for every phase-1 tile of the block whose H tile (TMA store) and scores are written
//     ticket[m]  hands out the block's routing chunks (one chunk = 512 / tpt tokens = one pass of the 16
//                epilogue warps) to whichever CTA asks first: the CTA that completed the block, or a CTA
//                that is about to need the block in phase 3
//     ready[m]   += 1 per routed chunk; a phase-3 A-tile producer waits for ready[m] == chunks per block
// All CTAs are co-resident (grid <= SM count, 1 CTA / SM) and every CTA finishes its phase-1 tiles, which
// never wait on another CTA, before it waits for anything, so the waits cannot deadlock.  The last CTA to
// exit zeroes the arrays again (the workspace must be zero before the first launch).
//
// Warp roles (640 threads):
//   warp 0      A-tile TMA producer (x in phase 1, H in phase 3 -- waits for ready[m])
//   warp 1      MMA issuer (leader CTA of the pair): tcgen05.mma.cta_group::2, accumulators in TMEM (2 stages)
//   warp 2      TMEM allocator, then the store / sync warp: TMA-stores finished output tiles from the smem
//               staging buffers, publishes done[m], claims routing chunks and posts them to the epilogue warps
//   warp 3      B-tile TMA producer (W1 / W2 slices; weights have no dependencies, so it runs ahead)
//   warps 4-19  epilogue: TMEM -> registers -> bias / exact GELU / product (packed f32x2 math) -> bf16 ->
//               smem staging; expert scores; routing chunks on request; split-K partials and reduction
#include "common.cuh"
#include "tcgen05.cuh"

#ifndef MOE_TRACE
#define MOE_TRACE 0
#endif

namespace moe {
namespace fused {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kAccStride = 256;
constexpr int kTmemCols = 512;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kNumThreads = kEpiWarp0 * 32 + kEpiThreads;
constexpr int kMaxStages = 8;
constexpr int kSmemLimit = 232448;
constexpr int kBiasBytesPerWarp = 128 * 4;          // 2 x 64 floats
constexpr int kSpartPerRow = 8;                      // partial / per-expert score slots per token row and tile
constexpr int kKeys = 16;                            // experts per routing thread
constexpr int kRouteWordsPerWarp = 32;
constexpr int kMaxExperts = 512;
constexpr int kSyncHeaderInts = 16;                  // [0] exit counter
constexpr int kMaxBlocks = 16384;                    // 128-row blocks the sync arrays can hold
constexpr size_t kSyncBytes = (kSyncHeaderInts + 3 * static_cast<size_t>(kMaxBlocks)) * 4;
constexpr size_t kSplitCounterBytes = 64 * 1024;

#if MOE_TRACE
__device__ unsigned long long g_trace[256 * 64];
#define TRACE(slot)                                                                                       \
  do {                                                                                                    \
    if (blockIdx.x < 256 && (slot) < 64) g_trace[blockIdx.x * 64 + (slot)] = tc::global_timer_ns();       \
  } while (0)
#else
#define TRACE(slot) \
  do {              \
  } while (0)
#endif

struct Barriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t hs_full[2];     // staging buffer written by the 16 epilogue warps
  uint64_t hs_empty[2];    // staging buffer stored by the sync warp
  uint64_t route_req;      // sync warp -> epilogue warps: a routing chunk is posted
  uint64_t route_done;     // 16 epilogue warps -> sync warp
  uint64_t fin;            // sync warp has nothing more to post
  uint32_t tmem_base;
  int req_block, req_chunk;
  int last_cta;
};

struct Shape {
  int T, d, h, E, es, k;
  int m_pairs;                                  // 256-row work units
  // phase 1
  int nv, n_tiles1, ks1, nkb1, items1;
  int experts_per_tile, chunks_per_expert, span;
  // phase 3
  int bn, n_tiles3, ks3, nkb3, split3, kb_per_slice3, items3;
  // pipeline
  int stages, slot_bytes, hs_bytes;             // hs_bytes: one phase-1 staging buffer (two of them; phase 3 uses both)
  // routing
  int tpt, tpt_log2, chunks_per_block;          // threads per token; chunk = 512 / tpt tokens
  int act, mask_h, count_begin, count_end;
  uint32_t es_magic;
};

struct Ptrs {
  const float* b1;
  const float* b2;
  float* scores;
  __nv_bfloat16* H;
  __nv_bfloat16* Y;
  const uint32_t* removed_bits;
  uint32_t* active_bits;
  int16_t* idx;
  unsigned long long* hist;
  int* sync;                // [kSyncHeaderInts] header | done[] | ticket[] | ready[]
  int* split_counters;
  float* split_partial;
};

// ------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(uint64_t p, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}
// packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2): one issue slot for two lanes of math
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// exact-erf GELU of two values: x Phi(x) = max(x, 0) - |x| * 0.5 erfc(|x| / sqrt 2), 0.5 erfc(t) = 2^P7(t) for
// t = min(|x| / sqrt 2, 4.3) (weighted minimax fit, max abs error 4e-7 on [-3, 3]); the polynomial runs as seven
// packed FFMA2, the rest is one MUFU.EX2 and three scalar ops per value.
__device__ __forceinline__ uint64_t gelu2(float x0, float x1) {
  const float t0 = fminf(fabsf(x0) * 0.70710678118654752f, 4.3f);
  const float t1 = fminf(fabsf(x1) * 0.70710678118654752f, 4.3f);
  const uint64_t t = pk2(t0, t1);
  uint64_t q = pk2(1.0664232831913978e-04f, 1.0664232831913978e-04f);
  q = fma2(q, t, pk2(-5.025442806072533e-04f, -5.025442806072533e-04f));
  q = fma2(q, t, pk2(-2.20537674613297e-03f, -2.20537674613297e-03f));
  q = fma2(q, t, pk2(2.9348013922572136e-02f, 2.9348013922572136e-02f));
  q = fma2(q, t, pk2(-1.4891357719898224e-01f, -1.4891357719898224e-01f));
  q = fma2(q, t, pk2(-9.183364510536194e-01f, -9.183364510536194e-01f));
  q = fma2(q, t, pk2(-1.6279140710830688f, -1.6279140710830688f));
  q = fma2(q, t, pk2(-0.9999999403953552f, -0.9999999403953552f));
  float q0, q1;
  unpk2(q, q0, q1);
  const float e0 = tc::ex2_approx(q0), e1 = tc::ex2_approx(q1);
  return pk2(fmaf(-fabsf(x0), e0, fmaxf(x0, 0.f)), fmaf(-fabsf(x1), e1, fmaxf(x1, 0.f)));
}

template <int ACT>
__device__ __forceinline__ uint64_t activate2(float x0, float x1) {
  if constexpr (ACT == MOE_ACT_GELU)
    return gelu2(x0, x1);
  else
    return pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
}

// order-preserving float -> uint32 key (ascending)
__device__ __forceinline__ uint32_t float_key(float s) {
  const uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ------------------------------------------------------------------------------------------ work list
struct Item {
  int m_blk;               // this CTA's 128-row block
  int n;                   // column tile
  int kb_begin, kb_end;    // k-block range
  int slice;
};
// phase-1 item `it` of pair `p` (P pairs): tile index p + it * P, row-block-major
__device__ __forceinline__ bool item1(const Shape& g, int it, int p, int P, int rm, Item& t) {
  const int i = p + it * P;
  if (i >= g.items1) return false;
  const int mp = i / g.n_tiles1;
  t.n = i - mp * g.n_tiles1;
  t.m_blk = 2 * mp + rm;
  t.kb_begin = 0;
  t.kb_end = g.nkb1;
  t.slice = 0;
  return true;
}
// phase-3 items continue the round-robin where phase 1 stopped, so that the per-pair item counts stay balanced
__device__ __forceinline__ bool item3(const Shape& g, int it, int p, int P, int rm, Item& t) {
  int first = p - g.items1 % P;
  if (first < 0) first += P;
  const int j = first + it * P;
  if (j >= g.items3) return false;
  t.slice = j % g.split3;
  const int r = j / g.split3;
  const int mp = r / g.n_tiles3;
  t.n = r - mp * g.n_tiles3;
  t.m_blk = 2 * mp + rm;
  t.kb_begin = t.slice * g.kb_per_slice3;
  t.kb_end = min(g.nkb3, t.kb_begin + g.kb_per_slice3);
  return true;
}

// ------------------------------------------------------------------------------------------ routing
// One chunk = 512 / tpt consecutive tokens; `tpt` consecutive lanes own one token, each lane 16 experts.
// Exact k-th largest score by bisection on the order-preserving integer keys (MSB first, starting below the
// key prefix the whole warp shares, stopping when every token of the warp has separated exactly k keys);
// ties on the k-th key go to the lowest expert ids.  Outputs: expert-set words, ascending labels, histogram
// (shared-memory bins), and write-only masking of H (16-byte zero stores, a whole warp per token row).
__device__ __forceinline__ void route_chunk(const Shape& g, const Ptrs& a, int tok0, int tok_end, int ew, int lane,
                                            uint32_t* s_words, unsigned int* s_hist) {
  const unsigned full = 0xffffffffu;
  const int tpt = g.tpt;
  const int tpw = 32 >> g.tpt_log2;            // tokens per warp
  const int part = lane & (tpt - 1);
  const int tl = lane >> g.tpt_log2;
  const int t = tok0 + ew * tpw + tl;
  const bool t_ok = t < tok_end;   // tok_end <= T: end of the 128-row block (a chunk never leaves its block)
  const int e0 = part * kKeys;
  const int E = g.E;

  uint32_t key[kKeys];
  uint32_t valid = 0u;
  {
    uint32_t removed = 0u;
    if (a.removed_bits != nullptr && e0 < E) removed = (__ldg(a.removed_bits + (e0 >> 5)) >> (e0 & 31)) & 0xffffu;
    const float* row = a.scores + static_cast<size_t>(t_ok ? t : 0) * E + e0;
    if ((E & 3) == 0) {
#pragma unroll
      for (int i4 = 0; i4 < kKeys / 4; ++i4) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool in = t_ok && (e0 + 4 * i4) < E;
        if (in) {
          s = __ldcg(reinterpret_cast<const float4*>(row + 4 * i4));
          valid |= 0xfu << (4 * i4);
        }
        const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = 4 * i4 + j;
          const float v = ((removed >> i) & 1u) ? 0.f : sv[j];   // zeroed pattern row => score exactly 0
          key[i] = in ? float_key(v) : 0u;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < kKeys; ++i) {
        const bool in = t_ok && (e0 + i) < E;
        float v = in ? __ldcg(row + i) : 0.f;
        if ((removed >> i) & 1u) v = 0.f;
        key[i] = in ? float_key(v) : 0u;
        if (in) valid |= 1u << i;
      }
    }
  }

  uint32_t sel = 0u;   // 16-bit mask over this lane's experts
  if (g.k >= E) {
    sel = valid;
  } else if (g.k > 0) {
    // common key prefix of the warp's valid keys
    uint32_t k_or = 0u, k_and = full;
#pragma unroll
    for (int i = 0; i < kKeys; ++i)
      if ((valid >> i) & 1u) {
        k_or |= key[i];
        k_and &= key[i];
      }
    k_or = __reduce_or_sync(full, k_or);
    k_and = __reduce_and_sync(full, k_and);
    const uint32_t diff = k_or ^ k_and;
    uint32_t prefix = 0u;
    bool exact = !t_ok;   // idle token slots never hold the warp back
    if (diff != 0u) {
      const int top = 31 - __clz(diff);
      prefix = (top == 31) ? 0u : (k_and & ~((2u << top) - 1u));
      for (int bit = top; bit >= 0; --bit) {
        const uint32_t thr = prefix | (1u << bit);
        int c = 0;
#pragma unroll
        for (int i = 0; i < kKeys; ++i) c += (key[i] >= thr) ? 1 : 0;
        for (int o = 1; o < tpt; o <<= 1) c += __shfl_xor_sync(full, c, o);
        if (c >= g.k) prefix = thr;
        exact = exact || (c == g.k);
        if (__all_sync(full, exact)) break;
      }
    } else {
      prefix = k_and;   // every valid key of the warp is identical
    }
    // keys above the k-th value, then ties on it from the lowest expert id
    uint32_t gt = 0u, eq = 0u;
#pragma unroll
    for (int i = 0; i < kKeys; ++i) {
      const bool v = (valid >> i) & 1u;
      gt |= (v && key[i] > prefix) ? (1u << i) : 0u;
      eq |= (v && key[i] == prefix) ? (1u << i) : 0u;
    }
    int n_gt = __popc(gt);
    for (int o = 1; o < tpt; o <<= 1) n_gt += __shfl_xor_sync(full, n_gt, o);
    const int n_eq = __popc(eq);
    int incl = n_eq;
    for (int o = 1; o < tpt; o <<= 1) {
      const int v = __shfl_up_sync(full, incl, o, tpt);
      if (part >= o) incl += v;
    }
    int take = g.k - n_gt - (incl - n_eq);
    take = take < 0 ? 0 : (take > n_eq ? n_eq : take);
    uint32_t ties = 0u, w = eq;
    for (int j = 0; j < take; ++j) {
      const uint32_t low = w & (0u - w);
      ties |= low;
      w ^= low;
    }
    sel = gt | ties;
  }

  uint32_t removed16 = 0u;
  if (a.removed_bits != nullptr && e0 < E) removed16 = (__ldg(a.removed_bits + (e0 >> 5)) >> (e0 & 31)) & 0xffffu;
  const uint32_t active = sel & ~removed16;

  // expert-set words: lanes with an even part own a 32-bit word
  const uint32_t hi = __shfl_down_sync(full, active, 1);
  const uint32_t word = (tpt > 1) ? (active | (hi << 16)) : active;
  const int words_per_token = (E + 31) >> 5;
  if ((part & 1) == 0 && (part >> 1) < words_per_token) {
    s_words[tl * words_per_token + (part >> 1)] = word;
    if (a.active_bits != nullptr && t_ok) a.active_bits[static_cast<size_t>(t) * words_per_token + (part >> 1)] = word;
  }

  if (a.idx != nullptr) {
    const int n_sel = __popc(sel);
    int incl = n_sel;
    for (int o = 1; o < tpt; o <<= 1) {
      const int v = __shfl_up_sync(full, incl, o, tpt);
      if (part >= o) incl += v;
    }
    if (t_ok) {
      int16_t* out = a.idx + static_cast<size_t>(t) * g.k + (incl - n_sel);
      uint32_t w = sel;
      while (w) {
        const int i = __ffs(w) - 1;
        *out++ = static_cast<int16_t>(e0 + i);
        w &= w - 1;
      }
    }
  }

  if (a.hist != nullptr && t_ok && t >= g.count_begin && t < g.count_end) {
    uint32_t w = sel;
    while (w) {
      const int i = __ffs(w) - 1;
      atomicAdd(&s_hist[e0 + i], 1u);
      w &= w - 1;
    }
  }
  __syncwarp();

  if (g.mask_h && g.k < E) {
    // write-only masking: zero the neurons of every expert outside the token's active set (H is never read);
    // 16-byte units, one token row per warp trip, 4-neuron groups never straddle an expert (es % 4 == 0)
    const int units = g.h >> 3;
    for (int tt = 0; tt < tpw; ++tt) {
      const int tok = tok0 + ew * tpw + tt;
      if (tok >= tok_end) break;
      const uint32_t* wtok = s_words + tt * words_per_token;
      uint4* hrow = reinterpret_cast<uint4*>(a.H + static_cast<size_t>(tok) * g.h);
      for (int u = lane; u < units; u += 32) {
        const uint32_t ea = __umulhi(static_cast<uint32_t>(u) << 3, g.es_magic);
        const uint32_t eb = __umulhi((static_cast<uint32_t>(u) << 3) + 4u, g.es_magic);
        const bool on_a = (wtok[ea >> 5] >> (ea & 31u)) & 1u;
        const bool on_b = (wtok[eb >> 5] >> (eb & 31u)) & 1u;
        if (!on_a && !on_b)
          hrow[u] = make_uint4(0u, 0u, 0u, 0u);
        else if (!on_a)
          reinterpret_cast<uint2*>(hrow + u)[0] = make_uint2(0u, 0u);
        else if (!on_b)
          reinterpret_cast<uint2*>(hrow + u)[1] = make_uint2(0u, 0u);
      }
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------ the kernel
// per-warp bias slices staged in shared memory: lanes load coalesced, everyone re-reads float4 broadcasts
__device__ __forceinline__ void stage_bias(float* sb, const float* b0, const float* b1, int n, int lane) {
  __syncwarp();
  for (int i = lane; i < 64; i += 32) {
    sb[i] = (b0 != nullptr && i < n) ? __ldg(b0 + i) : 0.f;
    sb[64 + i] = (b1 != nullptr && i < n) ? __ldg(b1 + i) : 0.f;
  }
  __syncwarp();
}

template <int kWords>
__device__ __forceinline__ void store_words(void* dst, const uint32_t* w) {
  if constexpr (kWords % 4 == 0) {
#pragma unroll
    for (int i = 0; i < kWords / 4; ++i)
      reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < kWords / 2; ++i) reinterpret_cast<uint2*>(dst)[i] = make_uint2(w[2 * i], w[2 * i + 1]);
  }
}

template <int CH, int ACT>
__device__ __forceinline__ void geglu_group(const Shape& g, uint32_t taddr, const float* sbias, __nv_bfloat16* hrow,
                                            float* spart_row, int col0, int cpg, int cg) {
  uint64_t score2 = pk2(0.f, 0.f);
  int chunk_in_expert = 0, e_slot = (g.chunks_per_expert > 0) ? cg * (cpg / g.es) : cg;
  for (int c = 0; c < cpg; c += CH) {
    uint32_t v[CH], gt[CH];
    tc::tmem_ld_cols<CH>(taddr + col0 + c, v);
    tc::tmem_ld_cols<CH>(taddr + g.nv + col0 + c, gt);
    tc::tmem_ld_wait();
    uint32_t hw[CH / 2];
#pragma unroll
    for (int i = 0; i < CH; i += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(sbias + c + i);
      const float4 bg = *reinterpret_cast<const float4*>(sbias + 64 + c + i);
      float ga, gb, gc, gd;
      unpk2(add2(pk2(__uint_as_float(gt[i]), __uint_as_float(gt[i + 1])), pk2(bg.x, bg.y)), ga, gb);
      unpk2(add2(pk2(__uint_as_float(gt[i + 2]), __uint_as_float(gt[i + 3])), pk2(bg.z, bg.w)), gc, gd);
      const uint64_t a01 = activate2<ACT>(ga, gb);
      const uint64_t a23 = activate2<ACT>(gc, gd);
      score2 = add2(score2, add2(a01, a23));
      const uint64_t v01 = add2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), pk2(bv.x, bv.y));
      const uint64_t v23 = add2(pk2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), pk2(bv.z, bv.w));
      float h0, h1, h2, h3;
      unpk2(mul2(v01, a01), h0, h1);
      unpk2(mul2(v23, a23), h2, h3);
      hw[i / 2] = pack_bf16x2(h0, h1);
      hw[i / 2 + 1] = pack_bf16x2(h2, h3);
    }
    store_words<CH / 2>(hrow + col0 + c, hw);
    if (g.chunks_per_expert > 0 && ++chunk_in_expert == g.chunks_per_expert) {
      float s0, s1;
      unpk2(score2, s0, s1);
      spart_row[e_slot++] = s0 + s1;
      score2 = pk2(0.f, 0.f);
      chunk_in_expert = 0;
    }
  }
  if (g.chunks_per_expert == 0) {
    float s0, s1;
    unpk2(score2, s0, s1);
    spart_row[cg] = s0 + s1;
  }
}

template <int CH>
__global__ void __launch_bounds__(kNumThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_hs, const __grid_constant__ CUtensorMap tmap_hl,
                 const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_y, const Shape g,
                 const Ptrs a) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TRACE(0);
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* hstage = smem + g.stages * g.slot_bytes;
  float* sbias_all = reinterpret_cast<float*>(hstage + 2 * g.hs_bytes);
  float* spart = sbias_all + kEpiWarps * (kBiasBytesPerWarp / 4);             // [2][128][kSpartPerRow]
  uint32_t* s_words_all = reinterpret_cast<uint32_t*>(spart + 2 * kBlockM * kSpartPerRow);
  unsigned int* s_hist = s_words_all + kEpiWarps * kRouteWordsPerWarp;        // [kMaxExperts]
  Barriers* bars = reinterpret_cast<Barriers*>(s_hist + kMaxExperts);

  const int rm = static_cast<int>(tc::cluster_ctarank());
  const int p = static_cast<int>(blockIdx.x) >> 1, P = static_cast<int>(gridDim.x) >> 1;
  int* const ws_done = a.sync + kSyncHeaderInts;
  int* const ws_ticket = ws_done + kMaxBlocks;
  int* const ws_ready = ws_ticket + kMaxBlocks;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tensormap(&tmap_x);
    tc::prefetch_tensormap(&tmap_w1);
    tc::prefetch_tensormap(&tmap_hl);
    tc::prefetch_tensormap(&tmap_w2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < g.stages; ++i) {
      tc::mbar_init(&bars->full[i], 1);
      tc::mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars->tmem_full[i], 1);
      tc::mbar_init(&bars->tmem_empty[i], 2 * kEpiWarps);   // both CTAs' epilogues release the leader's MMA thread
      tc::mbar_init(&bars->hs_full[i], kEpiWarps);
      tc::mbar_init(&bars->hs_empty[i], 1);
    }
    tc::mbar_init(&bars->route_req, 1);
    tc::mbar_init(&bars->route_done, kEpiWarps);
    tc::mbar_init(&bars->fin, 1);
    tc::fence_mbar_init();
  }
  if (warp == 2) {
    tc::tmem_alloc_2sm<kTmemCols>(&bars->tmem_base);
    if (lane == 0) {
      tc::prefetch_tensormap(&tmap_hs);
      tc::prefetch_tensormap(&tmap_y);
    }
  }
  for (int i = threadIdx.x; i < kMaxExperts; i += kNumThreads) s_hist[i] = 0u;
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::cluster_sync_all();
  tc::fence_after_thread_sync();
  if (threadIdx.x == 0) TRACE(1);
  pdl_wait();   // everything above overlapped the previous kernel's tail
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (threadIdx.x == 0) TRACE(2);
  const uint32_t tmem_base = bars->tmem_base;

  const uint32_t bytes1 = static_cast<uint32_t>(g.ks1 * (kABytes + g.nv * 128));
  const uint32_t bytes3 = static_cast<uint32_t>(g.ks3 * (kABytes + (g.bn / 2) * 128));

  if (warp == 0 || warp == 3) {
    // ================================================================== TMA producers (A: warp 0, B: warp 3)
    const bool do_a = warp == 0;
    int s = 0;
    uint32_t ph = 0;
    Item t;
    const uint32_t full_leader0 = tc::mapa_u32(&bars->full[0], 0);
    for (int phase = 0; phase < 2; ++phase) {
      const int ks = phase == 0 ? g.ks1 : g.ks3;
      for (int it = 0; phase == 0 ? item1(g, it, p, P, rm, t) : item3(g, it, p, P, rm, t); ++it) {
        if (phase == 1 && do_a) {
          // the block's H rows are complete and masked once every routing chunk of the block has been counted
          if (lane == 0) {
            const int* flag = ws_ready + t.m_blk;
            while (ld_acquire(flag) < g.chunks_per_block) {
            }
          }
          __syncwarp();
          fence_proxy_async_all();   // generic-proxy writes of other SMs (acquired above) -> this thread's TMA reads
        }
        for (int kb = t.kb_begin; kb < t.kb_end; kb += ks) {
          tc::mbar_wait(&bars->empty[s], ph ^ 1u);
          uint8_t* sa = smem + s * g.slot_bytes;
          uint8_t* sb = sa + ks * kABytes;
          if (tc::elect_one()) {
            const uint32_t full_leader = full_leader0 + static_cast<uint32_t>(s) * 8u;
            if (do_a) {
              if (rm == 0) tc::mbar_arrive_expect_tx(&bars->full[s], 2u * (phase == 0 ? bytes1 : bytes3));
              tc::tma_load_3d_2sm(sa, phase == 0 ? &tmap_x : &tmap_hl, full_leader, 0, t.m_blk * kBlockM, kb);
#if MOE_TRACE
              if (phase == 0 && it == 0 && kb == 0) TRACE(7);
#endif
            } else if (phase == 0) {
              // CTA 0 of the pair stages the value rows, CTA 1 the gate rows of the tile's neurons
              tc::tma_load_3d_2sm(sb, &tmap_w1, full_leader, 0, (rm == 0 ? 0 : g.h) + t.n * g.nv, kb);
            } else {
              tc::tma_load_3d_2sm(sb, &tmap_w2, full_leader, 0, t.n * g.bn + rm * (g.bn / 2), kb);
            }
          }
          __syncwarp();
          if (++s == g.stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer (leader CTA)
    if (rm == 0) {
      int s = 0;
      uint32_t ph = 0;
      int acc_it = 0;
      Item t;
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      for (int phase = 0; phase < 2; ++phase) {
        const int ks = phase == 0 ? g.ks1 : g.ks3;
        const uint32_t idesc = tc::umma_idesc_bf16_f32(2 * kBlockM, static_cast<uint32_t>(phase == 0 ? 2 * g.nv : g.bn));
        const uint32_t b_sub_bytes = static_cast<uint32_t>((phase == 0 ? g.nv : g.bn / 2) * 128);
        for (int it = 0; phase == 0 ? item1(g, it, p, P, rm, t) : item3(g, it, p, P, rm, t); ++it, ++acc_it) {
          const int as = acc_it & 1;
          tc::mbar_wait(&bars->tmem_empty[as], ((acc_it >> 1) & 1u) ^ 1u);
          tc::fence_after_thread_sync();
          const uint32_t d_tmem = tb + as * kAccStride;
          for (int kb = t.kb_begin; kb < t.kb_end; kb += ks) {
            tc::mbar_wait(&bars->full[s], ph);
            tc::fence_after_thread_sync();
            const uint32_t a_base = tc::smem_u32(smem + s * g.slot_bytes);
            const uint32_t b_base = a_base + ks * kABytes;
            const int n_sub = min(ks, t.kb_end - kb);
            const bool leader_lane = tc::elect_one();
            if (leader_lane) {
#if MOE_TRACE
              if (acc_it == 0 && kb == t.kb_begin) TRACE(3);
#endif
              for (int sub = 0; sub < n_sub; ++sub) {
                const uint32_t a_addr = a_base + sub * kABytes;
                const uint32_t b_addr = b_base + sub * b_sub_bytes;
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                  const uint64_t da = tc::umma_desc_kmajor_sw128(a_addr + k * kUmmaK * 2);
                  const uint64_t db = tc::umma_desc_kmajor_sw128(b_addr + k * kUmmaK * 2);
                  tc::umma_bf16_ss_2sm(d_tmem, da, db, idesc, (kb > t.kb_begin || sub != 0 || k != 0) ? 1u : 0u);
                }
              }
              tc::umma_commit_2sm_mc(&bars->empty[s], 0x3);   // frees the slot in both CTAs of the pair
            }
            __syncwarp();
            if (++s == g.stages) {
              s = 0;
              ph ^= 1u;
            }
          }
          if (tc::elect_one()) {
            tc::umma_commit_2sm_mc(&bars->tmem_full[as], 0x3);   // accumulator complete -> both epilogues
#if MOE_TRACE
            if (acc_it < 13) TRACE(8 + 4 * acc_it);
#endif
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 2) {
    // ================================================================== store / sync warp
    if (lane == 0) {
      int use[2] = {0, 0};        // completed uses of each staging buffer
      int st_it = 0;              // staged phase-1 tiles so far
      uint32_t r_posted = 0;
      auto claim_and_route = [&](int m_blk) {
        for (;;) {
          const int c = atomicAdd(ws_ticket + m_blk, 1);
          if (c >= g.chunks_per_block) break;
          bars->req_block = m_blk;
          bars->req_chunk = c;
          tc::mbar_arrive(&bars->route_req);
          tc::mbar_wait(&bars->route_done, r_posted & 1u);
          ++r_posted;
          __threadfence();
          atomicAdd(ws_ready + m_blk, 1);
        }
      };
      Item t;
      for (int it = 0; item1(g, it, p, P, rm, t); ++it, ++st_it) {
        const int buf = st_it & 1;
        tc::mbar_wait(&bars->hs_full[buf], use[buf] & 1u);
        tc::tma_store_2d(&tmap_hs, hstage + buf * g.hs_bytes, t.n * g.nv, t.m_blk * kBlockM);   // rows beyond T are clipped
        tc::tma_store_commit();
        tc::tma_store_wait<0>();          // H tile globally written (not only read out of smem)
        tc::mbar_arrive(&bars->hs_empty[buf]);
        ++use[buf];
        fence_proxy_async_all();
        __threadfence();                  // H tile + the tile's scores (ordered by hs_full) before the count
        const int prev = atomicAdd(ws_done + t.m_blk, 1);
        if (prev == g.n_tiles1 - 1) {
          __threadfence();                // acquire: the other CTAs' tiles of this block happen-before the routing
          claim_and_route(t.m_blk);
        }
      }
      for (int it = 0; item3(g, it, p, P, rm, t); ++it) {
        // make sure the block gets routed even if the CTA that completed it is busy: wait for its last
        // phase-1 tile, then take whatever chunks are still unclaimed
        {
          const int* flag = ws_done + t.m_blk;
          while (ld_acquire(flag) < g.n_tiles1) {
          }
        }
        claim_and_route(t.m_blk);
        if (g.split3 == 1) {
          tc::mbar_wait(&bars->hs_full[0], use[0] & 1u);
          tc::tma_store_2d(&tmap_y, hstage, t.n * g.bn, t.m_blk * kBlockM);   // clipped at T rows / d columns
          tc::tma_store_commit();
          tc::tma_store_wait_read<0>();
          tc::mbar_arrive(&bars->hs_empty[0]);
          ++use[0];
        }
      }
      tc::tma_store_wait<0>();
      tc::mbar_arrive(&bars->fin);
    }
    __syncwarp();
  } else {
    // ================================================================== epilogue warps
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int cg = ew >> 2;            // column group 0..3
    float* sbias = sbias_all + ew * (kBiasBytesPerWarp / 4);
    uint32_t* s_words = s_words_all + ew * kRouteWordsPerWarp;
    const int q_row = 32 * q + lane;
    uint32_t seen = 0;                 // routing requests served

    auto serve = [&]() {
      const int m_blk = bars->req_block, c = bars->req_chunk;
      route_chunk(g, a, m_blk * kBlockM + c * (kEpiThreads >> g.tpt_log2), min(g.T, (m_blk + 1) * kBlockM), ew, lane, s_words,
                  s_hist);
      ++seen;
      __threadfence();     // zero-writes / labels visible device-wide before the sync warp counts the chunk
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->route_done);
    };
    // wait on one of this CTA's barriers, serving routing requests meanwhile (a request posted by our own sync
    // warp may be what the awaited event transitively depends on)
    auto wait_or_serve = [&](uint64_t* bar, uint32_t parity) {
      for (;;) {
        if (__all_sync(0xffffffffu, tc::mbar_try_wait(bar, parity))) break;
        if (__all_sync(0xffffffffu, tc::mbar_try_wait(&bars->route_req, seen & 1u))) serve();
      }
    };

    int acc_it = 0;
    int use[2] = {0, 0};
    Item t;
    // ---------------------------------------------------------------- phase 1
    {
      const int cpg = g.nv / 4;
      const int col0 = cg * cpg;
      for (int it = 0; item1(g, it, p, P, rm, t); ++it, ++acc_it) {
        const int as = acc_it & 1;
        const int buf = it & 1;
        const int n_tile0 = t.n * g.nv;
        stage_bias(sbias, a.b1 != nullptr ? a.b1 + n_tile0 + col0 : nullptr,
                   a.b1 != nullptr ? a.b1 + g.h + n_tile0 + col0 : nullptr, cpg, lane);
        if (use[buf] > 0) wait_or_serve(&bars->hs_empty[buf], (use[buf] - 1) & 1u);   // staging buffer drained
        wait_or_serve(&bars->tmem_full[as], (acc_it >> 1) & 1u);
        tc::fence_after_thread_sync();
#if MOE_TRACE
        if (ew == 0 && lane == 0 && acc_it < 13) TRACE(8 + 4 * acc_it + 1);
#endif
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + as * kAccStride;
        __nv_bfloat16* hrow = reinterpret_cast<__nv_bfloat16*>(hstage + buf * g.hs_bytes) + q_row * g.nv;
        float* spart_row = spart + (buf * kBlockM + q_row) * kSpartPerRow;
        if (g.act == MOE_ACT_GELU)
          geglu_group<CH, MOE_ACT_GELU>(g, taddr, sbias, hrow, spart_row, col0, cpg, cg);
        else
          geglu_group<CH, MOE_ACT_RELU>(g, taddr, sbias, hrow, spart_row, col0, cpg, cg);
        // accumulator stage drained -> the leader's MMA thread may overwrite it
        tc::fence_before_thread_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_cluster(tc::mapa_u32(&bars->tmem_empty[as], 0));
        // expert scores of this row: the 4 column-group warps of the lane quarter meet, the first one writes
        tc::named_bar_sync(2 + q, 4 * 32);
        if (cg == 0) {
          const int row = t.m_blk * kBlockM + q_row;
          if (row < g.T) {
            float* dst = a.scores + static_cast<size_t>(row) * g.E + t.n * g.experts_per_tile;
            if (g.experts_per_tile == 4 && g.span == 1 && (g.E & 3) == 0) {
              *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(spart_row);
            } else {
              for (int e = 0; e < g.experts_per_tile; ++e) {
                float tot = 0.f;
                for (int j = 0; j < g.span; ++j) tot += spart_row[e * g.span + j];
                dst[e] = tot;
              }
            }
          }
        }
        // H tile: generic-proxy smem writes -> async proxy; the sync warp stores it and publishes the tile
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->hs_full[buf]);
        ++use[buf];
#if MOE_TRACE
        if (ew == 0 && lane == 0 && acc_it < 13) TRACE(8 + 4 * acc_it + 2);
#endif
      }
    }
    // ---------------------------------------------------------------- phase 3
    {
      const int cpg = g.bn / 4;
      const int col0 = cg * cpg;
      for (int it = 0; item3(g, it, p, P, rm, t); ++it, ++acc_it) {
        const int as = acc_it & 1;
        const int n0 = t.n * g.bn + col0;                  // first output column of this warp's group
        const int nvalid = max(0, min(cpg, g.d - n0));     // the last tile may overhang d
        stage_bias(sbias, a.b2 != nullptr ? a.b2 + n0 : nullptr, nullptr, nvalid, lane);
        if (g.split3 == 1) {
          // the whole staging area (both phase-1 buffers) holds one Y tile
          if (use[0] > 0) wait_or_serve(&bars->hs_empty[0], (use[0] - 1) & 1u);
          if (use[1] > 0) wait_or_serve(&bars->hs_empty[1], (use[1] - 1) & 1u);
        }
        wait_or_serve(&bars->tmem_full[as], (acc_it >> 1) & 1u);
        tc::fence_after_thread_sync();
#if MOE_TRACE
        if (ew == 0 && lane == 0 && acc_it < 13) TRACE(8 + 4 * acc_it + 1);
#endif
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + as * kAccStride + col0;
        const int row = t.m_blk * kBlockM + q_row;
        const bool row_ok = row < g.T;
        if (g.split3 == 1) {
          __nv_bfloat16* yrow = reinterpret_cast<__nv_bfloat16*>(hstage) + q_row * g.bn + col0;
          for (int c = 0; c < cpg; c += CH) {
            uint32_t acc[CH];
            tc::tmem_ld_cols<CH>(taddr + c, acc);
            tc::tmem_ld_wait();
            uint32_t yw[CH / 2];
#pragma unroll
            for (int i = 0; i < CH; i += 4) {
              const float4 b = *reinterpret_cast<const float4*>(sbias + c + i);
              yw[i / 2] = pack_bf16x2(__uint_as_float(acc[i]) + b.x, __uint_as_float(acc[i + 1]) + b.y);
              yw[i / 2 + 1] = pack_bf16x2(__uint_as_float(acc[i + 2]) + b.z, __uint_as_float(acc[i + 3]) + b.w);
            }
            store_words<CH / 2>(yrow + c, yw);
          }
          tc::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_cluster(tc::mapa_u32(&bars->tmem_empty[as], 0));
          tc::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bars->hs_full[0]);
          ++use[0];
        } else {
          // ---- split-K: park this slice's fp32 partial tile; the last slice to arrive reduces in slice order
          const size_t plane = static_cast<size_t>(g.T) * g.d;
          float* part = a.split_partial + t.slice * plane + static_cast<size_t>(row) * g.d + n0;
          for (int c = 0; c < cpg; c += CH) {
            uint32_t acc[CH];
            tc::tmem_ld_cols<CH>(taddr + c, acc);
            tc::tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int i = 0; i < CH; i += 4)
                if (c + i < nvalid)
                  __stcg(reinterpret_cast<float4*>(part + c + i),
                         make_float4(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1]), __uint_as_float(acc[i + 2]),
                                     __uint_as_float(acc[i + 3])));
            }
          }
          tc::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_cluster(tc::mapa_u32(&bars->tmem_empty[as], 0));
          __threadfence();
          tc::named_bar_sync(1, kEpiThreads);
          if (ew == 0 && lane == 0) {
            int* counter = a.split_counters + t.m_blk * g.n_tiles3 + t.n;
            const int prev = atomicAdd(counter, 1);
            const int last = prev == g.split3 - 1;
            if (last) *counter = 0;
            bars->last_cta = last;
          }
          tc::named_bar_sync(1, kEpiThreads);
          if (bars->last_cta) {
            __threadfence();
            const int et = ew * 32 + lane;
            const int cols4 = g.bn >> 2;
            const int tile_col0 = t.n * g.bn;
            const int tile_row0 = t.m_blk * kBlockM;
            for (int e = et; e < kBlockM * cols4; e += kEpiThreads) {
              const int r = e / cols4, c4 = (e - r * cols4) << 2;
              const int grow = tile_row0 + r, gcol = tile_col0 + c4;
              if (grow >= g.T || gcol >= g.d) continue;
              const float* src = a.split_partial + static_cast<size_t>(grow) * g.d + gcol;
              float4 pz[8];
#pragma unroll
              for (int sl = 0; sl < 8; ++sl)
                if (sl < g.split3) pz[sl] = __ldcg(reinterpret_cast<const float4*>(src + sl * plane));
              float4 sum = a.b2 != nullptr ? __ldg(reinterpret_cast<const float4*>(a.b2 + gcol)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int sl = 0; sl < 8; ++sl)
                if (sl < g.split3) {
                  sum.x += pz[sl].x;
                  sum.y += pz[sl].y;
                  sum.z += pz[sl].z;
                  sum.w += pz[sl].w;
                }
              *reinterpret_cast<uint2*>(a.Y + static_cast<size_t>(grow) * g.d + gcol) =
                  make_uint2(pack_bf16x2(sum.x, sum.y), pack_bf16x2(sum.z, sum.w));
            }
          }
          tc::named_bar_sync(1, kEpiThreads);
        }
#if MOE_TRACE
        if (ew == 0 && lane == 0 && acc_it < 13) TRACE(8 + 4 * acc_it + 2);
#endif
      }
    }
    // stay available for routing requests until the sync warp has nothing more to post
    wait_or_serve(&bars->fin, 0u);
#if MOE_TRACE
    if (ew == 0 && lane == 0) TRACE(4);
#endif
  }

  // ================================================================== teardown
  tc::fence_before_thread_sync();
  __syncthreads();
  if (threadIdx.x == 0) TRACE(5);
  if (a.hist != nullptr) {
    for (int i = threadIdx.x; i < g.E; i += kNumThreads)
      if (s_hist[i]) atomicAdd(a.hist + i, static_cast<unsigned long long>(s_hist[i]));
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const int prev = atomicAdd(a.sync, 1);
    bars->last_cta = (prev == static_cast<int>(gridDim.x) - 1);
  }
  tc::cluster_sync_all();   // nobody exits while the peer may still signal its smem (also a CTA barrier)
  tc::fence_after_thread_sync();
  if (bars->last_cta) {
    // every other CTA is past its last access: leave the sync arrays clean for the next launch
    const int n_blocks = 2 * g.m_pairs;
    for (int i = threadIdx.x; i < n_blocks; i += kNumThreads) {
      ws_done[i] = 0;
      ws_ticket[i] = 0;
      ws_ready[i] = 0;
    }
    if (threadIdx.x == 0) a.sync[0] = 0;
  }
  if (warp == 2) tc::tmem_dealloc_2sm<kTmemCols>(tmem_base);
  if (threadIdx.x == 0) TRACE(6);
}

// ------------------------------------------------------------------------------------------ host side
static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

static int ensure_smem(const void* kfn) {
  static const void* configured[16];
  static int n_configured = 0;
  for (int i = 0; i < n_configured; ++i)
    if (configured[i] == kfn) return MOE_OK;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e != cudaSuccess) return fail(MOE_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", kSmemLimit, cudaGetErrorString(e));
  if (n_configured < 16) configured[n_configured++] = kfn;
  return MOE_OK;
}

}  // namespace fused
}  // namespace moe

extern "C" {

size_t moe_ffn_fused_workspace_bytes(int T, int d, int h) {
  using namespace moe::fused;
  (void)h;
  size_t want = static_cast<size_t>(8) * static_cast<size_t>(T > 0 ? T : 0) * static_cast<size_t>(d > 0 ? d : 0) * 4;
  const size_t cap = static_cast<size_t>(64) << 20;
  if (want > cap) want = cap;
  return kSyncBytes + kSplitCounterBytes + want;
}

int moe_debug_trace_fused(unsigned long long* host_out, int n) {
  using namespace moe;
  MOE_REQUIRE(host_out != nullptr && n >= 0 && n <= 256 * 64, MOE_ERR_INVALID_ARGUMENT, "moe_debug_trace_fused: bad args");
#if MOE_TRACE
  cudaError_t e = cudaMemcpyFromSymbol(host_out, fused::g_trace, sizeof(unsigned long long) * n);
  if (e != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_debug_trace_fused: %s", cudaGetErrorString(e));
  void* sym = nullptr;
  if (cudaGetSymbolAddress(&sym, fused::g_trace) == cudaSuccess) cudaMemset(sym, 0, sizeof(unsigned long long) * 256 * 64);
  return MOE_OK;
#else
  return fail(MOE_ERR_UNSUPPORTED_SHAPE, "moe_debug_trace_fused: library built without MOE_TRACE");
#endif
}

int moe_ffn_fused(const void* x, const void* w1p, const float* b1p, const void* w2p, const float* b2, void* H,
                  float* scores, void* Y, const uint32_t* removed_bits, int k, uint32_t* active_bits, int16_t* idx,
                  unsigned long long* hist, int count_begin, int count_end, int T, int d, int h, int E, int es, int act,
                  int mask_h, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace moe;
  using namespace moe::fused;
  MOE_REQUIRE(x && w1p && w2p && H && scores && Y && workspace, MOE_ERR_INVALID_ARGUMENT,
              "moe_ffn_fused: NULL x / w1p / w2p / H / scores / Y / workspace");
  MOE_REQUIRE(T >= 0 && d >= 8 && h >= 8 && E >= 1 && es >= 1 && k >= 0 && k <= E, MOE_ERR_INVALID_ARGUMENT,
              "moe_ffn_fused: bad sizes T=%d d=%d h=%d E=%d es=%d k=%d", T, d, h, E, es, k);
  MOE_REQUIRE(static_cast<long long>(E) * es == h, MOE_ERR_INVALID_ARGUMENT, "moe_ffn_fused: E*es=%d*%d != h=%d", E, es, h);
  MOE_REQUIRE(act == MOE_ACT_GELU || act == MOE_ACT_RELU, MOE_ERR_INVALID_ARGUMENT, "moe_ffn_fused: act=%d", act);
  MOE_REQUIRE(d % 64 == 0 && h % 64 == 0, MOE_ERR_UNSUPPORTED_SHAPE,
              "moe_ffn_fused: d=%d and h=%d must be multiples of 64 (use the unfused kernels)", d, h);
  MOE_REQUIRE(es % 4 == 0 && E <= kMaxExperts && h < 65536, MOE_ERR_UNSUPPORTED_SHAPE,
              "moe_ffn_fused: needs es %% 4 == 0, E <= %d and h < 65536 (es=%d E=%d h=%d)", kMaxExperts, es, E, h);
  MOE_REQUIRE(((reinterpret_cast<uintptr_t>(H) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(scores) |
                reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
              MOE_ERR_INVALID_ARGUMENT, "moe_ffn_fused: H / Y / scores / workspace must be 16-byte aligned");
  MOE_REQUIRE(workspace_bytes >= kSyncBytes + kSplitCounterBytes, MOE_ERR_INVALID_ARGUMENT,
              "moe_ffn_fused: workspace of %zu bytes is smaller than moe_ffn_fused_workspace_bytes()", workspace_bytes);
  if (T == 0) return MOE_OK;

  Shape g = {};
  g.T = T;
  g.d = d;
  g.h = h;
  g.E = E;
  g.es = es;
  g.k = k;
  const int m_tiles = (T + kBlockM - 1) / kBlockM;
  g.m_pairs = (m_tiles + 1) / 2;
  MOE_REQUIRE(2 * g.m_pairs <= kMaxBlocks, MOE_ERR_UNSUPPORTED_SHAPE, "moe_ffn_fused: T=%d exceeds %d row blocks", T, kMaxBlocks);

  // ---- phase 1 tile: nv neuron pairs (UMMA N = 2 nv <= 256), 4 column groups of cpg = nv / 4 that hold whole
  // experts or a quarter / half of one
  int nv = 0;
  for (int cand = 128; cand >= 16; cand -= 8) {
    if (h % cand) continue;
    const int cpg = cand / 4;
    if (cpg % 4 || cpg > 64) continue;
    const bool whole = cpg % es == 0;
    const bool spans = es % cpg == 0 && (es / cpg == 2 || es / cpg == 4) && cand % es == 0;
    if (!(whole || spans)) continue;
    const int ept = cand / es, span = whole ? 1 : es / cpg;
    if (ept * span > kSpartPerRow) continue;
    nv = cand;
    break;
  }
  MOE_REQUIRE(nv > 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_ffn_fused: expert size %d unsupported for h=%d", es, h);
  const int cpg1 = nv / 4;
  int gcd_v = cpg1, tmp = es;
  while (tmp) {
    const int r = gcd_v % tmp;
    gcd_v = tmp;
    tmp = r;
  }
  const int ch1 = (gcd_v % 32 == 0) ? 32 : (gcd_v % 20 == 0) ? 20 : (gcd_v % 16 == 0) ? 16 : (gcd_v % 8 == 0) ? 8 : 4;
  g.nv = nv;
  g.n_tiles1 = h / nv;
  g.nkb1 = d / kBlockK;
  g.items1 = g.m_pairs * g.n_tiles1;
  g.experts_per_tile = nv / es;
  g.chunks_per_expert = (cpg1 % es == 0) ? es / ch1 : 0;
  g.span = (cpg1 % es == 0) ? 1 : es / cpg1;

  const int sms = sm_count();
  const int P = sms / 2;
  MOE_REQUIRE(P >= 1, MOE_ERR_CUDA, "moe_ffn_fused: needs at least 2 SMs");

  // ---- phase 3 tile width bn (divides into 4 column groups that are multiples of the epilogue chunk, fits the
  // staging area 2 x 128 x nv bf16, each CTA of the pair stages bn / 2 weight rows) and split-K factor
  g.nkb3 = h / kBlockK;
  size_t max_split_ws = 1;
  {
    const size_t per_slice = static_cast<size_t>(T) * d * 4;
    const size_t avail = workspace_bytes - kSyncBytes - kSplitCounterBytes;
    max_split_ws = per_slice ? avail / per_slice : 1;
    if (max_split_ws > 8) max_split_ws = 8;
    if (max_split_ws < 1) max_split_ws = 1;
    if (d % 4) max_split_ws = 1;
  }
  if (const char* e = getenv("MOE_FUSED_SPLIT")) {
    const int cap = atoi(e);
    if (cap >= 1 && static_cast<size_t>(cap) < max_split_ws) max_split_ws = cap;
  }
  int best_bn = 0, best_split = 1;
  double best_cost = 1e300;
  for (int bn = 256; bn >= 32; bn -= 16) {
    if (bn > 2 * nv) continue;                            // one Y tile must fit the staging area
    if ((bn / 2) % 8 || (bn / 4) % 4) continue;
    const int ch3 = ((bn / 4) % 32 == 0) ? 32 : ((bn / 4) % 20 == 0) ? 20 : ((bn / 4) % 16 == 0) ? 16 : ((bn / 4) % 8 == 0) ? 8 : 4;
    if (ch3 != ch1) continue;                             // one epilogue chunk width per kernel instantiation
    // (the last tile may overhang d: TMA zero-fills the missing weight rows and clips the Y store)
    const int n_tiles_c = (d + bn - 1) / bn;
    if (static_cast<size_t>(2 * g.m_pairs) * n_tiles_c * 4 > kSplitCounterBytes) continue;
    for (int sp = 1; sp <= static_cast<int>(max_split_ws); ++sp) {
      if (sp > 1 && g.nkb3 / sp < 8) break;
      int kb_per = (g.nkb3 + sp - 1) / sp;
      kb_per = (kb_per + 1) / 2 * 2;
      if (sp > 1 && (sp - 1) * kb_per >= g.nkb3) continue;
      const long long items = static_cast<long long>(g.m_pairs) * n_tiles_c * sp;
      const long long rounds = (items + P - 1) / P;
      const double per_kblock = 115.0 + (2.0 * bn > 220.0 ? 2.0 * bn : 220.0);
      const double cost = rounds * (kb_per * per_kblock + 2500.0) + (sp > 1 ? 6000.0 + 40.0 * sp * bn : 0.0);
      if (cost < best_cost * 0.999) {
        best_cost = cost;
        best_bn = bn;
        best_split = sp;
      }
    }
  }
  if (const char* e = getenv("MOE_FUSED_BN")) {
    const int bn = atoi(e);
    if (bn >= 32 && bn <= 2 * nv && bn % 16 == 0 && (bn / 2) % 8 == 0) {
      best_bn = bn;
      best_split = static_cast<int>(max_split_ws);
    }
  }
  MOE_REQUIRE(best_bn > 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_ffn_fused: no down-projection tile width for d=%d (nv=%d)", d, nv);
  g.bn = best_bn;
  g.n_tiles3 = (d + best_bn - 1) / best_bn;
  g.split3 = best_split;

  // ---- pipeline: k-blocks per stage (2 if at least 3 stages fit), ring slot = the larger of the two phases
  g.hs_bytes = kBlockM * nv * 2;
  const int fixed = 1024 + 2 * g.hs_bytes + kEpiWarps * kBiasBytesPerWarp + 2 * kBlockM * kSpartPerRow * 4 +
                    kEpiWarps * kRouteWordsPerWarp * 4 + kMaxExperts * 4 + static_cast<int>(sizeof(Barriers)) + 64;
  int ks = 2;
  if (const char* e = getenv("MOE_FUSED_KS")) ks = atoi(e) == 1 ? 1 : 2;
  for (; ks >= 1; --ks) {
    const int slot = ks * (kABytes + (nv > best_bn / 2 ? nv : best_bn / 2) * 128);
    const int stages = (kSmemLimit - fixed) / slot;
    if (stages >= 3 || ks == 1) {
      g.slot_bytes = slot;
      g.stages = stages > kMaxStages ? kMaxStages : stages;
      break;
    }
  }
  MOE_REQUIRE(g.stages >= 2, MOE_ERR_UNSUPPORTED_SHAPE, "moe_ffn_fused: tiles do not fit shared memory");
  g.ks1 = g.nkb1 >= 2 ? ks : 1;
  g.ks3 = ks;
  g.kb_per_slice3 = (g.nkb3 + g.split3 - 1) / g.split3;
  g.kb_per_slice3 = (g.kb_per_slice3 + g.ks3 - 1) / g.ks3 * g.ks3;
  while (g.split3 > 1 && (g.split3 - 1) * g.kb_per_slice3 >= g.nkb3) --g.split3;
  g.items3 = g.m_pairs * g.n_tiles3 * g.split3;

  // ---- routing geometry: 16 experts per thread, tpt (a power of two) threads per token
  int tpt = 1;
  while (tpt * kKeys < E) tpt <<= 1;
  MOE_REQUIRE(tpt <= 32, MOE_ERR_UNSUPPORTED_SHAPE, "moe_ffn_fused: E=%d too large", E);
  g.tpt = tpt;
  g.tpt_log2 = ilog2(tpt);
  g.chunks_per_block = kBlockM / (kEpiThreads / tpt) > 0 ? kBlockM / (kEpiThreads / tpt) : 1;
  MOE_REQUIRE(((E + 31) / 32) * (32 / tpt) <= kRouteWordsPerWarp, MOE_ERR_UNSUPPORTED_SHAPE, "moe_ffn_fused: E=%d", E);
  g.act = act;
  g.mask_h = mask_h;
  g.count_begin = count_begin;
  g.count_end = count_end;
  g.es_magic = static_cast<uint32_t>((0x100000000ull / static_cast<unsigned>(es)) + 1ull);

  if (getenv("MOE_DEBUG_PRINT"))
    fprintf(stderr,
            "[moe_ffn_fused] T=%d d=%d h=%d E=%d es=%d k=%d | nv=%d tiles1=%d ks1=%d | bn=%d tiles3=%d split=%d ks3=%d kb/slice=%d | "
            "stages=%d slot=%d | tpt=%d chunks/block=%d | items %d + %d on %d pairs\n",
            T, d, h, E, es, k, g.nv, g.n_tiles1, g.ks1, g.bn, g.n_tiles3, g.split3, g.ks3, g.kb_per_slice3, g.stages,
            g.slot_bytes, g.tpt, g.chunks_per_block, g.items1, g.items3, P);

  CUtensorMap tx, tw1, ths, thl, tw2, ty;
  int rc;
  if ((rc = make_tmap_bf16_kblocks(&tx, x, static_cast<uint64_t>(T), static_cast<uint64_t>(d), kBlockM, static_cast<uint32_t>(g.ks1)))) return rc;
  if ((rc = make_tmap_bf16_kblocks(&tw1, w1p, static_cast<uint64_t>(2) * h, static_cast<uint64_t>(d), static_cast<uint32_t>(nv),
                                   static_cast<uint32_t>(g.ks1))))
    return rc;
  if ((rc = make_tmap_bf16_2d(&ths, H, static_cast<uint64_t>(T), static_cast<uint64_t>(h), kBlockM, static_cast<uint32_t>(nv), false))) return rc;
  if ((rc = make_tmap_bf16_kblocks(&thl, H, static_cast<uint64_t>(T), static_cast<uint64_t>(h), kBlockM, static_cast<uint32_t>(g.ks3)))) return rc;
  if ((rc = make_tmap_bf16_kblocks(&tw2, w2p, static_cast<uint64_t>(d), static_cast<uint64_t>(h), static_cast<uint32_t>(g.bn / 2),
                                   static_cast<uint32_t>(g.ks3))))
    return rc;
  if ((rc = make_tmap_bf16_2d(&ty, Y, static_cast<uint64_t>(T), static_cast<uint64_t>(d), kBlockM, static_cast<uint32_t>(g.bn), false))) return rc;

  Ptrs a = {};
  a.b1 = b1p;
  a.b2 = b2;
  a.scores = scores;
  a.H = static_cast<__nv_bfloat16*>(H);
  a.Y = static_cast<__nv_bfloat16*>(Y);
  a.removed_bits = removed_bits;
  a.active_bits = active_bits;
  a.idx = idx;
  a.hist = hist;
  a.sync = static_cast<int*>(workspace);
  a.split_counters = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + kSyncBytes);
  a.split_partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + kSyncBytes + kSplitCounterBytes);

  const size_t smem = static_cast<size_t>(fixed) + static_cast<size_t>(g.stages) * g.slot_bytes;
  MOE_REQUIRE(smem <= static_cast<size_t>(kSmemLimit), MOE_ERR_UNSUPPORTED_SHAPE, "moe_ffn_fused: smem %zu", smem);

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(2 * P));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t le = cudaSuccess;
#define MOE_LAUNCH_FUSED(CHV)                                                                       \
  do {                                                                                              \
    rc = ensure_smem(reinterpret_cast<const void*>(ffn_fused_kernel<CHV>));                         \
    if (rc) return rc;                                                                              \
    le = cudaLaunchKernelEx(&cfg, ffn_fused_kernel<CHV>, tx, tw1, ths, thl, tw2, ty, g, a);         \
  } while (0)
  switch (ch1) {
    case 32: MOE_LAUNCH_FUSED(32); break;
    case 20: MOE_LAUNCH_FUSED(20); break;
    case 16: MOE_LAUNCH_FUSED(16); break;
    case 8: MOE_LAUNCH_FUSED(8); break;
    default: MOE_LAUNCH_FUSED(4); break;
  }
#undef MOE_LAUNCH_FUSED
  if (le != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_ffn_fused launch: %s", cudaGetErrorString(le));
  return check_launch("moe_ffn_fused");
}

}  // extern "C"
