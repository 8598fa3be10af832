// Per-token top-k routing shared by the fused layer kernel (ffn_fused.cu, routing stage) and the standalone
// router (router.cu, K2): several tokens per warp, exact k-th largest score by a bitonic sort (or, with a whole warp
// per token, an MSB-first bisection), ties to the lowest expert ids -- the same set torch.topk returns.
#pragma once
#include <stdint.h>
#include <cuda_bf16.h>

#ifndef MOE_ROUTE_TRACE
#define MOE_ROUTE_TRACE 0
#endif

namespace moe {
namespace route {

// order-preserving float -> uint32 key (ascending)
__device__ __forceinline__ uint32_t float_key(float s) {
  const uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ------------------------------------------------------------------------------------------ routing
// One chunk = 32 consecutive tokens of one 128-row block; 16 lanes own one token (2 tokens per warp), each lane
// KPT consecutive experts.  Exact k-th largest score by bisection on the order-preserving integer keys (MSB first,
// starting below the key prefix the whole warp shares, stopping as soon as both tokens of the warp have separated
// exactly k keys); ties on the k-th key go to the lowest expert ids.  Outputs: expert-set words (what the
// down-projection masks with), ascending labels, histogram (shared-memory bins) and, only when the caller wants
// the masked hidden state materialised, write-only zeroing of H (16-byte stores, a whole warp per token row).
//
// `G` / `A` are the caller's geometry / pointer structs (the fused layer kernel's Shape / Ptrs, the standalone
// router's RouterGeom / RouterPtrs): fields used are g.{lanes, lanes_log2, E, k, words, mask_h, h, es_magic,
// count_begin, count_end} and a.{scores, removed_bits, active_bits, idx, hist, H} (+ a.score_bias when kBias,
// + the per-lane running score maxima `mx` when kColmax: the standalone router's extras).
//
// The work is split in three stages so that a caller can order them: route_select (scores -> the lane's selection mask
// `sel`, expert-set words to shared and global memory), route_labels (ascending labels + histogram from `sel`) and
// route_zero_row (write-only masking of one token row of H from its expert-set words).  route_chunk runs them in the
// order select, labels, zero for the warp's own tokens (the standalone router); the fused layer kernel zeroes first,
// signals the block's consumers and writes the labels afterwards, off the critical path of the down-projection.
template <int KPT, bool kBias = false, bool kColmax = false, class G, class A>
__device__ __forceinline__ uint32_t route_select(const G& g, const A& a, int tok0, int tok_end, int ew, int lane,
                                                 uint32_t* s_words, float* mx = nullptr) {
  const unsigned full = 0xffffffffu;
  const int L = g.lanes;
  const int tpw = 32 >> g.lanes_log2;           // tokens per warp
  const int part = lane & (L - 1);
  const int tl = lane >> g.lanes_log2;
  // per-token sums over the token's L lanes with full-warp REDUX instructions: every token of the warp owns a
  // bit field of the reduced word (16 bits for 2 tokens per warp, 8 bits for 4 or 8 tokens -- 8 tokens take two
  // reductions); a field never overflows because a token has at most E <= 16 L experts.  (A reduction over a
  // per-token member mask compiles to a divergent slow path: ~550 cycles per bisection round.)
  const int fshift = (tpw <= 2) ? 16 * tl : 8 * (tl & 3);
  const uint32_t fmask = (tpw == 1) ? 0xffffffffu : (tpw == 2) ? 0xffffu : 0xffu;
  auto token_sum = [&](int c) -> int {
    const uint32_t v = static_cast<uint32_t>(c) << fshift;
    uint32_t r;
    if (tpw == 8) {
      const uint32_t r0 = __reduce_add_sync(full, tl < 4 ? v : 0u);
      const uint32_t r1 = __reduce_add_sync(full, tl < 4 ? 0u : v);
      r = tl < 4 ? r0 : r1;
    } else {
      r = __reduce_add_sync(full, v);
    }
    return static_cast<int>((r >> fshift) & fmask);
  };
  const int t = tok0 + tpw * ew + tl;
  const bool t_ok = t < tok_end;   // tok_end <= T: end of the 128-row block (a chunk never leaves its block)
  const int e0 = part * KPT;
  const int E = g.E;

  uint32_t key[KPT];
  uint32_t valid = 0u, removed = 0u;
  if (a.removed_bits != nullptr && e0 < E) removed = (__ldg(a.removed_bits + (e0 >> 5)) >> (e0 & 31)) & ((KPT == 32) ? ~0u : ((1u << KPT) - 1u));
  // standalone-router extras: per-expert score bias added before the selection (AddExperts), running maximum of the
  // raw scores per expert (ExpertPredictivity); a removed expert scores exactly 0 whatever its bias
  auto adjust = [&](int i, float raw, bool in) -> float {
    float v = raw + 0.f;      // -0.0 -> +0.0: the two compare equal, so they must tie (lowest expert id wins) like any equal pair
    if constexpr (kColmax) {
      if (in) mx[i] = fmaxf(mx[i], raw);
    }
    if constexpr (kBias) {
      if (in) v += __ldg(a.score_bias + e0 + i);
    }
    return ((removed >> i) & 1u) ? 0.f : v;   // zeroed pattern row => score exactly 0
  };
  {
    const float* row = a.scores + static_cast<size_t>(t_ok ? t : 0) * E + e0;
    if constexpr (KPT % 4 == 0) {
      if ((E & 3) == 0) {
#pragma unroll
        for (int i4 = 0; i4 < KPT / 4; ++i4) {
          float4 sc = make_float4(0.f, 0.f, 0.f, 0.f);
          const bool in = t_ok && (e0 + 4 * i4) < E;
          if (in) {
            sc = __ldcg(reinterpret_cast<const float4*>(row + 4 * i4));
            valid |= 0xfu << (4 * i4);
          }
          const float sv[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = 4 * i4 + j;
            const float v = adjust(i, sv[j], in);
            key[i] = in ? float_key(v) : 0u;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
          const bool in = t_ok && (e0 + i) < E;
          const float v = adjust(i, in ? __ldcg(row + i) : 0.f, in);
          key[i] = in ? float_key(v) : 0u;
          if (in) valid |= 1u << i;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < KPT; ++i) {
        const bool in = t_ok && (e0 + i) < E;
        const float v = adjust(i, in ? __ldcg(row + i) : 0.f, in);
        key[i] = in ? float_key(v) : 0u;
        if (in) valid |= 1u << i;
      }
    }
  }
#if MOE_ROUTE_TRACE
  if (ew == 0 && lane == 0) TRACE(56);
#endif

  uint32_t sel = 0u;   // KPT-bit mask over this lane's experts
  if (g.k >= E) {
    sel = valid;
  } else if (g.k > 0) {
    uint32_t prefix;
    if (L == 32) {
      // a whole warp per token: MSB-first bisection on the keys, one full-mask REDUX per round (~60 cycles of
      // dependent latency per round, ~20 rounds) -- a single warp walks the 36-stage sorting network of 256 keys in
      // ~4000 cycles, three times as long
      uint32_t k_or = 0u, k_and = full;
#pragma unroll
      for (int i = 0; i < KPT; ++i)
        if ((valid >> i) & 1u) {
          k_or |= key[i];
          k_and &= key[i];
        }
      k_or = __reduce_or_sync(full, k_or);
      k_and = __reduce_and_sync(full, k_and);
      const uint32_t diff = k_or ^ k_and;
      prefix = k_and;                       // every valid key identical
      if (diff != 0u && t_ok) {
        const int top = 31 - __clz(diff);
        prefix = (top == 31) ? 0u : (k_and & ~((2u << top) - 1u));
        for (int bit = top; bit >= 0; --bit) {
          const uint32_t thr = prefix | (1u << bit);
          int c = 0;
#pragma unroll
          for (int i = 0; i < KPT; ++i) c += (key[i] >= thr) ? 1 : 0;
          c = __reduce_add_sync(full, c);
          if (c >= g.k) prefix = thr;
          if (c == g.k) break;              // exactly k keys at or above the threshold: done (warp-uniform)
        }
      }
    } else {
    // k-th largest key of the token: bitonic sort of its N = KPT * L keys (element i = part * KPT + r; invalid
    // slots hold key 0 and sink to the bottom), compare-exchanges inside a lane for distances < KPT and through
    // shuffles beyond.  ~250 instructions and ~700 cycles per pass regardless of the data; the MSB-first bisection
    // this replaces needed ~20 dependent REDUX + VOTE rounds (3-6 us per pass).
    uint32_t srt[KPT];
#pragma unroll
    for (int r = 0; r < KPT; ++r) srt[r] = key[r];
    // Bitonic sort in the "flip" formulation: every compare-exchange keeps the minimum at the LOWER element index, so
    // there are no direction selects (the classic formulation costs two SELs per compare-exchange).  Merging blocks of
    // size k: first compare element i with its mirror image i ^ (k - 1) inside the block, then with i ^ j for
    // j = k / 4 ... 1.  Distances below KPT stay inside the lane (compile-time register pairs, min / max only);
    // beyond, the partner lane is part ^ mask and, for the mirror step, the register index is reversed.
#pragma unroll
    for (int k2 = 2; k2 <= KPT; k2 <<= 1) {
#pragma unroll
      for (int r = 0; r < KPT; ++r) {          // mirror step inside the lane
        const int q = r ^ (k2 - 1);
        if (r < q) {
          const uint32_t lo = min(srt[r], srt[q]), hi = max(srt[r], srt[q]);
          srt[r] = lo;
          srt[q] = hi;
        }
      }
#pragma unroll
      for (int j = k2 >> 2; j >= 1; j >>= 1) {
#pragma unroll
        for (int r = 0; r < KPT; ++r) {
          if ((r & j) == 0) {
            const uint32_t lo = min(srt[r], srt[r | j]), hi = max(srt[r], srt[r | j]);
            srt[r] = lo;
            srt[r | j] = hi;
          }
        }
      }
    }
    for (int lsize = 2; lsize <= L; lsize <<= 1) {        // blocks of k = lsize * KPT elements
      {
        // mirror step: lane part ^ (lsize - 1), register KPT - 1 - r; the lane with the lower index keeps the minima
        const bool keep_min = (part & (lsize >> 1)) == 0;
        uint32_t other[KPT];
#pragma unroll
        for (int r = 0; r < KPT; ++r) other[r] = __shfl_xor_sync(full, srt[KPT - 1 - r], lsize - 1);
#pragma unroll
        for (int r = 0; r < KPT; ++r) srt[r] = keep_min ? min(srt[r], other[r]) : max(srt[r], other[r]);
      }
      for (int ls = lsize >> 2; ls >= 1; ls >>= 1) {
        const bool keep_min = (part & ls) == 0;
#pragma unroll
        for (int r = 0; r < KPT; ++r) {
          const uint32_t other = __shfl_xor_sync(full, srt[r], ls);
          srt[r] = keep_min ? min(srt[r], other) : max(srt[r], other);
        }
      }
#pragma unroll
      for (int j = KPT >> 1; j >= 1; j >>= 1) {
#pragma unroll
        for (int r = 0; r < KPT; ++r) {
          if ((r & j) == 0) {
            const uint32_t lo = min(srt[r], srt[r | j]), hi = max(srt[r], srt[r | j]);
            srt[r] = lo;
            srt[r | j] = hi;
          }
        }
      }
    }
    // ascending over the token's lanes: the k-th largest sits at element N - k
    const int pos = KPT * L - g.k;
    uint32_t mine = 0u;
#pragma unroll
    for (int r = 0; r < KPT; ++r) mine = (r == (pos & (KPT - 1))) ? srt[r] : mine;
    prefix = __shfl_sync(full, mine, (lane & ~(L - 1)) + pos / KPT);
    }
    // keys above the k-th value, then ties on it from the lowest expert id
    uint32_t gt = 0u, eq = 0u;
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      const bool v = (valid >> i) & 1u;
      gt |= (v && key[i] > prefix) ? (1u << i) : 0u;
      eq |= (v && key[i] == prefix) ? (1u << i) : 0u;
    }
    const int n_gt = token_sum(__popc(gt));
    const int n_eq = __popc(eq);
    int incl = n_eq;
    for (int o = 1; o < L; o <<= 1) {
      const int v = __shfl_up_sync(full, incl, o, L);
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
#if MOE_ROUTE_TRACE
  if (ew == 0 && lane == 0) TRACE(57);
#endif

  // expert-set words: OR the lanes' masks of each 32-expert word together (butterfly inside the word's lanes)
  const uint32_t active = sel & ~removed;
  const int lanes_per_word = (32 / KPT) < L ? (32 / KPT) : L;
  uint32_t word = active << ((part * KPT) & 31);
  for (int o = 1; o < lanes_per_word; o <<= 1) word |= __shfl_xor_sync(full, word, o);
  const int widx = (part * KPT) >> 5;
  if ((part & (lanes_per_word - 1)) == 0 && widx < g.words) {
    s_words[tl * g.words + widx] = word;
    if (t_ok && a.active_bits != nullptr) a.active_bits[static_cast<size_t>(t) * g.words + widx] = word;
  }
  return sel;
}

// ascending labels and histogram of the warp's tokens from the lanes' selection masks
template <int KPT, class G, class A>
__device__ __forceinline__ void route_labels(const G& g, const A& a, uint32_t sel, int tok0, int tok_end, int ew, int lane,
                                             unsigned int* s_hist) {
  const unsigned full = 0xffffffffu;
  const int L = g.lanes;
  const int tpw = 32 >> g.lanes_log2;
  const int part = lane & (L - 1);
  const int tl = lane >> g.lanes_log2;
  const int t = tok0 + tpw * ew + tl;
  const bool t_ok = t < tok_end;
  const int e0 = part * KPT;
  if (a.idx != nullptr) {
    const int n_sel = __popc(sel);
    int incl = n_sel;
    for (int o = 1; o < L; o <<= 1) {
      const int v = __shfl_up_sync(full, incl, o, L);
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
}

// does the routing stage materialise the masked hidden state?  (k == E still masks the removed experts' neurons)
template <class G, class A>
__device__ __forceinline__ bool route_masks_h(const G& g, const A& a) {
  return g.mask_h && (g.k < g.E || a.removed_bits != nullptr);
}

// Zero the neurons of every expert outside token `tok`'s active set (expert-set words at `wtok`), 16-byte units
// [u_begin + lane, u_end) step 32 of its row of H: write-only, 4-neuron groups never straddle an expert because
// es % 4 == 0.  Two predicated 8-byte stores per unit and no branches: a three-way if / else (16-byte store, or either
// half) made the warp run each store flavour in turn.  (A variant that kept the expert words in registers and hoisted
// the two divisions per unit out of the loop was measured 1.5 % SLOWER end to end: at UNet batch 16 this stage is
// paced by the 7 TB/s of zero stores into L2, not by its instructions -- profiles/r02_route_ab.log.)
template <class G, class A>
__device__ __forceinline__ void route_zero_row(const G& g, const A& a, int tok, const uint32_t* wtok, int u_begin,
                                               int u_end, int lane) {
  uint4* hrow = reinterpret_cast<uint4*>(a.H + static_cast<size_t>(tok) * g.h);
#pragma unroll 4
  for (int u = u_begin + lane; u < u_end; u += 32) {
    const uint32_t ea = __umulhi(static_cast<uint32_t>(u) << 3, g.es_magic);
    const uint32_t eb = __umulhi((static_cast<uint32_t>(u) << 3) + 4u, g.es_magic);
    const bool on_a = (wtok[ea >> 5] >> (ea & 31u)) & 1u;
    const bool on_b = (wtok[eb >> 5] >> (eb & 31u)) & 1u;
    uint2* half = reinterpret_cast<uint2*>(hrow + u);
    if (!on_a) half[0] = make_uint2(0u, 0u);
    if (!on_b) half[1] = make_uint2(0u, 0u);
  }
}

template <int KPT, bool kBias = false, bool kColmax = false, class G, class A>
__device__ __forceinline__ void route_chunk(const G& g, const A& a, int tok0, int tok_end, int ew, int lane,
                                            uint32_t* s_words, unsigned int* s_hist, float* mx = nullptr) {
  const uint32_t sel = route_select<KPT, kBias, kColmax>(g, a, tok0, tok_end, ew, lane, s_words, mx);
  route_labels<KPT>(g, a, sel, tok0, tok_end, ew, lane, s_hist);
  __syncwarp();
#if MOE_ROUTE_TRACE
  if (ew == 0 && lane == 0) TRACE(58);
#endif
  if (route_masks_h(g, a)) {
    const int tpw = 32 >> g.lanes_log2;
    for (int tt = 0; tt < tpw; ++tt) {
      const int tok = tok0 + tpw * ew + tt;
      if (tok >= tok_end) break;
      route_zero_row(g, a, tok, s_words + tt * g.words, 0, g.h >> 3, lane);
    }
  }
  __syncwarp();
#if MOE_ROUTE_TRACE
  if (ew == 0 && lane == 0) TRACE(59);
#endif
}

// The fused layer kernel's order: select -> zero-writes -> signal -> labels.  `signal` tells the block's consumers that
// the chunk's rows of H are masked (called by every warp, only after the item's last chunk); labels and histogram
// follow it: nobody waits for them before the kernel ends.  (Spreading a token row's zero-writes over the warps that
// do not route -- chunks of 4 tokens at d = 1280 -- was measured 0.4 us SLOWER per layer call: the stage is a store
// round trip, not an instruction count, and the variant needs two more 512-thread barriers.  Publishing a block's
// scores under their own counter, ahead of its H tile stores, so that the selection overlaps the stores' completion
// was neutral to 0.8 us slower: the pair's last tile then pays two GPU-scope fences in a row.)
template <class G, class A, class Signal>
__device__ __forceinline__ void route_dispatch_ordered(const G& g, const A& a, int tok0, int tok_end, int ew, int lane,
                                                       uint32_t* s_words, unsigned int* s_hist, bool last_chunk,
                                                       Signal signal) {
  const bool routing_warp = ew < g.route_warps;   // small chunks (many consumers per block) use only the first warps
  uint32_t sel = 0u;
  if (routing_warp) {
    switch (g.kpt) {
      case 16: sel = route_select<16>(g, a, tok0, tok_end, ew, lane, s_words); break;
      case 8: sel = route_select<8>(g, a, tok0, tok_end, ew, lane, s_words); break;
      case 4: sel = route_select<4>(g, a, tok0, tok_end, ew, lane, s_words); break;
      case 2: sel = route_select<2>(g, a, tok0, tok_end, ew, lane, s_words); break;
      case 1: sel = route_select<1>(g, a, tok0, tok_end, ew, lane, s_words); break;
      default: break;
    }
  }
  __syncwarp();
#if MOE_ROUTE_TRACE
  if (ew == 0 && lane == 0) TRACE(58);
#endif
  if (routing_warp && route_masks_h(g, a)) {
    const int tpw = 32 >> g.lanes_log2;
    for (int tt = 0; tt < tpw; ++tt) {
      const int tok = tok0 + tpw * ew + tt;
      if (tok >= tok_end) break;
      route_zero_row(g, a, tok, s_words + tt * g.words, 0, g.h >> 3, lane);
    }
  }
  __syncwarp();
#if MOE_ROUTE_TRACE
  if (ew == 0 && lane == 0) TRACE(59);
#endif
  if (last_chunk) signal();
  if (routing_warp) {
    switch (g.kpt) {
      case 16: route_labels<16>(g, a, sel, tok0, tok_end, ew, lane, s_hist); break;
      case 8: route_labels<8>(g, a, sel, tok0, tok_end, ew, lane, s_hist); break;
      case 4: route_labels<4>(g, a, sel, tok0, tok_end, ew, lane, s_hist); break;
      case 2: route_labels<2>(g, a, sel, tok0, tok_end, ew, lane, s_hist); break;
      case 1: route_labels<1>(g, a, sel, tok0, tok_end, ew, lane, s_hist); break;
      default: break;
    }
    __syncwarp();   // the next chunk's selection overwrites s_words / reuses the lanes' registers in lock step
  }
}

}  // namespace route
}  // namespace moe
