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
// Cross-CTA dependencies go through a small global sync area, per 128-row block m:
//     done[m]    += 1 for every phase-1 tile of the block whose H tile (TMA store) and scores are written; 
//     routing    the block's phase-3 items ("consumers") split its routing chunks among themselves statically
//                (consumer r takes chunks r, r + consumers, ...) and route them with their 16 epilogue warps before
//                their own main loop; nothing is routed while phase 1 runs.  (A global work queue with
//                compare-and-swap tickets was tried: 148 CTAs hammering one L2 line serialised it completely.)
//     ready[m]   += 1 per routed chunk; a phase-3 A-tile producer waits for ready[m] == chunks per block
// Masking (the reference's gate[mask == 0] = 0) is write-only: the routing stage zeroes the neurons of inactive
// experts in H with 16-byte stores, a whole warp per token row, spread over the block's consumer CTAs.  (Masking
// the landed H tiles in shared memory inside the phase-3 pipeline was tried and measured: the extra
// wait / fence / cluster-arrive per stage cost ~1 us per stage, three times the main loop itself.)
// All CTAs are co-resident (grid <= SM count, 1 CTA / SM) and every CTA finishes its phase-1 tiles, which
// never wait on another CTA, before it waits for anything, so the waits cannot deadlock.  The last consumer of a
// block zeroes its record again (the workspace must be zero before the first launch).
//
// Warp roles (640 threads):
//   warp 0      A-tile TMA producer (x in phase 1, H in phase 3 -- waits for ready[m])
//   warp 1      MMA issuer (leader CTA of the pair): tcgen05.mma.cta_group::2, accumulators in TMEM (2 stages)
//   warp 2      TMEM allocator, then the store / sync warp: TMA-stores finished output tiles from the smem
//               staging buffers, publishes done[m], claims routing chunks and posts them to the epilogue warps
//   warp 3      B-tile TMA producer (W1 / W2 slices; weights have no dependencies, so it runs ahead)
//   warps 4-19  epilogue: TMEM -> registers -> bias / exact GELU / product (packed f32x2 math) -> bf16 ->
//               smem staging; expert scores; routing chunks on request; phase-3 maskers; split-K reduction
//
// Two experimental schedules of phase 1 are kept behind environment switches (both measured neutral to slower, both
// bit-identical to the default and parity-tested): MOE_FUSED_ARES=1 keeps a row block's x panels
// resident in shared memory and streams only W1 through the ring (contiguous runs of column tiles per CTA pair);
// MOE_FUSED_DIRECT=1 stores H straight from the epilogue threads and turns the staging buffers into a fourth ring
// slot.  Both run phase 1 on its own ring (barriers full_b / empty_b) and re-carve shared memory for phase 3 once
// every phase-1 MMA has completed (p1_done).
#include "common.cuh"
#include "tcgen05.cuh"

#ifndef MOE_TRACE
#define MOE_TRACE 0
#endif
#define MOE_ROUTE_TRACE MOE_TRACE

namespace moe {
namespace fused {
#if MOE_TRACE
__device__ unsigned long long g_trace[256 * 64];
#define TRACE(slot)                                                                                                                   \
  do {                                                                                                                                \
    if (blockIdx.x < 256 && (slot) < 64) ::moe::fused::g_trace[blockIdx.x * 64 + (slot)] = ::moe::tc::global_timer_ns();              \
  } while (0)
#else
#define TRACE(slot) \
  do {              \
  } while (0)
#endif
}  // namespace fused
}  // namespace moe

#include "route.cuh"      // (after TRACE: the routing code carries timeline stamps in the trace build)

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
constexpr int kMaxStages = 12;
constexpr int kSmemLimit = 232448;
constexpr int kBiasBytesPerWarp = 64 * 4;           // phase 1: 32 value + 32 gate biases (cpg <= 32), phase 3: up to 64 output biases
constexpr int kBiasGateOff = 32;
constexpr int kSpartPerRow = 8;                      // partial / per-expert score slots per token row and tile
constexpr int kMaxWords = 8;                         // expert-set words per token (E <= 256)
constexpr int kMaxExperts = 32 * kMaxWords;
// sync area (ints): header (unused), then one 128-byte record per 128-row block {done, ready, visits} (separate
// cache lines: pollers of different blocks do not serialise on one L2 line)
constexpr int kSyncHeaderInts = 32;
constexpr int kBlockRecInts = 32;
constexpr int kMaxBlocks = 2048;                     // 128-row blocks (T <= 262144 tokens)
constexpr size_t kSyncBytes = (kSyncHeaderInts + static_cast<size_t>(kMaxBlocks) * kBlockRecInts) * 4;
constexpr size_t kSplitCounterBytes = 64 * 1024;

struct Barriers {   // the mbarriers first, in this order: the kernel initialises them by index
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t hs_full[2];     // staging buffer written by the 16 epilogue warps
  uint64_t hs_empty[2];    // staging buffer stored by the sync warp
  uint64_t route_req;      // sync warp -> epilogue warps: a routing chunk is posted
  uint64_t route_done;     // 16 epilogue warps -> sync warp
  // resident-A mode of phase 1 (Shape::a_resident): its own ring barriers (B operand only), the resident A block
  // and the end of phase 1's tensor work
  uint64_t full_b[kMaxStages];
  uint64_t empty_b[kMaxStages];
  uint64_t a_full;         // a row block's x panels have landed (leader CTA's copy counts both CTAs' bytes)
  uint64_t a_empty;        // every MMA that reads the resident panels has completed
  uint64_t p1_done;        // every phase-1 MMA has completed: shared memory may be re-carved for phase 3
  uint32_t tmem_base;
  int req_block, req_chunk;
  int last_cta;
};

struct Shape {
  int T, d, h, E, es, k;
  int m_pairs;                                  // 256-row work units
  // phase 1
  int nv, n_tiles1, ks1, nkb1, items1;
  int step_m1, step_n1;                         // a pair's next phase-1 tile: (mp, n) += (step_m1, step_n1), carry n -> mp
  // resident-A mode (small d): a pair takes a contiguous run of tiles (same row block, consecutive column tiles), the
  // row block's x panels stay in shared memory for the whole run and only W1 streams through the ring
  int a_resident, a_res_bytes, stages1, slot1_bytes;
  // phase 1 on its own ring (barriers full_b / empty_b, `stages1` slots of `slot1_bytes` from byte `ring1_off`): set by
  // resident-A mode and by direct-H mode, where the epilogue stores H straight to global memory and the staging
  // buffers' space becomes a fourth ring slot for the phase whose main loop is bound by k-blocks in flight
  int sep_ring1, ring1_off, direct_h;
  int tail_off, spart_bytes;                    // small per-warp arrays + barriers start at tail_off
  int hs_off;                                   // the two staging buffers start here (after the operand ring area)
  int experts_per_tile, chunks_per_expert, span;
  // phase 3
  int bn, n_tiles3, ks3, nkb3, split3, kb_per_slice3, items3;
  long long split_plane4;                       // float4 elements between two K slices of the split-K partial tiles
  // pipeline
  int stages, slot_bytes, hs_bytes;             // hs_bytes: one phase-1 staging buffer (two of them; phase 3 uses both)
  // routing
  int lanes, lanes_log2;                        // lanes per token in the routing stage (4, 8 or 16)
  int kpt;                                      // experts per routing lane: ceil(E / lanes) rounded up to a power of two
  int route_warps;                              // epilogue warps that take part in routing a chunk (1..16)
  int chunk_tokens, chunks_per_block;           // chunk = route_warps * 32 / lanes tokens = one pass of those warps
  int words;                                    // expert-set words per token
  int act, mask_h, count_begin, count_end;
  int prefetch_weights;
  int rev3;                                     // phase-3 items walk the row blocks from the last one down
  int trace_p3;                                 // trace build only: the per-item stamps follow the phase-3 items
  int hist_slots;                               // shared-memory histogram bins: E rounded up to a multiple of 32
  int pub_batch;                                // phase-1 tiles published per GPU-scope release (2 .. 4)
  uint32_t es_magic;
};

struct Ptrs {
  const void* w1;           // raw weight pointers, only for the L2 prefetch before the dependency wait
  const void* w2;
  const float* b1;
  const float* b2;
  float* scores;
  __nv_bfloat16* H;
  __nv_bfloat16* Y;
  const uint32_t* removed_bits;
  uint32_t* active_bits;
  int16_t* idx;
  unsigned long long* hist;
  int* sync;                // header | per-block records | routing queue
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
// generic-proxy accesses to global memory (here: writes of other SMs that this thread has just acquired) before this
// thread's async-proxy (TMA) accesses to global memory: one FENCE.VIEW.ASYNC.G instead of the all-space form's
// GPU-scope membar + L1 invalidate
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// returns the old value; release at GPU scope (the partial tiles read afterwards come from L2 with ld.cg, so no L1
// invalidation -- the acquire half of an acq_rel atomic -- is needed)
__device__ __forceinline__ int atom_add_release(int* p, int v) {
  int old;
  asm volatile("atom.release.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
// counter += v with release semantics at GPU scope: every write this thread has performed or observed (the tile's
// scores through the hs_full barrier, the H tile through cp.async.bulk.wait_group) is visible to whoever acquires
// the new count.  One MEMBAR.ALL.GPU + RED; the __threadfence() + fence.proxy.async pair used before cost two
// GPU-scope membars and two L1 invalidations (1.6 us per tile, which made the store warp pace phase 1).
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Poll a cross-CTA counter until it reaches `target`.  A protocol bug or a workspace that was not zero (e.g. left
// dirty by an aborted launch) would otherwise hang the GPU until an external timeout; the watchdog turns a wait of
// more than 2 s into a trap (launch failure reported to the host), like tc::mbar_wait does for mbarriers.
__device__ __forceinline__ void poll_at_least(const int* flag, int target) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  // one lane per CTA polls its own block's 128-byte record: relaxed loads (an acquire load is LDG + CCTL.IVALL, an
  // L1 invalidation per iteration), one acquire load once the count is there
  for (;;) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= target) break;
    if ((++spins & 0xFFFFu) == 0) {
      const uint64_t now = tc::global_timer_ns();
      if (t0 == 0)
        t0 = now;
      else if (now - t0 > 2000000000ull)
        __trap();
    }
  }
  (void)ld_acquire(flag);   // the acquire (and its one L1 invalidation) once the count is there
}

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

// exact-erf GELU of two values: x Phi(x) = max(x, 0) - |x| * 0.5 erfc(|x| / sqrt 2), 0.5 erfc(u / sqrt 2) = 2^P7(u)
// for u = min(|x|, 4.3 sqrt 2) (weighted minimax fit, the 1 / sqrt 2 folded into the coefficients; max abs error
// 4e-7 on [-3, 3]); the polynomial runs as seven packed FFMA2, the rest is one MUFU.EX2 and three scalar ops per value.
// gemm_tc.cu's gelu_erf evaluates the same expression with scalar FFMAs (bit-identical results).
__device__ __forceinline__ uint64_t gelu2(float x0, float x1) {
  const float t0 = fminf(fabsf(x0), 6.081118318204309f);
  const float t1 = fminf(fabsf(x1), 6.081118318204309f);
  const uint64_t t = pk2(t0, t1);
  uint64_t q = pk2(9.425939424545504e-06f, 9.425939424545504e-06f);
  q = fma2(q, t, pk2(-6.281803507590666e-05f, -6.281803507590666e-05f));
  q = fma2(q, t, pk2(-3.898592258337885e-04f, -3.898592258337885e-04f));
  q = fma2(q, t, pk2(7.337003480643034e-03f, 7.337003480643034e-03f));
  q = fma2(q, t, pk2(-5.264890193939209e-02f, -5.264890193939209e-02f));
  q = fma2(q, t, pk2(-4.591682255268097e-01f, -4.591682255268097e-01f));
  q = fma2(q, t, pk2(-1.1511090993881226f, -1.1511090993881226f));
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
  else if constexpr (ACT == MOE_ACT_RELU)
    return pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
  else
    return pk2(x0, x1);   // profiling only (trace build): no activation math
}

// Every phase-3 item touches its block's record twice (its sync warp after the done-poll and its ready-add, its
// A-tile producer after the ready-wait); the last of these 2 x consumers visits zeroes the record, so the sync area
// is clean again when the kernel ends -- off the critical path, unlike a last-CTA-out sweep at kernel exit.
__device__ __forceinline__ void block_consumed(int* rec, int visits) {
  const int prev = atomicAdd(rec + 2, 1);
  if (prev == visits - 1) {
    rec[0] = 0;
    rec[1] = 0;
    rec[2] = 0;
  }
}

// ------------------------------------------------------------------------------------------ work list
struct Item {
  int m_blk;               // this CTA's 128-row block
  int n;                   // column tile
  int kb_begin, kb_end;    // k-block range
  int slice;
};
// phase-1 tiles of pair `p` (P pairs), row-block-major tile index: p, p + P, p + 2 P, ... (round-robin), or, in
// resident-A mode, a contiguous run [begin, end) of a balanced partition (the first items1 % P pairs hold one more)
__device__ __forceinline__ void range1(const Shape& g, int p, int P, int& begin, int& end, int& stride) {
  if (g.a_resident) {
    const int q = g.items1 / P, r = g.items1 - q * P;
    begin = p * q + min(p, r);
    end = begin + q + (p < r ? 1 : 0);
    stride = 1;
  } else {
    begin = p;
    end = g.items1;
    stride = P;
  }
}
__device__ __forceinline__ bool item1(const Shape& g, int it, int p, int P, int rm, Item& t) {
  int begin, end, stride;
  range1(g, p, P, begin, end, stride);
  const int i = begin + it * stride;
  if (i >= end) return false;
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
  int mp = r / g.n_tiles3;
  t.n = r - mp * g.n_tiles3;
  if (g.rev3) mp = g.m_pairs - 1 - mp;      // last-written row blocks first: their H tiles are still in L2
  t.m_blk = 2 * mp + rm;
  t.kb_begin = t.slice * g.kb_per_slice3;
  t.kb_end = min(g.nkb3, t.kb_begin + g.kb_per_slice3);
  return true;
}

using route::route_chunk;

// phase-1 tiles of a pair without divisions in the loop (the MMA warp walks its list on the critical path of every
// tile: item1() costs ~0.4 us of integer divisions per call)
struct Tile1Iter {
  int i, end, stride, mp, n, n_tiles1, step_m, step_n;
  __device__ __forceinline__ void init(const Shape& g, int p, int P) {
    range1(g, p, P, i, end, stride);
    n_tiles1 = g.n_tiles1;
    step_m = g.step_m1;
    step_n = g.step_n1;
    mp = i / n_tiles1;
    n = i - mp * n_tiles1;
  }
  __device__ __forceinline__ bool valid() const { return i < end; }
  __device__ __forceinline__ void next() {
    i += stride;
    mp += step_m;
    n += step_n;
    if (n >= n_tiles1) {
      n -= n_tiles1;
      ++mp;
    }
  }
};

// ------------------------------------------------------------------------------------------ the kernel
// per-warp bias slices staged in shared memory: lanes load coalesced, everyone re-reads float4 broadcasts
// (phase 1: n <= 32 value biases at sb[0 ..] and the gate biases at sb[kBiasGateOff ..]; phase 3: n <= 64 biases at sb[0 ..])
__device__ __forceinline__ void stage_bias(float* sb, const float* b0, const float* b1, int n, int lane, bool two_halves) {
  __syncwarp();
  if (two_halves) {
    sb[lane] = (b0 != nullptr && lane < n) ? __ldg(b0 + lane) : 0.f;
    sb[kBiasGateOff + lane] = (b1 != nullptr && lane < n) ? __ldg(b1 + lane) : 0.f;
  } else {
    for (int i = lane; i < 64; i += 32) sb[i] = (b0 != nullptr && i < n) ? __ldg(b0 + i) : 0.f;
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

// Output staging tiles (H in phase 1, Y in phase 3) are kept in the layout the TMA store expects with swizzling:
// 64-column panels ([128 rows][128 bytes], 128-byte swizzle) followed by one remainder panel of W % 64 columns
// (64- / 32-byte swizzle when it is 32 / 16 columns wide).  A thread owns one row, so a warp's store hits 32
// consecutive rows at the same column: in a dense [128][W] tile that is an 8- or 16-way bank conflict on every
// store, swizzled it is conflict-free.
struct StageLayout {
  int n_full;          // 64-column panels
  int rem_cols;        // columns of the remainder panel (0 if none)
  uint32_t rem_smask;  // swizzle mask of the remainder panel: ((byte >> 7) & rem_smask) << 4 is XORed in
};
__device__ __forceinline__ StageLayout stage_layout(int width) {
  StageLayout l;
  l.n_full = width >> 6;
  l.rem_cols = width & 63;
  l.rem_smask = l.rem_cols == 32 ? 3u : (l.rem_cols == 16 ? 1u : 0u);
  return l;
}
// shared-memory address of columns [x, x + 4) (x a multiple of 4, warp-uniform) of row r
__device__ __forceinline__ uint32_t stage_addr(uint32_t base, const StageLayout& l, int r, int x) {
  if (x < 64 * l.n_full)
    return base + (x >> 6) * (kBlockM * 128) + r * 128 + ((((x & 63) * 2)) ^ ((r & 7) << 4));
  const int pitch = l.rem_cols * 2;
  const uint32_t lin = static_cast<uint32_t>(r * pitch);
  return base + l.n_full * (kBlockM * 128) + lin + ((static_cast<uint32_t>(x - 64 * l.n_full) * 2u) ^ (((lin >> 7) & l.rem_smask) << 4));
}
// store kWords 32-bit words (2 columns each) starting at column x0: 16-byte stores on 8-column boundaries
template <int kWords>
__device__ __forceinline__ void stage_store(uint32_t base, const StageLayout& l, int r, int x0, const uint32_t* w) {
  static_assert(kWords % 2 == 0, "whole 4-column pieces");
  if ((x0 & 7) == 0) {
#pragma unroll
    for (int i = 0; i + 4 <= kWords; i += 4) tc::sts_b32x4(stage_addr(base, l, r, x0 + 2 * i), w[i], w[i + 1], w[i + 2], w[i + 3]);
    if constexpr (kWords % 4 != 0) tc::sts_b32x2(stage_addr(base, l, r, x0 + 2 * (kWords - 2)), w[kWords - 2], w[kWords - 1]);
  } else {
    tc::sts_b32x2(stage_addr(base, l, r, x0), w[0], w[1]);
#pragma unroll
    for (int i = 2; i + 4 <= kWords; i += 4) tc::sts_b32x4(stage_addr(base, l, r, x0 + 2 * i), w[i], w[i + 1], w[i + 2], w[i + 3]);
    if constexpr (kWords % 4 == 0) tc::sts_b32x2(stage_addr(base, l, r, x0 + 2 * (kWords - 2)), w[kWords - 2], w[kWords - 1]);
  }
}

// the same with the pieces' byte offsets computed once per kernel (they depend on the thread's row and column group
// only): `off[i]` = offset of 4-column piece i of the chunk; `aligned` = the chunk starts on an 8-column boundary
template <int kWords, bool aligned>
__device__ __forceinline__ void stage_store_pre(uint32_t base, const uint32_t* off, const uint32_t* w) {
  if constexpr (aligned) {
#pragma unroll
    for (int i = 0; i + 4 <= kWords; i += 4) tc::sts_b32x4(base + off[i / 2], w[i], w[i + 1], w[i + 2], w[i + 3]);
    if constexpr (kWords % 4 != 0) tc::sts_b32x2(base + off[kWords / 2 - 1], w[kWords - 2], w[kWords - 1]);
  } else {
    tc::sts_b32x2(base + off[0], w[0], w[1]);
#pragma unroll
    for (int i = 2; i + 4 <= kWords; i += 4) tc::sts_b32x4(base + off[i / 2], w[i], w[i + 1], w[i + 2], w[i + 3]);
    if constexpr (kWords % 4 == 0) tc::sts_b32x2(base + off[kWords / 2 - 1], w[kWords - 2], w[kWords - 1]);
  }
}

// store one score to global memory under a predicate, without a branch (a per-thread `if` around the store costs a
// BSSY / BSYNC pair per expert in the epilogue's inner loop)
__device__ __forceinline__ void st_global_if(float* p, float v, bool pred) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v),
               "r"(static_cast<int>(pred))
               : "memory");
}

// One thread's share of a phase-1 tile: row q_row, the cpg (<= 32) value columns of its column group and the matching
// gate columns, CH columns at a time.  The chunk loop is unrolled over its compile-time maximum, so the staging
// offsets stay in registers; everything that does not change from tile to tile is computed by the caller.
//   taddr_v / taddr_g  TMEM address of the group's first value / gate column (this warp's lane quarter)
//   cpe                chunks per expert when whole experts lie inside a column group, else 0
//   score_dst          global slot of the group's first expert score for this row (stored only if score_ok)
//   spart_slot         shared-memory slot of this group's partial sum (experts that span column groups)
// direct-H mode: kWords words (2 columns each) of one row straight to global memory; `aligned` = dst is 16-byte aligned
template <int kWords, bool aligned>
__device__ __forceinline__ void global_store_row(__nv_bfloat16* dst, const uint32_t* w) {
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  if constexpr (aligned) {
#pragma unroll
    for (int i = 0; i + 4 <= kWords; i += 4) *reinterpret_cast<uint4*>(d + i) = make_uint4(w[i], w[i + 1], w[i + 2], w[i + 3]);
    if constexpr (kWords % 4 != 0) *reinterpret_cast<uint2*>(d + kWords - 2) = make_uint2(w[kWords - 2], w[kWords - 1]);
  } else {
    *reinterpret_cast<uint2*>(d) = make_uint2(w[0], w[1]);
#pragma unroll
    for (int i = 2; i + 4 <= kWords; i += 4) *reinterpret_cast<uint4*>(d + i) = make_uint4(w[i], w[i + 1], w[i + 2], w[i + 3]);
    if constexpr (kWords % 4 == 0) *reinterpret_cast<uint2*>(d + kWords - 2) = make_uint2(w[kWords - 2], w[kWords - 1]);
  }
}

template <int CH, int ACT, bool kAligned, bool kDirect>
__device__ __forceinline__ void geglu_group(uint32_t taddr_v, uint32_t taddr_g, const float* sbias, uint32_t hbase,
                                            const uint32_t (&piece_off)[8], __nv_bfloat16* hrow, float* spart_slot,
                                            float* score_dst, bool score_ok, int cpg, int cpe) {
  uint64_t score2 = pk2(0.f, 0.f);
  int chunk_in_expert = 0;
  // a chunk's TMEM loads are split in two: the second half is in flight while the first half is processed
  // (TMEM reads run at ~16 B/clk per lane quarter -- a whole tile takes ~1300 cycles to read out)
  constexpr int kA = (CH >= 8) ? (CH / 2) / 4 * 4 : CH;
  constexpr int kMaxChunks = 32 / CH;
#pragma unroll
  for (int ci = 0; ci < kMaxChunks; ++ci) {
    const int c = ci * CH;
    if (ci > 0 && c >= cpg) break;
    uint32_t v[CH], gt[CH];
    tc::tmem_ld_cols<kA>(taddr_v + c, v);
    tc::tmem_ld_cols<kA>(taddr_g + c, gt);
    tc::tmem_ld_wait();
    if constexpr (kA < CH) {
      tc::tmem_ld_cols<CH - kA>(taddr_v + c + kA, v + kA);
      tc::tmem_ld_cols<CH - kA>(taddr_g + c + kA, gt + kA);
    }
    uint32_t hw[CH / 2];
#pragma unroll
    for (int i = 0; i < CH; i += 4) {
      if constexpr (kA < CH) {
        if (i == kA) tc::tmem_ld_wait();
      }
      const float4 bv = *reinterpret_cast<const float4*>(sbias + c + i);
      const float4 bg = *reinterpret_cast<const float4*>(sbias + kBiasGateOff + c + i);
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
    if constexpr (kDirect) {
      if (score_ok) global_store_row<CH / 2, kAligned>(hrow + c, hw);   // (score_ok: the row is inside the matrix)
    } else {
      stage_store_pre<CH / 2, kAligned>(hbase, &piece_off[(ci * CH) / 4], hw);
    }
    if (cpe > 0 && ++chunk_in_expert == cpe) {
      float s0, s1;
      unpk2(score2, s0, s1);
      // whole expert inside this thread's column group: its score goes straight to global memory
      st_global_if(score_dst, s0 + s1, score_ok);
      ++score_dst;
      score2 = pk2(0.f, 0.f);
      chunk_in_expert = 0;
    }
  }
  if (cpe == 0) {
    float s0, s1;
    unpk2(score2, s0, s1);
    *spart_slot = s0 + s1;
  }
}

template <int CH>
__global__ void __launch_bounds__(kNumThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_hs, const __grid_constant__ CUtensorMap tmap_hl,
                 const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_y,
                 const __grid_constant__ CUtensorMap tmap_hs_rem, const __grid_constant__ CUtensorMap tmap_y_rem,
                 const __grid_constant__ CUtensorMap tmap_xres,
                 const Shape g, const Ptrs a) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TRACE(0);
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* hstage = smem + g.hs_off;
  float* sbias_all = reinterpret_cast<float*>(smem + g.tail_off);
  float* spart = sbias_all + kEpiWarps * (kBiasBytesPerWarp / 4);             // [2][128][kSpartPerRow], or empty
  uint32_t* s_words_all = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(spart) + g.spart_bytes);   // [16 warps][16]: tokens per warp x words
  unsigned int* s_hist = s_words_all + kEpiWarps * 16;                        // [g.hist_slots] (E rounded up to 32)
  Barriers* bars = reinterpret_cast<Barriers*>(s_hist + g.hist_slots);

  const int rm = static_cast<int>(tc::cluster_ctarank());
  const int p = static_cast<int>(blockIdx.x) >> 1, P = static_cast<int>(gridDim.x) >> 1;
  int* const ws_rec = a.sync + kSyncHeaderInts;                        // per block: [0] done, [1] ready

  if (warp == 0 && lane == 0) {
    tc::prefetch_tensormap(&tmap_x);
    tc::prefetch_tensormap(&tmap_w1);
    tc::prefetch_tensormap(&tmap_hl);
    tc::prefetch_tensormap(&tmap_w2);
  }
  if (warp == 1) {
    // one barrier per lane (the struct starts with its mbarriers, in declaration order); the separate phase-1 ring
    // and its hand-over barriers only when a schedule uses them
    constexpr int kCore = 2 * kMaxStages + 10;                 // full, empty, tmem_*, hs_*, route_*
    constexpr int kAll = kCore + 2 * kMaxStages + 3;           // + full_b, empty_b, a_full, a_empty, p1_done
    uint64_t* const all = reinterpret_cast<uint64_t*>(bars);
    const int n_init = g.sep_ring1 ? kAll : kCore;
    for (int i = lane; i < n_init; i += 32) {
      uint32_t count = 1;
      if (i == 2 * kMaxStages + 2 || i == 2 * kMaxStages + 3) count = 2 * kEpiWarps;   // tmem_empty: both CTAs' epilogues
      if (i == 2 * kMaxStages + 4 || i == 2 * kMaxStages + 5) count = kEpiWarps;       // hs_full
      if (i == 2 * kMaxStages + 9) count = kEpiWarps;                                  // route_done
      tc::mbar_init(all + i, count);
    }
    tc::fence_mbar_init();
  }
  if (warp == 2) {
    tc::tmem_alloc_2sm<kTmemCols>(&bars->tmem_base);
    if (lane == 0) {
      tc::prefetch_tensormap(&tmap_hs);
      tc::prefetch_tensormap(&tmap_y);
      tc::prefetch_tensormap(&tmap_hs_rem);
      tc::prefetch_tensormap(&tmap_y_rem);
    }
  }
  for (int i = threadIdx.x; i < g.hist_slots; i += kNumThreads) s_hist[i] = 0u;
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::cluster_sync_all();
  tc::fence_after_thread_sync();
  if (threadIdx.x == 0) TRACE(1);
  // The weights do not depend on the previous kernel: while this CTA waits for it (CTAs of this grid become resident
  // as the predecessor's CTAs exit, up to ~20 us before its last one), pull this CTA's share of W1 and W2 into L2.
  if (warp == 3 && lane == 0 && g.prefetch_weights) {
    const size_t n1 = static_cast<size_t>(2) * g.h * g.d * 2, n2 = static_cast<size_t>(g.d) * g.h * 2;
    const size_t piece = 16384;
    const size_t pieces1 = (n1 + piece - 1) / piece, pieces2 = (n2 + piece - 1) / piece;
    for (size_t i = blockIdx.x; i < pieces1 + pieces2; i += gridDim.x) {
      const bool first = i < pieces1;
      const size_t off = (first ? i : i - pieces1) * piece;
      const size_t left = (first ? n1 : n2) - off;
      const uint32_t bytes = static_cast<uint32_t>(left < piece ? left : piece) & ~15u;
      const char* src = static_cast<const char*>(first ? a.w1 : a.w2) + off;
      if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
    }
  }
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail.  Only the warps whose FIRST
  // access to memory the predecessor may touch is not already ordered behind another warp's wait execute
  // griddepcontrol.wait: the A-tile producer (x, later the sync area) and the sync warp (sync area; a pair without
  // phase-1 tiles polls it straight away).  The B-tile producer only ever reads the weights and starts filling the
  // ring at once; the MMA thread, the epilogue warps and their stores follow the first x tile through mbarriers.
  if (warp == 0 || warp == 2) pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (threadIdx.x == 0) TRACE(2);
  const uint32_t tmem_base = bars->tmem_base;

  const uint32_t bytes1 = static_cast<uint32_t>(g.ks1 * (kABytes + g.nv * 128));
  const uint32_t bytes3 = static_cast<uint32_t>(g.ks3 * (kABytes + (g.bn / 2) * 128));

  if (warp == 0 || warp == 3) {
    // ================================================================== TMA producers (A: warp 0, B: warp 3)
    // both operands of both CTAs complete on the leader's full[s]
    const bool do_a = warp == 0;
    int s = 0;
    uint32_t ph = 0;
    Item t;
    const bool ares = g.a_resident != 0;
    for (int phase = 0; phase < 2; ++phase) {
      const int ks = phase == 0 ? g.ks1 : g.ks3;
      // resident-A mode: phase 1 has its own ring (B only, behind the resident panels) and barriers; phase 3 starts
      // from a fresh ring over the whole area once every phase-1 MMA has completed
      const bool res1 = ares && phase == 0;
      const bool sep1 = g.sep_ring1 != 0 && phase == 0;
      uint64_t* const full_bar = sep1 ? bars->full_b : bars->full;
      uint64_t* const empty_bar = sep1 ? bars->empty_b : bars->empty;
      const int n_stages = sep1 ? g.stages1 : g.stages;
      const int slot_bytes = sep1 ? g.slot1_bytes : g.slot_bytes;
      uint8_t* const ring = sep1 ? smem + g.ring1_off : smem;
      const uint32_t full_leader0 = tc::mapa_u32(full_bar, 0);
      if (g.sep_ring1 != 0 && phase == 1) {
        tc::mbar_wait(&bars->p1_done, 0u);
        s = 0;
        ph = 0;
      }
      int cur_mp = -1, a_runs = 0;
      for (int it = 0; phase == 0 ? item1(g, it, p, P, rm, t) : item3(g, it, p, P, rm, t); ++it) {
#if MOE_TRACE
        if (!g.trace_p3 && g.split3 == 1 && do_a && lane == 0 && phase == 0 && it == 0) TRACE(48);   // ramp: first item known (slot shared with the split-K stamps)
#endif
        if (phase == 1 && do_a) {
          // the block's H rows are complete and routed once every routing chunk of the block has been counted
          if (lane == 0) {
#if MOE_TRACE
            if (it == 0) TRACE(62);
#endif
            poll_at_least(ws_rec + t.m_blk * kBlockRecInts + 1, g.chunks_per_block);
#if MOE_TRACE
            if (it == 0) TRACE(63);
            if (g.trace_p3 && it < 8) TRACE(48 + it);
#endif
          }
          __syncwarp();
          fence_proxy_async_global();   // generic-proxy writes of other SMs (acquired above) -> this thread's TMA reads
        }
        if (res1 && do_a) {
          // one load per run of tiles on the same row block: all k-blocks of this CTA's 128 x rows
          if ((t.m_blk >> 1) != cur_mp) {
            cur_mp = t.m_blk >> 1;
            if (a_runs > 0) tc::mbar_wait(&bars->a_empty, (a_runs - 1) & 1u);   // the previous block's MMAs are done
            if (tc::elect_one()) {
              if (rm == 0) tc::mbar_arrive_expect_tx(&bars->a_full, 2u * static_cast<uint32_t>(g.a_res_bytes));
              tc::tma_load_3d_2sm(smem, &tmap_xres, tc::mapa_u32(&bars->a_full, 0), 0, t.m_blk * kBlockM, 0);
#if MOE_TRACE
              if (it == 0) TRACE(7);
#endif
            }
            __syncwarp();
            ++a_runs;
          }
          continue;
        }
        for (int kb = t.kb_begin; kb < t.kb_end; kb += ks) {
#if MOE_TRACE
          if (!g.trace_p3 && do_a && lane == 0 && phase == 0 && it == 4 && kb == t.kb_begin && g.items1 > 5 * P) TRACE(54);
          if (!g.trace_p3 && do_a && lane == 0 && phase == 1 && it == 0 && kb == t.kb_begin && g.items1 <= 5 * P) TRACE(54);
#endif
          tc::mbar_wait(&empty_bar[s], ph ^ 1u);
#if MOE_TRACE
          if (!g.trace_p3 && g.split3 == 1 && do_a && lane == 0 && phase == 0 && it == 0 && kb == 0) TRACE(49);   // ramp: first slot wait passed
#endif
#if MOE_TRACE
          if (!g.trace_p3 && do_a && lane == 0 && phase == 0 && it == 4 && kb + ks >= t.kb_end && g.items1 > 5 * P) TRACE(55);
          if (!g.trace_p3 && do_a && lane == 0 && phase == 1 && it == 0 && kb + ks >= t.kb_end && g.items1 <= 5 * P) TRACE(55);
#endif
          uint8_t* sa = ring + s * slot_bytes;
          uint8_t* sb = res1 ? sa : sa + ks * kABytes;
          if (tc::elect_one()) {
            const uint32_t full_leader = full_leader0 + static_cast<uint32_t>(s) * 8u;
            if (phase == 0) {
              if (do_a) {
                if (rm == 0) tc::mbar_arrive_expect_tx(&full_bar[s], 2u * bytes1);
                tc::tma_load_3d_2sm(sa, &tmap_x, full_leader, 0, t.m_blk * kBlockM, kb);
#if MOE_TRACE
                if (it == 0 && kb == 0) TRACE(7);
#endif
              } else {
                if (res1 && rm == 0) tc::mbar_arrive_expect_tx(&full_bar[s], 2u * static_cast<uint32_t>(g.slot1_bytes));
                // CTA 0 of the pair stages the value rows, CTA 1 the gate rows of the tile's neurons
                tc::tma_load_3d_2sm(sb, &tmap_w1, full_leader, 0, (rm == 0 ? 0 : g.h) + t.n * g.nv, kb);
              }
            } else if (do_a) {
              if (rm == 0) tc::mbar_arrive_expect_tx(&full_bar[s], 2u * bytes3);
              tc::tma_load_3d_2sm(sa, &tmap_hl, full_leader, 0, t.m_blk * kBlockM, kb);
            } else {
              tc::tma_load_3d_2sm(sb, &tmap_w2, full_leader, 0, t.n * g.bn + rm * (g.bn / 2), kb);
            }
          }
          __syncwarp();
          if (++s == n_stages) {
            s = 0;
            ph ^= 1u;
          }
        }
        // (the visit count of the block's record -- an atomic with a returned value, one L2 round trip -- after the
        // item's loads have been issued, not between the ready-wait and the first load)
        if (phase == 1 && do_a && lane == 0) block_consumed(ws_rec + t.m_blk * kBlockRecInts, 2 * g.n_tiles3 * g.split3);
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
      const bool ares = g.a_resident != 0;
      for (int phase = 0; phase < 2; ++phase) {
        const int ks = phase == 0 ? g.ks1 : g.ks3;
        const uint32_t idesc = tc::umma_idesc_bf16_f32(2 * kBlockM, static_cast<uint32_t>(phase == 0 ? 2 * g.nv : g.bn));
        const uint32_t b_sub_bytes = static_cast<uint32_t>((phase == 0 ? g.nv : g.bn / 2) * 128);
        const bool res1 = ares && phase == 0;
        const bool sep1 = g.sep_ring1 != 0 && phase == 0;
        uint64_t* const full_bar = sep1 ? bars->full_b : bars->full;
        uint64_t* const empty_bar = sep1 ? bars->empty_b : bars->empty;
        const int n_stages = sep1 ? g.stages1 : g.stages;
        const int slot_bytes = sep1 ? g.slot1_bytes : g.slot_bytes;
        const uint32_t ring = tc::smem_u32(smem) + (sep1 ? static_cast<uint32_t>(g.ring1_off) : 0u);
        const uint32_t a_res = tc::smem_u32(smem);
        if (g.sep_ring1 != 0 && phase == 1) {
          s = 0;
          ph = 0;
        }
        int cur_mp = -1, a_runs = 0;
        Tile1Iter it1;
        it1.init(g, p, P);
        for (int it = 0;; ++it, ++acc_it) {
          bool have;
          if (phase == 0) {
            have = it1.valid();
            if (have) {
              t.m_blk = 2 * it1.mp + rm;
              t.n = it1.n;
              t.kb_begin = 0;
              t.kb_end = g.nkb1;
              t.slice = 0;
              it1.next();          // (it1 now describes the pair's NEXT tile)
            }
          } else {
            have = item3(g, it, p, P, rm, t);
          }
          if (!have) break;
          const int as = acc_it & 1;
          bool last_of_run = false;
          if (res1) {
            if ((t.m_blk >> 1) != cur_mp) {
              cur_mp = t.m_blk >> 1;
              tc::mbar_wait(&bars->a_full, a_runs & 1u);
              ++a_runs;
            }
            last_of_run = !it1.valid() || it1.mp != cur_mp;
          }
#if MOE_TRACE
          if (!g.trace_p3 && lane == 0 && acc_it == 4 && g.items1 > 5 * P) TRACE(50);
          if (!g.trace_p3 && lane == 0 && phase == 1 && it == 0 && g.items1 <= 5 * P) TRACE(50);
#endif
          tc::mbar_wait(&bars->tmem_empty[as], ((acc_it >> 1) & 1u) ^ 1u);
          tc::fence_after_thread_sync();
#if MOE_TRACE
          if (!g.trace_p3 && lane == 0 && acc_it == 4 && g.items1 > 5 * P) TRACE(51);
          if (!g.trace_p3 && lane == 0 && phase == 1 && it == 0 && g.items1 <= 5 * P) TRACE(51);
#endif
          const uint32_t d_tmem = tb + as * kAccStride;
          for (int kb = t.kb_begin; kb < t.kb_end; kb += ks) {
            tc::mbar_wait(&full_bar[s], ph);
            tc::fence_after_thread_sync();
#if MOE_TRACE
            if (!g.trace_p3 && lane == 0 && acc_it == 4 && g.items1 > 5 * P) TRACE(kb == t.kb_begin ? 52 : 53);
            if (!g.trace_p3 && lane == 0 && phase == 1 && it == 0 && g.items1 <= 5 * P) TRACE(kb == t.kb_begin ? 52 : 53);
#endif
            const uint32_t slot = ring + s * slot_bytes;
            const uint32_t a_base = res1 ? a_res + static_cast<uint32_t>(kb) * kABytes : slot;
            const uint32_t b_base = res1 ? slot : slot + ks * kABytes;
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
              tc::umma_commit_2sm_mc(&empty_bar[s], 0x3);   // frees the slot in both CTAs of the pair
            }
            __syncwarp();
            if (++s == n_stages) {
              s = 0;
              ph ^= 1u;
            }
          }
          if (tc::elect_one()) {
            tc::umma_commit_2sm_mc(&bars->tmem_full[as], 0x3);   // accumulator complete -> both epilogues
            if (last_of_run) tc::umma_commit_2sm_mc(&bars->a_empty, 0x3);   // the resident panels may be replaced
#if MOE_TRACE
            {
              const int tix = g.trace_p3 ? (phase == 1 ? it : 99) : acc_it;
              if (tix < 8) TRACE(8 + 4 * tix);
            }
#endif
          }
          __syncwarp();
        }
        if (sep1 && tc::elect_one()) tc::umma_commit_2sm_mc(&bars->p1_done, 0x3);
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ================================================================== store / sync warp
    {
      // ---- phase 1 (whole warp): per tile, lane 0 issues the panel stores, releases the staging buffer as soon as
      // the stores have READ it, and then publishes the PREVIOUS tile (whose stores have fully completed by then:
      // wait_group 1).  Expert scores are written by the epilogue threads themselves (no barrier among the epilogue
      // warps, which therefore drift apart and overlap each other's TMEM reads and math); only experts that span
      // column groups are combined here from shared-memory partial sums.
      int st_it = 0;
      int use_p1[2] = {0, 0};
      // Publication of finished tiles (done[m] += 1) is BATCHED: a GPU-scope release costs ~1 us (MEMBAR.ALL.GPU waits
      // for every outstanding store of the SM), and paid once per tile it made this warp the pacemaker of phase 1
      // (2.5 - 2.7 us per tile against 2.2 us of tensor work at UNet batch 16: the in-kernel timeline showed the
      // "stored + signalled" stamps falling 0.4 us further behind with every tile).  Nobody consumes done[m] before
      // the block's LAST tile is in, so up to g.pub_batch - 1 finished tiles wait for one common fence.
      constexpr int kPubBatch = 8;
      int pend[kPubBatch] = {-1, -1, -1, -1, -1, -1, -1, -1};
      int n_pend = 0;
      Item t;
      for (int it = 0; item1(g, it, p, P, rm, t); ++it, ++st_it) {
        const int buf = st_it & 1;
        tc::mbar_wait(&bars->hs_full[buf], use_p1[buf] & 1u);
#if MOE_TRACE
        if (!g.trace_p3 && lane == 0 && it >= 2 && it < 4) TRACE(40 + 4 * (it - 2));
#endif
        if (lane == 0 && !g.direct_h) {   // (direct-H mode: the epilogue threads have already stored the tile)
          const uint8_t* src = hstage + buf * g.hs_bytes;
          const int n_full = g.nv >> 6;
          for (int pn = 0; pn < n_full; ++pn)
            tc::tma_store_2d(&tmap_hs, src + pn * (kBlockM * 128), t.n * g.nv + 64 * pn, t.m_blk * kBlockM);
          if (g.nv & 63) tc::tma_store_2d(&tmap_hs_rem, src + n_full * (kBlockM * 128), t.n * g.nv + 64 * n_full, t.m_blk * kBlockM);
          tc::tma_store_commit();
        }
        // experts that span column groups (es > nv / 4): combine the groups' partial sums from shared memory
        if (g.chunks_per_expert == 0) {
          const float* sp = spart + buf * kBlockM * kSpartPerRow;
          for (int r = lane; r < kBlockM; r += 32) {
            const int row = t.m_blk * kBlockM + r;
            if (row >= g.T) break;
            float* dst = a.scores + static_cast<size_t>(row) * g.E + t.n * g.experts_per_tile;
            const float* src = sp + r * kSpartPerRow;
            for (int e = 0; e < g.experts_per_tile; ++e) {
              float tot = 0.f;
              for (int j = 0; j < g.span; ++j) tot += src[e * g.span + j];
              dst[e] = tot;
            }
          }
          __threadfence();
        }
        __syncwarp();
        if (lane == 0) {
          tc::tma_store_wait_read<0>();
          tc::mbar_arrive(&bars->hs_empty[buf]);   // the epilogue warps may refill the buffer
#if MOE_TRACE
          if (!g.trace_p3 && it >= 2 && it < 4) TRACE(41 + 4 * (it - 2));
#endif
          Item nx;
          if (n_pend >= g.pub_batch - 1 && item1(g, it + 1, p, P, rm, nx)) {
            // the batch is full (and this is not the pair's last tile, whose publication follows the loop): every store
            // group but the one just committed is globally written -- one fence, then the counts
            tc::tma_store_wait<1>();
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
#pragma unroll
            for (int i = 0; i < kPubBatch - 1; ++i)
              if (i < n_pend) atomicAdd(ws_rec + pend[i] * kBlockRecInts, 1);
            n_pend = 0;
#if MOE_TRACE
            if (!g.trace_p3 && it >= 2 && it < 4) TRACE(42 + 4 * (it - 2));
            if (!g.trace_p3 && st_it - 1 < 8) TRACE(8 + 4 * (st_it - 1) + 3);
            if (!g.trace_p3 && it >= 2 && it < 4) TRACE(43 + 4 * (it - 2));
#endif
          }
#pragma unroll
          for (int i = 0; i < kPubBatch; ++i)
            if (i == n_pend) pend[i] = t.m_blk;
          ++n_pend;
        }
        ++use_p1[buf];
        __syncwarp();
      }
      __syncwarp();   // (span mode: the lanes' score stores above are already fenced)
      if (lane == 0 && n_pend > 0) {
        // the critical path of the phase: one wait for the outstanding stores, ONE release fence, then the counts of
        // the pair's last (up to kPubBatch) tiles
        tc::tma_store_wait<0>();
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
#pragma unroll
        for (int i = 0; i < kPubBatch; ++i)
          if (i < n_pend) atomicAdd(ws_rec + pend[i] * kBlockRecInts, 1);
#if MOE_TRACE
        if (!g.trace_p3 && st_it - 1 < 8) TRACE(8 + 4 * (st_it - 1) + 3);
#endif
      }
      __syncwarp();
    }
    if (lane == 0) {
      int use[2] = {0, 0};
      {   // staging-buffer uses of phase 1 (same count as above)
        Item tt;
        for (int it = 0; item1(g, it, p, P, rm, tt); ++it) ++use[it & 1];
      }
      uint32_t r_posted = 0;
      Item t;
      const int consumers = g.n_tiles3 * g.split3;      // phase-3 items per row block
      // Routing hand-shake of ONE phase-3 item: wait until every phase-1 tile of its block is published, let the
      // epilogue warps route this item's share of the block (consumer r of the block's `consumers` items takes chunks
      // r, r + consumers, ...), then count the share for the block's consumers.  It runs ONE ITEM AHEAD of the item's
      // tensor work: while the MMA thread is in item i's main loop the epilogue warps route item i + 1, so with several
      // items per pair (UNet batch 16: seven) a pair no longer alternates 6 us of routing with 6 us of MMAs.  (An item
      // waits only for phase-1 tiles and for the routing of its own block, so routing ahead cannot deadlock.)
      auto handshake = [&](const Item& ti) {
        poll_at_least(ws_rec + ti.m_blk * kBlockRecInts, g.n_tiles1);
        tc::mbar_arrive(&bars->route_req);
        tc::mbar_wait(&bars->route_done, r_posted & 1u);
        ++r_posted;
        int mine = 0;
        for (int c = ti.n * g.split3 + ti.slice; c < g.chunks_per_block; c += consumers) ++mine;
        // release at GPU scope: the epilogue threads' labels / zero-writes were ordered before route_done (mbarrier
        // arrive = release.cta, wait = acquire.cta), the release is cumulative over them
        if (mine > 0) red_release_add(ws_rec + ti.m_blk * kBlockRecInts + 1, mine);
        block_consumed(ws_rec + ti.m_blk * kBlockRecInts, 2 * consumers);
      };
      if (item3(g, 0, p, P, rm, t)) handshake(t);
      for (int it = 0; item3(g, it, p, P, rm, t); ++it) {
        Item nx;
        if (item3(g, it + 1, p, P, rm, nx)) handshake(nx);
        if (g.split3 == 1) {
          tc::mbar_wait(&bars->hs_full[0], use[0] & 1u);
          {   // clipped at T rows / d columns
            const int n_full = g.bn >> 6;
            for (int pn = 0; pn < n_full; ++pn)
              if (t.n * g.bn + 64 * pn < g.d)
                tc::tma_store_2d(&tmap_y, hstage + pn * (kBlockM * 128), t.n * g.bn + 64 * pn, t.m_blk * kBlockM);
            if ((g.bn & 63) && t.n * g.bn + 64 * n_full < g.d)
              tc::tma_store_2d(&tmap_y_rem, hstage + n_full * (kBlockM * 128), t.n * g.bn + 64 * n_full, t.m_blk * kBlockM);
          }
          tc::tma_store_commit();
          tc::tma_store_wait_read<0>();
          tc::mbar_arrive(&bars->hs_empty[0]);
          ++use[0];
        }
      }
      tc::tma_store_wait_read<0>();   // shared memory must outlive the reads; the writes complete with the kernel
    }
    __syncwarp();
  } else {
    // ================================================================== epilogue warps
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int cg = ew >> 2;            // column group 0..3
    float* sbias = sbias_all + ew * (kBiasBytesPerWarp / 4);
    uint32_t* s_words = s_words_all + ew * 16;
    const int q_row = 32 * q + lane;
    int acc_it = 0;
    int use0 = 0, use1 = 0;     // uses of the two staging buffers so far
    Item t;
    const StageLayout hl1 = stage_layout(g.nv), hl3 = stage_layout(g.bn);
    // ---------------------------------------------------------------- phase 1
    {
      const int cpg = g.nv / 4;
      const int col0 = cg * cpg;
      // per-thread constants of the loop: byte offsets of this thread's 4-column pieces inside a staging buffer
      // (row q_row, columns col0 ...) and whether the column group starts on an 8-column boundary
      uint32_t piece_off[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) piece_off[i] = (4 * i < cpg) ? stage_addr(0u, hl1, q_row, col0 + 4 * i) : 0u;
      const bool aligned = ((col0 | CH) & 7) == 0 || (col0 & 7) == 0;
      // tile-invariant pieces of the epilogue's addresses: the group's first expert inside a tile, this row's score
      // slot relative to the tile, the TMEM columns of the group
      const int cpe = g.chunks_per_expert;
      const int e_slot0 = (cpe > 0) ? cg * (cpg / g.es) : cg;
      const int act = g.act;
      const int n_tiles1 = g.n_tiles1, step_m1 = g.step_m1, step_n1 = g.step_n1;
      int i1, n_items1, i_stride;       // this pair's tile indices: i1, i1 + i_stride, ... < n_items1
      range1(g, p, P, i1, n_items1, i_stride);
      const long long score_row_stride = static_cast<long long>(kBlockM) * g.E;
      float* const score_row0 = a.scores + static_cast<long long>(rm * kBlockM + q_row) * g.E + e_slot0;
      const int experts_per_tile = g.experts_per_tile;
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + col0;
      const uint32_t hbase0 = tc::smem_u32(hstage);
      const uint32_t hs_bytes = g.hs_bytes;
      const int nv = g.nv;
      const bool direct_h = g.direct_h != 0;
      const long long h_row_stride = static_cast<long long>(kBlockM) * g.h;
      __nv_bfloat16* const h_row0 = a.H + static_cast<long long>(rm * kBlockM + q_row) * g.h + col0;
      const int rows_left0 = g.T - (rm * kBlockM + q_row);   // this row is inside the matrix iff rows_left0 > 256 mp
      // tile coordinates advance by a fixed step per iteration: no divisions in the loop
      int mp1 = i1 / n_tiles1, n1 = i1 - (i1 / n_tiles1) * n_tiles1;
      // bias slices: the first tile's are staged up front, every later tile's are fetched into registers while
      // the previous tile is being processed (a staged load per tile exposed ~0.8 us of global latency each time)
      if (i1 < n_items1)
        stage_bias(sbias, a.b1 != nullptr ? a.b1 + n1 * nv + col0 : nullptr,
                   a.b1 != nullptr ? a.b1 + g.h + n1 * nv + col0 : nullptr, cpg, lane, true);
      int it = 0;
      for (; i1 < n_items1; ++it, ++acc_it) {
        const int mp_cur = mp1, n_cur = n1;
        // next tile of this pair
        i1 += i_stride;
        mp1 += step_m1;
        n1 += step_n1;
        if (n1 >= n_tiles1) {
          n1 -= n_tiles1;
          ++mp1;
        }
        const int as = acc_it & 1;
        const int buf = it & 1;
        const int ub = it >> 1;          // earlier uses of this staging buffer
        float nb0 = 0.f, nb2 = 0.f;
        const bool has_next = i1 < n_items1;
        if (has_next && a.b1 != nullptr && lane < cpg) {
          const float* bv = a.b1 + n1 * nv + col0;
          nb0 = __ldg(bv + lane);
          nb2 = __ldg(bv + g.h + lane);
        }
        // this tile's addresses first: their dependent integer chains complete under the barrier waits below
        const uint32_t taddr = taddr0 + as * kAccStride;
        const uint32_t hbase = hbase0 + buf * hs_bytes;
        float* spart_slot = spart + (buf * kBlockM + q_row) * kSpartPerRow + cg;
        float* score_dst = score_row0 + 2 * mp_cur * score_row_stride + n_cur * experts_per_tile;
        const bool score_ok = rows_left0 > 2 * kBlockM * mp_cur;
        __nv_bfloat16* hrow = h_row0 + 2 * mp_cur * h_row_stride + n_cur * nv;   // direct-H mode only
        if (ub > 0) tc::mbar_wait(&bars->hs_empty[buf], (ub - 1) & 1u);   // staging buffer drained
        tc::mbar_wait(&bars->tmem_full[as], (acc_it >> 1) & 1u);
        tc::fence_after_thread_sync();
#if MOE_TRACE
        if (!g.trace_p3 && ew == 0 && lane == 0 && acc_it < 8) TRACE(8 + 4 * acc_it + 1);
#endif
        // (the alignment of the group's staging stores is warp-uniform: two instantiations, one uniform branch)
#define MOE_GEGLU_CALL(ACT_, AL_, DIR_) \
  geglu_group<CH, ACT_, AL_, DIR_>(taddr, taddr + nv, sbias, hbase, piece_off, hrow, spart_slot, score_dst, score_ok, cpg, cpe)
#define MOE_GEGLU_GROUP(ACT_)              \
  do {                                     \
    if (direct_h) {                        \
      if (aligned)                         \
        MOE_GEGLU_CALL(ACT_, true, true);  \
      else                                 \
        MOE_GEGLU_CALL(ACT_, false, true); \
    } else if (aligned) {                  \
      MOE_GEGLU_CALL(ACT_, true, false);   \
    } else {                               \
      MOE_GEGLU_CALL(ACT_, false, false);  \
    }                                      \
  } while (0)
        if (act == MOE_ACT_GELU) MOE_GEGLU_GROUP(MOE_ACT_GELU);
#if MOE_TRACE
        else if (act == 2) MOE_GEGLU_GROUP(2);
#endif
        else MOE_GEGLU_GROUP(MOE_ACT_RELU);
#undef MOE_GEGLU_CALL
#undef MOE_GEGLU_GROUP
        // accumulator stage drained -> the leader's MMA thread may overwrite it
        tc::fence_before_thread_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_cluster(tc::mapa_u32(&bars->tmem_empty[as], 0));
        if (has_next) {   // this warp is done reading the current slices
          sbias[lane] = nb0;
          sbias[kBiasGateOff + lane] = nb2;
        }
        // H tile: generic-proxy smem writes -> async proxy; the sync warp stores it and publishes the tile
        if (!direct_h) tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->hs_full[buf]);
#if MOE_TRACE
        if (!g.trace_p3 && ew == 0 && lane == 0 && acc_it < 8) TRACE(8 + 4 * acc_it + 2);
#endif
      }
      use0 = (it + 1) >> 1;
      use1 = it >> 1;
    }
    // ---------------------------------------------------------------- phase 3
    {
      const int cpg = g.bn / 4;
      const int col0 = cg * cpg;
      // ---- routing role, one item ahead of the item's epilogue (see the sync warp): once the sync warp has seen every
      // phase-1 tile of the block, route this item's share of the block's chunks
      auto route_item = [&](const Item& ti, int idx) {
        tc::mbar_wait(&bars->route_req, idx & 1u);
#if MOE_TRACE
        if (ew == 0 && lane == 0 && idx == 0) TRACE(60);
        if (g.trace_p3 && ew == 0 && lane == 0 && idx < 8) TRACE(40 + idx);
#endif
        // per chunk: select -> zero-writes -> (after the item's last chunk) signal -> labels / histogram.  The signal
        // follows a __syncwarp: every lane's zero-writes precede lane 0's arrive (release at CTA scope; the sync warp's
        // red.release then publishes them device-wide).  Labels and counters are kernel outputs nobody waits for.
        const int consumers = g.n_tiles3 * g.split3;
        const int c0 = ti.n * g.split3 + ti.slice;
        auto signal = [&] {
          if (lane == 0) tc::mbar_arrive(&bars->route_done);
#if MOE_TRACE
          if (ew == 0 && lane == 0) TRACE(61);
          if (g.trace_p3 && ew == 0 && lane == 0 && idx < 8) TRACE(8 + 4 * idx + 3);
#endif
        };
        if (c0 >= g.chunks_per_block) {       // more consumers than chunks: nothing to route for this item
          __syncwarp();
          signal();
        }
        for (int c = c0; c < g.chunks_per_block; c += consumers)
          route::route_dispatch_ordered(g, a, ti.m_blk * kBlockM + c * g.chunk_tokens, min(g.T, (ti.m_blk + 1) * kBlockM), ew,
                                        lane, s_words, s_hist, c + consumers >= g.chunks_per_block, signal);
      };
      if (item3(g, 0, p, P, rm, t)) route_item(t, 0);
      for (int it = 0; item3(g, it, p, P, rm, t); ++it, ++acc_it) {
        const int as = acc_it & 1;
        {
          Item nx;
          if (item3(g, it + 1, p, P, rm, nx)) route_item(nx, it + 1);
        }
        const int n0 = t.n * g.bn + col0;                  // first output column of this warp's group
        const int nvalid = max(0, min(cpg, g.d - n0));     // the last tile may overhang d
        stage_bias(sbias, a.b2 != nullptr ? a.b2 + n0 : nullptr, nullptr, nvalid, lane, false);
        if (g.split3 == 1) {
          // the whole staging area (both phase-1 buffers) holds one Y tile
          if (use0 > 0) tc::mbar_wait(&bars->hs_empty[0], (use0 - 1) & 1u);
          if (use1 > 0) tc::mbar_wait(&bars->hs_empty[1], (use1 - 1) & 1u);
        }
        tc::mbar_wait(&bars->tmem_full[as], (acc_it >> 1) & 1u);
        tc::fence_after_thread_sync();
#if MOE_TRACE
        if (ew == 0 && lane == 0 && (g.trace_p3 ? it : acc_it) < 8) TRACE(8 + 4 * (g.trace_p3 ? it : acc_it) + 1);
#endif
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + as * kAccStride + col0;
        const int row = t.m_blk * kBlockM + q_row;
        const bool row_ok = row < g.T;
        if (g.split3 == 1) {
          const uint32_t ybase = tc::smem_u32(hstage);
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
            stage_store<CH / 2>(ybase, hl3, q_row, col0 + c, yw);
          }
          tc::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_cluster(tc::mapa_u32(&bars->tmem_empty[as], 0));
          tc::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bars->hs_full[0]);
          ++use0;
        } else {
          // ---- split-K: park this slice's fp32 partial tile; the last slice to arrive reduces in slice order.
          // Partial tiles live in a tile-local layout [slice][tile][16-byte column chunk][row]: a thread owns a row, so a
          // warp's 32 stores of one chunk are 512 contiguous bytes (row-major partials were 32 half-filled sectors per
          // store: 3.5 us for a 40 KB tile), and the reducer reads them back with the row index fastest.
          const int cols4 = g.bn >> 2;
          const size_t tile4 = static_cast<size_t>(t.m_blk * g.n_tiles3 + t.n) * (cols4 * kBlockM);
          float4* part = reinterpret_cast<float4*>(a.split_partial) + t.slice * g.split_plane4 + tile4;
          for (int c = 0; c < cpg; c += CH) {
            uint32_t acc[CH];
            tc::tmem_ld_cols<CH>(taddr + c, acc);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < CH; i += 4)
              __stcg(part + ((col0 + c + i) >> 2) * kBlockM + q_row,
                     make_float4(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1]), __uint_as_float(acc[i + 2]),
                                 __uint_as_float(acc[i + 3])));
          }
          tc::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_cluster(tc::mapa_u32(&bars->tmem_empty[as], 0));
#if MOE_TRACE
          if (!g.trace_p3 && ew == 0 && lane == 0 && it == 0) TRACE(48);
#endif
          // all partial stores of the CTA, then ONE releasing increment: it publishes this slice's partial tile; the
          // last slice to arrive reads the other slices' tiles from L2 (ld.cg) after the barrier below.  (Letting every
          // slice wait for the others and reduce its share of the rows was measured 2 us SLOWER at d = 1280 / 512
          // tokens: all slices then end with the slowest one, plus a poll and an L2 round trip.)
          tc::named_bar_sync(1, kEpiThreads);
          if (ew == 0 && lane == 0) {
            int* counter = a.split_counters + t.m_blk * g.n_tiles3 + t.n;
            const int prev = atom_add_release(counter, 1);
            const int last = prev == g.split3 - 1;
            if (last) *counter = 0;
            bars->last_cta = last;
          }
          tc::named_bar_sync(1, kEpiThreads);
#if MOE_TRACE
          if (!g.trace_p3 && ew == 0 && lane == 0 && it == 0) TRACE(49);
#endif
          if (bars->last_cta) {
            // sum the slices (row index fastest: coalesced 16-byte loads), add b2, and stage the bf16 tile in the
            // swizzled staging layout; one thread then TMA-stores it (clipped at T rows / d columns)
            if (use0 > 0) tc::mbar_wait(&bars->hs_empty[0], (use0 - 1) & 1u);   // phase-1 stores have left the buffers
            if (use1 > 0) tc::mbar_wait(&bars->hs_empty[1], (use1 - 1) & 1u);
            const uint32_t ybase = tc::smem_u32(hstage);
            const float4* src = reinterpret_cast<const float4*>(a.split_partial) + tile4;
            const int tile_col0 = t.n * g.bn;
            // five chunks per thread and trip, the first two slices of all of them in flight together (one L2 round
            // trip per trip instead of one per chunk); further slices are added in order afterwards
            constexpr int kU = 5;
            const int total = kBlockM * cols4;
            for (int e0 = ew * 32 + lane; e0 < total; e0 += kU * kEpiThreads) {
              float4 p0[kU], p1[kU];
#pragma unroll
              for (int u = 0; u < kU; ++u) {
                const int e = e0 + u * kEpiThreads;
                p0[u] = p1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e < total) {
                  const float4* q = src + (e >> 7) * kBlockM + (e & (kBlockM - 1));
                  p0[u] = __ldcg(q);
                  if (g.split3 > 1) p1[u] = __ldcg(q + g.split_plane4);
                }
              }
#pragma unroll
              for (int u = 0; u < kU; ++u) {
                const int e = e0 + u * kEpiThreads;
                const int gcol = tile_col0 + 4 * (e >> 7);
                float4 sum = (e < total && a.b2 != nullptr && gcol < g.d) ? __ldg(reinterpret_cast<const float4*>(a.b2 + gcol))
                                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                sum.x += p0[u].x; sum.y += p0[u].y; sum.z += p0[u].z; sum.w += p0[u].w;     // slice order: deterministic
                sum.x += p1[u].x; sum.y += p1[u].y; sum.z += p1[u].z; sum.w += p1[u].w;
                p0[u] = sum;
              }
              // further slices one at a time, the trip's chunks of a slice in flight together (one L2 round trip per
              // slice; a load-and-add per chunk and slice was 2 x kU dependent round trips at split 4)
              for (int sl = 2; sl < g.split3; ++sl) {
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                  const int e = e0 + u * kEpiThreads;
                  if (e < total) p1[u] = __ldcg(src + sl * g.split_plane4 + (e >> 7) * kBlockM + (e & (kBlockM - 1)));
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                  p0[u].x += p1[u].x; p0[u].y += p1[u].y; p0[u].z += p1[u].z; p0[u].w += p1[u].w;
                }
              }
#pragma unroll
              for (int u = 0; u < kU; ++u) {
                const int e = e0 + u * kEpiThreads;
                if (e >= total) break;
                tc::sts_b32x2(stage_addr(ybase, hl3, e & (kBlockM - 1), 4 * (e >> 7)), pack_bf16x2(p0[u].x, p0[u].y),
                              pack_bf16x2(p0[u].z, p0[u].w));
              }
            }
            tc::fence_proxy_async_smem();
            tc::named_bar_sync(1, kEpiThreads);
            if (ew == 0 && lane == 0) {
              const int n_full = g.bn >> 6;
              for (int pn = 0; pn < n_full; ++pn)
                if (tile_col0 + 64 * pn < g.d)
                  tc::tma_store_2d(&tmap_y, hstage + pn * (kBlockM * 128), tile_col0 + 64 * pn, t.m_blk * kBlockM);
              if ((g.bn & 63) && tile_col0 + 64 * n_full < g.d)
                tc::tma_store_2d(&tmap_y_rem, hstage + n_full * (kBlockM * 128), tile_col0 + 64 * n_full, t.m_blk * kBlockM);
              tc::tma_store_commit();
              tc::tma_store_wait_read<0>();   // the staging area is free again (and may be reused by the next item)
            }
          }
          tc::named_bar_sync(1, kEpiThreads);
        }
#if MOE_TRACE
        if (ew == 0 && lane == 0 && (g.trace_p3 ? it : acc_it) < 8) TRACE(8 + 4 * (g.trace_p3 ? it : acc_it) + 2);
#endif
      }
    }
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
  tc::cluster_sync_relaxed();   // nobody exits while the peer may still signal its smem (also a CTA barrier)
  tc::fence_after_thread_sync();
  if (warp == 2) tc::tmem_dealloc_2sm<kTmemCols>(tmem_base);
  if (threadIdx.x == 0) TRACE(6);
}

// ------------------------------------------------------------------------------------------ host side
static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

static int ensure_smem(const void* kfn) { return ensure_dynamic_smem(kfn, kSmemLimit); }   // per (kernel, device)

// The kernel's CTAs wait on one another, so all `clusters` 2-CTA clusters must be able to be resident at once on
// this device (full-size B200: 74 clusters on 148 SMs; fewer under an MPS active-thread limit or a green context).
// Asked once per (kernel, device, shared-memory size).
static int clusters_fit(const void* kfn, cudaLaunchConfig_t* cfg, int clusters) {
  struct Entry { const void* fn; int dev; size_t smem; int max_clusters; };
  static Entry cache[64];
  static std::atomic<int> n_cache{0};
  const int dev = current_device();
  const int n = n_cache.load(std::memory_order_acquire);
  int have = -1;
  for (int i = 0; i < n; ++i)
    if (cache[i].fn == kfn && cache[i].dev == dev && cache[i].smem == cfg->dynamicSmemBytes) have = cache[i].max_clusters;
  if (have < 0) {
    int mc = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&mc, kfn, cfg);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return fail(MOE_ERR_CUDA, "cudaOccupancyMaxActiveClusters: %s", cudaGetErrorString(e));
    }
    have = mc;
    const int slot = n_cache.load(std::memory_order_relaxed);
    if (slot < 64) {
      cache[slot] = Entry{kfn, dev, cfg->dynamicSmemBytes, mc};
      n_cache.store(slot + 1, std::memory_order_release);
    }
  }
  if (have < clusters)
    return fail(MOE_ERR_UNSUPPORTED_SHAPE,
                "moe_ffn_fused: only %d of the %d CTA pairs can be co-resident on this device (reduced-SM context?); "
                "use the unfused kernels", have, clusters);
  return MOE_OK;
}

// Two fused launches must never be resident together: each would hold SMs while spinning on CTAs of its own grid
// that cannot be scheduled.  Launches on ONE stream are ordered anyway (programmatic dependent launch only overlaps
// the prologue); a launch on a different stream of the same device first waits for an event recorded behind the
// previous fused launch.  Streams under capture are left alone (cross-stream event waits would be captured as graph
// edges to work outside the graph): a captured graph serialises its own fused nodes through its stream order.
struct FusedOrder {
  cudaEvent_t ev = nullptr;
  cudaStream_t last = nullptr;
  bool any = false;
};
static FusedOrder g_order[kMaxDevices];

static int order_before_launch(cudaStream_t st) {
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) {
    (void)cudaGetLastError();
    return MOE_OK;
  }
  if (cap != cudaStreamCaptureStatusNone) return MOE_OK;
  FusedOrder& o = g_order[current_device()];
  if (o.any && o.last != st && o.ev != nullptr) {
    cudaError_t e = cudaStreamWaitEvent(st, o.ev, 0);
    if (e != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_ffn_fused: cudaStreamWaitEvent: %s", cudaGetErrorString(e));
  }
  return MOE_OK;
}

static void order_after_launch(cudaStream_t st) {
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    (void)cudaGetLastError();
    return;
  }
  FusedOrder& o = g_order[current_device()];
  if (o.ev == nullptr && cudaEventCreateWithFlags(&o.ev, cudaEventDisableTiming) != cudaSuccess) {
    (void)cudaGetLastError();
    o.ev = nullptr;
    return;
  }
  if (cudaEventRecord(o.ev, st) == cudaSuccess) {
    o.last = st;
    o.any = true;
  } else {
    (void)cudaGetLastError();
  }
}

}  // namespace fused
}  // namespace moe

extern "C" {

size_t moe_ffn_fused_workspace_bytes(int T, int d, int h) {
  using namespace moe::fused;
  (void)h;
  // up to 8 K slices of fp32 partial tiles: whole 256-row units x whole column tiles (at most 256 wide)
  const size_t rows = (static_cast<size_t>(T > 0 ? T : 0) + 255) / 256 * 256;
  const size_t cols = static_cast<size_t>(d > 0 ? d : 0) + 256;   // >= ceil(d / bn) * bn for every tile width bn <= 256
  size_t want = static_cast<size_t>(8) * rows * cols * 4;
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
  MOE_REQUIRE(T >= 0 && d >= 8 && h >= 8 && E >= 1 && es >= 1 && k >= 0 && k <= E, MOE_ERR_INVALID_ARGUMENT,
              "moe_ffn_fused: bad sizes T=%d d=%d h=%d E=%d es=%d k=%d", T, d, h, E, es, k);
  if (T == 0) return MOE_OK;   // an empty shard: nothing to compute, and empty tensors have no storage to point at
  MOE_REQUIRE(x && w1p && w2p && H && scores && Y && workspace, MOE_ERR_INVALID_ARGUMENT,
              "moe_ffn_fused: NULL x / w1p / w2p / H / scores / Y / workspace");
  MOE_REQUIRE(static_cast<long long>(E) * es == h, MOE_ERR_INVALID_ARGUMENT, "moe_ffn_fused: E*es=%d*%d != h=%d", E, es, h);
  MOE_REQUIRE(act == MOE_ACT_GELU || act == MOE_ACT_RELU || (MOE_TRACE && act == 2), MOE_ERR_INVALID_ARGUMENT,
              "moe_ffn_fused: act=%d", act);
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
    if (cpg % 4 || cpg > 32) continue;
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
  g.step_m1 = (sm_count() / 2) / g.n_tiles1;
  g.step_n1 = (sm_count() / 2) % g.n_tiles1;
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
    const size_t per_slice = (static_cast<size_t>(T) + 255) / 256 * 256 * (static_cast<size_t>(d) + 256) * 4;
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
      const double cost = rounds * (kb_per * per_kblock + 2500.0) + (sp > 1 ? 3000.0 + 20.0 * sp * bn : 0.0);
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
  g.spart_bytes = (g.chunks_per_expert == 0) ? 2 * kBlockM * kSpartPerRow * 4 : 0;   // only experts that span column groups
  g.hist_slots = (E + 31) / 32 * 32;
  const int small = kEpiWarps * kBiasBytesPerWarp + g.spart_bytes + kEpiWarps * 16 * 4 + g.hist_slots * 4 +
                    static_cast<int>(sizeof(Barriers)) + 64;
  const int fixed = 1024 + 2 * g.hs_bytes + small;
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
  if (const char* e = getenv("MOE_FUSED_STAGES")) {   // experiments only: a shallower ring
    const int v = atoi(e);
    if (v >= 2 && v < g.stages) g.stages = v;
  }
  MOE_REQUIRE(g.stages >= 2, MOE_ERR_UNSUPPORTED_SHAPE, "moe_ffn_fused: tiles do not fit shared memory");
  g.ks1 = g.nkb1 >= 2 ? ks : 1;
  g.ks3 = ks;
  // ---- resident-A mode of phase 1 (experimental, MOE_FUSED_ARES=1): a pair takes a contiguous run of column tiles
  // and loads its 128 x d block of x once per run instead of once per tile, which cuts the L2 -> SM traffic of phase 1
  // from 26 KB to 10 KB per k-block and CTA.  Measured slower (d = 320: 39.5 vs 37.2 us per layer call): the panels
  // take 80 KB away from the W1 ring, and with less than two tiles of W1 in flight the ring's round trip
  // (MMA complete -> slot free -> TMA -> landed) paces the tiles.  Off by default; kept for parity tests of the path.
  g.hs_off = g.stages * g.slot_bytes;
  g.tail_off = g.hs_off + 2 * g.hs_bytes;
  g.sep_ring1 = 0;
  g.ring1_off = 0;
  g.direct_h = 0;
  g.a_resident = 0;
  g.a_res_bytes = g.nkb1 * kABytes;
  g.slot1_bytes = g.ks1 * nv * 128;
  {
    // resident panels + the W1 ring over the WHOLE area in front of the staging buffers
    const int region = (kSmemLimit - fixed) / 1024 * 1024;
    int ares_ks = g.ks1;
    if (const char* e = getenv("MOE_FUSED_ARES_KS")) ares_ks = atoi(e) == 1 ? 1 : g.ks1;
    // MOE_FUSED_ARES=2: whole-tile slots (one barrier round trip and one commit per tile, two tiles of W1 in flight)
    if (const char* e = getenv("MOE_FUSED_ARES")) {
      if (atoi(e) == 2) ares_ks = g.nkb1;
    }
    const int slot1 = ares_ks * nv * 128;
    int stages1 = (region - g.a_res_bytes) / slot1;
    if (stages1 > kMaxStages) stages1 = kMaxStages;
    bool on = false;
    if (const char* e = getenv("MOE_FUSED_ARES")) on = atoi(e) != 0;
    if (on && stages1 >= 2 && g.items1 >= P && g.nkb1 <= 8) {
      g.a_resident = 1;
      g.ks1 = ares_ks;
      g.slot1_bytes = slot1;
      g.stages1 = stages1;
      g.hs_off = region > g.stages * g.slot_bytes ? region : g.stages * g.slot_bytes;
      g.tail_off = g.hs_off + 2 * g.hs_bytes;
      g.step_m1 = 0;
      g.step_n1 = 1;
      g.sep_ring1 = 1;
      g.ring1_off = g.a_res_bytes;
    }
  }
  // ---- direct-H mode of phase 1 (experimental, MOE_FUSED_DIRECT=1): the epilogue threads store their H rows straight
  // to global memory and the 2 x 20 KB of staging become a fourth ring slot for phase 1 (phase 3 keeps the staging for
  // its Y tiles).  Measured: d = 1280 / 512 tokens 39.7 vs 40.2 us, d = 640 35.6 vs 33.8 us -- a deeper ring does not
  // buy what a shallower one costs (2 slots: +6 us), so the L2 -> SM delivery rate, not the ring depth, is the limit,
  // and the uncoalesced row stores lengthen the epilogue.  Off by default; kept for parity tests of the path.
  if (!g.a_resident) {
    int stages_d = (kSmemLimit - 1024 - small) / g.slot_bytes;
    if (stages_d > kMaxStages) stages_d = kMaxStages;
    bool on = false;
    if (const char* e = getenv("MOE_FUSED_DIRECT")) on = atoi(e) != 0 && stages_d >= g.stages;
    if (on) {
      g.direct_h = 1;
      g.sep_ring1 = 1;
      g.ring1_off = 0;
      g.stages1 = stages_d;
      g.slot1_bytes = g.slot_bytes;
      if (g.stages1 * g.slot1_bytes > g.tail_off) g.tail_off = g.stages1 * g.slot1_bytes;
    }
  }
  g.kb_per_slice3 = (g.nkb3 + g.split3 - 1) / g.split3;
  g.kb_per_slice3 = (g.kb_per_slice3 + g.ks3 - 1) / g.ks3 * g.ks3;
  while (g.split3 > 1 && (g.split3 - 1) * g.kb_per_slice3 >= g.nkb3) --g.split3;
  g.items3 = g.m_pairs * g.n_tiles3 * g.split3;
  g.split_plane4 = static_cast<long long>(2 * g.m_pairs) * g.n_tiles3 * (g.bn / 4) * kBlockM;

  // ---- routing geometry: L lanes per token (chunk = 512 / L tokens, L / 4 chunks per 128-row block).  The
  // block's `consumers` phase-3 items share its chunks, so few consumers want large chunks; each lane holds
  // kpt = ceil(E / L) <= 16 experts.
  {
    const int consumers = g.n_tiles3 * g.split3;
    // a consumer waits only for items with a smaller index (acyclic) as long as one round covers a block
    MOE_REQUIRE(consumers <= P, MOE_ERR_UNSUPPORTED_SHAPE, "moe_ffn_fused: %d down-projection items per row block > %d CTA pairs",
                consumers, P);
    // measured (profiles/r01_fused_lanes_sweep.log): 64-token chunks for the two-consumer d = 320 shape, otherwise
    // one token per warp (32 lanes), which spreads a block over as many consumers as possible
    int L = consumers >= 4 ? 32 : (consumers >= 2 ? 8 : 4);
    while (L < 32 && (E + L - 1) / L > 16) L <<= 1;
    if (const char* e = getenv("MOE_FUSED_LANES")) {
      const int v = atoi(e);
      if ((v == 4 || v == 8 || v == 16 || v == 32) && (E + v - 1) / v <= 16) L = v;
    }
    int kpt = 1;
    while (kpt * L < E) kpt <<= 1;
    MOE_REQUIRE(kpt <= 16, MOE_ERR_UNSUPPORTED_SHAPE, "moe_ffn_fused: E=%d too large", E);
    g.lanes = L;
    g.lanes_log2 = ilog2(L);
    g.kpt = kpt;
    // chunk size: a block's 128 tokens divided among (up to 32 of) its consumers, at least one warp's worth
    int share = 1;
    while (share * 2 <= consumers && share < 32) share <<= 1;
    const int tpw = 32 / L;
    int warps = (kBlockM / share) / tpw;
    warps = warps < 1 ? 1 : (warps > kEpiWarps ? kEpiWarps : warps);
    if (const char* e = getenv("MOE_FUSED_ROUTE_WARPS")) {
      const int v = atoi(e);
      if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) warps = v;
    }
    g.route_warps = warps;
    g.chunk_tokens = tpw * warps;
    g.chunks_per_block = kBlockM / g.chunk_tokens;
    g.words = (E + 31) / 32;
  }
  g.act = act;
  g.mask_h = mask_h;
  g.count_begin = count_begin;
  g.count_end = count_end;
  g.es_magic = static_cast<uint32_t>((0x100000000ull / static_cast<unsigned>(es)) + 1ull);
  g.trace_p3 = getenv("MOE_TRACE_P3") != nullptr ? 1 : 0;
  g.prefetch_weights = 1;
  if (const char* e = getenv("MOE_FUSED_PREFETCH")) g.prefetch_weights = atoi(e) != 0;
  // tiles published per release: pairs with few tiles publish each one (early finishers start routing the early blocks),
  // pairs with many amortise the fence (measured: profiles/r02_sweep_publication_batch.log)
  // H larger than L2 (UNet batch 16 at d = 320: 168 MB): phase 1 leaves only the most recently written row blocks in L2;
  // phase 3 then starts with those, so that the routing stage's partial-sector zero stores and the first H loads hit L2
  g.rev3 = static_cast<size_t>(T) * h * 2 > (static_cast<size_t>(96) << 20) ? 1 : 0;
  if (const char* e = getenv("MOE_FUSED_REV3")) g.rev3 = atoi(e) != 0;
  g.pub_batch = (g.items1 >= 6 * P) ? 4 : 2;
  if (const char* e = getenv("MOE_FUSED_PUB")) {
    const int v = atoi(e);
    if (v >= 2 && v <= 8) g.pub_batch = v;
  }

  if (getenv("MOE_DEBUG_PRINT"))
    fprintf(stderr,
            "[moe_ffn_fused] T=%d d=%d h=%d E=%d es=%d k=%d | nv=%d tiles1=%d ks1=%d | bn=%d tiles3=%d split=%d ks3=%d kb/slice=%d | "
            "stages=%d slot=%d resident-A=%d direct-H=%d (ring1 %d x %d) | route lanes=%d kpt=%d warps=%d chunks/block=%d | items %d + %d on %d pairs\n",
            T, d, h, E, es, k, g.nv, g.n_tiles1, g.ks1, g.bn, g.n_tiles3, g.split3, g.ks3, g.kb_per_slice3, g.stages,
            g.slot_bytes, g.a_resident, g.direct_h, g.stages1, g.slot1_bytes, g.lanes, g.kpt, g.route_warps, g.chunks_per_block, g.items1, g.items3, P);

  CUtensorMap tx, tw1, ths, thl, tw2, ty, ths_rem, ty_rem, txres;
  int rc;
  if ((rc = make_tmap_bf16_kblocks(&txres, x, static_cast<uint64_t>(T), static_cast<uint64_t>(d), kBlockM,
                                   static_cast<uint32_t>(g.a_resident ? g.nkb1 : 1))))
    return rc;
  if ((rc = make_tmap_bf16_kblocks(&tx, x, static_cast<uint64_t>(T), static_cast<uint64_t>(d), kBlockM, static_cast<uint32_t>(g.ks1)))) return rc;
  if ((rc = make_tmap_bf16_kblocks(&tw1, w1p, static_cast<uint64_t>(2) * h, static_cast<uint64_t>(d), static_cast<uint32_t>(nv),
                                   static_cast<uint32_t>(g.ks1))))
    return rc;
  // staging-tile stores: 64-column panels with the 128-byte swizzle + a remainder panel (see StageLayout)
  auto rem_swizzle = [](int cols) { return cols == 32 ? 64 : (cols == 16 ? 32 : 0); };
  if ((rc = make_tmap_bf16_2d_sw(&ths, H, static_cast<uint64_t>(T), static_cast<uint64_t>(h), kBlockM, nv >= 64 ? 64u : static_cast<uint32_t>(nv),
                                 nv >= 64 ? 128 : rem_swizzle(nv))))
    return rc;
  if ((rc = make_tmap_bf16_2d_sw(&ths_rem, H, static_cast<uint64_t>(T), static_cast<uint64_t>(h), kBlockM,
                                 (nv & 63) ? static_cast<uint32_t>(nv & 63) : 64u, (nv & 63) ? rem_swizzle(nv & 63) : 128)))
    return rc;
  if ((rc = make_tmap_bf16_kblocks(&thl, H, static_cast<uint64_t>(T), static_cast<uint64_t>(h), kBlockM, static_cast<uint32_t>(g.ks3)))) return rc;
  if ((rc = make_tmap_bf16_kblocks(&tw2, w2p, static_cast<uint64_t>(d), static_cast<uint64_t>(h), static_cast<uint32_t>(g.bn / 2),
                                   static_cast<uint32_t>(g.ks3))))
    return rc;
  if ((rc = make_tmap_bf16_2d_sw(&ty, Y, static_cast<uint64_t>(T), static_cast<uint64_t>(d), kBlockM, g.bn >= 64 ? 64u : static_cast<uint32_t>(g.bn),
                                 g.bn >= 64 ? 128 : rem_swizzle(g.bn))))
    return rc;
  if ((rc = make_tmap_bf16_2d_sw(&ty_rem, Y, static_cast<uint64_t>(T), static_cast<uint64_t>(d), kBlockM,
                                 (g.bn & 63) ? static_cast<uint32_t>(g.bn & 63) : 64u, (g.bn & 63) ? rem_swizzle(g.bn & 63) : 128)))
    return rc;

  Ptrs a = {};
  a.w1 = w1p;
  a.w2 = w2p;
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

  const size_t smem = 1024u + static_cast<size_t>(g.tail_off) + static_cast<size_t>(small);
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
    rc = clusters_fit(reinterpret_cast<const void*>(ffn_fused_kernel<CHV>), &cfg, P);               \
    if (rc) return rc;                                                                              \
    rc = order_before_launch(cfg.stream);                                                           \
    if (rc) return rc;                                                                              \
    le = cudaLaunchKernelEx(&cfg, ffn_fused_kernel<CHV>, tx, tw1, ths, thl, tw2, ty, ths_rem, ty_rem, txres, g, a); \
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
  order_after_launch(cfg.stream);
  return check_launch("moe_ffn_fused");
}

}  // extern "C"
