// K1 (GEGLU up-projection, fused activation / override / expert-score epilogue) and
// K3 (down-projection) for sm_100a: persistent, warp-specialised tcgen05 GEMMs.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count;
//                                 operand slices are MULTICAST across a thread-block cluster)
//   warp 1      : MMA issuer     (one thread issues tcgen05.mma, accumulators in TMEM, 2 stages)
//   warp 2      : TMEM allocator
//   warp 3      : idle
//   warps 4-19  : epilogue       (tcgen05.ld -> registers -> fused math -> smem/global), four column
//                                 groups x four TMEM lane quarters, overlapped with the next tile's MMAs
//
// Operands are bf16, both K-major: A = activations [rows, K], B = weights [N, K] (nn.Linear layout),
// so D[m, n] = sum_k A[m, k] B[n, k] needs no transposes.  One UMMA per 16-wide K step covers the
// whole tile width; for K1 the B tile is [value rows | gate rows] so that value and gate
// accumulators of the same neurons sit side by side in one TMEM stage.
//
// Clusters.  Every launch measured L2->SM bound with 128 x 160 tiles (panel traffic = 2 M N K
// (1/BM + 1/BN) bytes), so CTAs are grouped in clusters of cn x cm: the cn CTAs of a cluster row work
// on the same 128 token rows and each loads 1/cn of the A tile and multicasts it to the row; the cm
// CTAs of a cluster column work on the same weight rows and each loads 1/cm of the B tile and
// multicasts it to the column.  A smem slot may be overwritten only when every CTA that receives
// the sender's slices has consumed it, so the MMA thread's tcgen05.commit multicasts its "slot free"
// arrive to all of its senders and each empty barrier expects cn + cm - 1 arrivals.
#include "common.cuh"
#include "tcgen05.cuh"

#include <stdlib.h>
#include <utility>

namespace moe {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                      // 64 bf16 = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KiB
constexpr int kAccStride = 256;                  // TMEM columns between the two accumulator stages
constexpr int kTmemCols = 512;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 16;
constexpr int kColGroups = kEpiWarps / 4;        // column groups per TMEM lane quarter
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kNumThreads = kEpiWarp0 * 32 + kEpiThreads;   // 640
constexpr int kMaxStages = 8;
constexpr int kSmemLimit = 232448;               // 227 KiB opt-in dynamic shared memory
constexpr int kBiasSmemPerWarp = 128 * 4;        // 2 x 64 floats
constexpr int kEpiBarrierId = 1;

// profiling counters (MOE_DEBUG_MODE bit 16): per CTA, cycles the producer / MMA threads spend per step
__device__ unsigned long long g_dbg_counters[256 * 8];
// timeline trace (MOE_DEBUG_MODE bit 32): per CTA, 64 globaltimer stamps (ns) at fixed points of the kernel
__device__ unsigned long long g_trace[256 * 64];

struct PipeBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t masked[kMaxStages];   // weight-masked down-projection: the masker warp has edited the landed B tile
  uint32_t tmem_base;
};

struct GemmShape {
  int rows;          // M (tokens)
  int k;             // reduction length
  int m_tiles, n_tiles;
  int tile_n;        // UMMA N (accumulator columns per stage) = B rows per stage
  int half_rows;     // K1: nv (value rows, then gate rows); K3: tile_n
  int second_off;    // K1: global row offset of the gate half (h); K3: 0
  int stages, stage_bytes;
  int cn, cm;        // cluster: cn CTAs share the A tile, cm CTAs share the B tile
  int a_box_rows;    // 128 / cn
  int b_box_rows;    // rows per B TMA box issued by one CTA
  int b_boxes;       // B boxes one CTA issues per stage
  int m_ctiles, n_ctiles;
  int extra_smem;    // bytes after the stage ring (epilogue staging)
  int ks;            // 64-wide k-blocks per pipeline stage (one barrier round trip per ks k-blocks)
  int split_k;       // K3: number of K slices per output tile (1 = no split)
  int kb_per_slice;  // 64-wide k-blocks per slice (a multiple of ks unless it is the last slice)
  int pair;          // 1: cta_group::2 -- the two CTAs of a cluster pair share one 256 x tile_n UMMA
  int debug;         // MOE_DEBUG_MODE bits: 1 = no TMA loads, 2 = no MMAs, 4 = no epilogue math (profiling only)
};

__device__ __forceinline__ void trace_at(const GemmShape& g, int slot) {
  if ((g.debug & 32) && blockIdx.x < 256 && slot < 64) g_trace[blockIdx.x * 64 + slot] = tc::global_timer_ns();
}

// exact GELU x * Phi(x) with Phi from 0.5 * erfc(u / sqrt2) = 2^(P7(u)), u = min(|x|, 4.3 sqrt2) (the 1 / sqrt2 is
// folded into the coefficients): branch-free, one MUFU.EX2 + 8 FFMA; max abs error 4e-7 on [-3, 3] (fp32 rounding
// level; the libdevice erff path costs ~2x the instructions and diverges).  Coefficients: weighted minimax fit.
__device__ __forceinline__ float gelu_erf(float x) {
  const float t = fminf(fabsf(x), 6.081118318204309f);
  float q = 9.425939424545504e-06f;
  q = fmaf(q, t, -6.281803507590666e-05f);
  q = fmaf(q, t, -3.898592258337885e-04f);
  q = fmaf(q, t, 7.337003480643034e-03f);
  q = fmaf(q, t, -5.264890193939209e-02f);
  q = fmaf(q, t, -4.591682255268097e-01f);
  q = fmaf(q, t, -1.1511090993881226f);
  q = fmaf(q, t, -0.9999999403953552f);
  const float e = tc::ex2_approx(q);               // 0.5 * erfc(t)
  // x * Phi(x) with Phi = 1 - e (x >= 0) or e (x < 0)  ==  max(x, 0) - |x| * e
  return fmaf(-fabsf(x), e, fmaxf(x, 0.f));
}

template <int ACT>
__device__ __forceinline__ float activate(float x) {
  if constexpr (ACT == MOE_ACT_GELU)
    return gelu_erf(x);
  else
    return fmaxf(x, 0.f);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// store kWords 32-bit words with the widest vectors the (known) alignment allows
template <int kWords>
__device__ __forceinline__ void store_words(__nv_bfloat16* dst, const uint32_t* w) {
  if constexpr (kWords % 4 == 0) {
#pragma unroll
    for (int i = 0; i < kWords / 4; ++i)
      reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < kWords / 2; ++i) reinterpret_cast<uint2*>(dst)[i] = make_uint2(w[2 * i], w[2 * i + 1]);
  }
}

// same, to a shared-memory address
template <int kWords>
__device__ __forceinline__ void store_words_smem(uint32_t addr, const uint32_t* w) {
  if constexpr (kWords % 4 == 0) {
#pragma unroll
    for (int i = 0; i < kWords / 4; ++i) tc::sts_b32x4(addr + 16 * i, w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < kWords / 2; ++i) tc::sts_b32x2(addr + 8 * i, w[2 * i], w[2 * i + 1]);
  }
}

// ------------------------------------------------------------------------------------------
// Tile schedule: cluster c takes cluster tiles c, c + n_clusters, ...; inside a cluster tile the
// CTA of rank (rn, rm) owns CTA tile (m = mct * cm + rm, n = nct * cn + rn).
// ------------------------------------------------------------------------------------------
struct TileCoord {
  int m_blk, n_blk;
  int slice, kb_begin, kb_end;   // split-K: this work item's K slice, in 64-wide k-blocks
};
__device__ __forceinline__ bool tile_at(const GemmShape& g, int it, int rn, int rm, TileCoord& t) {
  const int csize = g.cn * g.cm;
  int ct = static_cast<int>(blockIdx.x) / csize + it * (static_cast<int>(gridDim.x) / csize);
  if (ct >= g.m_ctiles * g.n_ctiles * g.split_k) return false;
  t.slice = ct % g.split_k;       // slices of one tile run side by side on different CTAs
  ct /= g.split_k;
  t.m_blk = (ct / g.n_ctiles) * g.cm + rm;
  t.n_blk = (ct % g.n_ctiles) * g.cn + rn;
  const int num_kb = (g.k + kBlockK - 1) / kBlockK;
  t.kb_begin = t.slice * g.kb_per_slice;
  t.kb_end = min(num_kb, t.kb_begin + g.kb_per_slice);
  return true;
}

// ------------------------------------------------------------------------------------------
// Shared mainloop roles
// ------------------------------------------------------------------------------------------
// Two producer threads (warp 0: A tiles + barrier arming, warp 3: B tiles): one thread pays ~200
// cycles per TMA instruction, which alone would pace the ring slower than the tensor pipe drains it.
template <bool PAIR>
__device__ __forceinline__ void producer_loop(const CUtensorMap* tmap_a, const CUtensorMap* tmap_b, uint8_t* smem,
                                              PipeBarriers* bars, const GemmShape& g, int rn, int rm, bool do_a) {
  uint16_t row_mask = 0, col_mask = 0;   // CTAs sharing my A tile / my B tile
  for (int j = 0; j < g.cn; ++j) row_mask |= static_cast<uint16_t>(1u << (rm * g.cn + j));
  for (int i = 0; i < g.cm; ++i) col_mask |= static_cast<uint16_t>(1u << (i * g.cn + rn));
  const int b_slice_rows = g.tile_n / g.cm;
  constexpr bool pair = PAIR;
  int s = 0;
  uint32_t ph = 0;
  TileCoord t;
  const bool prof = (g.debug & 16) != 0 && do_a;
  long long c_wait = 0, c_issue = 0, n_iter = 0;
  if (do_a && (threadIdx.x & 31) == 0) trace_at(g, 60);
  for (int it = 0; tile_at(g, it, rn, rm, t); ++it) {
    for (int kb = t.kb_begin; kb < t.kb_end; kb += g.ks) {
      const long long t0 = prof ? clock64() : 0;
      tc::mbar_wait(&bars->empty[s], ph ^ 1u);
      if (do_a && (threadIdx.x & 31) == 0 && it == 0 && kb == t.kb_begin) trace_at(g, 61);
      const long long t1 = prof ? clock64() : 0;
      c_wait += t1 - t0;
      ++n_iter;
      uint8_t* sa = smem + s * g.stage_bytes;
      uint8_t* sb = sa + g.ks * kABytes;
      if (tc::elect_one()) {
      if (g.debug & 1) {   // profiling: pretend the stage arrived
        if (do_a && (!pair || rm == 0)) tc::mbar_arrive(&bars->full[s]);
      } else if constexpr (pair) {
        // the leader's barrier collects the bytes of both CTAs of the pair
        const uint32_t full_leader = tc::mapa_u32(&bars->full[s], 0);
        if (do_a) {
          if (it == 0 && kb == t.kb_begin) trace_at(g, 62);
          if (rm == 0) tc::mbar_arrive_expect_tx(&bars->full[s], static_cast<uint32_t>(2 * g.stage_bytes));
          if (g.ks > 1)
            tc::tma_load_3d_2sm(sa, tmap_a, full_leader, 0, t.m_blk * kBlockM, kb);
          else
            tc::tma_load_2d_2sm(sa, tmap_a, full_leader, kb * kBlockK, t.m_blk * kBlockM);
        } else {
          const int j0 = rm * b_slice_rows;   // this CTA's half of the tile's weight rows
          const int grow = (j0 < g.half_rows) ? t.n_blk * g.half_rows + j0
                                              : g.second_off + t.n_blk * g.half_rows + (j0 - g.half_rows);
          if (g.ks > 1)
            tc::tma_load_3d_2sm(sb, tmap_b, full_leader, 0, grow, kb);
          else
            tc::tma_load_2d_2sm(sb, tmap_b, full_leader, kb * kBlockK, grow);
        }
      } else if (do_a) {
        // (non-pair) this CTA receives the whole stage (its own slices + its peers' multicasts)
        tc::mbar_arrive_expect_tx(&bars->full[s], static_cast<uint32_t>(g.stage_bytes));
        const int a_row0 = rn * g.a_box_rows;
        if (g.cn > 1)
          tc::tma_load_2d_mc(sa + a_row0 * 128, tmap_a, &bars->full[s], kb * kBlockK, t.m_blk * kBlockM + a_row0, row_mask);
        else if (g.ks > 1)
          tc::tma_load_3d(sa, tmap_a, &bars->full[s], 0, t.m_blk * kBlockM, kb);
        else
          tc::tma_load_2d(sa, tmap_a, &bars->full[s], kb * kBlockK, t.m_blk * kBlockM);
      } else {
        for (int b = 0; b < g.b_boxes; ++b) {
          const int j0 = rm * b_slice_rows + b * g.b_box_rows;   // row inside the stage's B tile
          const int grow = (j0 < g.half_rows) ? t.n_blk * g.half_rows + j0
                                              : g.second_off + t.n_blk * g.half_rows + (j0 - g.half_rows);
          if (g.cm > 1)
            tc::tma_load_2d_mc(sb + j0 * 128, tmap_b, &bars->full[s], kb * kBlockK, grow, col_mask);
          else if (g.ks > 1)
            tc::tma_load_3d(sb, tmap_b, &bars->full[s], 0, grow, kb);   // ks > 1 implies a single B box
          else
            tc::tma_load_2d(sb + j0 * 128, tmap_b, &bars->full[s], kb * kBlockK, grow);
        }
      }
      if (do_a && it == 0 && kb == t.kb_begin) trace_at(g, 7);
      }
      __syncwarp();
      if (prof) c_issue += clock64() - t1;
      if (++s == g.stages) {
        s = 0;
        ph ^= 1u;
      }
    }
  }
  if (prof && blockIdx.x < 256 && (threadIdx.x & 31) == 0) {
    g_dbg_counters[blockIdx.x * 8 + 0] = c_wait;
    g_dbg_counters[blockIdx.x * 8 + 1] = c_issue;
    g_dbg_counters[blockIdx.x * 8 + 2] = n_iter;
  }
}

template <bool PAIR, bool MASK = false>
__device__ __forceinline__ void mma_loop(uint8_t* smem, PipeBarriers* bars, const GemmShape& g, uint32_t tmem_base,
                                         int rn, int rm) {
  constexpr bool pair = PAIR;
  const uint32_t idesc = tc::umma_idesc_bf16_f32(pair ? 2 * kBlockM : kBlockM, static_cast<uint32_t>(g.tile_n));
  uint16_t sender_mask = 0;   // every CTA that writes into my smem: my cluster row and column
  for (int j = 0; j < g.cn; ++j) sender_mask |= static_cast<uint16_t>(1u << (rm * g.cn + j));
  for (int i = 0; i < g.cm; ++i) sender_mask |= static_cast<uint16_t>(1u << (i * g.cn + rn));
  const bool clustered = !pair && g.cn * g.cm > 1;
  int s = 0;
  uint32_t ph = 0;
  TileCoord t;
  const bool prof = (g.debug & 16) != 0;
  long long c_acc = 0, c_full = 0, c_mma = 0, c_commit = 0;
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);   // warp-uniform for the compiler
  for (int it = 0; tile_at(g, it, rn, rm, t); ++it) {
    const int as = it & 1;
    const uint32_t aph = (it >> 1) & 1u;
    const long long ta = prof ? clock64() : 0;
    tc::mbar_wait(&bars->tmem_empty[as], aph ^ 1u);
    tc::fence_after_thread_sync();
    if (prof) c_acc += clock64() - ta;
    const uint32_t d_tmem = tmem_base + as * kAccStride;
    for (int kb = t.kb_begin; kb < t.kb_end; kb += g.ks) {
      const long long t0 = prof ? clock64() : 0;
      // (weight-masked down-projection: the stage is ready once the masker warp has zeroed the masked weights)
      tc::mbar_wait(MASK ? &bars->masked[s] : &bars->full[s], ph);
      tc::fence_after_thread_sync();
      const long long t1 = prof ? clock64() : 0;
      c_full += t1 - t0;
      const uint32_t a_base = tc::smem_u32(smem + s * g.stage_bytes);
      const uint32_t b_base = a_base + g.ks * kABytes;
      const uint32_t b_sub_bytes = static_cast<uint32_t>((pair ? g.tile_n / 2 : g.tile_n) * 128);
      const int n_sub = min(g.ks, t.kb_end - kb);   // a slice's last stage may be partly empty (zero-filled)
      const bool leader_lane = tc::elect_one();
      if (leader_lane) {
        if (it == 0 && kb == t.kb_begin) trace_at(g, 3);
        for (int sub = 0; sub < n_sub; ++sub) {
          const uint32_t a_addr = a_base + sub * kABytes;
          const uint32_t b_addr = b_base + sub * b_sub_bytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            if (g.debug & 2) break;   // profiling: no tensor work
            const uint64_t da = tc::umma_desc_kmajor_sw128(a_addr + k * kUmmaK * 2);
            const uint64_t db = tc::umma_desc_kmajor_sw128(b_addr + k * kUmmaK * 2);
            const uint32_t acc = (kb > t.kb_begin || sub != 0 || k != 0) ? 1u : 0u;
            if constexpr (pair)
              tc::umma_bf16_ss_2sm(d_tmem, da, db, idesc, acc);
            else
              tc::umma_bf16_ss(d_tmem, da, db, idesc, acc);
          }
        }
      }
      const long long t2 = prof ? clock64() : 0;
      c_mma += t2 - t1;
      // slot free once these MMAs have read it: tell every CTA that sends into this slot
      // (tcgen05.commit must come from the thread that issued the MMAs)
      if (!leader_lane) {
      } else if (g.debug & 8) {
        tc::mbar_arrive(&bars->empty[s]);   // profiling: plain arrive instead of tcgen05.commit
      } else if constexpr (pair) {
        tc::umma_commit_2sm_mc(&bars->empty[s], 0x3);   // frees the slot in both CTAs of the pair
      } else {
        if (clustered)
          tc::umma_commit_mc(&bars->empty[s], sender_mask);
        else
          tc::umma_commit(&bars->empty[s]);
      }
      __syncwarp();
      if (prof) c_commit += clock64() - t2;
      if (++s == g.stages) {
        s = 0;
        ph ^= 1u;
      }
    }
    if (tc::elect_one()) {
      if constexpr (pair)
        tc::umma_commit_2sm_mc(&bars->tmem_full[as], 0x3);   // accumulator complete -> both epilogues
      else
        tc::umma_commit(&bars->tmem_full[as]);               // accumulator complete -> epilogue
      trace_at(g, 8 + 4 * it);
    }
    __syncwarp();
  }
  if (prof && blockIdx.x < 256 && (threadIdx.x & 31) == 0) {
    g_dbg_counters[blockIdx.x * 8 + 3] = c_acc;
    g_dbg_counters[blockIdx.x * 8 + 4] = c_full;
    g_dbg_counters[blockIdx.x * 8 + 5] = c_mma;
    g_dbg_counters[blockIdx.x * 8 + 6] = c_commit;
  }
}

template <bool PAIR>
__device__ __forceinline__ PipeBarriers* setup_pipeline(uint8_t*& smem, uint8_t*& extra, const GemmShape& g,
                                                        const CUtensorMap* ta, const CUtensorMap* tb) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by POINTER arithmetic (an integer round-trip would lose the shared address space and
  // turn every later access into a generic LD / ST)
  if (threadIdx.x == 0) trace_at(g, 0);
  smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  extra = smem + g.stages * g.stage_bytes;
  PipeBarriers* bars = reinterpret_cast<PipeBarriers*>(extra + g.extra_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::prefetch_tensormap(ta);
    tc::prefetch_tensormap(tb);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < g.stages; ++i) {
      tc::mbar_init(&bars->full[i], 1);
      tc::mbar_init(&bars->empty[i], PAIR ? 1u : static_cast<uint32_t>(g.cn + g.cm - 1));
      tc::mbar_init(&bars->masked[i], kEpiWarps);   // one arrive per epilogue warp (the maskers)
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars->tmem_full[i], 1);
      tc::mbar_init(&bars->tmem_empty[i], PAIR ? 2 * kEpiWarps : kEpiWarps);   // pair: both CTAs' epilogues
    }
    tc::fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (PAIR)
      tc::tmem_alloc_2sm<kTmemCols>(&bars->tmem_base);
    else
      tc::tmem_alloc<kTmemCols>(&bars->tmem_base);
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (g.cn * g.cm > 1) tc::cluster_sync_all();   // peers' barriers exist before anything is multicast
  tc::fence_after_thread_sync();
  // everything above overlapped the previous kernel's tail (PDL); from here on we touch its outputs
  if (threadIdx.x == 0) trace_at(g, 1);
  pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (threadIdx.x == 0) trace_at(g, 2);
  return bars;
}

template <bool PAIR>
__device__ __forceinline__ void teardown_pipeline(PipeBarriers* bars, const GemmShape& g) {
  tc::fence_before_thread_sync();
  __syncthreads();
  if (threadIdx.x == 0) trace_at(g, 5);
  if (g.cn * g.cm > 1) tc::cluster_sync_all();   // nobody exits while a peer may still signal its smem
  tc::fence_after_thread_sync();
  if (threadIdx.x == 0) trace_at(g, 6);
  if ((threadIdx.x >> 5) == 2) {
    if constexpr (PAIR)
      tc::tmem_dealloc_2sm<kTmemCols>(bars->tmem_base);
    else
      tc::tmem_dealloc<kTmemCols>(bars->tmem_base);
  }
}

// per-warp bias slices staged in shared memory: lanes load coalesced, everyone re-reads float4 broadcasts
__device__ __forceinline__ void stage_bias(float* sb, const float* b0, const float* b1, int n, int lane) {
  __syncwarp();
  for (int i = lane; i < 64; i += 32) {
    sb[i] = (b0 != nullptr && i < n) ? __ldg(b0 + i) : 0.f;
    sb[64 + i] = (b1 != nullptr && i < n) ? __ldg(b1 + i) : 0.f;
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// K1: GEGLU up-projection
// ------------------------------------------------------------------------------------------
struct GegluArgs {
  const float* b1;                 // [2h] or null
  const uint8_t* neuron_override;  // [h] or null          (FEAT path)
  float override_value;
  float* scores;                   // [T, E] or null
  __nv_bfloat16* gate_out;         // [T, h] or null       (FEAT path)
  int h, E, es, nv;                // nv = neuron pairs per tile (tile_n = 2 nv)
  // host-precomputed so the epilogue never divides: experts per tile, chunks per expert inside a column
  // group (0 if an expert spans groups), groups per expert (1 if experts fit a group)
  int experts_per_tile, chunks_per_expert, span;
};

// One column group (cpg = nv / 4 neuron pairs) of one tile for the token row this thread owns.
// CH divides both cpg and es, so a chunk never straddles an expert.  `e_first` is the id of the first
// expert this group touches.  sbias / hrow / spart are shared-space addresses.
template <int CH, int ACT, bool FEAT>
__device__ __forceinline__ void geglu_epilogue_group(const GegluArgs& a, uint32_t taddr, const float* sbias,
                                                     __nv_bfloat16* hrow, float* spart, int row, bool row_ok, int n0,
                                                     int col0, int cpg, int e_first, int cg, int q) {
  float score = 0.f;
  int chunk_in_expert = 0, e = e_first;
  float* score_row = a.scores != nullptr ? a.scores + static_cast<size_t>(row) * a.E : nullptr;
  for (int c = 0; c < cpg; c += CH) {
    uint32_t v[CH], g[CH];
    tc::tmem_ld_cols<CH>(taddr + col0 + c, v);
    tc::tmem_ld_cols<CH>(taddr + a.nv + col0 + c, g);
    tc::tmem_ld_wait();
    uint32_t hw[CH / 2];
    uint32_t gw[FEAT ? CH / 2 : 1];
#pragma unroll
    for (int i = 0; i < CH; i += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(sbias + c + i);
      const float4 bg = *reinterpret_cast<const float4*>(sbias + 64 + c + i);
      float gg[4] = {__uint_as_float(g[i]) + bg.x, __uint_as_float(g[i + 1]) + bg.y, __uint_as_float(g[i + 2]) + bg.z,
                     __uint_as_float(g[i + 3]) + bg.w};
      const float vv[4] = {__uint_as_float(v[i]) + bv.x, __uint_as_float(v[i + 1]) + bv.y,
                           __uint_as_float(v[i + 2]) + bv.z, __uint_as_float(v[i + 3]) + bv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        gg[j] = activate<ACT>(gg[j]);
        if constexpr (FEAT) {
          if (a.neuron_override != nullptr && __ldg(a.neuron_override + n0 + c + i + j)) gg[j] = a.override_value;
        }
        score += gg[j];
      }
      hw[i / 2] = pack_bf16x2(vv[0] * gg[0], vv[1] * gg[1]);
      hw[i / 2 + 1] = pack_bf16x2(vv[2] * gg[2], vv[3] * gg[3]);
      if constexpr (FEAT) {
        gw[i / 2] = pack_bf16x2(gg[0], gg[1]);
        gw[i / 2 + 1] = pack_bf16x2(gg[2], gg[3]);
      }
    }
    store_words<CH / 2>(hrow + col0 + c, hw);   // staged in smem; one TMA store per tile writes it out
    if constexpr (FEAT) {
      if (a.gate_out != nullptr && row_ok)
        store_words<CH / 2>(a.gate_out + static_cast<size_t>(row) * a.h + n0 + c, gw);
    }
    // experts that fit inside the group: flush a finished expert's score
    if (a.chunks_per_expert > 0 && ++chunk_in_expert == a.chunks_per_expert) {
      if (score_row != nullptr && row_ok) score_row[e] = score;
      score = 0.f;
      chunk_in_expert = 0;
      ++e;
    }
  }
  if (a.chunks_per_expert == 0 && a.scores != nullptr) {
    // one expert spans `span` adjacent column groups: combine the partial sums through smem
    volatile float* slot = spart + (32 * q + (threadIdx.x & 31)) * kColGroups;
    slot[cg] = score;
    tc::named_bar_sync(2 + q, 4 * 32);   // the 4 warps of this TMEM lane quarter
    if ((cg % a.span) == 0 && row_ok) {
      float tot = 0.f;
      for (int j = 0; j < a.span; ++j) tot += slot[cg + j];
      score_row[e_first] = tot;
    }
  }
}

template <int CH, bool FEAT, bool PAIR>
__global__ void __launch_bounds__(kNumThreads, 1)
geglu_up_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                const __grid_constant__ CUtensorMap tmap_h, const GemmShape g, const GegluArgs a, const int act) {
  uint8_t *smem, *extra;
  PipeBarriers* bars = setup_pipeline<PAIR>(smem, extra, g, &tmap_x, &tmap_w1);
  const uint32_t tmem_base = bars->tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = static_cast<int>(tc::cluster_ctarank());
  const int rn = crank % g.cn, rm = crank / g.cn;

  if (warp == 0 || warp == 3) {
    producer_loop<PAIR>(&tmap_x, &tmap_w1, smem, bars, g, rn, rm, warp == 0);   // warp-converged; one elected lane issues
  } else if (warp == 1) {
    if (!PAIR || rm == 0) mma_loop<PAIR>(smem, bars, g, tmem_base, rn, rm);   // pair: the leader CTA issues for both
  } else if (warp >= kEpiWarp0) {
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int cg = ew >> 2;            // column group 0..3
    const int cpg = a.nv / kColGroups;
    // epilogue smem: [2][128][nv] bf16 H staging | per-warp bias | partial scores
    uint8_t* hstage = extra;
    float* sbias = reinterpret_cast<float*>(extra + 2 * kBlockM * a.nv * 2) + ew * (kBiasSmemPerWarp / 4);
    float* spart = reinterpret_cast<float*>(extra + 2 * kBlockM * a.nv * 2 + kEpiWarps * kBiasSmemPerWarp);
    const bool store_thread = (ew == 0 && lane == 0);
    if (store_thread) tc::prefetch_tensormap(&tmap_h);
    const int q_row = 32 * q + lane;
    const int col0 = cg * cpg;
    const bool eprof = (g.debug & 16) != 0;
    long long e_wait = 0, e_math = 0, e_store = 0, e_tiles = 0;
    // first expert this column group touches inside a tile (no division in the loop below)
    const int e_in_tile = a.chunks_per_expert > 0 ? cg * (cpg / a.es) : cg / a.span;
    TileCoord t;
    for (int it = 0; tile_at(g, it, rn, rm, t); ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1u;
      const int n_tile0 = t.n_blk * a.nv;
      stage_bias(sbias, a.b1 != nullptr ? a.b1 + n_tile0 + col0 : nullptr,
                 a.b1 != nullptr ? a.b1 + a.h + n_tile0 + col0 : nullptr, cpg, lane);
      const long long te0 = eprof ? clock64() : 0;
      tc::mbar_wait(&bars->tmem_full[as], aph);
      tc::fence_after_thread_sync();
      const long long te1 = eprof ? clock64() : 0;
      if (store_thread) trace_at(g, 8 + 4 * it + 1);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + as * kAccStride;
      const int row = t.m_blk * kBlockM + q_row;
      const bool row_ok = row < g.rows;
      uint8_t* hbuf = hstage + (it & 1) * kBlockM * a.nv * 2;
      __nv_bfloat16* hrow = reinterpret_cast<__nv_bfloat16*>(hbuf) + q_row * a.nv;
      const int e_first = t.n_blk * a.experts_per_tile + e_in_tile;
      if (g.debug & 4) {
      } else if (act == MOE_ACT_GELU)
        geglu_epilogue_group<CH, MOE_ACT_GELU, FEAT>(a, taddr, sbias, hrow, spart, row, row_ok, n_tile0 + col0,
                                                     col0, cpg, e_first, cg, q);
      else
        geglu_epilogue_group<CH, MOE_ACT_RELU, FEAT>(a, taddr, sbias, hrow, spart, row, row_ok, n_tile0 + col0,
                                                     col0, cpg, e_first, cg, q);
      const long long te2 = eprof ? clock64() : 0;
      if (store_thread) trace_at(g, 8 + 4 * it + 2);
      // accumulator stage drained -> MMA may overwrite it
      tc::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR)
          tc::mbar_arrive_cluster(tc::mapa_u32(&bars->tmem_empty[as], 0));   // the leader's MMA thread waits
        else
          tc::mbar_arrive(&bars->tmem_empty[as]);
      }
      // H tile: generic-proxy smem writes -> async proxy, then one TMA store for the whole tile.
      // The store thread first waits until the previous tile's store has finished READING its
      // buffer, so after the barrier everybody may overwrite that buffer in the next iteration.
      tc::fence_proxy_async_smem();
      if (store_thread) tc::tma_store_wait_read<0>();
      tc::named_bar_sync(kEpiBarrierId, kEpiThreads);
      if (store_thread) {
        tc::tma_store_2d(&tmap_h, hbuf, n_tile0, t.m_blk * kBlockM);   // rows beyond T are clipped
        tc::tma_store_commit();
        trace_at(g, 8 + 4 * it + 3);
      }
      if (eprof) {
        e_wait += te1 - te0;
        e_math += te2 - te1;
        e_store += clock64() - te2;
        ++e_tiles;
      }
    }
    if (store_thread) tc::tma_store_wait<0>();
    if (store_thread) trace_at(g, 4);
    if (eprof && store_thread && blockIdx.x < 256) {
      g_dbg_counters[blockIdx.x * 8 + 7] = e_tiles;
      // pack the three epilogue sums into the CTA's slots 0..2 of a second bank (offset 1024)
      if (blockIdx.x < 128) {
        g_dbg_counters[1024 + blockIdx.x * 8 + 0] = e_wait;
        g_dbg_counters[1024 + blockIdx.x * 8 + 1] = e_math;
        g_dbg_counters[1024 + blockIdx.x * 8 + 2] = e_store;
        g_dbg_counters[1024 + blockIdx.x * 8 + 3] = e_tiles;
      }
    }
  }
  teardown_pipeline<PAIR>(bars, g);
}

// ------------------------------------------------------------------------------------------
// K3: down-projection
// ------------------------------------------------------------------------------------------
struct DownArgs {
  const float* b2;     // [d] or null
  __nv_bfloat16* Y;    // [T, d]
  int d;
  float* ws_partial;   // split-K: [split_k][T][d] fp32 partial sums
  int* ws_counters;    // split-K: one arrival counter per output tile (kept at zero between calls)
  const uint32_t* mask_bits;   // weight-masked form: bit (r * h + c) set = W2[r, c] is removed; else null
  int h;
};

// Weight-masked down-projection (WandaRemoveNeuronsFast): the 16 epilogue warps -- idle while a tile's main loop runs --
// edit every landed W2 tile in shared memory, a 16-bit zero store per set mask bit (Wanda masks are 2-12 % dense), between
// the TMA's `full` and the MMA thread's `masked` barrier, so the masked weights never exist in global memory: DRAM traffic
// per call = W2 + the mask bits (d h / 8 bytes).  Warp w owns rows [w * tile_n / 16, ...) of the B tile, two lanes per row
// (32 of a k-block's 64 mask bits each); the bit words of the NEXT stage are fetched before the current one is waited
// for.  Cost against the unmasked kernel: the epilogue of tile i no longer overlaps the main loop of tile i + 1 (the same
// warps mask first, then drain the accumulator).  (Round 2's first version used ONE masker warp: 2-8x slower than
// mask_weights + down_proj, profiles/r02_wanda_masked_k3.log.)
constexpr int kMaskSubs = 2;     // k-blocks per stage the masked form supports (the host caps ks)

struct MaskState {
  int s;
  uint32_t ph;
};

__device__ __forceinline__ void mask_fetch(const GemmShape& g, const DownArgs& a, const TileCoord& t, int kb, int row, int half,
                                           bool row_ok, uint32_t (&bits)[kMaskSubs]) {
  const int n_sub = min(g.ks, t.kb_end - kb);
  const int grow = t.n_blk * g.tile_n + row;
#pragma unroll
  for (int sub = 0; sub < kMaskSubs; ++sub)
    bits[sub] = (row_ok && sub < n_sub && grow < a.d)      // (rows beyond d are zero-filled by the TMA anyway)
                    ? __ldg(a.mask_bits + ((static_cast<size_t>(grow) * a.h) >> 5) + 2 * (kb + sub) + half)
                    : 0u;
}

// mask every stage of tile t's main loop as it lands (called by all lanes of every epilogue warp)
__device__ __forceinline__ void mask_tile(uint8_t* smem, PipeBarriers* bars, const GemmShape& g, const DownArgs& a,
                                          const TileCoord& t, int ew, int lane, MaskState& ms) {
  const int rpw = g.tile_n / kEpiWarps;                 // rows of the B tile per warp (tile_n is a multiple of 16)
  const int row = ew * rpw + (lane >> 1), half = lane & 1;
  const bool row_ok = (lane >> 1) < rpw;
  uint32_t cur[kMaskSubs], nxt[kMaskSubs];
  mask_fetch(g, a, t, t.kb_begin, row, half, row_ok, cur);
  for (int kb = t.kb_begin; kb < t.kb_end; kb += g.ks) {
    if (kb + g.ks < t.kb_end) mask_fetch(g, a, t, kb + g.ks, row, half, row_ok, nxt);
    tc::mbar_wait(&bars->full[ms.s], ms.ph);
    const uint32_t b_base = tc::smem_u32(smem + ms.s * g.stage_bytes) + g.ks * kABytes;
#pragma unroll
    for (int sub = 0; sub < kMaskSubs; ++sub) {
      // 128-byte-swizzled K-major tile: 16-byte chunk c of row j sits at chunk (c ^ (j & 7))
      const uint32_t rowaddr = b_base + static_cast<uint32_t>(sub * g.tile_n + row) * 128u;
      const uint32_t sw = static_cast<uint32_t>(row & 7);
      uint32_t w = cur[sub];
      while (w) {
        const uint32_t b = static_cast<uint32_t>(32 * half) + static_cast<uint32_t>(__ffs(w) - 1);
        w &= w - 1;
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(rowaddr + (((b >> 3) ^ sw) << 4) + ((b & 7u) << 1)), "h"(static_cast<unsigned short>(0)) : "memory");
      }
    }
    tc::fence_proxy_async_smem();      // generic-proxy edits -> the tensor core's async-proxy reads
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&bars->masked[ms.s]);
#pragma unroll
    for (int i = 0; i < kMaskSubs; ++i) cur[i] = nxt[i];
    if (++ms.s == g.stages) {
      ms.s = 0;
      ms.ph ^= 1u;
    }
  }
}

template <int CH, bool PAIR, bool MASK = false>
__global__ void __launch_bounds__(kNumThreads, 1)
down_proj_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_w2,
                 const GemmShape g, const DownArgs a) {
  uint8_t *smem, *extra;
  PipeBarriers* bars = setup_pipeline<PAIR>(smem, extra, g, &tmap_h, &tmap_w2);
  const uint32_t tmem_base = bars->tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = static_cast<int>(tc::cluster_ctarank());
  const int rn = crank % g.cn, rm = crank / g.cn;

  if (warp == 0 || warp == 3) {
    producer_loop<PAIR>(&tmap_h, &tmap_w2, smem, bars, g, rn, rm, warp == 0);   // warp-converged; one elected lane issues
  } else if (warp == 1) {
    if (!PAIR || rm == 0) mma_loop<PAIR, MASK>(smem, bars, g, tmem_base, rn, rm);   // pair: the leader CTA issues for both
  } else if (warp >= kEpiWarp0) {
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;
    const int cg = ew >> 2;
    const int cpg = g.tile_n / kColGroups;
    float* sbias = reinterpret_cast<float*>(extra) + ew * (kBiasSmemPerWarp / 4);
    TileCoord t;
    MaskState ms = {0, 0u};
    for (int it = 0; tile_at(g, it, rn, rm, t); ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1u;
      const int n0 = t.n_blk * g.tile_n + cg * cpg;    // first output column of this warp's group
      const int nvalid = max(0, min(cpg, a.d - n0));   // the last tile may overhang d
      if constexpr (MASK) mask_tile(smem, bars, g, a, t, ew, lane, ms);   // this tile's W2 stages, as they land
      stage_bias(sbias, a.b2 != nullptr ? a.b2 + n0 : nullptr, nullptr, nvalid, lane);
      tc::mbar_wait(&bars->tmem_full[as], aph);
      tc::fence_after_thread_sync();
      if (ew == 0 && lane == 0) trace_at(g, 8 + 4 * it + 1);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + as * kAccStride + cg * cpg;
      const int row = t.m_blk * kBlockM + 32 * q + lane;
      const bool row_ok = row < g.rows;
      if (g.split_k == 1) {
        for (int c = 0; c < cpg; c += CH) {
          if (g.debug & 4) break;
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
          if (row_ok) {
            __nv_bfloat16* dst = a.Y + static_cast<size_t>(row) * a.d + n0 + c;
            if (c + CH <= nvalid) {
              store_words<CH / 2>(dst, yw);
            } else {
#pragma unroll
              for (int i = 0; i < CH / 2; ++i)
                if (c + 2 * i < nvalid) *reinterpret_cast<uint32_t*>(dst + 2 * i) = yw[i];
            }
          }
        }
        tc::fence_before_thread_sync();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR)
            tc::mbar_arrive_cluster(tc::mapa_u32(&bars->tmem_empty[as], 0));   // the leader's MMA thread waits
          else
            tc::mbar_arrive(&bars->tmem_empty[as]);
        }
      } else {
        // ---- split-K: park this slice's fp32 partial tile, the last slice to arrive reduces
        const size_t plane = static_cast<size_t>(g.rows) * a.d;
        float* part = a.ws_partial + t.slice * plane + static_cast<size_t>(row) * a.d + n0;
        for (int c = 0; c < cpg; c += CH) {
          uint32_t acc[CH];
          tc::tmem_ld_cols<CH>(taddr + c, acc);
          tc::tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < CH; i += 4)
              if (c + i < nvalid)   // d and the group widths are multiples of 4
                __stcg(reinterpret_cast<float4*>(part + c + i),
                       make_float4(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1]), __uint_as_float(acc[i + 2]),
                                   __uint_as_float(acc[i + 3])));
          }
        }
        tc::fence_before_thread_sync();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR)
            tc::mbar_arrive_cluster(tc::mapa_u32(&bars->tmem_empty[as], 0));
          else
            tc::mbar_arrive(&bars->tmem_empty[as]);
        }
        __threadfence();                                   // partial tile visible device-wide ...
        tc::named_bar_sync(kEpiBarrierId, kEpiThreads);    // ... before the arrival is counted
        int* flag = reinterpret_cast<int*>(extra + kEpiWarps * kBiasSmemPerWarp);
        if (ew == 0 && lane == 0) {
          int* counter = a.ws_counters + t.m_blk * g.n_tiles + t.n_blk;
          const int prev = atomicAdd(counter, 1);
          const int last = prev == g.split_k - 1;
          if (last) *counter = 0;                          // leave the workspace clean for the next call
          *flag = last;
        }
        tc::named_bar_sync(kEpiBarrierId, kEpiThreads);
        if (*flag) {
          __threadfence();
          // coalesced reduction over the whole 128 x tile_n tile: 4 consecutive columns per thread, all
          // slices' loads in flight at once, summed in slice order (deterministic)
          const int et = (warp - kEpiWarp0) * 32 + lane;            // 0 .. 511
          const int cols4 = g.tile_n >> 2;                           // float4 groups per tile row
          const int tile_col0 = t.n_blk * g.tile_n;
          const int tile_row0 = t.m_blk * kBlockM;
          for (int e = et; e < kBlockM * cols4; e += kEpiThreads) {
            const int r = e / cols4, c4 = (e - r * cols4) << 2;
            const int grow = tile_row0 + r, gcol = tile_col0 + c4;
            if (grow >= g.rows || gcol >= a.d) continue;
            const float* src = a.ws_partial + static_cast<size_t>(grow) * a.d + gcol;
            float4 p[8];
#pragma unroll
            for (int sl = 0; sl < 8; ++sl)
              if (sl < g.split_k) p[sl] = __ldcg(reinterpret_cast<const float4*>(src + sl * plane));
            float4 sum = a.b2 != nullptr ? __ldg(reinterpret_cast<const float4*>(a.b2 + gcol)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int sl = 0; sl < 8; ++sl)
              if (sl < g.split_k) {
                sum.x += p[sl].x;
                sum.y += p[sl].y;
                sum.z += p[sl].z;
                sum.w += p[sl].w;
              }
            *reinterpret_cast<uint2*>(a.Y + static_cast<size_t>(grow) * a.d + gcol) =
                make_uint2(pack_bf16x2(sum.x, sum.y), pack_bf16x2(sum.z, sum.w));
          }
        }
        tc::named_bar_sync(kEpiBarrierId, kEpiThreads);    // flag may be rewritten by the next tile
      }
      if (ew == 0 && lane == 0) trace_at(g, 8 + 4 * it + 2);
    }
    if (ew == 0 && lane == 0) trace_at(g, 4);
  }
  teardown_pipeline<PAIR>(bars, g);
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
struct ClusterChoice {
  int cn, cm;
};

// L2->SM bandwidth and tensor rate used by the tiling heuristics (measured on B200)
constexpr double kL2BytesPerCycle = 3600.0;   // ~6.8 TB/s at 1.9 GHz, whole chip
static double mma_cycles_per_kstep(int tile_n) {
  const double tensor = tile_n / 2.0;                       // 128 x N x 16 at 4096 MAC/cycle/SM
  const double smem = (4096.0 + 32.0 * tile_n) / 128.0;     // A + B operand bytes at 128 B/cycle
  return tensor > smem ? tensor : smem;
}

static bool env_cluster(const char* name, ClusterChoice& c) {
  const char* e = getenv(name);
  if (!e) return false;
  int a = 0, b = 0;
  if (sscanf(e, "%d,%d", &a, &b) != 2 || a < 1 || b < 1 || a * b > 8) return false;
  c.cn = a;
  c.cm = b;
  return true;
}

// Pick the cluster shape minimising max(tensor time over waves, panel traffic / L2 bandwidth).
static ClusterChoice choose_cluster(const char* env, int m_tiles, int n_tiles, int tile_n, int half_rows, int k,
                                    double epi_cycles_per_tile) {
  ClusterChoice forced = {1, 1};
  const int cands[8][2] = {{1, 1}, {2, 1}, {1, 2}, {2, 2}, {4, 1}, {1, 4}, {4, 2}, {2, 4}};
  // Measured on B200 (profiles/r01_sweep_cluster.log): multicast does not pay -- the kernels are bound by the
  // single-thread TMA / MMA issue loops, not by L2 reads -- so it is used only when forced through `env`.
  const bool have_forced = env_cluster(env, forced);
  if (!have_forced) return ClusterChoice{1, 1};
  const int sms = sm_count();
  ClusterChoice best = {1, 1};
  double best_t = 1e300;
  for (auto& cd : cands) {
    const int cn = cd[0], cm = cd[1];
    if (have_forced && (cn != forced.cn || cm != forced.cm)) continue;
    if (n_tiles % cn) continue;                         // weight rows beyond N must never be addressed
    if (kBlockM % (8 * cn)) continue;
    const int slice = tile_n / cm;
    if (tile_n % cm || slice % 8) continue;             // 8-row swizzle groups stay whole
    if (slice > half_rows ? (slice % half_rows) : (half_rows % slice)) continue;   // boxes do not straddle halves
    if ((slice > half_rows ? half_rows : slice) > 256) continue;
    const int csize = cn * cm;
    const int n_clusters = sms / csize;
    if (n_clusters < 1) continue;
    const long long ctiles = static_cast<long long>((m_tiles + cm - 1) / cm) * (n_tiles / cn);
    const long long rounds = (ctiles + n_clusters - 1) / n_clusters;
    const double ksteps = (k + kUmmaK - 1) / kUmmaK;
    const double per_tile = ksteps * mma_cycles_per_kstep(tile_n);
    const double compute = rounds * (per_tile > epi_cycles_per_tile ? per_tile : epi_cycles_per_tile);
    const double bytes = static_cast<double>(ctiles) * (cm * kBlockM + cn * tile_n) * k * 2.0;
    const double traffic = bytes / kL2BytesPerCycle;
    // multicast is not free (cluster barriers, lock-step): 3 % per doubling as a tie-breaker
    double t = compute > traffic ? compute : traffic;
    for (int c = csize; c > 1; c >>= 1) t *= 1.03;
    if (t < best_t) {
      best_t = t;
      best = {cn, cm};
    }
  }
  return best;
}

// cta_group::2 pairs unless MOE_PAIR=0
static bool pair_enabled() {
  const char* e = getenv("MOE_PAIR");
  return !(e && atoi(e) == 0);
}

// k-blocks per pipeline stage: the deepest of {4, 2, 1} that still leaves >= 3 stages in shared memory
// (MOE_KS overrides).  Needs K % 64 == 0 (3-D k-block tensor maps) and a single B box per stage.
static int pick_ks(int k, int num_kb_per_item, int b_rows_per_cta, int extra_smem, bool allowed) {
  if (!allowed || k % kBlockK != 0) return 1;
  int forced = 0;
  if (const char* e = getenv("MOE_KS")) forced = atoi(e);
  for (int ks = 4; ks >= 1; ks >>= 1) {
    if (forced >= 1 && ks > forced) continue;
    if (ks > 1 && num_kb_per_item < 2 * ks) continue;
    const int stage = ks * (kABytes + b_rows_per_cta * 128);
    if ((kSmemLimit - 1024 - extra_smem - static_cast<int>(sizeof(PipeBarriers))) / stage >= 3) return ks;
  }
  return 1;
}

static int debug_mode() {
  const char* e = getenv("MOE_DEBUG_MODE");
  return e ? atoi(e) : 0;
}

static int finish_shape(GemmShape& g, int extra_smem) {
  g.debug = debug_mode();
  g.a_box_rows = kBlockM / g.cn;
  const int slice = g.tile_n / g.cm;
  g.b_box_rows = slice < g.half_rows ? slice : g.half_rows;
  g.b_boxes = slice / g.b_box_rows;
  g.m_ctiles = (g.m_tiles + g.cm - 1) / g.cm;
  g.n_ctiles = g.n_tiles / g.cn;
  if (g.ks < 1) g.ks = 1;
  g.stage_bytes = g.ks * (kABytes + (g.pair ? g.tile_n / 2 : g.tile_n) * 128);
  g.extra_smem = extra_smem;
  int stages = (kSmemLimit - 1024 - extra_smem - static_cast<int>(sizeof(PipeBarriers))) / g.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (const char* e = getenv("MOE_DEBUG_STAGES")) {
    const int cap = atoi(e);
    if (cap >= 2 && cap < stages) stages = cap;
  }
  if (stages < 2) return -1;
  g.stages = stages;
  return 0;
}

static size_t smem_bytes(const GemmShape& g) {
  return static_cast<size_t>(g.stages) * g.stage_bytes + 1024 + g.extra_smem + sizeof(PipeBarriers);
}

// opt in to the full 227 KiB of dynamic shared memory, once per kernel (keyed by entry address:
// all instantiations of one template share a function-pointer TYPE)
static int ensure_smem(const void* kfn, size_t bytes) {
  if (bytes > static_cast<size_t>(kSmemLimit)) return fail(MOE_ERR_UNSUPPORTED_SHAPE, "smem request %zu too large", bytes);
  return ensure_dynamic_smem(kfn, kSmemLimit);   // per (kernel, device)
}

template <typename... KArgs, typename... Args>
static int launch_clustered(void (*kernel)(KArgs...), const GemmShape& g, cudaStream_t st, Args&&... args) {
  const size_t smem = smem_bytes(g);
  int rc = ensure_smem(reinterpret_cast<const void*>(kernel), smem);
  if (rc) return rc;
  const int csize = g.cn * g.cm;
  const long long ctiles = static_cast<long long>(g.m_ctiles) * g.n_ctiles * g.split_k;
  long long n_clusters = sm_count() / csize;
  if (n_clusters > ctiles) n_clusters = ctiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(n_clusters * csize));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(csize);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
  if (e != cudaSuccess) return fail(MOE_ERR_CUDA, "cudaLaunchKernelEx(cluster %dx%d): %s", g.cn, g.cm, cudaGetErrorString(e));
  return MOE_OK;
}

}  // namespace moe

extern "C" {

int moe_debug_counters(unsigned long long* host_out, int n) {
  using namespace moe;
  MOE_REQUIRE(host_out != nullptr && n >= 0 && n <= 256 * 8, MOE_ERR_INVALID_ARGUMENT, "moe_debug_counters: bad args");
  cudaError_t e = cudaMemcpyFromSymbol(host_out, g_dbg_counters, sizeof(unsigned long long) * n);
  if (e != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_debug_counters: %s", cudaGetErrorString(e));
  return MOE_OK;
}

int moe_debug_trace(unsigned long long* host_out, int n) {
  using namespace moe;
  MOE_REQUIRE(host_out != nullptr && n >= 0 && n <= 256 * 64, MOE_ERR_INVALID_ARGUMENT, "moe_debug_trace: bad args");
  cudaError_t e = cudaMemcpyFromSymbol(host_out, g_trace, sizeof(unsigned long long) * n);
  if (e != cudaSuccess) return fail(MOE_ERR_CUDA, "moe_debug_trace: %s", cudaGetErrorString(e));
  void* sym = nullptr;   // clear, so that the next traced launch is seen alone
  if (cudaGetSymbolAddress(&sym, g_trace) == cudaSuccess) cudaMemset(sym, 0, sizeof(unsigned long long) * 256 * 64);
  return MOE_OK;
}

int moe_geglu_up(const void* x, const void* w1p, const float* b1p, const uint8_t* neuron_override,
                 float override_value, void* H, float* scores, void* gate_out, int T, int d, int h, int E,
                 int es, int act, void* stream) {
  using namespace moe;
  MOE_REQUIRE(T >= 0 && d >= 8 && h >= 8 && E >= 1 && es >= 1, MOE_ERR_INVALID_ARGUMENT,
              "moe_geglu_up: bad sizes T=%d d=%d h=%d E=%d es=%d", T, d, h, E, es);
  if (T == 0) return MOE_OK;   // empty tensors have no storage to point at
  MOE_REQUIRE(x && w1p && H, MOE_ERR_INVALID_ARGUMENT, "moe_geglu_up: NULL x / w1p / H");
  MOE_REQUIRE(static_cast<long long>(E) * es == h, MOE_ERR_INVALID_ARGUMENT, "moe_geglu_up: E*es=%d*%d != h=%d", E, es, h);
  MOE_REQUIRE(d % 8 == 0 && h % 8 == 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_geglu_up: d=%d and h=%d must be multiples of 8", d, h);
  MOE_REQUIRE(act == MOE_ACT_GELU || act == MOE_ACT_RELU, MOE_ERR_INVALID_ARGUMENT, "moe_geglu_up: act=%d", act);
  MOE_REQUIRE(es % 4 == 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_geglu_up: expert size %d unsupported (needs es %% 4 == 0)", es);
  MOE_REQUIRE(((reinterpret_cast<uintptr_t>(H) | reinterpret_cast<uintptr_t>(gate_out)) & 15) == 0,
              MOE_ERR_INVALID_ARGUMENT, "moe_geglu_up: H / gate_out must be 16-byte aligned");
  if (T == 0) return MOE_OK;
  // nv neuron pairs per tile (UMMA N = 2 nv <= 256) split over 4 epilogue column groups of cpg = nv/4:
  // a group holds whole experts (cpg % es == 0) or an expert spans 2 or 4 groups inside the tile.
  int nv = 0;
  for (int cand = 128; cand >= 16; cand -= 8) {
    if (h % cand || cand % 4) continue;
    const int cpg = cand / kColGroups;
    if (cpg % 4 || cpg > 64) continue;
    const bool whole = cpg % es == 0;
    const bool spans = es % cpg == 0 && (es / cpg == 2 || es / cpg == 4) && cand % es == 0;
    if (whole || spans) {
      nv = cand;
      break;
    }
  }
  MOE_REQUIRE(nv > 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_geglu_up: expert size %d unsupported for h=%d (no tile width)", es, h);
  const int cpg = nv / kColGroups;
  int gcd_v = cpg, tmp = es;
  while (tmp) {
    const int r = gcd_v % tmp;
    gcd_v = tmp;
    tmp = r;
  }
  const int ch = (gcd_v % 32 == 0) ? 32 : (gcd_v % 20 == 0) ? 20 : (gcd_v % 16 == 0) ? 16 : (gcd_v % 8 == 0) ? 8 : 4;

  GemmShape g = {};
  g.rows = T;
  g.k = d;
  g.m_tiles = (T + kBlockM - 1) / kBlockM;
  g.n_tiles = h / nv;
  g.tile_n = 2 * nv;
  g.half_rows = nv;
  g.second_off = h;
  g.split_k = 1;
  g.kb_per_slice = (d + kBlockK - 1) / kBlockK;
  ClusterChoice forced_cc;
  int pair_min_d = 64;
  if (const char* e = getenv("MOE_K1_PAIR_MIN_D")) pair_min_d = atoi(e);
  if (g.m_tiles >= 2 && d >= pair_min_d && pair_enabled() && !env_cluster("MOE_K1_CLUSTER", forced_cc)) {
    // cta_group::2: CTA 0 of the pair stages the value rows, CTA 1 the gate rows of the tile's weights
    g.pair = 1;
    g.cn = 1;
    g.cm = 2;
  } else {
    // epilogue issue cycles per tile and scheduler: 4 warps x nv/4 neuron pairs x ~19 slots
    const ClusterChoice cc = choose_cluster("MOE_K1_CLUSTER", g.m_tiles, g.n_tiles, g.tile_n, nv, d, 19.0 * nv);
    g.cn = cc.cn;
    g.cm = cc.cm;
  }
  const int extra = 2 * kBlockM * nv * 2 + kEpiWarps * kBiasSmemPerWarp + kBlockM * kColGroups * 4;
  g.ks = pick_ks(d, g.kb_per_slice, nv, extra, g.pair != 0);   // pair mode: one B box (value or gate rows) per CTA
  MOE_REQUIRE(finish_shape(g, extra) == 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_geglu_up: tile does not fit smem");
  CUtensorMap tx, tw, th;
  int rc;
  if (g.ks > 1) {
    rc = make_tmap_bf16_kblocks(&tx, x, static_cast<uint64_t>(T), static_cast<uint64_t>(d), kBlockM, static_cast<uint32_t>(g.ks));
    if (rc) return rc;
    rc = make_tmap_bf16_kblocks(&tw, w1p, static_cast<uint64_t>(2) * h, static_cast<uint64_t>(d), static_cast<uint32_t>(g.b_box_rows),
                                static_cast<uint32_t>(g.ks));
  } else {
    rc = make_tmap_bf16_2d(&tx, x, static_cast<uint64_t>(T), static_cast<uint64_t>(d), static_cast<uint32_t>(g.a_box_rows), kBlockK);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tw, w1p, static_cast<uint64_t>(2) * h, static_cast<uint64_t>(d), static_cast<uint32_t>(g.b_box_rows), kBlockK);
  }
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&th, H, static_cast<uint64_t>(T), static_cast<uint64_t>(h), kBlockM, static_cast<uint32_t>(nv), false);
  if (rc) return rc;
  GegluArgs a;
  a.b1 = b1p;
  a.neuron_override = neuron_override;
  a.override_value = override_value;
  a.scores = scores;
  a.gate_out = static_cast<__nv_bfloat16*>(gate_out);
  a.h = h;
  a.E = E;
  a.es = es;
  a.nv = nv;
  a.experts_per_tile = nv / es;
  a.chunks_per_expert = (cpg % es == 0) ? es / ch : 0;
  a.span = (cpg % es == 0) ? 1 : es / cpg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool feat = neuron_override != nullptr || gate_out != nullptr;
#define MOE_LAUNCH_GEGLU(CHV)                                                                              \
  rc = g.pair ? (feat ? launch_clustered(geglu_up_kernel<CHV, true, true>, g, st, tx, tw, th, g, a, act)   \
                      : launch_clustered(geglu_up_kernel<CHV, false, true>, g, st, tx, tw, th, g, a, act)) \
              : (feat ? launch_clustered(geglu_up_kernel<CHV, true, false>, g, st, tx, tw, th, g, a, act)  \
                      : launch_clustered(geglu_up_kernel<CHV, false, false>, g, st, tx, tw, th, g, a, act))
  switch (ch) {
    case 32: MOE_LAUNCH_GEGLU(32); break;
    case 20: MOE_LAUNCH_GEGLU(20); break;
    case 16: MOE_LAUNCH_GEGLU(16); break;
    case 8: MOE_LAUNCH_GEGLU(8); break;
    default: MOE_LAUNCH_GEGLU(4); break;
  }
#undef MOE_LAUNCH_GEGLU
  if (rc) return rc;
  return check_launch("moe_geglu_up");
}

static const int kCounterBytes = 64 * 1024;   // head of the split-K workspace: per-tile arrival counters

size_t moe_down_proj_workspace_bytes(int T, int h, int d) {
  // enough for 8 K-slices of fp32 partial sums, capped at 64 MiB, plus the counters
  (void)h;
  size_t want = static_cast<size_t>(8) * static_cast<size_t>(T > 0 ? T : 0) * static_cast<size_t>(d > 0 ? d : 0) * 4;
  const size_t cap = static_cast<size_t>(64) << 20;
  if (want > cap) want = cap;
  return kCounterBytes + want;
}

static int down_proj_impl(const void* H, const void* w2p, const uint32_t* mask_bits, const float* b2, void* Y, int T, int h,
                          int d, void* workspace, size_t workspace_bytes, void* stream);

int moe_down_proj(const void* H, const void* w2p, const float* b2, void* Y, int T, int h, int d, void* workspace,
                  size_t workspace_bytes, void* stream) {
  return down_proj_impl(H, w2p, nullptr, b2, Y, T, h, d, workspace, workspace_bytes, stream);
}

int moe_down_proj_masked(const void* H, const void* w2p, const uint32_t* mask_bits, const float* b2, void* Y, int T, int h,
                         int d, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace moe;
  MOE_REQUIRE(mask_bits != nullptr, MOE_ERR_INVALID_ARGUMENT, "moe_down_proj_masked: mask_bits is NULL (use moe_down_proj)");
  MOE_REQUIRE(h % 64 == 0, MOE_ERR_UNSUPPORTED_SHAPE,
              "moe_down_proj_masked: h=%d must be a multiple of 64 (use moe_mask_weights + moe_down_proj)", h);
  MOE_REQUIRE((reinterpret_cast<uintptr_t>(mask_bits) & 7) == 0, MOE_ERR_INVALID_ARGUMENT,
              "moe_down_proj_masked: mask_bits must be 8-byte aligned");
  return down_proj_impl(H, w2p, mask_bits, b2, Y, T, h, d, workspace, workspace_bytes, stream);
}

static int down_proj_impl(const void* H, const void* w2p, const uint32_t* mask_bits, const float* b2, void* Y, int T, int h,
                          int d, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace moe;
  const bool masked = mask_bits != nullptr;
  MOE_REQUIRE(T >= 0 && h >= 8 && d >= 8, MOE_ERR_INVALID_ARGUMENT, "moe_down_proj: bad sizes T=%d h=%d d=%d", T, h, d);
  if (T == 0) return MOE_OK;   // empty tensors have no storage to point at
  MOE_REQUIRE(H && w2p && Y, MOE_ERR_INVALID_ARGUMENT, "moe_down_proj: NULL H / w2p / Y");
  MOE_REQUIRE(h % 8 == 0 && d % 8 == 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_down_proj: h=%d and d=%d must be multiples of 8", h, d);
  MOE_REQUIRE((reinterpret_cast<uintptr_t>(Y) & 15) == 0, MOE_ERR_INVALID_ARGUMENT, "moe_down_proj: Y must be 16-byte aligned");
  MOE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, MOE_ERR_INVALID_ARGUMENT,
              "moe_down_proj: workspace must be 16-byte aligned");
  if (T == 0) return MOE_OK;
  // Tile width bn (a multiple of 16, <= 256; the last tile may overhang d by < 16 columns: TMA zero-fills the
  // missing weight rows and the epilogue guards its stores) and split-K factor S are chosen together.
  // Work units are CTA pairs (cta_group::2, 256 token rows) when there are at least two row tiles.  The cost
  // model (cycles) is fitted to profiles/r01_debug_clock*.log; every work item pays ~3000 for pipeline ramp +
  // epilogue; split-K adds a reduction pass proportional to slices x tile width.
  const int m_tiles = (T + kBlockM - 1) / kBlockM;
  const int sms = sm_count();
  ClusterChoice forced_cc = {1, 1};
  const bool forced = !masked && env_cluster("MOE_K3_CLUSTER", forced_cc);
  // (weight-masked form: single CTAs -- each masker warp signals its own CTA's MMA thread, no cluster-scope arrive)
  const bool pair = m_tiles >= 2 && pair_enabled() && !forced && !masked;
  const int units = pair ? sms / 2 : sms;
  const int m_units = pair ? (m_tiles + 1) / 2 : m_tiles;
  const int num_kb = (h + kBlockK - 1) / kBlockK;
  int max_split = 1;
  if (workspace != nullptr && workspace_bytes > static_cast<size_t>(kCounterBytes) && !forced && d % 4 == 0) {
    const size_t per_slice = static_cast<size_t>(T) * d * 4;
    const size_t fit = (workspace_bytes - kCounterBytes) / per_slice;
    max_split = fit > 8 ? 8 : static_cast<int>(fit);
    if (const char* e = getenv("MOE_K3_SPLIT")) {
      const int cap = atoi(e);
      if (cap >= 1 && cap < max_split) max_split = cap;
    }
    if (max_split < 1) max_split = 1;
  }
  int best = 0, best_split = 1;
  double best_cost = 1e300;
  for (int bn = 256; bn >= 16; bn -= 16) {
    if (bn - (d % bn ? d % bn : bn) >= 16) continue;
    if (pair && (bn / 2) % 8) continue;                  // each CTA of a pair stages bn/2 weight rows
    const int n_tiles_c = (d + bn - 1) / bn;
    if (static_cast<long long>(m_tiles + 1) * n_tiles_c * 4 > kCounterBytes) continue;
    const long long tiles = static_cast<long long>(m_units) * n_tiles_c;
    for (int sp = 1; sp <= max_split; ++sp) {
      if (sp > 1 && num_kb / sp < 8) break;              // keep at least 8 k-blocks per slice
      int kb_per = (num_kb + sp - 1) / sp;
      int ks_c = pick_ks(h, kb_per, pair ? bn / 2 : bn, kEpiWarps * kBiasSmemPerWarp + 16, !forced);
      if (masked && ks_c > kMaskSubs) ks_c = kMaskSubs;
      kb_per = (kb_per + ks_c - 1) / ks_c * ks_c;        // whole stages per slice
      if (sp > 1 && (sp - 1) * kb_per >= num_kb) continue;   // no empty slices
      const long long rounds = (tiles * sp + units - 1) / units;
      // measured: ~230 cycles of barrier wait + commit per stage round, plus 4 MMAs per k-block at
      // max(tensor, issue); the split-K epilogue (partials out, fence, reduction) is expensive
      const double per_kblock = 230.0 / ks_c + (2.0 * bn > 220.0 ? 2.0 * bn : 220.0);
      const double cost = rounds * (kb_per * per_kblock + 3000.0) + (sp > 1 ? 9000.0 + 40.0 * sp * bn : 0.0);
      if (cost < best_cost * 0.999) {
        best_cost = cost;
        best = bn;
        best_split = sp;
      }
    }
  }
  if (const char* e = getenv("MOE_K3_BN")) {   // tuning override: force the tile width (and MOE_K3_SPLIT the slices)
    const int bn = atoi(e);
    if (bn >= 16 && bn <= 256 && bn % 16 == 0 && (!pair || (bn / 2) % 8 == 0)) {
      best = bn;
      best_split = max_split;
      while (best_split > 1 && (best_split - 1) * ((num_kb + best_split - 1) / best_split) >= num_kb) --best_split;
    }
  }
  ClusterChoice best_cc = forced ? forced_cc : ClusterChoice{1, 1};
  if (pair) best_cc = ClusterChoice{1, 2};
  if (getenv("MOE_DEBUG_PRINT"))
    fprintf(stderr, "[moe_down_proj] T=%d h=%d d=%d -> bn=%d split<=%d pair=%d (model cost %.0f cycles)\n", T, h, d, best,
            best_split, pair ? 1 : 0, best_cost);
  MOE_REQUIRE(best > 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_down_proj: no tile width for d=%d", d);
  const int cpg = best / kColGroups;
  const int ch = (cpg % 32 == 0) ? 32 : (cpg % 20 == 0) ? 20 : (cpg % 16 == 0) ? 16 : (cpg % 8 == 0) ? 8 : 4;
  GemmShape g = {};
  g.rows = T;
  g.k = h;
  g.m_tiles = m_tiles;
  g.n_tiles = (d + best - 1) / best;
  g.tile_n = best;
  g.half_rows = best;
  g.second_off = 0;
  g.cn = best_cc.cn;
  g.cm = best_cc.cm;
  g.pair = pair ? 1 : 0;
  g.split_k = best_split;
  g.kb_per_slice = (num_kb + best_split - 1) / best_split;
  g.ks = pick_ks(h, g.kb_per_slice, pair ? best / 2 : best, kEpiWarps * kBiasSmemPerWarp + 16, !forced);
  if (masked && g.ks > kMaskSubs) g.ks = kMaskSubs;
  g.kb_per_slice = (g.kb_per_slice + g.ks - 1) / g.ks * g.ks;
  while (g.split_k > 1 && (g.split_k - 1) * g.kb_per_slice >= num_kb) --g.split_k;
  if (forced) {   // validate a forced multicast shape the same way the heuristic would
    const ClusterChoice ok = choose_cluster("MOE_K3_CLUSTER", m_tiles, g.n_tiles, best, best, h, 3.0 * best);
    g.cn = ok.cn;
    g.cm = ok.cm;
  }
  MOE_REQUIRE(finish_shape(g, kEpiWarps * kBiasSmemPerWarp + 16) == 0, MOE_ERR_UNSUPPORTED_SHAPE,
              "moe_down_proj: tile does not fit smem");
  CUtensorMap th, tw;
  int rc;
  if (g.ks > 1) {
    rc = make_tmap_bf16_kblocks(&th, H, static_cast<uint64_t>(T), static_cast<uint64_t>(h), kBlockM, static_cast<uint32_t>(g.ks));
    if (rc) return rc;
    rc = make_tmap_bf16_kblocks(&tw, w2p, static_cast<uint64_t>(d), static_cast<uint64_t>(h), static_cast<uint32_t>(g.b_box_rows),
                                static_cast<uint32_t>(g.ks));
  } else {
    rc = make_tmap_bf16_2d(&th, H, static_cast<uint64_t>(T), static_cast<uint64_t>(h), static_cast<uint32_t>(g.a_box_rows), kBlockK);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tw, w2p, static_cast<uint64_t>(d), static_cast<uint64_t>(h), static_cast<uint32_t>(g.b_box_rows), kBlockK);
  }
  if (rc) return rc;
  DownArgs a;
  a.b2 = b2;
  a.Y = static_cast<__nv_bfloat16*>(Y);
  a.d = d;
  a.ws_counters = static_cast<int*>(workspace);
  a.ws_partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + kCounterBytes);
  a.mask_bits = mask_bits;
  a.h = h;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (masked) MOE_REQUIRE(g.cn == 1 && g.cm == 1 && !g.pair && g.tile_n % kEpiWarps == 0 && g.ks <= kMaskSubs,
                          MOE_ERR_UNSUPPORTED_SHAPE, "moe_down_proj_masked: tile shape not supported");
#define MOE_LAUNCH_DOWN(CHV)                                                                   \
  rc = masked ? launch_clustered(down_proj_kernel<CHV, false, true>, g, st, th, tw, g, a)       \
       : g.pair ? launch_clustered(down_proj_kernel<CHV, true>, g, st, th, tw, g, a)            \
                : launch_clustered(down_proj_kernel<CHV, false>, g, st, th, tw, g, a)
  switch (ch) {
    case 32: MOE_LAUNCH_DOWN(32); break;
    case 20: MOE_LAUNCH_DOWN(20); break;
    case 16: MOE_LAUNCH_DOWN(16); break;
    case 8: MOE_LAUNCH_DOWN(8); break;
    default: MOE_LAUNCH_DOWN(4); break;
  }
#undef MOE_LAUNCH_DOWN
  if (rc) return rc;
  return check_launch("moe_down_proj");
}

}  // extern "C"
