// K1 (GEGLU up-projection, fused activation / override / expert-score epilogue) and
// K3 (down-projection) for sm_100a: persistent, warp-specialised tcgen05 GEMMs.
//
//   warp 0     : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count)
//   warp 1     : MMA issuer     (one thread issues tcgen05.mma, accumulators in TMEM, 2 stages)
//   warp 2     : TMEM allocator
//   warp 3     : idle
//   warps 4-11 : epilogue       (tcgen05.ld -> registers -> fused math -> global), two column groups
//                               x four TMEM lane quarters, overlapped with the next tile's MMAs
//
// Operands are bf16, both K-major: A = activations [rows, K], B = weights [N, K] (nn.Linear layout),
// so D[m, n] = sum_k A[m, k] B[n, k] needs no transposes.  One UMMA per 16-wide K step covers the
// whole tile width; for K1 the B tile is [value rows | gate rows] so that value and gate
// accumulators of the same neurons sit side by side in one TMEM stage.
#include "common.cuh"
#include "tcgen05.cuh"

namespace moe {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                      // 64 bf16 = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KiB
constexpr int kAccStride = 256;                  // TMEM columns between the two accumulator stages
constexpr int kTmemCols = 512;
constexpr int kNumThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kMaxStages = 8;
constexpr int kSmemLimit = 232448;               // 227 KiB opt-in dynamic shared memory

struct PipeBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

struct GemmShape {
  int rows;        // M (tokens)
  int k;           // reduction length
  int m_tiles;
  int n_tiles;
  int tile_n;      // UMMA N (accumulator columns per stage)
  int b_rows;      // rows per B TMA box (tile_n for K3, tile_n/2 for K1: two boxes per stage)
  int stages;
  int stage_bytes;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

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

// store kN packed bf16 pairs (kN/2 words) to a row pointer with the widest aligned vectors
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

// ------------------------------------------------------------------------------------------
// Shared mainloop roles
// ------------------------------------------------------------------------------------------
template <bool kTwoBBoxes>
__device__ __forceinline__ void producer_loop(const CUtensorMap* tmap_a, const CUtensorMap* tmap_b, uint8_t* smem,
                                              PipeBarriers* bars, const GemmShape& g, int b_row_offset2) {
  const int num_kb = (g.k + kBlockK - 1) / kBlockK;
  const int total = g.m_tiles * g.n_tiles;
  int s = 0;
  uint32_t ph = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int m_blk = tile / g.n_tiles, n_blk = tile % g.n_tiles;
    for (int kb = 0; kb < num_kb; ++kb) {
      tc::mbar_wait(&bars->empty[s], ph ^ 1u);
      uint8_t* sa = smem + s * g.stage_bytes;
      uint8_t* sb = sa + kABytes;
      tc::mbar_arrive_expect_tx(&bars->full[s], static_cast<uint32_t>(g.stage_bytes));
      tc::tma_load_2d(sa, tmap_a, &bars->full[s], kb * kBlockK, m_blk * kBlockM);
      tc::tma_load_2d(sb, tmap_b, &bars->full[s], kb * kBlockK, n_blk * g.b_rows);
      if constexpr (kTwoBBoxes)
        tc::tma_load_2d(sb + g.b_rows * 128, tmap_b, &bars->full[s], kb * kBlockK, b_row_offset2 + n_blk * g.b_rows);
      if (++s == g.stages) {
        s = 0;
        ph ^= 1u;
      }
    }
  }
}

__device__ __forceinline__ void mma_loop(uint8_t* smem, PipeBarriers* bars, const GemmShape& g, uint32_t tmem_base) {
  const int num_kb = (g.k + kBlockK - 1) / kBlockK;
  const int total = g.m_tiles * g.n_tiles;
  const uint32_t idesc = tc::umma_idesc_bf16_f32(kBlockM, static_cast<uint32_t>(g.tile_n));
  int s = 0;
  uint32_t ph = 0;
  int it = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
    const int as = it & 1;
    const uint32_t aph = (it >> 1) & 1u;
    tc::mbar_wait(&bars->tmem_empty[as], aph ^ 1u);
    tc::fence_after_thread_sync();
    const uint32_t d_tmem = tmem_base + as * kAccStride;
    for (int kb = 0; kb < num_kb; ++kb) {
      tc::mbar_wait(&bars->full[s], ph);
      tc::fence_after_thread_sync();
      const uint32_t a_addr = tc::smem_u32(smem + s * g.stage_bytes);
      const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
      for (int k = 0; k < kBlockK / kUmmaK; ++k) {
        const uint64_t da = tc::umma_desc_kmajor_sw128(a_addr + k * kUmmaK * 2);
        const uint64_t db = tc::umma_desc_kmajor_sw128(b_addr + k * kUmmaK * 2);
        tc::umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
      }
      tc::umma_commit(&bars->empty[s]);  // frees the smem slot once these MMAs have read it
      if (++s == g.stages) {
        s = 0;
        ph ^= 1u;
      }
    }
    tc::umma_commit(&bars->tmem_full[as]);  // accumulator complete -> epilogue
  }
}

__device__ __forceinline__ PipeBarriers* setup_pipeline(uint8_t*& smem, const GemmShape& g, const CUtensorMap* ta,
                                                        const CUtensorMap* tb) {
  extern __shared__ uint8_t smem_raw[];
  smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  PipeBarriers* bars = reinterpret_cast<PipeBarriers*>(smem + g.stages * g.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::prefetch_tensormap(ta);
    tc::prefetch_tensormap(tb);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < g.stages; ++i) {
      tc::mbar_init(&bars->full[i], 1);
      tc::mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars->tmem_full[i], 1);
      tc::mbar_init(&bars->tmem_empty[i], kEpiWarps);
    }
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc<kTmemCols>(&bars->tmem_base);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  return bars;
}

__device__ __forceinline__ void teardown_pipeline(PipeBarriers* bars) {
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  if ((threadIdx.x >> 5) == 2) tc::tmem_dealloc<kTmemCols>(bars->tmem_base);
}

// ------------------------------------------------------------------------------------------
// K1: GEGLU up-projection
// ------------------------------------------------------------------------------------------
struct GegluArgs {
  const float* b1;                 // [2h] or null
  const uint8_t* neuron_override;  // [h] or null
  float override_value;
  __nv_bfloat16* H;                // [T, h]
  float* scores;                   // [T, E] or null
  __nv_bfloat16* gate_out;         // [T, h] or null
  int h, E, es, nv;                // nv = neuron pairs per tile (tile_n = 2 nv)
};

template <int CH, int ACT>
__device__ __forceinline__ void geglu_epilogue_tile(const GegluArgs& a, uint32_t taddr, int row, bool row_ok,
                                                    int n_tile0, int col_begin, int col_end) {
  // this thread owns token `row`; columns [col_begin, col_end) of the tile are whole experts
  for (int c0 = col_begin; c0 < col_end; c0 += a.es) {
    float score = 0.f;
    const int n_exp = n_tile0 + c0;  // first packed neuron of this expert
    for (int c = 0; c < a.es; c += CH) {
      uint32_t v[CH], g[CH];
      tc::tmem_ld_cols<CH>(taddr + c0 + c, v);
      tc::tmem_ld_cols<CH>(taddr + a.nv + c0 + c, g);
      tc::tmem_ld_wait();
      const int n = n_exp + c;
      uint32_t hw[CH / 2], gw[CH / 2];
#pragma unroll
      for (int i = 0; i < CH; i += 2) {
        float g0 = __uint_as_float(g[i]), g1 = __uint_as_float(g[i + 1]);
        float v0 = __uint_as_float(v[i]), v1 = __uint_as_float(v[i + 1]);
        if (a.b1 != nullptr) {
          g0 += __ldg(a.b1 + a.h + n + i);
          g1 += __ldg(a.b1 + a.h + n + i + 1);
          v0 += __ldg(a.b1 + n + i);
          v1 += __ldg(a.b1 + n + i + 1);
        }
        g0 = activate<ACT>(g0);
        g1 = activate<ACT>(g1);
        if (a.neuron_override != nullptr) {
          if (__ldg(a.neuron_override + n + i)) g0 = a.override_value;
          if (__ldg(a.neuron_override + n + i + 1)) g1 = a.override_value;
        }
        score += g0;
        score += g1;
        hw[i / 2] = pack_bf16x2(v0 * g0, v1 * g1);
        gw[i / 2] = pack_bf16x2(g0, g1);
      }
      if (row_ok) {
        store_words<CH / 2>(a.H + static_cast<size_t>(row) * a.h + n, hw);
        if (a.gate_out != nullptr) store_words<CH / 2>(a.gate_out + static_cast<size_t>(row) * a.h + n, gw);
      }
    }
    if (a.scores != nullptr && row_ok) a.scores[static_cast<size_t>(row) * a.E + n_exp / a.es] = score;
  }
}

template <int CH>
__global__ void __launch_bounds__(kNumThreads, 1)
geglu_up_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                const GemmShape g, const GegluArgs a, const int act) {
  uint8_t* smem;
  PipeBarriers* bars = setup_pipeline(smem, g, &tmap_x, &tmap_w1);
  const uint32_t tmem_base = bars->tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) producer_loop<true>(&tmap_x, &tmap_w1, smem, bars, g, a.h);
  } else if (warp == 1) {
    if (lane == 0) mma_loop(smem, bars, g, tmem_base);
  } else if (warp >= kEpiWarp0) {
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int cg = (warp - kEpiWarp0) >> 2;      // column group (0/1)
    const int cols_per_group = a.nv / 2;
    const int total = g.m_tiles * g.n_tiles;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int m_blk = tile / g.n_tiles, n_blk = tile % g.n_tiles;
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1u;
      tc::mbar_wait(&bars->tmem_full[as], aph);
      tc::fence_after_thread_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + as * kAccStride;
      const int row = m_blk * kBlockM + 32 * q + lane;
      const bool row_ok = row < g.rows;
      const int cb = cg * cols_per_group, ce = cb + cols_per_group;
      if (act == MOE_ACT_GELU)
        geglu_epilogue_tile<CH, MOE_ACT_GELU>(a, taddr, row, row_ok, n_blk * a.nv, cb, ce);
      else
        geglu_epilogue_tile<CH, MOE_ACT_RELU>(a, taddr, row, row_ok, n_blk * a.nv, cb, ce);
      tc::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->tmem_empty[as]);
    }
  }
  teardown_pipeline(bars);
}

// ------------------------------------------------------------------------------------------
// K3: down-projection
// ------------------------------------------------------------------------------------------
struct DownArgs {
  const float* b2;     // [d] or null
  __nv_bfloat16* Y;    // [T, d]
  int d;
};

template <int CH>
__global__ void __launch_bounds__(kNumThreads, 1)
down_proj_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_w2,
                 const GemmShape g, const DownArgs a) {
  uint8_t* smem;
  PipeBarriers* bars = setup_pipeline(smem, g, &tmap_h, &tmap_w2);
  const uint32_t tmem_base = bars->tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) producer_loop<false>(&tmap_h, &tmap_w2, smem, bars, g, 0);
  } else if (warp == 1) {
    if (lane == 0) mma_loop(smem, bars, g, tmem_base);
  } else if (warp >= kEpiWarp0) {
    const int q = warp & 3;
    const int cg = (warp - kEpiWarp0) >> 2;
    const int cols_per_group = g.tile_n / 2;
    const int total = g.m_tiles * g.n_tiles;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int m_blk = tile / g.n_tiles, n_blk = tile % g.n_tiles;
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1u;
      tc::mbar_wait(&bars->tmem_full[as], aph);
      tc::fence_after_thread_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + as * kAccStride;
      const int row = m_blk * kBlockM + 32 * q + lane;
      const bool row_ok = row < g.rows;
      for (int c = cg * cols_per_group; c < (cg + 1) * cols_per_group; c += CH) {
        uint32_t acc[CH];
        tc::tmem_ld_cols<CH>(taddr + c, acc);
        tc::tmem_ld_wait();
        const int n = n_blk * g.tile_n + c;
        uint32_t yw[CH / 2];
#pragma unroll
        for (int i = 0; i < CH; i += 2) {
          float y0 = __uint_as_float(acc[i]), y1 = __uint_as_float(acc[i + 1]);
          if (a.b2 != nullptr) {  // indices clamped: the last tile may overhang d
            y0 += __ldg(a.b2 + min(n + i, a.d - 1));
            y1 += __ldg(a.b2 + min(n + i + 1, a.d - 1));
          }
          yw[i / 2] = pack_bf16x2(y0, y1);
        }
        if (row_ok) {
          __nv_bfloat16* dst = a.Y + static_cast<size_t>(row) * a.d + n;
          if (n + CH <= a.d) {
            store_words<CH / 2>(dst, yw);
          } else {  // last tile of a d that is not a multiple of the tile width
#pragma unroll
            for (int i = 0; i < CH / 2; ++i)
              if (n + 2 * i < a.d) *reinterpret_cast<uint32_t*>(dst + 2 * i) = yw[i];
          }
        }
      }
      tc::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->tmem_empty[as]);
    }
  }
  teardown_pipeline(bars);
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
static int fill_shape(GemmShape& g, int rows, int k, int n_total, int tile_n, int b_rows) {
  g.rows = rows;
  g.k = k;
  g.m_tiles = (rows + kBlockM - 1) / kBlockM;
  g.n_tiles = (n_total + b_rows - 1) / b_rows;
  g.tile_n = tile_n;
  g.b_rows = b_rows;
  g.stage_bytes = kABytes + tile_n * 128;
  int stages = (kSmemLimit - 1024 - static_cast<int>(sizeof(PipeBarriers))) / g.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return -1;
  g.stages = stages;
  return 0;
}

static size_t smem_bytes(const GemmShape& g) {
  return static_cast<size_t>(g.stages) * g.stage_bytes + 1024 + sizeof(PipeBarriers);
}

// opt in to the full 227 KiB of dynamic shared memory, once per kernel (keyed by entry address:
// all instantiations of one template share a function-pointer TYPE)
static int ensure_smem(const void* kfn, size_t bytes) {
  static const void* configured[32];
  static int n_configured = 0;
  if (bytes > static_cast<size_t>(kSmemLimit)) return fail(MOE_ERR_UNSUPPORTED_SHAPE, "smem request %zu too large", bytes);
  for (int i = 0; i < n_configured; ++i)
    if (configured[i] == kfn) return MOE_OK;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e != cudaSuccess) return fail(MOE_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", kSmemLimit, cudaGetErrorString(e));
  if (n_configured < 32) configured[n_configured++] = kfn;
  return MOE_OK;
}

}  // namespace moe

extern "C" {

int moe_geglu_up(const void* x, const void* w1p, const float* b1p, const uint8_t* neuron_override,
                 float override_value, void* H, float* scores, void* gate_out, int T, int d, int h, int E,
                 int es, int act, void* stream) {
  using namespace moe;
  MOE_REQUIRE(x && w1p && H, MOE_ERR_INVALID_ARGUMENT, "moe_geglu_up: NULL x / w1p / H");
  MOE_REQUIRE(T >= 0 && d >= 8 && h >= 8 && E >= 1 && es >= 1, MOE_ERR_INVALID_ARGUMENT,
              "moe_geglu_up: bad sizes T=%d d=%d h=%d E=%d es=%d", T, d, h, E, es);
  MOE_REQUIRE(static_cast<long long>(E) * es == h, MOE_ERR_INVALID_ARGUMENT, "moe_geglu_up: E*es=%d*%d != h=%d", E, es, h);
  MOE_REQUIRE(d % 8 == 0 && h % 8 == 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_geglu_up: d=%d and h=%d must be multiples of 8", d, h);
  MOE_REQUIRE(act == MOE_ACT_GELU || act == MOE_ACT_RELU, MOE_ERR_INVALID_ARGUMENT, "moe_geglu_up: act=%d", act);
  MOE_REQUIRE(es % 4 == 0 && es <= 64, MOE_ERR_UNSUPPORTED_SHAPE,
              "moe_geglu_up: expert size %d unsupported (needs es %% 4 == 0 and es <= 64)", es);
  MOE_REQUIRE(((reinterpret_cast<uintptr_t>(H) | reinterpret_cast<uintptr_t>(gate_out)) & 15) == 0,
              MOE_ERR_INVALID_ARGUMENT, "moe_geglu_up: H / gate_out must be 16-byte aligned");
  if (T == 0) return MOE_OK;
  // neuron pairs per tile: a multiple of 2*es (two epilogue column groups of whole experts) and of 8,
  // dividing h, at most 128 (UMMA N = 2*nv <= 256); largest wins.
  int nv = 0;
  for (int cand = 128; cand >= 8; cand -= 8)
    if (cand % (2 * es) == 0 && h % cand == 0) {
      nv = cand;
      break;
    }
  MOE_REQUIRE(nv > 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_geglu_up: no tile width for es=%d h=%d", es, h);
  const int ch = (es % 32 == 0) ? 32 : (es % 20 == 0) ? 20 : (es % 16 == 0) ? 16 : (es % 8 == 0) ? 8 : 4;

  GemmShape g;
  MOE_REQUIRE(fill_shape(g, T, d, h, 2 * nv, nv) == 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_geglu_up: tile does not fit smem");
  CUtensorMap tx, tw;
  int rc = make_tmap_bf16_2d(&tx, x, static_cast<uint64_t>(T), static_cast<uint64_t>(d), kBlockM, kBlockK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tw, w1p, static_cast<uint64_t>(2) * h, static_cast<uint64_t>(d), static_cast<uint32_t>(nv), kBlockK);
  if (rc) return rc;
  GegluArgs a;
  a.b1 = b1p;
  a.neuron_override = neuron_override;
  a.override_value = override_value;
  a.H = static_cast<__nv_bfloat16*>(H);
  a.scores = scores;
  a.gate_out = static_cast<__nv_bfloat16*>(gate_out);
  a.h = h;
  a.E = E;
  a.es = es;
  a.nv = nv;
  const int total = g.m_tiles * g.n_tiles;
  const int grid = total < sm_count() ? total : sm_count();
  const size_t smem = smem_bytes(g);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MOE_LAUNCH_GEGLU(CHV)                                              \
  do {                                                                     \
    rc = ensure_smem(reinterpret_cast<const void*>(&geglu_up_kernel<CHV>), smem);                          \
    if (rc) return rc;                                                     \
    geglu_up_kernel<CHV><<<grid, kNumThreads, smem, st>>>(tx, tw, g, a, act); \
  } while (0)
  switch (ch) {
    case 32: MOE_LAUNCH_GEGLU(32); break;
    case 20: MOE_LAUNCH_GEGLU(20); break;
    case 16: MOE_LAUNCH_GEGLU(16); break;
    case 8: MOE_LAUNCH_GEGLU(8); break;
    default: MOE_LAUNCH_GEGLU(4); break;
  }
#undef MOE_LAUNCH_GEGLU
  return check_launch("moe_geglu_up");
}

int moe_down_proj(const void* H, const void* w2p, const float* b2, void* Y, int T, int h, int d, void* stream) {
  using namespace moe;
  MOE_REQUIRE(H && w2p && Y, MOE_ERR_INVALID_ARGUMENT, "moe_down_proj: NULL H / w2p / Y");
  MOE_REQUIRE(T >= 0 && h >= 8 && d >= 8, MOE_ERR_INVALID_ARGUMENT, "moe_down_proj: bad sizes T=%d h=%d d=%d", T, h, d);
  MOE_REQUIRE(h % 8 == 0 && d % 8 == 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_down_proj: h=%d and d=%d must be multiples of 8", h, d);
  MOE_REQUIRE((reinterpret_cast<uintptr_t>(Y) & 15) == 0, MOE_ERR_INVALID_ARGUMENT, "moe_down_proj: Y must be 16-byte aligned");
  if (T == 0) return MOE_OK;
  // tile width: a multiple of 16, <= 256 (the last tile may overhang d: TMA zero-fills the missing
  // weight rows and the epilogue guards its stores); pick the best wave-quantised cost
  const int m_tiles = (T + kBlockM - 1) / kBlockM;
  const int sms = sm_count();
  int best = 0;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 16; bn -= 16) {
    if (bn - (d % bn ? d % bn : bn) >= 16) continue;  // never waste a whole 16-column group
    const long long tiles = static_cast<long long>(m_tiles) * ((d + bn - 1) / bn);
    const long long waves = (tiles + sms - 1) / sms;
    // per-tile MMA time ~ max(bn/2, smem-bound (4096 + 32 bn)/128) cycles per K step
    const double per_tile = bn / 2.0 > (4096.0 + 32.0 * bn) / 128.0 ? bn / 2.0 : (4096.0 + 32.0 * bn) / 128.0;
    const double cost = waves * per_tile;
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  MOE_REQUIRE(best > 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_down_proj: no tile width for d=%d", d);
  const int half = best / 2;
  const int ch = (half % 32 == 0) ? 32 : (half % 16 == 0) ? 16 : 8;
  GemmShape g;
  MOE_REQUIRE(fill_shape(g, T, h, d, best, best) == 0, MOE_ERR_UNSUPPORTED_SHAPE, "moe_down_proj: tile does not fit smem");
  CUtensorMap th, tw;
  int rc = make_tmap_bf16_2d(&th, H, static_cast<uint64_t>(T), static_cast<uint64_t>(h), kBlockM, kBlockK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tw, w2p, static_cast<uint64_t>(d), static_cast<uint64_t>(h), static_cast<uint32_t>(best), kBlockK);
  if (rc) return rc;
  DownArgs a;
  a.b2 = b2;
  a.Y = static_cast<__nv_bfloat16*>(Y);
  a.d = d;
  const int total = g.m_tiles * g.n_tiles;
  const int grid = total < sms ? total : sms;
  const size_t smem = smem_bytes(g);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MOE_LAUNCH_DOWN(CHV)                                          \
  do {                                                                \
    rc = ensure_smem(reinterpret_cast<const void*>(&down_proj_kernel<CHV>), smem);                    \
    if (rc) return rc;                                                \
    down_proj_kernel<CHV><<<grid, kNumThreads, smem, st>>>(th, tw, g, a); \
  } while (0)
  switch (ch) {
    case 32: MOE_LAUNCH_DOWN(32); break;
    case 16: MOE_LAUNCH_DOWN(16); break;
    default: MOE_LAUNCH_DOWN(8); break;
  }
#undef MOE_LAUNCH_DOWN
  return check_launch("moe_down_proj");
}

}  // extern "C"
