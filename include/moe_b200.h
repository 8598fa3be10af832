/*
 * moe_b200.h -- C ABI of libmoe_b200.so: the B200 (sm_100a) implementation of the MoEfied
 * GEGLU feed-forward hot path of ruchikachavhan/diffusion-models-moe.
 *
 * The reference has no FFI of its own (it is pure Python calling ATen from torch forward
 * hooks), so each entry point below names the reference hook code it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch) unless marked host;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs;
 *   - return value: 0 = ok, <0 = error (MOE_ERR_*); moe_last_error() gives the message of the
 *     last failing call on the calling thread;
 *   - bf16 tensors are row-major contiguous; "packed" neuron order means the inner (h)
 *     dimension has been permuted so that expert e owns neurons [e*es, (e+1)*es)
 *     (moe_b200.packing / helper.modify_ffn_to_experts do this once per model);
 *   - T = B*S tokens, d = model dim, h = GEGLU inner dim, E experts of es neurons (E*es == h),
 *     W = ceil(E/32) 32-bit words per token for expert bit sets (bit e%32 of word e/32).
 */
#ifndef MOE_B200_H
#define MOE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MOE_API __attribute__((visibility("default")))
#else
#define MOE_API
#endif

#define MOE_OK 0
#define MOE_ERR_INVALID_ARGUMENT (-1)
#define MOE_ERR_UNSUPPORTED_SHAPE (-2)
#define MOE_ERR_CUDA (-3)
#define MOE_ERR_NO_DEVICE (-4)

#define MOE_ACT_GELU 0 /* exact erf GELU (upstream diffusers GEGLU.gelu) */
#define MOE_ACT_RELU 1 /* sparsity/relufy_model.py:28-40 */

/* ABI version; bumped whenever a signature changes. */
MOE_API int moe_abi_version(void);
/* Message of the last error on this thread ("" if none). Host pointer, valid until the next call. */
MOE_API const char* moe_last_error(void);
/* Number of kernels this library has launched since load / since the last reset (bench bookkeeping). */
MOE_API long long moe_launch_count(void);
MOE_API void moe_reset_launch_count(void);

/*
 * K1 -- GEGLU up-projection with the gate activation, the per-neuron override and the
 * per-expert score segment-sum fused into the epilogue (tcgen05 / TMEM / TMA).
 *
 *   Y = x W1p^T + b1p;  v = Y[:, 0:h], g = act(Y[:, h:2h]);  g[:, n] = override_value where
 *   neuron_override[n] != 0;  scores[t, e] = sum_{n in expert e} g[t, n];  H = v * g.
 *
 * Replaces: neuron_receivers/moefy.py:11-13,18-20 (module.proj, chunk, module.gelu,
 * matmul(gate, patterns^T)); remove_skilled_neurons.py:30-42 (override = -0.17);
 * expert_activation.py:48-56; frequency_measure.py:44-51.
 *
 *   x      bf16 [T, d]        w1p  bf16 [2h, d]  (value rows first, then gate rows; packed order)
 *   b1p    f32  [2h] or NULL  neuron_override u8 [h] or NULL
 *   H      bf16 [T, h] out    scores f32 [T, E] out or NULL     gate_out bf16 [T, h] out or NULL
 * Constraints: d % 8 == 0, h % 8 == 0, E*es == h; the tile width is chosen from es
 * (lcm(es,8)-multiples); unsupported geometries return MOE_ERR_UNSUPPORTED_SHAPE.
 */
MOE_API int moe_geglu_up(const void* x, const void* w1p, const float* b1p, const uint8_t* neuron_override,
                 float override_value, void* H, float* scores, void* gate_out, int T, int d, int h, int E,
                 int es, int act, void* stream);

/*
 * K2 -- router: per-token top-k select over E expert scores, fused with the selection histogram, the score
 * column-max and the in-place zeroing of H for unselected / removed experts.  E <= 256: 4 / 8 / 16 lanes per token
 * (several tokens per warp, warp shuffles), exact k-th largest order-preserving key from a bitonic sort; larger E: one
 * warp per token, radix select over ballots.  Ties go to the lower expert id.
 *
 * Replaces: moefy.py:21-23 (topk, embedding(labels, patterns).sum, gate[mask==0]=0);
 * frequency_measure.py:52-57 (Python counter loop; integer counts instead of += 1/S);
 * remove_skilled_experts.py:29-49 (removed experts score exactly 0, still compete, own no
 * neurons); expert_activation.py:57 (max over tokens).
 *
 *   scores        f32 [T, E]
 *   removed_bits  u32 [W] or NULL: experts whose pattern row is zeroed for this (timestep, layer)
 *   active_bits   u32 [T, W] out or NULL: selected AND NOT removed (what owns live neurons)
 *   idx           i16 [T, k] out or NULL: selected expert ids, ascending (includes removed ones
 *                 that took a slot, exactly as the reference's `labels`)
 *   hist          u64 [E] accumulated (+= 1 per (token, selected expert)) for tokens in
 *                 [count_begin, count_end), or NULL
 *   score_colmax  f32 [E] max-accumulated over all T tokens (caller pre-fills with -inf), or NULL
 *   H             bf16 [T, h] zeroed in place where the neuron's expert is not active, or NULL
 */
MOE_API int moe_router_topk(const float* scores, const uint32_t* removed_bits, int k, uint32_t* active_bits,
                    int16_t* idx, unsigned long long* hist, float* score_colmax, void* H, int h, int es,
                    int T, int E, int count_begin, int count_end, void* stream);

/* Same with a per-expert bias added to the scores before the selection (score_bias f32 [E] or NULL; the column max
 * is taken on the unbiased scores).  Replaces add_skilled_experts.py:53-58 (score[:, idx] += 5 * std[idx]; top-k). */
MOE_API int moe_router_topk_biased(const float* scores, const float* score_bias, const uint32_t* removed_bits, int k,
                    uint32_t* active_bits, int16_t* idx, unsigned long long* hist, float* score_colmax, void* H,
                    int h, int es, int T, int E, int count_begin, int count_end, void* stream);

/*
 * Compacted token -> expert permutation (the north star's second router output).  From the expert-set words of
 * moe_router_topk: for every expert e the ascending list of the tokens whose active set contains e,
 *   perm_tokens[perm_offsets[e] + i], 0 <= i < perm_counts[e];
 * every list is padded to a multiple of `row_pad` rows (pad entries -1; perm_offsets[E] = total padded rows), and
 *   slot_pos[t * k + j] = row of token t's j-th active expert (ascending expert id), -1 for j >= its active count.
 * A stable counting sort without atomics: the lists are deterministic and in token order.
 * Replaces: the [T, k, h] gather `F.embedding(labels, patterns)` of moefy.py:22 as the data structure that says which
 * neurons of which token are live -- here E lists of token ids instead of T*k one-hot rows.
 *   active_bits u32 [T, W]   perm_offsets i32 [E + 1] out   perm_counts i32 [E] out
 *   perm_tokens i32 [moe_down_grouped_rows(T, k, E)] out (for row_pad <= 128)   slot_pos i32 [T, k] out
 *   workspace: moe_expert_permutation_workspace_bytes(T, E) bytes of device scratch (no initialisation needed)
 * moe_router_topk_perm = moe_router_topk (bits + labels + histogram) followed by the permutation, one call.
 */
MOE_API int moe_expert_permutation(const uint32_t* active_bits, int T, int E, int k, int row_pad, int* perm_offsets,
                    int* perm_counts, int* perm_tokens, int* slot_pos, void* workspace, size_t workspace_bytes, void* stream);
MOE_API size_t moe_expert_permutation_workspace_bytes(int T, int E);
MOE_API int moe_router_topk_perm(const float* scores, const uint32_t* removed_bits, int k, uint32_t* active_bits, int16_t* idx,
                    unsigned long long* hist, int* perm_offsets, int* perm_counts, int* perm_tokens, int* slot_pos,
                    int row_pad, int T, int E, int count_begin, int count_end, void* workspace, size_t workspace_bytes,
                    void* stream);

/*
 * Grouped / gathered down-projection: Y[t] = b2 + sum over the token's active experts e of
 * H[t, e*es : (e+1)*es] W2p[:, e*es : (e+1)*es]^T -- only the active experts' slices of H and W2 are read (tcgen05,
 * one 128-token tile of one expert's list per CTA, A rows gathered with cp.async, W2 slice by TMA), fp32 partial rows
 * combined per token in ascending expert order (deterministic).  Removed experts are absent from the lists, so the
 * removal mask is applied by construction.  H need not be masked.
 * Replaces: moefy.py:22-23 + the stock ff.net.2 Linear [upstream] (mask the gate, then a dense down-projection).
 *   H bf16 [T, h]   perm_offsets / perm_tokens / slot_pos: from moe_expert_permutation with row_pad = 128
 *   w2p bf16 [d, h]   b2 f32 [d] or NULL   Y bf16 [T, d] out
 *   workspace: moe_down_grouped_workspace_bytes(T, k, E, d) bytes (fp32 partial rows; no initialisation needed)
 * Requirements: es == 64 (one expert = one 64-wide bf16 k-block; BASELINE.json configs[0] geometry), d % 16 == 0;
 * otherwise MOE_ERR_UNSUPPORTED_SHAPE (use moe_down_proj on the masked H).
 */
MOE_API int moe_down_grouped(const void* H, const int* perm_offsets, const int* perm_tokens, const int* slot_pos,
                    const void* w2p, const float* b2, void* Y, int T, int h, int d, int E, int es, int k, void* workspace,
                    size_t workspace_bytes, void* stream);
MOE_API size_t moe_down_grouped_rows(int T, int k, int E);
MOE_API size_t moe_down_grouped_workspace_bytes(int T, int k, int E, int d);

/*
 * K3 -- down-projection Y = H W2p^T + b2 (tcgen05 / TMEM / TMA).  H has already been zeroed
 * for inactive experts by K2, so this is the dense-masked form; W2p may be the
 * Wanda-masked copy produced by moe_mask_weights.
 * Replaces: upstream FeedForward.net[2] (Linear(h, d)) and
 * remove_wanda_neurons_fast.py:69-83 (F.linear(x, W2*(1-M), b2)).
 *   H bf16 [T, h]   w2p bf16 [d, h] (columns in packed order)   b2 f32 [d] or NULL   Y bf16 [T, d] out
 * Small-T shapes split the K loop over several CTAs: fp32 partial tiles go to `workspace`, the last slice to
 * arrive sums them in a fixed order (deterministic), adds b2 and writes Y.
 */
MOE_API int moe_down_proj(const void* H, const void* w2p, const float* b2, void* Y, int T, int h, int d,
                  void* workspace, size_t workspace_bytes, void* stream);
/* K3 with an elementwise weight mask applied on the fly: Y = H (W2 * (1 - M))^T + b2, M given as bit words over the
 * row-major [d, h] weight (bit r*h + c; the layout of moe_mask_pack / moe_mask_union / moe_wanda_score_mask).  The
 * epilogue warps zero the masked weights of every landed W2 tile in shared memory, between the TMA and the tensor core,
 * so no masked copy of W2 is ever written: one launch, DRAM traffic = W2 + d*h/8 bytes of mask.
 * Replaces: remove_wanda_neurons_fast.py:72-77 (H2D of a dense int64 mask, W.clone(), W*(1-mask), F.linear) per call.
 * Requirements: h % 64 == 0, mask_bits 8-byte aligned; otherwise MOE_ERR_UNSUPPORTED_SHAPE and the caller uses
 * moe_mask_weights + moe_down_proj.  Workspace: as moe_down_proj. */
MOE_API int moe_down_proj_masked(const void* H, const void* w2p, const uint32_t* mask_bits, const float* b2, void* Y, int T,
                  int h, int d, void* workspace, size_t workspace_bytes, void* stream);
/* Recommended size of the optional split-K workspace of moe_down_proj (device memory, 16-byte aligned).
 * Its first 64 KiB are per-tile arrival counters: the caller zero-fills the buffer ONCE after allocating it;
 * the kernel leaves the counters at zero.  One workspace per concurrently running stream.  With
 * workspace == NULL the kernel never splits K (slow for T <= 512 with h >= 2560). */
MOE_API size_t moe_down_proj_workspace_bytes(int T, int h, int d);

/*
 * K4 -- standalone expert-frequency histogram over stored labels:
 * hist[e] += #occurrences of e in idx[0:n] (16-byte loads, per-warp shared-memory bins updated with
 * native shared atomics, one 64-bit global atomic per bin per CTA).
 * Replaces: frequency_measure.py:53-57 applied to saved labels; the quantity that
 * freq_expert_select.py:61-64 averages and that is all-reduced across GPUs.
 */
MOE_API int moe_hist_accumulate(const int16_t* idx, long long n, int E, unsigned long long* hist, void* stream);

/* Column max over tokens of a [T, C] f32 / bf16 matrix, max-accumulated into out[C]
 * (expert_activation.py:57 on scores; predictivity.py:49 on act(gate)). */
MOE_API int moe_colmax_f32(const float* m, int T, int C, float* out, void* stream);
MOE_API int moe_colmax_bf16(const void* m, int T, int C, float* out, void* stream);

/*
 * Removal-mask utilities (remove_wanda_neurons_fast.py:13-29,72-77; multi_concept_remover.py:43-53).
 *   moe_mask_pack:    dense u8 0/1 [n] -> bits u32 [ceil(n/32)]        (n % 32 tail handled)
 *   moe_mask_union:   out = a | b over n_words words (out may alias a or b)
 *   moe_mask_weights: W2m[r, c] = bit(r*h + c) ? 0 : W2[r, c]  for a [d, h] bf16 matrix
 *                     (bits index the row-major flattening; requires h % 32 == 0)
 */
MOE_API int moe_mask_pack(const uint8_t* dense, long long n, uint32_t* bits, void* stream);
MOE_API int moe_mask_union(const uint32_t* a, const uint32_t* b, uint32_t* out, long long n_words, void* stream);
MOE_API int moe_mask_weights(const void* w2, const uint32_t* bits, void* w2m, int d, int h, void* stream);

/* Column sums out[c] += sum_t m[t, c] over the rows with row_mask[t % period] != 0 (row_mask u8 [period] or NULL =
 * all rows).  Replaces get_experts.py:64-77 (score, optionally restricted to the bounding-box tokens of every batch
 * row, averaged over tokens before the top-k).  out f32 [C], accumulated (caller zero-fills). */
MOE_API int moe_colsum_f32(const float* m, int T, int C, const uint8_t* row_mask, int period, float* out, void* stream);
/* out[n] += sum_t (H[t, n] / max(||H[t, :]||_2, 1e-12))^2 -- the squared column norms of the row-normalised hidden
 * state.  Replaces wanda_receiver.py:47-53 + utils.py:330-337 (F.normalize(out, p=2, dim=1) then the incremental
 * column norm sqrt(old^2 + new^2)): keep the squares on the device, take the root when the norms are read.
 * H bf16 [T, h] (h % 8 == 0); out f32 [h], accumulated. */
MOE_API int moe_rownorm_colsumsq_bf16(const void* H, int T, int h, float* out, void* stream);

/* Wanda scoring of one (timestep, layer): bit (r*h + c) of `bits` := (|W2[r,c]| * norm_adj[c] > |W2[r,c]| * norm_base[c])
 * AND c is among the k = int(ratio * h) largest |W2[r,:]| * norm_adj of row r (ties on the k-th value: lowest column).
 * Replaces modularity/wanda.py:143-165 (two row-wise torch.sort over [d, h], scatter_, compare, CSR pickle).
 * w2 bf16 [d, h]; norm_* f32 [h]; bits u32 [d*h/32] out (the layout moe_mask_weights / moe_mask_union use). */
MOE_API int moe_wanda_score_mask(const void* w2, const float* norm_base, const float* norm_adj, int d, int h, int k,
                    uint32_t* bits, void* stream);
/* Union over timesteps by vote: out bit i := #{t : bit i of masks[t]} > threshold.  Replaces
 * benchmarks/save_union_over_time.py:192-205 (sum of T CSR masks > select_ratio * T).  masks u32 [T][n_words]. */
MOE_API int moe_mask_vote(const uint32_t* masks, int T, long long n_words, float threshold, uint32_t* out, void* stream);

/* The sampler step around the UNet call, one streaming kernel: classifier-free-guidance combine
 * and DDIM (eta = 0) update, x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev) eps with eps = eps_u + guidance (eps_c - eps_u)
 * and x0 = (x - sqrt(1 - a_t) eps) / sqrt(a_t).  Replaces the chain of elementwise ATen kernels of the stock pipeline
 * [upstream StableDiffusionPipeline.__call__ + DDIMScheduler.step, driven by base_receiver.py:73].  All four tensors
 * have n elements, f32 (is_bf16 = 0) or bf16 (is_bf16 = 1); x_prev may alias x.  alpha_* = cumulative alpha products. */
MOE_API int moe_cfg_ddim_step(const void* eps_uncond, const void* eps_cond, const void* x, void* x_prev, long long n,
                    int is_bf16, float guidance, float alpha_t, float alpha_prev, void* stream);

/*
 * Fused layer call -- the whole MoEfied GEGLU FFN of one BasicTransformerBlock in ONE persistent kernel:
 * up-projection + activation + product + expert scores (as moe_geglu_up), per-token top-k routing with the
 * removed-expert rule, histogram and write-only masking of H (as moe_router_topk), down-projection (as
 * moe_down_proj).  CTA pairs pass row blocks from phase to phase through counters in `workspace`, so the three
 * stages overlap and the layer pays one launch ramp instead of three.
 *
 * Replaces, for one hooked layer call: moefy.py:10-27 / remove_skilled_experts.py:29-49 /
 * frequency_measure.py:40-64 (hook body) followed by the stock ff.net.2 Linear [upstream].
 *
 *   x bf16 [T, d]   w1p bf16 [2h, d]   b1p f32 [2h] or NULL   w2p bf16 [d, h]   b2 f32 [d] or NULL
 *   H bf16 [T, h] out (masked hidden state)   scores f32 [T, E] out   Y bf16 [T, d] out
 *   removed_bits / k / active_bits / idx / hist / count_begin / count_end: as moe_router_topk
 *   act: MOE_ACT_GELU | MOE_ACT_RELU     mask_h: 0 = leave H unmasked (routing outputs only)
 *   workspace: device memory of moe_ffn_fused_workspace_bytes() bytes, 16-byte aligned, ZERO-FILLED ONCE after
 *              allocation (the kernel leaves its counters at zero); one workspace per concurrently running stream
 * Requirements: d % 64 == 0, h % 64 == 0, es % 4 == 0, E <= 512 and an expert size the tile shapes support;
 * otherwise MOE_ERR_UNSUPPORTED_SHAPE is returned and the caller uses the three separate entry points.
 * The kernel occupies every SM with one CTA and its CTAs wait on each other: the launch is refused
 * (MOE_ERR_UNSUPPORTED_SHAPE) when cudaOccupancyMaxActiveClusters says the CTA pairs cannot all be resident (MPS
 * thread limits, green contexts), and fused launches of one process on different streams of a device are ordered one
 * after the other by the library (events) -- two such grids resident together would wait for SMs the other holds.
 * A kernel of another library may run concurrently as long as it finishes (the pairs that found no SM start when it
 * does); a cross-CTA wait of more than 2 s traps instead of hanging the GPU.  One process per GPU is the tested
 * configuration; per-device caches (SM count, kernel attributes, ordering events) are keyed by the current device.
 */
MOE_API int moe_ffn_fused(const void* x, const void* w1p, const float* b1p, const void* w2p, const float* b2, void* H,
                  float* scores, void* Y, const uint32_t* removed_bits, int k, uint32_t* active_bits, int16_t* idx,
                  unsigned long long* hist, int count_begin, int count_end, int T, int d, int h, int E, int es, int act,
                  int mask_h, void* workspace, size_t workspace_bytes, void* stream);
MOE_API size_t moe_ffn_fused_workspace_bytes(int T, int d, int h);
/* Profiling hook of the fused kernel (library built with -DMOE_TRACE=1 only): per-CTA %globaltimer stamps. */
MOE_API int moe_debug_trace_fused(unsigned long long* host_out, int n);

/* Profiling hook: with the environment variable MOE_DEBUG_MODE bit 16 set, the GEMM kernels record
 * per-CTA cycle counts of their producer / MMA-issue loops; this copies the first n counters
 * (8 per CTA: empty-wait, TMA-issue, iterations, acc-wait, full-wait, MMA-issue, commit, -) to a HOST
 * buffer.  Synchronises the device.  Not part of the hot path. */
MOE_API int moe_debug_counters(unsigned long long* host_out, int n);
/* Profiling hook: with MOE_DEBUG_MODE bit 32 set, each GEMM CTA stamps %globaltimer (ns) at fixed points
 * (64 slots per CTA: entry, set-up done, first load, first MMA, per-tile commit / epilogue begin / end, exit);
 * this copies the first n stamps to a HOST buffer.  Synchronises the device.  Not part of the hot path. */
MOE_API int moe_debug_trace(unsigned long long* host_out, int n);

#ifdef __cplusplus
}
#endif
#endif /* MOE_B200_H */
