"""Thin tensor -> pointer wrappers over the C ABI.  Every function enqueues on the current
CUDA stream of the tensors' device and returns without synchronising."""
from typing import Optional

import torch

from . import _lib

ACT_GELU = 0
ACT_RELU = 1


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need(t: torch.Tensor, dtype, name: str, shape=None) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU fallback exists for this path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")


def launch_count() -> int:
    return int(_lib.load().moe_launch_count())


def reset_launch_count() -> None:
    _lib.load().moe_reset_launch_count()


def geglu_up(x, w1p, b1p, n_experts: int, expert_size: int, act: int = ACT_GELU, *, neuron_override=None,
             override_value: float = -0.17, want_scores: bool = True, want_gate: bool = False, out=None,
             scores_out=None):
    """K1.  x bf16 [T, d]; w1p bf16 [2h, d] packed; b1p f32 [2h] or None.
    Returns (H bf16 [T, h], scores f32 [T, E] | None, gate bf16 [T, h] | None)."""
    lib = _lib.load()
    T, d = x.shape
    h = w1p.shape[0] // 2
    _need(x, torch.bfloat16, "x")
    _need(w1p, torch.bfloat16, "w1p", (2 * h, d))
    if b1p is not None:
        _need(b1p, torch.float32, "b1p", (2 * h,))
    if neuron_override is not None:
        _need(neuron_override, torch.uint8, "neuron_override", (h,))
    H = out if out is not None else torch.empty((T, h), dtype=torch.bfloat16, device=x.device)
    _need(H, torch.bfloat16, "out", (T, h))
    scores = None
    if want_scores:
        scores = scores_out if scores_out is not None else torch.empty((T, n_experts), dtype=torch.float32,
                                                                       device=x.device)
        _need(scores, torch.float32, "scores_out", (T, n_experts))
    gate = torch.empty((T, h), dtype=torch.bfloat16, device=x.device) if want_gate else None
    with torch.cuda.device(x.device):
        rc = lib.moe_geglu_up(_ptr(x), _ptr(w1p), _ptr(b1p), _ptr(neuron_override), float(override_value), _ptr(H),
                              _ptr(scores), _ptr(gate), T, d, h, n_experts, expert_size, act, _stream(x))
    _lib.check(rc, "moe_geglu_up")
    return H, scores, gate


def router_topk(scores, k: int, *, removed_bits=None, want_bits: bool = True, want_idx: bool = False, hist=None,
                colmax_out=None, H=None, expert_size: int = 0, count_rows=(0, 0), bits_out=None, score_bias=None):
    """K2.  scores f32 [T, E].  Returns (active_bits i32 [T, W] | None, idx i16 [T, k] | None).
    `hist` (int64 [E]) and `colmax_out` (f32 [E], pre-filled with -inf) are accumulated in place;
    `H` (bf16 [T, E*expert_size]) is zeroed in place for inactive experts."""
    lib = _lib.load()
    _need(scores, torch.float32, "scores")
    T, E = scores.shape
    W = (E + 31) // 32
    dev = scores.device
    if removed_bits is not None:
        _need(removed_bits, torch.int32, "removed_bits", (W,))
    bits = None
    if want_bits:
        bits = bits_out if bits_out is not None else torch.empty((T, W), dtype=torch.int32, device=dev)
        _need(bits, torch.int32, "bits_out", (T, W))
    idx = torch.empty((T, k), dtype=torch.int16, device=dev) if want_idx else None
    if hist is not None:
        _need(hist, torch.int64, "hist", (E,))
    if colmax_out is not None:
        _need(colmax_out, torch.float32, "colmax_out", (E,))
    h = 0
    if H is not None:
        h = E * expert_size
        _need(H, torch.bfloat16, "H", (T, h))
    if score_bias is not None:
        _need(score_bias, torch.float32, "score_bias", (E,))
    with torch.cuda.device(dev):
        rc = lib.moe_router_topk_biased(_ptr(scores), _ptr(score_bias), _ptr(removed_bits), int(k), _ptr(bits), _ptr(idx),
                                        _ptr(hist), _ptr(colmax_out), _ptr(H), h, int(expert_size), T, E,
                                        int(count_rows[0]), int(count_rows[1]), _stream(scores))
    _lib.check(rc, "moe_router_topk")
    return bits, idx


_workspaces = {}


def splitk_workspace(device, stream_ptr: int, nbytes: int) -> torch.Tensor:
    """Zero-initialised split-K workspace, one per device, grown on demand (allocate it during warm-up,
    before any CUDA-graph capture).  The kernel keeps its counter header at zero, so the buffer is
    zero-filled only when (re)allocated.  Down-projections issued concurrently on several streams of
    one device must not share it: serialise them or call the C ABI with per-stream workspaces."""
    key = torch.device(device).index
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def down_proj(H, w2p, b2, out=None, mask_bits=None):
    """K3.  H bf16 [T, h]; w2p bf16 [d, h]; b2 f32 [d] or None -> Y bf16 [T, d].
    `mask_bits` (int32 [d*h/32], bit r*h + c set = weight removed): Y = H (W2 * (1 - M))^T + b2 in ONE launch, the
    mask applied to the W2 tiles in shared memory (moe_down_proj_masked; h % 64 == 0)."""
    lib = _lib.load()
    T, h = H.shape
    d = w2p.shape[0]
    _need(H, torch.bfloat16, "H")
    _need(w2p, torch.bfloat16, "w2p", (d, h))
    if b2 is not None:
        _need(b2, torch.float32, "b2", (d,))
    if mask_bits is not None:
        _need(mask_bits, torch.int32, "mask_bits", (d * h // 32,))
    Y = out if out is not None else torch.empty((T, d), dtype=torch.bfloat16, device=H.device)
    _need(Y, torch.bfloat16, "out", (T, d))
    with torch.cuda.device(H.device):
        st = _stream(H)
        ws = splitk_workspace(H.device, st, int(lib.moe_down_proj_workspace_bytes(T, h, d)))
        if mask_bits is None:
            rc = lib.moe_down_proj(_ptr(H), _ptr(w2p), _ptr(b2), _ptr(Y), T, h, d, _ptr(ws), ws.numel(), st)
        else:
            rc = lib.moe_down_proj_masked(_ptr(H), _ptr(w2p), _ptr(mask_bits), _ptr(b2), _ptr(Y), T, h, d, _ptr(ws),
                                          ws.numel(), st)
    _lib.check(rc, "moe_down_proj")
    return Y


_fused_workspaces = {}


def fused_workspace(device, nbytes: int, stream_ptr: int = 0) -> torch.Tensor:
    """Zero-initialised workspace of the fused layer kernel (sync counters + split-K partials), one per (device,
    stream), grown on demand; the kernel leaves its counters at zero (allocate during warm-up, before graph capture).
    Fused launches on different streams of one device never share counters, and the library additionally orders them
    one after the other (two persistent grids must not be resident together, see moe_b200.h)."""
    key = (torch.device(device).index, int(stream_ptr))
    ws = _fused_workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _fused_workspaces[key] = ws
    return ws


def ffn_fused(x, w1p, b1p, w2p, b2, n_experts: int, expert_size: int, k: int, act: int = ACT_GELU, *,
              removed_bits=None, want_bits: bool = False, want_idx: bool = False, hist=None, count_rows=(0, 0),
              mask_h: bool = True, H_out=None, scores_out=None, out=None, bits_out=None):
    """One hooked layer call in one launch: K1 -> routing -> K3 (moe_ffn_fused).
    Returns (Y bf16 [T, d], H bf16 [T, h] masked, scores f32 [T, E], active_bits | None, idx | None)."""
    lib = _lib.load()
    T, d = x.shape
    h = w1p.shape[0] // 2
    E = n_experts
    W = (E + 31) // 32
    dev = x.device
    _need(x, torch.bfloat16, "x")
    _need(w1p, torch.bfloat16, "w1p", (2 * h, d))
    _need(w2p, torch.bfloat16, "w2p", (d, h))
    if b1p is not None:
        _need(b1p, torch.float32, "b1p", (2 * h,))
    if b2 is not None:
        _need(b2, torch.float32, "b2", (d,))
    if removed_bits is not None:
        _need(removed_bits, torch.int32, "removed_bits", (W,))
    H = H_out if H_out is not None else torch.empty((T, h), dtype=torch.bfloat16, device=dev)
    _need(H, torch.bfloat16, "H_out", (T, h))
    scores = scores_out if scores_out is not None else torch.empty((T, E), dtype=torch.float32, device=dev)
    _need(scores, torch.float32, "scores_out", (T, E))
    Y = out if out is not None else torch.empty((T, d), dtype=torch.bfloat16, device=dev)
    _need(Y, torch.bfloat16, "out", (T, d))
    bits = None
    if want_bits:
        bits = bits_out if bits_out is not None else torch.empty((T, W), dtype=torch.int32, device=dev)
        _need(bits, torch.int32, "bits_out", (T, W))
    idx = torch.empty((T, k), dtype=torch.int16, device=dev) if want_idx else None
    if hist is not None:
        _need(hist, torch.int64, "hist", (E,))
    with torch.cuda.device(dev):
        ws = fused_workspace(dev, int(lib.moe_ffn_fused_workspace_bytes(T, d, h)), _stream(x))
        rc = lib.moe_ffn_fused(_ptr(x), _ptr(w1p), _ptr(b1p), _ptr(w2p), _ptr(b2), _ptr(H), _ptr(scores), _ptr(Y),
                               _ptr(removed_bits), int(k), _ptr(bits), _ptr(idx), _ptr(hist), int(count_rows[0]),
                               int(count_rows[1]), T, d, h, E, int(expert_size), int(act), 1 if mask_h else 0, _ptr(ws),
                               ws.numel(), _stream(x))
    _lib.check(rc, "moe_ffn_fused")
    return Y, H, scores, bits, idx


class ExpertPermutation:
    """Compacted token -> expert permutation (moe_expert_permutation): `tokens[offsets[e] : offsets[e] + counts[e]]`
    = ascending tokens whose active set holds expert e (lists padded to 128 rows with -1); `slot_pos[t, j]` = row of
    token t's j-th active expert, -1 beyond its active count."""

    def __init__(self, offsets, counts, tokens, slot_pos, k):
        self.offsets, self.counts, self.tokens, self.slot_pos, self.k = offsets, counts, tokens, slot_pos, k


def expert_permutation(active_bits, n_experts: int, k: int, row_pad: int = 128) -> ExpertPermutation:
    """active_bits int32 [T, W] (router_topk(..., want_bits=True)) -> ExpertPermutation (all int32, on the device)."""
    lib = _lib.load()
    T, W = active_bits.shape
    _need(active_bits, torch.int32, "active_bits", (T, (n_experts + 31) // 32))
    dev = active_bits.device
    offsets = torch.empty(n_experts + 1, dtype=torch.int32, device=dev)
    counts = torch.empty(n_experts, dtype=torch.int32, device=dev)
    tokens = torch.empty(int(lib.moe_down_grouped_rows(T, k, n_experts)), dtype=torch.int32, device=dev)
    slot_pos = torch.empty((T, max(k, 1)), dtype=torch.int32, device=dev)[:, :k]
    ws = torch.empty(int(lib.moe_expert_permutation_workspace_bytes(T, n_experts)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.moe_expert_permutation(_ptr(active_bits), T, n_experts, int(k), int(row_pad), _ptr(offsets), _ptr(counts),
                                        _ptr(tokens), _ptr(slot_pos) if k > 0 else 0, _ptr(ws), ws.numel(), _stream(active_bits))
    _lib.check(rc, "moe_expert_permutation")
    return ExpertPermutation(offsets, counts, tokens, slot_pos, k)


_grouped_ws = {}


def down_grouped(H, perm: ExpertPermutation, w2p, b2, n_experts: int, expert_size: int, out=None):
    """Grouped / gathered down-projection over the active experts only.  H bf16 [T, h] (need not be masked);
    w2p bf16 [d, h]; b2 f32 [d] or None -> Y bf16 [T, d].  Expert size 64 only (MoeLibraryError code -2 otherwise)."""
    lib = _lib.load()
    T, h = H.shape
    d = w2p.shape[0]
    _need(H, torch.bfloat16, "H")
    _need(w2p, torch.bfloat16, "w2p", (d, h))
    if b2 is not None:
        _need(b2, torch.float32, "b2", (d,))
    Y = out if out is not None else torch.empty((T, d), dtype=torch.bfloat16, device=H.device)
    _need(Y, torch.bfloat16, "out", (T, d))
    nbytes = int(lib.moe_down_grouped_workspace_bytes(T, perm.k, n_experts, d))
    key = (H.device.index, _stream(H))
    ws = _grouped_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=H.device)
        _grouped_ws[key] = ws
    with torch.cuda.device(H.device):
        rc = lib.moe_down_grouped(_ptr(H), _ptr(perm.offsets), _ptr(perm.tokens), _ptr(perm.slot_pos) if perm.k > 0 else _ptr(perm.offsets),
                                  _ptr(w2p), _ptr(b2), _ptr(Y), T, h, d, n_experts, int(expert_size), int(perm.k), _ptr(ws),
                                  ws.numel(), _stream(H))
    _lib.check(rc, "moe_down_grouped")
    return Y


def hist_accumulate(idx, n_experts: int, hist=None):
    """K4.  idx int16 (any shape) -> hist int64 [E] (accumulated in place if given)."""
    lib = _lib.load()
    _need(idx, torch.int16, "idx")
    if hist is None:
        hist = torch.zeros(n_experts, dtype=torch.int64, device=idx.device)
    _need(hist, torch.int64, "hist", (n_experts,))
    with torch.cuda.device(idx.device):
        rc = lib.moe_hist_accumulate(_ptr(idx), idx.numel(), n_experts, _ptr(hist), _stream(idx))
    _lib.check(rc, "moe_hist_accumulate")
    return hist


def colmax(m, out=None):
    """Column max over rows of a [T, C] f32 / bf16 matrix, max-accumulated into out (f32 [C])."""
    lib = _lib.load()
    if m.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("colmax: f32 or bf16 only")
    _need(m, m.dtype, "m")
    T, C = m.shape
    if out is None:
        out = torch.full((C,), float("-inf"), dtype=torch.float32, device=m.device)
    _need(out, torch.float32, "out", (C,))
    fn = lib.moe_colmax_f32 if m.dtype == torch.float32 else lib.moe_colmax_bf16
    with torch.cuda.device(m.device):
        rc = fn(_ptr(m), T, C, _ptr(out), _stream(m))
    _lib.check(rc, "moe_colmax")
    return out


def colsum(m, out=None, row_mask=None):
    """Column sums over rows of a [T, C] f32 matrix, accumulated into out (f32 [C], zero-filled if not given);
    `row_mask` uint8 [period] keeps only rows t with row_mask[t % period] != 0."""
    lib = _lib.load()
    _need(m, torch.float32, "m")
    T, C = m.shape
    if out is None:
        out = torch.zeros(C, dtype=torch.float32, device=m.device)
    _need(out, torch.float32, "out", (C,))
    period = 0
    if row_mask is not None:
        _need(row_mask, torch.uint8, "row_mask")
        period = row_mask.numel()
    with torch.cuda.device(m.device):
        rc = lib.moe_colsum_f32(_ptr(m), T, C, _ptr(row_mask), period, _ptr(out), _stream(m))
    _lib.check(rc, "moe_colsum_f32")
    return out


def rownorm_colsumsq(H, out=None):
    """Squared column norms of the row-normalised bf16 matrix H [T, h], accumulated into out (f32 [h])."""
    lib = _lib.load()
    _need(H, torch.bfloat16, "H")
    T, h = H.shape
    if out is None:
        out = torch.zeros(h, dtype=torch.float32, device=H.device)
    _need(out, torch.float32, "out", (h,))
    with torch.cuda.device(H.device):
        rc = lib.moe_rownorm_colsumsq_bf16(_ptr(H), T, h, _ptr(out), _stream(H))
    _lib.check(rc, "moe_rownorm_colsumsq_bf16")
    return out


def wanda_score_mask(w2, norm_base, norm_adj, k: int, out=None):
    """Wanda mask bits of one (timestep, layer): w2 bf16 [d, h], norms f32 [h] -> int32 [d*h/32]."""
    lib = _lib.load()
    d, h = w2.shape
    _need(w2, torch.bfloat16, "w2")
    _need(norm_base, torch.float32, "norm_base", (h,))
    _need(norm_adj, torch.float32, "norm_adj", (h,))
    if out is None:
        out = torch.empty(d * h // 32, dtype=torch.int32, device=w2.device)
    _need(out, torch.int32, "out", (d * h // 32,))
    with torch.cuda.device(w2.device):
        rc = lib.moe_wanda_score_mask(_ptr(w2), _ptr(norm_base), _ptr(norm_adj), d, h, int(k), _ptr(out), _stream(w2))
    _lib.check(rc, "moe_wanda_score_mask")
    return out


def mask_vote(masks, threshold: float, out=None):
    """masks int32 [T, n_words] -> int32 [n_words]: bit set where more than `threshold` of the T masks have it."""
    lib = _lib.load()
    _need(masks, torch.int32, "masks")
    Tn, n = masks.shape
    if out is None:
        out = torch.empty(n, dtype=torch.int32, device=masks.device)
    _need(out, torch.int32, "out", (n,))
    with torch.cuda.device(masks.device):
        rc = lib.moe_mask_vote(_ptr(masks), Tn, n, float(threshold), _ptr(out), _stream(masks))
    _lib.check(rc, "moe_mask_vote")
    return out


def mask_pack(dense):
    """dense uint8 0/1 (any shape, n elements) -> int32 [ceil(n/32)] bit words (bit i%32 of word i/32)."""
    lib = _lib.load()
    _need(dense, torch.uint8, "dense")
    n = dense.numel()
    bits = torch.empty(((n + 31) // 32,), dtype=torch.int32, device=dense.device)
    with torch.cuda.device(dense.device):
        rc = lib.moe_mask_pack(_ptr(dense), n, _ptr(bits), _stream(dense))
    _lib.check(rc, "moe_mask_pack")
    return bits


def mask_union(a, b, out=None):
    lib = _lib.load()
    _need(a, torch.int32, "a")
    _need(b, torch.int32, "b", a.shape)
    if out is None:
        out = torch.empty_like(a)
    _need(out, torch.int32, "out", a.shape)
    with torch.cuda.device(a.device):
        rc = lib.moe_mask_union(_ptr(a), _ptr(b), _ptr(out), a.numel(), _stream(a))
    _lib.check(rc, "moe_mask_union")
    return out


def mask_weights(w2, bits, out=None):
    """w2 bf16 [d, h]; bits int32 [d*h/32] -> masked copy (element zeroed where its bit is set)."""
    lib = _lib.load()
    d, h = w2.shape
    _need(w2, torch.bfloat16, "w2")
    _need(bits, torch.int32, "bits", (d * h // 32,))
    if out is None:
        out = torch.empty_like(w2)
    _need(out, torch.bfloat16, "out", (d, h))
    with torch.cuda.device(w2.device):
        rc = lib.moe_mask_weights(_ptr(w2), _ptr(bits), _ptr(out), d, h, _stream(w2))
    _lib.check(rc, "moe_mask_weights")
    return out


def cfg_ddim_step(eps_uncond, eps_cond, x, guidance: float, alpha_t: float, alpha_prev: float, out=None):
    """Classifier-free-guidance combine + DDIM (eta = 0) update in one kernel (moe_cfg_ddim_step).  f32 or bf16 tensors
    of equal shape; alpha_* are the scheduler's cumulative alpha products at this / the previous timestep."""
    lib = _lib.load()
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("cfg_ddim_step: f32 or bf16 latents only")
    for name, t in (("eps_uncond", eps_uncond), ("eps_cond", eps_cond), ("x", x)):
        _need(t, x.dtype, name, x.shape)
    if out is None:
        out = torch.empty_like(x)
    _need(out, x.dtype, "out", x.shape)
    with torch.cuda.device(x.device):
        rc = lib.moe_cfg_ddim_step(_ptr(eps_uncond), _ptr(eps_cond), _ptr(x), _ptr(out), x.numel(), 1 if x.dtype == torch.bfloat16 else 0,
                                   float(guidance), float(alpha_t), float(alpha_prev), _stream(x))
    _lib.check(rc, "moe_cfg_ddim_step")
    return out
