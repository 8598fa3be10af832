"""BaseNeuronReceiver: hook lifecycle of the reference (neuron_receivers/base_receiver.py:10-81)
with the hook arithmetic delegated to libmoe_b200.so.

Same public surface: ctor (seed, replace_fn, keep_nsfw, hook_module), `hook_fn`, `text_hook_fn`,
`remove_hooks`, `observe_activation(model, ann, bboxes=None) -> (out, gates)`, `test`.
Additions (keyword-only, defaults keep the reference behaviour):
  capture_gates       keep the per-call D2H copy of the gate (reference moefy.py:25); set False on
                      the fast path -- it is a forced device sync per layer call.
  skip_stock_forward  while hooked, replace the module's own forward by a stub so the stock GEGLU
                      is not computed just to be discarded (reference quirk SURVEY A.3 item 1).
  fuse_down_proj      while hooked, the sibling `ff.net.2` becomes a pass-through and the GEGLU hook returns the
                      finished FFN output Y [B, S, d]: routed receivers issue ONE `moe_ffn_fused` launch per layer
                      call (up-projection -> routing -> down-projection), the others run K1 (+ their statistic)
                      and the native K3 -- cuBLAS never runs on a hooked FFN.  Off: the hook returns H [B, S, h]
                      exactly as the reference's does and the stock Linear follows.

`gates` (reference moefy.py:25: `self.gates.append(gate.detach().cpu())`) are copied to pinned host memory
asynchronously; reading `receiver.gates` (or the end of `observe_activation`) waits for the copies, so the sampling
loop is not synchronised once per layer call.
"""
import numpy as np
import torch

from moe_b200 import ops
from moe_b200.ffn import attach_down, find_down_proj, get_state
from moe_b200.sd_modules import GEGLU, GELU  # noqa: F401


def sc(self, clip_input, images):
    return images, [False for _ in images]


class _NoSafetyChecker:
    """Placeholder with the attribute the reference monkey-patches (base_receiver.py:20-23)."""

    def forward(self, clip_input, images):
        return images, [False for _ in images]


def _safety_checker_class():
    try:  # pragma: no cover - only where diffusers exists
        from diffusers.pipelines.stable_diffusion import safety_checker
        return safety_checker.StableDiffusionSafetyChecker
    except ImportError:
        return _NoSafetyChecker


def _stub_forward(*args, **kwargs):
    return None


def _passthrough_forward(hidden_states, *args, **kwargs):
    """`ff.net.2.forward` while its GEGLU hook delivers the finished FFN output."""
    return hidden_states


class BaseNeuronReceiver:
    """Base class for storing and changing activation functions."""

    def __init__(self, seed=0, replace_fn=GEGLU, keep_nsfw=False, hook_module='unet', *, capture_gates=True,
                 skip_stock_forward=True, fuse_down_proj=True):
        self.seed = seed
        self._gates = []
        self._gate_copy_event = None
        self.hidden_states = []
        self.keep_nsfw = keep_nsfw
        self.safety_checker = _safety_checker_class()
        if self.keep_nsfw:
            self.safety_checker.forward = sc
        self.replace_fn = replace_fn
        self.hook_module = hook_module
        self.capture_gates = capture_gates
        self.skip_stock_forward = skip_stock_forward
        self.fuse_down_proj = fuse_down_proj
        self._stubbed = []
        self._fused_states = []

    # -- captured gates: pinned host tensors filled by async copies ---------------------------------
    @property
    def gates(self):
        if self._gate_copy_event is not None:
            self._gate_copy_event.synchronize()
            self._gate_copy_event = None
        return self._gates

    @gates.setter
    def gates(self, value):
        self._gates = value

    # -- to be provided by subclasses ---------------------------------------------------------
    def hook_fn(self, module, input, output):
        raise NotImplementedError

    def text_hook_fn(self, module, input, output):
        raise NotImplementedError

    # -- hook lifecycle -------------------------------------------------------------------------
    def _select_modules(self, model):
        if self.hook_module != 'unet':
            raise NotImplementedError("only the UNet FFN path is implemented natively (hook_module='unet')")
        return [(name, m) for name, m in model.unet.named_modules()
                if isinstance(m, self.replace_fn) and 'ff.net' in name]

    def _hook_function(self):
        return self.hook_fn

    def register_hooks(self, model, bboxes=None):
        hooks = []
        for name, module in self._select_modules(model):
            hooks.append(module.register_forward_hook(self._hook_function()))
            if hasattr(module, 'proj'):
                module.bounding_box = bboxes[name + '.proj.weight'] if bboxes is not None else None
            if self.skip_stock_forward and 'forward' not in module.__dict__:
                module.forward = _stub_forward
                self._stubbed.append(module)
            if self.fuse_down_proj and isinstance(module, GEGLU):
                self._fuse_down(model.unet, name, module)
        return hooks

    def _fuse_down(self, root, name, module):
        """Route the down-projection of a hooked GEGLU through the native kernels: `ff.net.2` passes its input
        through and the hook returns Y.  Needs the sibling Linear; FFNs without one keep the reference flow."""
        down = find_down_proj(root, name)
        if down is None or 'forward' in down.__dict__ or len(down._forward_hooks) > 0:
            return
        state = get_state(module)
        if state.down_module is not down:
            attach_down(module, state, down)
        if state.w2p is None or not state.w2p.is_cuda:
            return
        down.forward = _passthrough_forward
        self._stubbed.append(down)
        state.fused_down = True
        self._fused_states.append(state)

    def remove_hooks(self, hooks):
        for hook in hooks:
            hook.remove()
        for module in self._stubbed:
            module.__dict__.pop('forward', None)
        self._stubbed = []
        for state in self._fused_states:
            state.fused_down = False
        self._fused_states = []

    def _run_model(self, model, ann):
        try:
            return model(ann, safety_checker=self.safety_checker).images[0]
        except TypeError:
            return model(ann).images[0]

    def observe_activation(self, model, ann, bboxes=None):
        self.gates = []
        hooks = self.register_hooks(model, bboxes)
        try:
            # fix the seed to get the same output for every run (base_receiver.py:70-72)
            torch.manual_seed(self.seed)
            np.random.seed(self.seed)
            out = self._run_model(model, ann)
        finally:
            self.remove_hooks(hooks)
        return out, self.gates      # (reading `gates` waits for the last asynchronous gate copy)

    def test(self, model, ann='A brown dog in the snow'):
        raise NotImplementedError

    # -- shared by the GEGLU receivers ------------------------------------------------------------
    def _capture(self, gate_packed, state, lead_shape):
        """Reference: `self.gates.append(gate.detach().cpu())` with gate [B, S, h] in the ORIGINAL neuron order
        (the order of the label files and of every per-neuron artefact the reference writes)."""
        if gate_packed is None:
            return
        g = original_order(gate_packed, state)
        g = g.view(*lead_shape, g.shape[-1]).detach()
        if g.is_cuda:
            host = torch.empty(g.shape, dtype=g.dtype, pin_memory=True)
            host.copy_(g, non_blocking=True)
            self._gate_copy_event = torch.cuda.Event()
            self._gate_copy_event.record(torch.cuda.current_stream(g.device))
            self._gates.append(host)
        else:
            self._gates.append(g.cpu())

    @staticmethod
    def _finish(H, state, lead_shape, like):
        """Packed-order H [T, h] -> hook output: with the down-projection fused behind the hook, Y [B, S, d] from
        the native K3; otherwise H [B, S, h] in the order the following stock ff.net.2 expects."""
        if state.fused_down:
            out = ops.down_proj(H, state.w2p, state.b2)
        elif not state.weights_permuted_in_model:
            out = H[:, state.layout.inv_perm.to(H.device)]
        else:
            out = H
        out = out.view(*lead_shape, out.shape[-1])
        return out if out.dtype == like.dtype else out.to(like.dtype)


def original_order(t, state):
    """Per-neuron tensor [..., h] in the kernels' packed order -> the ORIGINAL neuron order of the model before
    MoEfication, whatever `weights_permuted_in_model` says: measured artefacts (gates, max activations, column
    norms) are interchangeable with the reference's and with what RemoveNeurons / WandaRemoveNeuronsFast expect."""
    lay = state.layout
    if lay.is_identity:
        return t
    return t[..., lay.inv_perm.to(t.device)]
