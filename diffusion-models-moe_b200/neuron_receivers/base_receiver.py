"""BaseNeuronReceiver: hook lifecycle of the reference (neuron_receivers/base_receiver.py:10-81)
with the hook arithmetic delegated to libmoe_b200.so.

Same public surface: ctor (seed, replace_fn, keep_nsfw, hook_module), `hook_fn`, `text_hook_fn`,
`remove_hooks`, `observe_activation(model, ann, bboxes=None) -> (out, gates)`, `test`.
Additions (keyword-only, defaults keep the reference behaviour):
  capture_gates       keep the per-call D2H copy of the gate (reference moefy.py:25); set False on
                      the fast path -- it is a forced device sync per layer call.
  skip_stock_forward  while hooked, replace the module's own forward by a stub so the stock GEGLU
                      is not computed just to be discarded (reference quirk SURVEY A.3 item 1).
"""
import numpy as np
import torch

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


class BaseNeuronReceiver:
    """Base class for storing and changing activation functions."""

    def __init__(self, seed=0, replace_fn=GEGLU, keep_nsfw=False, hook_module='unet', *, capture_gates=True,
                 skip_stock_forward=True):
        self.seed = seed
        self.gates = []
        self.hidden_states = []
        self.keep_nsfw = keep_nsfw
        self.safety_checker = _safety_checker_class()
        if self.keep_nsfw:
            self.safety_checker.forward = sc
        self.replace_fn = replace_fn
        self.hook_module = hook_module
        self.capture_gates = capture_gates
        self.skip_stock_forward = skip_stock_forward
        self._stubbed = []

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
        return hooks

    def remove_hooks(self, hooks):
        for hook in hooks:
            hook.remove()
        for module in self._stubbed:
            module.__dict__.pop('forward', None)
        self._stubbed = []

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
        return out, self.gates

    def test(self, model, ann='A brown dog in the snow'):
        raise NotImplementedError

    # -- shared by the GEGLU receivers ------------------------------------------------------------
    def _capture(self, gate_packed, state, lead_shape):
        """Reference: `self.gates.append(gate.detach().cpu())` with gate [B, S, h] in the model's
        neuron order."""
        if gate_packed is None:
            return
        g = gate_packed
        if not state.weights_permuted_in_model:
            g = g[:, state.layout.inv_perm.to(g.device)]
        self.gates.append(g.view(*lead_shape, g.shape[-1]).detach().cpu())

    @staticmethod
    def _finish(H, state, lead_shape, like):
        """Packed-order H [T, h] -> hook output [B, S, h] in the order the following ff.net.2 expects."""
        if not state.weights_permuted_in_model:
            H = H[:, state.layout.inv_perm.to(H.device)]
        out = H.view(*lead_shape, H.shape[-1])
        return out if out.dtype == like.dtype else out.to(like.dtype)
