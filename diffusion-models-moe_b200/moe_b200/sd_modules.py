"""Module classes the receivers hook.

If `diffusers` is importable its own GEGLU / GELU / LoRACompatibleLinear are used, so the
receivers drop into a real StableDiffusionPipeline.  It is not installable in this image
(no network), so otherwise the same module surface is defined here, with diffusers-identical
attribute names (`proj`, `gelu`, `net`) so that `named_modules()` yields the names the
reference filters on (`'ff.net' in name`, reference neuron_receivers/base_receiver.py:49-53).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

try:  # pragma: no cover - exercised only where diffusers exists
    from diffusers.models.activations import GEGLU, GELU  # type: ignore
    try:
        from diffusers.models.activations import LoRACompatibleLinear  # type: ignore
    except ImportError:
        from diffusers.models.lora import LoRACompatibleLinear  # type: ignore
    HAVE_DIFFUSERS = True
except ImportError:
    HAVE_DIFFUSERS = False

    class LoRACompatibleLinear(nn.Linear):
        """nn.Linear accepting the LoRA `scale` argument (reference calls `module.proj(x, 1.0)`)."""

        def forward(self, hidden_states, scale: float = 1.0):
            return F.linear(hidden_states, self.weight, self.bias)

    class GELU(nn.Module):
        def __init__(self, dim_in, dim_out, approximate="none"):
            super().__init__()
            self.proj = nn.Linear(dim_in, dim_out)
            self.approximate = approximate

        def gelu(self, gate):
            return F.gelu(gate, approximate=self.approximate)

        def forward(self, hidden_states):
            return self.gelu(self.proj(hidden_states))

    class GEGLU(nn.Module):
        """proj: Linear(d, 2h); value = first half, gate = second half; out = value * gelu(gate)."""

        def __init__(self, dim_in, dim_out):
            super().__init__()
            self.proj = LoRACompatibleLinear(dim_in, dim_out * 2)

        def gelu(self, gate):
            return F.gelu(gate)

        def forward(self, hidden_states, scale: float = 1.0):
            hidden_states, gate = self.proj(hidden_states, scale).chunk(2, dim=-1)
            return hidden_states * self.gelu(gate)


class FeedForward(nn.Module):
    """upstream FeedForward(dim, mult=4, activation_fn='geglu'): net = [GEGLU, Dropout(0), Linear]."""

    def __init__(self, dim: int, mult: int = 4, inner_dim=None):
        super().__init__()
        inner = dim * mult if inner_dim is None else inner_dim
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(0.0), LoRACompatibleLinear(inner, dim)])

    def forward(self, hidden_states, scale: float = 1.0):
        for module in self.net:
            if isinstance(module, (GEGLU, LoRACompatibleLinear)):
                hidden_states = module(hidden_states, scale)
            else:
                hidden_states = module(hidden_states)
        return hidden_states


class FFNTransformerBlock(nn.Module):
    """The slice of BasicTransformerBlock on the hot path: x + ff(norm3(x))."""

    def __init__(self, dim: int):
        super().__init__()
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, hidden_states):
        return hidden_states + self.ff(self.norm3(hidden_states))


class _Attention2D(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.transformer_blocks = nn.ModuleList([FFNTransformerBlock(dim)])

    def forward(self, x):
        return self.transformer_blocks[0](x)


class _Block(nn.Module):
    def __init__(self, dim: int, n_attn: int):
        super().__init__()
        self.attentions = nn.ModuleList([_Attention2D(dim) for _ in range(n_attn)])


# (block attribute, index or None, dim, n_attentions, tokens at 64x64 latents) in UNet firing order
SD15_FFN_PLAN = [
    ("down_blocks", 0, 320, 2, 4096), ("down_blocks", 1, 640, 2, 1024), ("down_blocks", 2, 1280, 2, 256),
    ("mid_block", None, 1280, 1, 64),
    ("up_blocks", 1, 1280, 3, 256), ("up_blocks", 2, 640, 3, 1024), ("up_blocks", 3, 320, 3, 4096),
]


def sd_ffn_shapes(latent_hw: int = 64):
    """[(name, d, h, tokens)] of the 16 transformer-block FFNs in firing order."""
    scale = (latent_hw * latent_hw) / 4096.0
    out = []
    for attr, idx, dim, n_attn, tokens in SD15_FFN_PLAN:
        for j in range(n_attn):
            prefix = f"{attr}.attentions.{j}" if idx is None else f"{attr}.{idx}.attentions.{j}"
            out.append((prefix + ".transformer_blocks.0.ff.net.0", dim, 4 * dim, int(tokens * scale)))
    return out


class FFNStackUNet(nn.Module):
    """The 16 transformer-block FFNs of the SD-1.5 / SD-2.1 UNet under diffusers-identical module
    names, WITHOUT the conv / attention layers around them (those are outside the hot path and
    stay stock PyTorch in a real pipeline).  forward() runs the FFN residual branch of every block
    once, in firing order, on per-layer hidden states -- one "UNet step" of the hot path."""

    def __init__(self, latent_hw: int = 64):
        super().__init__()
        self.latent_hw = latent_hw
        self.down_blocks = nn.ModuleList([_Block(320, 2), _Block(640, 2), _Block(1280, 2), nn.Module()])
        self.mid_block = _Block(1280, 1)
        self.up_blocks = nn.ModuleList([nn.Module(), _Block(1280, 3), _Block(640, 3), _Block(320, 3)])

    def ffn_blocks(self):
        out = []
        for attr, idx, dim, n_attn, tokens in SD15_FFN_PLAN:
            blk = getattr(self, attr) if idx is None else getattr(self, attr)[idx]
            for j in range(n_attn):
                out.append(blk.attentions[j].transformer_blocks[0])
        return out

    def forward(self, states):
        """states: list of 16 tensors [B, S_l, d_l]; returns the list after x + ff(norm3(x))."""
        return [blk(x) for blk, x in zip(self.ffn_blocks(), states)]


class SyntheticFFNPipeline:
    """Pipeline stand-in with the call surface the receivers use: `.unet`, `__call__(prompt, ...)`
    returning an object with `.images`.  One call = `num_inference_steps` UNet steps of the FFN
    stack at batch 2 per prompt (classifier-free guidance), on seeded synthetic hidden states.
    `.images` holds one [16]-list of final hidden states per prompt (there is no VAE here)."""

    class Output:
        def __init__(self, images):
            self.images = images

    def __init__(self, unet: FFNStackUNet, num_inference_steps: int = 50, device="cuda", dtype=torch.bfloat16):
        self.unet = unet.to(device=device, dtype=dtype)
        self.num_inference_steps = num_inference_steps
        self.device = torch.device(device)
        self.dtype = dtype

    def to(self, device):
        self.unet = self.unet.to(device)
        self.device = torch.device(device)
        return self

    @torch.no_grad()
    def __call__(self, prompt, safety_checker=None, num_inference_steps=None, **kwargs):
        prompts = prompt if isinstance(prompt, (list, tuple)) else [prompt]
        steps = num_inference_steps or self.num_inference_steps
        batch = 2 * len(prompts)
        shapes = sd_ffn_shapes(self.unet.latent_hw)
        # seeded synthetic hidden states, drawn on the pipeline's device (no host round trip per call)
        gen = torch.Generator(device=self.device).manual_seed(int(torch.initial_seed()) % (2 ** 31))
        states = [torch.randn(batch, s, d, generator=gen, device=self.device).to(self.dtype) for (_, d, _, s) in shapes]
        for _ in range(steps):
            # the residual stream of every block carries over to the next step (each block normalises its own
            # FFN input with norm3, so the FFN sees O(1) activations at every step)
            states = self.unet(states)
        per_prompt = [[x[2 * i:2 * i + 2] for x in states] for i in range(len(prompts))]
        return SyntheticFFNPipeline.Output(per_prompt)


class GraphedSampling:
    """One whole sampling run (`steps` UNet steps of the FFN stack for a batch of prompts) through a receiver's
    hooks, captured ONCE as a CUDA graph and replayed per prompt batch -- the launch-bound Python hook path
    (16 hooks x `steps` per prompt batch) runs only at capture time.

    The receiver's (timestep, layer) state machine advances during capture, so every captured layer call is bound
    to its own cell: the removed-expert bits of (t, l) and the histogram slice hist[t, l].  Replaying therefore
    accumulates the per-timestep expert counters of every prompt batch into the receiver's device histogram
    (reference flow: moefication/freq_expert_select.py:49-64, one `observe_activation` per prompt).

        gs = GraphedSampling(pipe, receiver, n_prompts=8, steps=50)
        gs.load_states(seed)     # this prompt batch's initial hidden states (any stream-ordered source)
        gs.replay()              # enqueue the whole sampling run; gs.states_out holds the final states
    """

    def __init__(self, pipe, receiver, n_prompts: int, steps: int = None, warmup_steps: int = None):
        self.pipe = pipe
        self.receiver = receiver
        self.steps = steps or pipe.num_inference_steps
        self.batch = 2 * n_prompts                      # classifier-free guidance: 2 UNet rows per prompt
        dev, dt = pipe.device, pipe.dtype
        shapes = sd_ffn_shapes(pipe.unet.latent_hw)
        self.states_in = [torch.zeros(self.batch, s, d, device=dev, dtype=dt) for (_, d, _, s) in shapes]
        self.states_out = None
        self.graph = None
        self._capture(warmup_steps)

    def load_states(self, seed: int):
        """Seeded synthetic hidden states of one prompt batch (identical whichever rank draws them)."""
        gen = torch.Generator(device=self.pipe.device).manual_seed(int(seed))
        for x in self.states_in:
            x.copy_(torch.randn(x.shape, generator=gen, device=x.device, dtype=torch.float32).to(x.dtype))

    @torch.no_grad()
    def _loop(self):
        states = self.states_in
        for _ in range(self.steps):
            states = self.pipe.unet(states)      # same loop as SyntheticFFNPipeline.__call__
        return states

    @torch.no_grad()
    def _capture(self, warmup_steps):
        rec = self.receiver
        hooks = rec.register_hooks(self.pipe)
        try:
            side = torch.cuda.Stream(device=self.pipe.device)
            side.wait_stream(torch.cuda.current_stream(self.pipe.device))
            with torch.cuda.stream(side):
                # eager pass: kernel attributes, workspaces, and every per-(t, l) device artefact of the receiver
                rec.reset_time_layer()
                self._loop()
                torch.cuda.synchronize()
                rec.reset_time_layer()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=side):
                    self.states_out = self._loop()
            torch.cuda.current_stream(self.pipe.device).wait_stream(side)
        finally:
            rec.remove_hooks(hooks)
        rec.reset_time_layer()

    def replay(self):
        self.graph.replay()
