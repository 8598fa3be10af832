"""Minimal `diffusers.models.activations` surface (see ../__init__.py)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class LoRACompatibleLinear(nn.Linear):
    """nn.Linear whose forward accepts the LoRA `scale` argument the reference passes
    (`module.proj(input[0], 1.0)`, reference neuron_receivers/moefy.py:11-12)."""

    def forward(self, hidden_states, scale: float = 1.0):
        return F.linear(hidden_states, self.weight, self.bias)


class GELU(nn.Module):
    """Plain GELU feed-forward activation block: proj then gelu (PixArt-style FFN)."""

    def __init__(self, dim_in, dim_out, approximate="none"):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)
        self.approximate = approximate

    def gelu(self, gate):
        return F.gelu(gate, approximate=self.approximate)

    def forward(self, hidden_states):
        return self.gelu(self.proj(hidden_states))


class GEGLU(nn.Module):
    """Gated GELU: proj to 2h, first half is the value, second half the gate."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = LoRACompatibleLinear(dim_in, dim_out * 2)

    def gelu(self, gate):
        return F.gelu(gate)

    def forward(self, hidden_states, scale: float = 1.0):
        hidden_states, gate = self.proj(hidden_states, scale).chunk(2, dim=-1)
        return hidden_states * self.gelu(gate)
