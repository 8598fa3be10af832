"""Stand-in for the (uninstallable, no network) `diffusers` package.

TEST INFRASTRUCTURE ONLY.  It exists so that the reference's own receiver files
(/root/reference/neuron_receivers/*.py) can be imported *unmodified* in the build
container to generate golden vectors (oracle/gen_golden.py).  It provides only the
module surface those files import; the arithmetic is the upstream definition
[upstream diffusers 0.2x, restated from its public behaviour]:
  GEGLU.proj = LoRACompatibleLinear(d, 2h); forward: value, gate = chunk(2); value * gelu(gate)
"""
