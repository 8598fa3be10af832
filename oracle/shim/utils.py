"""Stand-in for the reference's top-level `utils` module.

The head of /root/reference/utils.py imports diffusers pipelines / `sld`, which do not
exist here; its tail (from `class Average`, utils.py:233 onward: Average, StandardDev,
StatMeter, column-norm accumulators) is self-contained.  We exec that tail *from the
reference tree at run time* (nothing is copied into this repo)."""
import os
import json  # noqa: F401  (used by the exec'd StatMeter.save)
import numpy as np  # noqa: F401
import torch  # noqa: F401

_REF = os.environ.get("MOE_REFERENCE_ROOT", "/root/reference")
with open(os.path.join(_REF, "utils.py")) as _f:
    _src = _f.read()
_start = _src.index("class Average")
exec(compile(_src[_start:], os.path.join(_REF, "utils.py"), "exec"), globals())
