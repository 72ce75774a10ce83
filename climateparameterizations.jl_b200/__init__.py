"""B200-native NDE column engine for OceanParameterizations.jl's hot path (host-side mirror + C-ABI binding).

Import through `cpzload.load()` (alias `cpz_b200`); the CUDA work is done by `lib/libcpz.so` (built from csrc/).
"""
from . import desc, synthetic  # noqa: F401
from .desc import ClosureDesc, ModelDesc, NetDesc  # noqa: F401
