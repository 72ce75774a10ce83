"""B200-native NDE column engine for OceanParameterizations.jl's hot path (host-side mirror + C-ABI binding).

Import through `cpzload.load()` (alias `cpz_b200`); the CUDA work is done by `lib/libcpz.so` (built from csrc/).
    engine                 ctypes binding of include/cpz.h (Context, Model)
    ocean_parameterizations, wind_mixing, free_convection, flux   mirrors of the reference modules' NDE interface
    synthetic              seeded synthetic inputs
    parallel               column sharding + the torch.distributed allreduce hook
"""
from . import desc, flux, ocean_parameterizations, parallel, synthetic  # noqa: F401
from .desc import ClosureDesc, ModelDesc, NetDesc  # noqa: F401
