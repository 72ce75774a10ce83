"""ctypes mirror of include/cpz.h (the POD structs and enums that cross the C ABI).

Field order and types must match the header exactly; tests/test_abi.py checks sizeof()
against the value the library reports.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Sequence

MAX_LAYERS = 6
MAX_NETS = 3

# cpz_rhs_variant
RHS_TRAIN, RHS_INFER, RHS_FREE_CONVECTION = 0, 1, 2
# flags
FLAG_MPP = 1 << 0
FLAG_CA = 1 << 1
FLAG_ZERO_WEIGHTS = 1 << 2
FLAG_SMOOTH_NN = 1 << 3
FLAG_SMOOTH_RI = 1 << 4
FLAG_DIURNAL = 1 << 5
FLAG_CA_LITERAL_U = 1 << 6
FLAG_DIURNAL_UNSHIFTED = 1 << 7
FLAG_IMPLICIT_DIFFUSION = 1 << 8
# activations
ACT = {"identity": 0, "relu": 1, "mish": 2, "swish": 3, "leakyrelu": 4, "tanh": 5}
ACT_NAMES = {v: k for k, v in ACT.items()}
# integrators
INTEGRATOR = {"euler": 0, "rk4": 1, "tsit5": 2}
INTEGRATOR_NAMES = {v: k for k, v in INTEGRATOR.items()}
# evaluations of the RHS per integrator step (Tsit5 is FSAL: 6 new evaluations per step)
RHS_EVALS = {"euler": 1, "rk4": 4, "tsit5": 6}


class CNetDesc(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32),
        ("sizes", C.c_int32 * (MAX_LAYERS + 1)),
        ("act", C.c_int32 * MAX_LAYERS),
    ]


class CModelDesc(C.Structure):
    _fields_ = [
        ("Nz", C.c_int32),
        ("n_fields", C.c_int32),
        ("variant", C.c_int32),
        ("flags", C.c_uint32),
        ("n_nets", C.c_int32),
        ("nets", CNetDesc * MAX_NETS),
        ("H", C.c_float), ("tau", C.c_float), ("f", C.c_float), ("g", C.c_float), ("alpha", C.c_float),
        ("nu0", C.c_float), ("nu_m", C.c_float), ("Ric", C.c_float), ("dRi", C.c_float), ("Pr", C.c_float),
        ("kappa", C.c_float), ("eps", C.c_float),
        ("mu", C.c_float * 6),
        ("sigma", C.c_float * 6),
        ("K_ca", C.c_float),
        ("diurnal_period", C.c_float),
        ("integrator", C.c_int32),
        ("dt", C.c_float),
        ("t0", C.c_float),
        ("n_steps", C.c_int32),
        ("n_substeps", C.c_int32),
        ("save_stride", C.c_int32),
        ("ckpt_stride", C.c_int32),
    ]


class CClosureDesc(C.Structure):
    _fields_ = [
        ("Nx", C.c_int32), ("Ny", C.c_int32), ("Nz", C.c_int32),
        ("dz", C.c_float), ("dt", C.c_float), ("K", C.c_float),
        ("T_shift", C.c_float), ("T_div", C.c_float), ("mu_relax", C.c_float),
        ("T_mid", C.c_float), ("dT", C.c_float), ("Ly", C.c_float),
    ]


class CClosureUvtDesc(C.Structure):
    _fields_ = [
        ("Nx", C.c_int32), ("Ny", C.c_int32), ("Nz", C.c_int32),
        ("dz", C.c_float), ("dt", C.c_float),
        ("uw_top", C.c_float), ("vw_top", C.c_float), ("wT_top", C.c_float),
        ("convective_adjustment", C.c_int32), ("kappa_ca", C.c_float),
    ]


@dataclass
class NetDesc:
    """One Flux.Chain of Dense layers: sizes = [in, hidden..., out], acts per layer."""
    sizes: List[int]
    acts: List[str]

    def __post_init__(self):
        assert len(self.sizes) - 1 == len(self.acts) and 1 <= len(self.acts) <= MAX_LAYERS
        for a in self.acts:
            assert a in ACT, a

    @property
    def n_params(self) -> int:
        return sum(self.sizes[i] * self.sizes[i + 1] + self.sizes[i + 1] for i in range(len(self.acts)))

    @property
    def macs(self) -> int:
        return sum(self.sizes[i] * self.sizes[i + 1] for i in range(len(self.acts)))


@dataclass
class ModelDesc:
    """Python-side model description; `.to_c()` gives the struct that crosses the ABI.

    mu/sigma order: u, v, T, uw, vw, wT. Constants are stored as Python floats but are
    rounded to float32 when used (the engine and the oracle see identical float32 values).
    """
    Nz: int = 32
    n_fields: int = 3
    variant: int = RHS_TRAIN
    flags: int = 0
    nets: List[NetDesc] = field(default_factory=list)
    H: float = 256.0
    tau: float = 691200.0
    f: float = 1e-4
    g: float = 9.80665
    alpha: float = 2e-4
    nu0: float = 1e-4
    nu_m: float = 0.1
    Ric: float = 0.25
    dRi: float = 0.1
    Pr: float = 1.0
    kappa: float = 0.1
    eps: float = 1e-7
    mu: Sequence[float] = (0.0,) * 6
    sigma: Sequence[float] = (1.0,) * 6
    K_ca: float = 10.0
    diurnal_period: float = 86400.0
    integrator: str = "tsit5"
    dt: float = 1.0 / 1152.0
    t0: float = 0.0
    n_steps: int = 1
    n_substeps: int = 1
    save_stride: int = 1
    ckpt_stride: int = 1

    # ---- derived ----
    @property
    def S(self) -> int:
        return self.n_fields * self.Nz

    @property
    def n_bc(self) -> int:
        return 6 if self.n_fields == 3 else 2

    @property
    def n_params(self) -> int:
        return sum(n.n_params for n in self.nets)

    @property
    def n_saved(self) -> int:
        if self.save_stride <= 0:
            return 1
        return self.n_steps // self.save_stride + 1

    @property
    def rhs_evals_per_step(self) -> int:
        return RHS_EVALS[self.integrator] * self.n_substeps

    def has(self, flag: int) -> bool:
        return bool(self.flags & flag)

    def validate(self) -> None:
        assert 4 <= self.Nz <= 64
        assert self.n_fields in (1, 3)
        if self.n_fields == 1:
            assert self.variant == RHS_FREE_CONVECTION and len(self.nets) in (0, 1)
        else:
            assert self.variant in (RHS_TRAIN, RHS_INFER) and len(self.nets) in (0, 3)
        for n in self.nets:
            assert n.sizes[0] == self.S and n.sizes[-1] == self.Nz - 1, (n.sizes, self.S, self.Nz)
        assert self.integrator in INTEGRATOR
        assert self.n_steps >= 1 and self.n_substeps >= 1 and self.ckpt_stride >= 1 and self.save_stride >= 0

    def to_c(self) -> CModelDesc:
        self.validate()
        d = CModelDesc()
        d.Nz, d.n_fields, d.variant, d.flags, d.n_nets = self.Nz, self.n_fields, self.variant, self.flags, len(self.nets)
        for i, n in enumerate(self.nets):
            d.nets[i].n_layers = len(n.acts)
            for j, s in enumerate(n.sizes):
                d.nets[i].sizes[j] = s
            for j, a in enumerate(n.acts):
                d.nets[i].act[j] = ACT[a]
        for k in ("H", "tau", "f", "g", "alpha", "nu0", "nu_m", "Ric", "dRi", "Pr", "kappa", "eps", "K_ca",
                  "diurnal_period", "dt", "t0"):
            setattr(d, k, float(getattr(self, k)))
        for i in range(6):
            d.mu[i] = float(self.mu[i])
            d.sigma[i] = float(self.sigma[i])
        d.integrator = INTEGRATOR[self.integrator]
        d.n_steps, d.n_substeps, d.save_stride, d.ckpt_stride = self.n_steps, self.n_substeps, self.save_stride, self.ckpt_stride
        return d


@dataclass
class ClosureDesc:
    """Per-step closure in a 3-D host model (free_convection/double_gyre_nn.jl:27-62,110-168)."""
    Nx: int
    Ny: int
    Nz: int = 32
    dz: float = 62.5
    dt: float = 3600.0
    K: float = 10.0
    T_shift: float = 19.65
    T_div: float = 20.0
    mu_relax: float = 1.0 / 86400.0
    T_mid: float = 15.0
    dT: float = 30.0
    Ly: float = 6.0e6

    def to_c(self) -> CClosureDesc:
        c = CClosureDesc()
        for k in ("Nx", "Ny", "Nz"):
            setattr(c, k, int(getattr(self, k)))
        for k in ("dz", "dt", "K", "T_shift", "T_div", "mu_relax", "T_mid", "dT", "Ly"):
            setattr(c, k, float(getattr(self, k)))
        return c


@dataclass
class ClosureUvtDesc:
    """Per-step closure of the u/v/T NDE inside a host ocean model (wind_mixing/src/NDE_oceananigans.jl:17-101,288-344,380-405).
    Defaults: the reference's single-column run (Lz = 256 m over 32 levels, dt = 60 s, kappa_ca = 1)."""
    Nx: int
    Ny: int
    Nz: int = 32
    dz: float = 8.0
    dt: float = 60.0
    uw_top: float = 0.0
    vw_top: float = 0.0
    wT_top: float = 0.0
    convective_adjustment: bool = False
    kappa_ca: float = 1.0

    def to_c(self) -> CClosureUvtDesc:
        c = CClosureUvtDesc()
        for k in ("Nx", "Ny", "Nz"):
            setattr(c, k, int(getattr(self, k)))
        for k in ("dz", "dt", "uw_top", "vw_top", "wT_top", "kappa_ca"):
            setattr(c, k, float(getattr(self, k)))
        c.convective_adjustment = 1 if self.convective_adjustment else 0
        return c
