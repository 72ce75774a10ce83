"""Seeded synthetic inputs for the NDE column path. The reference ships no LES data, so every
parity test and benchmark runs on these. Generated on the host with numpy only; the CUDA engine, the oracle
and the CPU baseline all see identical float32 bits.

Magnitudes follow the LESbrary suite the reference trains on (wind_mixing/src/data_containers.jl:44-61:
Qu 2e-4..1e-3, Qb 1e-8..5e-8) and the constants of wind_mixing/train_NDE.jl:69-73 / NDE_training.jl:168.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .desc import (FLAG_CA, FLAG_MPP, FLAG_ZERO_WEIGHTS, RHS_FREE_CONVECTION, RHS_INFER, RHS_TRAIN, ModelDesc,
                   NetDesc)

# (mu, sigma) for u, v, T, uw, vw, wT  — SURVEY §8d "Scalings"
MU = (-0.01, 0.005, 19.8, -1e-4, 0.0, 5e-6)
SIGMA = (0.05, 0.03, 0.15, 1.5e-4, 5e-5, 6e-6)

# NN shapes that occur in the reference
NET_SHAPES = {
    # wind_mixing/train_NDE.jl:103
    "uvT_small": lambda Nz=32: NetDesc([3 * Nz, 50, 20, Nz - 1], ["mish", "mish", "identity"]),
    # wind_mixing/test_train_NDE.jl:33
    "uvT_test": lambda Nz=32, act="relu": NetDesc([3 * Nz, 400, Nz - 1], [act, "identity"]),
    # wind_mixing/construct_NN.jl:32-34
    "uvT_large": lambda Nz=32: NetDesc([3 * Nz, 400, 400, Nz - 1], ["relu", "relu", "identity"]),
    # free_convection/train_free_convection_nde.jl:119-121
    "T_only": lambda Nz=32: NetDesc([Nz, 4 * Nz, 4 * Nz, Nz - 1], ["relu", "relu", "identity"]),
}


def n_substeps_for(desc: ModelDesc, dz: Optional[float] = None) -> int:
    """Sub-steps so that nu_max*dt/(n_sub*dz^2) stays inside the explicit integrator's real-axis stability
    interval (0.8 Tsit5 / 0.65 RK4 / 0.45 Euler)."""
    lim = {"tsit5": 0.8, "rk4": 0.65, "euler": 0.45}[desc.integrator]
    dz = dz if dz is not None else desc.H / desc.Nz
    if desc.variant == RHS_FREE_CONVECTION:
        r = 0.0
        if desc.has(FLAG_CA):
            # non-dimensional K on a unit-depth grid: K*dt_hat*Nz^2 * (sigma_wT/sigma_T*tau/H)
            A = desc.sigma[5] / desc.sigma[2] * desc.tau / desc.H
            r += A * desc.K_ca * desc.dt * desc.Nz ** 2
        if desc.has(FLAG_MPP):  # mPP base in the T-only model (BASELINE config 1): nu_max/Pr on unstable faces, added to K
            r += (desc.nu0 + desc.nu_m) / desc.Pr * desc.dt * desc.tau / dz ** 2
            if desc.has(FLAG_CA):
                # two switched diffusivities (K at dT/dz < 0, nu_- at dT/dz < -eps) chatter at neutral faces; measured on
                # the FP64 oracle: the tangent of the fixed-step map blows up at 0.71 of the linear limit and is clean at 0.61
                lim *= 0.75
        return max(1, int(np.ceil(r / lim)))
    nu_max = desc.nu0 + desc.nu_m
    if desc.has(FLAG_CA):
        nu_max = max(nu_max, desc.kappa)
    r = nu_max * desc.dt * desc.tau / dz ** 2
    return max(1, int(np.ceil(r / lim)))


def wind_mixing_desc(variant: int = RHS_INFER, net: str = "uvT_small", Nz: int = 32, n_steps: int = 1152,
                     save_stride: int = 1, ckpt_stride: int = 9, integrator: str = "tsit5", flags: Optional[int] = None,
                     n_substeps: Optional[int] = None, **kw) -> ModelDesc:
    """u/v/T model with the constants of SURVEY §8d (H=256 m, tau=8 days, dt_hat=1/1152, mPP nu0=1e-4, nu_=0.1,
    Ric=0.25, dRi=0.1, Pr=1, kappa=0.1, eps=1e-7)."""
    if flags is None:
        flags = FLAG_MPP | FLAG_ZERO_WEIGHTS
    nets = [] if net is None else [NET_SHAPES[net](Nz) for _ in range(3)]
    d = ModelDesc(Nz=Nz, n_fields=3, variant=variant, flags=flags, nets=nets, H=256.0 * Nz / 32, tau=691200.0,
                  f=1e-4, g=9.80665, alpha=2e-4, nu0=1e-4, nu_m=0.1, Ric=0.25, dRi=0.1, Pr=1.0, kappa=0.1, eps=1e-7,
                  mu=MU, sigma=SIGMA, integrator=integrator, dt=1.0 / 1152.0, t0=0.0, n_steps=n_steps,
                  save_stride=save_stride, ckpt_stride=ckpt_stride)
    for k, v in kw.items():
        setattr(d, k, v)
    d.n_substeps = n_substeps if n_substeps is not None else n_substeps_for(d)
    return d


def free_convection_desc(ca: bool = True, Nz: int = 32, n_steps: int = 1152, save_stride: int = 9, ckpt_stride: int = 9,
                         integrator: str = "tsit5", net: Optional[str] = "T_only", n_substeps: Optional[int] = None,
                         mpp: bool = False, **kw) -> ModelDesc:
    """T-only model (free_convection): H = 100 m (free_convection/convective_adjustment.jl:70-72), tau = 8 days,
    Nt = 1153 frames so dt_hat = 1/1153 in the reference's tspan convention (free_convection_nde.jl:40);
    we keep 1/1152 so one step is one 600-s frame interval. mpp=True adds the mPP base diffusivity at u = v = 0
    (BASELINE config 1: "convective adjustment + mPP base"; constants of wind_mixing/train_NDE.jl:69-73)."""
    nets = [] if net is None else [NET_SHAPES[net](Nz)]
    d = ModelDesc(Nz=Nz, n_fields=1, variant=RHS_FREE_CONVECTION, flags=(FLAG_CA if ca else 0) | (FLAG_MPP if mpp else 0), nets=nets, H=100.0,
                  tau=691200.0, mu=MU, sigma=SIGMA, K_ca=10.0, integrator=integrator, dt=1.0 / 1152.0, n_steps=n_steps,
                  save_stride=save_stride, ckpt_stride=ckpt_stride)
    for k, v in kw.items():
        setattr(d, k, v)
    d.n_substeps = n_substeps if n_substeps is not None else n_substeps_for(d)
    return d


def theta_init(desc: ModelDesc, seed: int = 42, scale: float = 1e-5) -> np.ndarray:
    """Flux-default Glorot-uniform W, b = 0, in destructure order, times `scale` (the reference divides initial
    weights by 1e5 before NDE training: wind_mixing/train_NDE.jl:105-107). Per SURVEY §8d the same base weights
    are used for all three nets, as `re(weights ./ 1f5)` is in train_NDE.jl:105-107, unless distinct=True."""
    rng = np.random.default_rng(seed)
    parts = []
    for n in desc.nets:
        for i in range(len(n.acts)):
            fan_in, fan_out = n.sizes[i], n.sizes[i + 1]
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            W = rng.uniform(-lim, lim, size=(fan_out, fan_in)).astype(np.float32)
            parts.append(W.flatten(order="F"))
            parts.append(np.zeros(fan_out, dtype=np.float32))
    if not parts:
        return np.zeros(0, dtype=np.float32)
    return (np.concatenate(parts) * np.float32(scale)).astype(np.float32)


def theta_random(desc: ModelDesc, seed: int = 7, scale: float = 1.0, bias_scale: float = 0.1) -> np.ndarray:
    """Like theta_init but with non-zero biases, for parity tests where every parameter must matter."""
    rng = np.random.default_rng(seed)
    parts = []
    for n in desc.nets:
        for i in range(len(n.acts)):
            fan_in, fan_out = n.sizes[i], n.sizes[i + 1]
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            parts.append((rng.uniform(-lim, lim, size=fan_in * fan_out) * scale).astype(np.float32))
            parts.append((rng.uniform(-1, 1, size=fan_out) * bias_scale * scale).astype(np.float32))
    if not parts:
        return np.zeros(0, dtype=np.float32)
    return np.concatenate(parts).astype(np.float32)


FORCING_CASES = [(Qu, Qb) for Qu in (2e-4, 3.5e-4, 5e-4) for Qb in (1e-8, 2e-8, 3e-8)]  # wind_mixing/train_NDE_args.jl:39-60


def columns(desc: ModelDesc, ncol: int, seed: int = 1000, case_stride: Optional[int] = None
            ) -> Tuple[np.ndarray, np.ndarray]:
    """Initial profiles x0 [ncol, S] (scaled) and boundary fluxes bcs [ncol, n_bc] (scaled), SURVEY §8d.

    physical T(z) = 20 + 0.01 z (1+0.2 xi1) with a mixed layer of depth 20-60 m, u = 0.05 xi2 exp(z/30),
    v = 0.03 xi3 exp(z/30), plus 1e-3 sigma white noise per level (keeps shear non-zero, quirk Q4).
    Column i uses forcing case (i // case_stride) % 9 (default: blocks of ncol/9)."""
    rng = np.random.default_rng(seed)
    Nz, H = desc.Nz, desc.H
    dz = H / Nz
    z = -H + dz * (np.arange(Nz) + 0.5)  # centres, bottom first; z=0 at the surface
    xi = rng.standard_normal((ncol, 3))
    mld = rng.uniform(20.0, 60.0, size=(ncol, 1)) * (H / 256.0 if desc.n_fields == 3 else H / 100.0)
    zz = z[None, :]
    T = 20.0 + 0.01 * np.minimum(zz, -mld) * (1 + 0.2 * xi[:, 0:1])
    u = 0.05 * xi[:, 1:2] * np.exp(zz / 30.0)
    v = 0.03 * xi[:, 2:3] * np.exp(zz / 30.0)
    noise = rng.standard_normal((ncol, 3, Nz)) * 1e-3
    mu, sg = desc.mu, desc.sigma
    us = (u - mu[0]) / sg[0] + noise[:, 0]
    vs = (v - mu[1]) / sg[1] + noise[:, 1]
    Ts = (T - mu[2]) / sg[2] + noise[:, 2]
    if case_stride is None:
        case_stride = max(1, ncol // len(FORCING_CASES))
    case = (np.arange(ncol) // case_stride) % len(FORCING_CASES)
    Qu = np.array([FORCING_CASES[c][0] for c in case])
    Qb = np.array([FORCING_CASES[c][1] for c in case])
    s = lambda val, i: (val - mu[i]) / sg[i]
    if desc.n_fields == 3:
        x0 = np.concatenate([us, vs, Ts], axis=1)
        bcs = np.stack([np.full(ncol, s(0.0, 3)), s(-Qu, 3), np.full(ncol, s(0.0, 4)), np.full(ncol, s(0.0, 4)),
                        np.full(ncol, s(0.0, 5)), s(Qb / (desc.alpha * desc.g), 5)], axis=1)
    else:
        x0 = Ts
        bcs = np.stack([np.full(ncol, s(0.0, 5)), s(Qb / (desc.alpha * desc.g), 5)], axis=1)
    return np.ascontiguousarray(x0, dtype=np.float32), np.ascontiguousarray(bcs, dtype=np.float32)


def diurnal_Q(ncol: int, seed: int = 3000) -> np.ndarray:
    """Per-column diurnal buoyancy-flux amplitudes from the set in wind_mixing/src/data_containers.jl:138-152."""
    rng = np.random.default_rng(seed)
    return rng.choice(np.array([5.5e-8, 5e-8, 4e-8, 3.5e-8, 3e-8, 2e-8, 1e-8], dtype=np.float32), size=ncol)


def gyre_field(Nx: int, Ny: int, Nz: int = 32, seed: int = 4000, Ly: float = 6.0e6) -> Tuple[np.ndarray, np.ndarray]:
    """T [Nz, Ny, Nx] in deg C (0..30, stratified, weakly perturbed) and y [Ny] for the double-gyre closure
    (free_convection/double_gyre_nn.jl:62-71,116)."""
    rng = np.random.default_rng(seed)
    y = (-Ly / 2 + (np.arange(Ny) + 0.5) * Ly / Ny).astype(np.float32)
    k = (np.arange(Nz) + 0.5) / Nz
    base = 2.0 + 24.0 * k[:, None, None] ** 2 * (0.6 + 0.4 * (y[None, :, None] / Ly + 0.5))
    T = base + 0.3 * rng.standard_normal((Nz, Ny, Nx))
    return np.ascontiguousarray(T, dtype=np.float32), y


def uvt_fields(desc: ModelDesc, Nx: int, Ny: int, seed: int = 5000, unstable_every: int = 0
               ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Dimensional u, v [m/s] and T [deg C] fields [Nz, Ny, Nx] for the embedded u/v/T closure (NDE_oceananigans.jl): the
    profiles of `columns` un-scaled; every `unstable_every`-th column gets a statically unstable top third so that the
    convective-adjustment branch of the diffusivity acts."""
    x0, _ = columns(desc, Nx * Ny, seed=seed)
    N = desc.Nz
    f = [x0[:, q * N:(q + 1) * N].astype(np.float64) * desc.sigma[q] + desc.mu[q] for q in range(3)]
    if unstable_every > 0:
        k0 = 2 * N // 3
        f[2][::unstable_every, k0:] = f[2][::unstable_every, k0:][:, ::-1] - 0.02 * np.arange(N - k0)[None, :]
    to_field = lambda a: np.ascontiguousarray(a.T.reshape(N, Ny, Nx), dtype=np.float32)
    return to_field(f[0]), to_field(f[1]), to_field(f[2])
