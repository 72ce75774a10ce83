"""Host-side mirror of the FreeConvection NDE interface (free_convection/src/FreeConvection.jl:3-23), backed by libcpz.so.

FreeConvectionNDE / ConvectiveAdjustmentNDE build the T-only problem (free_convection_nde.jl:1-47,
convective_adjustment_nde.jl:1-57); FreeConvectionNDEParameters packs [bottom_flux, top_flux, σ_T, σ_wT, H, τ]
(:49-62); solve_nde integrates (solve.jl:1-51); train_neural_differential_equation is training.jl:44-74.
Batched over simulations; fixed-step explicit time stepping (see wind_mixing.py header).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import engine
from .desc import FLAG_CA, RHS_FREE_CONVECTION, ModelDesc
from .flux import ADAM, Chain, destructure
from .ocean_parameterizations import ZeroMeanUnitVarianceScaling
from .wind_mixing import _integrator

log = logging.getLogger("FreeConvection")


@dataclass
class FreeConvectionDataset:
    """One simulation's inputs as the NDE constructors read them from `ds` (free_convection_nde.jl:4-14,49-62):
    T [Nt, Nz] (deg C), wT bottom value and the imposed surface temperature flux, zf[0] (= -H), times [Nt] seconds."""
    T: np.ndarray
    wT_bottom: float
    temperature_flux: float
    H: float
    times: np.ndarray


@dataclass
class NDEProblem:
    """The ODEProblem returned by FreeConvectionNDE / ConvectiveAdjustmentNDE, reduced to what the engine needs."""
    nn: Chain
    convective_adjustment: bool
    Nz: int
    Nt: int
    H: float
    tau: float
    iterations: np.ndarray  # 0-based frame indices to save


def _problem(NN: Chain, ds: FreeConvectionDataset, iterations, ca: bool) -> NDEProblem:
    Nt, Nz = ds.T.shape
    its = np.arange(Nt) if iterations is None else np.asarray(iterations, dtype=np.int64)
    return NDEProblem(NN, ca, Nz, Nt, float(ds.H), float(ds.times[-1]), its)


def FreeConvectionNDE(NN: Chain, ds: FreeConvectionDataset, iterations=None) -> NDEProblem:
    return _problem(NN, ds, iterations, False)


def ConvectiveAdjustmentNDE(NN: Chain, ds: FreeConvectionDataset, iterations=None) -> NDEProblem:
    return _problem(NN, ds, iterations, True)


def FreeConvectionNDEParameters(ds: FreeConvectionDataset, T_scaling, wT_scaling) -> np.ndarray:
    """[bottom_flux, top_flux, σ_T, σ_wT, H, τ] (free_convection_nde.jl:49-62)"""
    return np.array([wT_scaling(ds.wT_bottom), wT_scaling(ds.temperature_flux), T_scaling.sigma, wT_scaling.sigma, ds.H,
                     ds.times[-1]], dtype=np.float32)


def _desc(nde: NDEProblem, T_scaling, wT_scaling, integrator: str, n_substeps: Optional[int], K: float = 10.0) -> ModelDesc:
    its = nde.iterations
    stride = int(its[1] - its[0]) if len(its) > 1 else 1
    if len(its) > 1 and np.any(np.diff(its) != stride):
        raise ValueError("the fixed-step engine needs uniformly strided iterations")
    n_steps = int(its[-1] - its[0]) if len(its) > 1 else 1
    mu = [0, 0, T_scaling.mu, 0, 0, wT_scaling.mu]
    sg = [1, 1, T_scaling.sigma, 1, 1, wT_scaling.sigma]
    d = ModelDesc(Nz=nde.Nz, n_fields=1, variant=RHS_FREE_CONVECTION, flags=(FLAG_CA if nde.convective_adjustment else 0),
                  nets=[nde.nn.net_desc()], H=nde.H, tau=nde.tau, mu=mu, sigma=sg, K_ca=K, integrator=integrator,
                  dt=1.0 / nde.Nt, t0=float(its[0]) / nde.Nt,  # tspan = (0, max(iterations)/Nt), free_convection_nde.jl:40
                  n_steps=n_steps, save_stride=stride, ckpt_stride=stride)
    if n_substeps is None:
        from .synthetic import n_substeps_for
        n_substeps = n_substeps_for(d)
    d.n_substeps = n_substeps
    return d


def solve_nde(ndes: Sequence[NDEProblem], NN: Chain, T0: np.ndarray, alg, nde_params: np.ndarray, T_scaling, wT_scaling,
              ctx: Optional[engine.Context] = None, n_substeps: Optional[int] = None) -> np.ndarray:
    """solve.jl:1-6 for a batch: T0 [n_sim, Nz] scaled, nde_params [n_sim, 6] -> scaled T [n_sim, n_saved, Nz]."""
    own = ctx is None
    ctx = ctx or engine.Context(0)
    d = _desc(ndes[0], T_scaling, wT_scaling, _integrator(alg), n_substeps)
    m = engine.Model(ctx, d, destructure(NN)[0])
    try:
        return m.solve(np.atleast_2d(T0), np.ascontiguousarray(np.atleast_2d(nde_params)[:, :2], dtype=np.float32))
    finally:
        m.close()
        if own:
            ctx.close()


def train_neural_differential_equation(NN: Chain, NDEType, algorithm, datasets: Dict[int, FreeConvectionDataset], T_scaling,
                                       wT_scaling, iterations, opt: ADAM, epochs: int, history: Optional[List[float]] = None,
                                       ctx: Optional[engine.Context] = None, n_substeps: Optional[int] = None) -> Chain:
    """training.jl:44-74: loss = Flux.mse over all simulations' saved frames; one ADAM step per epoch
    (Flux.train! over Iterators.repeated((), epochs)). Returns the trained Chain (the reference mutates NN in place)."""
    own = ctx is None
    ctx = ctx or engine.Context(0)
    ids = sorted(datasets.keys())
    ndes = [NDEType(NN, datasets[i], iterations) for i in ids]
    params = np.stack([FreeConvectionNDEParameters(datasets[i], T_scaling, wT_scaling) for i in ids])
    its = ndes[0].iterations
    T0 = np.stack([T_scaling(datasets[i].T[its[0]]) for i in ids]).astype(np.float32)
    true_sols = np.stack([T_scaling(datasets[i].T[its]) for i in ids]).astype(np.float32)
    d = _desc(ndes[0], T_scaling, wT_scaling, _integrator(algorithm), n_substeps)
    theta, re = destructure(NN)
    m = engine.Model(ctx, d, theta)
    w = np.array([0, 0, 1, 0, 0, 0], dtype=np.float32)
    bcs = np.ascontiguousarray(params[:, :2], dtype=np.float32)
    try:
        if opt.state:
            m.set_adam_state(opt.state["m"], opt.state["v"], opt.state["beta_pow"])
        for e in range(epochs):
            l = m.train_step(T0, bcs, true_sols, w, opt.eta, opt.beta[0], opt.beta[1], opt.eps)
            log.info("Training free convection NDE... MSE loss: %.12e", float(l[6]))
            if history is not None:
                history.append(float(l[6]))
        mm, vv, bp = m.adam_state()
        opt.state = {"m": mm, "v": vv, "beta_pow": bp}
        return re(m.get_theta())
    finally:
        m.close()
        if own:
            ctx.close()
