"""8f-1: implicit (backward-Euler, tridiagonal) treatment of the vertical-diffusion part of the RHS, NN flux explicit
(CPZ_FLAG_IMPLICIT_DIFFUSION) — the way the reference integrates inside Oceananigans (modified_pacanowski_philander!,
wind_mixing/src/NDE_oceananigans.jl:61-101; convective_adjustment!, free_convection/src/oceananigans_nn.jl:13-40).
CUDA through the C ABI vs the FP64 oracle (oracle/nde.py: implicit_diffusion) on identical inputs; tolerance 1e-4."""
import os

import numpy as np
import pytest
import torch

from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import FLAG_CA, FLAG_IMPLICIT_DIFFUSION as IMP, FLAG_MPP, FLAG_ZERO_WEIGHTS, RHS_INFER, RHS_TRAIN
from util import oracle_rhs, oracle_solve, rel_inf

pytestmark = pytest.mark.gpu
TOL = 1e-4


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)


CASES = {
    "uvT_infer": lambda **kw: syn.wind_mixing_desc(variant=RHS_INFER, flags=FLAG_MPP | FLAG_ZERO_WEIGHTS | IMP, n_substeps=1, **kw),
    "uvT_infer_ca_kappa10": lambda **kw: syn.wind_mixing_desc(variant=RHS_INFER, flags=FLAG_MPP | FLAG_CA | IMP, n_substeps=1, kappa=10.0, **kw),
    "uvT_train": lambda **kw: syn.wind_mixing_desc(variant=RHS_TRAIN, flags=FLAG_MPP | FLAG_ZERO_WEIGHTS | IMP, n_substeps=1, **kw),
    "nn_free": lambda **kw: syn.wind_mixing_desc(variant=RHS_INFER, net=None, flags=FLAG_MPP | FLAG_ZERO_WEIGHTS | IMP, n_substeps=1, **kw),
    "uvT_test_net": lambda **kw: syn.wind_mixing_desc(variant=RHS_INFER, net="uvT_test", flags=FLAG_MPP | FLAG_ZERO_WEIGHTS | IMP, n_substeps=1, **kw),
}


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("integrator", ["tsit5", "euler"])
def test_implicit_diffusion_solve_parity(ctx, case, integrator):
    """One step per 600-s frame interval, no stability sub-steps — including the reference's kappa = 10 (NDE_training.jl:168),
    which explicit Tsit5 could only take with >= 118 sub-steps."""
    d = CASES[case](n_steps=48, save_stride=4, integrator=integrator)
    th = syn.theta_random(d, scale=0.1) if d.nets else np.zeros(0, dtype=np.float32)
    x0, bcs = syn.columns(d, 77)
    m = engine.Model(ctx, d, th)
    desc = m.describe().splitlines()[0]
    got = m.solve(x0, bcs)
    with _env(CPZ_NO_TC="1"):
        simt = m.solve(x0, bcs)
    dx = m.rhs(x0, bcs, t=0.1)
    m.close()
    ref = oracle_solve(d, th, x0, bcs)
    floor = rel_inf(oracle_solve(d, th, x0, bcs, dtype=torch.float32), ref)
    e, e_s = rel_inf(got, ref), rel_inf(simt, ref)
    e_r = rel_inf(dx, oracle_rhs(d, th, x0, bcs, 0.1))
    print(f"implicit diffusion {case} {integrator}: {desc.split(':')[1][:28]} {e:.2e}  fp32 tiles {e_s:.2e}  fp32-oracle {floor:.2e}  explicit-part rhs {e_r:.2e}")
    assert np.isfinite(got).all()
    assert e_r <= 1e-5                                  # cpz_rhs returns the explicit part (what the stages evaluate)
    # the kappa = 10 convective-adjustment switch flips under rounding (the FP32 oracle itself is 3e-4 .. 2e-3 away); the tcgen05
    # kernel solves the tridiagonal systems by parallel cyclic reduction across the warp, a different rounding path than the
    # oracle's Thomas sweep, hence the wider multiple of the FP32-oracle floor for it
    assert e <= max(TOL, 5 * floor) and e_s <= max(TOL, 3 * floor)


@pytest.mark.parametrize("ca,mpp", [(True, False), (True, True)])
@pytest.mark.parametrize("ncol", [1, 200])
def test_implicit_diffusion_T_only(ctx, ca, mpp, ncol):
    """The T-only NDE with implicit convective adjustment (K = 10) [+ mPP base]: one step per frame instead of 4 / 15 sub-steps."""
    d = syn.free_convection_desc(ca=ca, mpp=mpp, n_steps=45, save_stride=9, n_substeps=1)
    d.flags |= IMP
    th = syn.theta_random(d, scale=0.3)
    x0, bcs = syn.columns(d, ncol)
    x0[:, 8:16] = x0[:, 15:7:-1].copy()
    m = engine.Model(ctx, d, th)
    got = m.solve(x0, bcs)
    m.close()
    ref = oracle_solve(d, th, x0, bcs)
    floor = rel_inf(oracle_solve(d, th, x0, bcs, dtype=torch.float32), ref)
    e = rel_inf(got, ref)
    print(f"implicit diffusion T-only ca={ca} mpp={mpp} ncol={ncol}: {e:.2e} (fp32-oracle {floor:.2e})")
    assert e <= max(TOL, 3 * floor)


def test_implicit_converges_to_the_explicit_solution(ctx):
    """Lie splitting is first order: halving the step halves the distance to the explicit Tsit5 solution."""
    d_ex = syn.wind_mixing_desc(variant=RHS_INFER, n_steps=36, save_stride=36)
    th = syn.theta_random(d_ex, scale=0.1)
    x0, bcs = syn.columns(d_ex, 40)
    m = engine.Model(ctx, d_ex, th)
    ref = m.solve(x0, bcs)[:, -1]
    m.close()
    errs = []
    for nsub in (1, 2, 4):
        d = syn.wind_mixing_desc(variant=RHS_INFER, flags=FLAG_MPP | FLAG_ZERO_WEIGHTS | IMP, n_steps=36, save_stride=36, n_substeps=nsub)
        m = engine.Model(ctx, d, th)
        errs.append(rel_inf(m.solve(x0, bcs)[:, -1], ref))
        m.close()
    print("implicit vs explicit, 1/2/4 sub-steps:", errs)
    assert errs[1] < 0.65 * errs[0] and errs[2] < 0.65 * errs[1]


# ---- discrete adjoint through the implicit step (VJP of the Thomas solve) ------------------------------------------------
W_GRAD = np.array([0.7, 0.7, 1.0, 3e-3, 3e-3, 5e-3], dtype=np.float32)
W_T = np.array([0, 0, 1.0, 0, 0, 0], dtype=np.float32)


def _grad_case(ctx, d, th, ncol, w, unstable=False, env=None):
    from util import oracle_loss_grad
    x0, bcs = syn.columns(d, ncol)
    if unstable:
        x0[:, 8:16] = x0[:, 15:7:-1].copy()
    rng = np.random.default_rng(0)
    th2 = (th * (1 + 0.3 * rng.standard_normal(th.shape))).astype(np.float32)
    tgt = (oracle_solve(d, th2, x0, bcs) + 0.02 * rng.standard_normal((ncol, d.n_saved, d.S))).astype(np.float32)
    m = engine.Model(ctx, d, th)
    with _env(**(env or {})):
        loss, grad = m.loss_grad(x0, bcs, tgt, w)
    m.close()
    tot, comps, g = oracle_loss_grad(d, th, x0, bcs, tgt, w)
    tot32, _, g32 = oracle_loss_grad(d, th, x0, bcs, tgt, w, dtype=torch.float32)
    return (abs(loss[6] - tot) / abs(tot), np.linalg.norm(grad - g) / np.linalg.norm(g), abs(tot32 - tot) / abs(tot),
            np.linalg.norm(g32 - g) / np.linalg.norm(g))


@pytest.mark.parametrize("ncol", [9, 70])
@pytest.mark.parametrize("variant,flags,kw", [(RHS_TRAIN, FLAG_MPP | FLAG_ZERO_WEIGHTS, {}), (RHS_INFER, FLAG_MPP | FLAG_CA, dict(kappa=1.0))])
@pytest.mark.parametrize("ckpt", [3, 1])
def test_implicit_gradient_uvT(ctx, ncol, variant, flags, kw, ckpt):
    d = syn.wind_mixing_desc(variant=variant, flags=flags | IMP, n_steps=12, save_stride=3, ckpt_stride=ckpt, n_substeps=1, **kw)
    th = syn.theta_random(d, scale=0.3)
    e_l, e_g, f_l, f_g = _grad_case(ctx, d, th, ncol, W_GRAD)
    # default: the tensor-core adjoint (its reverse sweep carries the VJP of the cyclic-reduction solve); CPZ_NO_TC_ADJ=1: the FP32
    # adjoint with the tcgen05 forward pass's stored tendencies, and with CPZ_NO_KSTORE=1 on top the recomputing FP32 path
    e_l2, e_g2, _, _ = _grad_case(ctx, d, th, ncol, W_GRAD, env=dict(CPZ_NO_TC_ADJ="1"))
    e_l3, e_g3, _, _ = _grad_case(ctx, d, th, ncol, W_GRAD, env=dict(CPZ_NO_TC_ADJ="1", CPZ_NO_KSTORE="1"))
    e_l4, e_g4, _, _ = _grad_case(ctx, d, th, ncol, W_GRAD, env=dict(CPZ_TC_SPLIT="0"))  # two column groups per CTA (large batches)
    assert e_l4 <= max(TOL, 3 * f_l) and e_g4 <= max(TOL, 3 * f_g), (e_l4, e_g4)
    m = engine.Model(ctx, d, th)
    desc = m.describe()
    m.close()
    print(f"implicit-diffusion gradient variant={variant} ncol={ncol} ckpt={ckpt}: loss {e_l:.2e} grad tensor-core {e_g:.2e} fp32 {e_g2:.2e} "
          f"fp32 recompute {e_g3:.2e} (fp32-oracle {f_g:.2e})")
    assert "adjoint kernels: tcgen05" in desc
    assert e_l <= max(TOL, 3 * f_l) and e_l2 <= max(TOL, 3 * f_l) and e_l3 <= max(TOL, 3 * f_l)
    assert e_g <= max(TOL, 3 * f_g) and e_g2 <= max(TOL, 3 * f_g) and e_g3 <= max(TOL, 3 * f_g)


@pytest.mark.parametrize("ncol", [1, 70])
def test_implicit_gradient_T_only_ca_mpp(ctx, ncol):
    """BASELINE config 1's model with the implicit treatment: ONE step per frame instead of 15 sub-steps."""
    d = syn.free_convection_desc(ca=True, mpp=True, n_steps=18, save_stride=9, ckpt_stride=9, n_substeps=1)
    d.flags |= IMP
    th = syn.theta_random(d, scale=0.3)
    e_l, e_g, f_l, f_g = _grad_case(ctx, d, th, ncol, W_T, unstable=True)
    print(f"implicit-diffusion gradient T-only CA+mPP ncol={ncol}: loss {e_l:.2e} grad {e_g:.2e} (fp32-oracle {f_g:.2e})")
    assert e_l <= max(TOL, 3 * f_l) and e_g <= max(TOL, 3 * f_g)


def test_implicit_mpp_parameter_gradient(ctx):
    """d loss / d (nu0, nu_m, dRi, Ric, Pr) through the implicit step of the NN-free base closure (8f-1 + 8f-3)."""
    from oracle import nde
    from util import t32, t64
    d = syn.wind_mixing_desc(variant=RHS_TRAIN, net=None, flags=FLAG_MPP | FLAG_ZERO_WEIGHTS | IMP, n_steps=12, save_stride=4, ckpt_stride=4, n_substeps=1)
    th = np.zeros(0, dtype=np.float32)
    x0, bcs = syn.columns(d, 70)
    d_true = syn.wind_mixing_desc(variant=RHS_TRAIN, net=None, flags=FLAG_MPP | FLAG_ZERO_WEIGHTS | IMP, n_steps=12, save_stride=4, ckpt_stride=4,
                                  n_substeps=1, nu_m=0.07, Ric=0.3, Pr=1.3, dRi=0.15, nu0=2e-4)
    tgt = oracle_solve(d_true, th, x0, bcs).astype(np.float32)
    m = engine.Model(ctx, d, th)
    loss, gp, _ = m.loss_grad_mpp(x0, bcs, tgt, W_GRAD)
    m.close()
    tot, g = nde.loss_grad_mpp(d, t64(th), t64(x0), t64(bcs), None, t64(tgt), W_GRAD)
    _, g32 = nde.loss_grad_mpp(d, t32(th), t32(x0), t32(bcs), None, t32(tgt), W_GRAD)
    p = np.array([d.nu0, d.nu_m, d.dRi, d.Ric, d.Pr])
    g, g32 = g.numpy(), g32.numpy().astype(np.float64)
    e_s = np.linalg.norm((gp - g) * p) / np.linalg.norm(g * p)
    f_s = np.linalg.norm((g32 - g) * p) / np.linalg.norm(g * p)
    print(f"implicit-diffusion mPP-parameter gradient: loss {abs(loss[6] - float(tot)) / abs(float(tot)):.2e} scaled-space L2 {e_s:.2e} (fp32-oracle {f_s:.2e}) "
          f"per parameter {np.array2string(np.abs(gp - g) / np.abs(g), precision=1)}")
    assert abs(loss[6] - float(tot)) / abs(float(tot)) <= TOL and e_s <= max(TOL, 3 * f_s)
