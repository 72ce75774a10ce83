"""BASELINE config 1: the T-only NDE with convective adjustment AND the mPP base diffusivity, and the small-batch
(4-column-tile) training pass.

The reference has no trainable "CA + mPP" model (`train_NDE` asserts they are exclusive and the stand-alone CA
branch reads an undefined kappa); BASELINE config 1 nevertheless names it, so the engine defines it as
    wT = [bottom; NN(T); top] - sigma_T/(sigma_wT H) nu/Pr dT/dz        (NDE_training.jl:114-139 at u = v = 0)
    dT/dt = sigma_wT/sigma_T tau/H (-D_c wT + D_c min(0, K dT/dz))         (convective_adjustment_nde.jl:41-47)
and the oracle restates exactly that (oracle/literal.py: rhs_free_convection, oracle/nde.py). Parity: RHS 1e-5,
profiles / loss / gradient 1e-4 (or the FP32-oracle floor), CUDA through the C ABI vs the FP64 oracle."""
import os

import numpy as np
import pytest
import torch

from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import RHS_TRAIN
from oracle import literal
from util import oracle_loss_grad, oracle_rhs, oracle_solve, rel_inf

pytestmark = pytest.mark.gpu
TOL = 1e-4
W_T = np.array([0, 0, 1.0, 0, 0, 0], dtype=np.float32)
W_GRAD = np.array([0.7, 0.7, 1.0, 3e-3, 3e-3, 5e-3], dtype=np.float32)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


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


def _unstable(x0):
    """turn part of every profile upside down so that both switches (dT/dz < 0 and Ri < Ric) act"""
    x0 = x0.copy()
    x0[:, 8:16] = x0[:, 15:7:-1].copy()
    return x0


@pytest.mark.parametrize("ca,mpp", [(True, True), (False, True)])
@pytest.mark.parametrize("ncol", [1, 50, 300])
def test_rhs_T_only_with_mpp_base(ctx, ca, mpp, ncol):
    d = syn.free_convection_desc(ca=ca, mpp=mpp)
    th = syn.theta_random(d, scale=1.0)
    x, bcs = syn.columns(d, ncol)
    x = _unstable(x)
    m = engine.Model(ctx, d, th)
    got = m.rhs(x, bcs)
    with _env(CPZ_NO_TC="1"):
        simt = m.rhs(x, bcs)
    m.close()
    ref = oracle_rhs(d, th, x, bcs, 0.0)
    lit = np.stack([literal.rhs(d, th, x[i], bcs[i], 0.0) for i in range(min(ncol, 4))])
    e, e_s = rel_inf(got, ref), rel_inf(simt, ref)
    print(f"T-only ca={ca} mpp={mpp} ncol={ncol}: rhs tcgen05 {e:.2e}  simt {e_s:.2e}  literal-vs-batched oracle {rel_inf(ref[:len(lit)], lit):.1e}")
    assert rel_inf(ref[:len(lit)], lit) <= 1e-12
    assert e <= 1e-5 and e_s <= 1e-5
    # the mPP term is there: it changes the tendencies
    d0 = syn.free_convection_desc(ca=ca, mpp=False)
    assert rel_inf(oracle_rhs(d0, th, x, bcs, 0.0), ref) > 1e-3


@pytest.mark.parametrize("ncol", [1, 70])
def test_solve_T_only_ca_mpp(ctx, ncol):
    d = syn.free_convection_desc(ca=True, mpp=True, n_steps=45, save_stride=9)
    th = syn.theta_random(d, scale=0.3)
    x0, bcs = syn.columns(d, ncol)
    x0 = _unstable(x0)
    m = engine.Model(ctx, d, th)
    got = m.solve(x0, bcs)
    with _env(CPZ_NO_TC="1"):
        simt = m.solve(x0, bcs)
    m.close()
    ref = oracle_solve(d, th, x0, bcs)
    floor = rel_inf(oracle_solve(d, th, x0, bcs, dtype=torch.float32), ref)
    e, e_s = rel_inf(got, ref), rel_inf(simt, ref)
    print(f"T-only CA+mPP solve ncol={ncol} ({d.n_substeps} sub-steps): tcgen05 {e:.2e}  simt {e_s:.2e}  fp32-oracle {floor:.2e}")
    assert e <= max(TOL, 3 * floor) and e_s <= max(TOL, 3 * floor)


def _grad_check(ctx, d, th, x0, bcs, w, env=None):
    rng = np.random.default_rng(0)
    th2 = (th * (1 + 0.3 * rng.standard_normal(th.shape))).astype(np.float32)
    tgt = oracle_solve(d, th2, x0, bcs).astype(np.float32)
    m = engine.Model(ctx, d, th)
    with _env(**(env or {})):
        loss, grad = m.loss_grad(x0, bcs, tgt, w)
    desc = m.describe()
    m.close()
    tot, comps, g = oracle_loss_grad(d, th, x0, bcs, tgt, w)
    g32 = oracle_loss_grad(d, th, x0, bcs, tgt, w, dtype=torch.float32)[2]
    floor = np.linalg.norm(g32 - g) / np.linalg.norm(g)
    e_l, e_g = abs(loss[6] - tot) / abs(tot), np.linalg.norm(grad - g) / np.linalg.norm(g)
    return e_l, e_g, floor, grad, desc


@pytest.mark.parametrize("ncol", [1, 9, 70])
def test_grad_T_only_ca_mpp(ctx, ncol):
    """ncol = 1 and 9 take the 4-column tiles, 70 the 32-column ones (forced), both against the FP64 oracle."""
    d = syn.free_convection_desc(ca=True, mpp=True, n_steps=18, save_stride=9, ckpt_stride=9)
    th = syn.theta_random(d, scale=0.3)
    x0, bcs = syn.columns(d, ncol)
    x0 = _unstable(x0)
    # default: ncol <= 32 runs one CTA per column (fc1_train_kernel); CPZ_FC1_MAX_NCOL=0 falls back to the 4-column tiles,
    # CPZ_SMALL_NCOL=0 on top of that to the 32-column tiles
    e_l0, e_g0, floor, g_fc1, desc = _grad_check(ctx, d, th, x0, bcs, W_T)
    e_l, e_g, _, g_small, _ = _grad_check(ctx, d, th, x0, bcs, W_T, env=dict(CPZ_FC1_MAX_NCOL="0"))
    e_l2, e_g2, _, g_big, _ = _grad_check(ctx, d, th, x0, bcs, W_T, env=dict(CPZ_FC1_MAX_NCOL="0", CPZ_SMALL_NCOL="0"))
    print(f"T-only CA+mPP gradient ncol={ncol}: one CTA per column loss {e_l0:.2e} grad {e_g0:.2e} | 4-column tiles loss {e_l:.2e} grad {e_g:.2e} | "
          f"32-column tiles loss {e_l2:.2e} grad {e_g2:.2e} (fp32-oracle {floor:.2e})")
    assert "small-batch training pass" in desc and "one CTA per column" in desc
    assert e_l0 <= TOL and e_l <= TOL and e_l2 <= TOL
    assert e_g0 <= max(TOL, floor) and e_g <= max(TOL, floor) and e_g2 <= max(TOL, floor)
    assert np.linalg.norm(g_small - g_big) <= max(TOL, floor) * np.linalg.norm(g_big)
    assert np.linalg.norm(g_fc1 - g_big) <= max(TOL, floor) * np.linalg.norm(g_big)


@pytest.mark.parametrize("h1,h2,act,integrator,flags", [(128, 128, "relu", "tsit5", "ca+mpp"), (50, 20, "mish", "rk4", "ca"), (33, 127, "tanh", "euler", "mpp"),
                                                      (8, 5, "swish", "tsit5", "none"), (64, 48, "leakyrelu", "tsit5", "ca+mpp"),
                                                      (128, 128, "relu", "tsit5", "ca+mpp+implicit"), (50, 20, "mish", "euler", "ca+implicit")])
def test_single_column_kernel_net_shapes_and_integrators(ctx, h1, h2, act, integrator, flags):
    """fc1_train_kernel pads every layer to 128-wide blocks in shared memory: odd widths, every activation, every tableau,
    final-state-only and strided saves, 1..5 columns — loss and gradient against the FP64 oracle and the tile kernels."""
    from cpz_b200.desc import NetDesc
    for ncol, save in ((1, 3), (5, 0)):
        d = syn.free_convection_desc(ca="ca" in flags, mpp="mpp" in flags, n_steps=12, save_stride=save, ckpt_stride=4, integrator=integrator, net=None)
        d.nets = [NetDesc([32, h1, h2, 31], [act, act, "identity"])]
        if "implicit" in flags:  # backward-Euler diffusion / convective adjustment at the start of every (single) sub-step
            from cpz_b200.desc import FLAG_IMPLICIT_DIFFUSION
            d.flags |= FLAG_IMPLICIT_DIFFUSION
            d.n_substeps = 1
        th = syn.theta_random(d, scale=0.4)
        x0, bcs = syn.columns(d, ncol)
        x0 = _unstable(x0)
        e_l, e_g, floor, g_fc1, desc = _grad_check(ctx, d, th, x0, bcs, W_T)
        _, e_gt, _, g_tile, _ = _grad_check(ctx, d, th, x0, bcs, W_T, env=dict(CPZ_FC1_MAX_NCOL="0"))
        print(f"single-column kernel {h1}x{h2} {act} {integrator} {flags} ncol={ncol} save={save}: loss {e_l:.2e} grad {e_g:.2e} (tiles {e_gt:.2e}, fp32-oracle {floor:.2e})")
        assert "one CTA per column" in desc
        if "implicit" in flags:
            # the K = 10 convective-adjustment rows make the backward-Euler solve ill-conditioned in FP32 (cyclic reduction here,
            # Thomas sweeps in the tile kernels, both a few 1e-4 from the FP64 oracle): the implicit tests' 3 x floor rule
            assert e_l <= 5e-4 and e_g <= max(TOL, 3 * floor) and e_gt <= max(TOL, 3 * floor)
            continue
        assert e_l <= TOL and e_g <= max(TOL, 1.5 * floor)  # the FP32 oracle itself sits at 1.3e-4 on the mPP-only Euler case
        assert np.linalg.norm(g_fc1 - g_tile) <= max(TOL, 2 * floor) * max(np.linalg.norm(g_tile), 1e-30)


@pytest.mark.parametrize("ncol,ckpt", [(1, 3), (9, 1), (18, 9), (45, 4)])
def test_grad_uvT_small_batches_take_small_tiles(ctx, ncol, ckpt):
    """The reference trains on 9-18 simulations: those batches run on 4-column tiles (tcgen05 forward + stored stage
    tendencies in the 32-column layout, FP32 adjoint on 4-column tiles) and must match the oracle and the 32-column path."""
    d = syn.wind_mixing_desc(variant=RHS_TRAIN, n_steps=18, save_stride=9, ckpt_stride=ckpt)
    th = syn.theta_random(d, scale=0.3)
    x0, bcs = syn.columns(d, ncol)
    e_l, e_g, floor, g_small, _ = _grad_check(ctx, d, th, x0, bcs, W_GRAD)
    _, e_gn, _, g_nok, _ = _grad_check(ctx, d, th, x0, bcs, W_GRAD, env=dict(CPZ_NO_KSTORE="1"))
    _, e_gs, _, g_simt, _ = _grad_check(ctx, d, th, x0, bcs, W_GRAD, env=dict(CPZ_NO_TC="1"))
    _, e_g2, _, g_big, _ = _grad_check(ctx, d, th, x0, bcs, W_GRAD, env=dict(CPZ_SMALL_NCOL="0"))
    print(f"u/v/T gradient ncol={ncol} ckpt={ckpt}: 4-col tiles {e_g:.2e} (no kstore {e_gn:.2e}, simt forward {e_gs:.2e}) | 32-col tiles {e_g2:.2e} (fp32-oracle {floor:.2e})")
    assert e_l <= TOL
    for e in (e_g, e_gn, e_gs, e_g2):
        assert e <= max(TOL, floor)


def test_config1_single_column_full_length(ctx):
    """BASELINE config 1 as stated: ONE synthetic column, 1152 steps, forward solve + loss gradient, vs the FP64 oracle
    answers of tests/golden/fullsize_config1_*.npz (make_fullsize.py)."""
    for name, scale in (("fullsize_config1_random", 0.1), ("fullsize_config1_init", 1e-5)):
        path = os.path.join(GOLDEN, name + ".npz")
        if not os.path.exists(path):
            pytest.skip("golden fixture missing: run tests/golden/make_fullsize.py")
        G = np.load(path)
        d = syn.free_convection_desc(ca=True, mpp=True, n_steps=1152, save_stride=9, ckpt_stride=9)
        th = syn.theta_init(d, seed=42, scale=scale) if scale < 1e-3 else syn.theta_random(d, scale=scale)
        x0, bcs = syn.columns(d, 1)
        m = engine.Model(ctx, d, th)
        traj = m.solve(x0, bcs)
        loss, grad = m.loss_grad(x0, bcs, G["targets"], W_T)
        m.close()
        e_t = rel_inf(traj, G["traj"])
        e_l = abs(loss[6] - G["loss"][6]) / abs(G["loss"][6])
        e_g = np.linalg.norm(grad - G["grad"]) / np.linalg.norm(G["grad"])
        print(f"config 1 ({name}, {d.n_substeps} sub-steps): profiles {e_t:.2e} (fp32-oracle {float(G['floor_traj']):.2e})  loss {e_l:.2e} "
              f"(fp32-oracle {float(G['floor_loss']):.2e})  gradient {e_g:.2e} (fp32-oracle {float(G['floor_grad']):.2e})")
        assert e_t <= max(TOL, 3 * float(G["floor_traj"]))
        assert e_l <= max(TOL, 3 * float(G["floor_loss"]))
        assert e_g <= max(TOL, 3 * float(G["floor_grad"]))
