"""The host-side mirror of the reference interface (train_NDE, solve_NDE_mutating, predict_NDE, FreeConvection NDEs)
driven end to end on the GPU, written the way the reference's own driver scripts call it
(wind_mixing/train_NDE.jl:103-191, free_convection/train_free_convection_nde.jl:119-266)."""
import ctypes as C

import numpy as np
import pytest
import torch

from cpz_b200 import engine, flux, free_convection as fc, synthetic as syn, wind_mixing as wm
from cpz_b200.desc import FLAG_MPP, FLAG_ZERO_WEIGHTS, RHS_INFER, RHS_TRAIN
from cpz_b200.ocean_parameterizations import ZeroMeanUnitVarianceScaling
from util import oracle_rhs, oracle_solve, rel_inf

pytestmark = pytest.mark.gpu
NAMES = ("u", "v", "T", "uw", "vw", "wT")


def _scalings():
    return {k: ZeroMeanUnitVarianceScaling(mu=syn.MU[i], sigma=syn.SIGMA[i]) for i, k in enumerate(NAMES)}


def _nets(seed, scale):
    rng = np.random.default_rng(seed)
    mk = lambda: flux.Chain(flux.Dense(96, 50, "mish", rng=rng), flux.Dense(50, 20, "mish", rng=rng), flux.Dense(20, 31, rng=rng)).scale(scale)
    return mk(), mk(), mk()


def _dataset(ctx, n_sim=12, n_frames=19):
    """synthetic 'LES': the inference solve of a hidden truth NDE, sampled every 600 s"""
    d = syn.wind_mixing_desc(variant=RHS_TRAIN, n_steps=n_frames - 1, save_stride=1)
    truth = _nets(123, 0.3)
    theta = np.concatenate([flux.destructure(n)[0] for n in truth])
    x0, bcs = syn.columns(d, n_sim)
    m = engine.Model(ctx, d, theta)
    uvT = m.solve(x0, bcs)
    m.close()
    return wm.ProfileData(uvT, np.arange(n_frames) * 600.0 * (1152 / 1152), bcs, _scalings(), np.linspace(-256.0, 0.0, 33)), d


def test_train_NDE_like_the_reference_script(ctx):
    data, d = _dataset(ctx)
    # tau of this short record is (n_frames-1)*600 s, so dt_hat differs from 1/1152: the mirror derives it from data.t
    uw, vw, wT = _nets(7, 1e-5)  # re(weights ./ 1f5), train_NDE.jl:105-107
    rec = wm.TrainingRecord()
    out = wm.train_NDE(uw, vw, wT, data, range(0, 19, 3), "Tsit5", [flux.ADAM(1e-3)], 1, maxiters=15,
                       modified_pacanowski_philander=True, zero_weights=True, train_gradient=True, nu_m=0.1, dRi=0.1,
                       training_fractions={"T": 0.8, "∂T∂z": 0.8, "profile": 0.5}, ctx=ctx, record=rec)
    assert len(out) == 3 and all(isinstance(n, flux.Chain) for n in out)
    assert len(rec.totals) == 15 and np.all(np.isfinite(rec.totals))
    assert min(rec.totals[5:]) < rec.totals[0]
    # loss fractions hold at the initial theta (wind_mixing/test/test_training_scaling.jl identities)
    L = rec.losses[0]
    np.testing.assert_allclose(L["T"] / (L["u"] + L["v"]), 4.0, rtol=1e-3)
    np.testing.assert_allclose((L["u"] + L["v"] + L["T"]) / (L["∂u∂z"] + L["∂v∂z"] + L["∂T∂z"]), 1.0, rtol=1e-3)
    assert rec.adam_state is not None and rec.adam_state["beta_pow"][0] == pytest.approx(0.9 ** 16, rel=1e-5)
    # returned nets are the best-seen parameters: not the initial ones
    assert np.abs(flux.destructure(out[2])[0] - flux.destructure(wT)[0]).max() > 0


def test_solve_NDE_mutating_and_predict_NDE(ctx):
    data, d = _dataset(ctx, n_sim=5, n_frames=10)
    uw, vw, wT = _nets(9, 0.3)
    constants = dict(H=256.0, tau=float(data.t[-1]), f=1e-4, Nz=32, g=9.80665, alpha=2e-4, nu0=1e-4, nu_m=0.1, Ric=0.25, dRi=0.1, Pr=1.0)
    conditions = dict(modified_pacanowski_philander=True, zero_weights=True)
    ts = np.arange(10) / 9.0
    sol = wm.solve_NDE_mutating(uw, vw, wT, data.scalings, constants, wm.prepare_BCs(data), None, data.uvT_scaled[:, 0], ts,
                                "Tsit5", conditions, ctx=ctx)
    assert sol.shape == (5, 10, 96)
    dd = syn.wind_mixing_desc(variant=RHS_INFER, n_steps=9, save_stride=1, tau=constants["tau"], dt=1 / 9.0)
    dd.n_substeps = wm.default_substeps(constants, conditions, 1 / 9.0, 32, "tsit5")
    theta = np.concatenate([flux.destructure(n)[0] for n in (uw, vw, wT)])
    ref = oracle_solve(dd, theta, data.uvT_scaled[:, 0], data.bcs_scaled)
    assert rel_inf(sol, ref) <= 1e-4
    dx = wm.predict_NDE(uw, vw, wT, data.uvT_scaled[:, 3], data.bcs_scaled, conditions, data.scalings, constants, ctx)
    dt = syn.wind_mixing_desc(variant=RHS_TRAIN, tau=constants["tau"])
    assert rel_inf(dx, oracle_rhs(dt, theta, data.uvT_scaled[:, 3], data.bcs_scaled)) <= 1e-5
    with pytest.raises(ValueError, match="ROCK4"):
        wm.solve_NDE_mutating(uw, vw, wT, data.scalings, constants, data.bcs_scaled, None, data.uvT_scaled[:, 0], ts, "ROCK4", conditions, ctx=ctx)


def test_free_convection_train_and_solve(ctx):
    rng = np.random.default_rng(3)
    T_scaling = ZeroMeanUnitVarianceScaling(mu=19.8, sigma=0.15)
    wT_scaling = ZeroMeanUnitVarianceScaling(mu=5e-6, sigma=6e-6)
    Nz, Nt = 32, 28
    z = -100 + (np.arange(Nz) + 0.5) * 100 / Nz
    datasets = {}
    for i in range(4):
        T0 = 20 + 0.01 * np.minimum(z, -rng.uniform(10, 40))
        T = T0[None, :] - 0.002 * np.linspace(0, 1, Nt)[:, None] * np.exp(z / 30)[None, :] * (1 + i)
        datasets[i + 1] = fc.FreeConvectionDataset(T=T, wT_bottom=0.0, temperature_flux=(1 + i) * 1e-5 / 2, H=100.0,
                                                   times=np.arange(Nt) * 600.0)
    NN = flux.Chain(flux.Dense(32, 128, "relu", rng=rng), flux.Dense(128, 128, "relu", rng=rng), flux.Dense(128, 31, rng=rng)).scale(1e-2)
    hist = []
    NN2 = fc.train_neural_differential_equation(NN, fc.ConvectiveAdjustmentNDE, "Tsit5", datasets, T_scaling, wT_scaling,
                                                range(0, 28, 9), flux.ADAM(1e-3), 12, history=hist, ctx=ctx)
    assert len(hist) == 12 and np.all(np.isfinite(hist)) and min(hist[4:]) < hist[0]
    ndes = [fc.ConvectiveAdjustmentNDE(NN2, datasets[i], range(0, 28, 9)) for i in sorted(datasets)]
    params = np.stack([fc.FreeConvectionNDEParameters(datasets[i], T_scaling, wT_scaling) for i in sorted(datasets)])
    T0 = np.stack([T_scaling(datasets[i].T[0]) for i in sorted(datasets)]).astype(np.float32)
    sol = fc.solve_nde(ndes, NN2, T0, "Tsit5", params, T_scaling, wT_scaling, ctx=ctx)
    assert sol.shape == (4, 4, 32) and np.isfinite(sol).all()
    np.testing.assert_array_equal(sol[:, 0], T0)


def test_allreduce_hook_plumbing_on_one_gpu(ctx):
    """A hook that doubles the packed buffer emulates two ranks holding identical shards: loss and gradient must be
    unchanged (sum of two equal packs, normalised by the doubled column count)."""
    d = syn.wind_mixing_desc(variant=RHS_TRAIN, n_steps=6, save_stride=3, ckpt_stride=3)
    th = syn.theta_random(d, scale=0.3)
    x0, bcs = syn.columns(d, 40)
    tgt = oracle_solve(d, (th * 1.2).astype(np.float32), x0, bcs).astype(np.float32)
    w = np.array([0.7, 0.7, 1.0, 3e-3, 3e-3, 5e-3], dtype=np.float32)
    m = engine.Model(ctx, d, th)
    l1, g1 = m.loss_grad(x0, bcs, tgt, w)
    m.close()
    c2 = engine.Context(0, torch.cuda.current_stream().cuda_stream or None)
    calls = []

    def hook(ptr, n, stream):
        from cpz_b200.parallel import _DevBuf
        t = torch.as_tensor(_DevBuf(ptr, n), device="cuda:0")
        calls.append(n)
        with torch.cuda.stream(torch.cuda.ExternalStream(stream)) if stream else torch.cuda.stream(torch.cuda.current_stream()):
            t.mul_(2.0)

    c2.set_allreduce(hook, 0, 2)
    m2 = engine.Model(c2, d, th)
    l2, g2 = m2.loss_grad(x0, bcs, tgt, w)
    m2.close(); c2.close()
    assert calls == [d.n_params + 16]  # [grad; 6 sums; ncol; pad; 5 mPP-parameter sums; pad]
    np.testing.assert_allclose(l2, l1, rtol=1e-6)
    np.testing.assert_allclose(g2, g1, rtol=1e-5, atol=1e-9)


def test_nonfinite_results_are_reported_not_hidden(ctx):
    """CPZ_ERR_NONFINITE: a NaN in the input propagates through relu (Julia's max(0, x) semantics, not fmaxf) into the
    final frame; the host flavour returns the results as computed AND the error code; the device flavour counts."""
    d = syn.wind_mixing_desc(variant=RHS_TRAIN, net=None, n_steps=4, save_stride=2)
    d.nets = [syn.NET_SHAPES["uvT_test"](32, "relu") for _ in range(3)]
    th = syn.theta_random(d, scale=0.3)
    x0, bcs = syn.columns(d, 5)
    m = engine.Model(ctx, d, th)
    ok = m.solve(x0, bcs)
    assert np.isfinite(ok).all()
    x0[3, 40] = np.nan
    out = np.zeros((5, d.n_saved, d.S), dtype=np.float32)
    with pytest.raises(engine.CpzError) as ei:
        m.solve(x0, bcs, out=out)
    assert ei.value.code == engine.ERR_NONFINITE
    assert np.isnan(out[3, -1]).any() and np.isfinite(out[[0, 1, 2, 4]]).all()   # IEEE semantics kept, other columns untouched
    before = ctx.nonfinite_count
    x0d, bcsd = torch.tensor(x0, device="cuda"), torch.tensor(bcs, device="cuda")
    traj = torch.empty((5, d.n_saved, d.S), device="cuda")
    m.solve_dev(x0d, bcsd, traj)
    assert ctx.nonfinite_count > before
    # NaN through every activation of the RHS: Julia propagates it, so must the kernels
    for act in ("relu", "leakyrelu", "tanh", "mish", "swish"):
        d.nets = [syn.NET_SHAPES["uvT_test"](32, act) for _ in range(3)]
        m2 = engine.Model(ctx, d, th)
        dx = m2.rhs(x0, bcs)
        m2.close()
        assert np.isnan(dx[3, 0]), act  # level 0 of u sees the NaN at level 8 of v only through the net's hidden layer
    m.close()


@pytest.mark.parametrize("case", ["uvT_train", "uvT_infer_ca", "T_only", "T_only_ca_mpp"])
def test_predict_flux_matches_the_oracle_and_the_rhs(ctx, case):
    """cpz_predict_flux = predict_flux (NDE_training.jl:83-147) / the wT reconstruction of solve_nde (solve.jl:35-48):
    D_c of the returned face fluxes, times -tau/H sigma_flux/sigma_q (plus Coriolis), must be cpz_rhs's tendency, and the
    T-only fluxes must equal [bottom; NN(T); top] - min(0, 10 dT/dz) evaluated with the oracle's Chain."""
    from cpz_b200.desc import FLAG_CA
    if case == "uvT_train":
        d = syn.wind_mixing_desc(variant=RHS_TRAIN)
    elif case == "uvT_infer_ca":
        d = syn.wind_mixing_desc(variant=RHS_INFER, flags=FLAG_MPP | FLAG_CA)
    else:
        d = syn.free_convection_desc(ca=True, mpp=(case == "T_only_ca_mpp"))
    th = syn.theta_random(d, scale=1.0)
    x, bcs = syn.columns(d, 45)
    if d.n_fields == 1:
        x[:, 8:16] = x[:, 15:7:-1].copy()
    m = engine.Model(ctx, d, th)
    F = m.predict_flux(x, bcs, t=0.1)
    dx = m.rhs(x, bcs, t=0.1)
    m.close()
    N = d.Nz
    assert F.shape == (45, d.n_fields, N + 1)
    ref = oracle_rhs(d, th, x, bcs, 0.1)
    for q in range(d.n_fields):
        qq = q if d.n_fields == 3 else 2
        A = d.tau / d.H * d.sigma[3 + qq] / d.sigma[qq]
        tend = -A * N * (F[:, q, 1:].astype(np.float64) - F[:, q, :-1])
        if d.n_fields == 3 and q == 0:
            tend += d.f * d.tau / d.sigma[0] * (d.sigma[1] * x[:, N:2 * N] + d.mu[1])
        if d.n_fields == 3 and q == 1:
            tend -= d.f * d.tau / d.sigma[1] * (d.sigma[0] * x[:, :N] + d.mu[0])
        scale = np.abs(ref).max()
        assert np.abs(tend - ref[:, q * N:(q + 1) * N]).max() / scale <= 2e-5, (case, q)
        assert np.abs(tend - dx[:, q * N:(q + 1) * N]).max() / scale <= 2e-5
    if case == "T_only":
        from oracle.flux_nn import chain_numpy
        nn = np.stack([chain_numpy(th.astype(np.float64), d.nets[0].sizes, d.nets[0].acts, x[i].astype(np.float64)) for i in range(45)])
        G = N * (x[:, 1:].astype(np.float64) - x[:, :-1])
        wT = np.concatenate([bcs[:, :1], nn - np.minimum(0, d.K_ca * G), bcs[:, 1:]], axis=1)
        assert rel_inf(F[:, 0], wT) <= 1e-5


def test_solve_nde_six_argument_method_and_causal_penalty(ctx):
    """solve_nde(ds, NN, NDEType, alg, T_scaling, wT_scaling) (solve.jl:8-51) returns unscaled T and the reconstructed wT;
    train_neural_differential_equation! accepts a causal penalty (training.jl:57-58, train_free_convection_nde.jl:190-195)."""
    rng = np.random.default_rng(3)
    T_scaling = ZeroMeanUnitVarianceScaling(mu=19.8, sigma=0.15)
    wT_scaling = ZeroMeanUnitVarianceScaling(mu=5e-6, sigma=6e-6)
    Nt, z = 19, -100 + (np.arange(32) + 0.5) * 100 / 32
    T = (20 + 0.01 * np.minimum(z, -30.0))[None, :] - 0.002 * np.linspace(0, 1, Nt)[:, None] * np.exp(z / 30)[None, :]
    ds = fc.FreeConvectionDataset(T=T, wT_bottom=0.0, temperature_flux=1e-5, H=100.0, times=np.arange(Nt) * 600.0)
    NN = flux.Chain(flux.Dense(32, 128, "relu", rng=rng), flux.Dense(128, 128, "relu", rng=rng), flux.Dense(128, 31, rng=rng)).scale(1e-2)
    out = fc.solve_nde_dataset(ds, NN, fc.ConvectiveAdjustmentNDE, "Tsit5", T_scaling, wT_scaling, ctx=ctx)
    assert out["T"].shape == (Nt, 32) and out["wT"].shape == (Nt, 33) and np.isfinite(out["wT"]).all()
    np.testing.assert_allclose(out["T"][0], T[0], rtol=1e-6)
    np.testing.assert_allclose(out["wT"][:, 0], 0.0, atol=1e-12)      # bottom flux
    np.testing.assert_allclose(out["wT"][:, -1], 1e-5, rtol=1e-5)     # imposed surface flux
    # causal penalty on the upper triangle of the first layer: the penalised weights shrink relative to plain training
    mask = np.triu(np.ones((128, 32), dtype=bool), k=1)
    pen = fc.masked_weight_penalty(NN, 0, mask)
    th0 = flux.destructure(NN)[0]
    v0, g0 = pen(th0)
    assert v0 > 0 and np.count_nonzero(g0) == mask.sum()
    datasets = {1: ds}
    hist = []
    lam = 1e3
    NNp = fc.train_neural_differential_equation(NN, fc.ConvectiveAdjustmentNDE, "Tsit5", datasets, T_scaling, wT_scaling, range(0, 19, 9),
                                                flux.ADAM(1e-3), 8, history=hist, ctx=ctx,
                                                causal_penalty=lambda th: tuple(lam * np.asarray(v) for v in pen(th)))
    NNu = fc.train_neural_differential_equation(NN, fc.ConvectiveAdjustmentNDE, "Tsit5", datasets, T_scaling, wT_scaling, range(0, 19, 9),
                                                flux.ADAM(1e-3), 8, ctx=ctx)
    vp, vu = pen(flux.destructure(NNp)[0])[0], pen(flux.destructure(NNu)[0])[0]
    print(f"causal penalty: initial {v0:.4e}, after 8 epochs with penalty {vp:.4e}, without {vu:.4e}; losses {hist[0]:.3e} -> {hist[-1]:.3e}")
    assert len(hist) == 8 and np.all(np.isfinite(hist)) and vp < vu and vp < v0
