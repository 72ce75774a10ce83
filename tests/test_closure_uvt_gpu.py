"""S7b parity: per-step closure of the u/v/T NDE embedded in a host ocean model — NN forcing chains + backward-Euler
modified Pacanowski–Philander step (wind_mixing/src/NDE_oceananigans.jl:17-101,288-344,380-405), CUDA vs the FP64 oracle."""
import numpy as np
import pytest

from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import ClosureUvtDesc, RHS_INFER, RHS_TRAIN
from oracle import literal, nde
from util import rel_inf, t64

pytestmark = pytest.mark.gpu


def _run(ctx, d, th, cd, u, v, T):
    m = engine.Model(ctx, d, th)
    dzf, out = m.closure_step_uvt(cd, u, v, T)
    m.close()
    return dzf, out


@pytest.mark.parametrize("kernel", ["tcgen05", "fp32-simt"])
@pytest.mark.parametrize("nx,ny,ca", [(64, 8, False), (64, 8, True), (50, 7, True), (3, 1, True), (1, 1, False)])
def test_closure_step_uvt(ctx, nx, ny, ca, kernel, monkeypatch):
    """Both kernels: the CLOSURE instantiation of the tcgen05 solve kernel (default for the production nets) and the FP32 SIMT
    tile kernel (CPZ_NO_TC=1; also what other net shapes run on)."""
    if kernel == "fp32-simt":
        monkeypatch.setenv("CPZ_NO_TC", "1")
    d = syn.wind_mixing_desc(variant=RHS_INFER)
    th = syn.theta_random(d, scale=0.5)
    u, v, T = syn.uvt_fields(d, nx, ny, unstable_every=3 if ca else 0)
    cd = ClosureUvtDesc(Nx=nx, Ny=ny, Nz=32, dz=d.H / 32, dt=60.0, uw_top=-1e-4, vw_top=2e-5, wT_top=3e-5, convective_adjustment=ca)
    dzf, out = _run(ctx, d, th, cd, u, v, T)
    dzf_ref, out_ref = nde.closure_step_uvt(d, t64(th), cd, t64(u), t64(v), t64(T))
    dzf_ref, out_ref = dzf_ref.numpy(), out_ref.numpy()
    # the implicit step moves the state by a small increment: compare the increments too, relative to the largest one
    inc, inc_ref = out.astype(np.float64) - np.stack([u, v, T]), out_ref - np.stack([u, v, T]).astype(np.float64)
    e = [rel_inf(dzf[q], dzf_ref[q]) for q in range(3)] + [rel_inf(out[q], out_ref[q]) for q in range(3)]
    e_inc = [float(np.abs(inc[q] - inc_ref[q]).max() / max(np.abs(inc_ref[q]).max(), 1e-30)) for q in range(3)]
    print(f"uvT closure [{kernel}] {nx}x{ny} ca={ca}: dz_flux {e[0]:.1e} {e[1]:.1e} {e[2]:.1e}  state {e[3]:.1e} {e[4]:.1e} {e[5]:.1e}  "
          f"increments {e_inc[0]:.1e} {e_inc[1]:.1e} {e_inc[2]:.1e} (largest |dT| {np.abs(inc_ref[2]).max():.2e})")
    assert max(e) <= 1e-5
    # increments are differences of O(1e-7)-accurate FP32 states: bounded by the state's rounding plus a relative part — 1e-5 on
    # the FP32 kernel (IEEE-accurate tanh), 1e-4 (the profile tolerance) on the tcgen05 kernel, whose diffusivities use the
    # MUFU.EX2 / MUFU.RCP forms of the solve kernels
    rel = 1e-5 if kernel == "fp32-simt" else 1e-4
    for q in range(3):
        assert np.abs(inc[q] - inc_ref[q]).max() <= 4e-7 * np.abs(np.stack([u, v, T])[q]).max() + rel * np.abs(inc_ref[q]).max()
    assert np.all(out[2][0] == T[0])  # T'[bottom] = T_bottom exactly
    if ca:
        assert np.abs(inc_ref[2]).max() > 1e-3  # the kappa_ca branch acted
    # one column against the line-by-line restatement (dense tridiagonal solve)
    j, i = 0, min(3, nx - 1)
    f_col = literal.uvt_forcing_column(d, th, cd, u[:, j, i], v[:, j, i], T[:, j, i])
    s_col = literal.modified_pacanowski_philander_step(d, cd, u[:, j, i], v[:, j, i], T[:, j, i])
    for q in range(3):
        assert np.abs(dzf[q][:, j, i] - f_col[q]).max() <= 1e-5 * np.abs(dzf_ref[q]).max()
        assert rel_inf(out[q][:, j, i], s_col[q]) <= 1e-5


def test_closure_step_uvt_large_nets_stream_weights(ctx):
    """400-wide nets do not fit shared memory: the kernel streams the weights from L2 (WS = false instantiation)."""
    d = syn.wind_mixing_desc(variant=RHS_INFER, net="uvT_large")
    th = syn.theta_random(d, scale=0.3)
    u, v, T = syn.uvt_fields(d, 40, 2)
    cd = ClosureUvtDesc(Nx=40, Ny=2, Nz=32, dz=d.H / 32, dt=60.0, uw_top=-1e-4, wT_top=1e-5)
    dzf, out = _run(ctx, d, th, cd, u, v, T)
    dzf_ref, out_ref = nde.closure_step_uvt(d, t64(th), cd, t64(u), t64(v), t64(T))
    assert rel_inf(dzf, dzf_ref.numpy()) <= 1e-5 and rel_inf(out, out_ref.numpy()) <= 1e-5


def test_closure_step_uvt_rejects_wrong_models(ctx):
    d = syn.free_convection_desc(ca=False)
    m = engine.Model(ctx, d, syn.theta_random(d))
    cd = ClosureUvtDesc(Nx=4, Ny=1)
    z = np.zeros((32, 1, 4), dtype=np.float32)
    with pytest.raises(engine.CpzError) as ei:
        m.closure_step_uvt(cd, z, z, z)
    assert ei.value.code == engine.ERR_INVALID
    m.close()


def test_oceananigans_mirror_runs_the_callback_sequence(ctx):
    """wind_mixing.oceananigans_modified_pacanowski_philander_nn: every transition frame k -> k+1 of the embedded run (callback,
    then the stand-in explicit update) equals the same step on the oracle from the same frame. (A free-running FP64 chain is
    not comparable: in the weakly stratified mixed layer the Richardson number amplifies a 1e-7 rounding of T ~ 20 into an
    O(1e-3) change of the diffusivity at the next step.)"""
    from cpz_b200 import wind_mixing as wm
    d = syn.wind_mixing_desc(variant=RHS_INFER)
    th = syn.theta_random(d, scale=0.3)
    u, v, T = syn.uvt_fields(d, 6, 1)
    cd = ClosureUvtDesc(Nx=6, Ny=1, Nz=32, dz=d.H / 32, dt=60.0, uw_top=-1e-4, wT_top=2e-5, convective_adjustment=True)
    flux = lambda t: 2e-5 * (1.0 + t / 600.0)  # a time-dependent top temperature flux (NDE_oceananigans.jl:129,332)
    m = engine.Model(ctx, d, th)
    frames = wm.oceananigans_modified_pacanowski_philander_nn(m, cd, u, v, T, n_iterations=12, output_every=1, f=d.f, wT_flux=flux)
    m.close()
    assert len(frames) == 13
    worst = 0.0
    for it in range(12):
        cd.wT_top = flux(it * cd.dt)
        uu, vv, TT = (t64(a) for a in frames[it])
        dzf, out = nde.closure_step_uvt(d, t64(th), cd, uu, vv, TT)
        ref = wm._standin_dynamics(out[0], out[1], out[2], dzf, cd.dt, d.f)
        worst = max([worst] + [rel_inf(frames[it + 1][q], ref[q].numpy()) for q in range(3)])
    print(f"embedded u/v/T run, 12 transitions: worst {worst:.1e}")
    assert worst <= 1e-5
    assert np.abs(frames[-1][2] - frames[0][2]).max() > 1e-4


def test_closure_uvt_weight_image_follows_theta_updates(ctx):
    """The kernel copies a cached device image of the shared-memory weight arena; cpz_set_theta must invalidate it."""
    d = syn.wind_mixing_desc(variant=RHS_INFER)
    u, v, T = syn.uvt_fields(d, 40, 2)
    cd = ClosureUvtDesc(Nx=40, Ny=2, Nz=32, dz=d.H / 32, dt=60.0, uw_top=-1e-4, wT_top=1e-5)
    th1, th2 = syn.theta_random(d, scale=0.5, seed=1), syn.theta_random(d, scale=0.5, seed=2)
    m = engine.Model(ctx, d, th1)
    f1, _ = m.closure_step_uvt(cd, u, v, T)
    m.set_theta(th2)
    f2, _ = m.closure_step_uvt(cd, u, v, T)
    m.close()
    r2, _ = nde.closure_step_uvt(d, t64(th2), cd, t64(u), t64(v), t64(T))
    assert rel_inf(f2, r2.numpy()) <= 1e-5
    assert rel_inf(f1, r2.numpy()) > 1e-3


@pytest.mark.parametrize("kernel", ["tcgen05", "fp32-simt"])
def test_closure_step_uvt_grid_spacing_and_time_step_differ_from_the_training_setup(ctx, kernel, monkeypatch):
    """The host model's dz and dt are its own (NDE_oceananigans.jl:104,117), not H/Nz and the training step: the tcgen05 kernel
    works in the model's scaled variables and must rebuild B_z and the non-dimensional sub-step from them."""
    if kernel == "fp32-simt":
        monkeypatch.setenv("CPZ_NO_TC", "1")
    d = syn.wind_mixing_desc(variant=RHS_TRAIN)  # a training-variant description (eps on the gradients) must not leak into the closure
    th = syn.theta_random(d, scale=0.5)
    u, v, T = syn.uvt_fields(d, 40, 3, unstable_every=4)
    cd = ClosureUvtDesc(Nx=40, Ny=3, Nz=32, dz=5.0, dt=600.0, uw_top=-2e-4, vw_top=1e-5, wT_top=-4e-5, convective_adjustment=True, kappa_ca=0.3)
    dzf, out = _run(ctx, d, th, cd, u, v, T)
    dzf_ref, out_ref = nde.closure_step_uvt(d, t64(th), cd, t64(u), t64(v), t64(T))
    e_f, e_s = rel_inf(dzf, dzf_ref.numpy()), rel_inf(out, out_ref.numpy())
    inc_ref = out_ref.numpy() - np.stack([u, v, T]).astype(np.float64)
    e_i = float(np.abs(out - np.stack([u, v, T]) - inc_ref).max() / np.abs(inc_ref).max())
    print(f"uvT closure [{kernel}] dz=5 dt=600 kappa_ca=0.3: dz_flux {e_f:.1e} state {e_s:.1e} increments {e_i:.1e}")
    assert e_f <= 1e-5 and e_s <= 1e-5 and e_i <= 2e-4
