"""S7 parity: per-step gyre closure (implicit convective adjustment + NN forcing), CUDA vs FP64 oracle."""
import numpy as np
import pytest
import torch

from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import ClosureDesc
from oracle import literal, nde
from util import rel_inf, t64

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nx,ny", [(64, 8), (50, 7), (3, 1)])
def test_closure_step(ctx, nx, ny):
    d = syn.free_convection_desc(ca=False)
    th = syn.theta_random(d, scale=1.0)
    T, y = syn.gyre_field(nx, ny, 32)
    # make a third of the columns statically unstable so the adjustment acts
    T[:, :, ::3] = T[::-1, :, ::3] * 1.0
    cd = ClosureDesc(Nx=nx, Ny=ny, Nz=32)
    m = engine.Model(ctx, d, th)
    forcing, T_out = m.closure_step(cd, T, y)
    m.close()
    f_ref, T_ref = nde.closure_step(d, t64(th), cd, t64(T), t64(y))
    e_f, e_T = rel_inf(forcing, f_ref.numpy()), rel_inf(T_out, T_ref.numpy())
    print(f"closure {nx}x{ny}: forcing {e_f:.2e}  T_out {e_T:.2e}")
    assert e_f <= 1e-5 and e_T <= 1e-5
    assert np.abs(T_out - T).max() > 1e-3  # the adjustment did something
    # one column against the line-by-line restatement
    col = (0, min(3, nx - 1))
    T_adj = literal.convective_adjustment_implicit(T[:, col[0], col[1]], cd.dt, cd.dz, cd.K)
    f_col = literal.gyre_closure_column(d, th, cd, T_adj, float(y[col[0]]))
    assert rel_inf(T_out[:, col[0], col[1]], T_adj) <= 1e-5
    assert np.abs(forcing[:, col[0], col[1]] - f_col).max() / np.abs(f_ref.numpy()).max() <= 1e-5


@pytest.mark.parametrize("h1,h2,act", [(128, 128, "relu"), (128, 128, "mish"), (64, 48, "tanh"), (20, 100, "swish"), (8, 8, "relu")])
def test_closure_step_tcgen05_net_shapes_and_simt_agreement(ctx, h1, h2, act):
    """The closure runs on the tcgen05 kernel (columns on the M side, activations in TMEM) for T-only nets up to
    128-wide; it must agree with the FP64 oracle to 1e-5 and with the FP32 SIMT kernel (CPZ_NO_TC=1)."""
    import os
    from cpz_b200.desc import NetDesc
    d = syn.free_convection_desc(ca=False, net=None)
    d.nets = [NetDesc([32, h1, h2, 31], [act, act, "identity"])]
    th = syn.theta_random(d, scale=1.0)
    nx, ny = 200, 3   # 600 columns: four full 128-column tiles and a ragged one
    T, y = syn.gyre_field(nx, ny, 32)
    T[:, :, ::3] = T[::-1, :, ::3] * 1.0
    cd = ClosureDesc(Nx=nx, Ny=ny, Nz=32)
    m = engine.Model(ctx, d, th)
    forcing, T_out = m.closure_step(cd, T, y)
    os.environ["CPZ_NO_TC"] = "1"
    try:
        f_simt, T_simt = m.closure_step(cd, T, y)
    finally:
        os.environ.pop("CPZ_NO_TC", None)
    m.close()
    f_ref, T_ref = nde.closure_step(d, t64(th), cd, t64(T), t64(y))
    e_f, e_T, e_s = rel_inf(forcing, f_ref.numpy()), rel_inf(T_out, T_ref.numpy()), rel_inf(f_simt, f_ref.numpy())
    print(f"closure tc {h1}x{h2} {act}: forcing {e_f:.2e} (simt {e_s:.2e})  T_out {e_T:.2e}")
    assert e_f <= 1e-5 and e_T <= 1e-5
    assert rel_inf(forcing, f_simt) <= 2e-5


def test_closure_weight_image_follows_theta_updates(ctx):
    """The tcgen05 closure caches its shared-memory weight image; cpz_set_theta must invalidate it."""
    d = syn.free_convection_desc(ca=False)
    cd = ClosureDesc(Nx=64, Ny=4, Nz=32)
    T, y = syn.gyre_field(64, 4, 32)
    th1, th2 = syn.theta_random(d, scale=1.0, seed=1), syn.theta_random(d, scale=1.0, seed=2)
    m = engine.Model(ctx, d, th1)
    f1, _ = m.closure_step(cd, T, y)
    m.set_theta(th2)
    f2, _ = m.closure_step(cd, T, y)
    m.close()
    r2, _ = nde.closure_step(d, t64(th2), cd, t64(T), t64(y))
    assert rel_inf(f2, r2.numpy()) <= 1e-5
    assert rel_inf(f1, r2.numpy()) > 1e-3
