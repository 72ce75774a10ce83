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
