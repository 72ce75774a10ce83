"""Committed golden vectors (tests/golden/make_golden.py): the oracle must reproduce them on CPU, the CUDA engine on GPU."""
import os

import numpy as np
import pytest

from util import oracle_loss_grad, oracle_rhs, oracle_solve, rel_inf

HERE = os.path.dirname(os.path.abspath(__file__))
import sys
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

NAMES = sorted(make_golden.CASES)


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    g, d = _load(name), make_golden.CASES[name]()
    assert rel_inf(oracle_rhs(d, g["theta"], g["x0"], g["bcs"], float(g["t_rhs"])), g["rhs"]) < 1e-12
    assert rel_inf(oracle_solve(d, g["theta"], g["x0"], g["bcs"]), g["traj"]) < 1e-11
    tot, comps, grad = oracle_loss_grad(d, g["theta"], g["x0"], g["bcs"], g["targets"], g["loss_w"])
    assert abs(tot - g["loss"][6]) / abs(g["loss"][6]) < 1e-10
    assert np.linalg.norm(grad - g["grad"]) / np.linalg.norm(g["grad"]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_reproduces_golden(ctx, name):
    from cpz_b200 import engine
    g, d = _load(name), make_golden.CASES[name]()
    m = engine.Model(ctx, d, g["theta"])
    e_rhs = rel_inf(m.rhs(g["x0"], g["bcs"], t=float(g["t_rhs"])), g["rhs"])
    e_traj = rel_inf(m.solve(g["x0"], g["bcs"]), g["traj"])
    loss, grad = m.loss_grad(g["x0"], g["bcs"], g["targets"], g["loss_w"])
    m.close()
    e_loss = abs(loss[6] - g["loss"][6]) / abs(g["loss"][6])
    e_grad = np.linalg.norm(grad - g["grad"]) / np.linalg.norm(g["grad"])
    # gradient: 1e-4, or the FP32 oracle's own distance from the FP64 golden where the RHS is non-smooth (see test_adjoint_gpu)
    import torch
    g32 = oracle_loss_grad(d, g["theta"], g["x0"], g["bcs"], g["targets"], g["loss_w"], dtype=torch.float32)[2]
    floor = np.linalg.norm(g32 - g["grad"]) / np.linalg.norm(g["grad"])
    print(f"golden {name}: rhs {e_rhs:.2e} traj {e_traj:.2e} loss {e_loss:.2e} grad {e_grad:.2e} (fp32-oracle grad {floor:.2e})")
    assert e_rhs <= 1e-5 and e_traj <= 1e-4 and e_loss <= 1e-4 and e_grad <= max(1e-4, floor)
