"""Generates tests/golden/*.npz: seeded inputs and FP64 outputs of the LITERAL (line-by-line) restatement of the
reference RHS, plus FP64 solve / loss / gradient outputs of the batched oracle.

The reference itself (Julia) cannot run in this image, so these vectors pin the oracle and the CUDA engine to each other
and to this commit's arithmetic — they are NOT reference outputs (parity unpinned, see oracle/__init__.py).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cpzload  # noqa: E402

cpzload.load()
from cpz_b200 import synthetic as syn  # noqa: E402
from cpz_b200.desc import FLAG_CA, FLAG_MPP, FLAG_ZERO_WEIGHTS, RHS_INFER, RHS_TRAIN  # noqa: E402
from oracle import literal  # noqa: E402
from util import oracle_loss_grad, oracle_solve  # noqa: E402

CASES = {
    "train_mpp_zw": lambda: syn.wind_mixing_desc(variant=RHS_TRAIN, flags=FLAG_MPP | FLAG_ZERO_WEIGHTS, n_steps=6, save_stride=3, ckpt_stride=3),
    "infer_mpp_ca": lambda: syn.wind_mixing_desc(variant=RHS_INFER, flags=FLAG_MPP | FLAG_CA, n_steps=6, save_stride=3, ckpt_stride=3),
    "free_convection_ca": lambda: syn.free_convection_desc(ca=True, n_steps=6, save_stride=3, ckpt_stride=3),
}
W = {3: np.array([0.7, 0.7, 1.0, 3e-3, 3e-3, 5e-3], dtype=np.float32), 1: np.array([0, 0, 1.0, 0, 0, 0], dtype=np.float32)}


def build(name):
    d = CASES[name]()
    th = syn.theta_random(d, seed=11, scale=0.3)
    x0, bcs = syn.columns(d, 6, seed=77)
    rhs = np.stack([literal.rhs(d, th, x0[i], bcs[i], 0.25) for i in range(6)])
    traj = oracle_solve(d, th, x0, bcs)
    # targets: a perturbed-theta solution plus seeded noise, so that (prediction - target) is well above FP32 rounding
    tgt = oracle_solve(d, (th * 1.25).astype(np.float32), x0, bcs)
    tgt = (tgt + 0.02 * np.random.default_rng(5).standard_normal(tgt.shape)).astype(np.float32)
    tot, comps, g = oracle_loss_grad(d, th, x0, bcs, tgt, W[d.n_fields])
    return dict(theta=th, x0=x0, bcs=bcs, t_rhs=0.25, rhs=rhs, traj=traj, targets=tgt, loss_w=W[d.n_fields].astype(np.float32),
                loss=np.concatenate([comps, [tot]]), grad=g)


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **build(name))
        print("wrote", name)
