"""world_size-2 gloo test of the data-parallel protocol: every rank packs [unnormalised grad; six
squared-error sums; column count], ONE sum-allreduce, identical finalisation on every rank — must equal the single-rank
full-batch loss and gradient. Runs the oracle on CPU (the device packs the same quantities)."""
import os
import socket
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(_HERE), _HERE):  # spawned workers re-import this module without conftest.py
    if _p not in sys.path:
        sys.path.insert(0, _p)
import cpzload  # noqa: E402

cpzload.load()

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cpz_b200 import parallel, synthetic as syn
from cpz_b200.desc import RHS_TRAIN
from oracle import nde
from util import oracle_loss_grad, oracle_solve, t64

NCOL = 11
W = np.array([0.7, 0.7, 1.0, 3e-3, 3e-3, 5e-3])


def _problem():
    d = syn.wind_mixing_desc(variant=RHS_TRAIN, n_steps=4, save_stride=2, ckpt_stride=2)
    th = syn.theta_random(d, scale=0.3)
    x0, bcs = syn.columns(d, NCOL)
    tgt = oracle_solve(d, (th * 1.2).astype(np.float32), x0, bcs)
    return d, th, x0, bcs, tgt


def _local_pack(d, th, x0, bcs, tgt):
    """what loss_grad_core leaves in b_red before the allreduce, computed with the oracle"""
    theta = t64(th).requires_grad_(True)
    traj = nde.solve(d, theta, t64(x0), t64(bcs), None)
    N, nt, n = d.Nz, traj.shape[1], traj.shape[0]
    sums = [c * (n * nt * (N if i < 3 else N + 1)) for i, c in enumerate(nde.loss_components(d, traj, t64(tgt)))]
    inv = [1.0 / (N * nt)] * 3 + [1.0 / ((N + 1) * nt)] * 3
    obj = sum(float(W[i]) * inv[i] * sums[i] for i in range(6))
    (g,) = torch.autograd.grad(obj, theta)
    return parallel.pack_local(g.numpy(), np.array([float(s) for s in sums]), n)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d, th, x0, bcs, tgt = _problem()
    lo, hi = parallel.shard_columns(NCOL, rank, world)
    buf = torch.tensor(_local_pack(d, th, x0[lo:hi], bcs[lo:hi], tgt[lo:hi]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    grad, loss = parallel.finalize(buf.numpy(), W, d.Nz, d.n_saved)
    out[rank] = (grad, loss)
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_rank():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    d, th, x0, bcs, tgt = _problem()
    tot, comps, g = oracle_loss_grad(d, th, x0, bcs, tgt, W)
    for r in range(2):
        grad, loss = out[r]
        assert np.linalg.norm(grad - g) / np.linalg.norm(g) < 1e-10
        np.testing.assert_allclose(loss[6], tot, rtol=1e-10)
        np.testing.assert_allclose(loss[:6], comps, rtol=1e-10, atol=1e-18)
    np.testing.assert_array_equal(out[0][0], out[1][0])  # every rank applies the identical update
