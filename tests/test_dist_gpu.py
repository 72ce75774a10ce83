"""Two ranks through libcpz's allreduce hook as `parallel.attach_torch_allreduce` installs it: every rank
runs cpz_loss_grad / cpz_train_step on its shard of the columns on a context with its own non-default stream, the hook
sums the packed [grad; sums; count] buffer over the process group ON THE STREAM THE ENGINE PASSES, and the result must
equal the single-process full-batch result. Both ranks share cuda:0 (the test box has one GPU), so the process group is
gloo — NCCL refuses two ranks on one device; bench.py --gpus N exercises the same hook over NCCL."""
import os
import socket
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(_HERE), _HERE):
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

pytestmark = pytest.mark.gpu
NCOL = 75
W = np.array([0.7, 0.7, 1.0, 3e-3, 3e-3, 5e-3], dtype=np.float32)


def _problem():
    d = syn.wind_mixing_desc(variant=RHS_TRAIN, n_steps=6, save_stride=3, ckpt_stride=3)
    th = syn.theta_random(d, scale=0.3)
    x0, bcs = syn.columns(d, NCOL)
    tgt = np.ascontiguousarray(np.repeat(x0[:, None, :], d.n_saved, axis=1) * np.float32(0.9))
    return d, th, x0, bcs, tgt


def _worker(rank, world, port, out):
    from cpz_b200 import engine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    side = torch.cuda.Stream()          # the context's stream is NOT torch's current stream
    ctx = engine.Context(0, side.cuda_stream)
    parallel.attach_torch_allreduce(ctx)
    d, th, x0, bcs, tgt = _problem()
    lo, hi = parallel.shard_columns(NCOL, rank, world)
    m = engine.Model(ctx, d, th)
    # keep torch's current (default) stream busy so that a collective enqueued there would race the engine's stream
    junk = torch.randn(4096, 4096, device="cuda")
    for _ in range(3):
        junk = junk @ junk * 1e-3
    loss, grad = m.loss_grad(x0[lo:hi], bcs[lo:hi], tgt[lo:hi], W)
    l2 = m.train_step(x0[lo:hi], bcs[lo:hi], tgt[lo:hi], W, 1e-3)
    out[rank] = (loss, grad, l2, m.get_theta())
    m.close(); ctx.close()
    dist.destroy_process_group()


def test_two_ranks_through_the_library_hook_equal_one_rank(ctx):
    from cpz_b200 import engine
    d, th, x0, bcs, tgt = _problem()
    m = engine.Model(ctx, d, th)
    loss1, grad1 = m.loss_grad(x0, bcs, tgt, W)
    m.train_step(x0, bcs, tgt, W, 1e-3)
    theta1 = m.get_theta()
    m.close()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for r in range(2):
        loss, grad, l2, theta = out[r]
        assert np.linalg.norm(grad - grad1) <= 1e-5 * np.linalg.norm(grad1)
        np.testing.assert_allclose(loss, loss1, rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(l2, loss1, rtol=1e-5, atol=1e-12)
        # ADAM's first step moves every parameter by lr * g / (|g| + eps): identical to 1e-2 of lr (components of the gradient
        # at the FP32 noise level see the different summation order of one vs two shards)
        assert np.abs(theta - theta1).max() <= 1e-2 * 1e-3
    np.testing.assert_array_equal(out[0][1], out[1][1])  # both ranks hold the identical summed gradient
    np.testing.assert_array_equal(out[0][3], out[1][3])  # ... and apply the identical update


def test_hook_refuses_the_legacy_default_stream(ctx):
    """attach_torch_allreduce cannot order a collective with kernels on stream 0; the hook says so instead of racing."""
    from cpz_b200 import engine
    calls = []

    class FakeCtx:
        device = 0

        def set_allreduce(self, fn, rank, world):
            calls.append(fn)

    import torch.distributed as dist_
    if not dist_.is_initialized():
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        dist_.init_process_group("gloo", rank=0, world_size=1)
        made = True
    else:
        made = False
    try:
        parallel.attach_torch_allreduce(FakeCtx())
        with pytest.raises(RuntimeError):
            calls[0](0, 4, 0)
    finally:
        if made:
            dist_.destroy_process_group()
