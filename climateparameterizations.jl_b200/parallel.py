"""Column sharding across ranks (one process per GPU) and the single gradient allreduce.

Forward / inference needs no communication: every rank integrates its contiguous block of columns.
Training sums one packed buffer [grad (P); six squared-error sums; column count; pad; five mPP-parameter gradient sums; pad]
(P + 16 floats) over ranks — the engine calls the
hook between the adjoint kernel and the fused ADAM update, on its own stream.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_columns(ncol: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; sizes differ by at most one column."""
    base, rem = divmod(ncol, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_local(grad_unnormalised: np.ndarray, sq_sums: np.ndarray, ncol_local: int) -> np.ndarray:
    """Host-side statement of what the device packs before the allreduce (cpz_capi_train.inc: loss_grad_core)."""
    return np.concatenate([np.asarray(grad_unnormalised, dtype=np.float64), np.asarray(sq_sums, dtype=np.float64)[:6],
                           [float(ncol_local), 0.0], np.zeros(8)])  # [P+8, P+13): mPP-parameter gradient sums (cpz_loss_grad_mpp)


def finalize(packed_sum: np.ndarray, loss_w: np.ndarray, Nz: int, n_saved: int):
    """Host-side statement of scale_kernel + finalize_loss_kernel: (grad, [6 weighted losses, total])."""
    P = len(packed_sum) - 16
    ncol = packed_sum[P + 6]
    grad = packed_sum[:P] / ncol
    inv = np.array([1.0 / (Nz * n_saved)] * 3 + [1.0 / ((Nz + 1) * n_saved)] * 3)
    comps = np.asarray(loss_w, dtype=np.float64) * packed_sum[P:P + 6] * inv / ncol
    return grad, np.concatenate([comps, [comps.sum()]])


class _DevBuf:
    """Wraps a raw device pointer for torch via __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


def attach_torch_allreduce(ctx, group=None) -> None:
    """Installs a torch.distributed sum-allreduce (NCCL on GPUs) as the context's hook. The collective is enqueued on
    the cudaStream_t the engine passes — the stream its adjoint / reduction kernels and the following scale + ADAM
    launches run on — so it is ordered with them whatever torch's current stream is. The context must have been
    created on an explicit stream (`Context(device, torch.cuda.Stream().cuda_stream)` or the library-owned one);
    the legacy default stream (handle 0) is refused because torch cannot wrap it as an ExternalStream."""
    import torch
    import torch.distributed as dist

    def hook(ptr: int, n: int, stream: int) -> None:
        if not stream:
            raise RuntimeError("allreduce hook called with the legacy default stream; create the cpz context on an "
                               "explicit (non-default) CUDA stream")
        t = torch.as_tensor(_DevBuf(ptr, n), device=f"cuda:{ctx.device}")
        with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=f"cuda:{ctx.device}")):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    ctx.set_allreduce(hook, dist.get_rank(group), dist.get_world_size(group))
