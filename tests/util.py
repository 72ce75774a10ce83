"""Shared helpers for the parity tests (oracle = checker only)."""
import numpy as np
import torch

from oracle import nde


def t64(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64)


def t32(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32)


def rel_inf(a, ref):
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300))


def oracle_rhs(desc, theta, x, bcs, t=0.0, Q=None, dtype=torch.float64):
    conv = (lambda a: torch.tensor(np.asarray(a), dtype=dtype))
    q = None if Q is None else conv(Q)
    return nde.rhs(desc, conv(theta), conv(x), conv(bcs), t, q).numpy()


def oracle_solve(desc, theta, x0, bcs, Q=None, dtype=torch.float64):
    conv = (lambda a: torch.tensor(np.asarray(a), dtype=dtype))
    q = None if Q is None else conv(Q)
    with torch.no_grad():
        return nde.solve(desc, conv(theta), conv(x0), conv(bcs), q).numpy()


def oracle_loss_grad(desc, theta, x0, bcs, targets, w, Q=None, dtype=torch.float64):
    conv = (lambda a: torch.tensor(np.asarray(a), dtype=dtype))
    q = None if Q is None else conv(Q)
    total, scaled, g = nde.loss_grad(desc, conv(theta), conv(x0), conv(bcs), q, conv(targets), w)
    return float(total), np.array([float(s) for s in scaled]), g.numpy()
