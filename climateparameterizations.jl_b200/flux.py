"""Host-side mirror of the Flux pieces the reference's NDE path touches: Dense, Chain, destructure / re, activations'
names, ADAM's hyper-parameter record. No arithmetic happens here — networks are only containers whose parameters are
handed to the engine in `Flux.destructure` order (wind_mixing/src/NDE_training.jl:11-13,37).

    NN = Chain(Dense(96, 50, "mish"), Dense(50, 20, "mish"), Dense(20, 31))      # wind_mixing/train_NDE.jl:103
    theta, re = destructure(NN);  NN2 = re(theta)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np

from .desc import ACT, NetDesc


class Dense:
    """Dense(in, out, σ): W is out×in (Glorot-uniform by default), b zeros — Flux 0.11 defaults."""

    def __init__(self, n_in: int, n_out: int, act: str = "identity", W: Optional[np.ndarray] = None,
                 b: Optional[np.ndarray] = None, rng: Optional[np.random.Generator] = None):
        assert act in ACT, act
        self.n_in, self.n_out, self.act = n_in, n_out, act
        if W is None:
            rng = rng or np.random.default_rng()
            lim = np.sqrt(6.0 / (n_in + n_out))
            W = rng.uniform(-lim, lim, size=(n_out, n_in))
        self.W = np.asarray(W, dtype=np.float32).reshape(n_out, n_in)
        self.b = np.zeros(n_out, dtype=np.float32) if b is None else np.asarray(b, dtype=np.float32).reshape(n_out)


class Chain:
    def __init__(self, *layers: Dense):
        assert layers
        for a, b in zip(layers[:-1], layers[1:]):
            assert a.n_out == b.n_in, "layer sizes do not chain"
        self.layers: List[Dense] = list(layers)

    @property
    def sizes(self) -> List[int]:
        return [self.layers[0].n_in] + [l.n_out for l in self.layers]

    @property
    def acts(self) -> List[str]:
        return [l.act for l in self.layers]

    def net_desc(self) -> NetDesc:
        return NetDesc(self.sizes, self.acts)

    def scale(self, s: float) -> "Chain":
        """re(weights ./ 1f5) idiom of wind_mixing/train_NDE.jl:105-107"""
        theta, re = destructure(self)
        return re(theta * np.float32(s))


def destructure(nn: Chain) -> Tuple[np.ndarray, Callable[[np.ndarray], Chain]]:
    """(theta, re): per layer vec(W) column-major (out×in) then b."""
    parts = []
    for l in nn.layers:
        parts.append(l.W.flatten(order="F"))
        parts.append(l.b)
    theta = np.concatenate(parts).astype(np.float32)
    shapes = [(l.n_in, l.n_out, l.act) for l in nn.layers]

    def re(th: np.ndarray) -> Chain:
        th = np.asarray(th, dtype=np.float32)
        assert th.shape == theta.shape
        out, off = [], 0
        for n_in, n_out, act in shapes:
            W = th[off:off + n_in * n_out].reshape(n_in, n_out).T
            off += n_in * n_out
            b = th[off:off + n_out]
            off += n_out
            out.append(Dense(n_in, n_out, act, W=W.copy(), b=b.copy()))
        return Chain(*out)

    return theta, re


@dataclass
class ADAM:
    """Flux.ADAM(η, (β1, β2)) — the hyper-parameters; the update itself runs on the device (cpz_train_step)."""
    eta: float = 1e-3
    beta: Tuple[float, float] = (0.9, 0.999)
    eps: float = 1e-8
    state: dict = field(default_factory=dict)  # {'m','v','beta_pow'} mirrored from the engine for checkpointing
