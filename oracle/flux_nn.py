"""Flux 0.11.6 Dense / Chain / destructure semantics, restated (oracle: test infrastructure only).

Flux is a third-party dependency of the reference (wind_mixing/Manifest.toml:523), not vendored under
/root/reference; its published behaviour is restated here and anchored on the reference's call sites:
    Flux.destructure(NN) -> (theta, re)      wind_mixing/src/NDE_training.jl:11-13
    re(theta) rebuilt every RHS call          wind_mixing/src/NDE_training.jl:62-64
    NN(x) with x the full scaled state        wind_mixing/src/NDE_training.jl:94-96
    Chain(Dense(96,50,mish),Dense(50,20,mish),Dense(20,31))   wind_mixing/train_NDE.jl:103
    Chain(Dense(Nz,4Nz,relu),Dense(4Nz,4Nz,relu),Dense(4Nz,Nz-1))   free_convection/train_free_convection_nde.jl:119-121

Dense(in,out,σ): y = σ.(W*x .+ b), W is out×in.  destructure concatenates, layer by layer, vec(W) (Julia
column-major: index = o + i*out) then b.  Activations follow NNlib 0.7.20 (wind_mixing/Manifest.toml:1222).
"""
import numpy as np
import torch


def act_torch(name: str, x: torch.Tensor) -> torch.Tensor:
    if name == "identity":
        return x
    if name == "relu":
        return torch.clamp_min(x, 0.0)
    if name == "mish":  # x * tanh(softplus(x))
        return x * torch.tanh(torch.nn.functional.softplus(x, beta=1.0, threshold=1e9))
    if name == "swish":  # x * sigmoid(x)
        return x * torch.sigmoid(x)
    if name == "leakyrelu":  # max(0.01x, x)
        return torch.maximum(0.01 * x, x)
    if name == "tanh":
        return torch.tanh(x)
    raise ValueError(name)


def act_numpy(name: str, x: np.ndarray) -> np.ndarray:
    if name == "identity":
        return x
    if name == "relu":
        return np.maximum(x, 0)
    if name == "mish":
        return x * np.tanh(np.logaddexp(0, x))
    if name == "swish":
        return x / (1 + np.exp(-x))
    if name == "leakyrelu":
        return np.maximum(0.01 * x, x)
    if name == "tanh":
        return np.tanh(x)
    raise ValueError(name)


def n_params(sizes) -> int:
    return sum(sizes[i] * sizes[i + 1] + sizes[i + 1] for i in range(len(sizes) - 1))


def glorot_theta(sizes, rng: np.random.Generator, scale: float = 1.0) -> np.ndarray:
    """Flux-default initial parameters in destructure order: W ~ U(±sqrt(6/(in+out))), b = 0, then * scale
    (the reference divides by 1e5 before NDE training: wind_mixing/train_NDE.jl:105-107)."""
    parts = []
    for i in range(len(sizes) - 1):
        fan_in, fan_out = sizes[i], sizes[i + 1]
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        W = rng.uniform(-lim, lim, size=(fan_out, fan_in)).astype(np.float32)
        parts.append(W.flatten(order="F"))  # column-major vec(W)
        parts.append(np.zeros(fan_out, dtype=np.float32))
    return (np.concatenate(parts) * np.float32(scale)).astype(np.float32)


def reconstruct(theta, sizes):
    """`re(theta)`: list of (W[out,in], b[out]) views in destructure order. Works for numpy arrays and torch tensors."""
    layers, off = [], 0
    for i in range(len(sizes) - 1):
        fan_in, fan_out = sizes[i], sizes[i + 1]
        w = theta[off:off + fan_in * fan_out]
        off += fan_in * fan_out
        b = theta[off:off + fan_out]
        off += fan_out
        # column-major (out x in)  ==  row-major (in x out) transposed
        W = w.reshape(fan_in, fan_out).T if isinstance(w, np.ndarray) else w.reshape(fan_in, fan_out).t()
        layers.append((W, b))
    assert off == len(theta)
    return layers


def destructure(layers) -> np.ndarray:
    parts = []
    for W, b in layers:
        parts.append(np.asarray(W).flatten(order="F"))
        parts.append(np.asarray(b))
    return np.concatenate(parts)


def chain_numpy(theta: np.ndarray, sizes, acts, x: np.ndarray) -> np.ndarray:
    """NN(x) for ONE input vector x (numpy), literal form σ.(W*x .+ b)."""
    h = x
    for (W, b), a in zip(reconstruct(theta, sizes), acts):
        h = act_numpy(a, W @ h + b)
    return h


def chain_torch(theta: torch.Tensor, sizes, acts, x: torch.Tensor) -> torch.Tensor:
    """NN over a batch: x [ncol, in] -> [ncol, out]."""
    h = x
    for (W, b), a in zip(reconstruct(theta, sizes), acts):
        h = act_torch(a, h @ W.t() + b)
    return h
