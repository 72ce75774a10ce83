"""Training-history files: mirror of wind_mixing/src/data_writing.jl:1-78.

The reference appends to a JLD2 (HDF5) file; this image has neither JLD2 nor h5py, so the same GROUP HIERARCHY and key
names ("training_info/...", "training_data/loss/total/<stage>/<count>", "training_data/neural_network/uw/<stage>/<count>",
"training_data/optimizer/state/<stage>/<count>", ...) are kept in a flat path -> array dictionary persisted as a numpy
.npz archive. Neural networks are stored as their destructured parameter vectors plus layer sizes / activations (what
`Flux.destructure` + the Chain's structure carry); the optimizer state is the (m, v, beta_pow) triple of
cpz_adam_get_state. Host-side I/O only: nothing here touches the GPU path."""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import numpy as np

from .flux import ADAM, Chain, destructure

LOSS_KEYS = ("u", "v", "T", "∂u∂z", "∂v∂z", "∂T∂z")


# One open history per path: the dictionary lives in memory between calls and is written back every `flush_every` appended
# iterations (and by flush()), so that a long run does not re-read and re-write the whole archive per iteration.
_OPEN: Dict[str, Dict[str, np.ndarray]] = {}
_DIRTY: Dict[str, int] = {}


def _load(path: str) -> Dict[str, np.ndarray]:
    path = os.path.abspath(path)
    if path not in _OPEN:
        with np.load(path, allow_pickle=False) as z:
            _OPEN[path] = {k: z[k] for k in z.files}
        _DIRTY[path] = 0
    return _OPEN[path]


def _store(path: str, d: Dict[str, np.ndarray], flush_every: int = 1) -> None:
    path = os.path.abspath(path)
    _OPEN[path] = d
    _DIRTY[path] = _DIRTY.get(path, 0) + 1
    if _DIRTY[path] >= flush_every:
        flush(path)


def flush(path: str) -> None:
    """Write the in-memory history of `path` to disk (atomic replace)."""
    path = os.path.abspath(path)
    if path in _OPEN:
        tmp = path + ".tmp.npz"
        np.savez(tmp, **_OPEN[path])
        os.replace(tmp, path)
        _DIRTY[path] = 0


def _put_chain(d: Dict[str, np.ndarray], key: str, nn: Chain) -> None:
    theta, _ = destructure(nn)
    d[key + "/theta"] = np.asarray(theta, dtype=np.float32)
    d[key + "/sizes"] = np.asarray(nn.sizes, dtype=np.int32)
    d[key + "/activations"] = np.asarray(nn.acts)


def write_metadata_NDE_training(FILE_PATH: str, train_files: Sequence[str], train_epochs, train_tranges, train_parameters: dict,
                                opts: Sequence[ADAM], uw_NN: Chain, vw_NN: Chain, wT_NN: Chain) -> None:
    """data_writing.jl:4-26 — creates the file with the training_info group (overwrites, as jldopen(..., "w") does)."""
    d: Dict[str, np.ndarray] = {}
    d["training_info/train_files"] = np.asarray(list(train_files))
    d["training_info/train_epochs"] = np.asarray(train_epochs)
    d["training_info/train_tranges"] = np.asarray(train_tranges)
    d["training_info/optimizers/eta"] = np.asarray([o.eta for o in opts], dtype=np.float64)
    d["training_info/optimizers/beta"] = np.asarray([o.beta for o in opts], dtype=np.float64)
    for k, v in train_parameters.items():
        d[f"training_info/parameters/{k}"] = np.asarray(v)
    _put_chain(d, "training_info/uw_neural_network", uw_NN)
    _put_chain(d, "training_info/vw_neural_network", vw_NN)
    _put_chain(d, "training_info/wT_neural_network", wT_NN)
    _OPEN.pop(os.path.abspath(FILE_PATH), None)
    _store(FILE_PATH, d)


def write_data_NDE_training(FILE_PATH: str, losses: Dict[str, float], loss_scalings: Dict[str, float], uw_NN: Chain, vw_NN: Chain,
                            wT_NN: Chain, stage, optimizer: ADAM, state: Optional[dict] = None, flush_every: int = 1) -> int:
    """data_writing.jl:28-78 — appends one iteration under .../<stage>/<count>; returns count (1-based, as in the reference).
    The archive on disk is rewritten every `flush_every` calls (1 = every call, the reference's behaviour) and by flush()."""
    d = _load(FILE_PATH)
    profile_loss = losses["u"] + losses["v"] + losses["T"]
    gradient_loss = losses["∂u∂z"] + losses["∂v∂z"] + losses["∂T∂z"]
    prefix = f"training_data/loss/total/{stage}/"
    count = sum(1 for k in d if k.startswith(prefix)) + 1
    if count == 1:
        for k in LOSS_KEYS:
            d[f"training_info/loss_scalings/{k}"] = np.float64(loss_scalings[k])
    d[f"training_data/loss/total/{stage}/{count}"] = np.float64(profile_loss + gradient_loss)
    d[f"training_data/loss/profile/{stage}/{count}"] = np.float64(profile_loss)
    d[f"training_data/loss/gradient/{stage}/{count}"] = np.float64(gradient_loss)
    for k in LOSS_KEYS:
        d[f"training_data/loss/{k}/{stage}/{count}"] = np.float64(losses[k])
    _put_chain(d, f"training_data/neural_network/uw/{stage}/{count}", uw_NN)
    _put_chain(d, f"training_data/neural_network/vw/{stage}/{count}", vw_NN)
    _put_chain(d, f"training_data/neural_network/wT/{stage}/{count}", wT_NN)
    d[f"training_data/optimizer/η/{stage}/{count}"] = np.float64(optimizer.eta)
    d[f"training_data/optimizer/β/{stage}/{count}"] = np.asarray(optimizer.beta, dtype=np.float64)
    if state:
        for k in ("m", "v", "beta_pow"):
            d[f"training_data/optimizer/state/{stage}/{count}/{k}"] = np.asarray(state[k], dtype=np.float32)
    _store(FILE_PATH, d, flush_every)
    return count


def read_training_history(FILE_PATH: str) -> Dict[str, np.ndarray]:
    """The whole file as a path -> array dictionary (what FileIO.load gives for a JLD2 history)."""
    flush(FILE_PATH)
    return dict(_load(FILE_PATH))


def loss_series(FILE_PATH: str, name: str = "total", stage=1) -> np.ndarray:
    """training_data/loss/<name>/<stage>/1..count as one array (the series the reference's plotting scripts read)."""
    d = _load(FILE_PATH)
    prefix = f"training_data/loss/{name}/{stage}/"
    n = sum(1 for k in d if k.startswith(prefix))
    return np.array([float(d[prefix + str(i)]) for i in range(1, n + 1)])
