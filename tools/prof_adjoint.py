"""Per-phase cycle counters of the adjoint kernel (CPZ_PROF=1), 32-column vs small tiles: python tools/prof_adjoint.py"""
import os, sys
os.environ["CPZ_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cpzload; cpzload.load()
from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import RHS_TRAIN

ctx = engine.Context(0)
w = np.array([1, 1, 1, 5e-3, 5e-3, 5e-3], dtype=np.float32)
for ncol, small in ((4736, "0"), (592, "100000"), (1, "100000")):
    os.environ["CPZ_SMALL_NCOL"] = small
    d = syn.wind_mixing_desc(variant=RHS_TRAIN, net="uvT_small", n_steps=90, save_stride=9, ckpt_stride=9)
    x, b = syn.columns(d, ncol, seed=1000)
    tg = np.ascontiguousarray(np.repeat(x[:, None, :], d.n_saved, axis=1))
    m = engine.Model(ctx, d, syn.theta_init(d, seed=42, scale=1e-5))
    print(f"--- ncol {ncol} small_ncol {small}", file=sys.stderr, flush=True)
    m.loss_grad(x, b, tg, w)
    m.close()
d = syn.free_convection_desc(ca=True, mpp=True, n_steps=36, save_stride=9, ckpt_stride=9)
x, b = syn.columns(d, 1)
tg = np.ascontiguousarray(np.repeat(x[:, None, :], d.n_saved, axis=1))
m = engine.Model(ctx, d, syn.theta_init(d, seed=42, scale=1e-5))
print("--- config 1 model, 1 column", file=sys.stderr, flush=True)
m.loss_grad(x, b, tg, np.array([0, 0, 1, 0, 0, 0], dtype=np.float32))
m.close()
