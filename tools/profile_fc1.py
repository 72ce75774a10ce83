"""Single-column T-only training pass (BASELINE config 1 model) for ncu: python tools/profile_fc1.py [--ncol 1] [--steps 36]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cpzload; cpzload.load()
from cpz_b200 import engine, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--ncol", type=int, default=1)
ap.add_argument("--steps", type=int, default=36)
a = ap.parse_args()
ctx = engine.Context(0)
d = syn.free_convection_desc(ca=True, mpp=True, n_steps=a.steps, save_stride=9, ckpt_stride=9)
m = engine.Model(ctx, d, syn.theta_init(d, seed=42, scale=1e-5))
x, b = syn.columns(d, a.ncol)
xd, bd = torch.tensor(x, device="cuda"), torch.tensor(b, device="cuda")
tg = torch.tensor(x, device="cuda")[:, None, :].repeat(1, d.n_saved, 1).contiguous() + 0.05
w = np.array([0, 0, 1, 0, 0, 0], dtype=np.float32)
loss = torch.zeros(8, device="cuda"); grad = torch.zeros(m.P, device="cuda")
for _ in range(2):
    m.loss_grad_dev(xd, bd, tg, w, loss, grad)
torch.cuda.synchronize()
print("done", d.n_substeps, float(loss[6]))
