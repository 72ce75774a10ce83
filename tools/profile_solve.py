"""Short forward solve for ncu: python tools/profile_solve.py [--ncol 4096] [--steps 4] [--net uvT_small] [--mode solve|grad]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cpzload; cpzload.load()
from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import RHS_INFER, RHS_TRAIN

ap = argparse.ArgumentParser()
ap.add_argument("--ncol", type=int, default=4096)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--net", default="uvT_small")
ap.add_argument("--mode", default="solve")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
ctx = engine.Context(0)
if a.mode == "solve":
    d = syn.wind_mixing_desc(variant=RHS_INFER, net=a.net, n_steps=a.steps, save_stride=1)
else:
    d = syn.wind_mixing_desc(variant=RHS_TRAIN, net=a.net, n_steps=a.steps, save_stride=max(1, a.steps // 2), ckpt_stride=max(1, a.steps // 2))
th = syn.theta_init(d)
m = engine.Model(ctx, d, th)
x0, bcs = syn.columns(d, a.ncol)
x0d, bcsd = torch.tensor(x0, device="cuda"), torch.tensor(bcs, device="cuda")
if a.mode == "solve":
    traj = torch.empty((a.ncol, d.n_saved, d.S), device="cuda")
    for _ in range(a.reps):
        m.solve_dev(x0d, bcsd, traj)
else:
    tgt = torch.zeros((a.ncol, d.n_saved, d.S), device="cuda")
    loss = torch.zeros(8, device="cuda"); grad = torch.zeros(m.P, device="cuda")
    w = np.array([1, 1, 1, 5e-3, 5e-3, 5e-3], dtype=np.float32)
    for _ in range(a.reps):
        m.loss_grad_dev(x0d, bcsd, tgt, w, loss, grad)
ctx.synchronize()
torch.cuda.synchronize()
print("done")
