"""Short T-only free-convection solve for ncu: python tools/profile_fc.py [--ncol 18944] [--steps 8]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cpzload; cpzload.load()
from cpz_b200 import engine, synthetic as syn
ap = argparse.ArgumentParser()
ap.add_argument("--ncol", type=int, default=18944)
ap.add_argument("--steps", type=int, default=8)
a = ap.parse_args()
ctx = engine.Context(0)
d = syn.free_convection_desc(ca=False, n_steps=a.steps, save_stride=max(1, a.steps // 2))
m = engine.Model(ctx, d, syn.theta_init(d))
x0, bcs = syn.columns(d, a.ncol)
x0d, bcsd = torch.tensor(x0, device="cuda"), torch.tensor(bcs, device="cuda")
traj = torch.empty((a.ncol, d.n_saved, d.S), device="cuda")
for _ in range(3):
    m.solve_dev(x0d, bcsd, traj)
ctx.synchronize()
print("done")
