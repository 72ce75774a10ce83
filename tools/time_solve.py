"""Wall-clock (CUDA-event) timing of the config-2 forward solve: python tools/time_solve.py [--ncol 4096] [--steps 1152] [--reps 3]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cpzload; cpzload.load()
from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import RHS_INFER
ap = argparse.ArgumentParser()
ap.add_argument("--ncol", type=int, default=4096)
ap.add_argument("--steps", type=int, default=1152)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--save", type=int, default=1)
a = ap.parse_args()
ctx = engine.Context(0)
d = syn.wind_mixing_desc(variant=RHS_INFER, n_steps=a.steps, save_stride=a.save)
m = engine.Model(ctx, d, syn.theta_init(d))
x0, bcs = syn.columns(d, a.ncol)
x0d, bcsd = torch.tensor(x0, device="cuda"), torch.tensor(bcs, device="cuda")
traj = torch.empty((a.ncol, d.n_saved, d.S), device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream)
with torch.cuda.stream(stream):
    for _ in range(2):
        m.solve_dev(x0d, bcsd, traj)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.reps):
        m.solve_dev(x0d, bcsd, traj)
    e1.record(stream)
e1.synchronize()
ms = e0.elapsed_time(e1) / a.reps
n_rhs = a.steps * d.n_substeps * 6
print(f"stagger={os.environ.get('CPZ_TC_STAGGER','default')} ms/solve {ms:.2f}  col-steps/s {a.ncol*a.steps/ms*1e3:.3e}  cycles/RHS@1.965GHz {ms*1e-3/n_rhs*1.965e9:.0f}")
