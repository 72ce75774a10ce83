"""Closure-step timing / ncu target: python tools/profile_closure.py [--nx 512] [--ny 64] [--reps 50]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cpzload; cpzload.load()
from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import ClosureDesc
ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=512)
ap.add_argument("--ny", type=int, default=64)
ap.add_argument("--reps", type=int, default=50)
a = ap.parse_args()
ctx = engine.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
d = syn.free_convection_desc(ca=False)
m = engine.Model(ctx, d, syn.theta_init(d))
T, y = syn.gyre_field(a.nx, a.ny, 32)
cd = ClosureDesc(Nx=a.nx, Ny=a.ny, Nz=32)
Td, yd = torch.tensor(T, device="cuda"), torch.tensor(y, device="cuda")
f_d, To_d = torch.empty_like(Td), torch.empty_like(Td)
with torch.cuda.stream(st):
    for _ in range(5):
        m.closure_step_dev(cd, Td, yd, f_d, To_d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(a.reps):
        m.closure_step_dev(cd, Td, yd, f_d, To_d)
    e1.record(st)
e1.synchronize()
ms = e0.elapsed_time(e1) / a.reps
n = a.nx * a.ny
print(f"closure {a.nx}x{a.ny}x32 ({n} columns, {(n+127)//128} tiles): {ms*1e3:.1f} us per call  {n/ms*1e3:.3e} col-steps/s  {3*n*128/ms/1e6:.1f} GB/s")
