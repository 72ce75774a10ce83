"""CUDA-event timing of the u/v/T closure step on a 512 x 64 x 32 slab: python tools/time_closure_uvt.py   (CPZ_NO_TC=1: FP32 kernel)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cpzload; cpzload.load()
from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import ClosureUvtDesc, RHS_INFER
ctx = engine.Context(0)
d = syn.wind_mixing_desc(variant=RHS_INFER)
m = engine.Model(ctx, d, syn.theta_init(d))
nx, ny = 512, 64
u, v, T = syn.uvt_fields(d, nx, ny, unstable_every=3)
cd = ClosureUvtDesc(Nx=nx, Ny=ny, Nz=32, dz=d.H / 32, dt=60.0, uw_top=-1e-4, wT_top=2e-5, convective_adjustment=True)
ud, vd, Td = (torch.tensor(a, device="cuda") for a in (u, v, T))
f = torch.empty((3,) + tuple(Td.shape), device="cuda"); o = torch.empty_like(f)
st = torch.cuda.ExternalStream(ctx.stream)
with torch.cuda.stream(st):
    for _ in range(5): m.closure_step_uvt_dev(cd, ud, vd, Td, f, o)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(50): m.closure_step_uvt_dev(cd, ud, vd, Td, f, o)
    e1.record(st)
e1.synchronize()
print("closure_uvt us per call", e0.elapsed_time(e1) / 50 * 1e3)
