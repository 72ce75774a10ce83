"""Per-GPU slices of BASELINE configs 1, 3, 4, 5 (config 2 is bench.py's headline): python tools/time_configs.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cpzload; cpzload.load()
from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import RHS_TRAIN, ClosureDesc

ctx = engine.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)


def timed(fn, reps=3, warm=1):
    with torch.cuda.stream(st):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            fn()
        e1.record(st)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


# config 1: single-column T-only NDE (convective adjustment), forward solve + loss gradient
d = syn.free_convection_desc(ca=True, n_steps=1152, save_stride=9, ckpt_stride=9)
th = syn.theta_init(d)
m = engine.Model(ctx, d, th)
x0, bcs = syn.columns(d, 1)
x0d, bcsd = torch.tensor(x0, device="cuda"), torch.tensor(bcs, device="cuda")
tgt = torch.zeros((1, d.n_saved, d.S), device="cuda")
loss = torch.zeros(8, device="cuda"); grad = torch.zeros(m.P, device="cuda")
w = np.array([0, 0, 1, 0, 0, 0], dtype=np.float32)
ms = timed(lambda: m.loss_grad_dev(x0d, bcsd, tgt, w, loss, grad))
print(f"config 1 (1 column T-only, CA, {d.n_steps} steps x{d.n_substeps} sub-steps, fwd + gradient): {ms:.1f} ms  {1152/ms*1e3:.3e} col-steps/s")
m.close()

# config 4 slice: FreeConvectionNDE inference, 131072 columns per GPU, save every 9th frame / final only
for save in (9, 0):
    d = syn.free_convection_desc(ca=False, n_steps=1152, save_stride=save)
    m = engine.Model(ctx, d, syn.theta_init(d))
    ncol = 131072
    x0, bcs = syn.columns(d, ncol)
    x0d, bcsd = torch.tensor(x0, device="cuda"), torch.tensor(bcs, device="cuda")
    traj = torch.empty((ncol, d.n_saved, d.S), device="cuda")
    ms = timed(lambda: m.solve_dev(x0d, bcsd, traj), reps=1, warm=1)
    print(f"config 4 slice (T-only inference, {ncol} columns, Tsit5 x{d.n_substeps}, save_stride {save}): {ms:.1f} ms  {ncol*1152/ms*1e3:.3e} col-steps/s")
    m.close()

# config 5 slice: gyre closure on a 512 x 64 x 32 y-slab, one call per host-model step
d = syn.free_convection_desc(ca=False)
m = engine.Model(ctx, d, syn.theta_init(d))
nx, ny = 512, 64
T, y = syn.gyre_field(nx, ny, 32)
cd = ClosureDesc(Nx=nx, Ny=ny, Nz=32)
Td, yd = torch.tensor(T, device="cuda"), torch.tensor(y, device="cuda")
f_d, To_d = torch.empty_like(Td), torch.empty_like(Td)
ms = timed(lambda: m.closure_step_dev(cd, Td, yd, f_d, To_d), reps=200, warm=20)
print(f"config 5 slice (closure step on {nx}x{ny}x32 = {nx*ny} columns): {ms*1e3:.1f} us per call  {nx*ny/ms*1e3:.3e} col-steps/s  {3*nx*ny*128/ms/1e6:.1f} GB/s algorithmic")
m.close()
