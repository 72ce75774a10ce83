"""Timing of the NN-free (mPP-only) forward solve: python tools/time_nnfree.py [ncol ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cpzload; cpzload.load()
from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import RHS_INFER
ctx = engine.Context(0)
for ncol in [int(v) for v in sys.argv[1:]] or [4096, 65536]:
    d = syn.wind_mixing_desc(variant=RHS_INFER, net=None, n_steps=1152, save_stride=1 if ncol <= 8192 else 9)
    m = engine.Model(ctx, d, np.zeros(0, dtype=np.float32))
    print(m.describe().splitlines()[0])
    x0, bcs = syn.columns(d, ncol)
    x0d, bcsd = torch.tensor(x0, device="cuda"), torch.tensor(bcs, device="cuda")
    traj = torch.empty((ncol, d.n_saved, d.S), device="cuda")
    st = torch.cuda.ExternalStream(ctx.stream)
    with torch.cuda.stream(st):
        for i in range(2):
            m.solve_dev(x0d, bcsd, traj)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(3):
            m.solve_dev(x0d, bcsd, traj)
        e1.record(st)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 3
    byt = ncol * d.n_saved * 96 * 4
    print(f"ncol {ncol} save_stride {d.save_stride}: {ms:.2f} ms  {ncol*1152/ms*1e3:.3e} col-steps/s  {byt/ms/1e6:.1f} GB/s")
    m.close()
