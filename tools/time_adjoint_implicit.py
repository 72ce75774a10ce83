"""Training iteration of config 3 with the diffusive flux implicit (one sub-step): python tools/time_adjoint_implicit.py [ncol]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cpzload; cpzload.load()
from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import RHS_TRAIN, FLAG_IMPLICIT_DIFFUSION
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 9216
ctx = engine.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
for imp in (1, 0):
    d = syn.wind_mixing_desc(variant=RHS_TRAIN, n_steps=1152, save_stride=9, ckpt_stride=9, **({"n_substeps": 1} if imp else {}))
    if imp: d.flags |= FLAG_IMPLICIT_DIFFUSION
    m = engine.Model(ctx, d, syn.theta_init(d))
    x0, bcs = syn.columns(d, ncol)
    x0d, bcsd = torch.tensor(x0, device="cuda"), torch.tensor(bcs, device="cuda")
    tg = torch.tensor(x0, device="cuda")[:, None, :].repeat(1, d.n_saved, 1).contiguous()
    w = np.array([1, 1, 1, 5e-3, 5e-3, 5e-3], dtype=np.float32)
    with torch.cuda.stream(st):
        m.train_step_dev(x0d, bcsd, tg, w, 3e-4)
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); loss = m.train_step_dev(x0d, bcsd, tg, w, 3e-4); e1.record(st); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
    print(("implicit x1" if imp else "explicit x2"), "ms per iteration:", [round(t, 1) for t in ts], "loss", float(loss[6]))
    m.close()
