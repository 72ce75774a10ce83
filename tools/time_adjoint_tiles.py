"""Training iteration (forward + adjoint + ADAM) of BASELINE config 3 at per-GPU shard sizes, 4-column vs 32-column adjoint
tiles and both checkpoint policies:  python tools/time_adjoint_tiles.py [ncol ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cpzload; cpzload.load()
from cpz_b200 import engine, synthetic as syn
from cpz_b200.desc import RHS_TRAIN

tstream = torch.cuda.Stream(); torch.cuda.set_stream(tstream)
ctx = engine.Context(0, tstream.cuda_stream)
w = np.array([1, 1, 1, 5e-3, 5e-3, 5e-3], dtype=np.float32)
sizes = [int(a) for a in sys.argv[1:]] or [1152, 2304, 4608]
for ncol in sizes:
    for ck in (9, 1):
        d = syn.wind_mixing_desc(variant=RHS_TRAIN, net="uvT_small", n_steps=1152, save_stride=9, ckpt_stride=ck)
        x, b = syn.columns(d, ncol, seed=1000)
        xd, bd = torch.tensor(x, device="cuda"), torch.tensor(b, device="cuda")
        tg = xd[:, None, :].repeat(1, d.n_saved, 1).contiguous()
        for small in ("0", "100000", "auto"):
            if small == "auto":
                os.environ.pop("CPZ_SMALL_NCOL", None)
            else:
                os.environ["CPZ_SMALL_NCOL"] = small
            m = engine.Model(ctx, d, syn.theta_init(d, seed=42, scale=1e-5))
            m.train_step_dev(xd, bd, tg, w, 3e-4)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m.train_step_dev(xd, bd, tg, w, 3e-4); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print(f"ncol {ncol:5d} ckpt_stride {ck} tiles {dict([('0', '32-col'), ('100000', '4-col '), ('auto', 'auto  ')])[small]}: {ms:8.1f} ms  {ncol*1152/ms*1e3:.3e} col-steps/s", flush=True)
            m.close()
