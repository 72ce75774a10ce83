"""Debug / A-B harness of the tensor-core adjoint: gradient of the same problem through (a) the FP32 SIMT adjoint
(CPZ_NO_TC_ADJ=1), (b) the tensor-core reverse sweep with the plain-FP32 reference contraction (CPZ_WGRAD_REF=1),
(c) the full tensor-core path, each against the FP64 oracle, with a per-layer breakdown of the error."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpzload  # noqa: E402

cpzload.load()
from cpz_b200 import engine, synthetic as syn  # noqa: E402
from cpz_b200.desc import RHS_TRAIN, RHS_INFER, FLAG_MPP, FLAG_CA  # noqa: E402
from util import oracle_loss_grad, oracle_solve  # noqa: E402

W = np.array([0.7, 0.7, 1.0, 3e-3, 3e-3, 5e-3], dtype=np.float32)


def layers(d):
    out, o = [], 0
    for q, n in enumerate(d.nets):
        sz = n.sizes
        for l in range(len(sz) - 1):
            out.append((f"net{q} W{l + 1}", o, o + sz[l] * sz[l + 1])); o += sz[l] * sz[l + 1]
            out.append((f"net{q} b{l + 1}", o, o + sz[l + 1])); o += sz[l + 1]
    return out


def run(ctx, d, th, x0, bcs, tgt, env):
    old = {k: os.environ.get(k) for k in ("CPZ_NO_TC_ADJ", "CPZ_WGRAD_REF")}
    for k in old:
        os.environ.pop(k, None)
    os.environ.update(env)
    m = engine.Model(ctx, d, th)
    t0 = time.time()
    loss, grad = m.loss_grad(x0, bcs, tgt, W)
    dt = time.time() - t0
    m.close()
    for k, v in old.items():
        os.environ.pop(k, None)
        if v is not None:
            os.environ[k] = v
    return loss, grad, dt


def main():
    ctx = engine.Context(0)
    cases = [
        ("train 18 steps ckpt 9, 45 cols", syn.wind_mixing_desc(variant=RHS_TRAIN, n_steps=18, save_stride=9, ckpt_stride=9), 45),
        ("train 12 steps ckpt 4, 100 cols", syn.wind_mixing_desc(variant=RHS_TRAIN, n_steps=12, save_stride=3, ckpt_stride=4), 100),
        ("infer+CA 12 steps ckpt 4, 40 cols", syn.wind_mixing_desc(variant=RHS_INFER, flags=FLAG_MPP | FLAG_CA, n_steps=12, save_stride=4, ckpt_stride=4), 40),
    ]
    which = sys.argv[1:] or ["simt", "ref", "tc"]
    for name, d, ncol in cases:
        th = syn.theta_random(d, scale=0.3)
        x0, bcs = syn.columns(d, ncol)
        rng = np.random.default_rng(0)
        th2 = (th * (1 + 0.3 * rng.standard_normal(th.shape))).astype(np.float32)
        tgt = oracle_solve(d, th2, x0, bcs).astype(np.float32)
        tot, comps, g = oracle_loss_grad(d, th, x0, bcs, tgt, W)
        print(f"== {name}: oracle loss {tot:.6e} |g| {np.linalg.norm(g):.4e}")
        for tag, env in (("simt", {"CPZ_NO_TC_ADJ": "1"}), ("ref", {"CPZ_WGRAD_REF": "1"}), ("tc", {})):
            if tag not in which:
                continue
            try:
                loss, grad, dt = run(ctx, d, th, x0, bcs, tgt, env)
            except Exception as e:  # noqa: BLE001
                print(f"  {tag}: FAILED {e}")
                continue
            e_g = np.linalg.norm(grad - g) / np.linalg.norm(g)
            print(f"  {tag:5s}: loss err {abs(loss[6] - tot) / abs(tot):.2e} comps {np.abs(loss[:6] - comps).max() / abs(tot):.2e} grad err {e_g:.3e} finite {np.isfinite(grad).all()} ({dt * 1e3:.0f} ms)")
            if e_g > 1e-4 or not np.isfinite(grad).all():
                for nm, a, b in layers(d):
                    ng = np.linalg.norm(g[a:b])
                    print(f"      {nm}: |g| {ng:.3e} err {np.linalg.norm(grad[a:b] - g[a:b]) / max(ng, 1e-30):.3e} nan {int(np.isnan(grad[a:b]).sum())}")
    ctx.close()


if __name__ == "__main__":
    main()
