#!/usr/bin/env python
"""bench.py — headline benchmark of the NDE column engine (BASELINE.json: column-steps/s at Nz=32).

Workload at every N: BASELINE config 2 per GPU — wind_mixing u/v/T NDE forward solve (inference RHS, mPP base,
3 x (96->50 mish->20 mish->31) nets, Tsit5 fixed step with the sub-steps explicit diffusion needs), 4,096 synthetic
columns x 1,152 steps, every frame saved (1.81 GB of trajectory per solve, larger than L2) — weak scaling: columns
shard across ranks with no data-path collective. One bench "step" = one full solve of the rank's column batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Prints ONE JSON line (rank 0). `value` = device-resident throughput, `e2e` = through the host-pointer C ABI call
(pinned host buffers, H2D + D2H inside the timed region). Extra objects: roofline (HBM, the contract's key),
roofline (tensor pipe of the tcgen05 kernel), roofline_hbm / roofline_compute (HBM and FP32-SIMT views of the same run), adjoint (fwd+adjoint training step,
BASELINE config 3 slice), nn_free (the HBM-fair mPP-only variant), cpu_baseline (the oracle on host cores).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NCOL = 4096
NSTEPS = 1152
FP32_LANES_PER_SM = 128
# dram__bytes_read.sum + dram__bytes_write.sum of one solve_tc_kernel launch at the bench configuration, from the ncu --set full
# capture summarised in profiles/ (None until captured)
TRAFFIC_NCU_BYTES = 1.769e9  # 10.9 MB read + 1.7584 GB written (profiles/r01_tc_solve_ncu_summary.txt)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured", d.get("bf16_tflops_sustained", 1400.0)
    return 6650.0, 1965.0, "fallback", 1400.0


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for i, nm in enumerate(names):
                    if r[5 + i].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(syn, RHS_INFER, n_steps):
    d = syn.wind_mixing_desc(variant=RHS_INFER, net="uvT_small", n_steps=n_steps, save_stride=1)
    theta = syn.theta_init(d, seed=42, scale=1e-5)
    return d, theta


def flops_per_colstep(d):
    mlp = 2 * sum(n.macs for n in d.nets)
    return (mlp + 1200) * d.rhs_evals_per_step


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path. The Julia reference cannot run here (no julia
    binary, no LES data), so this times the oracle port (restated-reference CPU) with all host threads on a bounded
    sample of the same workload."""
    if rank != 0:
        return
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cpzload
    cpzload.load()
    from cpz_b200 import synthetic as syn
    from cpz_b200.desc import RHS_INFER
    from oracle import nde
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ncol_s, nsteps_s = 512, 12
    d, theta = build_workload(syn, RHS_INFER, nsteps_s)
    x0, bcs = syn.columns(d, ncol_s)
    th, x, b = torch.tensor(theta), torch.tensor(x0), torch.tensor(bcs)
    times = []
    with torch.no_grad():
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            nde.solve(d, th, x, b, None)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    tot = sum(times)
    val = ncol_s * nsteps_s * len(times) / tot
    sample = f"{ncol_s} columns x {nsteps_s} steps (x{d.n_substeps} sub-steps, Tsit5) per step, FP32 torch-CPU oracle batched over columns"
    line = {
        "impl": "reference", "metric": "column-steps/sec (forward)", "value": val, "unit": "column-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE config 2 (wind_mixing u/v/T forward, Nz=32, small nets, Tsit5) — bounded sample",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": "column-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "column-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference Julia path cannot run in this image (no julia, no LES data); restated-reference CPU oracle timed instead",
    }
    emit(line)


def cpu_baseline(syn, RHS_INFER):
    """Bounded oracle run on the host cores (about 10-30 s)."""
    import torch
    from oracle import nde
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ncol_s, nsteps_s = 512, 12
    d, theta = build_workload(syn, RHS_INFER, nsteps_s)
    x0, bcs = syn.columns(d, ncol_s)
    th, x, b = torch.tensor(theta), torch.tensor(x0), torch.tensor(bcs)
    with torch.no_grad():
        nde.solve(d, th, x, b, None)  # warm-up
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 10.0 and reps < 20:
            nde.solve(d, th, x, b, None)
            reps += 1
        el = time.perf_counter() - t0
    return {"value": ncol_s * nsteps_s * reps / el, "unit": "column-steps/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x ({ncol_s} columns x {nsteps_s} steps x {d.n_substeps} sub-steps), FP32 torch-CPU oracle batched over columns"}


_REAL_STDOUT = None


def quiet_stdout():
    """Everything a library prints to fd 1 (e.g. NCCL's version banner) goes to stderr; the one JSON line is written to
    the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip adjoint / nn_free / cpu_baseline side measurements")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import cpzload
    cpzload.load()
    from cpz_b200 import engine, synthetic as syn
    from cpz_b200.desc import RHS_INFER, RHS_TRAIN

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # kernels are launched on a torch-owned stream so that torch.cuda.Event brackets them
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    ctx = engine.Context(local_rank, tstream.cuda_stream)
    d, theta = build_workload(syn, RHS_INFER, NSTEPS)
    model = engine.Model(ctx, d, theta)
    S, n_saved = d.S, d.n_saved
    x0, bcs = syn.columns(d, NCOL, seed=1000 + rank)
    x0_d, bcs_d = torch.tensor(x0, device="cuda"), torch.tensor(bcs, device="cuda")
    traj_d = torch.empty((NCOL, n_saved, S), dtype=torch.float32, device="cuda")

    # ---- device-resident timing (value) ----
    for _ in range(args.warmup):
        model.solve_dev(x0_d, bcs_d, traj_d)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_all0 = torch.cuda.Event(enable_timing=True); t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for e0, e1 in evs:
        e0.record()
        model.solve_dev(x0_d, bcs_d, traj_d)
        e1.record()
    t_all1.record()
    barrier()
    launches = ctx.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = t_all0.elapsed_time(t_all1)
    kern_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in evs]))
    t = torch.tensor([ms_total], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    colsteps_per_step = NCOL * NSTEPS * world
    value = colsteps_per_step * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the host-pointer C ABI (pinned host buffers; H2D + D2H inside the timed region) ----
    x0_h = torch.tensor(x0).pin_memory(); bcs_h = torch.tensor(bcs).pin_memory()
    traj_h = torch.empty((NCOL, n_saved, S), dtype=torch.float32).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))
    model.solve(x0_h.numpy(), bcs_h.numpy(), out=traj_h.numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        model.solve(x0_h.numpy(), bcs_h.numpy(), out=traj_h.numpy())  # synchronous on return
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = colsteps_per_step * e2e_steps / float(t.item())
    checksum = float(traj_h[:, -1].double().abs().mean())

    hbm_peak, sm_max_mhz, peak_src, bf16_peak = peaks()
    bytes_per_colstep = 4 * S * (n_saved / NSTEPS) + (4 * S + 4 * d.n_bc) / NSTEPS
    ach_gbs = bytes_per_colstep * NCOL * NSTEPS / (kern_ms * 1e-3) / 1e9
    fl = flops_per_colstep(d)
    ach_tf = fl * NCOL * NSTEPS / (kern_ms * 1e-3) / 1e12
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max_mhz
    peak_tf_max = n_sm * FP32_LANES_PER_SM * 2 * sm_max_mhz * 1e6 / 1e12
    peak_tf_obs = n_sm * FP32_LANES_PER_SM * 2 * sm_mhz * 1e6 / 1e12

    tc = "forward kernel: tcgen05" in model.describe()
    if tc:
        # dominant kernel: solve_tc_kernel (tcgen05 kind::tf32, 3xTF32). achieved = ALGORITHMIC tensor flops: every MAC of
        # the three MLPs is three TF32 MACs (hi*lo + lo*hi + hi*hi); peak = measured dense bf16 (sustained, the kernel
        # runs for tens of ms); TF32 runs at half the bf16 rate, so frac_of_tf32_peak = 2*frac.
        mlp_macs = sum(n.macs for n in d.nets)
        tf32_flops = 3 * 2 * mlp_macs * d.rhs_evals_per_step * NCOL * NSTEPS
        h1 = d.nets[0].sizes[1]; h2 = d.nets[0].sizes[2]
        mma_per_rhs_group = (72 if 3 * h1 > 128 else 36) + 9 * 7 + 9 * (3 if h2 <= 24 else 4)
        issued_flops = mma_per_rhs_group * (2 * 128 * 16 * 8) * (NCOL / 16) * d.rhs_evals_per_step * NSTEPS
        ach_tensor = tf32_flops / (kern_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach_tensor, "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach_tensor / bf16_peak,
                    "traffic": TRAFFIC_NCU_BYTES, "peak_source": peak_src + " (bf16_tflops_sustained)",
                    "kernel": "solve_tc_kernel<ACT_MISH,3>", "kernel_ms": kern_ms,
                    "algorithmic_flop_per_colstep": 3 * 2 * mlp_macs * d.rhs_evals_per_step,
                    "frac_of_tf32_peak": 2 * ach_tensor / bf16_peak,
                    "issued_tflops_incl_padding": issued_flops / (kern_ms * 1e-3) / 1e12,
                    "pipe_tensor_active_pct_ncu": 44.7,
                    "note": "3xTF32 on tcgen05 with M=128 x N=16 x K=8 MMAs; padding of the 150/60/93 output rows to 128-row "
                            "blocks makes issued flops 2.9x the algorithmic ones; the kernel is latency-bound (ncu: tensor pipe "
                            "active 45 %, issue slots 43 %), see profiles/r01_tc_solve_ncu_summary.txt"}
    else:
        roofline = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "traffic": None,
                    "peak_source": peak_src, "kernel": "solve_kernel<32,256,true>", "kernel_ms": kern_ms,
                    "algorithmic_bytes_per_colstep": bytes_per_colstep}
    line = {
        "metric": "column-steps/sec (forward)", "value": value, "unit": "column-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE config 2: wind_mixing u/v/T NDE forward solve, 4096 synthetic columns x 1152 steps per GPU, all frames saved",
                   "Nz": 32, "columns_per_gpu": NCOL, "n_steps": NSTEPS, "n_substeps": d.n_substeps, "integrator": d.integrator,
                   "rhs_evals_per_step": d.rhs_evals_per_step, "nets": "3 x (96->50 mish->20 mish->31), P=19563",
                   "rhs": "inference (solve_NDE_mutating), mPP base, zero_weights BCs",
                   "arithmetic": "FP32 state/stencil/Runge-Kutta; MLP as 3xTF32 on tcgen05 (hi/lo operand split, FP32 accumulate in TMEM), RHS within 6e-7 of the FP64 oracle",
                   "l2": "no explicit flush: each solve writes a 1.81 GB trajectory (> 126 MB L2)", "parallelism": f"columns x{world}"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "column-steps/s", "h2d_bytes_per_step": int(x0_h.numel() * 4 + bcs_h.numel() * 4),
                "d2h_bytes_per_step": int(traj_h.numel() * 4), "steps": e2e_steps, "final_state_checksum": checksum},
        "roofline": roofline,
        "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         "peak_source": peak_src, "algorithmic_bytes_per_colstep": bytes_per_colstep,
                         "note": "not the binding roof: arithmetic intensity ~ 1200 flop/B"},
        "roofline_compute": {"bound": "fp32_simt", "achieved": ach_tf, "unit": "TFLOP/s", "peak_at_max_clock": peak_tf_max,
                             "peak_at_observed_clock": peak_tf_obs, "frac": ach_tf / peak_tf_obs, "flop_per_colstep": fl,
                             "stage_evals_per_s": value * d.rhs_evals_per_step,
                             "note": "FP32-equivalent algorithmic flops against the FP32 SIMT peak (the roof of the non-tensor-core kernel)"},
    }
    if not args.no_extras:
        # ---- fwd + discrete adjoint + ADAM (BASELINE config 3: 9 forcing cases x 1024 columns, sharded over ranks) ----
        try:
            from cpz_b200 import parallel
            if dist is not None:
                parallel.attach_torch_allreduce(ctx)
            NC3 = 9216
            lo, hi = parallel.shard_columns(NC3, rank, world)
            d3 = syn.wind_mixing_desc(variant=RHS_TRAIN, net="uvT_small", n_steps=NSTEPS, save_stride=9, ckpt_stride=1)  # a checkpoint every step (4.1 GB of 180 GB): no segment recompute
            m3 = engine.Model(ctx, d3, syn.theta_init(d3, seed=42, scale=1e-5))
            x3, b3 = syn.columns(d3, NC3, seed=1000)
            x3_d, b3_d = torch.tensor(x3[lo:hi], device="cuda"), torch.tensor(b3[lo:hi], device="cuda")
            # targets: a smooth drift of the initial profiles (any target gives the same cost)
            tg = torch.tensor(x3[lo:hi], device="cuda")[:, None, :].repeat(1, d3.n_saved, 1).contiguous()
            tg += 0.05 * torch.linspace(0, 1, d3.n_saved, device="cuda")[None, :, None]
            w3 = np.array([1, 1, 1, 5e-3, 5e-3, 5e-3], dtype=np.float32)
            m3.train_step_dev(x3_d, b3_d, tg, w3, 3e-4)  # warm-up (allocates scratch)
            barrier()
            l0 = ctx.launch_count
            n3 = 2
            a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(n3):
                loss3 = m3.train_step_dev(x3_d, b3_d, tg, w3, 3e-4)
            a1.record()
            barrier()
            t3 = torch.tensor([a0.elapsed_time(a1)], device="cuda")
            if dist is not None:
                dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            ms3 = float(t3.item()) / n3
            bytes3 = 129.0
            line["adjoint"] = {
                "metric": "column-steps/sec (forward + discrete adjoint + ADAM)", "value": NC3 * NSTEPS / (ms3 * 1e-3),
                "unit": "column-steps/s", "ms_per_step": ms3, "scaling": "strong", "gpu_launches_per_step": int((ctx.launch_count - l0) / n3),
                "config": {"workload": "BASELINE config 3: 9 forcing cases x 1024 columns, training RHS, 1152 steps, 129 saved frames, ckpt_stride 1, Tsit5 x2 sub-steps, ADAM(3e-4), one allreduce of P+8 floats",
                           "columns_total": NC3, "columns_this_rank": hi - lo},
                "loss": float(loss3[6]),
                "roofline_hbm_frac": bytes3 * (hi - lo) * NSTEPS / (ms3 * 1e-3) / 1e9 / hbm_peak,
            }
            m3.close()
        except Exception as e:  # noqa: BLE001
            line["adjoint"] = {"error": repr(e)}
        # ---- NN-free mPP-only forward (SURVEY 8d '2-base', the HBM-fair variant) ----
        try:
            d0 = syn.wind_mixing_desc(variant=RHS_INFER, net=None, n_steps=NSTEPS, save_stride=1)
            m0 = engine.Model(ctx, d0, np.zeros(0, dtype=np.float32))
            m0.solve_dev(x0_d, bcs_d, traj_d)
            barrier()
            b0 = torch.cuda.Event(enable_timing=True); b1 = torch.cuda.Event(enable_timing=True)
            b0.record()
            for _ in range(3):
                m0.solve_dev(x0_d, bcs_d, traj_d)
            b1.record()
            barrier()
            ms0 = b0.elapsed_time(b1) / 3
            gb0 = bytes_per_colstep * NCOL * NSTEPS / (ms0 * 1e-3) / 1e9
            line["nn_free"] = {"metric": "column-steps/sec (forward, mPP only)", "value": NCOL * NSTEPS * world / (ms0 * 1e-3),
                               "ms_per_step": ms0, "roofline": {"bound": "hbm", "achieved": gb0, "peak": hbm_peak, "unit": "GB/s", "frac": gb0 / hbm_peak}}
            m0.close()
        except Exception as e:  # noqa: BLE001
            line["nn_free"] = {"error": repr(e)}
        # ---- BASELINE config 4 slice: FreeConvectionNDE inference, 131072 columns per GPU (tcgen05, columns on the M side) ----
        try:
            NC4 = 131072
            d4 = syn.free_convection_desc(ca=False, n_steps=NSTEPS, save_stride=9)
            m4 = engine.Model(ctx, d4, syn.theta_init(d4, seed=42, scale=1e-5))
            x4, b4 = syn.columns(d4, NC4, seed=1000 + rank)
            x4_d, b4_d = torch.tensor(x4, device="cuda"), torch.tensor(b4, device="cuda")
            t4_d = torch.empty((NC4, d4.n_saved, d4.S), dtype=torch.float32, device="cuda")
            m4.solve_dev(x4_d, b4_d, t4_d)
            barrier()
            a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(2):
                m4.solve_dev(x4_d, b4_d, t4_d)
            a1.record()
            barrier()
            t4 = torch.tensor([a0.elapsed_time(a1)], device="cuda")
            if dist is not None:
                dist.all_reduce(t4, op=dist.ReduceOp.MAX)
            ms4 = float(t4.item()) / 2
            macs4 = sum(n.macs for n in d4.nets)
            line["free_convection"] = {
                "metric": "column-steps/sec (forward, T-only FreeConvectionNDE)", "value": NC4 * NSTEPS * world / (ms4 * 1e-3),
                "ms_per_step": ms4, "scaling": "weak",
                "config": {"workload": "BASELINE config 4 slice: 131072 columns per GPU x 1152 steps, 32->128->128->31 relu, Tsit5, save every 9th frame",
                           "kernel": m4.describe().splitlines()[0]},
                "tensor_tflops_tf32_algorithmic": 3 * 2 * macs4 * d4.rhs_evals_per_step * NC4 * NSTEPS / (ms4 * 1e-3) / 1e12}
            m4.close()
        except Exception as e:  # noqa: BLE001
            line["free_convection"] = {"error": repr(e)}
        # ---- BASELINE config 5 slice: per-step closure on a 512 x 64 x 32 y-slab (one call per host-model step) ----
        try:
            from cpz_b200.desc import ClosureDesc
            d5 = syn.free_convection_desc(ca=False)
            m5 = engine.Model(ctx, d5, syn.theta_init(d5, seed=42, scale=1e-5))
            nx5, ny5 = 512, 64
            T5, y5 = syn.gyre_field(nx5, ny5, 32)
            cd5 = ClosureDesc(Nx=nx5, Ny=ny5, Nz=32)
            T5_d, y5_d = torch.tensor(T5, device="cuda"), torch.tensor(y5, device="cuda")
            f5_d, o5_d = torch.empty_like(T5_d), torch.empty_like(T5_d)
            for _ in range(20):
                m5.closure_step_dev(cd5, T5_d, y5_d, f5_d, o5_d)
            barrier()
            n5 = 500
            a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(n5):
                m5.closure_step_dev(cd5, T5_d, y5_d, f5_d, o5_d)
            a1.record()
            barrier()
            t5 = torch.tensor([a0.elapsed_time(a1)], device="cuda")
            if dist is not None:
                dist.all_reduce(t5, op=dist.ReduceOp.MAX)
            us5 = float(t5.item()) / n5 * 1e3
            gb5 = 3 * nx5 * ny5 * 128 / (us5 * 1e-6) / 1e9
            line["closure"] = {
                "metric": "column-steps/sec (NN closure applied every host-model step)", "value": nx5 * ny5 * world / (us5 * 1e-6),
                "us_per_call": us5, "scaling": "weak",
                "config": {"workload": "BASELINE config 5 slice: 512 x 64 x 32 y-slab per GPU, implicit convective adjustment + NN forcing, T read / T' + forcing written every call"},
                "roofline": {"bound": "hbm", "achieved": gb5, "peak": hbm_peak, "unit": "GB/s", "frac": gb5 / hbm_peak,
                             "algorithmic_bytes_per_colstep": 384}}
            m5.close()
        except Exception as e:  # noqa: BLE001
            line["closure"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_extras:  # the CPU baseline is reported at N = 1 only
        try:
            line["cpu_baseline"] = cpu_baseline(syn, RHS_INFER)
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": repr(e)}
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
