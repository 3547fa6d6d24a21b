#!/usr/bin/env python
"""Headline benchmark of the TorJ ray-tracing hot path (BASELINE.json: ray-steps/s & rays/s of an EC beam bundle).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (torchrun for N>1, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on all host cores (the Julia
                                                           # reference itself cannot run here: no Julia)

A "step" = one pass of the hot path over one bundle: BASELINE.json configs[2], a 65 543-ray Gaussian EC beam
(N_rings=66, min_az=14) on the 257x257 Solov'ev equilibrium, s_max=1.0 m, 1000 psi bins, X-mode 95 GHz.
Weak scaling: every rank traces its own 65 543-ray beam (poloidal steering angle shifted per rank); the only
exchange is one NCCL all-reduce of the [dP_dV | deposited | sum w] vector (reference src/solve.jl:233-240).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="config3_64k_ray_beam", N_rings=66, min_az=14, n_rays=65543, grid="257x257", s_max=1.0,
                n_psi=1000, f=95e9, mode=1, n_gl=24)
# algorithmic FP64 flops per unit (add/mul = 1, FMA = 2, div/sqrt/exp/Bessel call = 1: a lower bound)
F_RHS, F_ALPHA, F_HARM, F_RK_TSIT5, F_DEP = 660.0, 105.0, 1540.0, 440.0, 100.0


def algorithmic_flops(c):
    return F_RHS * c["n_rhs"] + F_ALPHA * c["n_alpha"] + F_HARM * c["n_harm"] + (F_RK_TSIT5 + F_DEP) * c["n_acc"]


def bundle_for_rank(rank, small=False):
    import torj_jl_b200 as tj
    pol = np.deg2rad(30.0 - 1.5 * rank)
    x0 = np.array([2.5, 0.0, 0.4])
    N0 = tj.pol_tor_angles_2_vector(pol, 0.0)
    nr, ma = (7, 20) if small else (WORKLOAD["N_rings"], WORKLOAD["min_az"])
    return tj.launch_peripheral_rays(x0, N0, 0.0174, 1 / 3.99, WORKLOAD["f"], N_rings=nr, min_azimuthal_points=ma)


def sweep_bundle(rank, world):
    """BASELINE.json configs[3]: 32 x 32 launcher angles (pol 10..40 deg, tor -15..15 deg) x 1 025-ray beams =
    1 049 600 rays; the 1 024 beams are dealt round-robin to the ranks (strong scaling; contiguous blocks would give
    every rank a different range of launcher angles, i.e. rays of different lengths)."""
    import torj_jl_b200 as tj
    from torj_jl_b200.distributed import shard_block_cyclic
    x0 = np.array([2.5, 0.0, 0.4])
    P, D, W = [], [], []
    for pol in np.deg2rad(np.linspace(10.0, 40.0, 32)):
        for tor in np.deg2rad(np.linspace(-15.0, 15.0, 32)):
            p, d, w = tj.launch_peripheral_rays(x0, tj.pol_tor_angles_2_vector(pol, tor), 0.0174, 1 / 3.99, WORKLOAD["f"],
                                                N_rings=7, min_azimuthal_points=20)
            P.append(p); D.append(d); W.append(w / 1024.0)
    P, D, W = np.concatenate(P), np.concatenate(D), np.concatenate(W)
    idx = shard_block_cyclic(len(W), 1025, rank, world)
    return P[idx], D[idx], W[idx], len(W)


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of k_trace per launch, from the committed ncu --set full capture of
    this command (profiles/ktrace_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ktrace_traffic.json")) as fh:
            return float(json.load(fh)["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_model():
    """`model name` of /proc/cpuinfo (the CPU baseline names the host it ran on)."""
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(rank_bundle, seconds_target=15.0):
    """Oracle port on all host cores over a bounded sample of the same bundle (rays spread evenly over it)."""
    import torj_jl_b200 as tj
    from oracle import torj_oracle as O
    cores = len(os.sched_getaffinity(0))  # torchrun exports OMP_NUM_THREADS=1; the oracle is told the count explicitly
    arr = tj.solovev_arrays(257, 257)
    opl = O.OraclePlasma(*arr.values())
    gl = np.polynomial.legendre.leggauss(WORKLOAD["n_gl"])
    pos, dirs, w = rank_bundle
    psi = np.linspace(0.0, 1.0, WORKLOAD["n_psi"])
    n_sample = int(max(cores * 2, min(len(w), round(seconds_target * cores / 0.3))))
    pick = np.linspace(0, len(w) - 1, n_sample).astype(int)
    t0 = time.perf_counter()
    # the reference's own deposition algorithm (spline roots per psi level, src/plasma.jl:91-151; psi(s) fitted once)
    r = opl.trace_bundle(pos[pick], dirs[pick], w[pick], WORKLOAD["f"], WORKLOAD["mode"], WORKLOAD["s_max"], psi, gl,
                         deposition="faithful", n_threads=cores)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    opl.trace_bundle(pos[pick[::4]], dirs[pick[::4]], w[pick[::4]], WORKLOAD["f"], WORKLOAD["mode"], WORKLOAD["s_max"], psi, gl,
                     deposition="streaming", n_threads=cores)
    dt_s = (time.perf_counter() - t1) * 4.0  # quarter sample, scaled: the GPU's deposition algorithm on the CPU
    return dict(value=r["counters"]["n_acc"] / dt, value_streaming_deposition=r["counters"]["n_acc"] / dt_s,
                unit="ray-steps/s", cores=cores, kind="port", cpu=cpu_model(),
                sample=f"{n_sample} of {len(w)} rays of the same bundle, evenly spaced, reference deposition algorithm, {dt:.1f} s",
                rays_per_s=n_sample / dt, seconds=dt, n_rays=n_sample, steps=int(r["counters"]["n_acc"]))


def run_reference_arm(args):
    """CPU arm: the reference's algorithm (oracle port; Julia is unavailable) on all host cores, bounded sample/step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bundle = bundle_for_rank(0)
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(bundle, seconds_target=6.0)
        if i >= args.warmup:
            vals.append(last)
    v = float(np.mean([x["value"] for x in vals]))
    ms = float(np.mean([x["seconds"] for x in vals])) * 1e3
    line = {"metric": "ray-steps/s", "value": v, "unit": "ray-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD["name"], **{k: WORKLOAD[k] for k in ("n_rays", "grid", "s_max", "n_psi", "f", "mode")},
                       "note": "Julia reference not runnable here (no Julia); CPU oracle port, OpenMP over rays"},
            "cpu_baseline": {"value": v, "unit": "ray-steps/s", "cores": last["cores"], "cpu": last["cpu"], "kind": "port",
                             "sample": last["sample"]},
            "rays_per_s": float(np.mean([x["rays_per_s"] for x in vals])),
            "e2e": {"value": v, "unit": "ray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--small", action="store_true", help="1 025-ray bundle (configs[1]) instead of 65 543 (profiling aid)")
    ap.add_argument("--workload", default="beam64k", choices=["beam64k", "sweep1m"],
                    help="beam64k: configs[2], weak scaling (default); sweep1m: configs[3], 1 049 600 rays, strong scaling")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--schedule", type=int, default=0, choices=[0, 1, 2],
                    help="torj_options.schedule: 0 automatic (default), 1 whole rays per lane, 2 segment hand-off")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    import torch
    import torch.distributed as dist
    import torj_jl_b200 as tj
    from torj_jl_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = tj.lib()
    tstream = torch.cuda.Stream()  # the library launches on this torch stream so torch events bracket its kernels
    torch.cuda.set_stream(tstream)
    ctx = _lib.context(local, tstream.cuda_stream)
    tj.abs_Al_init(WORKLOAD["n_gl"], ctx)
    pl = tj.Plasma(*tj.solovev_arrays(257, 257).values())
    ph = pl.handle(ctx)
    if args.workload == "sweep1m":
        pos, dirs, w, n_total = sweep_bundle(rank, world)
    else:
        pos, dirs, w = bundle_for_rank(rank, args.small)
    n = len(w)
    psi = np.ascontiguousarray(np.linspace(0.0, 1.0, WORKLOAD["n_psi"]))
    n_psi = len(psi)
    opt = _lib.default_options(schedule=args.schedule)
    dp = lambda a: a.ctypes.data_as(_lib.c_dp)

    # ---- resident bundle: inputs in HBM before the timed region ("value")
    posT, dirT = np.ascontiguousarray(pos.T), np.ascontiguousarray(dirs.T)
    fr = np.array([WORKLOAD["f"]]); md = np.array([WORKLOAD["mode"]], dtype=np.int32)
    bh = _lib.c_vp()
    _lib.check(L.torj_bundle_create(ctx, n, dp(posT), dp(dirT), dp(w), dp(fr), md.ctypes.data_as(_lib.c_ip), 0, C.byref(bh)))
    prof_t = None

    def device_step():
        nonlocal prof_t
        _lib.check(L.torj_bundle_trace(bh, ph, C.byref(opt), WORKLOAD["s_max"], n_psi, dp(psi)))
        if world > 1:
            if prof_t is None:
                # wrap the library's device buffer as a torch tensor (no copy) for the NCCL all-reduce
                class _DevBuf:
                    __cuda_array_interface__ = {"shape": (n_psi + 2,), "typestr": "<f8", "version": 2,
                                                "data": (int(L.torj_bundle_device_profile(bh)), False)}
                prof_t = torch.as_tensor(_DevBuf(), device=torch.device("cuda", local))
            dist.all_reduce(prof_t)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step()
    barrier()
    launches0 = L.torj_ctx_launch_count(ctx)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        device_step()
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = L.torj_ctx_launch_count(ctx) - launches0
    ms_total = ev[0].elapsed_time(ev[-1])
    cnt = _lib.TorjCounters()
    dep = C.c_double()
    status = np.zeros(n, dtype=np.int32)
    _lib.check(L.torj_bundle_results(bh, None, C.byref(dep), None, None, None, status.ctypes.data_as(_lib.c_ip), C.byref(cnt)))
    c = cnt.as_dict()
    rays_ok = torch.tensor([c["n_rays_ok"]], dtype=torch.float64, device="cuda")

    # ---- end to end through the reference-facing call with HOST buffers (H2D of the bundle, D2H of the results)
    def e2e_step():
        r = tj.trace_bundle(pl, pos, dirs, w, WORKLOAD["f"], WORKLOAD["mode"], WORKLOAD["s_max"], psi, ctx=ctx, options=opt)
        if world > 1:
            t = torch.from_numpy(np.concatenate([r["dP_dV"], [r["deposited_power"], w.sum()]])).cuda()
            dist.all_reduce(t)
            t.cpu()
        return r
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r_e2e = e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    h2d = 8 * (3 * n + 3 * n + n + 1 + n_psi) + 4
    d2h = 8 * (n_psi + 2 + 2 * n) + 4 * 2 * n + 48

    # ---- max over ranks
    t_ms = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    tot = torch.tensor([c["n_acc"], n, algorithmic_flops(c)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
        dist.all_reduce(rays_ok)
    ms_total, e2e_ms = t_ms.tolist()
    steps_all, rays_all, flops_all = tot.tolist()
    ms_per_step = ms_total / args.steps

    if rank == 0:
        # roofline of the dominant kernel (k_trace) on this rank: algorithmic FP64 flops / its launch duration.
        # k_ray_init and k_finalize are < 1 % of the step (profiles/), so the step time stands for the kernel.
        tf, tms = C.c_double(), C.c_double()
        _lib.check(L.torj_fp64_peak(ctx, 20000, C.byref(tf), C.byref(tms)))
        kms = C.c_double()
        _lib.check(L.torj_ctx_last_trace_ms(ctx, C.byref(kms)))  # k_trace alone, CUDA events on its own stream
        achieved = algorithmic_flops(c) / (ms_per_step * 1e-3) / 1e12
        roof = {"bound": "fp64", "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved / tf.value,
                "traffic": ncu_traffic(), "kernel": "k_trace<Tsit5>", "kernel_ms_last_launch": kms.value,
                "peak_source": "DFMA-chain microbenchmark torj_fp64_peak measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "algorithmic_flops_per_launch": algorithmic_flops(c), "counters": c}
        line = {"metric": "ray-steps/s", "value": steps_all / (ms_per_step * 1e-3), "unit": "ray-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong" if args.workload == "sweep1m" else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "config4_1M_ray_angle_sweep" if args.workload == "sweep1m" else
                           ("config2_1k_ray_beam" if args.small else WORKLOAD["name"]), "n_rays_per_gpu": n,
                           "grid": WORKLOAD["grid"], "s_max": WORKLOAD["s_max"], "n_psi": n_psi, "f": WORKLOAD["f"],
                           "mode": WORKLOAD["mode"], "scheme": "Tsit5", "schedule": args.schedule, "l2": "compute-bound kernel; tables 4.8 MB resident, "
                           "no L2 flush needed (inputs are re-read from HBM each step: ray state 7.3 MB)"},
                "rays_per_s": rays_all / (ms_per_step * 1e-3), "rays_total": int(rays_all), "rays_ok": int(rays_ok.item()),
                "e2e": {"value": steps_all / (e2e_ms * 1e-3), "unit": "ray-steps/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "rays_per_s": rays_all / (e2e_ms * 1e-3)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
                "absorbed_fraction": dep.value / max(1, world) if world > 1 else dep.value}
        if not args.no_cpu_baseline and world >= 1:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline((pos, dirs, w)).items() if k in ("value", "value_streaming_deposition", "unit", "cores", "cpu", "kind", "sample", "rays_per_s")}
        print(json.dumps(line), flush=True)
    L.torj_bundle_destroy(bh)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
