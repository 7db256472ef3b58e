#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched fire-spread step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE ForestFire.step over the whole batch of the workload (c2: 4096 envs of 14x14 per
GPU, Logs/14-sized constants, actions from the shared Philox ACTION stream, auto-reset on done).
Prints one JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.

  value      device-resident throughput: K steps issued as fused wf_rollout launches of `chunk`
             steps each (obs/reward/done of EVERY step are written to HBM), CUDA events, max over ranks
  e2e        same metric through the host-buffer C-ABI call wf_step_host: per step H2D actions,
             step, D2H obs + reward + done, synchronise -- wall clock
  roofline   warp_kernel, algorithmic bytes per SURVEY.md 8(d): 15 B per cell-update
  cpu_baseline  the C oracle (a port of the reference step) on one host core, bounded sample
--impl reference times the oracle port on all host threads (the Python reference itself cannot
travel to the GPU box); its line carries "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: 14x14 (Logs/14-sized constants) batched 4096 envs, random actions
    "c2": dict(n_envs=4096, meta=dict(width=14, height=14), chunk=64,
               desc="14x14 Logs/14-sized constants, 4096 envs/GPU, ACTION-stream random actions, auto-reset"),
    # configs[3]: 256x256, 1024 envs, wind enabled, multi-ignition stress of the stencil
    "c4": dict(n_envs=1024, meta=dict(width=256, height=256, wind=[0.85, (1, 0)], extra_ignitions=32), chunk=16,
               desc="256x256, wind [0.85,(1,0)], 32 extra ignitions, 1024 envs/GPU, random actions, auto-reset"),
    # configs[4]: 1024x1024 grid, 64 envs per GPU
    "c5": dict(n_envs=64, meta=dict(width=1024, height=1024, extra_ignitions=256), chunk=16,
               desc="1024x1024, no wind, 256 extra ignitions, 64 envs/GPU, random actions, auto-reset"),
}
BYTES_PER_CELL_UPDATE = 15  # 6 B state read + 6 B state write + 3 B uint8 observation


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def oracle_cfg(meta):
    return dict(seed=0, **meta)


def run_reference(args, wl):
    """The reference's CPU implementation of the path, as ported in oracle/ (test infrastructure:
    this is one of the two places bench.py may execute it), on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import wf_oracle as wo
    cores = os.cpu_count() or 1
    W, H = wl["meta"]["width"], wl["meta"]["height"]
    n_full = wl["n_envs"]
    probe = wo.OracleBatch(oracle_cfg(wl["meta"]), min(n_full, 256), cores)
    t0 = time.perf_counter(); n = probe.step(4); rate = n / (time.perf_counter() - t0)
    del probe
    # bounded sample: as many of the workload's envs as keep the whole run under ~90 s
    budget_env_steps = rate * 90.0
    n_envs = int(max(cores, min(n_full, budget_env_steps / max(1, args.steps + args.warmup))))
    batch = wo.OracleBatch(oracle_cfg(wl["meta"]), n_envs, cores)
    for _ in range(args.warmup):
        batch.step(1)
    t0 = time.perf_counter()
    done_steps = 0
    for _ in range(args.steps):
        done_steps += batch.step(1)
    dt = time.perf_counter() - t0
    value = done_steps / dt
    sample = f"{n_envs} of {n_full} envs x {args.steps} steps, ACTION-stream actions, reset on done, {cores} threads"
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": f"{args.workload}: {wl['desc']}", "grid": [W, H]},
        "cell_updates_per_sec": value * W * H,
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "C oracle port of Simulation/forest_fire.py + environment.py (validated step-by-step against the "
                "Python reference); the Python reference itself runs ~3.3e3 env-steps/s/core (BASELINE.md section 2)",
    }
    print(json.dumps(line), flush=True)


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from wildfire_control_python_b200 import BatchedForestFire

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N = wl["n_envs"]
    W, H = wl["meta"]["width"], wl["meta"]["height"]
    K, Wm = args.steps, args.warmup
    chunk = max(1, min(args.chunk or wl["chunk"], K))
    env = BatchedForestFire(N, device=dev, auto_reset=True, seed=0, env_id_base=rank * N, **wl["meta"])
    env.reset()
    fused = env.kernel_family == "warp"
    obs_buf = torch.empty((chunk, N, W, H, 3), dtype=torch.uint8, device=dev)
    rew_buf = torch.empty((chunk, N), dtype=torch.float64, device=dev)
    done_buf = torch.empty((chunk, N), dtype=torch.uint8, device=dev)
    tile_actions = None
    if not fused:  # tile family: one launch group per step, explicit random actions resident in HBM
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        tile_actions = torch.randint(0, 4, (max(K, Wm, 1), N), dtype=torch.int32, device=dev, generator=g)

    def run_steps(n_steps, offset=0):
        s = 0
        while s < n_steps:
            c = min(chunk, n_steps - s)
            out = (obs_buf[:c], rew_buf[:c], done_buf[:c])
            if fused:
                env.rollout(c, actions=None, out=out)
            else:
                env.rollout(c, actions=tile_actions[offset + s: offset + s + c], out=out)
            s += c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run_steps(max(Wm, 3))
    barrier()
    launches0 = env.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        run_steps(K)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = env.launch_count - launches0
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * N * K / (ms_max * 1e-3)

    # ---- per-step launches (policy-in-the-loop shape): one wf_step per step, actions resident in HBM
    Kp = min(K, 2000)
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    acts = torch.randint(0, 4, (Kp, N), dtype=torch.int32, device=dev, generator=g)
    for k in range(3):
        env.step(acts[k])
    barrier()
    ev0.record()
    for k in range(Kp):
        env.step(acts[k])
    ev1.record()
    barrier()
    t2 = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    per_step = {"value": world * N * Kp / (float(t2.item()) * 1e-3), "unit": "env-steps/s", "steps": Kp,
                "us_per_step": float(t2.item()) * 1e3 / Kp, "launches_per_step": 1 if fused else None,
                "issue": "python loop over BatchedForestFire.step"}
    # the same per-step launches replayed from a CUDA graph (no host work between launches)
    graph_info = None
    try:
        Kg = min(Kp, 200)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            env.step(acts[0])
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for k in range(Kg):
                    env.step(acts[k])
        torch.cuda.current_stream(dev).wait_stream(side)
        graph.replay()
        barrier()
        reps = max(1, Kp // Kg)
        ev0.record()
        for _ in range(reps):
            graph.replay()
        ev1.record()
        barrier()
        t3 = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        graph_info = {"value": world * N * Kg * reps / (float(t3.item()) * 1e-3), "unit": "env-steps/s",
                      "steps": Kg * reps, "us_per_step": float(t3.item()) * 1e3 / (Kg * reps)}
    except Exception as exc:  # graph capture is a convenience measurement, never the headline
        graph_info = {"error": repr(exc)[:200]}
    per_step["cuda_graph"] = graph_info

    # ---- end to end through the host-buffer C-ABI call
    Ke = min(K, 300)
    import numpy as np
    rng = np.random.default_rng(7 + rank)
    host_actions = rng.integers(0, 4, size=(Ke + 3, N), dtype=np.int32)
    for k in range(3):
        env.step_host(host_actions[k])
    barrier()
    t0 = time.perf_counter()
    for k in range(Ke):
        env.step_host(host_actions[3 + k])
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N * Ke / float(te.item())
    h2d = N * 4
    d2h = N * W * H * 3 + N * 8 + N

    clocks = clk.summary()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    per_gpu_steps = N * K / (ms_max * 1e-3)
    achieved = per_gpu_steps * W * H * BYTES_PER_CELL_UPDATE / 1e9
    own_bytes = W * H * 3 + 8 + 1 + (0 if fused else 4)  # what this layout must move per env-step (obs+reward+done)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get(f"{args.workload}_chunk{chunk}")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": "wf::warp_kernel" if fused else "wf::tile_step_kernel",
                "bytes_per_unit": W * H * BYTES_PER_CELL_UPDATE, "unit_name": "env-step",
                "units_per_launch": N * chunk, "avg_launch_ms": ms_max / max(1, launches),
                "own_layout_bytes_per_unit": own_bytes,
                "achieved_own_layout_gbs": per_gpu_steps * own_bytes / 1e9,
                "note": "c2 state (4096 envs x 0.8 KB) lives in registers/L2: launch- and issue-bound, not HBM-bound"
                if fused else "HBM-bound stencil"}

    cpu_baseline = None
    if world == 1:
        from oracle import wf_oracle as wo  # the checker, timed as the CPU baseline (allowed use)
        ob = wo.OracleBatch(oracle_cfg(wl["meta"]), min(N, 1024), 1)
        t0 = time.perf_counter(); n = ob.step(2); rate = n / (time.perf_counter() - t0)
        steps_cpu = max(1, int(rate * 12.0 / ob.n_envs))
        t0 = time.perf_counter(); n = ob.step(steps_cpu); dt = time.perf_counter() - t0
        cpu_baseline = {"value": n / dt, "unit": "env-steps/s", "cores": 1, "kind": "port",
                        "sample": f"{ob.n_envs} envs x {steps_cpu} steps of the same workload, 1 thread, {dt:.1f} s",
                        "host_cpus": os.cpu_count()}

    line = {
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K,
        "warmup": max(Wm, 3), "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32-bitplanes+u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "grid": [W, H], "envs_per_gpu": N,
                   "steps_per_launch": chunk, "kernel_family": env.kernel_family,
                   "l2": "state is register/L2 resident by design; the per-launch output stream "
                         f"({obs_buf.numel() / 1e6:.0f} MB obs) is larger than L2" if fused else "state + obs larger than L2"},
        "cell_updates_per_sec": value * W * H,
        "cell_updates_per_sec_per_gpu": value * W * H / world,
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": Ke, "api": "wf_step_host (pinned host buffers)"},
        "per_step_launch": per_step,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "clocks": clocks,
        "stats": env.stats(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--chunk", type=int, default=0, help="steps per fused launch (default: workload's)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
