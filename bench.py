#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched fire-spread step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE ForestFire.step over the whole batch of the workload (default c2 = BASELINE.json
configs[1]: 4096 envs of 14x14 per GPU, Logs/14-sized constants, actions from the shared Philox
ACTION stream, auto-reset on done).  Prints one JSON line (rank 0); DESIGN.md section 5 explains
every field.

  value         device-resident throughput over EXACTLY K timed steps (CUDA events, max over ranks).
                Both families run `steps_per_launch` steps per wf_rollout launch (warp family: state in
                registers; tile family: one thread-block cluster per env).  Obs/reward/done of EVERY step
                are written to HBM into a buffer larger than L2.
  e2e           the same metric through the host-buffer C-ABI call wf_step_host (page-locked host
                buffers): per step actions H2D, step, obs + reward + done D2H, synchronise; wall clock.
                For grids up to 32x32 the observation crosses PCIe as a bit stream (d2h_bytes_per_step
                counts those bytes) and the library's host threads expand it into the caller's uint8
                [N][W][H][3] buffer inside the timed region (obs_bytes_delivered_per_step).
  per_step_launch   one wf_step per step with device-resident actions (Python loop, and CUDA graph).
  roofline      dominant kernel; algorithmic bytes per SURVEY.md 8(d) (15 B per cell-update) and,
                beside it, the bytes this layout must move.
  cpu_baseline  the C oracle (a port of the reference step, oracle/) on ONE host core, bounded sample.
  secondary     (default workload only) the 256x256 stencil stress case c4, measured the same way.
--impl reference times the oracle port on all host threads (the Python reference cannot travel to
the GPU box); its line carries "impl": "reference".
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
    # 256 steps per launch: the ~15 us a launch costs beyond its steps (launch gap, state load/store, table set-up,
    # the tail of the slowest warp) falls from 9 % of a 64-step launch to 2.5 % (tools/ab_rollout.py, r01)
    "c2": dict(n_envs=4096, meta=dict(width=14, height=14), chunk=256,
               desc="14x14 Logs/14-sized constants, 4096 envs/GPU, ACTION-stream random actions, auto-reset"),
    # configs[2]: 14x14 batched 65536 envs (8192 per GPU on 8 GPUs) with a DQN-policy rollout via torch:
    # Flatten -> Dense(50, sigmoid) -> Dense(4) (DQN.py:209-233), fixed-seed weights, epsilon-greedy 0.1
    "c3": dict(n_envs=8192, meta=dict(width=14, height=14), chunk=1, policy=True,
               desc="14x14 Logs/14-sized constants, 8192 envs/GPU, torch MLP policy 588-50-4 in the loop (eps 0.1), auto-reset"),
    # configs[3]: 256x256, 1024 envs, wind enabled, multi-ignition stress of the stencil
    "c4": dict(n_envs=1024, meta=dict(width=256, height=256, wind=[0.85, (1, 0)], extra_ignitions=32), chunk=16,
               desc="256x256, wind [0.85,(1,0)], 32 extra ignitions, 1024 envs/GPU, ACTION-stream actions, auto-reset"),
    # configs[4]: 1024x1024 grid, 64 envs per GPU
    "c5": dict(n_envs=64, meta=dict(width=1024, height=1024, extra_ignitions=256), chunk=16,
               desc="1024x1024, no wind, 256 extra ignitions, 64 envs/GPU, ACTION-stream actions, auto-reset"),
}
BYTES_PER_CELL_UPDATE = 15  # 6 B state read + 6 B state write + 3 B uint8 observation
CPU_BASELINE_SECONDS = 12.0  # bounded sample of the cpu_baseline leg


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
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
            time.sleep(0.1)
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
    n_envs = int(max(min(cores, n_full), min(n_full, budget_env_steps / max(1, args.steps + args.warmup))))
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


class Dist:
    def __init__(self, gpus):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world == 1 and gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def measure(D: Dist, name: str, K: int, Wm: int, chunk_arg: int, no_graph: bool, with_e2e: bool, with_per_step: bool):
    """Time workload `name` on this rank's GPU; every rank calls this, results are max-reduced."""
    import numpy as np
    torch = D.torch
    from wildfire_control_python_b200 import BatchedForestFire

    wl = WORKLOADS[name]
    dev, rank, world = D.dev, D.rank, D.world
    N, W, H = wl["n_envs"], wl["meta"]["width"], wl["meta"]["height"]
    chunk = max(1, min(chunk_arg or wl["chunk"], K))
    env = BatchedForestFire(N, device=dev, auto_reset=True, seed=0, env_id_base=rank * N, **wl["meta"])
    env.reset()
    fused = True  # warp family: state in registers across the launch; tile family: one cluster per env, K steps per launch
    obs_buf = torch.empty((chunk, N, W, H, 3), dtype=torch.uint8, device=dev)
    rew_buf = torch.empty((chunk, N), dtype=torch.float64, device=dev)
    done_buf = torch.empty((chunk, N), dtype=torch.uint8, device=dev)
    out = (obs_buf, rew_buf, done_buf)

    graph, launches_per_chunk = None, 0  # both families run `chunk` steps per launch: nothing to capture
    replays = [0]

    def run_steps(n_steps):
        s = 0
        while s < n_steps:
            c = min(chunk, n_steps - s)
            if graph is not None and c == chunk:
                graph.replay()
                replays[0] += 1
            else:
                env.rollout(c, actions=None, out=(obs_buf[:c], rew_buf[:c], done_buf[:c]))
            s += c

    run_steps(max(Wm, 3))
    D.barrier()
    launches0, replays0 = env.launch_count, replays[0]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    ev0.record()
    run_steps(K)
    ev1.record()
    D.barrier()
    ms_max = D.max_over_ranks(ev0.elapsed_time(ev1))
    launches = env.launch_count - launches0 + (replays[0] - replays0) * launches_per_chunk
    res = {"name": name, "wl": wl, "N": N, "W": W, "H": H, "K": K, "chunk": chunk, "fused": fused, "graph": graph is not None,
           "ms": ms_max, "launches": int(launches), "value": world * N * K / (ms_max * 1e-3),
           "obs_mb": obs_buf.numel() / 1e6, "family": env.kernel_family, "state_bytes": env.state_bytes_per_env}

    if with_per_step:
        Kp = min(K, 2000)
        g = torch.Generator(device=dev).manual_seed(99 + rank)
        acts = torch.randint(0, 4, (Kp, N), dtype=torch.int32, device=dev, generator=g)
        for k in range(3):
            env.step(acts[k])
        D.barrier()
        ev0.record()
        for k in range(Kp):
            env.step(acts[k])
        ev1.record()
        D.barrier()
        t2 = D.max_over_ranks(ev0.elapsed_time(ev1))
        per_step = {"value": world * N * Kp / (t2 * 1e-3), "unit": "env-steps/s", "steps": Kp, "us_per_step": t2 * 1e3 / Kp,
                    "issue": "python loop over BatchedForestFire.step, actions resident in HBM"}
        try:  # the same per-step launches replayed from a CUDA graph (no host work between launches)
            Kg = max(2, min(Kp, 200) & ~1)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                env.step(acts[0]); env.step(acts[1 % Kp])
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, stream=side):
                    for k in range(Kg):
                        env.step(acts[k % Kp])
            torch.cuda.current_stream(dev).wait_stream(side)
            g2.replay()
            D.barrier()
            reps = max(1, Kp // Kg)
            ev0.record()
            for _ in range(reps):
                g2.replay()
            ev1.record()
            D.barrier()
            t3 = D.max_over_ranks(ev0.elapsed_time(ev1))
            per_step["cuda_graph"] = {"value": world * N * Kg * reps / (t3 * 1e-3), "unit": "env-steps/s", "steps": Kg * reps,
                                      "us_per_step": t3 * 1e3 / (Kg * reps),
                                      "note": "obs written to the same buffer every step (may stay in L2)"}
        except Exception as exc:  # a convenience measurement, never the headline
            per_step["cuda_graph"] = {"error": repr(exc)[:200]}
        res["per_step"] = per_step

    if with_e2e:
        Ke = min(K, 300)
        rng = np.random.default_rng(7 + rank)
        host_actions = rng.integers(0, 4, size=(Ke + 3, N), dtype=np.int32)
        for k in range(3):
            env.step_host(host_actions[k])
        D.barrier()
        t0 = time.perf_counter()
        for k in range(Ke):
            env.step_host(host_actions[3 + k])
        torch.cuda.synchronize()
        te = D.max_over_ranks(time.perf_counter() - t0)
        obs_bytes = N * W * H * 3
        if env.host_threads:  # packed path: one record of ceil(e * W*H*3 / 32) words per e envs (e = 2 if W <= 16 else 1)
            epw = 2 if W <= 16 else 1
            d2h_obs = ((N + epw - 1) // epw) * ((epw * W * H * 3 + 31) // 32) * 4
        else:
            d2h_obs = obs_bytes
        res["e2e"] = {"value": world * N * Ke / te, "unit": "env-steps/s", "h2d_bytes_per_step": N * 4,
                      "d2h_bytes_per_step": d2h_obs + N * 8 + N, "obs_bytes_delivered_per_step": obs_bytes,
                      "steps": Ke, "us_per_step": te * 1e6 / Ke,
                      "host_threads": env.host_threads,
                      "api": ("wf_step_host, page-locked host buffers: actions/reward/done zero-copy; observation sent as a bit "
                              "stream into mapped host memory and expanded to the uint8 array by the library's host threads"
                              if env.host_threads else
                              "wf_step_host, page-locked host buffers (actions/reward/done zero-copy, obs one DMA copy)")}
    res["stats"] = env.stats()
    env.close()
    del obs_buf, rew_buf, done_buf, out, env
    torch.cuda.empty_cache()
    return res


def measure_policy(D: Dist, name: str, K: int, Wm: int):
    """configs[2]: the reference's DQN network picks the actions (DQN.py:188-233), one wf_step per step.
    The whole step (obs -> float -> MLP -> epsilon-greedy -> wf_step) is captured in a CUDA graph."""
    torch = D.torch
    from wildfire_control_python_b200 import BatchedForestFire
    wl = WORKLOADS[name]
    dev, rank, world = D.dev, D.rank, D.world
    N, W, H = wl["n_envs"], wl["meta"]["width"], wl["meta"]["height"]
    env = BatchedForestFire(N, device=dev, auto_reset=True, seed=0, env_id_base=rank * N, **wl["meta"])
    obs = env.reset()
    # The policy: a network the REFERENCE trained (Models/14-sized/SARSA9-..., read by keras_h5 without h5py; the
    # file travels as a test fixture), else fixed-seed random weights of the same architecture.
    wpath = os.path.join(ROOT, "tests", "golden", "keras", "SARSA9-14s-10k-47298m-06-24-0808")
    if os.path.isfile(wpath) and (W, H) == (14, 14):
        from wildfire_control_python_b200.keras_h5 import read_keras_weights
        kw = read_keras_weights(wpath)
        w1, b1, w2, b2 = (torch.as_tensor(kw[k]).to(dev) for k in
                          ("dense_1/kernel:0", "dense_1/bias:0", "dense_2/kernel:0", "dense_2/bias:0"))
        policy_name = "the reference's trained DQN_SARSA network SARSA9-14s-10k-47298m (Keras weights)"
    else:
        g = torch.Generator(device=dev).manual_seed(1234)  # same weights on every rank
        w1 = torch.randn(W * H * 3, 50, device=dev, generator=g) * 0.05
        b1 = torch.zeros(50, device=dev)
        w2 = torch.randn(50, env.n_actions, device=dev, generator=g) * 0.05
        b2 = torch.zeros(env.n_actions, device=dev)
        policy_name = "random-weight network"
    actions = torch.zeros(N, dtype=torch.int32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(77 + rank)

    def one_step():
        q = torch.sigmoid(obs.view(N, -1).float() @ w1 + b1) @ w2 + b2
        greedy = q.argmax(1).to(torch.int32)
        explore = torch.rand(N, device=dev) < 0.1
        rnd = torch.randint(0, env.n_actions, (N,), device=dev, dtype=torch.int32)
        actions.copy_(torch.where(explore, rnd, greedy))
        env.step(actions)  # writes the handle's persistent obs buffer, which `obs` aliases

    for _ in range(3):
        one_step()
    G_STEPS = 16
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        one_step()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(G_STEPS):
                one_step()
    torch.cuda.current_stream(dev).wait_stream(side)
    reps_w, reps = max(1, Wm // G_STEPS), max(1, K // G_STEPS)
    for _ in range(reps_w):
        graph.replay()
    D.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        graph.replay()
    ev1.record()
    D.barrier()
    ms_torch = D.max_over_ranks(ev0.elapsed_time(ev1))
    Kt_torch = reps * G_STEPS
    via_torch = {"value": world * N * Kt_torch / (ms_torch * 1e-3), "unit": "env-steps/s", "steps": Kt_torch,
                 "us_per_step": ms_torch * 1e3 / Kt_torch,
                 "how": "obs -> float -> torch MLP -> eps-greedy -> wf_step, 16 steps per CUDA-graph replay"}
    # The same policy evaluated INSIDE the step kernel (WF_POLICY_MLP): first layer kept incrementally as a sum of
    # weight rows of the observation bits, 64 steps per launch, obs/reward/done of every step still written.
    chunk = 64
    env.set_policy_mlp(w1, b1, w2, b2, eps=0.1)
    out = (torch.empty((chunk, N, W, H, 3), dtype=torch.uint8, device=dev), torch.empty((chunk, N), dtype=torch.float64, device=dev),
           torch.empty((chunk, N), dtype=torch.uint8, device=dev))
    nl = max(1, K // chunk)
    for _ in range(max(1, Wm // chunk) + 2):
        env.rollout(chunk, policy="mlp", out=out)
    D.barrier()
    l0 = env.launch_count
    ev0.record()
    for _ in range(nl):
        env.rollout(chunk, policy="mlp", out=out)
    ev1.record()
    D.barrier()
    ms = D.max_over_ranks(ev0.elapsed_time(ev1))
    Kt = nl * chunk
    res = {"name": name, "wl": wl, "N": N, "W": W, "H": H, "K": Kt, "chunk": chunk, "fused": True, "graph": False, "ms": ms,
           "launches": env.launch_count - l0, "value": world * N * Kt / (ms * 1e-3), "obs_mb": out[0].numel() / 1e6,
           "family": env.kernel_family, "state_bytes": env.state_bytes_per_env, "stats": env.stats(), "via_torch": via_torch,
           "policy": f"MLP 588-50(sigmoid)-4 = {policy_name}, eps-greedy 0.1, evaluated inside the step kernel (WF_POLICY_MLP); "
                     "`via_torch` is the same policy through torch ops"}
    env.close()
    return res


def roofline_of(res, world):
    peak, peak_src = load_peaks()
    N, W, H, fused = res["N"], res["W"], res["H"], res["fused"]
    per_gpu_steps = res["value"] / world
    achieved = per_gpu_steps * W * H * BYTES_PER_CELL_UPDATE / 1e9
    if res["family"] == "warp":  # state stays in registers across a launch: only outputs (+ state once per launch) touch HBM
        own = W * H * 3 + 8 + 1 + 2 * res["state_bytes"] / res["chunk"]
        kernel = "wf::warp_kernel"
        note = (f"state ({N} envs x {res['state_bytes']} B) lives in registers across a launch and in L2 between launches: "
                "this kernel is issue/latency-bound (ncu r01: DRAM 7 %, issue slots 50 %, 2 048 warps for 592 scheduler "
                "slots); the HBM roofline is reported because the contract asks for it")
    else:  # dense pass per step: G, B, S read + S_next written (tick), F, I read + 96 B written (observation) per 32 cells
        own = W * H * (3 * 4 / 32 + 4 / 32 + 2 * 4 / 32 + 3)
        kernel = "wf::tile_rollout_kernel"
        note = ("HBM-bound: one cluster per env streams 6 plane words + 96 B of observation per 32 cells and step; "
                "fuel records / hit counters are touched only where the fire front is.  `achieved` uses SURVEY 8(d)'s 15 B per "
                "cell-update (byte-per-cell state read + written, 3 B observation); the bit-plane layout moves 3.75 B, so "
                "`frac` can exceed 1 -- `frac_own_layout` is the fraction of peak in bytes this layout really moves")
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get(f"{res['name']}_chunk{res['chunk'] if fused else 1}")
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src, "kernel": kernel, "bytes_per_unit": W * H * BYTES_PER_CELL_UPDATE, "unit_name": "env-step",
            "units_per_launch": N * (res["chunk"] if fused else 1),
            "avg_launch_ms": res["ms"] / max(1, res["launches"]) if fused else res["ms"] / res["K"],
            "own_layout_bytes_per_unit": own, "achieved_own_layout_gbs": per_gpu_steps * own / 1e9,
            "frac_own_layout": per_gpu_steps * own / 1e9 / peak, "note": note}


def run_ours(args, wl):
    D = Dist(args.gpus)
    with ClockSampler(D.local) as clk:
        if wl.get("policy"):
            res = measure_policy(D, args.workload, args.steps, args.warmup)
        else:
            res = measure(D, args.workload, args.steps, args.warmup, args.chunk, args.no_graph, not args.only_value,
                          not args.only_value)
        sec = None
        if args.workload == "c2" and not args.no_secondary and not args.only_value:
            sec = measure(D, "c4", 160, 32, 0, args.no_graph, False, False)
    clocks = clk.summary()
    if D.rank != 0:
        D.close()
        return
    world, N, W, H, K = D.world, res["N"], res["W"], res["H"], res["K"]

    cpu_baseline = None
    if world == 1 and not args.only_value:
        from oracle import wf_oracle as wo  # the checker, timed as the CPU baseline (allowed use)
        ob = wo.OracleBatch(oracle_cfg(wl["meta"]), min(N, 1024), 1)
        t0 = time.perf_counter(); n = ob.step(2); rate = n / (time.perf_counter() - t0)
        steps_cpu = max(1, int(rate * CPU_BASELINE_SECONDS / ob.n_envs))
        t0 = time.perf_counter(); n = ob.step(steps_cpu); dt = time.perf_counter() - t0
        cpu_baseline = {"value": n / dt, "unit": "env-steps/s", "cores": 1, "kind": "port",
                        "sample": f"{ob.n_envs} envs x {steps_cpu} steps of the same workload, 1 thread, {dt:.1f} s",
                        "host_cpus": os.cpu_count()}

    line = {
        "metric": "env_steps_per_sec", "value": res["value"], "unit": "env-steps/s", "n_gpus": world, "steps": K,
        "warmup": max(args.warmup, 3), "ms_per_step": res["ms"] / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32-bitplanes+u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "grid": [W, H], "envs_per_gpu": N,
                   "steps_per_launch": res["chunk"] if res["fused"] else None,
                   "steps_per_graph_replay": res["chunk"] if res["graph"] else None, "kernel_family": res["family"],
                   "l2": f"outputs of one launch/replay ({res['obs_mb']:.0f} MB obs) exceed the 126 MB L2; "
                         + ("state is register/L2 resident by design" if res["family"] == "warp"
                            else f"state ({res['state_bytes'] * res['N'] / 1e6:.0f} MB) exceeds L2 too")},
        "cell_updates_per_sec": res["value"] * W * H,
        "cell_updates_per_sec_per_gpu": res["value"] * W * H / world,
        "e2e": res.get("e2e"), "per_step_launch": res.get("per_step"),
        "gpu_launches": res["launches"],
        "roofline": roofline_of(res, world),
        "cpu_baseline": cpu_baseline,
        "clocks": clocks,
        "stats": res["stats"],
    }
    if "policy" in res:
        line["config"]["policy"] = res["policy"]
        line["steps"] = res["K"]
        line["via_torch"] = res["via_torch"]
    if sec is not None:
        line["secondary"] = {"c4": {
            "workload": WORKLOADS["c4"]["desc"], "value": sec["value"], "unit": "env-steps/s", "steps": sec["K"],
            "ms_per_step": sec["ms"] / sec["K"], "cell_updates_per_sec_per_gpu": sec["value"] * sec["W"] * sec["H"] / world,
            "gpu_launches": sec["launches"], "steps_per_graph_replay": sec["chunk"] if sec["graph"] else None,
            "roofline": roofline_of(sec, world)}}
    print(json.dumps(line), flush=True)
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--chunk", type=int, default=0, help="steps per fused launch / graph replay (default: workload's)")
    ap.add_argument("--no-graph", action="store_true", help="tile family: plain launches instead of CUDA-graph replay")
    ap.add_argument("--no-secondary", action="store_true", help="skip the c4 stencil measurement in the default line")
    ap.add_argument("--only-value", action="store_true",
                    help="profiling runs: only the fused-rollout measurement (no per-step, e2e, CPU baseline legs)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
