#!/usr/bin/env python
"""bench.py — env-steps/s of the batched `World` step with layered observations (BASELINE.json metric).

A "step" is one lockstep step of the hot path over one batch: 65,536 x `World.level(6)` (4 agents, 3 laser
sources, 4 gems) per GPU with device-sampled (Philox) random actions, auto-reset, and every per-step output
of `LLE.step` written: layered observation, state vector, availability mask, reward, done, events.

  python bench.py [--gpus N] [--steps K] [--warmup W]             our arm (CUDA, through the C ABI)
  python bench.py --impl reference [--gpus N] [--steps K] ...     the CPU arm: the oracle's restatement of the
        reference engine on all host cores (the Rust crate cannot be built in this image: no cargo/rustc)

For N > 1 launch with torchrun (one rank per GPU); envs are range-sharded over ranks (weak scaling, no
collective on the step path; one NCCL all-reduce(MAX) of the elapsed time and one of the episode counter).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LEVEL = 6
ENVS_PER_GPU = 65536
SEED = 2026
METRIC = "env-steps/s (layered obs, device-timed)"
UNIT = "env-steps/s"


def level_text(n: int) -> str:
    with open(os.path.join(ROOT, "lle_b200", "resources", "levels", f"lvl{n}")) as f:
        return f.read()


def algorithmic_bytes(A: int, G: int, C: int, H: int, W: int, R: int, record_bytes: int) -> dict:
    """B = OBS + ST + AV + AC + RDE + 2*S (bytes that must cross HBM per env-step)."""
    obs = 4 * C * H * W
    st = 4 * (3 * A + G)
    av = 5 * A
    ac = A
    rde = 4 * R + 1 + A + 1  # reward, done, events, err
    return dict(obs=obs, state=st, avail=av, actions=ac, reward_done_events=rde, record_rw=2 * record_bytes,
                total=obs + st + av + ac + rde + 2 * record_bytes)


def measured_peak() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram bytes per launch of the step kernel from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_baseline(sample_seconds: float = 12.0, n_envs: int = 4096, threads: int = 0) -> dict:
    """The oracle (CPU restatement of the reference engine + layered observation) on the host cores:
    one World per thread at a time, all outputs written each step, Philox actions, auto-reset."""
    from oracle import lle_oracle as lo

    vec = lo.OracleVec([level_text(LEVEL)], None, n_envs, seed=SEED, auto_reset=True)
    secs, used = vec.rollout(8, threads)  # warm-up + calibration
    rate = n_envs * 8 / max(secs, 1e-9)
    steps = max(8, int(sample_seconds * rate / n_envs))
    secs, used = vec.rollout(steps, threads)
    value = n_envs * steps / secs
    return {"value": value, "unit": UNIT, "cores": used, "kind": "port",
            "sample": f"{n_envs} envs x {steps} steps of level {LEVEL} in {secs:.2f} s, C++ restatement of the reference engine "
                      f"(Rust toolchain unavailable), one World per thread, {used} threads, layered obs + state + avail written each step",
            "seconds": secs}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" of this arm is a bounded sample of the same workload; K samples are timed
    k = max(1, min(args.steps, 5))
    w = 1 if args.warmup > 0 else 0
    t0 = time.time()
    for _ in range(w):
        cpu_baseline(sample_seconds=2.0)
    res = [cpu_baseline(sample_seconds=6.0) for _ in range(k)]
    value = statistics.mean(r["value"] for r in res)
    ms = 1000.0 * statistics.mean(r["seconds"] for r in res)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": k, "warmup": w,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 bitmasks, f32 encodings",
        "data": "synthetic",
        "config": {"workload": f"World.level({LEVEL}) 4 agents, 3 laser sources (BASELINE.json says 4), layered observations, "
                               f"random Philox actions, auto-reset; CPU arm: bounded samples of 4096 envs"},
        "cpu_baseline": {**{k2: v for k2, v in res[-1].items() if k2 != "seconds"}, "value": value},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line))


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import lle_b200

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world_size > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner (and any warning) to stdout unless told otherwise
        # (NCCL_DEBUG_FILE is honoured only above the VERSION level)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    from lle_b200.sharding import reduce_stats, shard_range

    begin, end = shard_range(args.envs * world_size, rank, world_size)  # weak scaling: args.envs worlds per GPU
    n_envs = end - begin
    vec = lle_b200.VecWorld(lle_b200.Map(level=LEVEL), n_envs, device=dev, seed=SEED, env_id_base=begin, auto_reset=True)
    A, G, C, H, W, R = vec.n_agents, vec.n_gems, vec.n_channels, vec.height, vec.width, vec.reward_dim
    K, Wm = args.steps, max(args.warmup, 3)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm: K fused steps back to back, no host sync inside
    for _ in range(Wm):
        vec.step(None)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = vec.launch_count
    vec.timing_begin()
    for _ in range(K):
        vec.step(None)
    ms_total, timed_launches = vec.timing_end()
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = vec.launch_count - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world_size * n_envs * K / (ms_max / 1e3)

    # ---------------- the same K steps as lle_vec_rollout launches (128 steps per launch, bit-identical results): reported
    # beside the headline as "rollout", not as the headline (the contract's step = one launch)
    RL = 128
    n_roll = max(1, K // RL)
    vec.rollout(RL)
    barrier()
    vec.timing_begin()
    for _ in range(n_roll):
        vec.rollout(RL)
    ms_roll, roll_launches = vec.timing_end()
    barrier()
    tr = torch.tensor([ms_roll], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
    rollout_value = world_size * n_envs * n_roll * RL / (float(tr.item()) / 1e3)

    # ---------------- end-to-end arm: host actions in (pinned, H2D), reward + done out (D2H), sync every step
    Ke = min(K, args.e2e_steps)
    vec.reset()
    vec.step_count = 10_000_000
    rec = torch.empty((Ke, n_envs, A), dtype=torch.int8).pin_memory()
    for s in range(Ke):  # record a valid action stream on the device, then replay it from the host
        vec.step(None)
        rec[s].copy_(vec.actions, non_blocking=True)
    torch.cuda.synchronize()
    vec.reset()
    vec.step_count = 10_000_000
    D = max(1, min(8, args.e2e_depth))
    reward_h = [torch.empty((n_envs, R), dtype=torch.float32).pin_memory() for _ in range(D)]
    done_h = [torch.empty((n_envs,), dtype=torch.uint8).pin_memory() for _ in range(D)]

    def e2e_sync() -> float:
        """lle_vec_step_host: H2D, step, D2H, stream sync — one call per step, nothing overlapped."""
        vec.reset()
        vec.step_count = 10_000_000
        barrier()
        t0 = time.perf_counter()
        n_done = 0
        for s in range(Ke):
            vec.step_host(rec[s], reward_h[0], done_h[0])
            n_done += int(done_h[0][0])  # the host consumes the result
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def e2e_pipelined() -> float:
        """lle_vec_pipeline_submit / _wait with D steps in flight: every step still takes its actions from pinned host
        memory and lands its reward + done in pinned host memory; the copies overlap the neighbouring steps' kernels."""
        vec.reset()
        vec.step_count = 10_000_000
        barrier()
        t0 = time.perf_counter()
        n_done = 0
        for s in range(Ke):
            if s >= D:
                vec.wait_host()
                n_done += int(done_h[(s - D) % D][0])  # the host consumes the result of step s - D
            vec.submit_host(rec[s], reward_h[s % D], done_h[s % D])
        for s in range(max(Ke - D, 0), Ke):
            vec.wait_host()
            n_done += int(done_h[s % D][0])
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def reduce_max(x: float) -> float:
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    for s in range(min(3, Ke)):
        vec.step_host(rec[s], reward_h[0], done_h[0])
    e2e_pipelined()  # warm-up (creates the pipeline's streams and ring)
    e2e_sync_value = world_size * n_envs * Ke / reduce_max(e2e_sync())
    e2e_value = world_size * n_envs * Ke / reduce_max(e2e_pipelined())
    assert int(vec.err.sum()) == 0, "replayed actions must be valid"

    # end-of-run stats reduction: the only collective on this path (NCCL all-reduce of a few counters)
    stats = reduce_stats(torch.stack([vec.done.sum().to(torch.int64), vec.reward.sum().to(torch.int64),
                                      torch.tensor(n_envs, dtype=torch.int64, device=dev)]))

    if rank == 0:
        bytes_env = algorithmic_bytes(A, G, C, H, W, R, record_bytes=vec.record_bytes)
        peak, peak_src = measured_peak()
        kernel_ms = ms_total / max(timed_launches, 1)
        achieved = bytes_env["total"] * n_envs / (kernel_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": K, "warmup": Wm,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u64 bitmasks, f32 encodings", "data": "synthetic",
            "config": {
                "workload": f"World.level({LEVEL}) 4 agents, 3 laser sources (BASELINE.json says 4; the level file has 3), "
                            f"{n_envs} batched envs per GPU with layered observations (BASELINE.json configs[1])",
                "envs_per_gpu": n_envs, "agents": A, "obs_shape": [C, H, W], "actions": "device Philox4x32-10, uniform over available",
                "auto_reset": True, "agent_env_steps_per_s": value * A,
                "l2": f"each step rewrites {n_envs * C * H * W * 4 / 1e6:.0f} MB of observations per GPU (> 126 MB L2); no flush needed",
                "sharding": "contiguous env ranges per rank, no collective on the step path",
                "launch": "one fused kernel launch per step (programmatic dependent launch between steps); "
                          "lle_vec_rollout(K) runs K steps in one launch with identical results",
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_envs * A, "d2h_bytes_per_step": n_envs * (4 * R + 1),
                    "steps": Ke, "pipeline_depth": D, "sync_value": e2e_sync_value,
                    "note": f"lle_vec_pipeline_submit/_wait, {D} steps in flight: every step copies its actions from pinned host memory "
                            "(H2D), runs the fused step and copies reward+done to pinned host memory (D2H), all inside the timed "
                            "region; the host reads each step's result. sync_value = lle_vec_step_host (same copies, stream sync "
                            "after every step, nothing overlapped). Observations stay in HBM (zero-copy DLPack hand-off)"},
            "gpu_launches": launches,
            "rollout": {"value": rollout_value, "unit": UNIT, "steps_per_launch": RL, "launches": int(roll_launches),
                        "ms_per_step": float(tr.item()) / (n_roll * RL),
                        "note": "lle_vec_rollout(128): the same steps, 128 per launch, ordered by the per-ticket epoch flags; device-timed"},
            "stats_allreduce": {"done_last_step": int(stats[0]), "reward_last_step": int(stats[1]), "envs_total": int(stats[2])},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_per_launch(), "peak_source": peak_src, "kernel": "lle_world_kernel<MODE_STEP, FAST>",
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_env_step": bytes_env,
                         "algorithmic_bytes_per_launch": bytes_env["total"] * n_envs},
        }
        if world_size == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline()
            cb.pop("seconds", None)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8192)
    ap.add_argument("--warmup", type=int, default=128)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--e2e-steps", type=int, default=2048)
    ap.add_argument("--e2e-depth", type=int, default=8, help="steps in flight in the pipelined end-to-end arm (1..8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
