"""bench.py -- env-steps/s of the B200 stepping engine on the BASELINE.json configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mnist|tinyloc|lidar_rooms|lidar_maze]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...     # the CPU restatement of the reference path on the host cores

A "step" is one vector step: every environment of the batch advances once through the fused kernel
(position update, resample, loss, reward, flags, autoreset).  `value` = env-steps/s with all inputs resident
in HBM (device-timed, max over ranks); `e2e` = the same through the host-buffer C-ABI call with the
host<->device copies inside the timed region.  Prints ONE JSON line (rank 0).
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
sys.path.insert(0, os.path.join(ROOT, "active-perception-gym_b200"))
sys.path.insert(0, ROOT)

# algorithmic bytes per env-step (DESIGN.md "Roofline"; SURVEY.md section 8d with the uint8 bank)
WORKLOADS = {
    # name: (description, per-GPU envs, bytes per env-step)
    "mnist": dict(desc="MNIST-v0 shape: synthetic 60000x28x28x1 u8 bank, ImageClassificationVectorEnv, K=10, S=5, T=16",
                  n_envs=65536, bytes_per_step=300),
    "tinyloc": dict(desc="TinyImageNetLoc-v0 shape: synthetic 100000x64x64x3 u8 bank, ImageLocalizationVectorEnv, S=10, T=16",
                    n_envs=131072, bytes_per_step=1930),
    "lidar_rooms": dict(desc="LIDARLocRooms-v0: 4096 procedural 32x32 rooms maps (reference generator), 64 beams, range 5, T=100",
                        n_envs=262144, bytes_per_step=510),
    "lidar_maze": dict(desc="LIDARLocMaze-v0 large: 256 procedural 63x63 mazes (reference generator), 256 beams, range 5, T=100",
                       n_envs=131072, bytes_per_step=1800),
}
LIDAR = {"lidar_rooms": ("maps_rooms32.npz", 64), "lidar_maze": ("maps_maze63.npz", 256)}
GOLDEN = os.path.join(ROOT, "tests", "golden")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0,
                    help="timed steps (default: per workload, ~0.4 s of device time: mnist 20400, tinyloc 3400, "
                         "lidar_rooms 680, lidar_maze 340)")
    ap.add_argument("--warmup", type=int, default=34)
    ap.add_argument("--workload", default="mnist", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="override per-GPU num_envs")
    ap.add_argument("--bank", type=int, default=0, help="override bank length")
    ap.add_argument("--e2e-steps", type=int, default=68)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt-blend", action="store_true", help="skip the extra float32-blend timing of image workloads")
    ap.add_argument("--blend", default="float64", choices=["float64", "float32"],
                    help="image workloads: float64 = the reference's blend arithmetic (bit-exact pixels, default); "
                         "float32 = opt-in blend within the 1e-5 relative bar")
    args = ap.parse_args()
    if args.steps <= 0:
        args.steps = {"mnist": 20400, "tinyloc": 3400, "lidar_rooms": 680, "lidar_maze": 340}[args.workload]
    return args


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms; stop(t0, t1) keeps the samples that arrived inside the
    timed region [t0, t1] (wall clock) and falls back to all of them, saying so, when the region was too short."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (t, ln) in self.lines if t0 is not None and t0 <= t <= t1 + 0.05]
        in_region = len(inside) > 0
        for ln in (inside if in_region else [ln for (_, ln) in self.lines]):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "sampled": "inside the timed region" if in_region else
                "around the timed region (shorter than the 50 ms sampling period)", "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU side
def cpu_oracle_run(workload: str, n_envs: int, steps: int, seed: int = 0):
    """One process of the CPU restatement (oracle/) on a bounded sample; returns env-steps/s."""
    from oracle.datasets import synthetic_u8_bank
    from oracle.image_oracle import ImageEnvOracle

    if workload == "mnist":
        imgs, labels = synthetic_u8_bank(seed=0, M=4096, H=28, W=28, C=1, K=10)
        env = ImageEnvOracle(imgs, labels, n_envs, task="cls", num_classes=10, track_overlay=False)
        pw = 10
    elif workload == "tinyloc":
        imgs, labels = synthetic_u8_bank(seed=0, M=512, H=64, W=64, C=3, K=200)
        env = ImageEnvOracle(imgs, labels, n_envs, task="loc", num_classes=200, sensor_size=(10, 10), track_overlay=False)
        pw = 2
    elif workload in LIDAR:
        return cpu_lidar_run(workload, n_envs, steps, seed)
    else:
        raise ValueError(workload)
    rng = np.random.default_rng(seed)
    if env.task == "loc":
        env.reset(seed=seed, explicit_targets=rng.uniform(-1, 1, (n_envs, 2)).astype(np.float32))
    else:
        env.reset(seed=seed)
    a = rng.uniform(-1, 1, (steps, n_envs, 2)).astype(np.float32)
    p = rng.standard_normal((steps, n_envs, pw)).astype(np.float32)
    for s in range(3):
        env.step(a[s], p[s])
    t0 = time.perf_counter()
    for s in range(steps):
        env.step(a[s], p[s])
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def cpu_lidar_run(workload: str, n_envs: int, steps: int, seed: int = 0):
    from oracle.lidar_oracle import LidarVecOracle

    fname, beams = LIDAR[workload]
    z = np.load(os.path.join(GOLDEN, fname))
    W = int(z["width"])
    x = np.arange(W)
    maps = ((z["bits"][:64, :, x >> 5] >> (x & 31).astype(np.uint32)) & 1).astype(bool)
    env = LidarVecOracle(maps, n_envs, beam_count=beams)
    env.reset(seed=seed)
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, (steps + 3, n_envs, 2)).astype(np.float32)
    p = rng.uniform(-1, 1, (steps + 3, n_envs, 2)).astype(np.float32)
    for s in range(3):
        env.step(a[s], p[s])
    t0 = time.perf_counter()
    for s in range(3, steps + 3):
        env.step(a[s], p[s])
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def _cpu_worker(args):
    return cpu_oracle_run(*args)


def cpu_baseline(workload: str, procs: int, n_envs: int, steps: int):
    if procs <= 1:
        v, dt = cpu_oracle_run(workload, n_envs, steps)
        return v, dt
    import multiprocessing as mp

    with mp.get_context("spawn").Pool(procs) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(workload, n_envs, steps, i) for i in range(procs)])
        wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    n_envs = {"mnist": 4096, "tinyloc": 256, "lidar_rooms": 1024, "lidar_maze": 128}[args.workload]
    steps_per = 17
    vals = []
    for _ in range(max(1, args.warmup // 17)):
        cpu_baseline(args.workload, procs, n_envs, 4)
    t0 = time.perf_counter()
    reps = max(1, min(args.steps // 17, 6))
    for _ in range(reps):
        v, _ = cpu_baseline(args.workload, procs, n_envs, steps_per)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.median(vals))
    sample = f"{procs} processes x {n_envs} envs x {steps_per} steps of the CPU restatement (oracle/), median of {reps}"
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_envs * procs / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "num_envs_per_gpu": w["n_envs"]},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU side
def build_env(args, rank, world, device):
    import ap_gym_b200 as apg

    w = WORKLOADS[args.workload]
    n = args.envs or w["n_envs"]
    if args.workload == "mnist":
        ds = apg.SyntheticImageDataset("mnist", length=args.bank or None, seed=0)
        cfg = apg.ImagePerceptionConfig(dataset=ds)
        env = apg.ImageClassificationVectorEnv(n, cfg, device=device, env_offset=rank * n, global_num_envs=world * n,
                                               check_errors="never", payload_ring=64, blend=args.blend)
        pw = 10
    elif args.workload in LIDAR:
        fname, beams = LIDAR[args.workload]
        bank = apg.FloorMapBank.load_npz(os.path.join(GOLDEN, fname), args.bank or None)
        env = apg.LIDARLocalization2DVectorEnv(n, bank, lidar_beam_count=beams, device=device, env_offset=rank * n,
                                               check_errors="never")
        pw = 2
    else:
        ds = apg.SyntheticImageDataset("tinyimagenet", length=args.bank or None, seed=0)
        cfg = apg.ImagePerceptionConfig(dataset=ds, sensor_size=(10, 10))
        env = apg.ImageLocalizationVectorEnv(n, cfg, device=device, env_offset=rank * n, global_num_envs=world * n,
                                             check_errors="never", payload_ring=64, blend=args.blend)
        pw = 2
    return env, n, pw


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from ap_gym_b200 import _capi

    w = WORKLOADS[args.workload]
    env, n, pw = build_env(args, rank, world, local_rank)
    is_lidar = args.workload in LIDAR
    T = 16
    K, W = args.steps, args.warmup
    dev = torch.device("cuda", local_rank)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    # distinct inputs for every step: streamed from HBM, total larger than L2
    chunk = 17 * 3  # steps per rollout call (3 episodes)
    pool = min(K + W, 7 * chunk)  # distinct per-step inputs; longer runs walk the pool cyclically (it is >> L2)
    acts = torch.rand((pool, n, 2), generator=gen, device=dev) * 2 - 1
    if pw == 2:
        preds = torch.rand((pool, n, 2), generator=gen, device=dev) * 2 - 1
    else:
        preds = torch.randn((pool, n, pw), generator=gen, device=dev)
    t_reset0 = time.perf_counter()
    env.reset(seed=0)
    env.sync()
    reset_s = time.perf_counter() - t_reset0

    def run(s0, s1, env=env):
        s = s0
        while s < s1:
            a = s % pool
            m = min(chunk, s1 - s, pool - a)
            env.rollout(acts[a:a + m], preds[a:a + m])
            s += m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up.  Autoreset payloads are NOT staged ahead: the engine draws each one on the device, in-stream, right
    # before the step that consumes it (numpy-identical streams, csrc/apg_payload.cu), inside the timed region.
    run(0, W)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _capi.lib().apg_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record()
    run(W, W + K)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = _capi.lib().apg_launch_count() - l0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * n * K / (ms_max * 1e-3)

    # ---- image workloads: the same timed region with the opt-in float32 blend (pixels within 1e-5 relative of the
    # reference instead of bit-exact; tests/test_image_gpu.py holds both modes to their bars)
    alt = None
    if not is_lidar and args.blend == "float64" and not args.no_alt_blend:
        import copy
        a2 = copy.copy(args)
        a2.blend = "float32"
        env2, _, _ = build_env(a2, rank, world, local_rank)
        env2.reset(seed=0)
        run(0, W, env2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(W, W + K, env2)
        e1.record()
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms2 = float(t2.item())
        alt = {"blend": "float32", "value": world * n * K / (ms2 * 1e-3), "ms_per_step": ms2 / K,
               "parity": "glimpse pixels within 1e-5 relative of the reference (no absolute slack), everything else bit-exact"}
        env2.close()
        del env2

    # ---- e2e: host buffers through the C-ABI call, copies inside the timed region
    Ke = args.e2e_steps
    h_a = torch.empty((n, 2), dtype=torch.float32).pin_memory()
    h_p = torch.empty((n, pw), dtype=torch.float32).pin_memory()
    h_r = torch.empty((n,), dtype=torch.float32).pin_memory()
    h_te = torch.empty((n,), dtype=torch.uint8).pin_memory()
    h_tr = torch.empty((n,), dtype=torch.uint8).pin_memory()
    h_a.copy_(acts[0].cpu())
    h_p.copy_(preds[0].cpu())
    lib = _capi.lib()

    def e2e_step():
        if is_lidar:
            _capi.check(lib.apg_lidar_step_host(env._h, h_a.data_ptr(), h_p.data_ptr(), h_r.data_ptr(), h_te.data_ptr(),
                                                h_tr.data_ptr(), None))
            return
        _capi.check(lib.apg_image_step_host(env._h, h_a.data_ptr(), h_p.data_ptr(), h_r.data_ptr(), h_te.data_ptr(),
                                            h_tr.data_ptr(), None))

    for _ in range(5):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * n * Ke / float(t_e.item())
    h2d = n * (2 + pw) * 4
    d2h = n * (4 + 1 + 1)

    # ---- final NCCL reduce of the device-side episode statistics (the only collective of the data path)
    stats = env.episode_stats(reduce=True)
    env.sync()
    # sticky device-side error bits (NaN input, tap outside the image grid) of everything this process stepped: the
    # benchmark runs with check_errors="never", so look once at the end -- a non-zero value invalidates the run
    import ctypes
    flags = ctypes.c_uint32(0)
    _capi.check((lib.apg_lidar_poll_error if is_lidar else lib.apg_image_poll_error)(env._h, ctypes.byref(flags), 0))
    t_f = torch.tensor([int(flags.value)], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t_f, op=dist.ReduceOp.MAX)
    error_flags = int(t_f.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        try:  # dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if args.workload in tj and n == w["n_envs"]:
                traffic = tj[args.workload]["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        bytes_per_launch = w["bytes_per_step"] * n
        avg_launch_s = (ms_max * 1e-3) / K
        achieved = bytes_per_launch / avg_launch_s / 1e9
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "num_envs_per_gpu": n, "global_num_envs": world * n, "sharding": f"env-index x{world}", "blend": None if is_lidar else args.blend, "explicit_reset_s": round(reset_s, 3),
                       "l2": "per-step inputs come from a pool of %d distinct steps streamed from HBM (%.0f MB > 126 MB L2, walked cyclically); engine-owned obs buffers are rewritten every step" % (pool, acts.numel() * 4 / 1e6 + preds.numel() * 4 / 1e6)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "api": ("apg_lidar_step_host" if is_lidar else "apg_image_step_host") + " (pinned host action/prediction in, reward/terminated/truncated out)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                         "kernel": "lidar_step_kernel" if is_lidar else
                         ["image_step_kernel (warp blend)", "image_step_kernel (thread blend)",
                          "image_step_pipelined_kernel"][lib.apg_image_variant(env._h)], "bytes_per_env_step": w["bytes_per_step"],
                         "avg_launch_us": avg_launch_s * 1e6, "launches_per_step": launches / K,
                         "note": "achieved = algorithmic bytes of one step / device time of one step; with 2 launches per "
                                 "step (image_head_kernel + image_step_kernel, warp path) the step time covers both"},
            "episode_stats_allreduce": stats,
            "error_flags": error_flags,
        }
        if error_flags:
            sys.stderr.write("bench.py: device error flags 0x%x were raised during the run -- result INVALID\n" % error_flags)
        if alt is not None:
            alt["roofline_frac"] = (bytes_per_launch / (alt["ms_per_step"] * 1e-3) / 1e9) / peak
            line["alt_blend"] = alt
        if not args.no_cpu_baseline and world == 1:
            n_cpu, st_cpu = {"mnist": (4096, 680), "tinyloc": (256, 1020), "lidar_rooms": (2048, 303), "lidar_maze": (256, 808)}[args.workload]
            v, dt = cpu_oracle_run(args.workload, n_cpu, st_cpu)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": 1, "kind": "port",
                                    "sample": f"CPU restatement (oracle/), 1 process, N={n_cpu} x {st_cpu} steps, {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
