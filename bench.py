"""bench.py -- env-steps/s of the B200 stepping engine on the BASELINE.json configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mnist|tinyloc|lidar_rooms|lidar_maze] [--only]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...     # the CPU restatement of the reference path on the host cores

A "step" is one vector step: every environment of the batch advances once through the fused kernels (position
update, resample, loss, reward, flags, autoreset).  `value` = env-steps/s with all inputs resident in HBM
(device-timed with CUDA events, max over ranks); `e2e` = the same through the host-buffer C-ABI call with the
host<->device copies inside the timed region.  Prints ONE JSON line (rank 0): the headline workload (default `mnist` =
BASELINE.json configs[1]) at the top level and every workload of BASELINE.json's configs 2-5 under `workloads`, each
measured under the same --steps/--warmup discipline (`--only` keeps the headline workload alone).

Before the W warm-up steps every workload runs an UNTIMED priming phase (at least two whole episode cycles and at
least --prime-ms of device work): first-use costs (module loads, the first autoreset payload) and the clock ramp of
an idle GPU are not part of a step, and the driver's --warmup 5 is shorter than one 17-step image episode cycle.
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

# algorithmic bytes per env-step (SURVEY.md section 8d with the uint8 bank), frozen
WORKLOADS = {
    "mnist": dict(desc="MNIST-v0 shape: synthetic 60000x28x28x1 u8 bank, ImageClassificationVectorEnv, K=10, S=5, T=16",
                  n_envs=65536, bytes_per_step=300, cycle=17, default_steps=20400),
    "tinyloc": dict(desc="TinyImageNetLoc-v0 shape: synthetic 100000x64x64x3 u8 bank, ImageLocalizationVectorEnv, S=10, T=16",
                    n_envs=131072, bytes_per_step=1930, cycle=17, default_steps=6800),
    "lidar_rooms": dict(desc="LIDARLocRooms-v0: 4096 procedural 32x32 rooms maps (reference generator), 64 beams, range 5, T=100",
                        n_envs=262144, bytes_per_step=510, cycle=101, default_steps=808),
    "lidar_maze": dict(desc="LIDARLocMaze-v0 large: 256 procedural 63x63 mazes (reference generator), 256 beams, range 5, T=100",
                       n_envs=131072, bytes_per_step=1800, cycle=101, default_steps=404),
}
ORDER = ["mnist", "tinyloc", "lidar_rooms", "lidar_maze"]
LIDAR = {"lidar_rooms": ("maps_rooms32.npz", 64), "lidar_maze": ("maps_maze63.npz", 256)}
GOLDEN = os.path.join(ROOT, "tests", "golden")
CPU_SAMPLE = {"mnist": (4096, 340), "tinyloc": (256, 510), "lidar_rooms": (2048, 101), "lidar_maze": (256, 202)}
REF_ARM_ENVS = {"mnist": 4096, "tinyloc": 256, "lidar_rooms": 1024, "lidar_maze": 128}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0,
                    help="timed steps (default: per workload, ~0.3 s of device time)")
    ap.add_argument("--warmup", type=int, default=34)
    ap.add_argument("--workload", default="mnist", choices=sorted(WORKLOADS), help="headline workload of the line")
    ap.add_argument("--only", action="store_true", help="measure the headline workload only (no `workloads` table)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="override per-GPU num_envs")
    ap.add_argument("--bank", type=int, default=0, help="override bank length")
    ap.add_argument("--e2e-steps", type=int, default=34)
    ap.add_argument("--prime-ms", type=float, default=60.0, help="minimum device time of the untimed priming phase")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lidar-maps", default="bank", choices=["bank", "procedural"],
                    help="LIDAR workloads: bank = a finite bank of reference-generated maps (tests/golden), procedural = "
                         "the reference's datasets of len 2**32, every reset generates its map on the device")
    ap.add_argument("--blend", default="float64", choices=["float64", "float32"],
                    help="image workloads: float64 = the reference's blend arithmetic (bit-exact pixels, default); "
                         "float32 = opt-in blend within the 1e-5 relative bar")
    return ap.parse_args()


def workload_config(name: str, n: int, world: int, blend: str | None, lidar_maps: str = "bank"):
    """The `config` object of a line -- the SAME keys in the GPU arm and in the reference arm."""
    w = WORKLOADS[name]
    return {"workload": w["desc"], "name": name, "num_envs_per_gpu": n, "global_num_envs": world * n,
            "sharding": f"env-index x{world}", "blend": None if name in LIDAR else blend,
            "maps": lidar_maps if name in LIDAR else None}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms for the whole run; window(t0, t1) summarises the samples
    that arrived inside [t0, t1] (wall clock) and falls back to the nearest ones, saying so, when the region was shorter
    than the sampling period."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

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

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        deadline = time.time() + 0.3
        while time.time() < deadline and not any(t >= t1 for (t, _) in self.lines):
            time.sleep(0.02)
        inside = [ln for (t, ln) in self.lines if t0 <= t <= t1 + 0.06]
        how = "inside the timed region"
        if not inside:
            near = sorted(self.lines, key=lambda x: min(abs(x[0] - t0), abs(x[0] - t1)))[:3]
            inside = [ln for (_, ln) in near]
            how = "the 3 samples nearest to the timed region (shorter than the 50 ms sampling period)"
        sm, mx, reasons = [], [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(self.NAMES, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "sampled": how, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU side
def cpu_oracle_run(workload: str, n_envs: int, steps: int, seed: int = 0):
    """One process of the CPU restatement (oracle/) on a bounded sample; returns (env-steps/s, seconds)."""
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


def cpu_pool_run(pool, workload: str, procs: int, n_envs: int, steps: int):
    """`procs` independent processes of the restatement, each with its own `n_envs` environments; the rate is all
    env-steps over the wall time of the slowest worker."""
    t0 = time.perf_counter()
    pool.map(_cpu_worker, [(workload, n_envs, steps, i) for i in range(procs)])
    wall = time.perf_counter() - t0
    return procs * n_envs * steps / wall, wall


def real_reference_figure(workload: str):
    """env-steps/s of the UNMODIFIED reference (/root/reference through the test-only gymnasium shim, one process -- the
    reference's vector env is single-threaded by design).  Live where /root/reference exists (the build container),
    else the figure recorded there (profiles/r02_real_reference.json), labelled as such; LIDAR has none (no GEOS)."""
    rec = None
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r02_real_reference.json"))).get(workload)
    except (OSError, ValueError):
        pass
    if not os.path.isdir("/root/reference/ap_gym") or workload in LIDAR:
        if rec is not None:
            rec = dict(rec, source="recorded in the build container (profiles/r02_real_reference.json); "
                                   "/root/reference does not exist on the GPU box")
        return rec
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_real_reference.py"), workload],
                           capture_output=True, text=True, timeout=600)
        out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        out["source"] = "live: unmodified /root/reference under oracle/shim, 1 process"
        return out
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": repr(exc)[:200]} if rec is None else rec


def run_reference_arm(args):
    """The reference's CPU implementation of the path on the host cores: the oracle port on all cores, each timed step =
    one vector step of every worker's bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    name = args.workload
    w = WORKLOADS[name]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    n_envs = REF_ARM_ENVS[name]
    K = args.steps if args.steps > 0 else 2 * w["cycle"]
    W = max(args.warmup, 0)
    with mp.get_context("spawn").Pool(procs) as pool:
        if W:
            cpu_pool_run(pool, name, procs, n_envs, W)
        value, wall = cpu_pool_run(pool, name, procs, n_envs, K)
    sample = (f"{procs} processes x {n_envs} envs x {K} steps of the CPU restatement (oracle/); every process also "
              f"pays its own setup + 3 untimed steps inside the wall time")
    cpu = {"value": value, "unit": "env-steps/s", "cores": procs, "kind": "port", "sample": sample}
    rr = real_reference_figure(name)
    if rr is not None:
        cpu["real_reference"] = rr
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * wall / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(name, args.envs or w["n_envs"], int(os.environ.get("WORLD_SIZE", "1")), args.blend,
                                  args.lidar_maps),
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU side
def build_env(name, args, rank, world, device):
    import ap_gym_b200 as apg

    w = WORKLOADS[name]
    n = args.envs or w["n_envs"]
    if name == "mnist":
        ds = apg.SyntheticImageDataset("mnist", length=args.bank or None, seed=0)
        cfg = apg.ImagePerceptionConfig(dataset=ds)
        env = apg.ImageClassificationVectorEnv(n, cfg, device=device, env_offset=rank * n, global_num_envs=world * n,
                                               check_errors="never", payload_ring=8, blend=args.blend)
        pw = 10
    elif name in LIDAR:
        fname, beams = LIDAR[name]
        if args.lidar_maps == "procedural":  # the reference's datasets of len 2**32: maps generated on the device per reset
            bank = apg.FloorMapDatasetRooms(device=device) if name == "lidar_rooms" else \
                apg.FloorMapDatasetMaze(63, 63, device=device)
        else:
            bank = apg.FloorMapBank.load_npz(os.path.join(GOLDEN, fname), args.bank or None)
        env = apg.LIDARLocalization2DVectorEnv(n, bank, lidar_beam_count=beams, device=device, env_offset=rank * n,
                                               check_errors="never")
        pw = 2
    else:
        ds = apg.SyntheticImageDataset("tinyimagenet", length=args.bank or None, seed=0)
        cfg = apg.ImagePerceptionConfig(dataset=ds, sensor_size=(10, 10))
        env = apg.ImageLocalizationVectorEnv(n, cfg, device=device, env_offset=rank * n, global_num_envs=world * n,
                                             check_errors="never", payload_ring=8, blend=args.blend)
        pw = 2
    return env, n, pw


def kernel_facts(name):
    """Per-kernel facts measured under ncu and committed under profiles/ (DRAM traffic per launch, warp-instructions per
    env-step): read, not measured, here."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_kernel_facts.json"))).get(name, {})
    except (OSError, ValueError):
        return {}


def run_workload(name, args, rank, world, local_rank, dist, sampler, peaks):
    import torch

    from ap_gym_b200 import _capi

    w = WORKLOADS[name]
    is_lidar = name in LIDAR
    env, n, pw = build_env(name, args, rank, world, local_rank)
    lib = _capi.lib()
    K = args.steps if args.steps > 0 else w["default_steps"]
    W = args.warmup
    dev = torch.device("cuda", local_rank)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    # distinct inputs for every step, streamed from HBM (the pool is much larger than L2 and walked cyclically)
    chunk = 51 if not is_lidar else 101  # steps per rollout call
    pool = min(K + W + 2 * w["cycle"], 7 * 51)
    acts = torch.rand((pool, n, 2), generator=gen, device=dev) * 2 - 1
    if pw == 2:
        preds = torch.rand((pool, n, 2), generator=gen, device=dev) * 2 - 1
    else:
        preds = torch.randn((pool, n, pw), generator=gen, device=dev)
    t_reset0 = time.perf_counter()
    env.reset(seed=0)
    env.sync()
    reset_s = time.perf_counter() - t_reset0
    cursor = [0]

    def run(steps):
        done = 0
        while done < steps:
            a = cursor[0] % pool
            m = min(chunk, steps - done, pool - a)
            env.rollout(acts[a:a + m], preds[a:a + m])
            cursor[0] += m
            done += m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---- priming (untimed): whole episode cycles until at least --prime-ms of device time has gone by.  Autoreset
    # payloads are NOT staged ahead anywhere in this file: the engine draws them on the device, in-stream
    # (numpy-identical streams) -- each rollout launch is followed by the draws that replace the payloads it consumed,
    # so every timed window draws as many payloads as it uses.
    prime_steps = 0
    p0, p1 = ev(), ev()
    p0.record()
    while True:
        run(2 * w["cycle"])
        prime_steps += 2 * w["cycle"]
        p1.record()
        p1.synchronize()
        if p0.elapsed_time(p1) >= args.prime_ms or prime_steps >= 200 * w["cycle"]:
            break
    # one untimed rehearsal of the timed window (barrier, K steps, barrier): first-use costs of the barrier -> launch path
    # (communicator set-up, cold host code) are warm-up, not throughput
    barrier()
    run(K)
    prime_steps += K
    barrier()
    run(W)
    barrier()
    l0 = lib.apg_launch_count()
    ev0, ev1 = ev(), ev()
    barrier()
    t_wall0 = time.time()
    ev0.record()
    t_host0 = time.perf_counter()
    run(K)
    host_enqueue_us = (time.perf_counter() - t_host0) * 1e6
    ev1.record()
    back_to_back_ms = None
    if os.environ.get("APG_BENCH_DIAG"):  # a second window queued behind the first: device time without the idle-start gap
        ev2 = ev()
        run(K)
        ev2.record()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    if os.environ.get("APG_BENCH_DIAG"):
        back_to_back_ms = ev1.elapsed_time(ev2)
        sys.stderr.write("diag %s rank %d: timed window %.1f us, host enqueue %.1f us, same window queued back to back %.1f us\n"
                         % (name, rank, ms * 1e3, host_enqueue_us, back_to_back_ms * 1e3))
    launches = lib.apg_launch_count() - l0
    if os.environ.get("APG_BENCH_DIAG"):  # repeated windows: is the first one after the communicator came up special?
        for rep in range(6):
            if rep < 4:
                barrier()
            else:
                torch.cuda.synchronize()
            d0, d1 = ev(), ev()
            d0.record()
            th = time.perf_counter()
            run(K)
            th = (time.perf_counter() - th) * 1e6
            d1.record()
            torch.cuda.synchronize()
            sys.stderr.write("diag %s rank %d: repeat %d (%s) window %.1f us, host enqueue %.1f us\n"
                             % (name, rank, rep, "barrier" if rep < 4 else "sync only", d0.elapsed_time(d1) * 1e3, th))
        barrier()
    clocks = sampler.window(t_wall0, t_wall1) if rank == 0 else None
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        all_ms = [torch.zeros_like(t_ms) for _ in range(world)]
        dist.all_gather(all_ms, t_ms)
        per_rank_ms = [float(x.item()) for x in all_ms]
    else:
        per_rank_ms = [ms]
    ms_max = max(per_rank_ms)
    value = world * n * K / (ms_max * 1e-3)

    # ---- e2e: host buffers through the C-ABI call (the reference-facing step), copies inside the timed region;
    # measured twice: reward / flags only, and with the observation (obs["glimpse"] / obs["lidar"]) copied back too
    Ke = args.e2e_steps
    obs_elems = n * (env._B if is_lidar else env._S0 * env._S1 * env._C)
    # every step copies ITS OWN inputs: a pool of distinct page-locked host steps (taken from the device pool), walked
    # cyclically -- one fixed action repeated every step would drive the LIDAR robots into the walls (a different,
    # slower state distribution than the device-timed rollout's)
    hp = min(16, pool)
    h_as = [acts[i].cpu().pin_memory() for i in range(hp)]
    h_ps = [preds[i].cpu().pin_memory() for i in range(hp)]
    h_r = torch.empty((n,), dtype=torch.float32).pin_memory()
    h_te = torch.empty((n,), dtype=torch.uint8).pin_memory()
    h_tr = torch.empty((n,), dtype=torch.uint8).pin_memory()
    h_obs = torch.empty((obs_elems,), dtype=torch.float32).pin_memory()
    step_host = lib.apg_lidar_step_host if is_lidar else lib.apg_image_step_host

    e2e_cursor = [0]

    def e2e(with_obs):
        o = h_obs.data_ptr() if with_obs else None

        def one():
            i = e2e_cursor[0] % hp
            e2e_cursor[0] += 1
            _capi.check(step_host(env._h, h_as[i].data_ptr(), h_ps[i].data_ptr(), h_r.data_ptr(), h_te.data_ptr(),
                                  h_tr.data_ptr(), o))

        for _ in range(5):
            one()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            one()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * n * Ke / float(t.item())

    e2e_value = e2e(False)
    e2e_obs_value = e2e(True)
    h2d = n * (2 + pw) * 4
    d2h = n * (4 + 1 + 1)

    # ---- final NCCL reduce of the device-side episode statistics (the only collective of the data path)
    stats = env.episode_stats(reduce=True)
    env.sync()
    # sticky device-side error bits (NaN input, tap outside the image grid) of everything this process stepped: the
    # benchmark runs with check_errors="never", so look once at the end -- a non-zero value invalidates the run
    flags = C.c_uint32(0)
    _capi.check((lib.apg_lidar_poll_error if is_lidar else lib.apg_image_poll_error)(env._h, C.byref(flags), 0))
    t_f = torch.tensor([int(flags.value)], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t_f, op=dist.ReduceOp.MAX)
    error_flags = int(t_f.item())
    variant = None if is_lidar else lib.apg_image_variant(env._h)
    env.close()
    del env, acts, preds
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    peak = float(peaks.get("hbm_gbs", 6650.0))
    facts = kernel_facts(name)
    bytes_per_launch = w["bytes_per_step"] * n
    step_s = (ms_max * 1e-3) / K
    achieved = bytes_per_launch / step_s / 1e9
    if is_lidar:
        kernel = "lidar_step_kernel"
    else:
        kernel = {0: "image_head_kernel + image_step_kernel (staged tiles, warp blend)",
                  1: "image_step_kernel (staged tiles, thread blend)",
                  3: "tma_step_kernel (TMA tiles, persistent warps, three environments per pass)",
                  4: "image_rollout_kernel (persistent thread-per-env rollout)"}.get(variant, str(variant))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": facts.get("dram_bytes_per_step"),
                "algorithmic_bytes_per_launch": bytes_per_launch, "bytes_per_env_step": w["bytes_per_step"],
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                "kernel": kernel, "avg_step_us": step_s * 1e6, "launches_per_step": launches / K,
                "note": "achieved = algorithmic bytes of one vector step / device time of one vector step (all launches "
                        "of the step together); traffic = dram read+write bytes per vector step from the committed ncu "
                        "--set full capture (profiles/r02_kernel_facts.json)"}
    if is_lidar:
        # the ray walk is bound by instruction issue, not by bytes: the ALU-side figure is the achieved
        # warp-instruction rate against 4 issue slots x SMs x clock, with the instruction count per env-step from ncu
        ipe = facts.get("warp_instr_per_env_step")
        sm_clock = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        issue_peak = 4.0 * sms * sm_clock * 1e6
        roofline = dict(roofline, bound="issue", hbm_frac=roofline["frac"], hbm_achieved=achieved, hbm_peak=peak,
                        unit="warp-instr/s", peak=issue_peak,
                        achieved=None if ipe is None else ipe * (value / world),
                        frac=None if ipe is None else ipe * (value / world) / issue_peak,
                        warp_instr_per_env_step=ipe,
                        note="issue-bound kernel: achieved = warp-instructions per env-step (ncu, profiles/) x "
                             "env-steps/s per GPU; peak = 4 issue slots x SMs x SM clock; hbm_* keep the byte-side figures")
    res = {
        "value": value, "unit": "env-steps/s", "steps": K, "warmup": W, "ms_per_step": ms_max / K,
        "per_rank_ms": per_rank_ms, "primed_steps": prime_steps,
        "config": workload_config(name, n, world, args.blend, args.lidar_maps),
        "timing": {"explicit_reset_s": round(reset_s, 3),
                   "l2": "per-step inputs come from a pool of %d distinct steps streamed from HBM (%.0f MB > 126 MB L2, "
                         "walked cyclically); engine-owned observation buffers are rewritten every step" %
                         (pool, pool * n * (2 + pw) * 4 / 1e6),
                   "priming": "%d untimed steps (whole episode cycles, >= %.0f ms of device time) before the %d warm-up "
                              "steps" % (prime_steps, args.prime_ms, W)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                "api": ("apg_lidar_step_host" if is_lidar else "apg_image_step_host") +
                       " (pinned host action/prediction in, reward/terminated/truncated out)",
                "inputs": "%d distinct page-locked host steps walked cyclically (each step copies its own inputs)" % hp,
                "with_obs": {"value": e2e_obs_value, "d2h_bytes_per_step": d2h + obs_elems * 4,
                             "what": "the same call with the observation (" + ("lidar" if is_lidar else "glimpse") +
                                     ") copied back to pinned host memory as well"}},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "episode_stats_allreduce": stats,
        "error_flags": error_flags,
    }
    if error_flags:
        sys.stderr.write("bench.py: device error flags 0x%x were raised during %s -- result INVALID\n" % (error_flags, name))
    if not args.no_cpu_baseline and world == 1:
        n_cpu, st_cpu = CPU_SAMPLE[name]
        v, dt = cpu_oracle_run(name, n_cpu, st_cpu)
        res["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": 1, "kind": "port",
                               "sample": f"CPU restatement (oracle/), 1 process, N={n_cpu} x {st_cpu} steps, {dt:.1f} s"}
        rr = real_reference_figure(name)
        if rr is not None:
            res["cpu_baseline"]["real_reference"] = rr
    return res


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
        # bring the communicator up now: its lazy set-up (first collective) otherwise lands right in front of the first
        # workload's timed window and slows that window's launches (measured: +80 us on a 175 us window)
        for _ in range(3):
            dist.barrier()
        torch.cuda.synchronize()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    names = [args.workload] if args.only else [args.workload] + [x for x in ORDER if x != args.workload]
    results = {}
    for name in names:
        results[name] = run_workload(name, args, rank, world, local_rank, dist, sampler, peaks)
    sampler.stop()
    if rank == 0:
        head = results[args.workload]
        line = {"metric": "env-steps/sec", "value": head["value"], "unit": "env-steps/s", "n_gpus": world,
                "steps": head["steps"], "warmup": head["warmup"], "ms_per_step": head["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
        for k in ("config", "timing", "clocks", "e2e", "gpu_launches", "roofline", "per_rank_ms", "primed_steps",
                  "episode_stats_allreduce", "error_flags", "cpu_baseline"):
            if k in head:
                line[k] = head[k]
        line["gpu_launches"] = sum(r["gpu_launches"] for r in results.values())
        line["gpu_launches_headline"] = head["gpu_launches"]
        line["error_flags"] = max(r["error_flags"] for r in results.values())
        if not args.only:
            line["workloads"] = results
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
