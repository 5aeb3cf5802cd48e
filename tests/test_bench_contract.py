"""CPU: the reference arm of bench.py (the CPU restatement on the host cores) runs here and prints ONE JSON line with
the keys the measurement contract names; the GPU arm's line is checked for the same keys on a B200 (-m gpu)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"]


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = _run(["--impl", "reference", "--steps", "17", "--warmup", "3"], 600)
    for k in COMMON:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["unit"] == "env-steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert sorted(d["config"]) == sorted(["workload", "name", "num_envs_per_gpu", "global_num_envs", "sharding", "blend", "maps"])
    assert d["steps"] == 17 and d["warmup"] == 3


@pytest.mark.gpu
def test_gpu_arm_json_line():
    d = _run(["--steps", "51", "--warmup", "17", "--e2e-steps", "5", "--only"], 900)
    for k in COMMON + ["clocks", "roofline", "error_flags"]:
        assert k in d, k
    assert "impl" not in d or d["impl"] != "reference"
    assert d["n_gpus"] == 1 and d["steps"] == 51 and d["value"] > 0 and d["gpu_launches"] >= 1
    assert d["error_flags"] == 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert len(d["per_rank_ms"]) == 1 and d["primed_steps"] >= 34
    assert "with_obs" in d["e2e"] and d["e2e"]["with_obs"]["value"] > 0
    assert sorted(d["config"]) == sorted(["workload", "name", "num_envs_per_gpu", "global_num_envs", "sharding", "blend", "maps"])


@pytest.mark.gpu
def test_gpu_arm_workloads_table():
    """The default line carries every BASELINE.json workload (small env counts here to keep the test short)."""
    d = _run(["--steps", "20", "--warmup", "5", "--e2e-steps", "3", "--envs", "8192", "--bank", "2048", "--prime-ms", "5",
              "--no-cpu-baseline"], 900)
    assert sorted(d["workloads"]) == ["lidar_maze", "lidar_rooms", "mnist", "tinyloc"]
    for name, w in d["workloads"].items():
        assert w["value"] > 0 and w["steps"] == 20 and w["error_flags"] == 0, name
        assert w["roofline"]["bound"] == ("issue" if name.startswith("lidar") else "hbm"), name
        assert w["e2e"]["value"] > 0 and w["e2e"]["with_obs"]["value"] > 0, name
    assert d["value"] == d["workloads"]["mnist"]["value"]
