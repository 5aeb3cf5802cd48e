"""Prints the README results table from the committed bench lines (profiles/r02_bench_{1,2,4,8}gpu.json,
profiles/r02_bench_reference_arm.json, profiles/r02_long_windows.jsonl)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(n):
    return json.loads(open(os.path.join(P, f"r02_bench_{n}gpu.json")).read().strip().splitlines()[-1])


def main():
    b = {n: load(n) for n in (1, 2, 4, 8) if os.path.exists(os.path.join(P, f"r02_bench_{n}gpu.json"))}
    one = b[1]
    print("| workload (per GPU) | µs / vector step | env-steps/s, 1 GPU | roofline | fraction | end to end (host buffers) | … with observation | 2 GPUs | 4 GPUs | 8 GPUs |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    names = {"mnist": "MNIST shape, N = 65 536", "tinyloc": "TinyImageNetLoc shape, N = 131 072",
             "lidar_rooms": "LIDARLocRooms, 64 beams, N = 262 144", "lidar_maze": "LIDARLocMaze 63², 256 beams, N = 131 072"}
    for k, w in one["workloads"].items():
        r = w["roofline"]
        bound = "HBM (algorithmic bytes)" if r["bound"] == "hbm" else "instruction issue"
        multi = " | ".join(f"{b[n]['workloads'][k]['value']:.3g}" if n in b else "—" for n in (2, 4, 8))
        print(f"| {names.get(k, k)} | {w['ms_per_step'] * 1e3:.1f} | {w['value']:.3g} | {bound} | {r['frac']:.2f} | "
              f"{w['e2e']['value']:.3g} | {w['e2e']['with_obs']['value']:.3g} | {multi} |")
    ref = json.loads(open(os.path.join(P, "r02_bench_reference_arm.json")).read().strip().splitlines()[-1])
    cb = ref["cpu_baseline"]
    print()
    print(f"Reference arm (`bench.py --impl reference`, MNIST shape, same box): {ref['value']:.3g} env-steps/s on {cb['cores']} host "
          f"processes of the CPU restatement; the unmodified reference itself runs at {cb['real_reference']['value']:.0f} "
          f"env-steps/s per core ({cb['real_reference']['sample']}).")
    lw = os.path.join(P, "r02_long_windows.jsonl")
    if os.path.exists(lw):
        parts = []
        for line in open(lw):
            w = json.loads(line)
            parts.append(f"{w['config']['name']} {w['ms_per_step'] * 1e3:.2f} µs (fraction {w['roofline']['frac']:.2f})")
        print("At 200-step windows (launch ramp and tail amortised): " + ", ".join(parts) + ".")


if __name__ == "__main__":
    main()
