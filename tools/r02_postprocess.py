"""Turns the round-2 GPU captures in gpurun_out/ into the committed evidence under profiles/:

    python tools/r02_postprocess.py

  profiles/r02_<workload>.txt                 counter summary of the ncu --set full capture (tools/ncu_summary.py)
  profiles/r02_<workload>_source_lines.txt    warp-instructions / stall samples per source line (tools/ncu_source_lines.py)
  profiles/r02_kernel_facts.json              per workload: DRAM bytes per vector step and warp-instructions per env-step of
                                              the captured launch (bench.py reads it into roofline.traffic / the LIDAR
                                              issue-side roofline)
  profiles/r02_sass_evidence.txt              SASS mnemonics of the shipped library: UTMALDG / UBLKCP / LDGSTS / STG.E.128
Inputs: gpurun_out/prof_r02_<workload>.ncu-rep captured with
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <n> -c 1 ... bench.py --only --workload <w> --steps 20
(the captured launch is the timed 20-step rollout window for the image workloads, one vector step for LIDAR).
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")
# workload: (envs per GPU, vector steps inside the captured launch)
WORK = {"mnist": (65536, 20), "tinyloc": (131072, 20), "lidar_rooms": (262144, 1), "lidar_maze": (131072, 1)}


def raw_metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    return col, units, data


def to_bytes(value, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(value) * mult


def main():
    facts = {}
    for w, (n, steps) in WORK.items():
        rep = os.path.join(GP, f"prof_r02_{w}.ncu-rep")
        if not os.path.exists(rep):
            print("missing", rep)
            continue
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, os.path.join(OUT, f"r02_{w}.txt")],
                       stdout=subprocess.DEVNULL)
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_lines.py"), rep,
                        os.path.join(OUT, f"r02_{w}_source_lines.txt"), str(n * steps)], stdout=subprocess.DEVNULL)
        col, units, data = raw_metrics(rep)
        r = data[-1]
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        inst = float(r[col["inst_executed"]]) if "inst_executed" in col else float(r[col["smsp__inst_executed.sum"]])
        dur_i = col["gpu__time_duration.sum"]
        dur_us = float(r[dur_i]) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[dur_i], 1)
        facts[w] = {
            "kernel": r[col["Kernel Name"]][:100],
            "vector_steps_in_captured_launch": steps,
            "dram_bytes_per_step": (rd + wr) / steps,
            "dram_read_bytes_per_step": rd / steps,
            "dram_write_bytes_per_step": wr / steps,
            "warp_instr_per_env_step": inst / (n * steps),
            "ncu_duration_us_per_step": dur_us / steps,
            "ipc": float(r[col["sm__inst_executed.avg.per_cycle_elapsed"]]),
            "source": f"ncu --set full --clock-control none, profiles/r02_{w}.txt",
        }
    json.dump(facts, open(os.path.join(OUT, "r02_kernel_facts.json"), "w"), indent=1)
    print(json.dumps(facts, indent=1))
    lib = os.path.join(ROOT, "active-perception-gym_b200", "ap_gym_b200", "_lib", "libapg_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    lines = ["# SASS evidence of the shipped library (cuobjdump -sass ap_gym_b200/_lib/libapg_b200.so)",
             "# counts of the mnemonics that identify TMA / bulk copies / cp.async / 128-bit stores, per kernel", ""]
    cur, counts = None, {}
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = {}
            continue
        if cur is None:
            continue
        for key in ("UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "STG.E.128", "STG.E ", "RED", "LDS.64", "DMUL", "DADD"):
            if key in ln:
                counts[cur][key.strip()] = counts[cur].get(key.strip(), 0) + 1
    for fn, c in counts.items():
        if any(k in c for k in ("UTMALDG", "UBLKCP", "LDGSTS", "STG.E.128")) or "lidar" in fn:
            short = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:110]
            lines.append(f"{short}\n    " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
    lines.append("")
    lines.append("# excerpt: every bulk copy (UBLKCP global<->shared), tensor-map copy (UTMALDG, APG_TMA_BLOCKED=0 builds only) and mbarrier operation of the library")
    for ln in sass.splitlines():
        if "UTMALDG" in ln or "UBLKCP" in ln or "SYNCS.ARRIVE.TRANS64" in ln or "SYNCS.PHASECHK" in ln:
            lines.append(ln.rstrip()[:120])
    open(os.path.join(OUT, "r02_sass_evidence.txt"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
