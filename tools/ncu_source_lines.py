"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep captured with
`ncu --set full --import-source on` (sources compiled with -lineinfo).

    python tools/ncu_source_lines.py gpurun_out/prof.ncu-rep profiles/<name>_source_lines.txt <envs per launch>
"""
import csv
import subprocess
import sys


def main(rep, out, n_units):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    per = []
    for r in rows[3:]:
        if len(r) > 8 and r[0].isdigit() and r[7].isdigit():
            per.append((int(r[0]), int(r[7]), int(r[4]) if r[4].isdigit() else 0, r[1].strip()[:96]))
    inst = sum(p[1] for p in per) or 1
    samp = sum(p[2] for p in per) or 1
    lines = [f"# {rep.split('/')[-1]}: warp-instructions executed and warp-stall samples per source line",
             f"# kernel: {rows[1][1] if len(rows) > 1 and len(rows[1]) > 1 else '?'}",
             f"# total {inst} warp-instructions = {inst / n_units:.1f} per env-step ({n_units} envs per launch), {samp} samples",
             "# line  instr/env  instr%  samples%  source"]
    for ln, i, s, src in per:
        if i / inst >= 0.008 or s / samp >= 0.015:
            lines.append(f"{ln:5d} {i / n_units:9.1f} {100 * i / inst:6.1f} {100 * s / samp:8.1f}  {src}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:8]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]))
