"""Shared-memory wavefronts per opcode of an `ncu --set full --import-source on` capture:

    python tools/ncu_shared_wavefronts.py <capture.ncu-rep> <units per launch>

(the l1tex data stage serves one wavefront per cycle and SM: wavefronts x units / SMs / clock = the LSU floor of the kernel)
"""
import csv
import subprocess
import sys


def main(rep, units):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        src = r[col["Source"]].strip().split()
        if not src:
            continue
        op = src[1] if src[0].startswith("@") else src[0]
        w = float(r[col["L1 Wavefronts Shared"]] or 0)
        if w > 0:
            a = agg.setdefault(op, [0.0, 0.0, 0.0])
            a[0] += w
            a[1] += float(r[col["L1 Wavefronts Shared Ideal"]] or 0)
            a[2] += float(r[col["Instructions Executed"]] or 0)
    tot = 0.0
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%-22s wavefronts/unit %7.1f  ideal %7.1f  instr/unit %6.1f" % (k, v[0] / units, v[1] / units, v[2] / units))
        tot += v[0]
    print("%-22s wavefronts/unit %7.1f" % ("total", tot / units))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]))
