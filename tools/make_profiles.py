"""Turns the raw ncu output of a GPU session (gpurun_out/) into the tracked summaries under profiles/:

  profiles/<tag>_launches.txt        per-kernel totals of the `--metrics gpu__time_duration.sum` launch list of one
                                     eager decoder step (cold-cache, serialised: shares, not absolute times)
  profiles/<tag>_ncu_<kernel>.txt    key metrics of every captured launch of the `--set full` reports
  profiles/r01_ncu_traffic.json      measured DRAM bytes (read + write) per launch, averaged per kernel family;
                                     bench.py copies these into roofline.traffic

usage: python tools/make_profiles.py <tag> [gpurun_out]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_metrics import KEYS  # noqa: E402


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    tag = sys.argv[1]
    src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)
    lc = os.path.join(src, "launches.csv")
    if os.path.exists(lc):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), lc], capture_output=True,
                             text=True).stdout
        with open(os.path.join(out, f"{tag}_launches.txt"), "w") as f:
            f.write("# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python bench.py --profile\n"
                    "# (one eager decoder step, batch 16, large-v2; cold-cache serialised launches)\n" + txt)
    traffic = {}
    for rep, fam in (("prof_xattn.ncu-rep", "cross_attention_rowhead_kernel"), ("prof_gemm.ncu-rep", "woq_gemm_tc_kernel")):
        path = os.path.join(src, rep)
        if not os.path.exists(path):
            continue
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        lines = [f"# ncu --set full --clock-control none --import-source on -k regex:{fam.split('_kernel')[0]} python bench.py --profile",
                 f"# source report: {rep} (not tracked)"]
        tot, n = 0.0, 0
        for r in rows[2:]:
            lines.append("== " + r[hdr.index("Kernel Name")][:100] + "  grid " + r[hdr.index("Grid Size")])
            for k in KEYS:
                if k in hdr:
                    lines.append(f"   {k:85s} {r[hdr.index(k)]:>14s} {units[hdr.index(k)]}")
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            tot += to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
            n += 1
        if n:
            traffic[fam] = tot / n
            lines.append(f"# average DRAM traffic per launch over {n} launches: {tot / n / 1e6:.3f} MB")
        with open(os.path.join(out, f"{tag}_ncu_{fam}.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")
    if traffic:
        with open(os.path.join(out, "r01_ncu_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    print("profiles written:", sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
