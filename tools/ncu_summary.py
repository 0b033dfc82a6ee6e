#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (tracked).
  python tools/ncu_summary.py launches gpurun_out/r01_launches.csv profiles/r01_launches_summary.txt
  python tools/ncu_summary.py raw gpurun_out/r01_prof_gemm.ncu-rep profiles/r01_gemm_ncu.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_uniform.sum", "smsp__cycles_active.avg"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r"^void |hb::\(anonymous namespace\)::|unnamed>::", "", name).replace("<", "<").strip()


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1)
        a = agg.setdefault(short(row["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as o:
        o.write(f"# ncu --metrics gpu__time_duration.sum launch list: {src}\n# per-launch times are cold-cache and serialised: compare SHARES\n")
        o.write(f"total_ms {tot / 1e6:.3f}\n{'share%':>8} {'launches':>8} {'avg_us':>10}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write(f"{v[1] / tot * 100:8.2f} {v[0]:8d} {v[1] / v[0] / 1e3:10.1f}  {k}\n")
    print(open(dst).read())


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in KEEP or h == "Kernel Name"]
    with open(dst, "w") as o:
        o.write(f"# ncu --set full --clock-control none: {src}\n")
        for r in rows[2:]:
            o.write("\n")
            for i in idx:
                val = short(r[i]) if hdr[i] == "Kernel Name" else r[i]
                o.write(f"{hdr[i]:<70} {val} {units[i]}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
