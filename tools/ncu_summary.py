#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (tracked).
  python tools/ncu_summary.py launches gpurun_out/r01_launches.csv profiles/r01_launches_summary.txt
  python tools/ncu_summary.py raw gpurun_out/r01_prof_gemm.ncu-rep profiles/r01_gemm_ncu.txt
  python tools/ncu_summary.py stalls gpurun_out/attn.ncu-rep profiles/r01_attn_stalls.txt
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


def stalls(src, dst):
    """Source-level warp-stall sampling of a `--set full --import-source on` capture: the hottest SASS instructions, every
    mbarrier wait (fast-path executions, spin-loop executions, samples) and the samples grouped by per-role execution count.
    This is the view that showed the attention MMA issuer never waiting for its inputs while the tensor pipe idled
    (per-instruction waterfall loops under `if (lane == 0)`)."""
    out = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    ia, isrc, iall, iex = (hdr.index(h) for h in ("Address", "Source", "Warp Stall Sampling (All Samples)", "Instructions Executed"))
    num = lambda x: int(x or 0)
    tot = sum(num(r[iall]) for r in data)
    by_exec = collections.Counter()
    for r in data:
        by_exec[num(r[iex])] += num(r[iall])
    with open(dst, "w") as o:
        o.write(f"# ncu source page: {src}\n# kernel: {short(rows[0][1])}\ntotal samples {tot}\n\n")
        o.write("samples by instruction execution count (one count per role / loop level):\n")
        for k, v in by_exec.most_common(8):
            o.write(f"  executed {k:>10} times: {v:>8} samples ({100.0 * v / max(tot, 1):.1f} %)\n")
        o.write("\nhottest instructions:\n")
        for r in sorted(data, key=lambda r: -num(r[iall]))[:25]:
            o.write(f"  {r[ia][-6:]} {num(r[iall]):>8} samples  executed {num(r[iex]):>10}  {r[isrc][:100]}\n")
        o.write("\nmbarrier waits (SYNCS.PHASECHK.TRYWAIT; a fast-path try is followed by its spin loop):\n")
        for i, r in enumerate(data):
            if "SYNCS.PHASECHK" in r[isrc] and num(r[iex]) > 0:
                nxt = data[i + 1] if i + 1 < len(data) else r
                m = re.search(r"\+0x([0-9a-f]+)\]", r[isrc])
                o.write(f"  {r[ia][-6:]} executed {num(r[iex]):>10}  samples {num(r[iall]) + num(nxt[iall]):>7}  smem+0x{m.group(1) if m else '?'}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "stalls": stalls}[sys.argv[1]](sys.argv[2], sys.argv[3])
