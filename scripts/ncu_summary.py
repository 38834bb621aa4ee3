"""Turns ncu captures brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python scripts/ncu_summary.py launches <launches.csv> <out.md>      # per-kernel device-time shares of one bench run
  python scripts/ncu_summary.py full <prof.ncu-rep> <out.md>          # key metrics of a `--set full` capture
  python scripts/ncu_summary.py traffic <prof.ncu-rep> <out.json>     # DRAM bytes per launch per kernel
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "gpc__cycles_elapsed.avg.per_second", "sm__cycles_active.avg"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0][-70:] + (" grid=" + r[gi] if gi is not None else "")
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): `--metrics gpu__time_duration.sum --clock-control none` — cold-cache, serialised: compare SHARES\n\n")
        f.write("| kernel | launches | avg us | share % |\n|---|---:|---:|---:|\n")
        for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{n}` | {len(v)} | {sum(v) / len(v):.2f} | {100 * sum(v) / tot:.1f} |\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"## `{name[:120]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")


def traffic(src, dst):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged per kernel name -> JSON (bench.py's roofline.traffic)."""
    import json
    import re
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"^void ", "", r[hdr.index("Kernel Name")]).split("(")[0].split("<")[0].split("::")[-1]
        b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        t = float(r[it]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[it]]
        agg.setdefault(name, []).append((b, t))
    res = {"source": f"ncu --set full --clock-control none ({src})", "kernels": {
        k: {"dram_bytes_per_launch": sum(b for b, _ in v) / len(v), "launches": len(v), "avg_us_under_ncu": sum(t for _, t in v) / len(v)}
        for k, v in agg.items()}}
    json.dump(res, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
