"""Turns gpurun_out/*.csv / *.ncu-rep (scratch) into the committed summaries under profiles/.
  python profiles/summarize.py launches gpurun_out/launches_r01.csv  profiles/r01_launches.txt
  python profiles/summarize.py ncu      gpurun_out/prof_pcg_r01.ncu-rep profiles/r01_ncu_pcg.txt
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__throughput.avg.pct_of_peak_sustained_active",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(src, dst):
    rows = list(csv.reader(open(src, errors="ignore")))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        k = re.sub(r"\(.*", "", d["Kernel Name"])
        v = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(d["Metric Unit"], 1e-6)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as fh:
        fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (per-launch times are cold-cache and serialised: compare SHARES)\n")
        fh.write(f"# source: {src}; {len(data)} launches, {tot:.3f} ms total\n")
        fh.write(f"{'kernel':58s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share_%':>8s}\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            fh.write(f"{k:58s} {a[0]:8d} {a[1]:10.3f} {a[1] / a[0] * 1e3:10.1f} {a[1] / tot * 100:8.1f}\n")
    print(open(dst).read())


def ncu(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none --import-source on ; source: {src}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            fh.write(f"\n== {d['Kernel Name'][:110]}\n")
            for m in METRICS:
                if m in d:
                    fh.write(f"  {m:82s} {d[m]:>16s} {units[hdr.index(m)]}\n")
            try:
                t = float(d["gpu__time_duration.sum"].replace(",", ""))
                tu = units[hdr.index("gpu__time_duration.sum")]
                t_s = t * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "second": 1.0}.get(tu, 1e-6)
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                b = sum(float(d[k].replace(",", "")) * scale.get(units[hdr.index(k)], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                fh.write(f"  {'-> dram traffic (read+write)':82s} {b / 1e9:16.4f} GB\n")
                fh.write(f"  {'-> dram traffic / duration':82s} {b / t_s / 1e9:16.1f} GB/s\n")
            except (KeyError, ValueError):
                pass
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "ncu": ncu}[sys.argv[1]](sys.argv[2], sys.argv[3])
