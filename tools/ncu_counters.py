#!/usr/bin/env python
"""Per-kernel counters beyond profiles/summarize.py, read from an existing .ncu-rep (no GPU needed): which unit is busy, how many warps wait for what.

    python tools/ncu_counters.py gpurun_out/r02_s4_ncu_fvm.ncu-rep [kernel-name substring ...] > profiles/r02_ncu_fvm_counters.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg", "l1tex__t_sectors.sum", "l1tex__t_sectors_lookup_hit.sum", "l1tex__t_sectors_lookup_miss.sum",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.sum", "lts__t_sectors_srcunit_tex_lookup_hit.sum",
    "lts__t_sectors_srcunit_tex_lookup_miss.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep, pats = sys.argv[1], sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {}
    for i, h in enumerate(hdr):
        col.setdefault(h.split("TriageCompute.")[-1], i)       # some metrics carry a section prefix
    name_i = hdr.index("Kernel Name")
    print(f"# {rep}: ncu --set full --clock-control none, counters read with `ncu -i ... --page raw --csv` (tools/ncu_counters.py)")
    seen = set()
    for r in rows[2:]:
        k = r[name_i]
        if pats and not any(p in k for p in pats):
            continue
        if k in seen:
            continue
        seen.add(k)
        print(f"\n== {k}")
        for m in METRICS:
            if m in col and r[col[m]] != "":
                print(f"  {m:90s} {r[col[m]]:>20s} {units[col[m]]}")


if __name__ == "__main__":
    main()
