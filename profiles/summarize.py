#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep captures into the small text summaries committed under profiles/.
usage: python profiles/summarize.py <report.ncu-rep> <out.txt> [title]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(vals, units)))
    lines = [f"# {title}", f"# source: ncu --set full --clock-control none --import-source on  ({rep})",
             f"kernel: {d.get('Kernel Name', ('?',))[0]}", ""]
    for k in WANT:
        if k in d:
            lines.append(f"{k:75s} {d[k][0]:>18s} {d[k][1]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    hi = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
    if hi:
        h = srows[hi[0]]
        ix = {n: i for i, n in enumerate(h)}
        data = [r for r in srows[hi[0] + 1:] if len(r) == len(h)]
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
        stall = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        agg = {c: sum(int(r[ix[c]] or 0) for r in data) for c in stall}
        lines += ["", "warp stall samples by reason (all SASS lines):"]
        for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
            lines.append(f"  {c:28s} {100 * v / tot:5.1f} %")
        lines += ["", "top SASS lines by stall samples:"]
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]:
            top = max(stall, key=lambda c: int(r[ix[c]] or 0))
            lines.append(f"  {100 * int(r[ix['# Samples']]) / tot:5.1f} %  {r[ix['Source']][:64]:64s} {top}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
