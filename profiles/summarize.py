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
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]


UNITS = {"l1tex": "l1tex__throughput.avg.pct_of_peak_sustained_active", "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
         "l2": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "fma_pipe": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
         "alu_pipe": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"}


def raw_metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    return dict(zip(rows[0], zip(rows[2], rows[1])))


def num(d, key):
    v, unit = d[key]
    x = float(v.replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)


def evidence_json():
    """usage: summarize.py --json <workload> <out.json> <stage>=<report.ncu-rep> ...
    Writes the file bench.py reads its `roofline.traffic` from, stamped with the hash of the kernel sources of THIS tree
    (run it right after the capture, on the tree that was profiled)."""
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    workload, out = sys.argv[2], sys.argv[3]
    kernels = {}
    for spec in sys.argv[4:]:
        stage, rep = spec.split("=", 1)
        d = raw_metrics(rep)
        pct = {u: round(float(d[m][0].replace(",", "")), 1) for u, m in UNITS.items() if m in d}
        kernels[stage] = {"kernel": d["Kernel Name"][0].split("(")[0].split("::")[-1],
                          "dram_bytes": int(num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum")),
                          "time_us_under_ncu": round(num(d, "gpu__time_duration.sum") / (1e3 if d["gpu__time_duration.sum"][1] == "ns" else 1.0), 1),
                          "bound_unit": max(pct, key=pct.get), "units_pct": pct,
                          "l2_hit_pct": round(float(d["lts__t_sector_hit_rate.pct"][0]), 1),
                          "l1_hit_pct": round(float(d["l1tex__t_sector_hit_rate.pct"][0]), 1),
                          "lsu_wavefronts": int(num(d, "l1tex__data_pipe_lsu_wavefronts.sum")) if "l1tex__data_pipe_lsu_wavefronts.sum" in d else None,
                          "inst_executed": int(num(d, "smsp__inst_executed.sum")), "report": os.path.basename(rep)}
    doc = {"workload": workload, "source_hash": bench.kernel_source_hash(),
           "source": "ncu --set full --clock-control none, one launch per kernel; summaries beside this file", "kernels": kernels}
    open(out, "w").write(json.dumps(doc, indent=1) + "\n")
    print("wrote", out)


def main():
    if sys.argv[1] == "--json":
        return evidence_json()
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
