#!/usr/bin/env python
"""Per-kernel shares of a step from an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).
usage: python profiles/launch_shares.py <launches.csv> <out.txt> [skip_first_launches]"""
import collections
import csv
import sys


def main():
    src, out = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = list(csv.reader(open(src)))
    hdr = next(r for r in rows if r and r[0] == "ID")
    ix = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows if len(r) == len(hdr) and r[0].isdigit() and int(r[0]) >= skip]
    t = collections.defaultdict(list)
    for r in data:
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        t[r[ix["Kernel Name"]].split("(")[0].replace("void ", "")].append(us)
    total = sum(sum(v) for v in t.values())
    lines = [f"# {src}: launches {skip}.. of `python bench.py --steps 10 --warmup 10 --no-cpu-baseline` (tank 8M drop)",
             "# per-launch times under ncu are cold-cache and serialised: compare SHARES with the bench line's stage table"]
    for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"{k:44s} n={len(v):4d} avg={sum(v) / len(v):9.1f} us share={100 * sum(v) / total:5.1f}%")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
