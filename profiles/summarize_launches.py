#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`):
one row per kernel with launches, total time and share.  Usage: python profiles/summarize_launches.py X.csv out.md "title" """
import csv
import sys
from collections import OrderedDict


def main():
    src, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ci = {h: i for i, h in enumerate(hdr)}
    acc = OrderedDict()
    for r in rows:
        if r is hdr or len(r) != len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "").replace("crfgpu::<unnamed>::", "").replace("crfgpu::", "").replace("unnamed>::", "").replace("(bool)", "").replace("(int)", "")
        unit, val = r[ci["Metric Unit"]], float(r[ci["Metric Value"]].replace(",", ""))
        ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        a = acc.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ms
    total = sum(v[1] for v in acc.values())
    n = sum(v[0] for v in acc.values())
    lines = [f"# ncu launch list{': ' + title if title else ''}", "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold cache, serialised: compare SHARES, not absolutes); raw csv: `"
             + src.split("/")[-1] + "`.", "", f"{n} launches captured, {total:.2f} ms of kernel time.", "",
             "| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
    for k, (c, ms) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {c} | {ms:.3f} | {100 * ms / total:.1f} % | {1000 * ms / c:.1f} |")
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
