#!/usr/bin/env python
"""Summarise an `ncu --set full` report: one row per distinct kernel (first profiled launch), with the counters the
roofline in bench.py / DESIGN.md refers to.  Usage:
    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/rN_name.md [profiles/traffic.json]
The optional JSON maps kernel name -> DRAM bytes (read+write) of one launch; bench.py reads it for `roofline.traffic`."""
import csv
import io
import json
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("lts__t_sectors.sum", "L2 sectors"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_size", "cluster")]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out_md = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    seen, lines, traffic = set(), [], {}
    lines.append("| kernel | " + " | ".join(n for _, n in KEYS) + " |")
    lines.append("|---|" + "---|" * len(KEYS))
    for r in rows[2:]:
        name = r[ci["Kernel Name"]]
        short = name.replace("void ", "").replace("crfgpu::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("unnamed>::", "")
        short = short.replace("(bool)", "").replace("(int)", "").split("(")[0]
        key = (short, r[ci.get("launch__grid_size", 0)])
        if key in seen:
            continue
        seen.add(key)
        cells = []
        for k, _ in KEYS:
            if k in ci and r[ci[k]] != "":
                cells.append(f"{r[ci[k]]} {units[ci[k]]}".strip())
            else:
                cells.append("-")
        lines.append(f"| `{short}` | " + " | ".join(cells) + " |")
        try:
            rd = float(r[ci["dram__bytes_read.sum"]]) * TO_BYTES[units[ci["dram__bytes_read.sum"]]]
            wr = float(r[ci["dram__bytes_write.sum"]]) * TO_BYTES[units[ci["dram__bytes_write.sum"]]]
            traffic.setdefault(short, rd + wr)
        except Exception:
            pass
    open(out_md, "w").write(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n\n(first profiled launch of each kernel; times are under the profiler, cold cache, serialised)\n\n" + "\n".join(lines) + "\n")
    if len(sys.argv) > 3:
        old = {}
        try:
            old = json.load(open(sys.argv[3]))
        except Exception:
            pass
        old.update(traffic)
        json.dump(old, open(sys.argv[3], "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
