#!/usr/bin/env python
"""Summarise an `ncu --set full` report: one line per captured launch with duration, DRAM traffic and the
achieved DRAM bandwidth against MEASURED_PEAKS.json, SM / tensor-pipe activity and occupancy.

    python tools/ncu_summary.py profiles/r02_batch.ncu-rep [--md out.md] [--algo tools/algo_bytes.json]

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU).  `--algo` maps kernel-name substrings
to the ALGORITHMIC bytes of one launch of the benchmark batch, from which the achieved
algorithmic GB/s and its fraction of the measured HBM peak are derived."""
from __future__ import annotations

import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WANT = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active": "tensor_pct2",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}


def to_float(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return None


def unit_scale(metric, unit):
    u = (unit or "").lower()
    if metric == "gpu__time_duration.sum":
        return {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}.get(u, 1e-3)
    if metric.startswith("dram__bytes"):
        return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "b": 1.0, "kb": 1e3, "mb": 1e6, "gb": 1e9}.get(u, 1.0)
    return 1.0


def load(rep):
    if rep.endswith(".csv"):  # already exported with `ncu -i <rep> --page raw --csv`
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(header)}
    res = []
    for r in rows[2:]:
        if len(r) < len(header):
            continue
        d = {"name": r[idx["Kernel Name"]], "id": r[idx["ID"]]}
        for m, k in WANT.items():
            if m in idx:
                v = to_float(r[idx[m]])
                d[k] = None if v is None else v * unit_scale(m, units[idx[m]])
        res.append(d)
    return res


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    return re.sub(r"ocrs::\w+::(\(anonymous namespace\)::|<unnamed>::)?", "", name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--md")
    ap.add_argument("--algo")
    ap.add_argument("--stages", help="tools/stage_map.json: also print one line per pipeline stage")
    ap.add_argument("--stages-md")
    a = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = peaks["hbm_gbs"]
    algo = json.load(open(a.algo)) if a.algo else {}
    rows = load(a.rep)
    lines = ["| # | kernel | grid x block | regs | duration us | DRAM read MB | DRAM write MB | DRAM GB/s | of measured %.1f GB/s | algorithmic MB | algorithmic GB/s (frac) | SM %% | tensor pipe %% | L2 hit %% | warps active %% |" % hbm,
             "|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for d in rows:
        dur = d.get("dur") or 0.0
        rd, wr = d.get("rd") or 0.0, d.get("wr") or 0.0
        gbs = (rd + wr) / (dur * 1e-6) / 1e9 if dur else 0.0
        nm = short(d["name"])
        ab = None
        for k, v in algo.items():
            if k in nm:
                ab = v
        a_gbs = ab / (dur * 1e-6) / 1e9 if (ab and dur) else None
        tp = d.get("tensor_pct") if d.get("tensor_pct") is not None else d.get("tensor_pct2")
        lines.append("| %s | `%s` | %s x %s | %s | %.1f | %.2f | %.2f | %.0f | %.3f | %s | %s | %s | %s | %s | %s |" % (
            d["id"], nm[:70], int(d.get("grid") or 0), int(d.get("block") or 0), int(d.get("regs") or 0), dur, rd / 1e6, wr / 1e6, gbs,
            gbs / hbm, "%.2f" % (ab / 1e6) if ab else "", "%.0f (%.3f)" % (a_gbs, a_gbs / hbm) if a_gbs else "",
            "%.1f" % d["sm_pct"] if d.get("sm_pct") is not None else "", "%.1f" % tp if tp is not None else "",
            "%.1f" % d["l2_hit"] if d.get("l2_hit") is not None else "", "%.1f" % d["occ"] if d.get("occ") is not None else ""))
    if a.stages:
        sm = json.load(open(a.stages))["stages"]
        sl = ["| stage | kernels (launches in the batch) | time us | DRAM MB (read + write, ncu) | DRAM GB/s | of measured %.1f GB/s | algorithmic MB | algorithmic GB/s | of measured peak | note |" % hbm,
              "|---|---|---:|---:|---:|---:|---:|---:|---:|---|"]
        for st in sm:
            sel = [d for d in rows if any(k in short(d["name"]) for k in st["kernels"])]
            if not sel:
                continue
            dur = sum(d.get("dur") or 0.0 for d in sel)
            byt = sum((d.get("rd") or 0.0) + (d.get("wr") or 0.0) for d in sel)
            gbs = byt / (dur * 1e-6) / 1e9 if dur else 0.0
            ab = st.get("algo_bytes")
            a_gbs = ab / (dur * 1e-6) / 1e9 if (ab and dur) else None
            sl.append("| %s | %s (%d) | %.1f | %.1f | %.0f | %.3f | %s | %s | %s | %s |" % (
                st["stage"], ", ".join("`%s`" % k for k in st["kernels"]), len(sel), dur, byt / 1e6, gbs, gbs / hbm,
                "%.1f" % (ab / 1e6) if ab else "--", "%.0f" % a_gbs if a_gbs else "--", "%.3f" % (a_gbs / hbm) if a_gbs else "--", st.get("note", "")))
        stext = "\n".join(sl) + "\n"
        if a.stages_md:
            open(a.stages_md, "w").write(stext)
        sys.stdout.write(stext + "\n")
    text = "\n".join(lines) + "\n"
    if a.md:
        open(a.md, "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
