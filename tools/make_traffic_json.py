#!/usr/bin/env python
"""profiles/r02_kernel_traffic.json from the raw CSV of an `ncu --set full` capture of ONE benchmark batch:
per profiler operator class (the names bench.py's roofline uses) the DRAM bytes (read + write) per launch.

    python tools/make_traffic_json.py profiles/r02w_batch_raw.csv.gz "ncu --set full ... (profiles/...)" """
from __future__ import annotations

import csv
import gzip
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = {
    "rec/conv3x3_tc_kernel": ["conv3x3_ws_kernel", "conv3x3_res_kernel", "conv3x3_halo_kernel", "conv3x3_tc_kernel", "conv3x3_pair_kernel"],
    "rec/GRU recurrence(packed)": ["gru_cluster3_kernel", "gru_cluster_kernel"],
    "rec/stem_kernel": ["stem_kernel"],
}


def main():
    path, source = sys.argv[1], sys.argv[2]
    raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(header)}

    def scale(u):
        return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)

    out = {"source": source, "ops": {}}
    for op, names in OPS.items():
        sel = [r for r in rows[2:] if len(r) >= len(header) and any(n in r[idx["Kernel Name"]] for n in names)]
        if not sel:
            continue
        tot = 0.0
        dur = 0.0
        for r in sel:
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[idx[m]].replace(",", "")) * scale(units[idx[m]])
            dur += float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1e-3)
        out["ops"][op] = {"launches_in_capture": len(sel), "dram_bytes_per_launch": tot / len(sel),
                          "duration_us_per_launch_under_ncu": dur / len(sel)}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_kernel_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
