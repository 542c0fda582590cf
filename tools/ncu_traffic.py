#!/usr/bin/env python3
"""profiles/ncu_traffic.json from `ncu --page raw --csv` exports of `--set full` captures: DRAM bytes (read + write)
per work unit of the dominant kernels, which bench.py multiplies by the live unit count of a launch to fill
`roofline.traffic` / `roofline_ntt.traffic` (so the number follows the committed capture, not a constant in bench.py).

    python tools/ncu_traffic.py KEY=CSV:UNITS[:KERNEL_REGEX] ...
      KEY     entry name bench.py looks up (k_msm_accum_g1, k_ntt_pass)
      CSV     raw-page CSV of one capture (profiles/*.csv)
      UNITS   work units of the captured launch (mixed additions; field elements read + written by the pass)
Existing entries of the JSON are kept unless overwritten."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_traffic.json")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    cur = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for arg in sys.argv[1:]:
        key, rest = arg.split("=", 1)
        parts = rest.split(":")
        path, units = parts[0], float(parts[1])
        rx = re.compile(parts[2]) if len(parts) > 2 else None
        rows = list(csv.reader(open(path)))
        hdr, unit_row = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        row = next(r for r in rows[2:] if rx is None or rx.search(r[col["Kernel Name"]]))

        def val(name):
            i = col[name]
            return float(row[i]) * SCALE[unit_row[i]]

        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        dur_i = col["gpu__time_duration.sum"]
        cur[key] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "units": units,
                    "dram_bytes_per_unit": (rd + wr) / units, "kernel": row[col["Kernel Name"]][:120],
                    "duration": f"{row[dur_i]} {unit_row[dur_i]}", "source": os.path.relpath(path, ROOT)}
    json.dump(cur, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(cur, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
