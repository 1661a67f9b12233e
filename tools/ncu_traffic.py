#!/usr/bin/env python
"""profiles/force_kernel_traffic.json from an `ncu --set full ... --page raw --csv` export of the pair kernel: DRAM bytes per
launch, tied to the kernel sources it was taken from (bench.py prints `traffic: null` when the hash differs).
   python tools/ncu_traffic.py gpurun_out/<name>_raw.csv profiles/<committed copy>_raw.csv"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(path, committed):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    recs = [r for r in rows[2:] if "k_force_tile" in r[ki]]
    if not recs:
        raise SystemExit("no k_force_tile launch in " + path)

    def col(name, r):
        v, u = float(r[hdr.index(name)].replace(",", "")), units[hdr.index(name)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    recs = recs[:1]  # one launch, as exported (tests/test_profiles.py finds exactly these numbers in the committed file)
    rd = col("dram__bytes_read.sum", recs[0])
    wr = col("dram__bytes_write.sum", recs[0])
    t_us = float(recs[0][hdr.index("gpu__time_duration.sum")].replace(",", ""))
    old = {}
    tp = os.path.join(ROOT, "profiles", "force_kernel_traffic.json")
    if os.path.exists(tp):
        old = json.load(open(tp))
    out = {
        "kernel": recs[0][ki].split("(")[0].replace("void ", ""),
        "source": committed + " (ncu --set full --clock-control none, N=1e6, 1 B200, first captured launch)",
        "source_sha16": bench.kernel_source_sha16(),
        "dram_bytes_read_per_launch": int(rd), "dram_bytes_write_per_launch": int(wr), "dram_bytes_per_launch": int(rd + wr),
        "kernel_us_under_ncu": t_us,
        "note": "the kernel reads 16-bit tile-local rows (2 B per listed neighbour, padded to whole 32-entry passes) where the "
                "algorithmic count of SURVEY 8d charges 4 B; positions arrive once per chunk as bulk copies of the tile",
        "gather_kernel": old.get("gather_kernel", {}),
    }
    json.dump(out, open(tp, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
