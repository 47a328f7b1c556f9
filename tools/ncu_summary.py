"""Summarise an `ncu -i <rep> --page raw --csv` dump into one JSON record per kernel launch.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > gpurun_out/x_raw.csv
    python tools/ncu_summary.py gpurun_out/x_raw.csv profiles/x_summary.json

Units vary per row in ncu's CSV (byte / Kbyte / Mbyte / Gbyte, us / ms): everything is converted to MB and ms.
"""
import csv
import json
import sys

SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}
TIME = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}
PICK = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "gpu__dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm__throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__grid_size": "grid",
    "launch__registers_per_thread": "regs",
}


def main(src, dst):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[start], rows[start + 1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[start + 2:]:
        if len(r) < len(hdr):
            continue
        rec = {"kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("mla::", "")}
        for metric, name in PICK.items():
            if metric not in col:
                continue
            v, u = r[col[metric]], units[col[metric]]
            try:
                v = float(v.replace(",", ""))
            except ValueError:
                continue
            if name == "time":
                rec["time_ms"] = round(v * TIME.get(u, 1.0), 5)
            elif name in ("dram_read", "dram_write"):
                rec[name + "_MB"] = round(v * SCALE.get(u, 1.0), 3)
            else:
                rec[name] = round(v, 2)
        if "time_ms" in rec and "dram_read_MB" in rec:
            rec["dram_GBs"] = round((rec["dram_read_MB"] + rec.get("dram_write_MB", 0.0)) / rec["time_ms"], 1)
        out.append(rec)
    json.dump(out, open(dst, "w"), indent=1)
    for rec in out:
        print(rec)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
