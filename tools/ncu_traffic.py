#!/usr/bin/env python
"""tools/ncu_traffic.py OUT TAG -- collects dram bytes per launch from the ncu CSVs tools/gpu_evidence.sh wrote
into a JSON list (profiles/ncu_traffic.json is a copy of it; bench.py reads `roofline.traffic` from there).

Two CSV shapes are understood: the raw page of a `--set full` report (one column per metric, a units row) and
the `--metrics ... --csv` log (one row per metric with "Metric Name", "Metric Unit", "Metric Value")."""
import csv
import json
import os
import sys

SCALE = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "": 1.0,
         "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def num(v):
    return float(v.replace(",", ""))


def read_metrics(path):
    """-> {metric name: value in base units} for the FIRST profiled launch of the file, plus 'kernel' and 'grid'."""
    rows = [r for r in csv.reader(open(path, newline="")) if r]
    hdr = next((r for r in rows if "Kernel Name" in r), None)
    if hdr is None:
        return None
    col = {name: k for k, name in enumerate(hdr)}
    body = rows[rows.index(hdr) + 1:]
    out = {}
    if "Metric Name" in col:                       # long format
        first_id = None
        for r in body:
            if len(r) < len(hdr):
                continue
            if first_id is None:
                first_id = r[col["ID"]]
            if r[col["ID"]] != first_id:
                break
            out["kernel"] = r[col["Kernel Name"]]
            out["grid"] = r[col.get("Grid Size", 0)] if "Grid Size" in col else ""
            unit = r[col["Metric Unit"]].lower()
            try:
                out[r[col["Metric Name"]]] = num(r[col["Metric Value"]]) * SCALE.get(unit, 1.0)
            except ValueError:
                pass
    else:                                          # raw page: units row, then one row per launch
        units, data = body[0], body[1]
        out["kernel"] = data[col["Kernel Name"]]
        out["grid"] = data[col["Grid Size"]] if "Grid Size" in col else ""
        for name, k in col.items():
            try:
                out[name] = num(data[k]) * SCALE.get(units[k].lower(), 1.0)
            except (ValueError, IndexError):
                pass
    return out


def entry(path, kind, n, note, tag):
    if not os.path.exists(path):
        return None
    m = read_metrics(path)
    if not m or "dram__bytes_read.sum" not in m:
        return None
    e = {"kernel_kind": kind, "kernel": m.get("kernel"), "grid": m.get("grid"), "n": n, "segments": 32, "chain": 2048,
         "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_written": m["dram__bytes_write.sum"],
         "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
         "algorithmic_bytes_per_launch": 80.0 * n, "note": note,
         "source": f"profiles/{tag}_{os.path.basename(path).split(tag + '_', 1)[-1]} (ncu, dram__bytes_read.sum + dram__bytes_write.sum)"}
    e["traffic_over_algorithmic"] = e["dram_bytes_per_launch"] / e["algorithmic_bytes_per_launch"]
    if "lts__t_bytes.sum" in m:
        e["l2_bytes"] = m["lts__t_bytes.sum"]
    if "gpu__time_duration.sum" in m:
        e["duration_ms_under_ncu"] = m["gpu__time_duration.sum"] * 1e3
    return e


def main():
    out, tag = sys.argv[1], sys.argv[2]
    p = lambda name: os.path.join(out, f"{tag}_{name}")
    entries = [
        entry(p("force_full.csv"), "force", 262144, "default: scratch ring in L2 (unsharded fused step)", tag),
        entry(p("force_4m_dram.csv"), "force", 4194304, "default: scratch ring in L2", tag),
        entry(p("force_noring_dram.csv"), "force_noring", 262144, "MAPC_RING=0: one scratch slot per target block (A/B)", tag),
        entry(p("well_full.csv"), "well", 4194304, "well_step_kernel (CSMain as shipped)", tag),
    ]
    json.dump([e for e in entries if e], sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
