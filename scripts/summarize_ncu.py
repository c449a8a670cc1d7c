#!/usr/bin/env python3
"""Summarises an `ncu --metrics ... --csv` launch log (one row per launch and metric) into per-kernel totals:
launches, device time, DRAM bytes, and the time-weighted tensor-pipe / issue / DRAM utilisation.
usage: summarize_ncu.py <log.csv> <out.json>"""
import csv
import json
import re
import sys
from collections import defaultdict


def main(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) >= 15 and r[0].isdigit()]
    launches = defaultdict(dict)
    for r in rows:
        kid, name, metric, unit, val = int(r[0]), r[4], r[12], r[13], r[14]
        launches[kid]["name"] = re.sub(r"\(.*", "", name).replace("void ", "").replace("unnamed>::", "").strip()
        try:
            v = float(val.replace(",", ""))
        except ValueError:
            continue
        if unit in ("ns", "nsecond"):
            v /= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e3
        elif unit in ("Mbyte",):
            v *= 1e6
        elif unit in ("Kbyte",):
            v *= 1e3
        elif unit in ("Gbyte",):
            v *= 1e9
        launches[kid][metric] = v
    agg = defaultdict(lambda: defaultdict(float))
    for kid, m in launches.items():
        a = agg[m["name"]]
        t = m.get("gpu__time_duration.sum", 0.0)
        a["launches"] += 1
        a["time_us"] += t
        a["dram_read_bytes"] += m.get("dram__bytes_read.sum", 0.0)
        a["dram_write_bytes"] += m.get("dram__bytes_write.sum", 0.0)
        for k, short in (("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
                         ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_elapsed_pct"),
                         ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                         ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
                         ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct")):
            a["_w_" + short] += m.get(k, 0.0) * t
    res = {}
    for name, a in agg.items():
        t = a["time_us"]
        res[name] = {"launches": int(a["launches"]), "time_us": round(t, 2), "dram_read_bytes": a["dram_read_bytes"], "dram_write_bytes": a["dram_write_bytes"]}
        for k in list(a):
            if k.startswith("_w_"):
                res[name][k[3:] + "_time_weighted"] = round(a[k] / t, 3) if t else 0.0
    total = sum(v["time_us"] for v in res.values())
    res["_total"] = {"time_us": round(total, 2), "launches": sum(v["launches"] for v in res.values()),
                     "dram_bytes": sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in res.values())}
    conv = [v for k, v in res.items() if k.startswith("conv_halo") or k.startswith("conv_tc")]
    ct = sum(v["time_us"] for v in conv)
    if ct:
        res["_conv_family"] = {"time_us": round(ct, 2), "launches": sum(v["launches"] for v in conv),
                               "dram_bytes": sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in conv),
                               "tensor_pipe_active_pct_time_weighted": round(sum(v.get("tensor_pipe_active_pct_time_weighted", 0) * v["time_us"] for v in conv) / ct, 3),
                               "issue_active_pct_time_weighted": round(sum(v.get("issue_active_pct_time_weighted", 0) * v["time_us"] for v in conv) / ct, 3),
                               "dram_throughput_pct_time_weighted": round(sum(v.get("dram_throughput_pct_time_weighted", 0) * v["time_us"] for v in conv) / ct, 3)}
    json.dump(res, open(out, "w"), indent=1)
    for k, v in sorted(res.items(), key=lambda kv: -kv[1].get("time_us", 0)):
        print(f"{k[:60]:60s}", v)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
