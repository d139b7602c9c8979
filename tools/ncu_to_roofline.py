#!/usr/bin/env python3
"""Digest `ncu -i X.ncu-rep --page raw --csv` exports into profiles/roofline_inputs.json, the only place bench.py takes
executed-instruction and DRAM-traffic counters from (they cannot be measured inside a timed run).

    python tools/ncu_to_roofline.py --leaf profiles/r2_leaf_hash_raw.csv --perms 268435456 [--lde profiles/r2_lde_raw.csv]

--leaf: one k_leaf_hash launch of the bench workload (2^23 leaves x 32 permutations = 268 435 456 permutations)
--lde : all LDE kernels (k_ingest*, k_pass*, k_lde_*) of ONE commit of the bench workload; their DRAM bytes are summed
"""
import argparse
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6,
        "s": 1e9, "second": 1e9, "nsecond": 1.0}


def rows(path):
    """ncu `--page raw --csv` (header, a line of units, then one line per launch) or `--csv --log-file` of a --metrics run
    (one line per launch and metric: "Metric Name", "Metric Unit", "Metric Value") -> [{column: value in base units}]"""
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = list(csv.reader(lines))
    hdr = rd[0]
    if "Metric Name" in hdr:  # long format: pivot
        out = {}
        for r in rd[1:]:
            d = dict(zip(hdr, r))
            row = out.setdefault(d["ID"], {"Kernel Name": d["Kernel Name"]})
            row[d["Metric Name"]] = float(d["Metric Value"].replace(",", "")) * UNIT.get(d["Metric Unit"], 1.0)
        return list(out.values())
    units = rd[1]
    res = []
    for r in rd[2:]:
        if len(r) != len(hdr):
            continue
        d = {}
        for k, u, v in zip(hdr, units, r):
            try:
                d[k] = float(v.replace(",", "")) * UNIT.get(u, 1.0)
            except ValueError:
                d[k] = v
        res.append(d)
    return res


def num(x):
    return float(x)


ap = argparse.ArgumentParser()
ap.add_argument("--leaf")
ap.add_argument("--perms", type=int, default=268435456)
ap.add_argument("--lde")
ap.add_argument("--lde-launches", type=int, default=7, help="launches of ONE commit (the capture holds warm-up commits too: the last N are used)")
args = ap.parse_args()
out_path = os.path.join(ROOT, "profiles", "roofline_inputs.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
if args.leaf:
    r = [x for x in rows(args.leaf) if "k_leaf_hash" in x.get("Kernel Name", "")][0]
    inst = num(r["smsp__inst_executed.sum"])
    out["k_leaf_hash"] = {"source": os.path.relpath(args.leaf, ROOT), "permutations": args.perms,
                          "warp_instructions": inst, "thread_instructions_per_permutation": round(inst * 32 / args.perms, 1),
                          "dram_bytes": num(r["dram__bytes_read.sum"]) + num(r["dram__bytes_write.sum"]),
                          "gpu_time_ns": num(r["gpu__time_duration.sum"])}
if args.lde:
    rs = [x for x in rows(args.lde) if any(k in x.get("Kernel Name", "") for k in ("k_ingest", "k_pass", "k_lde", "k_scale", "k_ntt"))][-args.lde_launches:]
    out["lde"] = {"source": os.path.relpath(args.lde, ROOT), "launches": len(rs),
                  "dram_bytes": sum(num(x["dram__bytes_read.sum"]) + num(x["dram__bytes_write.sum"]) for x in rs),
                  "gpu_time_ns": sum(num(x["gpu__time_duration.sum"]) for x in rs)}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
