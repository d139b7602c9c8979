"""Golden Merkle roots of the bench workload (bench_workload.trace_numpy: 2^k x 256, seed 0xB200 + rank), computed on the
host by the CPU oracle: the tuned implementation oracle/fast_commit.c (itself checked word for word against the slow
restatement in tests/test_oracle_fast_commit.py) at the full 2^22 x 256 size, and BOTH implementations at 2^16 x 256.
Writes tests/golden/bench_roots.json.  ~1 minute and ~17 GiB of host memory for the full size.
    python scripts/gen_bench_root.py [--log-rows 22]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench_workload as W  # noqa: E402
import oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log-rows", type=int, default=22)
ap.add_argument("--ranks", type=int, default=2, help="seeds 0xB200 + r for r < ranks (bench.py --gpus N uses one per rank)")
args = ap.parse_args()
out = {"generator": "bench_workload.trace_numpy", "cols": 256, "roots": {}}
small = W.trace_numpy(1 << 16, 256)
slow = oracle.PcsData([small]).root
fast = oracle.fast_pcs_commit(small)[0]
assert (slow == fast).all()
out["roots"]["log_rows=16,seed=0xB200"] = [int(x) for x in slow]
for r in range(args.ranks):
    t = time.time()
    m = W.trace_numpy(1 << args.log_rows, 256, seed=W.SEED + r)
    root, _, ph = oracle.fast_pcs_commit(m)
    print(f"seed 0xB200+{r}: root {root.tolist()}  ({time.time() - t:.1f} s, phases {ph})", flush=True)
    out["roots"][f"log_rows={args.log_rows},seed=0xB200+{r}"] = [int(x) for x in root]
    del m
oracle.fast_release()
path = os.path.join(ROOT, "tests", "golden", "bench_roots.json")
old = {}
if os.path.exists(path):
    old = json.load(open(path)).get("roots", {})
old.update(out["roots"])
out["roots"] = old
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)
