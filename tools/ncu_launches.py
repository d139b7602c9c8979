"""Digest an `ncu --csv --metrics ...` log (one row per launch and metric) into one line per launch."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, d = None, collections.OrderedDict()
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            rec = dict(zip(hdr, r))
            k = (int(rec["ID"]), rec["Kernel Name"][:60], rec["Grid Size"], rec["Block Size"])
            d.setdefault(k, {})[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
    for k, v in d.items():
        t = v.get("gpu__time_duration.sum", 0) / 1e6
        b = (v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0)) / 1e9
        inst = v.get("smsp__inst_executed.sum", 0)
        print(f"{k[0]:3d} {k[1]:60s} {k[2]:16s} {k[3]:12s} {t:8.3f} ms {b:7.2f} GB {b / t if t else 0:6.2f} TB/s  warp-inst {inst:.3e}")


if __name__ == "__main__":
    main(sys.argv[1])
