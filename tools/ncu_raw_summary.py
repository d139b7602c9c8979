"""One line per launch from `ncu -i X.ncu-rep --page raw --csv`: time, registers, occupancy, issue, pipes, DRAM, top stalls."""
import csv
import sys


def f(d, k):
    try:
        return float(d.get(k, "0").replace(",", ""))
    except ValueError:
        return 0.0


def main(path, pat=""):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if pat and pat not in d["Kernel Name"]:
            continue
        st = sorted(((f(d, h), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:4]
        print(f"{d['ID']:>3} {d['Kernel Name'][:34]:34s} grid {d['Grid Size']:>14s} blk {d['Block Size']:>11s} {f(d, 'gpu__time_duration.sum'):9.1f} {rows[1][hdr.index('gpu__time_duration.sum')]}"
              f" regs {d.get('launch__registers_per_thread', '?'):>3} warps% {f(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}"
              f" issue% {f(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} fmaH% {f(d, 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'):5.1f}"
              f" alu% {f(d, 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'):5.1f} lsu% {f(d, 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):5.1f}"
              f" dram% {f(d, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} inst {f(d, 'smsp__inst_executed.sum'):.3e} | " + " ".join(f"{n}={v:.2f}" for v, n in st))


if __name__ == "__main__":
    main(*sys.argv[1:])
