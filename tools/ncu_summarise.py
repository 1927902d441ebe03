"""Turns `ncu -i step.ncu-rep --page raw --csv` of one profiled step (tools/prof_step.py) into the per-launch table kept
under profiles/: one row per kernel launch, named after the engine op that issued it.

  python tools/ncu_summarise.py raw.csv ops.txt out.csv

ops.txt = the stderr of tools/prof_step.py (lines "op <i> <name>"); launch i of the profiled step is op i.
"""
import csv
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
]


def main():
    raw, ops_txt, out = sys.argv[1:4]
    ops = {}
    for line in open(ops_txt):
        p = line.split()
        if len(p) >= 3 and p[0] == "op" and p[1].isdigit():
            ops[int(p[1])] = p[2]
    rows = [r for r in csv.reader(open(raw)) if r]
    # the raw page: a header row, a units row, then one row per launch
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    head, units, body = rows[hi], rows[hi + 1], rows[hi + 2:]
    kn = head.index("Kernel Name")
    cols = [(k, head.index(k)) for k in KEEP if k in head]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "op", "Kernel Name"] + [k for k, _ in cols])
        w.writerow(["", "", ""] + [units[i] for _, i in cols])
        for n, r in enumerate(body):
            name = r[kn].split("(")[0]
            w.writerow([n, ops.get(n, ""), name] + [r[i] for _, i in cols])
    print(f"{len(body)} launches -> {out}")


if __name__ == "__main__":
    main()
