"""`ncu -i rep --page source --csv` -> the N most-sampled SASS instructions with their dominant stall reasons.

  python tools/ncu_source_top.py source.csv out.csv [N]
"""
import csv
import sys


def main():
    src, out = sys.argv[1:3]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = [r for r in csv.reader(open(src)) if r]
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    head, body = rows[hi], rows[hi + 1:]
    i_src, i_smp, i_exec = head.index("Source"), head.index("# Samples"), head.index("Instructions Executed")
    stalls = [(h, i) for i, h in enumerate(head) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[i_smp] or 0) for r in body)
    recs = []
    for k, r in enumerate(body):
        s = int(r[i_smp] or 0)
        if s:
            top = sorted(((int(r[i] or 0), h) for h, i in stalls), reverse=True)[:2]
            recs.append((s, k, r[i_exec], r[i_src].strip(), ";".join(f"{h}={v}" for v, h in top if v)))
    recs.sort(reverse=True)
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["sass_index", "samples", "pct_of_samples", "instructions_executed", "sass", "top_stalls"])
        for s, k, ex, sass, st in recs[:n]:
            w.writerow([k, s, round(100.0 * s / max(total, 1), 2), ex, sass, st])
    print(f"{total} samples, top {min(n, len(recs))} -> {out}")


if __name__ == "__main__":
    main()
