"""`ncu -i rep --page source --csv --print-source cuda,sass` -> per kernel, the CUDA source lines with the most warp-stall
samples (and the instructions executed on them).

  python tools/ncu_lines_top.py cuda_sass.csv [N]
"""
import csv
import sys


def main():
    src = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    kern, fpath, head = None, None, None
    per = {}
    for r in csv.reader(open(src)):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            kern = r[1].split("(")[0]
            continue
        if r[0] == "Line No":
            head = r
            continue
        if head is None or not r[0].isdigit():
            continue
        d = per.setdefault(kern, {})
        off = len(r) - len(head)                       # inline-asm source lines carry quotes / commas of their own
        key = (fpath, int(r[0]), ",".join(r[1:2 + off]).strip())
        try:
            smp = int(r[6 + off] or 0)
            ex = int(r[7 + off] or 0)
        except ValueError:
            continue
        a = d.setdefault(key, [0, 0])
        a[0] += smp
        a[1] += ex
    for k, d in per.items():
        tot = sum(v[0] for v in d.values())
        print(f"== {k}: {tot} samples")
        for (f, ln, text), (s, ex) in sorted(d.items(), key=lambda t: -t[1][0])[:n]:
            print(f"  {100.0 * s / max(tot, 1):5.1f}%  {ex:>9}  {f}:{ln}  {text[:110]}")


if __name__ == "__main__":
    main()
