"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): markdown table of the LAST
`--last` launches (one step), kernels of this library only unless --all.   python tools/summarize_launches.py file.csv"""
import argparse
import csv
import re
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--last", type=int, default=0, help="only the last N launches (0 = all)")
    ap.add_argument("--all", action="store_true", help="include kernels that are not this library's")
    a = ap.parse_args()
    rows = [r for r in csv.reader(open(a.csv)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = []
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)     # -> us
        launches.append((re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("tvg::", "").replace("(anonymous namespace)::", ""), v))
    if a.last:
        launches = launches[-a.last:]
    if not a.all:
        launches = [l for l in launches if not l[0].startswith("at::")]
    tot = sum(v for _, v in launches)
    agg = {}
    for k, v in launches:
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += v
    print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% |" % (k, n, v, 100 * v / tot))
    print("| total | %d | %.1f | |" % (len(launches), tot))


if __name__ == "__main__":
    sys.exit(main())
