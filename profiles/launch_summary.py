"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, max and share."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    d = collections.defaultdict(list)
    regs = {}
    for r in rows[1:]:
        name = r[ki].split("(")[0][:48]
        if r[mi] == "gpu__time_duration.sum":
            d[name].append(float(r[vi].replace(",", "")))
        elif r[mi] == "launch__registers_per_thread":
            regs[name] = r[vi]
    total = sum(sum(v) for v in d.values())
    print(f"{'kernel':50s} {'n':>5s} {'mean_us':>9s} {'max_us':>9s} {'share':>6s} regs")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:50s} {len(v):5d} {sum(v) / len(v) / 1e3:9.2f} {max(v) / 1e3:9.2f} {sum(v) / total:6.3f} {regs.get(k, '')}")


if __name__ == "__main__":
    main(sys.argv[1])
