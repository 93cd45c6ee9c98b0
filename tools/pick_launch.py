"""Reads an `ncu --metrics gpu__time_duration.sum --csv` launch list and prints, for a kernel-name
regex, the 0-based index (among launches matching the regex) of the longest launch; with --summary
prints per-kernel totals (count, total ms, share) as a markdown table."""
import csv
import re
import sys


def rows(path):
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        yield r["Kernel Name"], v * scale


def main():
    path = sys.argv[1]
    if sys.argv[2] == "--summary":
        tot = {}
        for k, ms in rows(path):
            k = re.sub(r"\(.*", "", k)
            k = re.sub(r"^void\s+", "", k)
            c = tot.setdefault(k, [0, 0.0, 0.0])
            c[0] += 1; c[1] += ms; c[2] = max(c[2], ms)
        total = sum(v[1] for v in tot.values())
        print("| kernel | launches | total ms | share | longest ms |\n|---|---|---|---|---|")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            print("| `%s` | %d | %.3f | %.1f%% | %.3f |" % (k, v[0], v[1], 100 * v[1] / total, v[2]))
        print("| **total** | %d | %.3f | 100%% | |" % (sum(v[0] for v in tot.values()), total))
        return
    pat = re.compile(sys.argv[2])
    best, idx, i = -1.0, 0, 0
    for k, ms in rows(path):
        if pat.search(k):
            if ms > best:
                best, idx = ms, i
            i += 1
    print(idx)


if __name__ == "__main__":
    main()
