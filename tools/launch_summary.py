"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total
time and share.  Usage: python tools/launch_summary.py launches.csv [out.md] [title]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("clskd::<unnamed>::", "")
        rows.append((name, float(row["Metric Value"].replace(",", "")), row["Grid Size"], row["ID"]))
    return rows


def main():
    rows = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, v, _, _ in rows:
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v for _, v in agg.values())
    out = ["| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        out.append("| `%s` | %d | %.3f | %.1f%% |" % (k[:95], c, v / 1e6, 100 * v / tot))
    head = "%d launches, %.1f ms of kernel time" % (len(rows), tot / 1e6)
    text = "\n".join(out)
    if len(sys.argv) > 2:
        title = sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]
        with open(sys.argv[2], "w") as f:
            f.write("# %s\n\n%s (cold-cache, serialised: compare shares).\n\n%s\n" % (title, head, text))
    print(head)
    print(text)


if __name__ == "__main__":
    main()
