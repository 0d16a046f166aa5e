"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count / avg / share."""
import collections, csv, sys
def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"][:72]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print(f"{k:72s} n={len(v):3d} avg={sum(v)/len(v):8.1f}us min={min(v):7.1f} max={max(v):7.1f} share={100*sum(v)/tot:5.1f}%")
    print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
if __name__ == "__main__":
    main(sys.argv[1])
