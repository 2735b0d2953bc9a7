"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (name, grid, block)."""
import collections, csv, re, sys

def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
        key = (name[:64], row.get("Grid Size"), row.get("Block Size"))
        agg[key][0] += 1
        agg[key][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {n} launches, {tot:.1f} us total (serialised, cold-cache ncu timings: compare shares)")
    print(f"{'total_us':>11} {'share':>6} {'n':>5} {'avg_us':>9}  kernel grid block")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:11.1f} {100 * v[1] / tot:5.1f}% {v[0]:5d} {v[1] / v[0]:9.1f}  {k[0]} {k[1]} {k[2]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
