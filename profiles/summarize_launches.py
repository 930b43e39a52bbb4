"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [skip_launches] > profiles/<name>.md"""
import collections, csv, re, sys

def main(path, skip=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"])) for r in csv.DictReader(lines)
            if r.get("Metric Name") == "gpu__time_duration.sum"]
    rows = rows[skip:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        k = re.sub(r"\(.*", "", k)
        k = re.sub(r"^void ", "", k)[:90]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    print(f"launches: {len(rows)}  total device time: {tot/1e6:.3f} ms (ncu-serialised, cold-cache; compare shares)\n")
    print("| share | ms | launches | kernel |\n|---:|---:|---:|---|")
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if v / tot < 0.0005:
            continue
        print(f"| {v/tot*100:.2f}% | {v/1e6:.3f} | {c} | `{k}` |")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
