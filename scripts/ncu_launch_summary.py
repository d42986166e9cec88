"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (run here, no GPU needed)."""
import collections, csv, re, sys
path = sys.argv[1]
skip_idle = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0  # launches shorter than this many us are early-outs
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    act = [x for x in v if x >= skip_idle] or v
    print(f"{k[:64]:64s} n={len(v):3d} total={sum(v):10.1f}us share={100*sum(v)/tot:5.1f}%  mean(active {len(act)})={sum(act)/len(act):9.1f}us")
