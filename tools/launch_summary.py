"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): kernel,launches,total_ns,share,mean_ns."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
h = rows[hdr]; ik = h.index("Kernel Name"); iv = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    try:
        agg.setdefault(r[ik].split("(")[0][:70], []).append(float(r[iv].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
print("kernel,launches,total_ns,share,mean_ns")
for k, v in agg.items():
    print(f"{k},{len(v)},{int(sum(v))},{sum(v) / tot:.3f},{int(sum(v) / len(v))}")
