"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv, sys, re, collections
path = sys.argv[1]
rows = [l for l in open(path) if l.startswith('"')]
rd = csv.DictReader(rows)
tot = collections.OrderedDict(); cnt = collections.Counter()
for r in rd:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
    name = re.sub(r"at::.*?elementwise_kernel.*", "torch elementwise/fill", name)
    ns = float(r["Metric Value"].replace(",", ""))
    tot[name] = tot.get(name, 0.0) + ns; cnt[name] += 1
s = sum(tot.values())
print(f"{'kernel':55s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:55]:55s} {cnt[k]:8d} {v/1e6:10.3f} {v/cnt[k]/1e3:10.1f} {100*v/s:6.1f}%")
print(f"{'sum':55s} {sum(cnt.values()):8d} {s/1e6:10.3f}")
