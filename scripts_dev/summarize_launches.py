"""ncu launch list (csv, --metrics gpu__time_duration.sum) -> markdown table of per-kernel totals and shares.
usage: summarize_launches.py <launches.csv> <title> <command line> > summary.md"""
import csv, sys, collections
path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
h = rows[0]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
tot = 0.0
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * (1e3 if r[ui] == "ms" else 1.0)  # -> us
    name = r[ki].split("(")[0][:80]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"# {title}\n\n`{cmd}`\n({len(rows)-1} launches; cold-cache, serialised under ncu: compare SHARES with `kernel_share_of_step` of the bench line, not absolutes).\n")
print("| kernel | launches | total us | share | us/launch |\n|---|---:|---:|---:|---:|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {n} | {c} | {t:.1f} | {100*t/tot:.1f}% | {t/c:.1f} |")
