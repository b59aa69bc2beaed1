"""Summarise an ncu launch list (gpu__time_duration) of bench.py: last step, share per kernel.
usage: python tools/launch_summary.py launches.csv n_steps_total"""
import csv, re, sys, collections
path, nsteps = sys.argv[1], int(sys.argv[2])
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
per = len(rows) // nsteps
last = rows[-per:]
agg = collections.defaultdict(lambda: [0.0, 0])
for x in last:
    k = re.sub(r"\(.*", "", re.sub(r"<.*", "", x["Kernel Name"]))[:80]
    agg[k][0] += float(x["Metric Value"].replace(",", "")); agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"{len(rows)} launches total, {per} per step; last step: {tot/1000:.1f} us of kernel time")
print("| share | us/step | launches | kernel |\n|---:|---:|---:|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| {v[0]/tot*100:.1f}% | {v[0]/1000:.1f} | {v[1]} | `{k}` |")
