"""Launch list of ONE step from an ncu `--metrics gpu__time_duration.sum --csv` log of `bench.py --workload rgb|sam`:
the kernels between the last two L2 flushes (at::FillFunctor of 65536 blocks) of the timed loop.
usage: python tools/launch_windows.py launches.csv > table.md"""
import collections, csv, re, sys

rows = list(csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith("==")]))
name = lambda r: re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("sanerf::", "")
flush = [i for i, r in enumerate(rows) if "FillFunctor" in r["Kernel Name"] and r["Grid Size"].startswith("(65536")]
wins = [rows[a + 1:b] for a, b in zip(flush[:-1], flush[1:]) if b - a > 5]
win = wins[-1]
at = [sum(1 for r in w if "at::" in r["Kernel Name"]) for w in wins]
us = [float(r["Metric Value"].replace(",", "")) / 1e3 for r in win]
tot = sum(us)
print(f"{len(win)} launches, {tot:.0f} us serialised (cold caches, one kernel at a time: compare shares, not absolutes); "
      f"`at::` launches per window of the timed loop: {at}\n")
print("| # | kernel | grid | us | share |\n|---:|---|---|---:|---:|")
for i, (r, u) in enumerate(zip(win, us)):
    print(f"| {i} | `{name(r)}` | {r['Grid Size']} | {u:.1f} | {u / tot * 100:.1f} % |")
agg = collections.defaultdict(lambda: [0.0, 0])
for r, u in zip(win, us):
    agg[name(r)][0] += u; agg[name(r)][1] += 1
print("\nBy kernel:\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|")
for k, (u, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| `{k}` | {n} | {u:.1f} | {u / tot * 100:.1f} % |")
