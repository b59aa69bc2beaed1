"""Top stalled SASS instructions of a kernel from an ncu report: python tools/ncu_top.py rep.ncu-rep regex [n]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--launch-skip", "0", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
tot = sum(int(r["# Samples"] or 0) for r in rows)
print("total samples", tot, "instructions", len(rows))
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
agg = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
print("stall mix:", {k: f"{100*v/max(1,sum(agg.values())):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for i, r in sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"] or 0))[:n]:
    s = int(r["# Samples"] or 0)
    top = sorted(((int(r[c] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{100*s/tot:5.1f}% idx={i:5d} exec={r['Instructions Executed']:>8} {r['Source'].strip()[:90]:90s} {top}")
