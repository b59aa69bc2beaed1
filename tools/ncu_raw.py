"""Per-launch key metrics from an ncu report: python tools/ncu_raw.py rep.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
hdr = r[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("lts__t_bytes.sum", "l2_bytes_MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]
idx = [(hdr.index(a), b) for a, b in want if a in hdr]
units = r[1]
for row in r[2:]:
    d = {}
    for i, b in idx:
        v = row[i]
        if b == "kernel":
            v = v.split("(")[0][-40:]
        elif b.endswith("_MB"):
            u = units[i]
            f = float(v.replace(",", ""))
            f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
            v = f"{f:.1f}"
        elif b == "us":
            u = units[i]
            f = float(v.replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
            v = f"{f:.1f}"
        else:
            try: v = f"{float(v.replace(',', '')):.1f}"
            except ValueError: pass
        d[b] = v
    print(" ".join(f"{k}={v}" for k, v in d.items()))
