"""Compact per-launch table from an `ncu --page raw --csv` export:  python tools/ncu_table.py file_raw.csv [name-filter]
(`--traffic <size>`: the conv_tc DRAM-traffic summary bench.py reports as roofline.traffic)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
flt = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
hdr, units, data = rows[0], rows[1], rows[2:]
if "--traffic" in sys.argv:  # python tools/ncu_table.py raw.csv --traffic <size>: JSON for profiles/conv_traffic.json (bench.py reads it)
    import json

    size = sys.argv[sys.argv.index("--traffic") + 1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = [float(r[ir].replace(",", "")) * scale.get(units[ir], 1) + float(r[iw].replace(",", "")) * scale.get(units[iw], 1)
           for r in data if "conv_tc_kernel" in r[ik]]
    print(json.dumps({size: {"dram_bytes_per_launch": sum(tot) / max(len(tot), 1), "launches": len(tot),
                             "dram_bytes_per_iteration": sum(tot), "source": "ncu --set full, every conv_tc launch of one feval"}}))
    sys.exit(0)
cols = [("Kernel Name", "kernel", 34), ("gpu__time_duration.sum", "us", 9),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 8),
        ("dram__bytes_read.sum", "rdMB", 9), ("dram__bytes_write.sum", "wrMB", 9),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 7),
        ("lts__t_bytes.sum", "L2MB", 9), ("launch__registers_per_thread", "regs", 5),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 6),
        ("launch__grid_size", "grid", 7), ("smsp__cycles_active.avg", "cycles", 10)]
idx = [(hdr.index(c), n, w) for c, n, w in cols if c in hdr]
def mb(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
print(" ".join(n.rjust(w) for _, n, w in idx))
for r in data:
    if flt and flt not in r[hdr.index("Kernel Name")]:
        continue
    out = []
    for i, n, w in idx:
        v = r[i]
        if n == "kernel":
            v = v.replace("void ", "").replace("maua::<unnamed>::", "").replace("unnamed>::", "").split("(")[0][:w]
        elif n in ("rdMB", "wrMB", "L2MB"):
            v = f"{mb(v, units[i]):.1f}"
        elif n == "us":
            f = float(v.replace(",", ""))
            f *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[i], 1)
            v = f"{f:.1f}"
        else:
            try:
                v = f"{float(v.replace(',', '')):.1f}"
            except ValueError:
                pass
        out.append(v.rjust(w))
    print(" ".join(out))
