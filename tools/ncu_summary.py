"""Condense `ncu --csv --metrics ...` logs into one line per kernel launch:  python tools/ncu_summary.py gpurun_out/x_*.csv"""
import csv
import sys

TO_BASE = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
           "hz": 1.0, "Khz": 1e3, "Mhz": 1e6, "Ghz": 1e9, "%": 1.0}

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    if not rows:
        print(path, "(empty)")
        continue
    hdr = rows[0]
    ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    ui = hdr.index("Metric Unit")
    out = {}
    for r in rows[1:]:
        name = r[ki].split("<")[1].split(">")[0][-28:] if "<" in r[ki] else r[ki][:28]
        out.setdefault((int(r[idi]), name), {})[r[mi]] = float(r[vi].replace(",", "")) * TO_BASE.get(r[ui], 1.0)
    print(path)
    for (i, name), m in sorted(out.items()):
        print(f"   {name:30s} {m.get('gpu__time_duration.sum', 0) * 1e3:8.3f} ms  "
              f"read {m.get('dram__bytes_read.sum', 0) * 1e-9:7.2f} GB  write {m.get('dram__bytes_write.sum', 0) * 1e-9:6.2f} GB  "
              f"L2 hit {m.get('lts__t_sector_hit_rate.pct', 0):5.1f} %  "
              f"clk {m.get('sm__cycles_elapsed.avg.per_second', 0) * 1e-9:5.3f} GHz  "
              f"tensor {m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f} %")
