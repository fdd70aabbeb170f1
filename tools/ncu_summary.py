"""Condense `ncu --csv --metrics ...` logs into one line per kernel launch:  python tools/ncu_summary.py gpurun_out/x_*.csv"""
import csv
import sys

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    if not rows:
        print(path, "(empty)")
        continue
    hdr = rows[0]
    ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    out = {}
    for r in rows[1:]:
        out.setdefault((int(r[idi]), r[ki].split("<")[1].split(">")[0][-28:] if "<" in r[ki] else r[ki][:28]), {})[r[mi]] = float(r[vi].replace(",", ""))
    print(path)
    for (i, name), m in sorted(out.items()):
        rd, wr = m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0)
        scale = 1e-9 if rd > 1e6 else 1.0  # bytes or already GB
        print(f"   {name:30s} {m.get('gpu__time_duration.sum', 0) * (1e-6 if m.get('gpu__time_duration.sum', 0) > 1e5 else 1):8.3f} ms  "
              f"read {rd * scale:7.2f} GB  write {wr * scale:6.2f} GB  L2 hit {m.get('lts__t_sector_hit_rate.pct', 0):5.1f} %  "
              f"clk {m.get('sm__cycles_elapsed.avg.per_second', 0) * (1e-9 if m.get('sm__cycles_elapsed.avg.per_second', 0) > 1e6 else 1):5.3f} GHz  "
              f"tensor {m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f} %")
