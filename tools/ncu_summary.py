"""Summarise an `ncu -i X.ncu-rep --page raw --csv` dump as a markdown table (one row per captured kernel).
usage: python tools/ncu_summary.py raw.csv > profiles/<name>.md"""
import csv, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("time us", "gpu__time_duration.sum", 1e-3), ("DRAM read MB", "dram__bytes_read.sum", None), ("DRAM write MB", "dram__bytes_write.sum", None),
        ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
        ("tcgen05 bf16 ops % of peak", "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", 1),
        ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1), ("DRAM throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("regs/thread", "launch__registers_per_thread", 1), ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("L2 hit %", "lts__t_sector_hit_rate.pct", 1), ("IPC", "sm__inst_executed.avg.per_cycle_elapsed", 1)]
units = rows[1]
print("| kernel | " + " | ".join(c[0] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in data:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("f3d::", "")
    out = []
    for label, key, scale in cols:
        if key not in ix or r[ix[key]] in ("", "n/a", "no data"):
            out.append("-")
            continue
        v = float(r[ix[key]].replace(",", ""))
        if scale is None:  # bytes with a unit column
            u = units[ix[key]].lower()
            v *= {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
        else:
            if label == "time us":
                u = units[ix[key]].lower()
                v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
            else:
                v *= scale
        out.append("%.2f" % v)
    print("| " + name + " | " + " | ".join(out) + " |")
