"""Summarise an `ncu -i X.ncu-rep --page raw --csv` dump as a markdown table (one row per captured kernel): time, DRAM bytes, the
pipe / memory-level utilisations and the two dominant warp-stall reasons -- i.e. which resource bounds each kernel.

usage: python tools/ncu_summary.py raw.csv [--traffic-json profiles/ncu_traffic.json --capture "<what was captured>"] > profiles/<name>.md
--traffic-json writes dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel (first instance) for bench.py's roofline.traffic."""
import csv, json, sys

args = sys.argv[1:]
traffic_json = capture = None
if "--traffic-json" in args:
    i = args.index("--traffic-json"); traffic_json = args[i + 1]; del args[i:i + 2]
if "--capture" in args:
    i = args.index("--capture"); capture = args[i + 1]; del args[i:i + 2]
rows = list(csv.reader(open(args[0])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("time us", "gpu__time_duration.sum", "time"), ("DRAM read MB", "dram__bytes_read.sum", "bytes"), ("DRAM write MB", "dram__bytes_write.sum", "bytes"),
        ("DRAM thr %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("L2 thr %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("L1 thr %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
        ("issue slots %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1), ("IPC", "sm__inst_executed.avg.per_cycle_elapsed", 1),
        ("eligible warps / cycle", "smsp__warps_eligible.avg.per_cycle_active", 1), ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("regs", "launch__registers_per_thread", 1), ("L2 hit %", "lts__t_sector_hit_rate.pct", 1)]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]


def val(r, key, kind):
    if key not in ix or r[ix[key]] in ("", "n/a", "no data"):
        return None
    v = float(r[ix[key]].replace(",", ""))
    u = units[ix[key]].lower()
    if kind == "bytes":
        return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
    if kind == "time":
        return v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
    return v * kind


print("| kernel | " + " | ".join(c[0] for c in cols) + " | top stalls (warps stalled per issue) |")
print("|---|" + "---|" * (len(cols) + 1))
traffic = {}
for r in data:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("f3d::", "")
    out = []
    for label, key, kind in cols:
        v = val(r, key, kind)
        out.append("-" if v is None else ("%.2f" % v))
    st = sorted(((float(r[ix[h]].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stalls
                 if r[ix[h]] not in ("", "n/a", "no data")), reverse=True)[:2]
    print("| " + name + " | " + " | ".join(out) + " | " + ", ".join("%s %.2f" % (n, v) for v, n in st) + " |")
    rd, wr = val(r, "dram__bytes_read.sum", "bytes"), val(r, "dram__bytes_write.sum", "bytes")
    if rd is not None and name not in traffic:
        key = name
        if name.startswith("post_tc_kernel<0>"): key = "post_tc_kernel<detector>"
        if name.startswith("post_tc_kernel<1>"): key = "post_tc_kernel<descriptor>"
        key = key.split("<")[0] if key.startswith(("fps_group_kernel", "lin_tc", "bn_", "wgrad")) else key
        traffic.setdefault(key, {"dram_bytes": round((rd + (wr or 0.0)) * 1e6)})
if traffic_json:
    json.dump({"capture": capture or args[0], "unit": "dram__bytes_read.sum + dram__bytes_write.sum per launch, bytes", "kernels": traffic},
              open(traffic_json, "w"), indent=1)
