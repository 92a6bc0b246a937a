"""Per-kernel DRAM traffic of one forward+backward step from an ncu metrics pass:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/traffic_fb.csv python bench.py --steps 1 --warmup 3 --bwd-steps 1 --no-cpu-baseline ...
  python scripts/traffic_table.py gpurun_out/traffic_fb.csv > profiles/r1_dram_traffic_fwd_bwd.txt
"maps" = (read + written bytes) / one 32-channel fp32 map of the bench batch (629 MB): what each kernel really moves,
to hold against its algorithmic map count."""
import csv, io, re, sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
by = {}
for r in csv.DictReader(io.StringIO("".join(lines))):
    i = int(r["ID"]); d = by.setdefault(i, {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if u in ("ns", "nsecond") else v * {"us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
    else:
        d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
ids = sorted(by)
short = lambda n: re.sub(r"\(.*$", "", re.sub(r"paif::", "", re.sub(r"^void\s+", "", n)))
names = [short(by[i]["name"]) for i in ids]
end = max(k for k, n in enumerate(names) if n.startswith("stem_backward_kernel"))
start = [k for k, n in enumerate(names[:end]) if n.startswith("stem_forward")][-2]
MAP = 16 * 480 * 640 * 32 * 4.0
print("one forward + backward-to-input step, batch 16 x 480x640 (ncu, cold cache, serialised launches)")
print("%-46s %8s %7s %7s %7s" % ("kernel", "us", "rd_GB", "wr_GB", "maps"))
tot = [0.0, 0.0, 0.0]
for k in range(start, end + 1):
    d = by[ids[k]]
    rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    tot[0] += d["us"]; tot[1] += rd; tot[2] += wr
    print("%-46s %8.1f %7.3f %7.3f %7.2f" % (names[k][:46], d["us"], rd / 1e9, wr / 1e9, (rd + wr) / MAP))
print("%-46s %8.1f %7.3f %7.3f %7.2f" % ("total", tot[0], tot[1] / 1e9, tot[2] / 1e9, (tot[1] + tot[2]) / MAP))
