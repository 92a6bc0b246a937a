"""Turn the raw ncu outputs of scripts/prof_r1.sh (gpurun_out/) into the summaries committed under profiles/.

  python scripts/summarize_profiles.py gpurun_out profiles r1

* launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum --csv)  -> <tag>_ncu_launches.csv (our kernels only)
                                                                        <tag>_ncu_launch_shares.txt (one forward step)
* prof_conv_<tag>.ncu-rep, prof_gf_<tag>.ncu-rep (ncu --set full)     -> <tag>_ncu_full_summary.{txt,json}
                                                                        conv_traffic.json (read by bench.py)
Needs the `ncu` CLI to read the .ncu-rep files (no GPU)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

OURS = re.compile(r"paif::|conv_tc_kernel|gf_|stem_|dilconv|spa_|eca_|out_forward|out_border|channel_pool|confusion")


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"paif::", "", name)
    return re.sub(r"\(.*$", "", name)


def read_launch_csv(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            us = v / 1e3 if unit in ("ns", "nsecond") else v * {"us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
            rows.append((int(r["ID"]), r["Kernel Name"], us))
    return rows


def launch_shares(rows, out_csv, out_txt):
    ours = [(i, short(n), us) for (i, n, us) in rows if OURS.search(n)]
    with open(out_csv, "w") as f:
        f.write("id,kernel,duration_us\n")
        for i, n, us in ours:
            f.write('%d,"%s",%.2f\n' % (i, n, us))
    # one forward step = from one stem_forward pair to the next out_border kernel
    names = [n for _, n, _ in ours]
    ends = [k for k, n in enumerate(names) if n.startswith("out_border_kernel")]
    if len(ends) < 2:
        step = ours
    else:
        step = ours[ends[-2] + 1: ends[-1] + 1]
    agg = {}
    for _, n, us in step:
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + us)
    tot = sum(t for _, t in agg.values())
    with open(out_txt, "w") as f:
        f.write("ncu launch list (gpu__time_duration.sum, --clock-control none), one resident forward step of bench.py\n")
        f.write("cold-cache, serialised launches: compare SHARES with the CUDA-event table, not absolutes\n\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-70s x%-3d %8.3f ms %5.1f%%\n" % (n, c, t / 1e3, 100 * t / tot))
        f.write("total %.3f ms over %d launches\n" % (tot / 1e3, len(step)))


def read_rep(path):
    if path.endswith(".csv"):                 # `ncu -i x.ncu-rep --page raw --csv` already run on the GPU box
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines(True) if l.startswith('"')]
    rd = csv.reader(io.StringIO("".join(lines)))
    head, units = next(rd), next(rd)
    return [dict(zip(head, r)) for r in rd], dict(zip(head, units))


def pick(row, units, key, want_unit=None):
    for k, v in row.items():
        if k.endswith(key) and v not in ("", "n/a"):
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            u = units.get(k, "")
            if want_unit == "byte":
                x *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)
            if want_unit == "us":
                x *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
            return x
    return float("nan")


def full_summary(reps, out_txt, out_json, traffic_json, note):
    recs = []
    for path in reps:
        if not os.path.exists(path):
            continue
        rows, units = read_rep(path)
        for r in rows:
            recs.append({
                "kernel": short(r["Kernel Name"]), "grid": r.get("Grid Size", ""),
                "dur_us": pick(r, units, "gpu__time_duration.sum", "us"),
                "dram_read_bytes": pick(r, units, "dram__bytes_read.sum", "byte"),
                "dram_write_bytes": pick(r, units, "dram__bytes_write.sum", "byte"),
                "dram_pct": pick(r, units, "dram__bytes_read.sum.pct_of_peak_sustained_elapsed")
                + pick(r, units, "dram__bytes_write.sum.pct_of_peak_sustained_elapsed"),
                "tensor_pct": pick(r, units, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "warps_pct": pick(r, units, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "issue_pct": pick(r, units, "sm__inst_issued.avg.pct_of_peak_sustained_active"),
                "regs": pick(r, units, "launch__registers_per_thread"),
            })
    with open(out_json, "w") as f:
        json.dump({"note": note, "kernels": recs}, f, indent=1)
    with open(out_txt, "w") as f:
        f.write(note + "\n")
        f.write("%-38s %8s %8s %8s %7s %8s %7s %7s %5s  %s\n" % ("kernel", "dur_us", "rd_GB", "wr_GB", "dram%", "tensor%", "warps%", "issue%", "regs", "grid"))
        for r in recs:
            f.write("%-38s %8.1f %8.3f %8.3f %7.1f %8.1f %7.1f %7.1f %5.0f  %s\n" % (
                r["kernel"][:38], r["dur_us"], r["dram_read_bytes"] / 1e9, r["dram_write_bytes"] / 1e9, r["dram_pct"],
                r["tensor_pct"], r["warps_pct"], r["issue_pct"], r["regs"], r["grid"]))
    # the 32-cout convolutions (what bench.py's roofline counts); the single-output stem_out instance (<..., 16>) is listed but not averaged
    conv = [r for r in recs if r["kernel"].startswith("conv_tc_kernel") and not r["kernel"].rstrip().endswith(", 16>")]
    if conv and traffic_json:
        per = sum(r["dram_read_bytes"] + r["dram_write_bytes"] for r in conv) / len(conv)
        old = {}
        if os.path.exists(traffic_json):
            old = json.load(open(traffic_json))
        old.update({"kernel": "conv_tc_kernel (%d launches of one forward step, batch 16 x 480x640)" % len(conv),
                    "dram_bytes_per_launch": per,
                    "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, %s" % os.path.basename(out_json)})
        json.dump(old, open(traffic_json, "w"), indent=1)


def main():
    src, dst, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    lc = os.path.join(src, "launches_%s.csv" % tag)
    if os.path.exists(lc):
        launch_shares(read_launch_csv(lc), os.path.join(dst, "%s_ncu_launches.csv" % tag),
                      os.path.join(dst, "%s_ncu_launch_shares.txt" % tag))
    r2 = [(k, os.path.join(src, "prof_%s_%s.raw.csv" % (k, tag))) for k in ("fwd", "bwd", "bf16")]
    if any(os.path.exists(f) for _, f in r2):            # scripts/prof_r2.sh layout: one capture per step kind
        what = {"fwd": "one forward step, batch 16 x 480x640, fp32 storage (scripts/one_step.py fwd)",
                "bwd": "one forward + backward-to-input step, batch 8 x 480x640 (scripts/one_step.py bwd)",
                "bf16": "one forward step, batch 16 x 480x640, bf16 storage (scripts/one_step.py bf16)"}
        for k, f in r2:
            if os.path.exists(f):
                full_summary([f], os.path.join(dst, "%s_ncu_full_%s.txt" % (tag, k)), os.path.join(dst, "%s_ncu_full_%s.json" % (tag, k)),
                             os.path.join(dst, "conv_traffic.json") if k == "fwd" else None,
                             "ncu --set full --clock-control none, every kernel of %s" % what[k])
        return
    full_summary([os.path.join(src, "prof_conv_%s.ncu-rep" % tag), os.path.join(src, "prof_gf_%s.ncu-rep" % tag)],
                 os.path.join(dst, "%s_ncu_full_summary.txt" % tag), os.path.join(dst, "%s_ncu_full_summary.json" % tag),
                 os.path.join(dst, "conv_traffic.json"),
                 "ncu --set full --clock-control none, bench.py forward step (batch 16 x 480x640), %s kernels" % tag)


if __name__ == "__main__":
    main()
