"""Summaries of ncu reports for profiles/: launch list shares and key raw metrics."""
import csv, collections, subprocess, sys, json, io

def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit() or (r and r[0] == "ID")]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        t = float(r[vi].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        if unit == "us": t *= 1e3
        elif unit == "ms": t *= 1e6
        elif unit == "s": t *= 1e9
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    out = ["kernel | launches | total ms | share", "---|---|---|---"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%s | %d | %.3f | %.1f%%" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))
    out.append("all | %d | %.3f | 100%%" % (sum(v[0] for v in agg.values()), tot / 1e6))
    return "\n".join(out)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]

def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for i, h in enumerate(hdr):
        if h in WANT or h == "Kernel Name":
            d[h] = (vals[i], units[i])
    return d

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launches(sys.argv[2]))
    else:
        d = raw(sys.argv[2])
        for k in ["Kernel Name"] + WANT:
            if k in d:
                print("%s | %s %s" % (k, d[k][0], d[k][1]))
