"""One line per kernel from `ncu -i X.ncu-rep --page raw --csv`: duration, DRAM bytes and % of peak, issue-slot use,
occupancy, L2 hit rate — launches of the same kernel summed.  usage: python tools/ncu_raw_table.py raw.csv"""
import csv, collections, sys
csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def f(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return default
def scaled(r, name):
    v = f(r, name); u = units[col[name]] if name in col else ""
    m = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1,
         "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1}
    return v * m.get(u, 1)
agg = collections.OrderedDict()
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "inst": 0.0, "issue": [], "occ": [], "l2": [], "dram": [], "regs": 0})
    a["n"] += 1
    a["t"] += scaled(r, "gpu__time_duration.sum")
    a["rd"] += scaled(r, "dram__bytes_read.sum"); a["wr"] += scaled(r, "dram__bytes_write.sum")
    a["inst"] += f(r, "smsp__inst_executed.sum")
    a["issue"].append(f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", f(r, "sm__inst_executed.avg.pct_of_peak_sustained_active")))
    a["occ"].append(f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"))
    a["l2"].append(f(r, "lts__t_sector_hit_rate.pct"))
    a["dram"].append(f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", f(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed")))
    a["regs"] = int(f(r, "launch__registers_per_thread"))
print("%-34s %4s %9s %9s %8s %7s %6s %6s %6s %5s" % ("kernel", "n", "ms", "dram MB", "GB/s", "dram%", "issue%", "occ%", "L2hit", "regs"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    mb = (a["rd"] + a["wr"]) / 1e6
    avg = lambda v: sum(v) / max(1, len(v))
    print("%-34s %4d %9.3f %9.1f %8.0f %7.1f %6.1f %6.1f %6.1f %5d" % (k[:34], a["n"], a["t"] * 1e3, mb, mb / 1e3 / a["t"] if a["t"] else 0,
                                                                 avg(a["dram"]), avg(a["issue"]), avg(a["occ"]), avg(a["l2"]), a["regs"]))
