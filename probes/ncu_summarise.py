# Turns ncu outputs into the small tracked summaries under profiles/.
#   python probes/ncu_summarise.py rep <file.ncu-rep> <out.csv>            selected metrics of every captured launch (one column per launch)
#   python probes/ncu_summarise.py launches <launches.csv> <out.csv> [n] [regex]   per-(kernel, grid) totals of the LAST n launches whose
#                                                                                   name matches regex (default: all launches)
import csv, io, re, subprocess, sys

KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch_%d" % i for i in range(len(data))])
        for name in KEEP:
            if name in hdr:
                i = hdr.index(name)
                w.writerow([name, units[i]] + [r[i] for r in data])


def launches(path, out, last=None, pattern=None):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    data = [r for r in rows[hi + 1:] if len(r) > mv and (pattern is None or re.search(pattern, r[kn]))]
    if last:
        data = data[-last:]
    agg, order = {}, []
    for r in data:
        key = (re.sub(r"\(.*", "", r[kn]).replace("void nla::", "").replace("void ", ""), r[gs])
        if key not in agg:
            agg[key] = [0, 0.0]
            order.append(key)
        agg[key][0] += 1
        agg[key][1] += float(r[mv].replace(",", "")) / 1e6   # ns -> ms
    total = sum(v[1] for v in agg.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "launches", "total_ms", "share_pct"])
        for key in order:
            c, ms = agg[key]
            w.writerow([key[0], key[1], c, "%.3f" % ms, "%.1f" % (100 * ms / total)])
        w.writerow(["total", "", sum(v[0] for v in agg.values()), "%.3f" % total, "100"])


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        rep(sys.argv[2], sys.argv[3])
    else:
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 and int(sys.argv[4]) > 0 else None, sys.argv[5] if len(sys.argv) > 5 else None)
