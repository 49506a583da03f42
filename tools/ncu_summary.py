"""Summarise ncu outputs into profiles/ (run in the build container):
   python tools/ncu_summary.py launches gpurun_out/launches_c5.csv  > profiles/<name>.txt
   python tools/ncu_summary.py raw gpurun_out/prof_main_c5.ncu-rep  > profiles/<name>.txt
"""
import collections, csv, subprocess, sys

def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(d["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6}.get(d["Metric Unit"], 1.0)
            k = d["Kernel Name"].split("(")[0][:70]
            a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
            a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)")
    print("%-72s %6s %12s %7s %10s %10s %10s" % ("kernel", "n", "total_us", "share", "avg_us", "min_us", "max_us"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %6d %12.1f %6.1f%% %10.1f %10.1f %10.1f" % (k, a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[2], a[3]))
    print("%-72s %6s %12.1f" % ("TOTAL", "", tot))

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]

def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("## kernel:", d.get("Kernel Name", "?")[:100], " grid", d.get("Grid Size"), " block", d.get("Block Size"))
        for h, u, v in zip(hdr, units, vals):
            if h in WANT or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                print("%-95s %-16s %s" % (h, u, v))

if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
