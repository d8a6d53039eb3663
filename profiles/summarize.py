"""Turns gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.
usage: python profiles/summarize.py launches <launches.csv> | full <report.ncu-rep>"""
import csv
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [x["Kernel Name"] for x in rows]
    last = [i for i, n in enumerate(names) if "k_prepare" in n or "k_filter_first" in n][-1]
    # one step = clear()/reset kernels are outside; from k_filter_first to the end of the tally
    seg = rows[last:]
    tot = sum(float(x["Metric Value"]) for x in seg)
    print("# one step of bench.py (last step in the capture), ncu gpu__time_duration.sum per launch, --clock-control none")
    print("# (cold-cache, serialised: compare SHARES, not absolutes)")
    for x in seg:
        t = float(x["Metric Value"])
        print("%-70s grid=%-12s %10.1f us  %5.1f%%" % (x["Kernel Name"][:70], x["Grid Size"].replace(" ", ""), t / 1e3, 100 * t / tot))
    print("total %.1f us over %d launches" % (tot / 1e3, len(seg)))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, unit, vals = rows[0], rows[1], rows[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__inst_executed.sum",
            "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "lts__t_bytes.sum",
            "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
    for i, h in enumerate(hdr):
        if h in want:
            print("%-70s %-12s %s" % (h, unit[i], vals[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
