"""Key metrics of every kernel in an ncu report, one line each (for profiles/*.txt).
usage: python tools/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
ci = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print(r[ci["Kernel Name"]])
    for w in WANT:
        if w in ci:
            print("    %-70s %s %s" % (w, r[ci[w]], units[ci[w]]))
