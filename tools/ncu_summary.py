"""Condense an .ncu-rep (ncu --set full) into one CSV row per profiled launch with the metrics the roofline notes quote.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.csv"""
import csv, io, subprocess, sys
KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = [(k, hdr.index(k)) for k in KEEP if k in hdr]
w = csv.writer(sys.stdout)
w.writerow([f"{k} [{units[i]}]" if units[i] else k for k, i in cols])
for r in rows[2:]:
    w.writerow([r[i] for _, i in cols])
