"""Summarise every kernel of an .ncu-rep (ncu --set full) into the text table committed under profiles/.
usage: python tools/ncu_chain_summary.py report.ncu-rep "header line" > profiles/xxx.ncu.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]

rep = sys.argv[1]
print(sys.argv[2] if len(sys.argv) > 2 else rep)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("-----")
    print("%-70s %s" % ("Kernel Name", r[hdr.index("Kernel Name")]))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%-70s %s %s" % (k, r[i], units[i]))
