"""Summarise every kernel of an .ncu-rep (ncu --set full) into the text table committed under profiles/ and, with
--json, into the machine-readable summary bench.py quotes in `roofline.traffic` / `ncu_evidence` (so that no ncu number is
ever a literal in bench.py).
usage: python tools/ncu_chain_summary.py report.ncu-rep "header line" [--json profiles/chain_kernels_rNN.json --batch 1024] > profiles/xxx.ncu.txt"""
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

if "--json" in sys.argv:
    import json
    path = sys.argv[sys.argv.index("--json") + 1]
    batch = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0,
             "nsecond": 1e-6, "second": 1e3}

    def val(r, k):
        if k not in hdr:
            return None
        i = hdr.index(k)
        return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)

    tags = (("ofdm_rx", "fft"), ("chest_kernel", "chest"), ("pdsch_llr_dematch", "demap"), ("turbo_decode", "turbo"), ("tb_assemble", "tb"), ("tdec_deinterleave", "deint"))
    out = {"source": "profiles/%s (ncu --set full --clock-control none; regenerate with tools/ncu_chain_summary.py)" %
                     path.split("/")[-1].replace(".json", ".ncu.txt"), "batch": batch}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        for pat, tag in tags:
            if pat in name and tag not in out:
                rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
                out[tag] = {"kernel": name.split("(")[0], "launch_ms_under_ncu": val(r, "gpu__time_duration.sum"),
                            "dram_bytes": (rd or 0.0) + (wr or 0.0),
                            "alu_pipe_pct": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                            "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                            "l1tex_pct": val(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                            "registers_per_thread": val(r, "launch__registers_per_thread"),
                            "block_size": val(r, "launch__block_size"), "grid_size": val(r, "launch__grid_size")}
    json.dump(out, open(path, "w"), indent=1)

