"""Summarise an .ncu-rep (raw page + source page) into text: key metrics, dynamic opcode mix, stall reasons."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, v = r[0], r[-1]
keys = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'sm__cycles_elapsed.max', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct',
        'l1tex__throughput.avg.pct', 'launch__grid_size', 'launch__block_size', 'sm__cycles_active.avg', 'smsp__cycles_active.avg', 'sm__pipe_alu_cycles_active',
        'sm__pipe_fmaheavy', 'sm__pipe_fmalite', 'sm__inst_executed_pipe_fmaheavy', 'sm__inst_executed_pipe_fmalite', 'sm__inst_executed_pipe_uniform']
print("== kernel:", v[h.index("Kernel Name")] if "Kernel Name" in h else "?")
for i, k in enumerate(h):
    if any(k.startswith(x) for x in keys) and not k.endswith(('.min', '.max.pct', '.sum.pct_of_peak_sustained_active')):
        print("%-75s %s" % (k, v[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {x: i for i, x in enumerate(hdr)}
tot, byop, st = 0, {}, {}
cols = [x for x in hdr if x.startswith('stall_') and 'Not Issued' not in x]
for r_ in data:
    e = r_[ix['Instructions Executed']]
    if not e.isdigit():
        continue
    e = int(e)
    t = r_[ix['Source']].split()
    op = (t[1] if t and t[0].startswith('@') else (t[0] if t else '?')).rstrip(';')
    byop[op] = byop.get(op, 0) + e
    tot += e
    for c in cols:
        if r_[ix[c]].isdigit():
            st[c] = st.get(c, 0) + int(r_[ix[c]])
print("== dynamic warp instructions:", tot)
for k, x in sorted(byop.items(), key=lambda z: -z[1])[:24]:
    print("  %-22s %12d %5.1f%%" % (k, x, 100 * x / tot))
s = sum(st.values()) or 1
print("== stall samples")
for k, x in sorted(st.items(), key=lambda z: -z[1])[:9]:
    print("  %-24s %8d %5.1f%%" % (k, x, 100 * x / s))
