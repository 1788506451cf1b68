"""Summarise an .ncu-rep (raw page) into the handful of numbers we steer by."""
import csv, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sass__inst_executed_register_spilling', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_elapsed.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'lts__t_bytes.sum']
stall = [h for h in hdr if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')]
out = []
for r in data:
    d = {}
    for k in keys:
        if k in hdr:
            d[k + ' [' + units[hdr.index(k)] + ']'] = r[hdr.index(k)]
    st = sorted(((float(r[hdr.index(h)]), h.split('issue_stalled_')[1].split('_per_issue')[0]) for h in stall), reverse=True)[:7]
    d['stalls(warps per issue)'] = {n: round(v, 2) for v, n in st}
    out.append(d)
print(json.dumps(out, indent=1))
