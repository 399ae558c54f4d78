"""Per-instruction stall summary of one kernel in an ncu report (needs -lineinfo builds and --import-source on).
usage: python scripts/ncu_stalls.py <report.ncu-rep> <kernel regex> [top N]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '-k', f'regex:{pat}'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'lts__t_sectors_op_write.sum',
        'lts__t_sectors_op_read.sum', 'smsp__cycles_active.avg']
for r in rows[2:3]:
    for m in want:
        if m in hdr:
            print(f'{m:75s} {r[hdr.index(m)]} {rows[1][hdr.index(m)]}')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', f'regex:{pat}'], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr = r[1]
rows = r[2:]
iS, isrc, iex = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(x[iS]) for x in rows)
print('total samples', tot)
agg = {}
for x in rows:
    for i in stalls:
        agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(x[i])
print({k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for n, x in enumerate(rows):
    x.append(n)
for x in sorted(rows, key=lambda x: -int(x[iS]))[:top]:
    st = {hdr[i][6:]: int(x[i]) for i in stalls if int(x[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{x[-1]:5d} {int(x[iS]):6d} {100*int(x[iS])/tot:5.1f}% ex={x[iex]:>9} {x[isrc].strip()[:64]:64s} {st}")
