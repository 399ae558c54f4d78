"""Turn an ncu report into the markdown summary kept under profiles/ (run where ncu is installed).

usage: python scripts/ncu_summary.py <report.ncu-rep> <out.md> "<title>" "<command line that produced it>"
"""
import csv
import io
import subprocess
import sys

METRICS = ['Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
           'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
           'launch__shared_mem_per_block_dynamic', 'launch__cluster_dim_x',
           'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']


def main():
    rep, out, title, cmd = sys.argv[1:5]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    name_i = hdr.index('Kernel Name')
    seen = set()
    with open(out, 'w') as f:
        f.write(f'# {title}\n\nCommand: `{cmd}`\n\nPer-launch values (first captured launch of each kernel).\n')
        for r in body:
            k = r[name_i]
            if k in seen:
                continue
            seen.add(k)
            f.write(f'\n## {k.split("(")[0]}\n\n| metric | value | unit |\n|---|---|---|\n')
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f'| {m} | {r[i]} | {units[i]} |\n')


if __name__ == '__main__':
    main()
