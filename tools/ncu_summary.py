"""Key metrics of an `ncu --set full` report as text:  python tools/ncu_summary.py rep.ncu-rep [title]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, v = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum']
if len(sys.argv) > 2:
    print(sys.argv[2])
print('report:', rep.split('/')[-1])
for w in want:
    for i, n in enumerate(h):
        if n == w:
            print('  %-78s %s %s' % (n, v[i][:110], units[i]))
