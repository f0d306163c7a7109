"""Turn ncu output into the small text summaries kept under profiles/.

    python scripts/ncu_summary.py launches <launches.csv> [top]      per-kernel totals / shares of a launch list
    python scripts/ncu_summary.py report <file.ncu-rep>              key metrics of every kernel in a --set full report
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
STALLS = 'smsp__pcsamp_warps_issue_stalled_'


def launches(path, top=30):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u in ('ns', 'nsecond') else v * 1e3 if u in ('ms', 'msecond') else v
        a = agg[row['Kernel Name']]
        a[0] += 1
        a[1] += v
        tot += v
    print('# %s: %d launches, %.1f us total (ncu per-launch times: cold cache, serialised)' % (
        path, sum(a[0] for a in agg.values()), tot))
    print('#   total_us  launches  us/launch  share  kernel')
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%10.1f %8d %10.1f %6.1f%%  %s' % (t, n, t / n, 100 * t / tot, k[:110]))


def report(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('## %s' % r[hdr.index('Kernel Name')][:120])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('   %-75s %s %s' % (k, r[i], units[i]))
        st = [(hdr[i][len(STALLS):], float(r[i] or 0)) for i in range(len(hdr)) if hdr[i].startswith(STALLS)
              and not hdr[i].endswith('_not_issued')]
        tot = sum(v for _, v in st) or 1.0
        print('   stall samples: ' + ', '.join('%s %.0f%%' % (k, 100 * v / tot) for k, v in sorted(st, key=lambda kv: -kv[1])[:6]))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
    else:
        report(sys.argv[2])
