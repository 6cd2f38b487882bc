"""Summarise ncu outputs into profiles/ (text): launch list -> per-kernel totals/shares; .ncu-rep -> key metrics."""
import csv, collections, subprocess, sys

def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[1:]:
        try:
            t = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        k = r[ki].split('(')[0][-60:]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += t; tot += t
    with open(out, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold cache, serialised: compare SHARES)\n')
        f.write('# source: %s ; total %.1f us over %d launches\n' % (path, tot / 1e3, sum(a[0] for a in agg.values())))
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write('%-62s n=%5d total=%11.1f us avg=%9.1f us share=%5.1f%%\n' % (k, n, t / 1e3, t / 1e3 / n, 100 * t / tot))

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput', 'dram__cycles_active',
        'sm__pipe_tensor_cycles_active', 'sm__throughput.avg.pct', 'sm__warps_active.avg.pct', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed_pipe_xu', 'sm__inst_executed_pipe_fma', 'sm__inst_executed_pipe_fp64',
        'l1tex__data_bank_conflicts', 'lts__t_bytes.sum', 'sm__pipe_fma_cycles_active', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'smsp__warp_issue_stalled', 'launch__shared_mem_per_block', 'launch__occupancy_limit']

def report(rep, out):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    with open(out, 'w') as f:
        f.write('# ncu --set full --clock-control none ; source: %s\n' % rep)
        for r in rows[2:]:
            f.write('## kernel: %s\n' % r[hdr.index('Kernel Name')][:100])
            for i, h in enumerate(hdr):
                if any(k in h for k in KEYS):
                    f.write('%-70s %s %s\n' % (h, r[i], rows[1][i]))

if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        report(sys.argv[2], sys.argv[3])
