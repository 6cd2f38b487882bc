import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 22
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]; data = rows[2:]
si = hdr.index('# Samples'); src = hdr.index('Source')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[si]) for r in data)
agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stalls}
print('total samples', tot, sorted(agg.items(), key=lambda x: -x[1])[:7])
for r in sorted(data, key=lambda r: -int(r[si]))[:topn]:
    reasons = sorted([(int(r[i]), hdr[i]) for i in stalls], reverse=True)[:2]
    print(r[si].rjust(7), r[src][:64].ljust(64), reasons)
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
for k in ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers', 'launch__registers_per_thread',
          'smsp__inst_executed_pipe_xu.sum', 'sm__pipe_tensor_subpipe', 'dram__bytes_read.sum', 'lts__t_sectors_op_read.sum', 'sm__pipe_tc']:
    for i, name in enumerate(h):
        if k in name:
            print(name, rr[2][i], rr[1][i])
