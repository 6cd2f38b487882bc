"""Top stall sites of one kernel launch in an .ncu-rep (source page).  Usage: ncu_stalls.py rep [topn] [section]
The source page of a multi-launch report is a sequence of sections (per launch: SASS view, then the
source-correlated view when -lineinfo was used); `section` picks one (default 0)."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 22
section = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
print('sections:', len(starts) - 1)
lo, hi = starts[section], starts[section + 1]
hdr = rows[lo + 1]; data = [r for r in rows[lo + 2:hi] if len(r) >= len(hdr) - 1]
si = hdr.index('# Samples'); src = hdr.index('Source')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[si]) for r in data)
agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stalls}
print('total samples', tot, sorted(agg.items(), key=lambda x: -x[1])[:7])
for r in sorted(data, key=lambda r: -int(r[si]))[:topn]:
    reasons = sorted([(int(r[i]), hdr[i]) for i in stalls], reverse=True)[:2]
    print(r[si].rjust(7), r[src][:72].ljust(72), reasons)
