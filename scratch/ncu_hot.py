"""Top sampled SASS instructions of one launch: python scratch/ncu_hot.py rep kernel_regex launch_skip [n]"""
import csv, io, subprocess, sys
rep, rx, skip = sys.argv[1], sys.argv[2], sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx, '--launch-skip', skip, '--launch-count', '1'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:150])
hdr = rows[1]
isrc, isamp, iexec = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[isamp].isdigit()]
tot = sum(int(r[isamp]) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {}
for r in data:
    for i in stalls:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:n]
for i in sorted(order):
    r = data[i]
    top = sorted(((int(r[j] or 0), hdr[j]) for j in stalls), reverse=True)[:2]
    print(f'{i:5d} {int(r[isamp]):6d} {100*int(r[isamp])/max(tot,1):5.1f}% x{r[iexec]:>8} {r[isrc].strip()[:70]:70s} {top}')
