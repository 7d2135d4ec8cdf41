"""Sample distribution by straight-line code region of one launch: python scratch/ncu_regions.py rep kernel_regex launch_skip"""
import csv, io, subprocess, sys, collections
rep, rx, skip = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx, '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, isamp, iexec = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) == len(hdr) and r[isamp].isdigit()]
data = data[:len(data)//2]
tot = sum(int(r[isamp]) for r in data)
# group consecutive instructions by exec count
groups = []
for i, r in enumerate(data):
    e = int(r[iexec])
    if groups and groups[-1][0] == e: groups[-1][1] += int(r[isamp]); groups[-1][2] += 1
    else: groups.append([e, int(r[isamp]), 1, i])
print('total', tot, 'instr', len(data), 'total warp instr', sum(int(r[iexec]) for r in data))
for e, sm, n, i0 in groups:
    if sm > tot * 0.01: print(f'start {i0:5d} n_instr {n:4d} exec {e:8d} samples {sm:6d} {100*sm/tot:5.1f}%  instr-issues {e*n:10d}')
