"""Opcode mix and stall summary of one launch from an .ncu-rep: python scratch/ncu_opmix.py rep kernel_regex launch_skip"""
import csv, io, subprocess, sys, collections
rep, rx, skip = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx, '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, isamp, iexec = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) == len(hdr) and r[isamp].isdigit()]
data = data[:len(data)//2]
mix = collections.Counter(); samp = collections.Counter()
for r in data:
    toks = r[isrc].split()
    op = toks[0] if not toks[0].startswith('@') else toks[1]
    op = op.split('.')[0]
    mix[op] += int(r[iexec]); samp[op] += int(r[isamp])
tot = sum(mix.values()); ts = sum(samp.values())
print('total warp instr', tot, 'samples', ts)
for op, c in mix.most_common(22): print(f'{op:10s} {c:10d} {100*c/tot:5.1f}%   samples {100*samp[op]/ts:5.1f}%')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.Counter()
for r in data:
    for i in stalls: agg[hdr[i]] += int(r[i] or 0)
print({k: v for k, v in agg.most_common(8)})
