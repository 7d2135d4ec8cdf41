"""Summarise an .ncu-rep (raw page) into a compact table: python scratch/ncu_summary.py file.ncu-rep [regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [('gpu__time_duration.sum', 'us'), ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'blk'),
        ('launch__occupancy_limit_registers', 'occR'), ('launch__occupancy_limit_shared_mem', 'occS'),
        ('sm__warps_active.avg.per_cycle_active', 'warps'), ('sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'issue%'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('dram__bytes_read.sum', 'rdMB'), ('dram__bytes_write.sum', 'wrMB'),
        ('lts__t_sector_hit_rate.pct', 'L2hit%'), ('smsp__inst_executed.sum', 'inst')]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
ik = hdr.index('Kernel Name')
print('kernel'.ljust(34) + ' '.join(n.rjust(9) for _, n in idx))
for r in data:
    if pat and not pat.search(r[ik]):
        continue
    name = re.sub(r'^void |lk::|tc::|\(.*$', '', r[ik])[:33]
    vals = []
    for i, n in idx:
        v = r[i]
        try:
            f = float(v.replace(',', ''))
            v = f'{f:.0f}' if f >= 1000 else f'{f:.2f}'.rstrip('0').rstrip('.')
        except ValueError:
            pass
        vals.append(v[:9].rjust(9))
    print(name.ljust(34) + ' '.join(vals))
