"""Timeline of one lk_tc_chain launch (CTA 0): LK_CHAIN_TRACE=1 python scratch/chain_trace.py M"""
import ctypes, os, sys
os.environ['LK_CHAIN_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legommenders_b200 import ops, _lib
M = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator().manual_seed(0)
a = torch.randn(M, 256, generator=g).cuda()
ws = [(torch.randn(256, 256, generator=g) / 16).cuda() for _ in range(3)]
bs = [torch.randn(256, generator=g).cuda() for _ in range(3)]
dot = torch.randn(256, generator=g).cuda()
A = ops.split_planes(a); W = [ops.split_planes(w) for w in ws]
for _ in range(3):
    ops.tc_chain(A, [dict(w=W[0], bias=bs[0], want_planes=True), dict(w=W[1], bias=bs[1], want_f32=True, want_planes=True),
                     dict(w=W[2], bias=bs[2], act=1, want_f32=True, dotvec=dot)])
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 1024)()
n = _lib.load().lk_tc_chain_trace(ctypes.cast(buf, ctypes.c_void_p), 1024)
t = list(buf)
t0 = t[3 * 256]
ev = []
for i in range(256):
    if t[i]: ev.append((t[i] - t0, 'mma  gemm %d kb %d %s' % (i // 12, i // 3 % 4, ['A ready', 'B ready', 'issued'][i % 3])))
    if t[256 + i]:
        n_, k = divmod(i, 12)
        name = {0: 'prefetched', 1: 'tfull', 10: 'done'}.get(k) or ('chunk %d %s' % ((k - 2) // 2, ['tmem loaded', 'a_epi arrive'][(k - 2) % 2]))
        ev.append((t[256 + i] - t0, 'epi  gemm %d %s' % (n_, name)))
    if t[512 + i] and i % 4 == 0: ev.append((t[512 + i] - t0, 'prod stage (gemm %d kb %d)' % (i // 16, i // 4 % 4)))
for c, name in sorted(ev): print('%8d clk  %7.2f us  %s' % (c, c / 1.965e3, name))
