"""Time lk_tc_chain against the three separate lk_tc_gemm launches it replaces (CUDA events, L2 flushed between reps by size)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legommenders_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 42587
g = torch.Generator().manual_seed(0)
a = torch.randn(M, 256, generator=g).cuda()
ws = [(torch.randn(256, 256, generator=g) / 16).cuda() for _ in range(3)]
bs = [torch.randn(256, generator=g).cuda() for _ in range(3)]
dot = torch.randn(256, generator=g).cuda()
A = ops.split_planes(a); W = [ops.split_planes(w) for w in ws]
def chain():
    return ops.tc_chain(A, [dict(w=W[0], bias=bs[0], want_planes=True), dict(w=W[1], bias=bs[1], want_f32=True, want_planes=True),
                            dict(w=W[2], bias=bs[2], act=1, want_f32=True, dotvec=dot)])
def chain_b():
    return ops.tc_chain(A, [dict(w=W[0], addsrc=a, want_planes=True, want_colsum=True), dict(w=W[1], want_planes=True, want_colsum=True),
                            dict(w=W[2], want_f32=True)], b_mn=True)
def sep():
    _, p0, _ = ops.tc_gemm_ex(A, W[0], M, 256, 256, store_c=False, bias=bs[0], want_planes=True)
    y1, p1, _ = ops.tc_gemm_ex(p0, W[1], M, 256, 256, bias=bs[1], want_planes=True)
    y2, _, _ = ops.tc_gemm_ex(p1, W[2], M, 256, 256, bias=bs[2], act=1)
    return y2
for name, fn in (('chain fwd', chain), ('chain bwd', chain_b), ('3 x tc_gemm', sep)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'{name}: M={M} {ms * 1e3:.1f} us  {3 * 2 * M * 256 * 256 / ms / 1e9:.1f} TFLOP/s algorithmic')
