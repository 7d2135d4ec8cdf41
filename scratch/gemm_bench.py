import sys, torch, time
sys.path.insert(0, '/root/repo')
from legommenders_b200 import ops, _lib
lib = _lib.load()
def bench(M, N, K, ws, mn=False, act=0, reps=20):
    lib.lk_tc_set_weight_stationary(ws)
    if mn:
        a = ops.split_planes(torch.randn(K, M, device='cuda')); b = ops.split_planes(torch.randn(K, N, device='cuda'))
    else:
        a = ops.split_planes(torch.randn(M, K, device='cuda')); b = ops.split_planes(torch.randn(N, K, device='cuda'))
    out = torch.empty(M, N, device='cuda')
    bias = torch.randn(N, device='cuda')
    for _ in range(3): ops.tc_gemm(a, b, mn, M, N, K, out=out, bias=None if mn else bias, act=act)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.tc_gemm(a, b, mn, M, N, K, out=out, bias=None if mn else bias, act=act)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / reps
    print(f'M={M} N={N} K={K} mn={mn} ws={ws} act={act}: {us:8.1f} us  {2*M*N*K/us/1e6:7.1f} TFLOP/s alg ({6*M*N*K/us/1e6:7.1f} bf16)', flush=True)
for M in (40000, 116160):
    for (N, K) in ((768, 256), (256, 256), (256, 304)):
        for ws in (0, 1):
            bench(M, N, K, ws)
    bench(M, 256, 256, 1, act=1)
    bench(M, 256, 768, 0)
    bench(256, 256, M, 0, mn=True)
    bench(768, 256, M, 0, mn=True)
# split kernel
x = torch.randn(40000, 768, device='cuda')
for _ in range(3): ops.split_planes(x, colsum=True)
torch.cuda.synchronize(); t=time.time()
for _ in range(20): ops.split_planes(x, colsum=True)
torch.cuda.synchronize(); dt=(time.time()-t)/20
print(f'split 40000x768 + colsum: {dt*1e6:.1f} us  {40000*768*8/dt/1e9:.0f} GB/s')
