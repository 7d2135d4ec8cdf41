import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legommenders_b200 import ops
U, N, k = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 10
g = torch.Generator(device='cuda').manual_seed(0)
Up = ops.split_planes(torch.empty((U, 256), device='cuda').normal_(generator=g))
Ip = ops.split_planes(torch.empty((N, 256), device='cuda').normal_(generator=g))
for _ in range(2): ops.sweep_topk(Up, Ip, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): ops.sweep_topk(Up, Ip, k)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f'sweep_topk U={U} N={N}: {ms:.2f} ms, {2.0*U*N*256/ms/1e9:.1f} TFLOP/s algorithmic, {U*N/ms/1e6:.1f} G scores/s')
