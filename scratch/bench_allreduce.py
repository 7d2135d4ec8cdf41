"""The gradient all-reduce alone (no skew from the step): lk_allreduce_p2p fused / bracketed vs ncclAllReduce on the 3.49 MB bucket.
torchrun --nproc-per-node N scratch/bench_allreduce.py   (LK_AR_THREADS sweeps the block size of the fused kernel)"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legommenders_b200.trainer import FlatAdam          # noqa: E402

rank, W = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=dev)
n = int(os.environ.get('AR_FLOATS', 872448))


def timeit(fn, iters=200):
    for _ in range(20):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t), 2)


out = {}
for mode in ('fused', 'fused_multimem', 'bracketed', 'nccl'):
    os.environ['LK_P2P_ALLREDUCE'] = '0' if mode == 'nccl' else '1'
    os.environ['LK_P2P_FUSED_BARRIER'] = '1' if mode.startswith('fused') else '0'
    os.environ['LK_P2P_MULTIMEM'] = '1' if mode == 'fused_multimem' else '0'
    m = torch.nn.Linear(n, 1, bias=False).to(dev)
    opt = FlatAdam(m)
    opt.grad.fill_(1.0)
    if mode == 'fused_multimem':
        out['multicast'] = bool(opt._mc_ptr)
        opt.grad.fill_(float(rank + 1)); opt.allreduce(); torch.cuda.synchronize()
        out['multimem_sum_ok'] = bool((opt.grad == W * (W + 1) / 2).all())
    out[mode + '_us'] = timeit(opt.allreduce)
    if mode.startswith('fused') and os.environ.get('LK_AR_TRACE'):
        from legommenders_b200 import _lib
        tr = torch.zeros(148 * 6, dtype=torch.int64, device=dev)
        _lib.load().lk_allreduce_set_trace(tr.data_ptr())
        for _ in range(5):
            opt.allreduce()
        torch.cuda.synchronize()
        _lib.load().lk_allreduce_set_trace(None)
        t = tr.view(148, 6).cpu()
        t = t[t[:, 5] > 0]
        out[mode + '_clk_median'] = dict(zip(['dep', 'rdv_in', 'reduce', 'fence', 'rdv_out', 'total'], t.median(0).values.tolist()))
        out[mode + '_clk_max'] = dict(zip(['dep', 'rdv_in', 'reduce', 'fence', 'rdv_out', 'total'], t.max(0).values.tolist()))
    del opt, m
if rank == 0:
    print(json.dumps(dict(W=W, floats=n, threads=os.environ.get('LK_AR_THREADS', 'auto'), **out)))
dist.destroy_process_group()
