"""One launch (after a warm-up) of every HBM-bound kernel of the path at its BASELINE.json size, for `ncu --set full`:
   ncu --set full --clock-control none -k regex:'gather_pool|gather_split|cached_scores|index_rows|additive_pool|group_metrics|resample|shard_' \
       -o gpurun_out/hbm python scratch/prof_hbm.py
Each section writes >126 MB between warm-up and the profiled launch so that the L2 starts cold."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from legommenders_b200 import ops, _lib
from legommenders_b200._lib import call, ptr
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
g = torch.Generator(device='cpu').manual_seed(1)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def cold(): flush.fill_(1); torch.cuda.synchronize()

# (1) gather + masked mean pool over a 4M x 300 table (config 4 size), uniform ids: HBM-resident rows
E = 300
big = torch.empty((4_000_000, E), dtype=torch.float32, device=dev).normal_(0, 0.4)
ids = torch.randint(0, big.shape[0], (260_952, 30), generator=g).to(dev); ids[:, 24:] = -1
ops.gather_pool(ids, None, big); cold(); ops.gather_pool(ids, None, big)
# (1b) gather straight into split-bf16 planes (embedding stage of a training step: 42k token rows, Zipf ids over 400k rows)
from legommenders_b200.synth import zipf_probs
tab = torch.empty((400_000, E), dtype=torch.float32, device=dev).normal_(0, 0.4)
tok = torch.multinomial(torch.from_numpy(zipf_probs(400_000)).float(), 42_587, replacement=True, generator=g).to(dev)
hi = torch.empty((42_587, 304), dtype=torch.bfloat16, device=dev); lo = torch.empty_like(hi)
for _ in range(2):
    cold(); call('lk_gather_split_bf16', ptr(tok), ptr(tab), tab.shape[0], ptr(hi), ptr(lo), tok.numel(), E, 304)
del big, ids
# (5) cached-eval scoring: MIND-small validation shape, rows sorted by user (as stored) and shuffled
D, NU, NI, R = 256, 91_935, 65_238, 2_631_436
U = torch.randn(NU, D, generator=g).to(dev); I = torch.randn(NI, D, generator=g).to(dev)
uid = torch.sort(torch.randint(0, NU, (R,), generator=g)).values.to(dev); iid = torch.randint(0, NI, (R,), generator=g).to(dev)
out = torch.empty(R, dtype=torch.float32, device=dev)
ops.cached_scores(U, I, uid, iid, out=out); cold(); ops.cached_scores(U, I, uid, iid, out=out)
perm = torch.randperm(R, generator=g).to(dev)
cold(); ops.cached_scores(U, I, uid[perm], iid, out=out)
cold(); ops.index_rows(I, iid[:500_000])
# group metrics
from legommenders_b200 import Env
from legommenders_b200.metrics import MetricPool
Env.use_cuda(0)
lab = (torch.rand(R, generator=g) < 0.1).long().to(dev)
pool = MetricPool.parse(['GAUC', 'MRR', 'NDCG@1', 'NDCG@5', 'NDCG@10'])
pool.calculate(out, lab, uid); cold(); pool.calculate(out, lab, uid)
# (3) additive pooling over the item rows of a step (planes + row-dot partials in, as the fused chain leaves them)
T, N = 42_587, 3_520
cu = torch.zeros(N + 1, dtype=torch.int32); cu[1:] = torch.cumsum(torch.full((N,), T // N, dtype=torch.int32), 0); cu[-1] = T
cu = cu.to(dev)
lin = torch.randn(T, D, generator=g).to(dev); P = ops.split_planes(lin)
hid = torch.tanh(torch.randn(T, D, generator=g)).to(dev); w2 = torch.randn(D, generator=g).to(dev)
sp = torch.randn(T, 4, generator=g).to(dev) * 0.1
rep = torch.empty(N, D, device=dev); alpha = torch.empty(T, device=dev)
drep = torch.randn(N, D, generator=g).to(dev); dlin = torch.empty(T, D, device=dev); dpre = torch.empty(T, D, device=dev); dw2p = torch.empty(N, D, device=dev)
for _ in range(2):
    cold(); call('lk_additive_pool_fwd_planes', ptr(P.hi), ptr(P.lo), P.ld, ptr(sp), ptr(cu), ptr(rep), ptr(alpha), N, 40, D)
    cold(); call('lk_additive_pool_bwd_planes', ptr(P.hi), ptr(P.lo), P.ld, ptr(hid), ptr(w2), ptr(alpha), ptr(cu), ptr(drep), ptr(dlin), ptr(dpre), None, None, 0, None, ptr(dw2p), N, 40, D, D)
torch.cuda.synchronize()
print('done')
