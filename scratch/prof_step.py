"""Light driver for ncu captures: the bench's batch shape (64 impressions, title 30, history 50, hidden 256) on a SMALL world
(3k items, 30k words) so that start-up and ncu's per-replay memory save/restore stay short.  Not a bench."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as ge

ge.build()
from legommenders_b200 import Env, builder
from legommenders_b200.batching import BatchBuilder, tree_to_device
from legommenders_b200.synth import MindWorld
from legommenders_b200.trainer import FlatAdam, NativeNRMSStep

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
autograd = len(sys.argv) > 2 and sys.argv[2] == 'autograd'
torch.manual_seed(1)
world = MindWorld(n_items=3000, n_words=30000, n_users=2000, n_train=4096, n_eval_groups=100, seed=2023)
model, resampler, cfg = builder.build_model(world, 'nrms', hidden=256, heads=8, additive=256, dropout=0.1, neg_count=4)
opt = FlatAdam(model, lr=1e-3)
Env.train()
model.train()
bb = BatchBuilder(resampler, world, neg_count=4, seed=1)
rng = np.random.default_rng(7)
batches = [tree_to_device(bb.train_batch(rng.integers(0, world.n_train, size=64)), Env.device, non_blocking=False) for _ in range(2)]
native = None if autograd else NativeNRMSStep(model, opt)
for i in range(steps):
    if native is not None:
        loss = native.step(batches[i % 2])
    else:
        opt.zero_grad()
        loss = model(batch=batches[i % 2])
        loss.backward()
        opt.step()
torch.cuda.synchronize()
print('loss', float(loss))
