"""Mint golden vectors from the LIVE reference (run in the build container only; needs /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

For every case the reference's own Legommender / Resampler / ReprCacher / MetricPool run on CPU fp32 (eval-mode
equivalent: all dropouts 0) over a seeded synthetic world, with parameters from the deterministic numpy recipe
`synth.init_state`, and the inputs/outputs are written to tests/golden/<case>.npz.  The worlds and parameters are
NOT stored: tests rebuild them from the same seeds (see tests/golden/cases.py).
"""
from __future__ import annotations

import copy
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import torch  # noqa: E402

import ref_harness as rh  # noqa: E402
from cases import CASES, make_world, make_state, flatten_tree, sample_strided  # noqa: E402


def run_case(name: str, c: dict):
    torch.manual_seed(0)
    torch.set_num_threads(4)
    world, llm = make_world(c)
    model, resampler, cfg, Env = rh.build_reference(world, c['kind'], hidden=c['hidden'], heads=c['heads'],
                                                    additive=c['additive'], dropout=0.0,
                                                    use_neg_sampling=c.get('use_neg_sampling', True), llm_item_table=llm)
    from loader.data_set import DataSet
    from torch.utils.data import DataLoader
    from utils.metrics import MetricPool

    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    state = make_state(c, world, shapes, llm)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    model.train()   # dropouts are 0; train() keeps MHA on the same (non-fast-path) code the trainer runs

    out = {}
    # ---- one training batch through Resampler -> default_collate -----------------------------------------
    Env.train()
    random.seed(c['seed'])
    if c.get('use_neg_sampling', True):
        ut = world.train_table()
    else:
        ut = world.eval_table()
    loader = DataLoader(DataSet(ut, resampler), batch_size=c['batch'], num_workers=0, shuffle=False)
    batch = next(iter(loader))
    for k, v in flatten_tree(batch).items():
        out['batch/' + k] = v.numpy()

    loss = model(batch=copy.deepcopy(batch))
    loss.backward()
    out['loss'] = np.float32(loss.item())
    for n, p in model.named_parameters():
        if p.requires_grad:
            g = p.grad.detach().numpy() if p.grad is not None else np.zeros(tuple(p.shape), dtype=np.float32)   # unused parameter (CNNCat's linear)
            if c.get('full_grads', True):
                out['grad/' + n] = g
            else:
                out['gradnorm/' + n] = np.float32(np.linalg.norm(g.astype(np.float64)))
                out['gradmax/' + n] = np.float32(np.abs(g).max())
                out['gradsample/' + n] = sample_strided(g)

    # ---- scores + intermediate representations for the same batch ------------------------------------------
    Env.test()
    with torch.no_grad():
        b2 = copy.deepcopy(batch)
        scores = model(batch=b2)
        out['scores'] = scores.numpy()
        b3 = copy.deepcopy(batch)
        if cfg.use_item_content:
            out['items'] = model.get_item_content(b3, 'item_id').numpy()
        out['user'] = model.get_user_content(copy.deepcopy(batch)).numpy()
    Env.train()

    # ---- cached evaluation ---------------------------------------------------------------------------------------
    if c.get('cached_eval', False):
        Env.test()
        model.eval()
        fast = DataSet(world.fast_table(), resampler)
        model.cacher.cache(item_contents=resampler.item_cache, user_contents=fast)
        if cfg.use_item_content:
            out['item_repr'] = model.cacher.item.repr.detach().numpy()
        out['user_repr'] = model.cacher.user.repr.detach().numpy()
        ev = DataLoader(DataSet(world.eval_table(), resampler), batch_size=64, num_workers=0, shuffle=False)
        sc, lb, gr = [], [], []
        with torch.no_grad():
            for eb in ev:
                gr.extend(eb['user_id'].tolist())
                lb.extend(eb['click'].tolist())
                s = model(batch=eb)
                sc.extend(s.reshape(-1).tolist())
        out['eval_scores'] = np.asarray(sc, dtype=np.float32)
        out['eval_labels'] = np.asarray(lb, dtype=np.int64)
        out['eval_groups'] = np.asarray(gr, dtype=np.int64)
        names = ['GAUC', 'MRR', 'NDCG@1', 'NDCG@5', 'NDCG@10']
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            vals = MetricPool.parse(names).calculate(sc, lb, gr)
        out['metrics'] = np.asarray([vals[n] for n in names], dtype=np.float64)
        model.cacher.clean()
        Env.train()

    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: loss={out["loss"]:.6f} -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)')


if __name__ == '__main__':
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        run_case(name, c)
