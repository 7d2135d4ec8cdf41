"""Pin the oracle: its outputs must reproduce the committed vectors minted from the LIVE reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import cases
import helpers
from oracle import lego_oracle as O

TOL = 2e-5   # CPU fp32 vs CPU fp32 (different association inside nn.MultiheadAttention's fused path)
TOL_G = 1e-4  # gradients: two fp32 CPU implementations differ by up to 7e-5 of a tensor's max through the LayerNorm stacks of the BERT-style encoders


@pytest.mark.parametrize('name', list(cases.CASES))
def test_forward_backward_matches_reference(name):
    c = cases.CASES[name]
    g = cases.load(name)
    world, llm = cases.make_world(c)
    batch = cases.unflatten_batch(g)
    r = helpers.oracle_run(c, world, llm, batch)
    assert abs(r['loss'] - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    assert helpers.normwise(r['scores'], g['scores']) < TOL
    assert helpers.normwise(r['user'], g['user']) < TOL
    if 'items' in g.files:
        assert helpers.normwise(r['items'], g['items']) < TOL
    scale = max(np.abs(v).max() for v in r['grads'].values())
    for k, v in r['grads'].items():
        if 'grad/' + k in g.files:
            ref = g['grad/' + k]
            assert np.abs(v - ref).max() <= helpers.grad_bound(c, np.abs(ref).max(), scale, TOL_G), k
        else:
            assert abs(np.linalg.norm(v.astype(np.float64)) - float(g['gradnorm/' + k])) <= 1e-4 * max(float(g['gradnorm/' + k]), 5e-2 * scale), k
            ref = g['gradsample/' + k]
            assert np.abs(cases.sample_strided(v) - ref).max() <= helpers.grad_bound(c, float(g['gradmax/' + k]), scale, TOL_G), k


@pytest.mark.parametrize('name', [n for n, c in cases.CASES.items() if c.get('cached_eval')])
def test_cached_eval_matches_reference(name):
    c = cases.CASES[name]
    g = cases.load(name)
    world, llm = cases.make_world(c)
    _, state = helpers.oracle_state(c, world, llm)
    spec = helpers.case_spec(c, world)
    item_repr, user_repr, scores = helpers.oracle_cached_eval(c, world, llm, state, spec)
    if item_repr is not None:
        assert np.abs(item_repr.numpy() - g['item_repr']).max() < 1e-5
    # users with an empty history never occur (history length >= 1), so every row is comparable
    assert np.abs(user_repr.numpy() - g['user_repr']).max() < 1e-5
    assert helpers.normwise(scores.numpy(), g['eval_scores']) < TOL
    assert np.array_equal(world.eval_click, g['eval_labels'])
    assert np.array_equal(world.eval_users, g['eval_groups'])
    m = O.metric_pool(g['eval_scores'], g['eval_labels'], g['eval_groups'])
    for v, ref in zip(m.values(), g['metrics']):
        assert round(v, 4) == round(float(ref), 4)


def test_layouts_match_reference_batch():
    """a1/a2/a17: the oracle's integer restatements reproduce the reference batch bit for bit."""
    for name in ('nrms_small', 'naml_small'):
        c = cases.CASES[name]
        g = cases.load(name)
        world, _ = cases.make_world(c)
        trees = helpers.item_trees(world, c['kind'])
        users = g['batch/user_id']
        cand_title = g['batch/item_id/input_ids/' + world.title_col]
        hist_title = g['batch/history/input_ids/' + world.title_col]
        for b, u in enumerate(users):
            ids, mask = O.pad_history(world.histories[int(u)], world.hist_len)
            assert np.array_equal(mask, g['batch/__clicks_mask__'][b])
            for t, it in enumerate(ids):
                assert np.array_equal(trees[int(it)]['input_ids'][world.title_col], hist_title[b, t])
            # candidate 0 is the positive of the b-th training row
            assert np.array_equal(trees[int(world.train_pos[b])]['input_ids'][world.title_col], cand_title[b, 0])
        if c['kind'] == 'nrms':
            am = g['batch/history/attention_mask']
            sp = g['batch/history/input_ids/' + O.SPECIAL_VOCAB]
            for b, u in enumerate(users):
                ids, _ = O.pad_history(world.histories[int(u)], world.hist_len)
                for t, it in enumerate(ids):
                    assert np.array_equal(trees[int(it)]['attention_mask'], am[b, t])
                    assert np.array_equal(trees[int(it)]['input_ids'][O.SPECIAL_VOCAB], sp[b, t])


def test_metrics_match_sklearn():
    from sklearn.metrics import ndcg_score, roc_auc_score
    rng = np.random.default_rng(0)
    for trial in range(50):
        n = int(rng.integers(2, 40))
        s = np.round(rng.standard_normal(n), 1 if trial % 2 else 6)   # odd trials force score ties
        y = (rng.random(n) < 0.3).astype(np.int64)
        y[0], y[1] = 1, 0
        assert abs(O.auc(s, y) - roc_auc_score(y, s)) < 1e-12
        for k in (1, 5, 10):
            assert abs(O.ndcg(s, y, k) - ndcg_score([y], [s], k=k)) < 1e-12
        order = sorted(range(n), key=lambda i: s[i], reverse=True)
        yt = [y[i] for i in order]
        ref = sum(yt[i] / (i + 1) for i in range(n)) / sum(yt)
        assert abs(O.mrr(s, y) - ref) < 1e-12
