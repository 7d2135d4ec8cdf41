"""a17 (Resampler semantics) on the host: the vectorised BatchBuilder against the REFERENCE's own batch for the same candidates, and the
candidate draw of the device Resampler (`lk_resample_reference`, the inline function the kernel executes) against the semantics of
loader/resampler.py:139-193 restated in the oracle."""
import ctypes
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import cases
from oracle import lego_oracle as O


def item_layouts(world, kind):
    """Per-item inputer outputs (Resampler.item_cache) from this package's host inputers; no device needed."""
    from legommenders_b200.inputer.concat_inputer import ConcatInputer
    from legommenders_b200.inputer.simple_inputer import SimpleInputer
    ut = world.item_table()
    inputs = [world.title_col, 'category']
    inp = ConcatInputer(use_cls_token=False, use_sep_token=True, ut=ut, inputs=inputs, eh=None) if kind == 'nrms' else \
        SimpleInputer(ut=ut, inputs=inputs, eh=None)
    return [inp(ut[i]) for i in range(len(ut))]


@pytest.mark.parametrize('name', ['nrms_small', 'naml_small', 'nrms_full'])
def test_batch_builder_equals_reference_batch(name):
    """BatchBuilder.train_batch (the vectorised host builder) reproduces the reference Resampler + default_collate batch BIT FOR BIT when it is
    given the candidates the reference drew (recovered from the golden batch by matching token rows)."""
    from legommenders_b200.batching import BatchBuilder
    c = cases.CASES[name]
    g = cases.load(name)
    world, _ = cases.make_world(c)
    layouts = item_layouts(world, c['kind'])
    title = 'batch/item_id/input_ids/' + world.title_col
    cat = 'batch/item_id/input_ids/category'
    B, C = g[title].shape[:2]
    key = {}
    for i, lay in enumerate(layouts):
        key.setdefault((tuple(lay['input_ids'][world.title_col].tolist()), tuple(lay['input_ids']['category'].tolist())), i)
    cand = np.array([[key[(tuple(g[title][b, j].tolist()), tuple(g[cat][b, j].tolist()))] for j in range(C)] for b in range(B)])
    rows = g['batch/index']
    assert np.array_equal(cand[:, 0], [key[(tuple(layouts[p]['input_ids'][world.title_col].tolist()),
                          tuple(layouts[p]['input_ids']['category'].tolist()))] for p in world.train_pos[rows]])      # candidate 0 = the positive
    bb = BatchBuilder(SimpleNamespace(item_cache=layouts), world, neg_count=C - 1, pin=False)
    batch = bb.train_batch(rows, cand=cand)
    flat = cases.flatten_tree(batch)
    assert sorted('batch/' + k for k in flat) == sorted(k for k in g.files if k.startswith('batch/'))
    for k, v in flat.items():
        assert v.dtype == torch.int64 and np.array_equal(v.numpy(), g['batch/' + k]), k


def draw(lib, seed, row, pos, negs, K, n_items):
    negs = np.ascontiguousarray(negs, dtype=np.int64)
    out = (ctypes.c_int64 * (K + 1))()
    rc = lib.lk_resample_reference(seed, row, pos, negs.ctypes.data, len(negs), K, n_items, ctypes.cast(out, ctypes.c_void_p))
    assert rc == 0, lib.lk_last_error()
    return list(out)


def test_candidate_draw_semantics():
    """[pos, negs...]; negatives = min(K, len) DISTINCT POSITIONS of the true-negative list + uniform ids for the rest (resampler.py:160-171,
    oracle.candidates); deterministic in (seed, row); different rows / seeds give different draws."""
    from legommenders_b200 import _lib
    lib = _lib.load()
    n_items, K = 1000, 4
    rng = np.random.default_rng(0)
    for n_negs in (0, 1, 3, 4, 5, 37, 100):
        negs = np.arange(5000, 5000 + n_negs)            # distinct values outside the random-id range: provenance is visible
        for row in range(50):
            cand = draw(lib, 7, row, 123, negs, K, n_items)
            k = min(K, n_negs)
            true_part, rand_part = cand[1:1 + k], cand[1 + k:]
            assert cand[0] == 123 and len(cand) == K + 1
            assert len(set(true_part)) == k and all(5000 <= t < 5000 + n_negs for t in true_part)
            assert all(0 <= r < n_items for r in rand_part)
            assert cand == O.candidates(123, true_part, rand_part).tolist()                     # order restated by the oracle
            assert cand == draw(lib, 7, row, 123, negs, K, n_items)                            # pure function of (seed, row)
    a = [tuple(draw(lib, 7, r, 1, np.arange(5000, 5050), K, n_items)) for r in range(200)]
    b = [tuple(draw(lib, 8, r, 1, np.arange(5000, 5050), K, n_items)) for r in range(200)]
    assert len(set(a)) > 190 and sum(x == y for x, y in zip(a, b)) < 5


def test_candidate_draw_is_uniform():
    """random.sample semantics: every position equally likely in every slot (chi-square over 40k draws), uniform fill ids."""
    from legommenders_b200 import _lib
    lib = _lib.load()
    K, n_negs, R = 4, 10, 40000
    negs = np.arange(100, 100 + n_negs)
    counts = np.zeros((K, n_negs))
    for row in range(R):
        cand = draw(lib, 99, row, 0, negs, K, 50)
        for j in range(K):
            counts[j, cand[1 + j] - 100] += 1
    expected = R / n_negs
    chi2 = ((counts - expected) ** 2 / expected).sum(axis=1)       # 9 degrees of freedom per slot: P(chi2 > 33) < 1e-4
    assert (chi2 < 33).all(), chi2
    fill = np.zeros(50)
    for row in range(R):
        fill[draw(lib, 5, row, 0, np.zeros(0, dtype=np.int64), 1, 50)[1]] += 1
    assert (((fill - R / 50) ** 2) / (R / 50)).sum() < 100          # 49 dof: P(> 100) < 1e-4
