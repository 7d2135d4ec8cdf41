"""Shared parity plumbing: run the oracle on a golden case, compare tensors norm-wise."""
from __future__ import annotations

import copy
from collections import OrderedDict

import numpy as np
import torch

import cases
from oracle import lego_oracle as O


def normwise(a, b) -> float:
    """max|a-b| / max|b| (SURVEY §8c parity protocol); 0 when both are identically zero."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max() if b.size else 0.0
    num = np.abs(a - b).max() if b.size else 0.0
    return float(num / den) if den > 0 else float(num)


STRICT_GRAD_FACTOR = 5.0


def grad_bound(c, ref_max: float, scale: float, tol: float) -> float:
    """Allowed max-abs error of one parameter gradient.

    Cases flagged `strict_grads` ("hot" states: every gradient is a few percent of the largest or more) are held PER TENSOR, relative to
    that tensor's own max, with no floor.  The factor is what the arithmetic supports and was measured (scratch/grad_errors.py, profiles/
    r2_01_gradient_precision.md): the tensor-core contractions carry operands as two bf16 planes (16 mantissa bits; loss and logits agree
    with fp64 to 1e-6..1e-5, inside the 1e-4 bar of BASELINE.json), and the additive-attention backward subtracts nearly equal row dots
    (ds = alpha*(x.g - sum alpha x.g)), which amplifies those 1e-5 forward differences to 1e-4..3.3e-4 on W1 / b1 / w2 gradients.  The exact
    fp32 SIMT path (LK_TC=0) holds 7e-6 on the same tensors; tests/test_gpu_model.py::test_exact_fp32_mode_gradients pins that.
    The plain-init cases keep a floor of 5 % of the largest gradient: their additive-attention gradients are ~1e-7 of it, below fp32
    resolution of the sums they come from."""
    if c.get('strict_grads'):
        return STRICT_GRAD_FACTOR * tol * ref_max
    return tol * max(ref_max, 5e-2 * scale)


def check_grads(c, grads: dict, ref_grads: dict, tol: float, golden=None):
    """grads vs the oracle's full gradients and, when given, the golden file's (full tensors or strided samples)."""
    scale = max(np.abs(v).max() for v in ref_grads.values())
    assert set(grads) == set(ref_grads)
    for k, ref in ref_grads.items():
        err = np.abs(grads[k] - ref).max()
        assert err <= grad_bound(c, np.abs(ref).max(), scale, tol), (k, err, np.abs(ref).max(), scale)
        if golden is None:
            continue
        if 'grad/' + k in golden.files:
            gref = golden['grad/' + k]
            assert np.abs(grads[k] - gref).max() <= grad_bound(c, np.abs(gref).max(), scale, tol), k
        else:
            gmax = float(golden['gradmax/' + k])
            assert np.abs(cases.sample_strided(grads[k]) - golden['gradsample/' + k]).max() <= grad_bound(c, gmax, scale, tol), k


def case_spec(c, world):
    return O.ModelSpec(c['kind'], c['heads'], {world.title_col: world.word_vocab, 'category': 'category'},
                       use_neg_sampling=c.get('use_neg_sampling', True), layers=c.get('layers', 0))


def state_shapes(c, world, llm=None):
    """Parameter shapes of the reference model for a case, without importing the reference."""
    E = llm.shape[1] if c['kind'] == 'llmid' else world.embed_dim
    return O.state_shapes(c['kind'], c['hidden'], c['additive'], E, world.n_words, world.n_cats, world.n_items, layers=c.get('layers', 0),
                          codes=c.get('codes', c['heads'] if c['kind'] == 'fastformer' else 0), code_dim=c.get('code_dim', 0))


def oracle_state(c, world, llm=None, dtype=torch.float32):
    np_state = cases.make_state(c, world, state_shapes(c, world, llm), llm)
    state = {}
    for k, v in np_state.items():
        t = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
        frozen = k.endswith('.embedding.weight')
        state[k] = t.requires_grad_(not frozen)
    return np_state, state


def oracle_run(c, world, llm, batch, dtype=torch.float32):
    """loss + grads + scores + item/user representations from the oracle for one batch."""
    _, state = oracle_state(c, world, llm, dtype)
    spec = case_spec(c, world)
    want = {}
    loss = O.forward(state, spec, copy.deepcopy(batch), want=want)
    loss.backward()
    # a parameter the forward never touches (CNNCat's inherited `linear`) has no gradient: zeros, as an optimiser would see it
    grads = {k: (v.grad.detach().numpy() if v.grad is not None else np.zeros(tuple(v.shape), dtype=np.float32)) for k, v in state.items()
             if v.requires_grad}
    with torch.no_grad():
        scores = O.forward(state, spec, copy.deepcopy(batch), return_scores=True)
    return dict(loss=float(loss.item()), grads=grads, scores=scores.numpy(), items=want['items'].detach().numpy(),
                user=want['user'].detach().numpy(), state=state, spec=spec)


def item_trees(world, kind, title_len=None):
    """Per-item token layouts from the oracle's restatement of the inputers (a1/a2)."""
    inputs = [world.title_col, 'category']
    max_lens = {world.title_col: world.title_len, 'category': None}
    tab = world.item_table()
    out = []
    for i in range(len(tab)):
        s = tab[i]
        if kind in ('nrms', 'miner', 'fastformer'):          # ConcatInputer operators; Fastformer's yaml switches the SEP tokens off
            out.append(O.concat_layout(s, inputs, max_lens, use_cls_token=False, use_sep_token=kind != 'fastformer'))
        else:
            out.append(O.simple_layout(s, inputs, max_lens))
    return out


def oracle_cached_eval(c, world, llm, state, spec):
    hist = np.stack([O.pad_history(h, world.hist_len)[0] for h in world.histories])
    hmask = np.stack([O.pad_history(h, world.hist_len)[1] for h in world.histories])
    hist_t, hmask_t = torch.from_numpy(hist), torch.from_numpy(hmask)
    item_repr = None
    if c['kind'] != 'llmid':
        item_repr = O.build_item_cache(state, spec, item_trees(world, c['kind']))
    else:
        # id path: history ids keep -1 padding through ConcatInputer (resampler.py:219-221 -> user_inputer(sample))
        hist_t = torch.where(hmask_t > 0, hist_t, torch.full_like(hist_t, -1))
    user_repr = O.build_user_cache(state, spec, item_repr, hist_t, hmask_t)
    if c['kind'] == 'llmid':
        with torch.no_grad():
            item_all = O.table_lookup(state, spec.item_vocab, torch.arange(world.n_items))
        scores = O.cached_scores(user_repr, item_all, torch.from_numpy(world.eval_users), torch.from_numpy(world.eval_items))
    else:
        scores = O.cached_scores(user_repr, item_repr, torch.from_numpy(world.eval_users), torch.from_numpy(world.eval_items))
    return item_repr, user_repr, scores
