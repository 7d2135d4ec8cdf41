"""Oracle against the LIVE reference modules (only where /root/reference exists, i.e. the build container).

The committed goldens (tests/golden/*.npz) are what travels; this test re-derives one small case from the reference's own
Legommender / Resampler on the spot, so a stale golden or a drifted oracle shows up here first.
"""
import copy
import os
import random
import sys

import numpy as np
import pytest
import torch

import cases
import helpers

REF = os.environ.get('LEGO_REF', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'model')), reason='reference tree not present on this box')


@pytest.mark.parametrize('name', ['nrms_small', 'naml_small'])
def test_oracle_matches_live_reference(name):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    import ref_harness as rh
    c = cases.CASES[name]
    torch.manual_seed(0)
    world, llm = cases.make_world(c)
    model, resampler, cfg, Env = rh.build_reference(world, c['kind'], hidden=c['hidden'], heads=c['heads'], additive=c['additive'],
                                                    dropout=0.0, use_neg_sampling=c.get('use_neg_sampling', True), llm_item_table=llm)
    from loader.data_set import DataSet
    from torch.utils.data import DataLoader
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    state = cases.make_state(c, world, shapes, llm)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    model.train()
    Env.train()
    random.seed(c['seed'] + 99)      # a different draw than the committed golden
    batch = next(iter(DataLoader(DataSet(world.train_table(), resampler), batch_size=c['batch'], num_workers=0, shuffle=False)))
    loss = model(batch=copy.deepcopy(batch))
    loss.backward()
    ora = helpers.oracle_run(c, world, llm, batch)
    assert abs(ora['loss'] - loss.item()) <= 1e-6 * abs(loss.item())
    refs = {n: p.grad.numpy() for n, p in model.named_parameters() if p.requires_grad}
    scale = max(np.abs(v).max() for v in refs.values())
    for n, ref in refs.items():   # per tensor, with a floor for tensors whose gradient is pure cancellation noise (~1e-9)
        assert np.abs(ora['grads'][n] - ref).max() <= 1e-5 * max(np.abs(ref).max(), 5e-2 * scale), n
