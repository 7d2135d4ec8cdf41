"""The drop-in boundary seen FROM THE REFERENCE SIDE: the reference's own ClassHub discovers the B200 plugin classes, the
reference's own LegoConfig + Legommender build around them, and the state dict carries the reference's key names.

Runs only where the reference tree exists (this build container: /root/reference, or $LEGO_REF); the GPU box has no reference tree, so
the GPU twin below is skipped there and the committed goldens (minted from the same tree) carry the numbers instead.
"""
import copy
import os
import sys
import textwrap

import numpy as np
import pytest
import torch

import cases
import helpers
import ref_harness as rh

needs_ref = pytest.mark.skipif(not rh.available(), reason='reference tree not present')

PLUGIN_FILES = {
    'b200plug/operators/b200attention_operator.py': "B200AttentionOperator = integration.plugin('attention')",
    'b200plug/operators/b200cnn_operator.py': "B200CNNOperator = integration.plugin('cnn')",
    'b200plug/operators/b200ada_operator.py': "B200AdaOperator = integration.plugin('ada')",
    'b200plug/operators/b200pooling_operator.py': "B200PoolingOperator = integration.plugin('pooling')",
    'b200plug/operators/b200cnncat_operator.py': "B200CNNCatOperator = integration.plugin('cnncat')",
    'b200plug/operators/b200gru_operator.py': "B200GRUOperator = integration.plugin('gru')",
    'b200plug/operators/b200transformer_operator.py': "B200TransformerOperator = integration.plugin('transformer')",
    'b200plug/operators/b200fastformer_operator.py': "B200FastformerOperator = integration.plugin('fastformer')",
    'b200plug/operators/b200polyattention_operator.py': "B200PolyAttentionOperator = integration.plugin('polyattention')",
    'b200plug/predictors/b200dot_predictor.py': "B200DotPredictor = integration.plugin('dot', kind='predictor')",
    'b200plug/predictors/b200miner_predictor.py': "B200MINERPredictor = integration.plugin('miner', kind='predictor')",
}


@pytest.fixture
def plugin_tree(tmp_path, monkeypatch):
    """The files INTEGRATION.md tells a maintainer to add, written to a scratch package (the reference tree is read-only) and discovered
    with the reference's real ClassHub."""
    rh.activate()
    for rel, line in PLUGIN_FILES.items():
        p = tmp_path / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(textwrap.dedent(f'''\
            from legommenders_b200 import integration
            {line}
        '''))
    monkeypatch.chdir(tmp_path)                 # ClassHub globs a path relative to the working directory (class_hub.py:131-140)
    monkeypatch.syspath_prepend(str(tmp_path))
    from legommenders_b200 import integration
    yield tmp_path
    integration.unbind_env()
    for m in [m for m in sys.modules if m.startswith('b200plug')]:
        del sys.modules[m]


def discover():
    from loader.class_hub import ClassHub
    from model.operators.base_operator import BaseOperator
    from model.predictors.base_predictor import BasePredictor
    ops_hub = ClassHub(BaseOperator, os.path.join('b200plug', 'operators'), 'Operator')
    pred_hub = ClassHub(BasePredictor, os.path.join('b200plug', 'predictors'), 'Predictor')
    return ops_hub, pred_hub


@needs_ref
def test_classhub_discovers_plugins(plugin_tree):
    from model.operators.base_operator import BaseOperator as RefOp
    from legommenders_b200.operators.attention_operator import AttentionOperator
    ops_hub, pred_hub = discover()
    assert sorted(ops_hub.list()) == ['b200ada', 'b200attention', 'b200cnn', 'b200cnncat', 'b200fastformer', 'b200gru', 'b200polyattention',
                                      'b200pooling', 'b200transformer']      # yaml: meta.item: B200Attention -> lower()
    assert sorted(pred_hub.list()) == ['b200dot', 'b200miner']
    cls = ops_hub('B200Attention')
    assert issubclass(cls, RefOp) and issubclass(cls, AttentionOperator)
    assert not issubclass(AttentionOperator, RefOp)                                 # why the shim is needed at all


def build_through_reference(world, c, item_cls, user_cls, pred_cls):
    """loader/manager.py:139-153's build order with the REFERENCE's LegoConfig / Legommender and this package's EmbeddingHub."""
    from loader.column_map import ColumnMap
    from loader.env import Env as RefEnv
    from model.lego_config import LegoConfig
    from model.legommender import Legommender
    from legommenders_b200 import integration
    RefEnv.device = torch.device('cuda', 0) if torch.cuda.is_available() else torch.device('cpu')
    RefEnv.simple_dev = False
    RefEnv.train()
    for f in (RefEnv.set_item_cache, RefEnv.set_user_cache, RefEnv.set_lm_cache):
        f(False)
    integration.bind_reference_env()
    item_cfg = dict(num_attention_heads=c['heads'], attention_dropout=0.0, additive_hidden_size=c['additive'],
                    inputer_config=dict(use_cls_token=False, use_sep_token=True))
    user_cfg = dict(num_attention_heads=c['heads'], attention_dropout=0.0, additive_hidden_size=c['additive'],
                    inputer_config=dict(use_cls_token=False, use_sep_token=False))
    cfg = LegoConfig(hidden_size=c['hidden'], user_config=user_cfg, item_config=item_cfg, neg_count=4, use_neg_sampling=True,
                     use_item_content=True, item_page_size=0)
    cfg.set_component_classes(item_cls, user_cls, pred_cls)
    inputs = [world.title_col, 'category']
    cfg.set_item_ut(world.item_table(), inputs)
    cfg.set_user_ut(world.user_table(), ['history'])
    cfg.set_column_map(ColumnMap(neg_col='neg'))
    eh = integration.EmbeddingHub(embedding_dim=cfg.item_hidden_size, transformation='auto', transformation_dropout=0.0)
    eh.load_pretrained_embedding(world.word_table, vocab_name=world.word_vocab, frozen=True)
    eh.register_ut(world.item_table(), inputs)
    cfg.set_embedding_hub(eh)
    cfg.build_components()
    cfg.register_inputer_vocabs()
    return Legommender(cfg), cfg


@needs_ref
def test_reference_model_builds_around_plugins(plugin_tree):
    from loader.env import Env as RefEnv
    from legommenders_b200 import Env
    c = cases.CASES['nrms_small']
    world, llm = cases.make_world(c)
    ops_hub, pred_hub = discover()
    model, cfg = build_through_reference(world, c, ops_hub('b200attention'), ops_hub('b200attention'), pred_hub('b200dot'))
    assert type(model).__module__ == 'model.legommender'                              # the reference's own model class
    assert set(model.state_dict()) == set(helpers.state_shapes(c, world))             # the reference's parameter names
    for k, shape in helpers.state_shapes(c, world).items():
        assert tuple(model.state_dict()[k].shape) == tuple(shape), k
    assert model.item_op.inputer.max_sequence_len == world.title_len + 1 + 2
    # one global state: the reference's phase / cache switches are what this package's modules see
    RefEnv.test()
    assert Env.is_testing and not Env.is_training
    Env.set_item_cache(True)
    assert RefEnv.item_cache is True
    RefEnv.set_item_cache(False)
    RefEnv.train()
    assert Env.is_training and Env.device == RefEnv.device


@needs_ref
@pytest.mark.gpu
def test_reference_model_runs_on_plugins(plugin_tree):
    """GPU twin: the reference's Legommender.forward drives the B200 operators / hub / predictor; loss and gradients equal the golden."""
    c = cases.CASES['nrms_small']
    g = cases.load('nrms_small')
    world, llm = cases.make_world(c)
    ops_hub, pred_hub = discover()
    model, cfg = build_through_reference(world, c, ops_hub('b200attention'), ops_hub('b200attention'), pred_hub('b200dot'))
    np_state, _ = helpers.oracle_state(c, world, llm)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in np_state.items()})
    model.to('cuda')
    model.train()
    loss = model(batch=copy.deepcopy(cases.unflatten_batch(g)))
    loss.backward()
    assert abs(loss.item() - float(g['loss'])) <= 1e-4 * abs(float(g['loss']))
    for n, p in model.named_parameters():
        if p.requires_grad:
            ref = g['grad/' + n]
            assert np.abs(p.grad.cpu().numpy() - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-3), n
