"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares (no compute without a GPU),
host-side layouts / resampling follow the oracle's integer restatement, and the product refuses to run without CUDA."""
import os
import random
import re
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import cases
import helpers
from oracle import lego_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'legommenders_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(lk_\w+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from legommenders_b200 import _lib
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
        assert name in _lib.SIGNATURES, f'{name} has no ctypes signature'
    assert set(_lib.SIGNATURES) == set(declared)
    assert b'sm_100a' in lib.lk_version()


def test_ctypes_signatures_match_the_header_arity_and_structs():
    """The binding is hand-written: every entry point's argument COUNT must equal the header's, and the two structs that cross the
    boundary by pointer must have the layout the C compiler gives them (sizes checked against the documented field lists)."""
    import ctypes
    from legommenders_b200 import _lib, ops
    text = open(os.path.join(ROOT, 'include', 'legommenders_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    decls = re.findall(r'\b(?:int|int64_t|size_t|const char\*|void|unsigned long long)\s+(lk_\w+)\s*\(([^;{]*?)\)\s*;', text, flags=re.S)
    assert len(decls) == len(_lib.SIGNATURES)
    for name, args in decls:
        a = args.strip()
        n = 0 if a in ('', 'void') else len(a.split(','))
        assert len(_lib.SIGNATURES[name][0]) == n, f'{name}: header has {n} arguments, the binding {len(_lib.SIGNATURES[name][0])}'
    # struct lk_gemm_epilogue: 17 fields; lk_split_seg: 7 x 8 bytes; lk_colsum_job: 5 x 8 + int (padded to 48)
    assert ctypes.sizeof(ops.SplitSeg) == 56 and ctypes.sizeof(ops.ColsumJob) == 48
    fields = [f[0] for f in ops.GemmEpilogue._fields_]
    body = re.search(r'typedef struct lk_gemm_epilogue \{(.*?)\} lk_gemm_epilogue;', text, flags=re.S).group(1)
    declared = re.findall(r'(\w+);', body)
    assert fields == declared, (fields, declared)
    fields = [f[0] for f in ops.ChainStage._fields_]
    body = re.search(r'typedef struct lk_chain_stage \{(.*?)\} lk_chain_stage;', text, flags=re.S).group(1)
    assert fields == re.findall(r'(\w+);', body)


def test_workspace_queries_do_not_need_a_gpu():
    from legommenders_b200 import _lib
    assert _lib.query('lk_linear_bwd_weight_workspace_bytes', 116160, 256, 256) > 256 * 256 * 4
    assert _lib.query('lk_scatter_add_workspace_bytes', 1000, 18, 256) > 0
    assert _lib.query('lk_colsum_workspace_bytes', 1000, 256) > 0


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from legommenders_b200 import Env, builder
    world, _ = cases.make_world(cases.CASES['nrms_small'])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        builder.build_model(world, 'nrms', hidden=64, additive=32)
    with pytest.raises(RuntimeError):
        from legommenders_b200 import ops
        ops.gather_add(None, torch.zeros(3, dtype=torch.long), None, torch.zeros(4, 8))


def test_inputer_layouts_match_oracle():
    from legommenders_b200.inputer import ConcatInputer, SimpleInputer
    world, _ = cases.make_world(cases.CASES['nrms_small'])
    ut = world.item_table()
    inputs = [world.title_col, 'category']
    max_lens = {world.title_col: world.title_len, 'category': None}
    for cls_tok, sep_tok in [(False, True), (True, True), (True, False), (False, False)]:
        inp = ConcatInputer(ut=ut, inputs=inputs, eh=None, use_cls_token=cls_tok, use_sep_token=sep_tok)
        assert inp.max_sequence_len == world.title_len + 1 + int(cls_tok) + 2 * int(sep_tok)
        for i in range(0, len(ut), 7):
            got = inp(dict(ut[i]))
            ref = O.concat_layout(ut[i], inputs, max_lens, cls_tok, sep_tok)
            assert list(got['input_ids']) == list(ref['input_ids'])
            for k in ref['input_ids']:
                assert got['input_ids'][k].dtype == torch.int64
                assert np.array_equal(got['input_ids'][k].numpy(), ref['input_ids'][k])
            assert np.array_equal(got['attention_mask'].numpy(), ref['attention_mask'])
    sinp = SimpleInputer(ut=ut, inputs=inputs, eh=None)
    for i in range(0, len(ut), 7):
        got = sinp(dict(ut[i]))
        ref = O.simple_layout(ut[i], inputs, max_lens)
        for k in ref['input_ids']:
            assert np.array_equal(got['input_ids'][k].numpy(), ref['input_ids'][k])
            assert np.array_equal(got['attention_mask'][k].numpy(), ref['attention_mask'][k])


def _fake_config(world, kind='nrms', neg_count=4):
    from legommenders_b200 import ColumnMap
    from legommenders_b200.inputer import ConcatInputer, SimpleInputer
    item_ut, user_ut = world.item_table(), world.user_table()
    inputs = [world.title_col, 'category']
    if kind == 'nrms':
        item_inp = ConcatInputer(ut=item_ut, inputs=inputs, eh=None, use_cls_token=False, use_sep_token=True)
    else:
        item_inp = SimpleInputer(ut=item_ut, inputs=inputs, eh=None)
    user_inp = ConcatInputer(ut=user_ut, inputs=['history'], eh=None, use_cls_token=False, use_sep_token=False)
    return SimpleNamespace(use_item_content=True, cm=ColumnMap(neg_col='neg'), item_ut=item_ut, user_ut=user_ut,
                           item_operator=SimpleNamespace(inputer=item_inp), user_operator=SimpleNamespace(inputer=user_inp),
                           user_operator_class=SimpleNamespace(flatten_mode=False), use_neg_sampling=True, neg_count=neg_count)


@pytest.mark.parametrize('name', ['nrms_small', 'naml_small'])
def test_resampler_reproduces_reference_batch(name):
    """Bit-exact indices/masks: same `random` seed -> the committed reference batch."""
    from torch.utils.data import DataLoader
    from legommenders_b200 import DataSet, Env, Resampler
    c = cases.CASES[name]
    g = cases.load(name)
    world, _ = cases.make_world(c)
    Env.train()
    Env.set_item_cache(False)
    Env.set_user_cache(False)
    rs = Resampler(_fake_config(world, c['kind']))
    random.seed(c['seed'])
    batch = next(iter(DataLoader(DataSet(world.train_table(), rs), batch_size=c['batch'], num_workers=0, shuffle=False)))
    flat = cases.flatten_tree(batch)
    assert sorted('batch/' + k for k in flat) == sorted(k for k in g.files if k.startswith('batch/'))
    for k, v in flat.items():
        assert np.array_equal(v.numpy(), g['batch/' + k]), k


def test_resampler_cache_modes_and_candidate_order():
    from legommenders_b200 import DataSet, Env, Resampler
    world, _ = cases.make_world(cases.CASES['nrms_small'])
    rs = Resampler(_fake_config(world))
    tab = world.train_table()
    Env.train()
    Env.set_item_cache(False)
    Env.set_user_cache(False)
    random.seed(1)
    s = rs(dict({k: (list(v) if isinstance(v, list) else v) for k, v in tab[3].items()}))
    assert s['item_id']['attention_mask'].shape[0] == 5 and 'neg' not in s
    # item cache on: ids only, candidate 0 is the positive, negatives come from the user's neg list first
    Env.set_item_cache(True)
    random.seed(1)
    row = {k: (list(v) if isinstance(v, list) else v) for k, v in tab[3].items()}
    negs = list(row['neg'])
    s = rs(row)
    ids = s['item_id'].tolist()
    assert ids[0] == tab[3]['item_id'] and len(ids) == 5
    if len(negs) >= 4:
        assert all(i in negs for i in ids[1:])
    h, m = O.pad_history(tab[3]['history'], world.hist_len)
    assert np.array_equal(s['history'].numpy(), h) and np.array_equal(s['__clicks_mask__'].numpy(), m)
    # user cache on: history is dropped; evaluation: no negative sampling
    Env.set_user_cache(True)
    Env.test()
    s = rs({k: (list(v) if isinstance(v, list) else v) for k, v in tab[3].items()})
    assert 'history' not in s and s['item_id'].tolist() == [tab[3]['item_id']]
    Env.set_item_cache(False)
    Env.set_user_cache(False)
    Env.train()


def test_embedding_hub_errors_and_names():
    from legommenders_b200 import EmbeddingHub, Env
    from legommenders_b200.synth import Vocab
    with pytest.raises(ValueError):
        EmbeddingHub(64, 'bogus', 0.1)
    eh = EmbeddingHub(64, 'auto', 0.1)
    with pytest.raises(ValueError):
        eh.load_pretrained_embedding(np.zeros((4, 8), np.float32))
    with pytest.raises(ValueError):
        eh.load_pretrained_embedding(np.zeros((4, 8), np.float32), vocab_name='a', col_name='b')
    Env.device = torch.device('cpu')            # registration only allocates parameters; no kernel runs
    eh.load_pretrained_embedding(np.zeros((10, 48), np.float32), vocab_name='glove', frozen=True)
    eh.register_vocab(Vocab('glove', 10))
    eh.register_vocab(Vocab('category', 18))
    with pytest.raises(ValueError, match='conflict'):
        eh.register_vocab(Vocab('category', 19))
    names = {k for k, _ in eh.vocab_table.named_parameters()}
    assert names == {'glove.embedding.weight', 'glove.linear.weight', 'glove.linear.bias', 'category.weight'}
    assert not eh.vocab_table['glove'].embedding.weight.requires_grad
    eh2 = EmbeddingHub(48, 'auto', 0.1)        # same width -> no projection
    eh2.load_pretrained_embedding(np.zeros((10, 48), np.float32), vocab_name='glove', frozen=False)
    eh2.register_vocab(Vocab('glove', 10))
    assert {k for k, _ in eh2.vocab_table.named_parameters()} == {'glove.weight'}
    assert eh2.vocab_table['glove'].weight.requires_grad
    with pytest.raises(ValueError, match='does not match'):
        eh3 = EmbeddingHub(48, 'auto', 0.1)
        eh3.load_pretrained_embedding(np.zeros((10, 48), np.float32), vocab_name='glove')
        eh3.register_vocab(Vocab('glove', 11))
    Env.device = None


def test_checkpoint_interchange_with_torch_adam(tmp_path):
    """ FlatAdam / LinearWarmupSchedule state dicts are the reference's torch.optim.Adam / LambdaLR ones: a reference
    checkpoint loads into the flat buffers, and a checkpoint written here loads into torch.optim.Adam (base_lego.py:228-265)."""
    from transformers import get_linear_schedule_with_warmup
    from legommenders_b200.trainer import FlatAdam, LinearWarmupSchedule, load_checkpoint, save_checkpoint
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    ref[0].bias.requires_grad_(False)                                # a frozen parameter is skipped by both optimisers
    ropt = torch.optim.Adam(filter(lambda p: p.requires_grad, ref.parameters()), lr=1e-3)
    rsch = get_linear_schedule_with_warmup(ropt, num_warmup_steps=3, num_training_steps=10)
    lrs = []
    for i in range(4):
        ropt.zero_grad()
        ref(torch.full((2, 6), float(i + 1))).sum().backward()
        ropt.step(); rsch.step()
        lrs.append(rsch.get_last_lr()[0])
    path = str(tmp_path / 'ref.pt')
    torch.save(dict(model=ref.state_dict(), optimizer=ropt.state_dict(), scheduler=rsch.state_dict()), path)

    mine = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    mine[0].bias.requires_grad_(False)
    opt = FlatAdam(mine, lr=1e-3)
    sch = LinearWarmupSchedule(opt, 3, 10)
    assert [round(sch.base_lr * sch.factor(k), 12) for k in range(1, 5)] == [round(x, 12) for x in lrs]
    load_checkpoint(path, mine, opt, sch)
    for a, b in zip(mine.parameters(), ref.parameters()):
        assert torch.equal(a, b)
        assert a.data_ptr() >= opt.flat.data_ptr() or not a.requires_grad       # still views of the flat buffer
    assert opt.step_count == 4 and abs(opt.lr - lrs[-1]) < 1e-15
    rstate = ropt.state_dict()['state']
    for i, (p, off, n) in enumerate(opt._slices()):
        assert torch.equal(opt.m[off:off + n].view_as(p), rstate[i]['exp_avg'])
        assert torch.equal(opt.v[off:off + n].view_as(p), rstate[i]['exp_avg_sq'])
    # and back: our checkpoint loads into the reference's optimiser / scheduler objects
    out = str(tmp_path / 'mine.pt')
    save_checkpoint(out, mine, opt, sch)
    sd = torch.load(out, weights_only=False)
    fresh = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    fresh[0].bias.requires_grad_(False)
    fresh.load_state_dict(sd['model'], strict=True)
    fopt = torch.optim.Adam(filter(lambda p: p.requires_grad, fresh.parameters()), lr=1.0)
    fopt.load_state_dict(sd['optimizer'])
    fsch = get_linear_schedule_with_warmup(fopt, num_warmup_steps=3, num_training_steps=10)
    fsch.load_state_dict(sd['scheduler'])
    assert fsch.last_epoch == 4 and abs(fsch.get_last_lr()[0] - lrs[-1]) < 1e-15
    fs = fopt.state_dict()['state']
    assert all(torch.equal(fs[i]['exp_avg'], rstate[i]['exp_avg']) and float(fs[i]['step']) == 4.0 for i in rstate)
    with pytest.raises(ValueError):
        opt.load_state_dict(dict(state={}, param_groups=[dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, params=[0])]))
