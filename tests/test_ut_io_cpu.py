"""UniTok dataset directories: write a synthetic world in the layout `legommenders_b200.ut_io` restates, read it back, and
drive the host-side batch path from the loaded tables.  No GPU.  (The layout itself is unpinned — the `unitok` package is absent here; see
the module's header.)"""
import json
import os
import pickle

import numpy as np
import pytest

from legommenders_b200 import ut_io
from legommenders_b200.synth import MindWorld


@pytest.fixture(scope='module')
def world():
    return MindWorld(n_items=300, n_words=700, n_users=90, n_train=256, n_eval_groups=20, title_len=12, hist_len=9, embed_dim=16, seed=5)


def test_table_round_trip(world, tmp_path):
    t = world.item_table()
    ut_io.save_table(t, str(tmp_path / 'items'))
    assert sorted(os.listdir(tmp_path / 'items')) == ['category.vocab', 'data.pkl', 'glove.vocab', 'item_id.vocab', 'meta.json']
    meta = json.load(open(tmp_path / 'items' / 'meta.json'))
    assert [f['name'] for f in meta['features']] == ['item_id', 'title@glove', 'category']
    assert [f['key'] for f in meta['features']] == [True, False, False]
    back = ut_io.load_table(str(tmp_path / 'items'))
    assert len(back) == len(t) and back.key_feature.name == 'item_id'
    for name, f in t.meta.features.items():
        g = back.meta.features[name]
        assert (g.name, g.max_len, g.tokenizer.vocab.name, g.tokenizer.vocab.size) == (f.name, f.max_len, f.tokenizer.vocab.name,
                                                                                       f.tokenizer.vocab.size)
    assert back.meta.features['title@glove'].tokenizer.vocab is not back.meta.features['category'].tokenizer.vocab
    for i in (0, 17, len(t) - 1):
        assert back[i] == t[i]


def test_reader_accepts_the_older_layouts(world, tmp_path):
    """unitok < 4.3 calls features 'jobs' and names the key at top level; 3.x keeps data.npy and text vocabularies."""
    d = tmp_path / 'users'
    ut_io.save_table(world.user_table(), str(d))
    meta = json.load(open(d / 'meta.json'))
    for f in meta['features']:
        f.pop('key')
    meta['jobs'] = meta.pop('features')
    meta['key_job'] = 'user_id'
    json.dump(meta, open(d / 'meta.json', 'w'))
    data = pickle.load(open(d / 'data.pkl', 'rb'))
    os.remove(d / 'data.pkl')
    np.save(d / 'data.npy', np.array(data, dtype=object), allow_pickle=True)
    toks = pickle.load(open(d / 'user_id.vocab', 'rb'))
    os.remove(d / 'user_id.vocab')
    with open(d / 'tok.user_id.dat', 'w') as f:
        f.write(''.join(f'{t}\n' for t in toks))
    back = ut_io.load_table(str(d))
    assert back.key_feature.name == 'user_id' and len(back) == world.n_users
    assert back[3]['history'] == world.histories[3].tolist()
    assert back.meta.features['user_id'].tokenizer.vocab.size == world.n_users


def test_reader_rejects_inconsistent_directories(world, tmp_path):
    d = tmp_path / 'items'
    ut_io.save_table(world.item_table(), str(d))
    data = pickle.load(open(d / 'data.pkl', 'rb'))
    data['category'] = data['category'][:-1]
    pickle.dump(data, open(d / 'data.pkl', 'wb'))
    with pytest.raises(ValueError, match='samples'):
        ut_io.load_table(str(d))
    pickle.dump(list(range(5)), open(d / 'category.vocab', 'wb'))
    with pytest.raises(ValueError, match='tokens'):
        ut_io.load_table(str(d))


def test_dir_world_feeds_the_batch_path(world, tmp_path):
    """The loaded world gives the same training batches (candidates, histories, token layout) as the world it was written from."""
    from types import SimpleNamespace
    from legommenders_b200.batching import BatchBuilder
    from legommenders_b200.inputer.concat_inputer import ConcatInputer
    root = str(tmp_path / 'mind')
    ut_io.save_world(world, root)
    dw = ut_io.DirWorld(root, word_table=os.path.join(root, 'glove.npy'))
    assert (dw.n_items, dw.n_users, dw.n_words, dw.title_len, dw.hist_len, dw.n_cats) == (300, 90, 700, 12, 9, world.n_cats)
    assert dw.word_vocab == 'glove' and dw.embed_dim == 16 and np.array_equal(dw.word_table, world.word_table)
    assert np.array_equal(dw.train_users, world.train_users) and np.array_equal(dw.train_pos, world.train_pos)
    assert np.array_equal(dw.eval_click, world.eval_click) and np.array_equal(dw.eval_items, world.eval_items)
    assert all(np.array_equal(a, b) for a, b in zip(dw.histories, world.histories))
    assert all(np.array_equal(a, b) for a, b in zip(dw.negs, world.negs))
    rows = np.arange(24)
    batches = []
    for w in (world, dw):
        ut = w.item_table()
        inp = ConcatInputer(use_cls_token=False, use_sep_token=True, ut=ut, inputs=[w.title_col, 'category'], eh=None)
        layouts = [inp(ut[i]) for i in range(len(ut))]                       # Resampler.item_cache without a device
        batches.append(BatchBuilder(SimpleNamespace(item_cache=layouts), w, neg_count=4, seed=3, pin=False).train_batch(rows))

    def same(a, b):
        if isinstance(a, dict):
            return set(a) == set(b) and all(same(a[k], b[k]) for k in a)
        if hasattr(a, 'shape'):
            return a.shape == b.shape and bool((np.asarray(a) == np.asarray(b)).all())
        return a == b
    assert same(batches[0], batches[1])
