"""Where does a parity case diverge?  items / user / scores / loss of the CUDA path against the oracle: python scratch/diag_case.py <case>"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import numpy as np, torch
import cases, helpers
from legommenders_b200 import Env, builder
name = sys.argv[1]
c = cases.CASES[name]; g = cases.load(name)
world, llm = cases.make_world(c)
model, _, _ = builder.build_model(world, c['kind'], hidden=c['hidden'], heads=c['heads'], additive=c['additive'], dropout=0.0)
np_state, _ = helpers.oracle_state(c, world, llm)
builder.load_state(model, np_state)
batch = cases.unflatten_batch(g)
ora = helpers.oracle_run(c, world, llm, batch)
Env.test(); model.eval()
with torch.no_grad():
    items = model.get_item_content(copy.deepcopy(batch), 'item_id').cpu().numpy()
    user = model.get_user_content(copy.deepcopy(batch)).cpu().numpy()
    scores = model(batch=copy.deepcopy(batch)).cpu().numpy()
print('items ', helpers.normwise(items, ora['items']), items.shape)
print('user  ', helpers.normwise(user, ora['user']), user.shape)
print('scores', helpers.normwise(scores, ora['scores']), scores.shape)
if user.ndim == 3:
    for b in range(user.shape[0]):
        print(' user row', b, helpers.normwise(user[b], ora['user'][b]), 'hist len', int(batch['__clicks_mask__'][b].sum()))
