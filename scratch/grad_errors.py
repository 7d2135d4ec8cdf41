"""Per-tensor gradient error (max|d| / max|ref|, ref = fp64 oracle) of the autograd path and the native step driver on the parity cases."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import numpy as np, torch
import cases, helpers
from legommenders_b200 import Env, builder
from legommenders_b200.trainer import FlatAdam, NativeNRMSStep

for name in sys.argv[1:]:
    c = cases.CASES[name]; g = cases.load(name)
    world, llm = cases.make_world(c)
    model, _, _ = builder.build_model(world, c['kind'], hidden=c['hidden'], heads=c['heads'], additive=c['additive'], dropout=0.0)
    np_state, _ = helpers.oracle_state(c, world, llm)
    builder.load_state(model, np_state)
    batch = cases.unflatten_batch(g)
    ref = helpers.oracle_run(c, world, llm, batch, dtype=torch.float64)
    Env.train(); model.train()
    loss = model(batch=copy.deepcopy(batch)); loss.backward()
    ga = {n: p.grad.detach().cpu().numpy().copy() for n, p in model.named_parameters() if p.requires_grad}
    opt = FlatAdam(model, lr=1e-3); native = NativeNRMSStep(model, opt); opt.zero_grad()
    nl = native.fwd_bwd(copy.deepcopy(batch), training=False).item()
    gn = {n: p.grad.detach().cpu().numpy().copy() for n, p in model.named_parameters() if p.requires_grad}
    print(name, 'loss rel err autograd %.2e native %.2e' % (abs(loss.item() - ref['loss']) / abs(ref['loss']), abs(nl - ref['loss']) / abs(ref['loss'])))
    for k, r in ref['grads'].items():
        m = np.abs(r).max()
        l2 = np.linalg.norm(r)
        print('  %-60s max/max: autograd %.2e native %.2e   l2: autograd %.2e native %.2e' % (k, np.abs(ga[k] - r).max() / m, np.abs(gn[k] - r).max() / m,
              np.linalg.norm(ga[k] - r) / l2, np.linalg.norm(gn[k] - r) / l2))
