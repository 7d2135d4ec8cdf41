import sys, time, copy
sys.path.insert(0,'.'); sys.path.insert(0,'tests/golden')
import torch
import torch.nn.functional as F
from legommenders_b200.synth import MindWorld
import ref_harness as rh
from oracle import lego_oracle as O
w = MindWorld(n_items=300, n_words=800, n_users=80, n_train=64, n_eval_groups=20)

def split2(a):
    hi = a.bfloat16().float(); lo = (a-hi).bfloat16().float(); return hi, lo
def split3(a):
    hi = a.bfloat16().float(); r=a-hi; mid=r.bfloat16().float(); lo=(r-mid).bfloat16().float(); return hi,mid,lo
MODE='x3'
def smm(a, b):  # a [M,K] @ b [K,N]
    if MODE=='x1':
        return a.bfloat16().float() @ b.bfloat16().float()
    if MODE=='x3':
        ah,al = split2(a); bh,bl = split2(b)
        return ah@bh + (ah@bl + al@bh)
    if MODE=='x6':
        a0,a1,a2 = split3(a); b0,b1,b2=split3(b)
        return a0@b0 + (a0@b1+a1@b0) + (a0@b2+a1@b1+a2@b0)
    if MODE=='tf32':
        def t(x): return (x.view(torch.int32) & ~0x1fff).view(torch.float32)
        return t(a)@t(b)
    return a@b
class SLin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.hasb = b is not None
        y = smm(x.reshape(-1,x.shape[-1]), w.t()).reshape(*x.shape[:-1], w.shape[0])
        return y + b if b is not None else y
    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1]); x2 = x.reshape(-1, x.shape[-1])
        return smm(g2, w).reshape(x.shape), smm(g2.t().contiguous(), x2), (g2.sum(0) if ctx.hasb else None)
orig = F.linear
def run(mode, m, dtype=torch.float32):
    global MODE
    MODE=mode
    torch.manual_seed(1)
    model, rs, cfg, Env = rh.build_reference(w, m)
    from loader.data_set import DataSet
    from torch.utils.data import DataLoader
    import random; random.seed(5)
    ds = DataSet(w.train_table(), rs)
    batch = next(iter(DataLoader(ds, batch_size=32, num_workers=0)))
    state = {k: v.detach().clone().to(dtype if v.is_floating_point() else v.dtype).requires_grad_(v.requires_grad) for k,v in model.state_dict(keep_vars=True).items()}
    spec = O.ModelSpec(m, 8, {w.title_col:'glove','category':'category'})
    if mode not in ('fp32','fp64'):
        O.F.linear = lambda x,w_,b=None: SLin.apply(x,w_,b)
    else:
        O.F.linear = orig
    want={}
    sc = O.forward(state, spec, copy.deepcopy(batch), return_scores=True, want=want)
    loss = F.cross_entropy(sc, torch.zeros(sc.shape[0],dtype=torch.long))
    loss.backward()
    O.F.linear = orig
    return sc.detach().double(), loss.item(), {k:v.grad.double() for k,v in state.items() if v.requires_grad}
for m in ['nrms','naml']:
    ref = run('fp64', m, torch.float64)
    for mode in ['fp32','x1','tf32','x3','x6']:
        r = run(mode, m)
        e_sc = ((r[0]-ref[0]).abs().max()/ref[0].abs().max()).item()
        e_g = max(((r[2][k]-ref[2][k]).abs().max()/ (ref[2][k].abs().max()+1e-30)).item() for k in ref[2] if ref[2][k].abs().max()>1e-7)
        print(m, mode, 'scores normwise %.2e'%e_sc, 'loss rel %.2e'%(abs(r[1]-ref[1])/ref[1]), 'worst grad normwise %.2e'%e_g, flush=True)
