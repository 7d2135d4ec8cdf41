"""Training-step plumbing: flat parameter/gradient buffers + the fused Adam kernel, optional NCCL gradient allreduce.

Mirrors what base_lego.py:175-223 + trainer.py:190-204 do around `loss.backward()`: Adam(lr, betas=(0.9,0.999), eps=1e-8)
over all trainable parameters, stepped every batch (accumulate_batch = 1).  All trainable parameters are re-pointed into
one contiguous fp32 buffer (and their .grad into another) so that the optimiser is ONE kernel launch and data-parallel
training needs ONE allreduce.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


class FlatAdam:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, process_group=None):
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        dev = self.params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]      # keep every slice 16-byte aligned
        total = sum(sizes)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.symm = None
        self.grad = self._peer_bucket(total, dev) if (self.world > 1 and dev.type == 'cuda') else None
        if self.grad is None:
            self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p, sz in zip(self.params, sizes):
            n = p.numel()
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)
            p.grad = self.grad[off:off + n].view_as(p)
            off += sz
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        self.time_allreduce = None

    def _peer_bucket(self, total, dev):
        """The gradient bucket in symmetric memory (peer-mapped over NVLink by torch.distributed's rendezvous: plumbing) so that the all-reduce
        can be this library's one-kernel slice reduction (lk_allreduce_p2p).  None -> NCCL all-reduce (LK_P2P_ALLREDUCE=0, or no peer access)."""
        import ctypes
        import os
        if os.environ.get('LK_P2P_ALLREDUCE', '1') == '0':
            return None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            group = self.group if self.group is not None else dist.group.WORLD
            grad = symm_mem.empty(total, dtype=torch.float32, device=dev)
            grad.zero_()
            self.symm = symm_mem.rendezvous(grad, group=group)
            self._peer_ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in self.symm.buffer_ptrs])
            self._flag_ptrs = None
            self._mc_ptr = None
            if os.environ.get('LK_P2P_FUSED_BARRIER', '1') != '0':
                self._flags = symm_mem.empty(4096, dtype=torch.int32, device=dev)       # LK_ALLREDUCE_FLAG_WORDS
                self._flags.zero_()
                fh = symm_mem.rendezvous(self._flags, group=group)
                self._flag_ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in fh.buffer_ptrs])
                self._epoch = 0
                mc = int(getattr(self.symm, 'multicast_ptr', 0) or 0)
                if mc and os.environ.get('LK_P2P_MULTIMEM', '1') != '0':
                    self._mc_ptr = mc                              # NVSwitch multicast mapping: in-switch reduction + broadcast
            torch.cuda.synchronize(dev)
            dist.barrier(group=group)                           # every rank's flags are zero before anybody signals
            return grad
        except Exception as e:   # noqa: BLE001
            import sys
            print(f'[legommenders_b200] peer-memory gradient bucket unavailable ({type(e).__name__}: {e}); using ncclAllReduce', file=sys.stderr)
            self.symm = None
            return None

    def _allreduce_now(self):
        if self.symm is not None:
            from ._lib import call
            import ctypes
            if self._flag_ptrs is not None:                    # one launch: rendezvous, slice reduction, rendezvous
                self._epoch += 1
                call('lk_allreduce_p2p', ctypes.addressof(self._peer_ptrs), ctypes.addressof(self._flag_ptrs), self._mc_ptr, self._epoch, self.symm.rank,
                     self.world, self.grad.numel(), 1.0)
                return
            self.symm.barrier(channel=0)                       # every rank has written its gradients
            call('lk_allreduce_p2p', ctypes.addressof(self._peer_ptrs), None, None, 0, self.symm.rank, self.world, self.grad.numel(), 1.0)
            self.symm.barrier(channel=1)                       # every slice's sum is visible everywhere
        else:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.group)

    def zero_grad(self):
        self.grad.zero_()
        for p in self.params:            # autograd accumulates into the existing views
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr():
                raise RuntimeError('parameter .grad was detached from the flat buffer')

    def allreduce(self):
        """Batch data-parallel: one all-reduce (sum) of the flat bucket — lk_allreduce_p2p over NVLink peer memory when the bucket is in
        symmetric memory, ncclAllReduce otherwise; the 1/world mean is folded into the Adam kernel."""
        if self.world > 1:
            if self.time_allreduce is not None:        # bench.py: CUDA events around the collective (local gradients ready -> reduced)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                self._allreduce_now()
                e1.record()
                self.time_allreduce.append((e0, e1))
            else:
                self._allreduce_now()

    def step(self, lr: float = None):
        self.step_count += 1
        self.allreduce()
        ops.adam_step(self.flat, self.grad, self.m, self.v, self.step_count, lr=self.lr if lr is None else lr,
                      beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, grad_scale=1.0 / self.world)
        ops.invalidate_weight_planes()   # the kernel updated the parameters behind torch's version counters

    # ---- checkpoint compatibility with the reference's torch.optim.Adam (base_lego.py:198-204, 228-267) ----------------
    def _slices(self):
        off = 0
        for p in self.params:
            n = p.numel()
            yield p, off, n
            off += (n + 3) // 4 * 4

    def state_dict(self) -> dict:
        """The state dict `torch.optim.Adam(filter(requires_grad, model.parameters()))` would hold after the same steps: per-parameter
        `step` / `exp_avg` / `exp_avg_sq` keyed by the parameter's position, one param group."""
        state = {}
        if self.step_count > 0:
            for i, (p, off, n) in enumerate(self._slices()):
                state[i] = dict(step=torch.tensor(float(self.step_count)), exp_avg=self.m[off:off + n].view_as(p).clone(),
                                exp_avg_sq=self.v[off:off + n].view_as(p).clone())
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                     capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False, params=list(range(len(self.params))))
        return dict(state=state, param_groups=[group])

    def load_state_dict(self, sd: dict):
        """Accepts a `torch.optim.Adam` state dict of the same parameter list (a reference checkpoint's `optimizer` entry)."""
        groups = sd['param_groups']
        if len(groups) != 1:
            raise ValueError('FlatAdam holds one parameter group (the reference\'s two-group item_lr set-up is not on this path)')
        g = groups[0]
        if len(g['params']) != len(self.params):
            raise ValueError(f'optimizer state has {len(g["params"])} parameters, the model has {len(self.params)} trainable ones')
        self.lr, self.betas, self.eps = float(g['lr']), tuple(g['betas']), float(g['eps'])
        if g.get('weight_decay', 0) or g.get('amsgrad', False) or g.get('maximize', False):
            raise ValueError('weight_decay / amsgrad / maximize are not used by the reference and not supported')
        self.m.zero_(); self.v.zero_()
        steps = set()
        for pos, (p, off, n) in zip(g['params'], self._slices()):
            st = sd['state'].get(pos)
            if st is None:
                continue
            if tuple(st['exp_avg'].shape) != tuple(p.shape):
                raise ValueError(f'optimizer state {pos} has shape {tuple(st["exp_avg"].shape)}, parameter has {tuple(p.shape)}')
            self.m[off:off + n].copy_(st['exp_avg'].reshape(-1))
            self.v[off:off + n].copy_(st['exp_avg_sq'].reshape(-1))
            steps.add(int(float(st['step'])))
        if len(steps) > 1:
            raise ValueError(f'parameters were stepped a different number of times: {sorted(steps)}')
        self.step_count = steps.pop() if steps else 0


class LinearWarmupSchedule:
    """`transformers.get_linear_schedule_with_warmup` (base_lego.py:211-223): lr = base * step / n_warmup while warming up, then a linear
    decay to 0 at `n_training`.  `step()` after every optimiser step; state dict interchangeable with the reference's LambdaLR."""

    def __init__(self, opt: FlatAdam, n_warmup: int, n_training: int):
        self.opt, self.n_warmup, self.n_training = opt, int(n_warmup), int(n_training)
        self.base_lr = opt.lr
        self.last_epoch = 0
        opt.lr = self.base_lr * self.factor(0)

    def factor(self, step: int) -> float:
        if step < self.n_warmup:
            return float(step) / float(max(1, self.n_warmup))
        return max(0.0, float(self.n_training - step) / float(max(1, self.n_training - self.n_warmup)))

    def step(self):
        self.last_epoch += 1
        self.opt.lr = self.base_lr * self.factor(self.last_epoch)

    def get_last_lr(self):
        return [self.opt.lr]

    def state_dict(self) -> dict:
        return dict(base_lrs=[self.base_lr], last_epoch=self.last_epoch, verbose=False, _step_count=self.last_epoch + 1,
                    _get_lr_called_within_step=False, _last_lr=[self.opt.lr], lr_lambdas=[None])

    def load_state_dict(self, sd: dict):
        self.base_lr = float(sd['base_lrs'][0])
        self.last_epoch = int(sd['last_epoch'])
        self.opt.lr = self.base_lr * self.factor(self.last_epoch)


def save_checkpoint(path: str, model: torch.nn.Module, opt: FlatAdam, scheduler: LinearWarmupSchedule = None):
    """base_lego.py:255-265 — `{model, optimizer, scheduler}` with the reference's key names, loadable by the reference."""
    sd = dict(model={k: v.detach().clone() for k, v in model.state_dict().items()}, optimizer=opt.state_dict(),
              scheduler=scheduler.state_dict() if scheduler is not None else {})
    torch.save(sd, path)


def load_checkpoint(path: str, model: torch.nn.Module, opt: FlatAdam = None, scheduler: LinearWarmupSchedule = None, strict: bool = True,
                    model_only: bool = False):
    """base_lego.py:228-253.  Parameters live in FlatAdam's flat buffer: values are copied INTO the existing storage (views stay valid)."""
    sd = torch.load(path, map_location=next(model.parameters()).device, weights_only=False)
    missing = model.load_state_dict(sd['model'], strict=strict)
    ops.invalidate_weight_planes()
    if not model_only:
        if opt is not None:
            opt.load_state_dict(sd['optimizer'])
        if scheduler is not None and sd.get('scheduler'):
            scheduler.load_state_dict(sd['scheduler'])
    return missing


class NativeNRMSStep:
    """Training step of the NRMS configuration through the native driver `lk_nrms_fwd_bwd` (csrc/lk_nrms_step.cu):
    host packing -> one C-ABI call for forward+backward -> (allreduce) -> one Adam launch.  Same kernels and the same
    mathematics as the autograd path in legommender.py; it only removes the per-launch Python/autograd cost."""

    ENC = ['multi_head_attention.in_proj_weight', 'multi_head_attention.in_proj_bias', 'multi_head_attention.out_proj.weight',
           'multi_head_attention.out_proj.bias', 'linear.weight', 'linear.bias', 'additive_attention.encoder.0.weight',
           'additive_attention.encoder.0.bias', 'additive_attention.encoder.2.weight']

    def __init__(self, model, opt: FlatAdam, sharded_table=None):
        """sharded_table: a sharding.ShardedTable holding the frozen title-word table row-sharded over the ranks (config 4).
        Every step then exchanges the batch's DISTINCT token ids (all-to-all), receives their rows (all-to-all) and gathers
        from that compact per-step table; the replicated table of the model is not touched."""
        import numpy as np
        from . import _lib
        from .embedding_hub import Table, Transformation
        from .inputer.concat_inputer import ConcatInputer
        from .operators.attention_operator import AttentionOperator
        cfgm = model.config
        if not (isinstance(model.item_op, AttentionOperator) and isinstance(model.user_op, AttentionOperator)
                and isinstance(model.item_op.inputer, ConcatInputer) and model.use_neg_sampling and cfgm.use_item_content):
            raise ValueError('NativeNRMSStep needs the NRMS configuration (Attention item/user operators, negative sampling)')
        inp = model.item_op.inputer
        if len(inp.inputs) != 2 or not inp.use_sep_token:
            raise ValueError('NativeNRMSStep expects item inputs [title, category] with SEP tokens')
        self.title_col, self.cat_col, self.special_col = inp.inputs[0], inp.inputs[1], inp.vocab.name
        tv = inp.ut.meta.features[self.title_col].tokenizer.vocab.name
        cv = inp.ut.meta.features[self.cat_col].tokenizer.vocab.name
        ttab, ctab, stab = model.eh(tv), model.eh(cv), model.eh(self.special_col)
        if not (isinstance(ttab, Transformation) and isinstance(ctab, Table) and isinstance(stab, Table)):
            raise ValueError('NativeNRMSStep expects a projected (pretrained) title table and plain category / special tables')
        if ttab.embedding.weight.requires_grad:
            raise ValueError('NativeNRMSStep expects the pretrained title table to be frozen')
        self.model, self.opt = model, opt
        self.glove = ttab.embedding.weight
        self.sharded_table = sharded_table
        if sharded_table is not None and sharded_table.local.shape[1] != self.glove.shape[1]:
            raise ValueError('sharded table width differs from the title embedding width')
        mha = model.item_op.multi_head_attention
        self.D, self.heads = mha.embed_dim, mha.num_heads
        self.A = model.item_op.additive_attention.hidden_size
        self.E = self.glove.shape[1]
        self.n_cats, self.n_special = ctab.weight.shape[0], stab.weight.shape[0]
        self.drop_embed, self.drop_attn = float(ttab.p), float(mha.dropout)
        if model.user_op.multi_head_attention.num_heads != self.heads or model.user_op.multi_head_attention.dropout != mha.dropout:
            raise ValueError('NativeNRMSStep expects identical head count / dropout in both encoders')
        named = dict(model.named_parameters())
        names = [f'embedding_vocab_table.{tv}.linear.weight', f'embedding_vocab_table.{tv}.linear.bias',
                 f'embedding_vocab_table.{cv}.weight', f'embedding_vocab_table.{self.special_col}.weight']
        names += [f'item_op.{n}' for n in self.ENC] + [f'user_op.{n}' for n in self.ENC]
        base = opt.flat.data_ptr()
        offs = []
        for n in names:
            p = named[n]
            off = p.data_ptr() - base
            if off < 0 or off % 4 or off // 4 + p.numel() > opt.flat.numel():
                raise ValueError(f'{n} does not live in the flat parameter buffer')
            offs.append(off // 4)
        if len(named) - 1 != len(names):   # every trainable parameter must be covered (the frozen table is the only other one)
            extra = set(k for k, v in named.items() if v.requires_grad) - set(names)
            if extra:
                raise ValueError(f'parameters not handled by the native step: {sorted(extra)}')
        self.offsets = np.asarray(offs, dtype=np.int64)
        self._lib = _lib
        self.arena = None
        self.loss = torch.zeros((), dtype=torch.float32, device=opt.flat.device)
        self.calls = 0
        self.check_every = 64

    def _ensure_arena(self, T, N, B, C):
        """Grow-only arena sized by the driver's own sizing pass, with 20% head-room so that it is queried rarely."""
        cap = getattr(self, '_cap', None)
        if self.arena is None or T > cap[0] or N > cap[1] or B > cap[2] or C > cap[3]:
            cap = (int(T * 1.2) + 64, int(N * 1.2) + 8, B, C)
            need = self._lib.query('lk_nrms_arena_bytes', cap[0], cap[1], B, C, self.D, self.A, self.E, self.heads, self.n_cats, self.n_special)
            self.arena = torch.empty(int(need), dtype=torch.uint8, device=self.opt.flat.device)
            self._cap = cap

    def pack(self, batch):
        """Host-side integer bookkeeping (packing.py); cached on device-resident batches."""
        from .packing import pack_offsets, pack_tokens
        meta = batch.get('__lk_packed__')
        if meta is None:
            cm = self.model.cm
            cand, hist = batch[cm.item_col], batch[cm.history_col]
            B, C, S = cand['attention_mask'].shape
            H = hist['attention_mask'].shape[1]
            clicks = batch[cm.mask_col]
            ids = {c: torch.cat([cand['input_ids'][c].reshape(B * C, S), hist['input_ids'][c].reshape(B * H, S)])
                   for c in cand['input_ids']}
            mask = torch.cat([cand['attention_mask'].reshape(B * C, S), hist['attention_mask'].reshape(B * H, S)])
            valid = torch.cat([torch.ones(B * C, dtype=clicks.dtype, device=clicks.device), clicks.reshape(-1)])
            pk = pack_tokens(ids, mask, valid, keep_empty=False)
            cu_u, max_u = pack_offsets(clicks)
            meta = (pk, cu_u, max_u, B, C)
            if mask.is_cuda:
                batch['__lk_packed__'] = meta
        return meta

    def fwd_bwd(self, batch, training=True):
        """Enqueue forward + backward; gradients land in opt.grad, the loss (device scalar) is returned."""
        from ._lib import call, ptr
        pk, cu_u, max_u, B, C = self.pack(batch)
        self._ensure_arena(pk.rows, pk.n, B, C)
        self.calls += 1
        seed = (torch.initial_seed() * 1000003 + self.calls) & ((1 << 60) - 1)
        de, da = (self.drop_embed, self.drop_attn) if training else (0.0, 0.0)
        title_ids, table = pk.ids[self.title_col], self.glove
        if self.sharded_table is not None:
            # config 4: rows of this step's distinct tokens arrive over NVLink; `inverse` (-1 where the position is unset) indexes them
            table, title_ids = self.sharded_table.lookup_unique(title_ids)
            if table.shape[0] == 0:
                table = table.new_zeros((1, table.shape[1]))
            if self.calls % self.check_every == 0 and hasattr(self.sharded_table, 'check'):
                self.sharded_table.check()               # bucket overflow of the device lookup plan (a device->host read, hence rarely)
        call('lk_nrms_fwd_bwd', ptr(title_ids), ptr(pk.ids[self.cat_col]), ptr(pk.ids[self.special_col]), ptr(pk.cu),
             pk.n, pk.rows, pk.max_len, ptr(cu_u), B, C, max_u, ptr(table), table.shape[0], ptr(self.opt.flat), ptr(self.opt.grad),
             self.offsets.ctypes.data, self.D, self.heads, self.A, self.E, self.n_cats, self.n_special, float(de), float(da), int(seed),
             ptr(self.loss), None, ptr(self.arena), self.arena.numel())
        return self.loss

    def step(self, batch, lr=None):
        loss = self.fwd_bwd(batch, training=True)
        self.opt.step(lr)
        return loss
