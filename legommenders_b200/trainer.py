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
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1

    def zero_grad(self):
        self.grad.zero_()
        for p in self.params:            # autograd accumulates into the existing views
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr():
                raise RuntimeError('parameter .grad was detached from the flat buffer')

    def allreduce(self):
        """Batch data-parallel: one flat-bucket NCCL allreduce (sum); the 1/world mean is folded into the Adam kernel."""
        if self.world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.group)

    def step(self, lr: float = None):
        self.step_count += 1
        self.allreduce()
        ops.adam_step(self.flat, self.grad, self.m, self.v, self.step_count, lr=self.lr if lr is None else lr,
                      beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, grad_scale=1.0 / self.world)
        ops.invalidate_weight_planes()   # the kernel updated the parameters behind torch's version counters
