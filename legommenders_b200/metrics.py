"""MetricPool on the device (mirror of utils/metrics.py:248-369 for the group metrics the reference evaluates with).

`MetricPool.parse(['GAUC', 'MRR', 'NDCG@1', 'NDCG@5', 'NDCG@10']).calculate(scores, labels, groups)` keeps the reference's call
shape and result (`OrderedDict(name -> float)`), but the per-group work is ONE kernel (`lk_group_metrics`,
csrc/lk_metrics.cu) instead of a pandas groupby + a process pool calling sklearn per group.
Names follow utils/metrics.py: GAUC (:100-108), MRR (:144-160, the "modified" MRR), NDCG@k (:223-235).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Sequence

import numpy as np
import torch

from ._lib import call, ptr, query, raise_on_bad_ids, workspace
from .env import Env

SUPPORTED = ('GAUC', 'MRR', 'NDCG')


class MetricPool:
    def __init__(self, names: Sequence[str]):
        self.names: List[str] = list(names)
        self.ks: List[int] = []
        for n in self.names:
            base, _, k = n.partition('@')
            if base.upper() not in SUPPORTED:
                raise ValueError(f'metric {n} is not a group metric of this path (supported: GAUC, MRR, NDCG@k)')
            if base.upper() == 'NDCG':
                if not k:
                    raise ValueError('NDCG needs a cut-off, e.g. NDCG@10')
                if int(k) not in self.ks:
                    self.ks.append(int(k))
        if len(self.ks) > 8:
            raise ValueError('at most 8 distinct NDCG cut-offs')
        self.values = OrderedDict()
        self.group = True

    @classmethod
    def parse(cls, names: Sequence[str]):
        return cls(names)

    def calculate(self, scores, labels, groups, group_worker: int = 5, per_group: bool = False):
        """scores: fp32 [R]; labels, groups: int64 [R] — device tensors are used in place, anything else is copied up once."""
        dev = Env.device
        s = torch.as_tensor(scores, dtype=torch.float32).to(dev).contiguous()
        y = torch.as_tensor(labels).to(dev).to(torch.int64).contiguous()          # widened on the device, not on the host
        g = torch.as_tensor(groups).to(dev).to(torch.int64).contiguous()
        R = s.numel()
        if not (y.numel() == R and g.numel() == R):
            raise ValueError('scores, labels and groups must have the same length')
        if R == 0:
            raise ValueError('no rows to evaluate')
        nk = len(self.ks)
        n_disc = max(self.ks + [1]) + 1
        disc = np.concatenate([[0.0], np.cumsum(1.0 / np.log2(np.arange(n_disc - 1) + 2.0))])   # sklearn's discount, fp64
        disc_d = torch.from_numpy(disc).to(dev)
        ks = np.asarray(self.ks + [0] * (8 - nk), dtype=np.int32)
        out = torch.empty(3 + nk, dtype=torch.float64, device=dev)
        pg = torch.empty((2 + nk, R), dtype=torch.float32, device=dev) if per_group else None
        nbytes = query('lk_group_metrics_workspace_bytes', R, nk)
        ws = workspace(nbytes, dev, 'metrics')
        call('lk_group_metrics', ptr(s), ptr(y), ptr(g), R, ks.ctypes.data, nk, ptr(disc_d), n_disc, ptr(out), ptr(pg), ptr(ws), ws.numel())
        host = out.cpu().numpy()          # the one device->host read of the evaluation
        raise_on_bad_ids('evaluation')     # the stream is drained here anyway: surface out-of-range user / item ids as IndexError
        self.n_groups = int(host[2 + nk])
        self.values = OrderedDict()
        for n in self.names:
            base, _, k = n.partition('@')
            b = base.upper()
            v = host[0] if b == 'GAUC' else host[1] if b == 'MRR' else host[2 + self.ks.index(int(k))]
            self.values[n] = float(np.float32(v))
        if per_group:
            self.per_group = pg[:, :self.n_groups]
        return self.values
