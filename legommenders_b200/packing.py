"""Padding-free (packed / varlen) sequence layout.

The reference pads every item to S tokens and every history to H items, computes the pad positions, and masks them
away afterwards (pad keys get -inf in MHA, pad positions get weight 0 in AdditiveAttention, pad history slots hold the
encoding of item 0 and are masked in the user encoder — model/operators/attention_operator.py:49-58,
model/common/attention.py:31-38, loader/resampler.py:209-218).  None of those positions can influence a loss, a score or a
gradient, so the B200 path packs the valid tokens of the valid items into contiguous rows and runs every kernel on the
packed rows with int32 cumulative offsets.  MIND-small shape: 116,160 padded token rows per step -> ~40,000 packed rows.

All bookkeeping here is integer work on the id/mask tensors of the batch (done on the host when the batch is still on
the host, i.e. before the H2D copy, so it adds no device synchronisation).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import torch

from .env import Env


class Packed:
    __slots__ = ('ids', 'cu', 'n', 'rows', 'max_len')

    def __init__(self, ids, cu, n, rows, max_len):
        self.ids, self.cu, self.n, self.rows, self.max_len = ids, cu, n, rows, max_len


def pack_tokens(ids_by_col: Dict[str, torch.Tensor], mask: torch.Tensor, item_valid: Optional[torch.Tensor] = None,
                keep_empty: bool = False) -> Packed:
    """ids_by_col[c]: [N,S] int64, mask: [N,S] (1 = real token), item_valid: [N] or None.
    Returns packed ids per column [T] and cu [n+1] over the items that have at least one token (all items if keep_empty)."""
    m = mask > 0
    if item_valid is not None:
        m = m & (item_valid.reshape(-1, 1) > 0)
    lens = m.sum(dim=1)
    lens_h = lens.cpu() if lens.is_cuda else lens
    if not keep_empty:
        lens_h = lens_h[lens_h > 0]
    cu = torch.zeros(lens_h.numel() + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(lens_h, 0)
    rows = torch.nonzero(m.reshape(-1), as_tuple=False).reshape(-1)
    ids = OrderedDict((c, v.reshape(-1)[rows].to(Env.device, non_blocking=True)) for c, v in ids_by_col.items())
    max_len = int(lens_h.max().item()) if lens_h.numel() else 0
    return Packed(ids, cu.to(Env.device, non_blocking=True), lens_h.numel(), int(cu[-1].item()), max_len)


def pack_offsets(valid: torch.Tensor):
    """valid: [B,H] (1 = real history slot) -> (cu int32 [B+1] on device, max_len)."""
    lens = (valid > 0).sum(dim=1)
    lens_h = lens.cpu() if lens.is_cuda else lens
    cu = torch.zeros(lens_h.numel() + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(lens_h, 0)
    return cu.to(Env.device, non_blocking=True), int(lens_h.max().item()) if lens_h.numel() else 0
