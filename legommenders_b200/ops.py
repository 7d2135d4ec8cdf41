"""Differentiable host-side wrappers over the C-ABI kernels (torch.autograd.Function = plumbing only).

Every forward/backward here is one or more calls into liblegommenders_b200.so on the current stream;
no torch math runs on the hot path.  Each op cites the reference call site it replaces.
"""
from __future__ import annotations

import ctypes
import os
import weakref

import torch
from torch.autograd import Function

from . import _lib
from ._lib import call, id_violation_counter, ptr, query, raise_on_bad_ids, workspace

ACT_NONE, ACT_TANH, ACT_RELU = 0, 1, 2
POOL_MEAN, POOL_MAX, POOL_SUM = 0, 1, 2


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda:
        raise RuntimeError(f'expected a CUDA float32 tensor, got {t.dtype} on {t.device}')
    return t if t.is_contiguous() else t.contiguous()


def _i64(t):
    if t is None:
        return None
    if t.dtype != torch.int64 or not t.is_cuda:
        raise RuntimeError(f'expected a CUDA int64 tensor, got {t.dtype} on {t.device}')
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------------------
# dense contractions.  Two engines behind one interface:
#   * tensor-core path (lk_tc_gemm: tcgen05 + TMEM + TMA on split-bf16 planes) for every shape it supports
#   * exact-fp32 SIMT path (lk_linear_*) for tiny / odd shapes and when LK_TC=0
# ----------------------------------------------------------------------------------------------------
USE_TC = os.environ.get('LK_TC', '1') != '0'
TC_MIN_ROWS = 256


class Planes:
    """Split-bf16 image of an fp32 matrix [rows, cols]: hi/lo bf16 planes with pitch `ld` (multiple of 8, zero padded)."""
    __slots__ = ('hi', 'lo', 'rows', 'cols', 'ld')

    def __init__(self, hi, lo, rows, cols, ld):
        self.hi, self.lo, self.rows, self.cols, self.ld = hi, lo, rows, cols, ld


def split_planes(x2: torch.Tensor, transpose: bool = False, colsum: bool = False):
    """-> Planes, or (Planes, column sums of x2) when colsum=True (bias gradient fused into the split pass)."""
    rows, cols = x2.shape
    orows, ocols = (cols, rows) if transpose else (rows, cols)
    ld = (ocols + 7) // 8 * 8
    buf = torch.empty((2, orows, ld), dtype=torch.bfloat16, device=x2.device)
    cs = ws = None
    nbytes = 0
    if colsum:
        cs = torch.empty((cols,), dtype=torch.float32, device=x2.device)
        nbytes = query('lk_split_bf16_workspace_bytes', rows, cols)
        ws = workspace(nbytes, x2.device, 'colsum')
    call('lk_split_bf16', ptr(x2), rows, cols, x2.stride(0), ptr(buf[0]), ptr(buf[1]), ld, int(transpose), ptr(cs), ptr(ws),
         ws.numel() if ws is not None else 0)
    planes = Planes(buf[0], buf[1], orows, ocols, ld)
    return (planes, cs) if colsum else planes


class SplitSeg(ctypes.Structure):
    """struct lk_split_seg (include/legommenders_b200.h)"""
    _fields_ = [('X', ctypes.c_void_p), ('hi', ctypes.c_void_p), ('lo', ctypes.c_void_p), ('rows', ctypes.c_int64),
                ('cols', ctypes.c_int64), ('ld_in', ctypes.c_int64), ('ld_out', ctypes.c_int64)]


def im2col_planes(x2: torch.Tensor, S: int, taps: int) -> Planes:
    """Planes of the im2col image [rows, taps*C] of token rows x2 [rows, C] (sequences of S rows, zero beyond their ends)."""
    rows, C = x2.shape
    ld = (taps * C + 7) // 8 * 8
    buf = torch.empty((2, rows, ld), dtype=torch.bfloat16, device=x2.device)
    call('lk_im2col_split_bf16', ptr(x2), rows, S, C, taps, ptr(buf[0]), ptr(buf[1]), ld)
    return Planes(buf[0], buf[1], rows, taps * C, ld)


def split_planes_multi(mats):
    """Planes of several fp32 matrices (e.g. all weights of a step) in ONE launch (lk_split_bf16_multi)."""
    segs = (SplitSeg * len(mats))()
    planes = []
    for sg, x in zip(segs, mats):
        rows, cols = x.shape
        ld = (cols + 7) // 8 * 8
        buf = torch.empty((2, rows, ld), dtype=torch.bfloat16, device=x.device)
        planes.append(Planes(buf[0], buf[1], rows, cols, ld))
        sg.X, sg.hi, sg.lo, sg.rows, sg.cols, sg.ld_in, sg.ld_out = ptr(x), ptr(buf[0]), ptr(buf[1]), rows, cols, x.stride(0), ld
    call('lk_split_bf16_multi', ctypes.addressof(segs), len(mats))
    return planes


class ColsumJob(ctypes.Structure):
    """struct lk_colsum_job (include/legommenders_b200.h)"""
    _fields_ = [('part', ctypes.c_void_p), ('out', ctypes.c_void_p), ('nparts', ctypes.c_int64), ('cols', ctypes.c_int64),
                ('stride', ctypes.c_int64), ('accumulate', ctypes.c_int)]


def colsum_finish_multi(jobs):
    """jobs: list of (part [nparts, stride>=cols] fp32, out [cols] fp32, accumulate) -> every out[c] (+)= sum_i part[i, c], ONE launch."""
    arr = (ColsumJob * len(jobs))()
    for a, (part, out, acc) in zip(arr, jobs):
        a.part, a.out, a.nparts, a.cols, a.stride, a.accumulate = ptr(part), ptr(out), part.shape[0], out.numel(), part.stride(0), int(acc)
    call('lk_colsum_finish_multi', ctypes.addressof(arr), len(jobs))


_wplanes = {}


def weight_planes(w: torch.Tensor, transpose: bool = False) -> Planes:
    """Planes of a parameter, cached per tensor OBJECT until it is modified in place (optimizer steps bump `_version`;
    kernels that update parameters behind torch's back call invalidate_weight_planes)."""
    key = (id(w), transpose)
    ent = _wplanes.get(key)
    if ent is not None:
        ver, ref, dptr, planes = ent
        if ref() is w and ver == w._version and dptr == w.data_ptr():
            return planes
    if len(_wplanes) > 256:
        for k in [k for k, e in _wplanes.items() if e[1]() is None]:
            del _wplanes[k]
    planes = split_planes(w.detach(), transpose)
    _wplanes[key] = (w._version, weakref.ref(w), w.data_ptr(), planes)
    return planes


def invalidate_weight_planes():
    _wplanes.clear()


CONV_TC_FWD = os.environ.get('LK_CONV_TC_FWD', '0') == '1'


def tc_ok(M, N, K) -> bool:
    return USE_TC and M >= TC_MIN_ROWS and N % 4 == 0 and K % 4 == 0 and N >= 16 and K >= 16


def tc_gemm(A: Planes, B: Planes, mn_major: bool, GM, GN, GK, out=None, bias=None, rowmask=None, act=0, drop_p=0.0, seed=0,
            accumulate=False):
    y = out if out is not None else torch.empty((GM, GN), dtype=torch.float32, device=A.hi.device)
    nbytes = query('lk_tc_gemm_workspace_bytes', GM, GN, GK)
    ws = workspace(nbytes, y.device, 'tc')
    call('lk_tc_gemm', ptr(A.hi), ptr(A.lo), A.ld, int(mn_major), ptr(B.hi), ptr(B.lo), B.ld, int(mn_major), ptr(y), y.stride(0),
         GM, GN, GK, ptr(bias), ptr(rowmask), act, float(drop_p), int(seed), int(accumulate), ptr(ws), ws.numel())
    return y


class GemmEpilogue(ctypes.Structure):
    """struct lk_gemm_epilogue (include/legommenders_b200.h)"""
    _fields_ = [('bias', ctypes.c_void_p), ('rowmask', ctypes.c_void_p), ('rowmask_is_ids', ctypes.c_int), ('act', ctypes.c_int),
                ('drop_p', ctypes.c_float), ('seed', ctypes.c_uint64), ('accumulate', ctypes.c_int), ('store_c_off', ctypes.c_int),
                ('add_ids0', ctypes.c_void_p), ('add_tab0', ctypes.c_void_p), ('add_ids1', ctypes.c_void_p), ('add_tab1', ctypes.c_void_p),
                ('out_hi', ctypes.c_void_p), ('out_lo', ctypes.c_void_p), ('ld_planes', ctypes.c_int64), ('colsum', ctypes.c_void_p),
                ('colsum_part', ctypes.c_void_p)]


class ChainStage(ctypes.Structure):
    """struct lk_chain_stage (include/legommenders_b200.h)"""
    _fields_ = [('w_hi', ctypes.c_void_p), ('w_lo', ctypes.c_void_p), ('ldw', ctypes.c_int64), ('bias', ctypes.c_void_p),
                ('addsrc', ctypes.c_void_p), ('act', ctypes.c_int), ('out_f32', ctypes.c_void_p), ('out_hi', ctypes.c_void_p),
                ('out_lo', ctypes.c_void_p), ('ld_planes', ctypes.c_int64), ('colsum_part', ctypes.c_void_p), ('dotvec', ctypes.c_void_p),
                ('rowdot_part', ctypes.c_void_p)]


def tc_chain(A: Planes, stages, b_mn=False):
    """lk_tc_chain: up to three chained [M,256]x[256,256] contractions with the intermediates kept on chip.
    stages: list of dicts(w=Planes, bias=, addsrc=, act=, want_f32=, want_planes=, want_colsum=, dotvec=).
    Returns per stage a dict with the requested outputs (f32, planes, colsum_part [tiles*4,256], rowdot_part [M,4])."""
    M = A.rows
    dev = A.hi.device
    arr = (ChainStage * len(stages))()
    outs = []
    for a, s in zip(arr, stages):
        o = {}
        w = s['w']
        a.w_hi, a.w_lo, a.ldw = ptr(w.hi), ptr(w.lo), w.ld
        a.bias, a.addsrc, a.act = ptr(s.get('bias')), ptr(s.get('addsrc')), int(s.get('act', 0))
        if s.get('want_f32'):
            o['f32'] = torch.empty((M, 256), dtype=torch.float32, device=dev)
            a.out_f32 = ptr(o['f32'])
        if s.get('want_planes'):
            o['planes'] = Planes(torch.empty((M, 256), dtype=torch.bfloat16, device=dev), torch.empty((M, 256), dtype=torch.bfloat16, device=dev),
                                 M, 256, 256)
            a.out_hi, a.out_lo, a.ld_planes = ptr(o['planes'].hi), ptr(o['planes'].lo), 256
        if s.get('want_colsum'):
            o['colsum_part'] = torch.zeros(((M + 127) // 128 * 4, 256), dtype=torch.float32, device=dev)
            a.colsum_part = ptr(o['colsum_part'])
        if s.get('dotvec') is not None:
            o['rowdot_part'] = torch.zeros((M, 4), dtype=torch.float32, device=dev)
            a.dotvec, a.rowdot_part = ptr(s['dotvec']), ptr(o['rowdot_part'])
        outs.append(o)
    call('lk_tc_chain', ptr(A.hi), ptr(A.lo), A.ld, M, ctypes.addressof(arr), len(stages), int(b_mn))
    return outs


def merge_topk(vals: torch.Tensor, idx: torch.Tensor, k: int):
    """k-way selection over per-range (or per-rank) candidates: vals / idx [U, C] -> ([U, k], [U, k]); descending score, ties towards the
    smaller item index (the order torch.topk / a stable descending sort of the full score row gives)."""
    o1 = torch.argsort(idx, dim=1, stable=True)
    v1, i1 = torch.gather(vals, 1, o1), torch.gather(idx, 1, o1)
    o2 = torch.argsort(v1, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(v1, 1, o2), torch.gather(i1, 1, o2)


def sweep_topk(Up: Planes, Ip: Planes, k: int = 10, item_offset: int = 0):
    """lk_sweep_topk + range merge: for every user row of `Up` the k best rows of `Ip` by dot product, scores never materialised.
    -> (scores [U, k] fp32, item ids [U, k] int64 = row in Ip + item_offset)."""
    U, N = Up.rows, Ip.rows
    dev = Up.hi.device
    R = query('lk_sweep_ranges', U, N)
    pv = torch.empty((R, U, k), dtype=torch.float32, device=dev)
    pi = torch.empty((R, U, k), dtype=torch.int32, device=dev)
    call('lk_sweep_topk', ptr(Up.hi), ptr(Up.lo), Up.ld, U, ptr(Ip.hi), ptr(Ip.lo), Ip.ld, N, Up.cols, k, R, ptr(pv), ptr(pi))
    vals, idx = merge_topk(pv.permute(1, 0, 2).reshape(U, R * k), pi.permute(1, 0, 2).reshape(U, R * k).to(torch.int64), k)
    return vals, idx + item_offset


def tc_gemm_ex(A: Planes, B: Planes, GM, GN, GK, b_mn=False, a_mn=False, out=None, store_c=True, bias=None, rowmask=None,
               rowmask_is_ids=False, act=0, drop_p=0.0, seed=0, accumulate=False, add0=None, add1=None, want_planes=False,
               want_colsum=False):
    """lk_tc_gemm_ex: the contraction with its fused epilogue (row addends from small tables, split-bf16 plane output, column
    sums).  add0/add1 = (ids int64 [GM], table fp32 [*, GN]).  Returns (C or None, Planes or None, colsum or None)."""
    dev = A.hi.device
    y = out
    if y is None and store_c:
        y = torch.empty((GM, GN), dtype=torch.float32, device=dev)
    ep = GemmEpilogue()
    keep = [bias, rowmask, add0, add1]
    ep.bias, ep.rowmask, ep.rowmask_is_ids, ep.act = ptr(bias), ptr(rowmask), int(rowmask_is_ids), act
    ep.drop_p, ep.seed, ep.accumulate, ep.store_c_off = float(drop_p), int(seed), int(accumulate), int(not store_c)
    if add0 is not None:
        ep.add_ids0, ep.add_tab0 = ptr(add0[0]), ptr(add0[1])
    if add1 is not None:
        ep.add_ids1, ep.add_tab1 = ptr(add1[0]), ptr(add1[1])
    planes = cs = None
    if want_planes:
        ld = (GN + 7) // 8 * 8
        buf = torch.zeros((2, GM, ld), dtype=torch.bfloat16, device=dev)
        planes = Planes(buf[0], buf[1], GM, GN, ld)
        ep.out_hi, ep.out_lo, ep.ld_planes = ptr(buf[0]), ptr(buf[1]), ld
    if want_colsum:
        cs = torch.empty((GN,), dtype=torch.float32, device=dev)
        ep.colsum = ptr(cs)
    nbytes = query('lk_tc_gemm_workspace_bytes', GM, GN, GK)
    ws = workspace(nbytes, dev, 'tc')
    call('lk_tc_gemm_ex', ptr(A.hi), ptr(A.lo), A.ld, int(a_mn), ptr(B.hi), ptr(B.lo), B.ld, int(b_mn), ptr(y),
         y.stride(0) if y is not None else GN, GM, GN, GK, ctypes.addressof(ep), ptr(ws), ws.numel())
    del keep
    return y, planes, cs


def linear_fwd_raw(x2, w, b, rowmask, act, out=None, accumulate=False, drop_p=0.0, seed=0, xp=None):
    """y = epilogue(x·Wᵀ).  Returns (y, planes-of-x or None)."""
    M, K = x2.shape
    N = w.shape[0]
    if tc_ok(M, N, K):
        xp = xp if xp is not None else split_planes(x2)
        y = tc_gemm(xp, weight_planes(w), False, M, N, K, out=out, bias=b, rowmask=rowmask, act=act, drop_p=drop_p, seed=seed,
                    accumulate=accumulate)
        return y, xp
    y = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=x2.device)
    call('lk_linear_fwd', ptr(x2), ptr(w), ptr(b), ptr(rowmask), ptr(y), M, N, K, act, int(accumulate), float(drop_p), int(seed))
    return y, None


def linear_bwd_data_raw(dy2, w, out=None, accumulate=False, dyp=None):
    """dX = dY·W  (contraction over N: B operand is Wᵀ's planes, K-major)."""
    M, N = dy2.shape
    K = w.shape[1]
    if tc_ok(M, K, N):
        dyp = dyp if dyp is not None else split_planes(dy2)
        return tc_gemm(dyp, weight_planes(w, transpose=True), False, M, K, N, out=out, accumulate=accumulate)
    dx = out if out is not None else torch.empty((M, K), dtype=torch.float32, device=dy2.device)
    call('lk_linear_bwd_data', ptr(dy2), ptr(w), ptr(dx), M, N, K, int(accumulate))
    return dx


def linear_bwd_weight_raw(dy2, x2, want_bias=True, dyp=None, xp=None, db=None):
    """dW = dYᵀ·X (contraction over the M token rows: both operands MN-major), db = column sums of dY."""
    M, N = dy2.shape
    K = x2.shape[1] if x2 is not None else xp.cols
    if tc_ok(M, N, K):
        if dyp is None:
            dyp, db = split_planes(dy2, colsum=True) if want_bias else (split_planes(dy2), None)
        xp = xp if xp is not None else split_planes(x2)
        dw = tc_gemm(dyp, xp, True, N, K, M)
        if want_bias and db is None:
            db = colsum_raw(dy2)
        return dw, db
    if x2 is None:
        raise RuntimeError('linear_bwd_weight_raw: fp32 input required for the SIMT path')
    dw = torch.empty((N, K), dtype=torch.float32, device=dy2.device)
    db = torch.empty((N,), dtype=torch.float32, device=dy2.device) if want_bias else None
    nbytes = query('lk_linear_bwd_weight_workspace_bytes', M, N, K)
    ws = workspace(nbytes, dy2.device, 'wgrad')
    call('lk_linear_bwd_weight', ptr(dy2), ptr(x2), ptr(dw), ptr(db), M, N, K, 0, ptr(ws), ws.numel())
    return dw, db


def colsum_raw(x2):
    M, N = x2.shape
    out = torch.empty((N,), dtype=torch.float32, device=x2.device)
    nbytes = query('lk_colsum_workspace_bytes', M, N)
    ws = workspace(nbytes, x2.device, 'colsum')
    call('lk_colsum', ptr(x2), ptr(out), M, N, 0, ptr(ws), ws.numel())
    return out


def act_bwd_raw(dy2, y2, rowmask, act, drop_p=0.0, seed=0):
    M, N = dy2.shape
    out = torch.empty_like(dy2)
    call('lk_act_bwd', ptr(dy2), ptr(y2), ptr(rowmask), ptr(out), M, N, act, float(drop_p), int(seed))
    return out


class _Linear(Function):
    """y = dropout(act(x·Wᵀ + b)) * rowmask — nn.Linear call sites (embedding_hub.py:95-96,
    attention_operator.py:56, cnn_operator.py:62, attention.py:17-19) and MHA in/out projections."""

    @staticmethod
    def forward(ctx, x, w, b, rowmask, act, drop_p, seed):
        x, w = _f32(x), _f32(w)
        b = _f32(b) if b is not None else None
        x2 = x.reshape(-1, x.shape[-1])
        rm = _i64(rowmask.reshape(-1)) if rowmask is not None else None
        y, xp = linear_fwd_raw(x2, w, b, rm, act, drop_p=drop_p, seed=seed)
        ctx.act, ctx.drop_p, ctx.seed, ctx.has_b = act, drop_p, seed, b is not None
        plain = act == ACT_NONE and rm is None and drop_p == 0.0
        ctx.plain = plain
        ctx.xp = xp                      # split-bf16 image of x (same bytes as x) reused by the weight gradient
        ctx.save_for_backward(x2 if xp is None else None, w, None if plain else y, rm)
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w, y, rm = ctx.saved_tensors
        dy2 = _f32(dy).reshape(-1, dy.shape[-1])
        if not ctx.plain:
            dy2 = act_bwd_raw(dy2, y, rm, ctx.act, ctx.drop_p, ctx.seed)
        dx = dw = db = None
        M, N = dy2.shape
        K = w.shape[1]
        dyp = dbias = None
        if tc_ok(M, K, N) or tc_ok(M, N, K):
            dyp, dbias = split_planes(dy2, colsum=True) if ctx.has_b else (split_planes(dy2), None)
        if ctx.needs_input_grad[0]:
            dx = linear_bwd_data_raw(dy2, w, dyp=dyp).view(*dy.shape[:-1], K)
        if ctx.needs_input_grad[1] or (ctx.has_b and ctx.needs_input_grad[2]):
            dw, db = linear_bwd_weight_raw(dy2, x2, want_bias=ctx.has_b, dyp=dyp, xp=ctx.xp, db=dbias)
        ctx.xp = None
        return dx, dw, db, None, None, None, None


def linear(x, w, b=None, rowmask=None, act=ACT_NONE, drop_p=0.0, seed=0):
    return _Linear.apply(x, w, b, rowmask, act, drop_p, seed)


# ----------------------------------------------------------------------------------------------------
# embedding gather
# ----------------------------------------------------------------------------------------------------
def valid_mask(ids: torch.Tensor) -> torch.Tensor:
    """mask = (ids > -1).long() — concat_inputer.py:108."""
    ids = _i64(ids)
    out = torch.empty_like(ids)
    call('lk_valid_mask', ptr(ids), ptr(out), ids.numel())
    return out


def scatter_add_rows(ids, mask, src2, table_shape, scale=None, row_div=1):
    """dtable = Σ_p valid(p)·scale[p]·src[p // row_div] at row ids[p] (sorted segmented reduction)."""
    V, E = table_shape
    P = ids.numel()
    dt = torch.empty((V, E), dtype=torch.float32, device=src2.device)
    nbytes = query('lk_scatter_add_workspace_bytes', P, V, E)
    ws = workspace(nbytes, src2.device, 'scatter')
    call('lk_scatter_add_sorted', ptr(ids), ptr(mask), ptr(src2), ptr(scale), row_div, ptr(dt), P, V, E, 0, ptr(ws), ws.numel())
    return dt


class _GatherAdd(Function):
    """out = base + valid·table[ids] (base may be None) — aten::embedding + mask multiply + add at
    concat_inputer.py:105-113 / simple_inputer.py:51-64."""

    @staticmethod
    def forward(ctx, base, ids, mask, table):
        ids, mask, table = _i64(ids), _i64(mask), _f32(table)
        M, E = ids.numel(), table.shape[1]
        if base is None:
            out = torch.empty((*ids.shape, E), dtype=torch.float32, device=table.device)
            acc = 0
        else:
            out = _f32(base).clone()
            acc = 1
        id_violation_counter(table.device)
        call('lk_gather_rows', ptr(ids), ptr(mask), ptr(table), table.shape[0], ptr(out), M, E, acc)
        ctx.save_for_backward(ids, mask)
        ctx.tshape = tuple(table.shape)
        ctx.has_base = base is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        ids, mask = ctx.saved_tensors
        dbase = dout if (ctx.has_base and ctx.needs_input_grad[0]) else None
        dtable = None
        if ctx.needs_input_grad[3]:
            d2 = _f32(dout).reshape(-1, dout.shape[-1])
            dtable = scatter_add_rows(ids.reshape(-1), None if mask is None else mask.reshape(-1), d2, ctx.tshape)
        return dbase, None, None, dtable


def gather_add(base, ids, mask, table):
    return _GatherAdd.apply(base, ids, mask, table)


class _GatherPool(Function):
    """Fused gather + masked pooling (north_star piece 1; pooling_operator.py:46-56 over an nn.Embedding)."""

    @staticmethod
    def forward(ctx, ids, mask, table, mode):
        ids, mask, table = _i64(ids), _i64(mask), _f32(table)
        N, S = ids.shape
        E = table.shape[1]
        out = torch.empty((N, E), dtype=torch.float32, device=table.device)
        id_violation_counter(table.device)
        call('lk_gather_pool', ptr(ids), ptr(mask), ptr(table), table.shape[0], ptr(out), N, S, E, mode)
        ctx.save_for_backward(ids, mask)
        ctx.tshape, ctx.mode = tuple(table.shape), mode
        return out

    @staticmethod
    def backward(ctx, dout):
        ids, mask = ctx.saved_tensors
        if not ctx.needs_input_grad[2]:
            return None, None, None, None
        if ctx.mode == POOL_MAX:
            raise RuntimeError('gather_pool: backward of max pooling is not implemented')
        N, S = ids.shape
        scale = None
        if ctx.mode == POOL_MEAN:
            m = mask if mask is not None else valid_mask(ids)
            cnt = m.sum(dim=1, keepdim=True).to(torch.float32)   # tiny [N,1] host-side bookkeeping
            scale = (1.0 / (cnt + 1e-8)).expand(N, S).contiguous().reshape(-1)
        dtable = scatter_add_rows(ids.reshape(-1), None if mask is None else mask.reshape(-1),
                                  _f32(dout), ctx.tshape, scale=scale, row_div=S)
        return None, None, dtable, None


def gather_pool(ids, mask, table, mode=POOL_MEAN):
    return _GatherPool.apply(ids, mask, table, mode)


# ----------------------------------------------------------------------------------------------------
# multi-head self-attention core
# ----------------------------------------------------------------------------------------------------
class _MHACore(Function):
    """softmax((q·dh^-0.5)kᵀ + key padding)·v per head — nn.MultiheadAttention, attention_operator.py:49-55.
    Dense: qkv [N,S,3D] + mask [N,S].  Packed: qkv [T,3D] + cu int32 [N+1] (rows of sequence n = cu[n]..cu[n+1])."""

    @staticmethod
    def forward(ctx, qkv, mask, cu, max_len, heads, drop_p, seed):
        qkv = _f32(qkv)
        mask = _i64(mask) if mask is not None else None
        D = qkv.shape[-1] // 3
        if cu is None:
            N, S = qkv.shape[0], qkv.shape[1]
            rows = N * S
        else:
            N, S, rows = cu.numel() - 1, int(max_len), qkv.shape[0]
        out = torch.empty((*qkv.shape[:-1], D), dtype=torch.float32, device=qkv.device)
        lse = torch.empty((rows, heads), dtype=torch.float32, device=qkv.device)
        call('lk_mha_fwd', ptr(qkv), ptr(mask), ptr(cu), ptr(out), None, None, ptr(lse), N, S, D, heads, float(drop_p), int(seed))
        ctx.save_for_backward(qkv, mask, cu, lse, out)
        ctx.dims = (N, S, D, heads, drop_p, seed)
        return out

    @staticmethod
    def backward(ctx, dctx):
        qkv, mask, cu, lse, out = ctx.saved_tensors
        N, S, D, heads, drop_p, seed = ctx.dims
        dqkv = torch.empty_like(qkv)
        call('lk_mha_bwd', ptr(qkv), ptr(mask), ptr(cu), ptr(out), ptr(lse), ptr(_f32(dctx)), ptr(dqkv), None, None, None, N, S, D, heads,
             float(drop_p), int(seed))
        return dqkv, None, None, None, None, None, None


def mha_core(qkv, mask, heads, drop_p=0.0, seed=0, cu=None, max_len=None):
    return _MHACore.apply(qkv, mask, cu, max_len, heads, drop_p, seed)


# ----------------------------------------------------------------------------------------------------
# additive attention
# ----------------------------------------------------------------------------------------------------
class _AdditiveAttention(Function):
    """model/common/attention.py:23-38: W1 GEMM + tanh, then the fused score/exp/mask/normalise/pool kernel.
    Dense: x [N,S,D] + mask [N,S] (or None).  Packed: x [T,D] + cu int32 [N+1]."""

    @staticmethod
    def forward(ctx, x, mask, cu, max_len, w1, b1, w2):
        x, w1, b1, w2 = _f32(x), _f32(w1), _f32(b1), _f32(w2)
        mask = _i64(mask) if mask is not None else None
        D, A = x.shape[-1], w1.shape[0]
        if cu is None:
            N, S = x.shape[0], x.shape[1]
        else:
            N, S = cu.numel() - 1, int(max_len)
        x2 = x.reshape(-1, D)
        rows = x2.shape[0]
        hid, xp = linear_fwd_raw(x2, w1, b1, None, ACT_TANH)
        ctx.xp = xp
        out = torch.empty((N, D), dtype=torch.float32, device=x.device)
        alpha = torch.empty((rows,), dtype=torch.float32, device=x.device)
        call('lk_additive_pool_fwd', ptr(x2), ptr(hid), ptr(w2), ptr(mask), ptr(cu), ptr(out), ptr(alpha), N, S, D, A)
        ctx.save_for_backward(x2, hid, alpha, w1, w2, cu)
        ctx.shape = (N, S, D, A, tuple(x.shape))
        return out

    @staticmethod
    def backward(ctx, dout):
        x2, hid, alpha, w1, w2, cu = ctx.saved_tensors
        N, S, D, A, xshape = ctx.shape
        dev = x2.device
        rows = x2.shape[0]
        dx = torch.empty((rows, D), dtype=torch.float32, device=dev)
        dpre = torch.empty((rows, A), dtype=torch.float32, device=dev)
        dw2p = torch.empty((N, A), dtype=torch.float32, device=dev)
        call('lk_additive_pool_bwd', ptr(x2), ptr(hid), ptr(w2), ptr(alpha), ptr(cu), ptr(_f32(dout)), ptr(dx), ptr(dpre), ptr(dw2p),
             N, S, D, A, 0)
        dw2 = colsum_raw(dw2p).view(1, A)
        dpp = db1 = None
        if tc_ok(rows, D, A) or tc_ok(rows, A, D):
            dpp, db1 = split_planes(dpre, colsum=True)
        dw1, db1 = linear_bwd_weight_raw(dpre, x2, dyp=dpp, xp=ctx.xp, db=db1)
        linear_bwd_data_raw(dpre, w1, out=dx, accumulate=True, dyp=dpp)
        ctx.xp = None
        return dx.view(xshape), None, None, None, dw1, db1, dw2


def additive_attention(x, mask, w1, b1, w2, cu=None, max_len=None):
    return _AdditiveAttention.apply(x, mask, cu, max_len, w1, b1, w2)


# ----------------------------------------------------------------------------------------------------
# LayerNorm (+ residual), GELU, dropout — the row-wise pieces of the BERT-style blocks (Transformer / Fastformer operators)
# ----------------------------------------------------------------------------------------------------
class _LayerNorm(Function):
    """y = LayerNorm(x + res) * w + b over the last axis (BertSelfOutput / BertOutput / BertEmbeddings, eps 1e-12)."""

    @staticmethod
    def forward(ctx, x, res, w, b, eps):
        x, w, b = _f32(x), _f32(w), _f32(b)
        D = x.shape[-1]
        x2 = x.reshape(-1, D)
        r2 = _f32(res).reshape(-1, D) if res is not None else None
        rows = x2.shape[0]
        dev = x.device
        y = torch.empty_like(x2)
        xs = torch.empty_like(x2) if r2 is not None else None
        mean = torch.empty(rows, dtype=torch.float32, device=dev)
        rstd = torch.empty(rows, dtype=torch.float32, device=dev)
        call('lk_layernorm_fwd', ptr(x2), ptr(r2), ptr(w), ptr(b), ptr(y), ptr(xs), ptr(mean), ptr(rstd), rows, D, float(eps))
        ctx.save_for_backward(xs if xs is not None else x2, w, mean, rstd)
        ctx.has_res, ctx.shape = res is not None, tuple(x.shape)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        xs, w, mean, rstd = ctx.saved_tensors
        rows, D = xs.shape
        dy2 = _f32(dy).reshape(rows, D)
        dx = torch.empty_like(xs)
        nparts = query('lk_layernorm_bwd_parts', rows)
        parts = torch.empty((nparts, 2 * D), dtype=torch.float32, device=xs.device)
        call('lk_layernorm_bwd', ptr(dy2), ptr(xs), ptr(w), ptr(mean), ptr(rstd), ptr(dx), ptr(parts), rows, D)
        dwb = colsum_raw(parts)
        dxv = dx.view(ctx.shape)
        return dxv, (dxv if ctx.has_res else None), dwb[:D].clone(), dwb[D:].clone(), None


def layernorm(x, w, b, eps=1e-12, res=None):
    return _LayerNorm.apply(x, res, w, b, eps)


class _Gelu(Function):
    @staticmethod
    def forward(ctx, x):
        x = _f32(x)
        y = torch.empty_like(x)
        call('lk_gelu', ptr(x), None, ptr(y), x.numel(), 0)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, = ctx.saved_tensors
        dx = torch.empty_like(x)
        call('lk_gelu', ptr(x), ptr(_f32(dy)), ptr(dx), x.numel(), 1)
        return dx


def gelu(x):
    """Exact (erf) GELU — transformers' ACT2FN['gelu'] (BertIntermediate) and F.gelu (miner_predictor.py:36)."""
    return _Gelu.apply(x)


class _Dropout(Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        x = _f32(x)
        y = torch.empty_like(x)
        call('lk_dropout', ptr(x), ptr(y), x.numel(), float(p), int(seed))
        ctx.ps = (float(p), int(seed))
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _f32(dy)
        dx = torch.empty_like(dy)
        call('lk_dropout', ptr(dy), ptr(dx), dy.numel(), ctx.ps[0], ctx.ps[1])
        return dx, None, None


def dropout(x, p, seed):
    """nn.Dropout(p) with the library's counter-based stream; identity for p == 0."""
    return x if p <= 0.0 else _Dropout.apply(x, p, seed)


# ----------------------------------------------------------------------------------------------------
# MINER: poly-attention pooling and the target-aware predictor
# ----------------------------------------------------------------------------------------------------
class _PolyPool(Function):
    """out[b, c, :] = sum_s softmax_s(mask ? logits[b, s, c] : 1e-30) x[b, s, :]  (poly_attention_operator.py:52-56)."""

    @staticmethod
    def forward(ctx, logits, mask, x):
        logits, x, mask = _f32(logits), _f32(x), _i64(mask)
        B, S, C = logits.shape
        D = x.shape[-1]
        out = torch.empty((B, C, D), dtype=torch.float32, device=x.device)
        w = torch.empty((B, C, S), dtype=torch.float32, device=x.device)
        call('lk_poly_pool_fwd', ptr(logits), ptr(mask), ptr(x), ptr(out), ptr(w), B, S, C, D)
        ctx.save_for_backward(w, mask, x)
        return out

    @staticmethod
    def backward(ctx, dout):
        w, mask, x = ctx.saved_tensors
        B, C, S = w.shape
        D = x.shape[-1]
        dx = torch.empty_like(x)
        dlogits = torch.empty((B, S, C), dtype=torch.float32, device=x.device)
        call('lk_poly_pool_bwd', ptr(_f32(dout)), ptr(w), ptr(mask), ptr(x), ptr(dx), ptr(dlogits), B, S, C, D)
        return dlogits, None, dx


def poly_pool(logits, mask, x):
    return _PolyPool.apply(logits, mask, x)


MINER_MODES = {'weighted': 0, 'max': 1, 'mean': 2}


class _MinerScore(Function):
    """miner_predictor.py:50-62: scores = items·userᵀ; weighted: sum_c softmax_c(items·projᵀ) * scores | max_c | mean_c."""

    @staticmethod
    def forward(ctx, user, proj, items, mode):
        user, items = _f32(user), _f32(items)
        proj = _f32(proj) if proj is not None else None
        B, C, D = user.shape
        K1 = items.shape[1]
        dev = user.device
        out = torch.empty((B, K1), dtype=torch.float32, device=dev)
        sc = torch.empty((B, K1, C), dtype=torch.float32, device=dev)
        wt = torch.empty((B, K1, C), dtype=torch.float32, device=dev)
        call('lk_miner_fwd', ptr(user), ptr(proj), ptr(items), ptr(out), ptr(sc), ptr(wt), B, K1, C, D, mode)
        ctx.save_for_backward(user, proj if proj is not None else user, items, sc, wt)
        ctx.mode, ctx.has_proj = mode, proj is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        user, proj, items, sc, wt = ctx.saved_tensors
        B, C, D = user.shape
        K1 = items.shape[1]
        du, dp, di = torch.empty_like(user), torch.empty_like(user), torch.empty_like(items)
        call('lk_miner_bwd', ptr(_f32(dout)), ptr(user), ptr(proj), ptr(items), ptr(sc), ptr(wt), ptr(du), ptr(dp), ptr(di), B, K1, C, D, ctx.mode)
        return du, (dp if ctx.has_proj else None), di, None


def miner_score(user, proj, items, mode):
    return _MinerScore.apply(user, proj, items, MINER_MODES[mode] if isinstance(mode, str) else mode)


# ----------------------------------------------------------------------------------------------------
# Fastformer: per-head softmax pooling, broadcast product, add
# ----------------------------------------------------------------------------------------------------
class _HeadPool(Function):
    """out[b, h*dh + j] = sum_s softmax_s(score[b, s, h] * scale - 10000 * (1 - mask[b, s])) v[b, s, h*dh + j]  (fastformer.py:103-116, 125-132)."""

    @staticmethod
    def forward(ctx, score, mask, v, scale):
        score, v, mask = _f32(score), _f32(v), _i64(mask)
        B, S, H = score.shape
        D = v.shape[-1]
        out = torch.empty((B, D), dtype=torch.float32, device=v.device)
        w = torch.empty((B, H, S), dtype=torch.float32, device=v.device)
        call('lk_head_pool_fwd', ptr(score), ptr(mask), ptr(v), ptr(out), ptr(w), B, S, H, D, float(scale))
        ctx.save_for_backward(w, v)
        ctx.scale = float(scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        w, v = ctx.saved_tensors
        B, H, S = w.shape
        D = v.shape[-1]
        dv = torch.empty_like(v)
        dscore = torch.empty((B, S, H), dtype=torch.float32, device=v.device)
        call('lk_head_pool_bwd', ptr(_f32(dout)), ptr(w), ptr(v), ptr(dv), ptr(dscore), B, S, H, D, ctx.scale)
        return dscore, None, dv, None


def head_pool(score, mask, v, scale):
    return _HeadPool.apply(score, mask, v, scale)


class _BcastMul(Function):
    """y[b, s, :] = a[b, s, :] * v[b, :]"""

    @staticmethod
    def forward(ctx, a, v):
        a, v = _f32(a), _f32(v)
        B, S, D = a.shape
        y = torch.empty_like(a)
        call('lk_bcast_mul', ptr(a), ptr(v), ptr(y), B, S, D)
        ctx.save_for_backward(a, v)
        return y

    @staticmethod
    def backward(ctx, dy):
        a, v = ctx.saved_tensors
        B, S, D = a.shape
        dy = _f32(dy)
        da = torch.empty_like(a)
        dv = torch.empty_like(v)
        call('lk_bcast_mul', ptr(dy), ptr(v), ptr(da), B, S, D)
        call('lk_bcast_mul_dv', ptr(dy), ptr(a), ptr(dv), B, S, D)
        return da, dv


def bcast_mul(a, v):
    return _BcastMul.apply(a, v)


class _Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _f32(a), _f32(b)
        y = torch.empty_like(a)
        call('lk_add', ptr(a), ptr(b), ptr(y), a.numel())
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def add(a, b):
    return _Add.apply(a, b)


# ----------------------------------------------------------------------------------------------------
# GRU over padded sequences -> last hidden state (LSTUR user encoder)
# ----------------------------------------------------------------------------------------------------
class _GRULast(Function):
    """nn.GRU(1 layer, batch_first) over pack_padded_sequence(x, lengths) -> h at each sequence's last valid step
    (model/operators/gru_operator.py:40-52).  Input projection and all weight / input gradients are contractions (tensor cores when the
    shapes allow); the recurrence itself is lk_gru_fwd / lk_gru_bwd."""

    @staticmethod
    def forward(ctx, x, lengths, w_ih, w_hh, b_ih, b_hh):
        x, w_ih, w_hh, b_ih, b_hh = _f32(x), _f32(w_ih), _f32(w_hh), _f32(b_ih), _f32(b_hh)
        B, S, I = x.shape
        H = w_hh.shape[1]
        dev = x.device
        x2 = x.reshape(B * S, I)
        gi, xp = linear_fwd_raw(x2, w_ih, b_ih, None, ACT_NONE)
        lens = lengths.to(dev, torch.int32).contiguous()
        whhT = w_hh.t().contiguous()                      # layout only: [H, 3H] so that a warp reads consecutive recurrent weights
        last = torch.empty((B, H), dtype=torch.float32, device=dev)
        hs = torch.zeros((B, S, H), dtype=torch.float32, device=dev)        # rows beyond a sequence's length stay zero
        gates = torch.zeros((B, S, 3 * H), dtype=torch.float32, device=dev)
        hnp = torch.zeros((B, S, H), dtype=torch.float32, device=dev)
        call('lk_gru_fwd', ptr(gi), ptr(whhT), ptr(b_hh), ptr(lens), ptr(last), ptr(hs), ptr(gates), ptr(hnp), B, S, H)
        ctx.save_for_backward(x2, lens, w_ih, w_hh, hs, gates, hnp)
        ctx.xp, ctx.dims = xp, (B, S, I, H)
        return last

    @staticmethod
    def backward(ctx, dlast):
        x2, lens, w_ih, w_hh, hs, gates, hnp = ctx.saved_tensors
        B, S, I, H = ctx.dims
        dev = x2.device
        dgi = torch.empty((B * S, 3 * H), dtype=torch.float32, device=dev)
        dgh = torch.empty((B * S, 3 * H), dtype=torch.float32, device=dev)
        call('lk_gru_bwd', ptr(_f32(dlast)), ptr(w_hh), ptr(lens), ptr(hs), ptr(gates), ptr(hnp), ptr(dgi), ptr(dgh), B, S, H)
        dw_ih, db_ih = linear_bwd_weight_raw(dgi, x2, xp=ctx.xp)
        hprev = torch.cat([torch.zeros((B, 1, H), dtype=torch.float32, device=dev), hs[:, :-1]], dim=1).reshape(B * S, H)   # h_{t-1}: a shift
        dw_hh, db_hh = linear_bwd_weight_raw(dgh, hprev)
        dx = linear_bwd_data_raw(dgi, w_ih)
        ctx.xp = None
        return dx.view(B, S, I), None, dw_ih, dw_hh, db_ih, db_hh


def gru_last_hidden(x, lengths, w_ih, w_hh, b_ih, b_hh):
    return _GRULast.apply(x, lengths, w_ih, w_hh, b_ih, b_hh)


# ----------------------------------------------------------------------------------------------------
# Conv1d('same') + ReLU + mask (NAML)
# ----------------------------------------------------------------------------------------------------
class _Conv1dReluMask(Function):
    """dropout(relu(conv1d(x, W, b, 'same')) * mask) over [N,S,C] — cnn_operator.py:54-58."""

    @staticmethod
    def forward(ctx, x, w, b, mask, drop_p, seed):
        x, w, b = _f32(x), _f32(w), _f32(b)
        N, S, Cin = x.shape
        Cout, _, taps = w.shape
        wr = w.permute(0, 2, 1).contiguous().view(Cout, taps * Cin)       # Wr[o, j*Cin+i] = W[o,i,j] (layout only)
        rm = _i64(mask.reshape(-1)) if mask is not None else None
        x2 = x.reshape(N * S, Cin).contiguous()
        # Backward (two thirds of the work) always runs as tcgen05 contractions over im2col planes when the shape allows.  The forward does
        # so only on request (CONV_TC_FWD): ReLU's gate is a step function of the pre-activation, and the ~1e-5 relative error of the
        # split-bf16 product flips it for a few hundred of the ~1e8 elements of a NAML batch (|pre| < 1e-5) — each flip is an O(1) error
        # in one dPre element, which shows up at 5e-3 in small bias gradients (measured on naml_full).  The fp32 FFMA forward flips
        # 100x fewer gates, like any fp32 implementation of the reference does against another.
        ctx.tc = tc_ok(N * S, Cout, taps * Cin) and Cin % 4 == 0 and Cout % 4 == 0
        if ctx.tc and CONV_TC_FWD:
            y = tc_gemm(im2col_planes(x2, S, taps), split_planes(wr), False, N * S, Cout, taps * Cin, bias=b, rowmask=rm, act=ACT_RELU,
                        drop_p=drop_p, seed=seed)
        else:
            y = torch.empty((N * S, Cout), dtype=torch.float32, device=x.device)
            call('lk_conv1d_fwd', ptr(x2), ptr(wr), ptr(b), ptr(rm), ptr(y), N * S, S, Cin, Cout, taps, ACT_RELU, float(drop_p), int(seed))
        ctx.save_for_backward(x2, w, y, rm)
        ctx.dims = (N, S, Cin, Cout, taps)
        ctx.drop_p, ctx.seed = drop_p, seed
        return y.view(N, S, Cout)

    @staticmethod
    def backward(ctx, dy):
        x2, w, y, rm = ctx.saved_tensors
        N, S, Cin, Cout, taps = ctx.dims
        dy2 = act_bwd_raw(_f32(dy).reshape(N * S, Cout), y, rm, ACT_RELU, ctx.drop_p, ctx.seed)
        dx = dw = db = None
        if ctx.tc:
            if ctx.needs_input_grad[0]:
                wd = w.flip(2).permute(1, 2, 0).contiguous().view(Cin, taps * Cout)   # Wd[i, j*Cout+o] = W[o,i,taps-1-j]
                dx = tc_gemm(im2col_planes(dy2, S, taps), split_planes(wd), False, N * S, Cin, taps * Cout).view(N, S, Cin)
            if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
                dyp, db = split_planes(dy2, colsum=True)
                dwr = tc_gemm(dyp, im2col_planes(x2, S, taps), True, Cout, taps * Cin, N * S)
                dw = dwr.view(Cout, taps, Cin).permute(0, 2, 1).contiguous()
            return dx, dw, db, None, None, None
        if ctx.needs_input_grad[0]:
            wd = w.flip(2).permute(1, 2, 0).contiguous().view(Cin, taps * Cout)   # Wd[i, j*Cout+o] = W[o,i,taps-1-j]
            dx = torch.empty((N * S, Cin), dtype=torch.float32, device=dy.device)
            call('lk_conv1d_bwd_data', ptr(dy2), ptr(wd), ptr(dx), N * S, S, Cin, Cout, taps, 0)
            dx = dx.view(N, S, Cin)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dwr = torch.empty((Cout, taps * Cin), dtype=torch.float32, device=dy.device)
            db = torch.empty((Cout,), dtype=torch.float32, device=dy.device)
            nbytes = query('lk_conv1d_bwd_weight_workspace_bytes', N * S, Cin, Cout, taps)
            ws = workspace(nbytes, dy.device, 'wgrad')
            call('lk_conv1d_bwd_weight', ptr(dy2), ptr(x2), ptr(dwr), ptr(db), N * S, S, Cin, Cout, taps, 0, ptr(ws), ws.numel())
            dw = dwr.view(Cout, taps, Cin).permute(0, 2, 1).contiguous()
        return dx, dw, db, None, None, None


def conv1d_relu_mask(x, w, b, mask, drop_p=0.0, seed=0):
    return _Conv1dReluMask.apply(x, w, b, mask, drop_p, seed)


# ----------------------------------------------------------------------------------------------------
# masked pooling of gathered embeddings
# ----------------------------------------------------------------------------------------------------
class _MaskedPool(Function):
    """pooling_operator.py:46-56 on an [N,S,D] tensor."""

    @staticmethod
    def forward(ctx, x, mask, mode):
        x, mask = _f32(x), _i64(mask)
        N, S, D = x.shape
        out = torch.empty((N, D), dtype=torch.float32, device=x.device)
        call('lk_masked_pool', ptr(x), ptr(mask), ptr(out), N, S, D, mode)
        ctx.save_for_backward(mask)
        ctx.dims, ctx.mode = (N, S, D), mode
        return out

    @staticmethod
    def backward(ctx, dout):
        (mask,) = ctx.saved_tensors
        if ctx.mode != POOL_MEAN:
            raise RuntimeError('masked_pool: backward of max pooling is not implemented')
        N, S, D = ctx.dims
        dx = torch.empty((N, S, D), dtype=torch.float32, device=dout.device)
        call('lk_masked_mean_pool_bwd', ptr(_f32(dout)), ptr(mask), ptr(dx), N, S, D)
        return dx, None, None


def masked_pool(x, mask, mode=POOL_MEAN):
    return _MaskedPool.apply(x, mask, mode)


# ----------------------------------------------------------------------------------------------------
# scoring + loss
# ----------------------------------------------------------------------------------------------------
class _DotScores(Function):
    """scores[b,c] = <u[b], v[b,c]> — legommender.py:279-283 + dot_predictor.py:10 (no materialised repeat)."""

    @staticmethod
    def forward(ctx, user, items):
        user, items = _f32(user), _f32(items)
        B, C, D = items.shape
        scores = torch.empty((B, C), dtype=torch.float32, device=user.device)
        call('lk_dot_scores', ptr(user), ptr(items), ptr(scores), B, C, D)
        ctx.save_for_backward(user, items)
        return scores

    @staticmethod
    def backward(ctx, ds):
        user, items = ctx.saved_tensors
        B, C, D = items.shape
        du, dv = torch.empty_like(user), torch.empty_like(items)
        call('lk_dot_bwd', ptr(user), ptr(items), ptr(_f32(ds)), ptr(du), ptr(dv), B, C, D)
        return du, dv


def dot_scores(user, items):
    return _DotScores.apply(user, items)


class _DotCE(Function):
    """loss = CrossEntropy(dot scores, label 0), mean — legommender.py:252-254, 263 fused with the predictor."""

    @staticmethod
    def forward(ctx, user, items):
        user, items = _f32(user), _f32(items)
        B, C, D = items.shape
        dev = user.device
        scores = torch.empty((B, C), dtype=torch.float32, device=dev)
        probs = torch.empty((B, C), dtype=torch.float32, device=dev)
        rowloss = torch.empty((B,), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        call('lk_dot_ce_fwd', ptr(user), ptr(items), ptr(scores), ptr(probs), ptr(rowloss), ptr(loss), B, C, D)
        ctx.save_for_backward(user, items, probs)
        ctx.mark_non_differentiable(scores)
        return loss, scores

    @staticmethod
    def backward(ctx, dloss, _dscores):
        user, items, probs = ctx.saved_tensors
        B, C, D = items.shape
        du, dv = torch.empty_like(user), torch.empty_like(items)
        call('lk_dot_ce_bwd', ptr(user), ptr(items), ptr(probs), ptr(_f32(dloss)), ptr(du), ptr(dv), B, C, D)
        return du, dv


def dot_ce_loss(user, items):
    """returns (loss, scores)"""
    return _DotCE.apply(user, items)


class _DotBCE(Function):
    """loss = BCEWithLogits(<u,v>, click), mean — legommender.py:256-257, 285-290."""

    @staticmethod
    def forward(ctx, user, item, label):
        user, item, label = _f32(user), _f32(item), _f32(label)
        B, D = user.shape
        dev = user.device
        scores = torch.empty((B,), dtype=torch.float32, device=dev)
        rowloss = torch.empty((B,), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        call('lk_dot_bce_fwd', ptr(user), ptr(item), ptr(label), ptr(scores), ptr(rowloss), ptr(loss), B, D)
        ctx.save_for_backward(user, item, label, scores)
        ctx.mark_non_differentiable(scores)
        return loss, scores

    @staticmethod
    def backward(ctx, dloss, _ds):
        user, item, label, scores = ctx.saved_tensors
        B, D = user.shape
        dz = torch.empty_like(scores)
        du, dv = torch.empty_like(user), torch.empty_like(item)
        call('lk_dot_bce_bwd', ptr(user), ptr(item), ptr(label), ptr(scores), ptr(_f32(dloss)), ptr(dz), ptr(du), ptr(dv), B, D)
        return du, dv, None


def dot_bce_loss(user, item, label):
    return _DotBCE.apply(user, item, label)


# ----------------------------------------------------------------------------------------------------
# cached evaluation (no autograd: caches are detached — fast_item_pager.py:143, fast_user_pager.py:133)
# ----------------------------------------------------------------------------------------------------
def cached_scores(user_repr, item_repr, user_ids, item_ids, out=None):
    """score[r] = <U[uid[r]], I[iid[r]]> — legommender.py:153-157, 202-203 + dot."""
    user_repr, item_repr = _f32(user_repr), _f32(item_repr)
    user_ids, item_ids = _i64(user_ids.reshape(-1)), _i64(item_ids.reshape(-1))
    R, D = user_ids.numel(), user_repr.shape[1]
    if out is None:
        out = torch.empty((R,), dtype=torch.float32, device=user_repr.device)
    id_violation_counter(user_repr.device)
    call('lk_cached_scores', ptr(user_repr), user_repr.shape[0], ptr(item_repr), item_repr.shape[0], ptr(user_ids), ptr(item_ids), ptr(out), R, D)
    return out


def index_rows(table, ids):
    """table[ids] for cache indexing — legommender.py:153-157."""
    table, flat = _f32(table), _i64(ids.reshape(-1))
    out = torch.empty((flat.numel(), table.shape[1]), dtype=torch.float32, device=table.device)
    id_violation_counter(table.device)
    call('lk_index_rows', ptr(table), table.shape[0], ptr(flat), ptr(out), flat.numel(), table.shape[1])
    return out.view(*ids.shape, table.shape[1])


def adam_step(p, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """torch.optim.Adam defaults (base_lego.py:198-204) over one flat fp32 buffer."""
    for t in (p, g, m, v):
        if t.data_ptr() % 16:
            raise RuntimeError('adam_step: buffers must be 16-byte aligned')
    call('lk_adam_step', ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps), int(step),
         float(grad_scale))
