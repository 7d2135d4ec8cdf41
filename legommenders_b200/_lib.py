"""ctypes binding of the C-ABI in include/legommenders_b200.h.

The product path has NO fallback: if the shared library is missing (or a call fails) a RuntimeError is
raised.  Tensors cross the boundary as raw device pointers + sizes; torch is only the owner of the memory
and of the current stream.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'liblegommenders_b200.so')

_P, _Q, _I, _F, _U, _Z = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_uint64, ctypes.c_size_t
_T = {'p': _P, 'q': _Q, 'i': _I, 'f': _F, 'u': _U, 'z': _Z, 's': _P}

# name -> (argument codes, restype code).  p pointer, q int64, i int, f float, u uint64, z size_t, s stream
SIGNATURES = {
    'lk_version': ('', 'c'),
    'lk_last_error': ('', 'c'),
    'lk_device_ok': ('', 'i'),
    'lk_launch_count': ('', 'u'),
    'lk_profile_enable': ('i', 'v'),
    'lk_profile_collect': ('pipppi', 'i'),
    'lk_set_id_violation_counter': ('p', 'v'),
    'lk_gather_rows': ('pppqpqqis', 'i'),
    'lk_gather_pool': ('pppqpqqqis', 'i'),
    'lk_gather_split_bf16': ('ppqppqqqs', 'i'),
    'lk_scatter_add_workspace_bytes': ('qqq', 'z'),
    'lk_scatter_add_sorted': ('ppppqpqqqipzs', 'i'),
    'lk_pack_item_tokens': ('ppippqqs', 'i'),
    'lk_concat_embed_bwd_blocks': ('q', 'q'),
    'lk_concat_embed_bwd_workspace_bytes': ('qqqq', 'z'),
    'lk_concat_embed_bwd': ('ppppqqqqfuppqppppzs', 'i'),
    'lk_linear_fwd': ('pppppqqqiifus', 'i'),
    'lk_act_bwd': ('ppppqqifus', 'i'),
    'lk_valid_mask': ('ppqs', 'i'),
    'lk_linear_bwd_data': ('pppqqqis', 'i'),
    'lk_linear_bwd_weight_workspace_bytes': ('qqq', 'z'),
    'lk_linear_bwd_weight': ('ppppqqqipzs', 'i'),
    'lk_splitk_reduce': ('ppqqqiis', 'i'),
    'lk_split_bf16_workspace_bytes': ('qq', 'z'),
    'lk_split_bf16': ('pqqqppqippzs', 'i'),
    'lk_im2col_split_bf16': ('pqqqippqs', 'i'),
    'lk_split_bf16_multi': ('pis', 'i'),
    'lk_split_bf16_partial': ('pqqqppqps', 'i'),
    'lk_colsum_finish_multi': ('pis', 'i'),
    'lk_tc_gemm_workspace_bytes': ('qqq', 'z'),
    'lk_tc_gemm_ex': ('ppqippqipqqqqppzs', 'i'),
    'lk_tc_gemm': ('ppqippqipqqqqppifuipzs', 'i'),
    'lk_colsum_workspace_bytes': ('qq', 'z'),
    'lk_colsum': ('ppqqipzs', 'i'),
    'lk_conv1d_fwd': ('pppppqqqqiifus', 'i'),
    'lk_conv1d_bwd_data': ('pppqqqqiis', 'i'),
    'lk_conv1d_bwd_weight_workspace_bytes': ('qqqi', 'z'),
    'lk_conv1d_bwd_weight': ('ppppqqqqiipzs', 'i'),
    'lk_mha_fwd': ('pppppppqqqqfus', 'i'),
    'lk_mha_bwd': ('ppppppppppqqqqfus', 'i'),
    'lk_additive_pool_fwd': ('pppppppqqqqs', 'i'),
    'lk_additive_pool_fwd_planes': ('ppqpppp' + 'qqqs', 'i'),
    'lk_additive_pool_bwd_planes': ('ppqppppppp' + 'ppqp' + 'p' + 'qqqqs', 'i'),
    'lk_additive_pool_bwd': ('pppppppppqqqqis', 'i'),
    'lk_masked_pool': ('pppqqqis', 'i'),
    'lk_masked_mean_pool_bwd': ('pppqqqs', 'i'),
    'lk_dot_scores': ('pppqqqs', 'i'),
    'lk_dot_ce_fwd': ('ppppppqqqs', 'i'),
    'lk_dot_ce_bwd': ('ppppppqqqs', 'i'),
    'lk_dot_bwd': ('pppppqqqs', 'i'),
    'lk_dot_bce_fwd': ('ppppppqqs', 'i'),
    'lk_dot_bce_bwd': ('ppppppppqqs', 'i'),
    'lk_cached_scores': ('pqpqpppqqs', 'i'),
    'lk_index_rows': ('pqppqqs', 'i'),
    'lk_group_metrics_workspace_bytes': ('qi', 'z'),
    'lk_group_metrics': ('pppqpipqpppzs', 'i'),
    'lk_adam_step': ('ppppqffffqfs', 'i'),
    'lk_fill_f32': ('pfqs', 'i'),
    'lk_shard_hash_slots': ('q', 'q'),
    'lk_shard_plan': ('pqiqppqppps', 'i'),
    'lk_shard_inverse': ('pqppqps', 'i'),
    'lk_shard_gather': ('pqipqqps', 'i'),
    'lk_allreduce_set_trace': ('p', 'i'),
    'lk_allreduce_p2p': ('pppiiiqfs', 'i'),
    'lk_resample_batch': ('pqiuppppppp' + 'qqq' + 'ppppp' + 'qs', 'i'),
    'lk_resample_reference': ('uqqpqiqp', 'i'),
    'lk_sweep_ranges': ('qq', 'q'),
    'lk_sweep_topk': ('ppqqppqqqiqpps', 'i'),
    'lk_layernorm_fwd': ('pppppppp' + 'qqfs', 'i'),
    'lk_layernorm_bwd_parts': ('q', 'q'),
    'lk_layernorm_bwd': ('ppppppp' + 'qqs', 'i'),
    'lk_gelu': ('pppqis', 'i'),
    'lk_dropout': ('ppqfus', 'i'),
    'lk_poly_pool_fwd': ('ppppp' + 'qqqqs', 'i'),
    'lk_poly_pool_bwd': ('pppppp' + 'qqqqs', 'i'),
    'lk_miner_fwd': ('pppppp' + 'qqqqis', 'i'),
    'lk_miner_bwd': ('ppppppppp' + 'qqqqis', 'i'),
    'lk_head_pool_fwd': ('ppppp' + 'qqqqfs', 'i'),
    'lk_head_pool_bwd': ('ppppp' + 'qqqqfs', 'i'),
    'lk_bcast_mul': ('pppqqqs', 'i'),
    'lk_bcast_mul_dv': ('pppqqqs', 'i'),
    'lk_add': ('pppqs', 'i'),
    'lk_gru_fwd': ('pppppppp' + 'qqqs', 'i'),
    'lk_gru_bwd': ('pppppppp' + 'qqqs', 'i'),
    'lk_tc_chain': ('ppqqpiis', 'i'),
    'lk_tc_chain_trace': ('pi', 'i'),
    'lk_nrms_arena_bytes': ('qqqqqqqqqq', 'z'),
    'lk_nrms_fwd_bwd': ('ppppqqqpqqqpqppp' + 'qqqqqq' + 'ffu' + 'pppzs', 'i'),
}

_lib = None
_lock = threading.Lock()


def load() -> ctypes.CDLL:
    """Load the CUDA library (once).  Raises if it has not been built — there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a).  legommenders_b200 has no CPU / eager fallback.')
        lib = ctypes.CDLL(LIB_PATH)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the header and the library disagree
            fn.argtypes = [_T[a] for a in args]
            fn.restype = ctypes.c_char_p if res == 'c' else (None if res == 'v' else _T[res])
        _lib = lib
    return _lib


def ptr(t):
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


_profile = None   # when a list: (name, flops, start_event, stop_event) per C-ABI call (bench.py's per-kernel timing)

# algorithmic flops of the dense-contraction entry points, from their (M, N, K) arguments
_FLOPS = {
    'lk_linear_fwd': lambda a: 2 * a[5] * a[6] * a[7],
    'lk_linear_bwd_data': lambda a: 2 * a[3] * a[4] * a[5],
    'lk_linear_bwd_weight': lambda a: 2 * a[4] * a[5] * a[6],
    'lk_tc_gemm': lambda a: 2 * a[10] * a[11] * a[12],
    'lk_conv1d_fwd': lambda a: 2 * a[5] * a[7] * a[8] * a[9],
    'lk_conv1d_bwd_data': lambda a: 2 * a[3] * a[5] * a[6] * a[7],
    'lk_conv1d_bwd_weight': lambda a: 2 * a[4] * a[6] * a[7] * a[8],
}


def call(name: str, *args):
    """Invoke an int-returning entry point on the current stream; raise on a non-zero status."""
    lib = load()
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream())
        e1.record()
        f = _FLOPS.get(name)
        _profile.append((name, f(args) if f else 0, e0, e1))
    else:
        rc = getattr(lib, name)(*args, stream())
    if rc != 0:
        raise RuntimeError(f'{name} failed ({rc}): {lib.lk_last_error().decode()}')
    if _CHECK_NOW and _viol:
        raise_on_bad_ids(name)


def profile_begin():
    global _profile
    _profile = []
    return _profile


def profile_end(records):
    """-> ({entry point: total ms}, {'ms','flops','calls'} of the dense-contraction entry points)."""
    global _profile
    _profile = None
    torch.cuda.synchronize()
    shares, per = {}, {}
    for name, flops, e0, e1 in records:
        ms = e0.elapsed_time(e1)
        shares[name] = shares.get(name, 0.0) + ms
        if flops:
            g = per.setdefault(name, dict(ms=0.0, flops=0, calls=0, name=name))
            g['ms'] += ms
            g['flops'] += flops
            g['calls'] += 1
    gemm = max(per.values(), key=lambda g: g['ms']) if per else dict(ms=0.0, flops=0, calls=0, name=None)
    return {k: round(v, 4) for k, v in sorted(shares.items(), key=lambda kv: -kv[1])}, gemm


def native_profile_begin():
    """Start per-sub-call CUDA-event timing inside the native step drivers (lk_profile_enable)."""
    load().lk_profile_enable(1)


def native_profile_end():
    """-> ({sub-call: total ms}, {'ms','flops','calls','name'} of the slowest dense-contraction entry point)."""
    lib = load()
    cap, stride = 64, 40
    names = ctypes.create_string_buffer(cap * stride)
    ms = (ctypes.c_float * cap)()
    flops = (ctypes.c_double * cap)()
    calls = (ctypes.c_int * cap)()
    n = lib.lk_profile_collect(ctypes.cast(names, ctypes.c_void_p), stride, ctypes.cast(ms, ctypes.c_void_p),
                               ctypes.cast(flops, ctypes.c_void_p), ctypes.cast(calls, ctypes.c_void_p), cap)
    lib.lk_profile_enable(0)
    if n < 0:
        raise RuntimeError(f'lk_profile_collect failed: {lib.lk_last_error().decode()}')
    shares, gemms = {}, []
    for i in range(n):
        name = names.raw[i * stride:(i + 1) * stride].split(b'\0', 1)[0].decode()
        shares[name] = round(float(ms[i]), 4)
        if flops[i] > 0:
            gemms.append(dict(ms=float(ms[i]), flops=float(flops[i]), calls=int(calls[i]), name=name))
    # every dense contraction of the driver goes through lk_tc_gemm (labelled per shape): aggregate them
    gemm = dict(ms=sum(g['ms'] for g in gemms), flops=sum(g['flops'] for g in gemms), calls=sum(g['calls'] for g in gemms),
                name='lk_tc_gemm' if gemms else None, per_shape={g['name']: dict(ms=round(g['ms'], 4), calls=g['calls'],
                tflops=round(g['flops'] / max(g['ms'], 1e-9) / 1e9, 1)) for g in gemms})
    return dict(sorted(shares.items(), key=lambda kv: -kv[1])), gemm


_viol = {}


def id_violation_counter(device=None) -> torch.Tensor:
    """The device int32 counter the kernels add out-of-range ids to (one per device, registered with the library on first use)."""
    device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    c = _viol.get(key)
    if c is None:
        c = torch.zeros(1, dtype=torch.int32, device=torch.device('cuda', key))
        _viol[key] = c
    load().lk_set_id_violation_counter(c.data_ptr())
    return c


def raise_on_bad_ids(where: str = ''):
    """Read the counter (a device->host sync: call it where the host synchronises anyway) and raise the reference's IndexError."""
    bad = 0
    for c in _viol.values():
        n = int(c.item())
        if n:
            c.zero_()
            bad += n
    if bad:
        raise IndexError(f'{bad} id(s) out of range in a gather / index / scoring kernel{" (" + where + ")" if where else ""}')


_CHECK_NOW = os.environ.get('LK_CHECK_IDS', '0') == '1'      # debug: synchronise and check after every id-consuming call


def query(name: str, *args) -> int:
    return int(getattr(load(), name)(*args))


_ws = {}


def workspace(nbytes: int, device, tag: str = 'default') -> torch.Tensor:
    """Grow-only scratch buffer per (device, tag); the C ABI never allocates."""
    key = (str(device), tag)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf
