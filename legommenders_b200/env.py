"""Global phase flags of a Legommenders run (mirror of loader/env.py:4-61 — same names, same meaning).

`Resampler` reads `item_cache` / `user_cache` / `lm_cache` (possibly in forked workers), the model reads the
phase flags to pick loss vs scores (model/legommender.py:260-263), and every module reads `device`.
"""
import torch

_BRIDGED = ('device', 'simple_dev', 'is_training', 'is_evaluating', 'is_testing', 'item_cache', 'user_cache', 'lm_cache')


class _EnvMeta(type):
    """When bound to the reference's `loader.env.Env` (integration.bind_reference_env) the phase flags, the device and the cache flags
    read and write THROUGH to it, so a run driven by the reference's trainer and this package's kernels share one global state."""
    _target = None

    def __getattribute__(cls, name):
        if name in _BRIDGED:
            target = type.__getattribute__(cls, '_target')
            if target is not None:
                return getattr(target, name)
        return type.__getattribute__(cls, name)

    def __setattr__(cls, name, value):
        if name in _BRIDGED and type.__getattribute__(cls, '_target') is not None:
            setattr(type.__getattribute__(cls, '_target'), name, value)
            return
        type.__setattr__(cls, name, value)


class Env(metaclass=_EnvMeta):
    device = None          # assigned directly, as base_lego.py:120 does
    simple_dev = False
    UNSET = -1

    is_training = True
    is_evaluating = False
    is_testing = False

    item_cache = False
    user_cache = False
    lm_cache = False

    @classmethod
    def _phase(cls, training, evaluating, testing):
        cls.is_training, cls.is_evaluating, cls.is_testing = training, evaluating, testing

    @classmethod
    def train(cls):
        cls._phase(True, False, False)

    @classmethod
    def dev(cls):
        cls._phase(False, True, False)

    @classmethod
    def test(cls):
        cls._phase(False, False, True)

    @classmethod
    def set_device(cls, device):
        cls._device = device   # reference quirk kept: writes `_device`, not `device` (loader/env.py:47-49)

    @classmethod
    def set_item_cache(cls, flag):
        cls.item_cache = flag

    @classmethod
    def set_user_cache(cls, flag):
        cls.user_cache = flag

    @classmethod
    def set_lm_cache(cls, flag):
        cls.lm_cache = flag

    @classmethod
    def bind(cls, target=None):
        """Delegate the shared flags to another Env class (the reference's); bind(None) restores the local state."""
        type.__setattr__(cls, '_target', target)

    @classmethod
    def use_cuda(cls, index: int = 0):
        """Select cuda:<index>; the B200 path has no CPU mode (the reference's `--cuda -1` is the oracle's job)."""
        if not torch.cuda.is_available():
            raise RuntimeError('legommenders_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        cls.device = torch.device('cuda', index)
        torch.cuda.set_device(cls.device)
        return cls.device
