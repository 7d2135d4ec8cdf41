"""Import stub for `pigmento` (absent here); only used by tests/golden/make_golden.py."""


class _Pnt:
    def __call__(self, *a, **k):
        pass

    def __getattr__(self, item):
        return lambda *a, **k: None


pnt = _Pnt()
