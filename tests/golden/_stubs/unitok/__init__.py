"""Import stub for the third-party `unitok` package (absent here), used ONLY when the live
reference under /root/reference is imported to mint golden vectors (tests/golden/make_golden.py)."""


class Symbol:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return self.name


class Vocab:
    def __init__(self, name, size=0):
        self.name = name
        self._tokens = []
        self._size = size

    def append(self, tok):
        self._tokens.append(tok)
        self._size = max(self._size, len(self._tokens))
        return len(self._tokens) - 1

    @property
    def size(self):
        return self._size

    def __len__(self):
        return self._size


class Feature:
    pass


class UniTok:
    pass


class Meta:
    pass


class State:
    pass
