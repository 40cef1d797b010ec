"""Minimal eager type_check: expressions evaluate to plain Python values."""


class InvalidType(Exception):
    pass


class _Dtype(object):
    def __init__(self, dt):
        self._dt = dt
        self.char = dt.char
        self.kind = dt.kind

    def __eq__(self, o):
        return self._dt == o


class _TypeInfo(object):
    def __init__(self, arr):
        self.shape = tuple(arr.shape)
        self.ndim = arr.ndim
        self.dtype = _Dtype(arr.dtype)


class _TypeInfoTuple(tuple):
    def size(self):
        return len(self)


def get_types(arrays):
    return _TypeInfoTuple(_TypeInfo(a) for a in arrays)


def expect(*conds):
    for i, c in enumerate(conds):
        if not bool(c):
            raise InvalidType('type_check.expect: condition %d failed' % i)
