def load_npz(path, obj):
    raise RuntimeError('chainer shim: load_npz is not available')


def save_npz(path, obj):
    raise RuntimeError('chainer shim: save_npz is not available')
