class _Config(object):
    train = True
    use_cudnn = 'never'


config = _Config()
