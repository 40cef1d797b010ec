def check_unexpected_kwargs(kwargs, **unexpected):
    for k, msg in unexpected.items():
        if k in kwargs:
            raise ValueError(msg)


def assert_kwargs_empty(kwargs):
    if kwargs:
        raise TypeError('unexpected keyword arguments: %s' % sorted(kwargs))
