from .variable import Variable, as_variable


class FunctionNode(object):
    """Only `apply` with identity-like forward is needed (models/utils.py CPU2GPU/GPU2CPU are
    never reached on the numpy path: transform.py:76 tests `xp != np`)."""

    def apply(self, inputs):
        vs = [as_variable(x) for x in inputs]
        outs = self.forward(tuple(v.data for v in vs))
        return tuple(Variable(o) for o in outs)
