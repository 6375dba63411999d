"""paddle.decomposition: decompose composite ops into primitives. Parity: python/paddle/decomposition/.
Eager design: composite ops already run as fused kernels or primitive torch ops, so `decompose` is the identity on
programs; `register_decomp` keeps a registry for static.passes."""
_REG = {}


def register_decomp(op_type):
    def deco(fn):
        _REG[op_type] = fn
        return fn

    return deco


def decompose(program, src_vars=None, blacklist=frozenset(), whitelist=frozenset()):
    return src_vars if src_vars is not None else program


def get_decomp_rule(op_type):
    return _REG.get(op_type)
