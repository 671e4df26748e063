"""``cg`` plugin -- drop-in for reference ``betty/hypergradient/cg.py:8-70``."""
from .. import _native as N
from .. import engine as E


def cg(vector, curr, prev, sync):
    """Best-response-Jacobian x vector by K conjugate-gradient steps for H x = v, with the
    reference's exact ``cg_alpha`` placement.  alpha/beta never leave the device."""
    assert len(curr.paths) == 0, "cg method is not supported for higher order MLO!"
    call = E.HypergradientCall(curr, "cg")
    return call.finish(prev, call.solve(vector), sync)
