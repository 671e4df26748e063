import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def flat(ts):
    return torch.cat([t.detach().reshape(-1).double().cpu() for t in ts])


def rel_l2(a, b):
    a, b = flat(a), flat(b)
    return float((a - b).norm() / (b.norm() + 1e-300))


def assert_close(eng, ref, rtol, what=""):
    """SURVEY.md §8(c) tolerance protocol: relative L2 and allclose(rtol, atol=rtol*|ref|_inf)."""
    e, r = flat(eng), flat(ref)
    rl2 = float((e - r).norm() / (r.norm() + 1e-300))
    assert rl2 <= rtol, f"{what}: rel-L2 {rl2:.3e} > {rtol:.1e}"
    atol = rtol * float(r.abs().max())
    assert torch.allclose(e, r, rtol=rtol, atol=atol), f"{what}: allclose(rtol={rtol}, atol={atol:.3e}) failed; max abs diff {float((e-r).abs().max()):.3e}"
    return rl2


def load_golden(case):
    return torch.load(os.path.join(GOLDEN, case + ".pt"), weights_only=False)


def checksum(wl):
    s = 0.0
    for t in list(wl.lower.module.parameters()) + list(wl.upper.module.parameters()) + list(wl.vector):
        s += float(t.detach().double().sum())
    for b in wl.lower.cur_batch:
        if torch.is_tensor(b):
            s += float(b.double().sum())
    return s


def to_double(wl):
    """fp64 copy of a workload in place (modules, floating inputs, direction)."""
    wl.lower.module.double()
    wl.upper.module.double()
    wl.lower.cur_batch = tuple(b.double() if torch.is_tensor(b) and b.is_floating_point() else b for b in wl.lower.cur_batch)
    wl.vector = tuple(v.double() for v in wl.vector)
    return wl
