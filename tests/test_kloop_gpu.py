"""GPU tests for the flat-arena kernels K1-K4 (csrc/kloop.cu), called through the C ABI."""
import pytest
import torch

from betty_b200 import _native as N
from betty_b200.arena import ArenaLayout, ChunkTable, pack, stream_ptr
from betty_b200.engine import Workspace

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, generator=g).to(DEV)


@pytest.mark.parametrize("n", [4, 1000 * 4, 62008, 3_000_000])
def test_neumann_update(n):
    v, p, hv = _rand(n, 1), _rand(n, 2), _rand(n, 3)
    v0, p0 = v.clone(), p.clone()
    N.call("bb_neumann_update", v.data_ptr(), p.data_ptr(), hv.data_ptr(), 0.3, 0.0, n, stream_ptr())
    want_v = v0 - 0.3 * hv
    # the kernel contracts v - alpha*hv into one FMA; torch rounds the product first
    assert torch.allclose(v, want_v, rtol=1e-6, atol=1e-6)
    assert torch.allclose(p, p0 + want_v, rtol=1e-6, atol=2e-6)
    # with a declared c*I shift
    v2, p2 = v0.clone(), p0.clone()
    N.call("bb_neumann_update", v2.data_ptr(), p2.data_ptr(), hv.data_ptr(), 0.3, 2.0, n, stream_ptr())
    assert torch.allclose(v2, v0 - 0.3 * (hv + 2.0 * v0), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n", [8, 62008, 5_000_000])
@pytest.mark.parametrize("cg_alpha", [1.0, 0.1])
def test_cg_iteration_matches_reference_arithmetic(n, cg_alpha):
    """One full reference iteration (cg.py:42-53) in torch vs the three kernels."""
    r, p, hp, x = _rand(n, 1), _rand(n, 2), _rand(n, 3), _rand(n, 4)
    hp = hp + 2 * p  # keep (hp.p) away from 0
    ws = Workspace(torch.device(DEV))
    # torch restatement
    rr = torch.dot(r, r)
    alpha = rr / torch.dot(cg_alpha * hp, p)
    x_new = x + alpha * p
    r_new = r - alpha * hp
    beta = torch.dot(r_new, r_new) / rr
    p_new = r_new + beta * p
    s = stream_ptr()
    N.call("bb_cg_dots", r.data_ptr(), hp.data_ptr(), p.data_ptr(), cg_alpha, 0.0, 1, n, ws.ptr, s)
    N.call("bb_cg_update_xr", x.data_ptr(), r.data_ptr(), p.data_ptr(), hp.data_ptr(), 0.0, n, ws.ptr, s)
    N.call("bb_cg_update_p", p.data_ptr(), r.data_ptr(), n, ws.ptr, s)
    sc = ws.scalars.cpu()
    assert abs(float(sc[3]) - float(alpha)) <= 2e-5 * abs(float(alpha))
    assert abs(float(sc[4]) - float(beta)) <= 2e-5 * abs(float(beta))
    assert abs(float(sc[0]) - float(torch.dot(r_new, r_new))) <= 2e-5 * float(sc[0])  # rr rolled forward
    for got, want in ((x, x_new), (r, r_new), (p, p_new)):
        assert float((got - want).norm() / want.norm()) < 2e-6
    # determinism: same inputs -> bit-identical scalars
    r2, p2, hp2, x2 = _rand(n, 1), _rand(n, 2), _rand(n, 3) + 2 * _rand(n, 2), _rand(n, 4)
    ws2 = Workspace(torch.device(DEV))
    N.call("bb_cg_dots", r2.data_ptr(), hp2.data_ptr(), p2.data_ptr(), cg_alpha, 0.0, 1, n, ws2.ptr, s)
    assert float(ws2.scalars[3]) == float(sc[3])


@pytest.mark.parametrize("n", [62008, 5_000_000])
def test_cg_iteration_with_a_folded_identity_term(n):
    """`shift`: the kernels see hp WITHOUT a declared c*I curvature term and add shift*p themselves -- same iteration
    as handing them hp + shift*p (reference cg.py:42-53 on the full Hessian-vector product)."""
    shift, cg_alpha = 0.1, 1.0
    r, p, hp, x = _rand(n, 1), _rand(n, 2), _rand(n, 3), _rand(n, 4)
    hp = hp + 2 * p
    full = hp + shift * p
    outs = []
    for h, sh in ((hp, shift), (full, 0.0)):
        rr_, pp_, xx_ = r.clone(), p.clone(), x.clone()
        ws = Workspace(torch.device(DEV))
        s = stream_ptr()
        N.call("bb_cg_dots", rr_.data_ptr(), h.data_ptr(), pp_.data_ptr(), cg_alpha, sh, 1, n, ws.ptr, s)
        N.call("bb_cg_update_xr", xx_.data_ptr(), rr_.data_ptr(), pp_.data_ptr(), h.data_ptr(), sh, n, ws.ptr, s)
        N.call("bb_cg_update_p", pp_.data_ptr(), rr_.data_ptr(), n, ws.ptr, s)
        outs.append((xx_, rr_, pp_, ws.scalars.cpu().clone()))
    for a, b in zip(outs[0][:3], outs[1][:3]):
        assert float((a - b).norm() / b.norm()) < 2e-6
    assert abs(float(outs[0][3][3]) - float(outs[1][3][3])) <= 2e-6 * abs(float(outs[1][3][3]))   # alpha
    assert abs(float(outs[0][3][4]) - float(outs[1][3][4])) <= 2e-5 * abs(float(outs[1][3][4]))   # beta


def test_pack_views_roundtrip_and_padding():
    shapes = [(6, 3, 5, 5), (6,), (120, 400), (7,), ()]
    ts = [_rand(int(torch.Size(s).numel()), i).view(s) for i, s in enumerate(shapes)]
    lay = ArenaLayout.like(ts)
    assert lay.total % 4 == 0 and all(o % 4 == 0 for o in lay.offsets)
    flat = lay.new(torch.device(DEV))
    pack(lay, ts, flat)
    for t, v in zip(ts, lay.views(flat)):
        assert torch.equal(t, v)
    assert float(flat.sum()) == pytest.approx(float(sum(t.double().sum() for t in ts)), rel=1e-5)


def test_fd_kernels():
    shapes = [(33, 7), (5,), (70000,)]
    w = [_rand(int(torch.Size(s).numel()), 10 + i).view(s) for i, s in enumerate(shapes)]
    v = [_rand(int(torch.Size(s).numel()), 20 + i).view(s) for i, s in enumerate(shapes)]
    w0 = [t.clone() for t in w]
    dev = torch.device(DEV)
    ws = Workspace(dev)
    s = stream_ptr()
    tab = ChunkTable([t.data_ptr() for t in v], [t.data_ptr() for t in w], [t.numel() for t in v], dev)
    N.call("bb_mt_sumsq", tab.ptr, tab.n, ws.ptr, s)
    N.call("bb_fd_eps", ws.ptr, 0.01, s)
    nrm = torch.cat([t.reshape(-1) for t in v]).norm()
    eps = 0.01 / (float(nrm) + 1e-15)
    assert float(ws.scalars[6]) == pytest.approx(eps, rel=1e-6)
    assert float(ws.scalars[7]) == pytest.approx(1 / (2 * eps), rel=1e-6)
    N.call("bb_mt_sumsq", tab.ptr, tab.n, ws.ptr, s)  # slots were reset: same answer again
    assert float(ws.scalars[5]) == pytest.approx(float(nrm) ** 2, rel=1e-6)
    N.call("bb_mt_axpby", tab.ptr, tab.n, 1.0, ws.scalar_ptr(6), 1.0, s)
    for a, b, c in zip(w, w0, v):
        assert torch.allclose(a, b + eps * c, rtol=1e-6, atol=1e-7)
    N.call("bb_mt_axpby", tab.ptr, tab.n, -2.0, ws.scalar_ptr(6), 1.0, s)
    N.call("bb_mt_axpby", tab.ptr, tab.n, 1.0, ws.scalar_ptr(6), 1.0, s)
    for a, b in zip(w, w0):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)
    gm = [_rand(t.numel(), 30 + i).view(t.shape) for i, t in enumerate(w)]
    gp = [_rand(t.numel(), 40 + i).view(t.shape) for i, t in enumerate(w)]
    want = [(a - b) / (2 * eps) for a, b in zip(gm, gp)]
    tc = ChunkTable([t.data_ptr() for t in gm], [t.data_ptr() for t in gp], [t.numel() for t in gm], dev)
    N.call("bb_mt_fd_combine", tc.ptr, tc.n, ws.ptr, s)
    for a, b in zip(gm, want):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5 * float(b.abs().max()))
