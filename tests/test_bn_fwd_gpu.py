"""Native prologue BatchNorm forward (csrc/bn_fwd.cu) against aten.native_batch_norm on the same GPU and
against an fp64 restatement; and the recorder path (betty_b200/trace.py) that substitutes it inside the lower forward."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref64(x, w, b, eps):
    xd = x.double()
    mean = xd.mean(dim=(0, 2, 3))
    var = xd.var(dim=(0, 2, 3), unbiased=False)
    invstd = 1.0 / torch.sqrt(var + eps)
    g = w.double() if w is not None else torch.ones_like(mean)
    bb = b.double() if b is not None else torch.zeros_like(mean)
    y = g[None, :, None, None] * (xd - mean[None, :, None, None]) * invstd[None, :, None, None] + bb[None, :, None, None]
    n = x.numel() // x.shape[1]
    return y, mean, invstd, var * n / (n - 1)


@pytest.mark.parametrize("shape", [(32, 64, 84, 84), (16, 64, 42, 42), (8, 64, 21, 21), (5, 3, 7, 9), (2, 130, 16, 16),
                                   (1, 8, 64, 64)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("affine", [True, False])
def test_bn_forward_matches_aten_and_fp64(shape, dtype, affine):
    from betty_b200 import _native as N

    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * 1.7 + 0.6).to(dev).to(dtype)
    c = shape[1]
    w = (torch.rand(c, generator=g) + 0.5).to(dev) if affine else None
    b = torch.randn(c, generator=g).to(dev) if affine else None
    eps = 1e-5
    y = torch.empty_like(x)
    stats = torch.empty((3, c), dtype=torch.float32, device=dev)
    S = N.lib().bb_bn_forward_splits(shape[0], c)
    assert 1 <= S <= max(1, shape[0])
    ws = torch.empty(2 * c * S, dtype=torch.float64, device=dev)
    N.call("bb_bn_forward", x.data_ptr(), 0 if dtype == torch.float32 else 1, w.data_ptr() if affine else None,
           b.data_ptr() if affine else None, eps, y.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
           stats[2].data_ptr(), ws.data_ptr(), shape[0], c, shape[2] * shape[3], torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    y64, m64, i64, v64 = _ref64(x, w, b, eps)
    assert torch.allclose(stats[0].double(), m64, rtol=1e-6, atol=1e-7)
    assert torch.allclose(stats[1].double(), i64, rtol=1e-6, atol=0)
    assert torch.allclose(stats[2].double(), v64, rtol=1e-6, atol=0)
    ya, ma, ia = torch.ops.aten.native_batch_norm(x, w, b, None, None, True, 0.1, eps)
    assert torch.allclose(stats[0], ma, rtol=1e-5, atol=1e-6) and torch.allclose(stats[1], ia, rtol=1e-5, atol=0)
    if dtype == torch.float32:
        assert torch.allclose(y.double(), y64, rtol=1e-5, atol=1e-5)
        assert torch.allclose(y, ya, rtol=1e-5, atol=1e-5)
    else:
        # one bf16 rounding of the fp32 result: within half an ulp of the exact value (+ fp32 noise), and at most one
        # bf16 ulp from aten's own rounding of the same expression
        err = (y.double() - y64).abs()
        assert bool((err <= y64.abs() * 2.0 ** -8 + 1e-6).all())
        assert bool(((y.float() - ya.float()).abs() <= ya.float().abs() * 2.0 ** -7 + 1e-6).all())
        assert float((y != ya).float().mean()) < 0.02
    # bit-reproducible
    y2 = torch.empty_like(x)
    stats2 = torch.empty_like(stats)
    N.call("bb_bn_forward", x.data_ptr(), 0 if dtype == torch.float32 else 1, w.data_ptr() if affine else None,
           b.data_ptr() if affine else None, eps, y2.data_ptr(), stats2[0].data_ptr(), stats2[1].data_ptr(),
           stats2[2].data_ptr(), ws.data_ptr(), shape[0], c, shape[2] * shape[3], torch.cuda.current_stream().cuda_stream)
    assert torch.equal(y, y2) and torch.equal(stats, stats2)


class _Net(torch.nn.Module):
    def __init__(self, track):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, 16, 3, padding=1)
        self.bn = torch.nn.BatchNorm2d(16, track_running_stats=track)
        self.fc = torch.nn.Linear(16, 4)

    def forward(self, x):
        h = torch.relu(self.bn(self.conv(x)))
        return self.fc(h.mean(dim=(2, 3)))


@pytest.mark.parametrize("track", [False, True])
@pytest.mark.parametrize("autocast", [False, True])
def test_recorder_substitutes_batch_norm(track, autocast, monkeypatch):
    import copy

    from betty_b200 import trace as T

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = _Net(track).to(dev)
    twin = copy.deepcopy(net)
    x = torch.randn(32, 3, 24, 24, device=dev)
    y = torch.randint(0, 4, (32,), device=dev)

    def step(m):
        # fp32 inputs would take aten.cudnn_batch_norm (fast already, left alone): force the native op the recorder handles
        with torch.backends.cudnn.flags(enabled=False), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            return torch.nn.functional.cross_entropy(m(x).float(), y)

    monkeypatch.setattr(T, "native_bn_min_numel", 1)
    before = T.native_bn_calls
    loss, tape = T.record_tape(lambda: step(net), list(net.parameters()))
    assert T.native_bn_calls == before + 1
    assert any(op.name in T._BN_FWD_OPS for op in tape.ops)
    want = step(twin)
    tol = 2e-2 if autocast else 1e-5
    assert abs(float(loss) - float(want)) <= tol * abs(float(want))
    ga = torch.autograd.grad(loss, list(net.parameters()))
    gb = torch.autograd.grad(want, list(twin.parameters()))
    for a, b in zip(ga, gb):
        assert float((a - b).norm()) <= tol * float(b.norm()) + 1e-6
    if track:
        assert torch.allclose(net.bn.running_mean, twin.bn.running_mean, rtol=1e-4, atol=1e-5)
        assert torch.allclose(net.bn.running_var, twin.bn.running_var, rtol=1e-4, atol=1e-5)
    # switched off -> PyTorch's own kernel runs
    monkeypatch.setattr(T, "native_bn_min_numel", 0)
    T.record_tape(lambda: step(net), list(net.parameters()))
    assert T.native_bn_calls == before + 1
