"""Plan cache (engine.PlanCache / HvpPlan.rebind): a second plugin call whose forward has the same tape signature
reuses the native plan -- descriptors, buffers, tensor maps and the captured K-loop graph -- with the new forward's
values copied behind the pointers the plan holds.  Every value changes between the calls here (parameters, batch,
direction, upper parameters), so a stale constant anywhere shows up as a mismatch against a cache-less call."""
import pytest
import torch

from betty_b200 import _native as N
from betty_b200 import engine as E
from betty_b200 import hypergradient as H
from betty_b200 import workloads as W
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu

CASES = {
    "logistic_neumann": ("logistic_regression_hpo", dict(method="neumann", K=5)),
    "mlp_cg": ("mlp_reweight", dict(method="cg", K=5)),
    "lenet_cg": ("learning_to_reweight", dict(method="cg", batch=16, K=6)),
    "fourconv_neumann": ("implicit_maml", dict(method="neumann", n=10, hidden=16, K=6, alpha=0.01)),
    "fourconv_mini_bf16": ("implicit_maml", dict(method="neumann", n=4, hidden=64, image="miniimagenet", K=6, alpha=0.01, precision="bf16")),
    "roberta_cg": ("bert_data_reweighting", dict(method="cg", batch=3, seq=9, K=4, tiny=True)),
    # (Neumann: CG on bf16 products amplifies the run-to-run order of the split-K atomics beyond any fixed tolerance)
    "roberta_bf16_neumann": ("bert_data_reweighting", dict(method="neumann", batch=4, seq=40, K=4, alpha=0.02, tiny=True, tiny_hidden=256, precision="bf16")),
}


def _perturb(wl, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    with torch.no_grad():
        for p in list(wl.lower.module.parameters()) + list(wl.upper.module.parameters()):
            p.add_(0.05 * torch.randn(p.shape, generator=g, device="cuda") * (p.abs().mean() + 1e-3))
    batch = []
    for b in wl.lower.cur_batch:
        if torch.is_tensor(b) and b.is_floating_point():
            b = b + 0.3 * torch.randn(b.shape, generator=g, device="cuda")
        elif torch.is_tensor(b) and b.dtype == torch.long and b.dim() == 1:
            b = b[torch.randperm(b.numel(), generator=g, device="cuda")]
        batch.append(b)
    wl.lower.cur_batch = tuple(batch)
    wl.vector = tuple(torch.randn(v.shape, generator=g, device="cuda") for v in wl.vector)


@pytest.fixture
def fresh_cache(monkeypatch):
    E.plan_cache.clear()
    E.plan_cache.hits = E.plan_cache.misses = 0
    yield E.plan_cache
    E.plan_cache.clear()


@pytest.mark.parametrize("case", sorted(CASES))
def test_cached_plan_serves_new_values(case, fresh_cache, monkeypatch):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fac, kw = CASES[case]
    method = kw["method"]
    fn = H.jvp_fn_mapping[method]
    wl = W.FACTORIES[fac](device="cuda", **kw)
    results = []
    for step in range(3):
        if step:
            _perturb(wl, 100 + step)
        results.append([t.clone() for t in fn(wl.vector, wl.lower, wl.upper, False)])
        # what a cache-less engine gives on exactly the same values
        monkeypatch.setenv("BB200_PLAN_CACHE", "0")
        want = fn(wl.vector, wl.lower, wl.upper, False)
        monkeypatch.delenv("BB200_PLAN_CACHE")
        tol = 1e-2 if kw.get("precision") == "bf16" else 2e-5   # same kernels, same values: only atomics order differs
        assert_close(results[-1], want, tol, f"{case} call {step}")
    assert fresh_cache.misses == 1 and fresh_cache.hits == 2, (fresh_cache.hits, fresh_cache.misses)
    plan = next(iter(fresh_cache.entries.values())).plan
    assert N.lib().bb_plan_graph_captures(plan.handle) <= 2     # one K-loop graph (+ one bare H.d graph for the epilogue)
    # and the three calls really were different problems
    assert not torch.allclose(results[0][0], results[1][0], rtol=1e-3, atol=0)


def test_cache_is_bypassed_while_a_call_still_holds_the_plan(fresh_cache):
    wl = W.mlp_reweight(device="cuda", method="cg", K=3)
    a = E.HypergradientCall(wl.lower, "cg")
    b = E.HypergradientCall(wl.lower, "cg")            # same signature, but `a` has not finished
    assert a.hvp is not b.hvp
    xa = [t.clone() for t in a.solve(wl.vector)]
    xb = [t.clone() for t in b.solve(wl.vector)]
    assert_close(xa, xb, 1e-6, "two live calls")
    a.release(); b.release()
    c = E.HypergradientCall(wl.lower, "cg")
    assert c.hvp is a.hvp or c.hvp is b.hvp
    c.release()


def test_data_dependent_refusal_is_rechecked_on_reuse(fresh_cache):
    from betty_b200.ir import UnsupportedGraph

    wl = W.mlp_reweight(device="cuda", method="cg", K=2)

    def step(p, batch):
        x, y = batch
        return torch.nn.functional.cross_entropy(p.module(x), y, ignore_index=-100)

    wl.lower._training_step = step
    H.cg(wl.vector, wl.lower, wl.upper, False)
    x, y = wl.lower.cur_batch
    y = y.clone()
    y[0] = -100                                         # same shapes, but now a target is ignored
    wl.lower.cur_batch = (x, y)
    with pytest.raises(UnsupportedGraph):
        H.cg(wl.vector, wl.lower, wl.upper, False)
