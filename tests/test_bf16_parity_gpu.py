"""bf16 parity at the north_star bar: the full plugin call against the reference algorithm (oracle.ref_port: torch
autograd double backward) run UNDER THE SAME AUTOCAST ON THE SAME GPU, rel-L2 <= 1e-2 and
allclose(rtol=1e-2, atol=1e-2 |ref|_inf)  (BASELINE.json north_star: "rtol 1e-2 bf16"; autocast scope = reference
problem.py:327-332, K-loops neumann.py:59-66 / cg.py:34-56).

Shapes follow the two bf16 headline configs: config 3 (4-conv backbone, 64 channels -- the TMA tensor-core
convolution path -- Omniglot and mini-ImageNet inputs, Neumann K=20 alpha=0.01) and config 5 (RoBERTa blocks at
hidden 256 / 40 tokens -- tensor-core Linear and batched attention products -- CG K=10).  Each case also prints both
results' distance from the fp64 evaluation of the same problem, i.e. how much of the gap is the reference's own
bf16 noise."""
import pytest
import torch

from betty_b200 import hypergradient as H
from betty_b200 import workloads as W
from oracle import ref_port
from tests.helpers import assert_close, rel_l2, to_double

pytestmark = pytest.mark.gpu

CASES = {
    # config 3 shapes
    "fourconv_omniglot_neumann": ("implicit_maml", dict(method="neumann", n=8, hidden=64, K=20, alpha=0.01)),
    "fourconv_omniglot_n25_neumann": ("implicit_maml", dict(method="neumann", n=25, hidden=64, K=20, alpha=0.01)),
    "fourconv_mini_neumann": ("implicit_maml", dict(method="neumann", n=6, hidden=64, image="miniimagenet", K=20, alpha=0.01)),
    "fourconv_mini_cg": ("implicit_maml", dict(method="cg", n=6, hidden=64, image="miniimagenet", K=3, alpha=1.0)),
    # config 5 shapes
    "roberta_h256_cg": ("bert_data_reweighting", dict(method="cg", batch=4, seq=40, K=10, tiny=True, tiny_hidden=256)),
    "roberta_h256_neumann": ("bert_data_reweighting", dict(method="neumann", batch=4, seq=40, K=10, alpha=0.05, tiny=True, tiny_hidden=256)),
    "roberta_h128_b8_cg": ("bert_data_reweighting", dict(method="cg", batch=8, seq=16, K=10, tiny=True, tiny_hidden=128)),
    # Linear / weighted-CE only
    "mlp_w256_cg": ("mlp_reweight", dict(method="cg", batch=256, din=192, hidden=256, classes=64, K=8)),
}


@pytest.mark.parametrize("precision", ["bf16"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_bf16_plugin_matches_reference_under_the_same_autocast(case, precision):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fac, kw = CASES[case]
    method = kw["method"]
    wl = W.FACTORIES[fac](device="cuda", precision=precision, **kw)
    want = ref_port.METHODS[method](wl.vector, wl.lower, wl.upper, False)     # torch autograd under bf16 autocast
    got = H.jvp_fn_mapping[method](wl.vector, wl.lower, wl.upper, False)
    w64 = to_double(W.FACTORIES[fac](device="cuda", precision="fp32", **kw))
    exact = ref_port.METHODS[method](w64.vector, w64.lower, w64.upper, False)
    print(f"[bf16 parity] {case}: engine-vs-reference(bf16) {rel_l2(got, want):.3e}   engine-vs-fp64 "
          f"{rel_l2(got, exact):.3e}   reference(bf16)-vs-fp64 {rel_l2(want, exact):.3e}")
    assert_close(got, want, 1e-2, case)
