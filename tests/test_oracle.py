"""CPU tests: the oracle port against (i) golden vectors produced by the real reference
(oracle/make_golden.py) and (ii) the fp64 dense restatement."""
import glob
import os

import numpy as np
import pytest
import torch

from betty_b200 import workloads as W
from oracle import dense, ref_port
from tests.helpers import GOLDEN, assert_close, checksum, flat, load_golden

CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, "*.pt")))
# more model families (ResNet, mini-ImageNet 4-conv, a longer CG) pinned for the oracle only: the GPU parity suite
# globs tests/golden/, these live in tests/golden_cpu/
GOLDEN_CPU = os.path.join(os.path.dirname(GOLDEN), "golden_cpu")
CPU_CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_CPU, "*.pt")))


@pytest.mark.parametrize("case", CASES)
def test_port_matches_reference_golden(case):
    rec = load_golden(case)
    torch.set_num_threads(1)
    wl = W.FACTORIES[rec["factory"]](device="cpu", **rec["kwargs"])
    assert abs(checksum(wl) - rec["checksum"]) <= 1e-6 * max(1.0, abs(rec["checksum"])), "seeded inputs drifted"
    hg = ref_port.METHODS[rec["method"]](wl.vector, wl.lower, wl.upper, False)
    assert_close(hg, rec["hypergrad"], 1e-5, case)
    if "ihvp" in rec:
        x = ref_port.k_loop_only(rec["method"], wl.vector, wl.lower)
        assert_close(x, rec["ihvp"], 1e-5, case + " ihvp")


@pytest.mark.parametrize("case", CPU_CASES)
def test_port_matches_reference_golden_more_families(case):
    rec = torch.load(os.path.join(GOLDEN_CPU, case + ".pt"), weights_only=False)
    torch.set_num_threads(1)
    wl = W.FACTORIES[rec["factory"]](device="cpu", **rec["kwargs"])
    assert abs(checksum(wl) - rec["checksum"]) <= 1e-6 * max(1.0, abs(rec["checksum"])), "seeded inputs drifted"
    hg = ref_port.METHODS[rec["method"]](wl.vector, wl.lower, wl.upper, False)
    assert_close(hg, rec["hypergrad"], 1e-5, case)
    if "ihvp" in rec:
        assert_close(ref_port.k_loop_only(rec["method"], wl.vector, wl.lower), rec["ihvp"], 1e-5, case + " ihvp")


def test_port_sync_accumulates_into_upper_grads():
    wl = W.logistic_hpo(method="cg", K=4)
    want = ref_port.cg(wl.vector, wl.lower, wl.upper, False)
    wl.upper.zero_grad()
    assert ref_port.cg(wl.vector, wl.lower, wl.upper, True) is None
    got = [p.grad for p in wl.upper.trainable_parameters()]
    assert_close(got, want, 1e-6, "sync path")


@pytest.mark.parametrize("method,K,alpha", [("neumann", 5, 1.0), ("cg", 3, 0.1), ("cg", 20, 1.0)])
def test_port_matches_dense_fp64(method, K, alpha):
    wl = W.logistic_hpo(method=method, K=K, alpha=alpha)
    H, M = dense.dense_blocks(wl.lower, wl.upper)
    v = flat(wl.vector).numpy()
    x = dense.neumann_dense(H, v, K, alpha) if method == "neumann" else dense.cg_dense(H, v, K, alpha)
    want = dense.hypergradient_dense(M, x)
    got = ref_port.METHODS[method](wl.vector, wl.lower, wl.upper, False)
    assert_close(got, [torch.from_numpy(want)], 2e-4, f"dense {method}")
    if method == "cg" and K == 20:
        exact = -(M.T @ np.linalg.solve(H, v))
        assert np.linalg.norm(want - exact) / np.linalg.norm(exact) < 1e-10


def test_darts_is_minus_mixed_block_times_v():
    wl = W.logistic_hpo(method="darts")
    H, M = dense.dense_blocks(wl.lower, wl.upper)
    got = ref_port.darts(wl.vector, wl.lower, wl.upper, False)
    want = -(M.T @ flat(wl.vector).numpy())
    assert_close(got, [torch.from_numpy(want)], 5e-3, "darts ~ -M^T v")
