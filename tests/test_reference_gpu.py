"""The drop-in boundary exercised with the REAL reference objects on CUDA.

``oracle/_ref`` is the unmodified reference (mirrored by oracle/fetch_ref.sh; it travels to the GPU box).  Real
``betty.problems.ImplicitProblem`` objects are wired by a real ``betty.engine.Engine``; the hypergradient goes through
the reference's own ``betty.hypergradient.get_grads`` -- first with its own table (torch autograd on the same GPU),
then after ``betty_b200.install()`` rebinds the table.  fp32 bar: 1e-4 (BASELINE.json north_star).

Also: the reference's own regression suite of this path (test/test_regression.py: darts / cg / neumann,
``loss < 0.48``) run unmodified on the rebound table, and (2 GPUs) the ``sync=True`` all-reduce through a real
DDP-wrapped upper module over NCCL.
"""
import importlib.util
import os
import socket
import unittest

import pytest
import torch
import torch.nn.functional as F

from betty_b200 import workloads as W
from oracle import reference as R
from tests.helpers import assert_close, rel_l2

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(), reason="oracle/_ref not fetched (bash oracle/fetch_ref.sh)")]


def _ce_upper_step(p, batch):
    x, y = batch[0], batch[-1]
    return F.cross_entropy(p.peers["lower"].module(x), y.long())


def _logistic_upper_step(p, batch):
    x, y = batch
    return F.binary_cross_entropy_with_logits(p.peers["lower"].module(x)[0], y)


CASES = {
    "logistic_neumann": ("logistic_regression_hpo", dict(method="neumann", K=5), _logistic_upper_step),
    "logistic_cg_quirk": ("logistic_regression_hpo", dict(method="cg", K=3, alpha=0.1), _logistic_upper_step),
    "logistic_darts": ("logistic_regression_hpo", dict(method="darts"), _logistic_upper_step),
    "mlp_cg": ("mlp_reweight", dict(method="cg", K=5), _ce_upper_step),
    "lenet_cg": ("learning_to_reweight", dict(method="cg", batch=32, K=8), _ce_upper_step),
    "fourconv_neumann": ("implicit_maml", dict(method="neumann", n=10, hidden=16, K=10), _ce_upper_step),
}


@pytest.fixture
def reference_table():
    """The reference's plugin table, restored after the test whatever ``install()`` did to it."""
    R.load()
    import betty.hypergradient as RH

    keep = dict(RH.jvp_fn_mapping)
    yield RH
    RH.jvp_fn_mapping.clear()
    RH.jvp_fn_mapping.update(keep)


@pytest.mark.parametrize("case", sorted(CASES))
def test_install_is_a_drop_in_for_the_real_engine(case, reference_table):
    import betty_b200

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fac, kw, upper_step = CASES[case]
    wl = W.FACTORIES[fac](device="cpu", **kw)
    engine, upper, lower = R.real_problems(wl, upper_step, strategy="gpu")
    assert next(lower.module.parameters()).is_cuda and lower.cur_batch[0].is_cuda
    want = R.hypergradient_through_reference(upper, lower)          # the reference's own plugin, same GPU
    # the reference's fp32 result on a GPU is not bit-reproducible (atomics in cuDNN / index_add backward); where two
    # runs of the REFERENCE differ by more than 2e-5 the 1e-4 bar is widened by exactly that measured self-distance
    # (profiles/r02_parity_noise.md) -- on every case seen so far it is not
    again = R.hypergradient_through_reference(upper, lower)
    self_dist = rel_l2(again, want)
    tol = 1e-4 if self_dist <= 2e-5 else 1e-4 + self_dist
    print(f"[reference parity] {case}: reference-vs-reference {self_dist:.3e}, tolerance {tol:.2e}")
    table = betty_b200.install()
    assert table is reference_table.jvp_fn_mapping and table["cg"].__module__.startswith("betty_b200")
    got = R.hypergradient_through_reference(upper, lower)
    assert_close(got, want, tol, case)
    # sync=True: accumulate into .grad (through autograd.backward, so a DDP reducer would fire) and return None
    for p in upper.trainable_parameters():
        p.grad = None
    assert R.hypergradient_through_reference(upper, lower, do_sync=True) is None
    assert_close([p.grad for p in upper.trainable_parameters()], want, tol, case + " sync")


def test_reference_regression_suite_on_the_rebound_table(reference_table):
    """test/test_regression.py of the reference (2000 iterations, unroll 100, `loss < 0.48`), unmodified, with the
    three plugin entries replaced by the B200 engine."""
    import betty_b200

    path = os.path.join(R.REF_ROOT, "test", "test_regression.py")
    spec = importlib.util.spec_from_file_location("betty_reference_test_regression", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    betty_b200.install()
    suite = unittest.TestSuite()
    for name in ("test_darts", "test_cg", "test_neumann"):
        suite.addTest(mod.RegressionTest(name))
    result = unittest.TextTestRunner(verbosity=0).run(suite)
    assert result.testsRun == 3 and result.wasSuccessful(), (result.failures, result.errors)


# ---- 2 GPUs: sync=True through a real DDP-wrapped upper module over NCCL --------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ddp_worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import betty_b200
    from betty_b200 import workloads as W
    from oracle import reference as R

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    R.load()
    wl = W.mlp_reweight(device="cpu", method="cg", K=4, seed=0)      # same parameters on every rank
    g = torch.Generator().manual_seed(100 + rank)                     # rank-specific batch
    wl.lower.cur_batch = (torch.randn(64, 32, generator=g), torch.randint(0, 10, (64,), generator=g))
    # real problems + Engine on this rank's GPU; the upper module DDP-wrapped the way the reference's
    # strategy="distributed" does it (problem.py:218-224) -- that strategy itself needs a torch DataLoader to
    # re-shard, which a fixed synthetic batch is not, so the wrap is done here
    torch.cuda.set_device(rank)
    torch.distributed.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    engine, upper, lower = R.real_problems(wl, _ce_upper_step, strategy="default")
    upper.module = torch.nn.parallel.DistributedDataParallel(upper.module, device_ids=[rank])
    out = {}
    for tag in ("reference", "b200"):
        if tag == "b200":
            betty_b200.install()
        local = R.hypergradient_through_reference(upper, lower, do_sync=False)
        for p in upper.trainable_parameters():
            p.grad = None
        assert R.hypergradient_through_reference(upper, lower, do_sync=True) is None
        out[tag] = {"local": [t.detach().cpu() for t in local],
                    "synced": [p.grad.detach().cpu().clone() for p in upper.trainable_parameters()]}
    torch.save(out, os.path.join(out_dir, f"r{rank}.pt"))
    torch.distributed.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sync_allreduce_through_real_ddp_over_nccl(tmp_path):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_ddp_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    recs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    for tag in ("reference", "b200"):
        for a, b in zip(recs[0][tag]["synced"], recs[1][tag]["synced"]):
            assert torch.equal(a, b), tag                                   # identical after the reducer
        mean = [(a + b) / 2 for a, b in zip(recs[0][tag]["local"], recs[1][tag]["local"])]
        assert rel_l2(recs[0][tag]["synced"], mean) < 1e-5, tag             # = average of the local solves
        assert rel_l2(recs[0][tag]["local"], recs[1][tag]["local"]) > 1e-3  # no hidden communication in the K-loop
    assert_close(recs[0]["b200"]["synced"], recs[0]["reference"]["synced"], 1e-4, "DDP sync")
