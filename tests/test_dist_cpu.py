"""world_size-2 gloo test (CPU) of the N>1 host path: every rank solves its own local
system; the only exchange is the DDP all-reduce(avg) fired by the sync epilogue
(engine.mixed_product -> torch.autograd.backward(in_grad, inputs=lambda, grad_tensors=-x)).
The K-loop itself needs CUDA, so x comes from the oracle here; mixed_product is device independent."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from betty_b200 import engine as E
    from betty_b200 import workloads as W
    from oracle import ref_port

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    # same parameters on every rank (seed 0), rank-specific batch (the reference's strided sampler)
    wl = W.mlp_reweight(device="cpu", method="cg", K=4, seed=0)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(64, 32, generator=g)
    y = torch.randint(0, 10, (64,), generator=g)
    wl.lower.cur_batch = (x, y)
    wl.upper.module = torch.nn.parallel.DistributedDataParallel(wl.upper.module)
    # local solve (no communication): x_rank = approx H_rank^-1 v
    in_grad = ref_port.lower_gradient(wl.lower)
    xs = ref_port.cg_solve(list(wl.vector), ref_port.make_hvp(in_grad, wl.lower.parameters()), 4, 1.0)
    local = [-t for t in torch.autograd.grad(in_grad, wl.upper.trainable_parameters(), grad_outputs=xs, retain_graph=True)]
    assert E.mixed_product(in_grad, wl.upper, xs, True) is None
    synced = [p.grad.clone() for p in wl.upper.trainable_parameters()]

    # the native epilogue (boundary seeds -> first-order backward through the DDP-wrapped upper module) must give
    # the same all-reduced result; the seeds come from the torch interpreter here (the CUDA plan needs a GPU)
    from betty_b200.ir import lower_tape
    from betty_b200.trace import record_tape
    from oracle.plan_interp import Interp

    for p in wl.upper.trainable_parameters():
        p.grad = None
    params = wl.lower.trainable_parameters()
    loss, tape = record_tape(lambda: wl.lower.training_step_exec(wl.lower.cur_batch), params)
    it = Interp(lower_tape(tape), torch.float32)
    it.base_backward()
    assert E.chain_boundary_seeds(it.mixed_seeds([t.detach() for t in xs]), wl.upper, True) is None
    synced_native = [p.grad.clone() for p in wl.upper.trainable_parameters()]
    torch.save({"local": local, "synced": synced, "synced_native": synced_native}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_sync_epilogue_allreduces_local_solves(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    recs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    for a, b in zip(recs[0]["synced"], recs[1]["synced"]):
        assert torch.equal(a, b)                       # identical after the reducer
    for i, s in enumerate(recs[0]["synced"]):
        mean = (recs[0]["local"][i] + recs[1]["local"][i]) / 2
        assert torch.allclose(s, mean, rtol=1e-5, atol=1e-7)
    for a, b in zip(recs[0]["synced_native"], recs[0]["synced"]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)
    for a, b in zip(recs[0]["synced_native"], recs[1]["synced_native"]):
        assert torch.equal(a, b)
    # and the local solves really differ (no hidden communication in the K-loop)
    assert not torch.allclose(recs[0]["local"][0], recs[1]["local"][0])
