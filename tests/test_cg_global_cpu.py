"""Global-batch CG with rank-sharded vectors (betty_b200/hypergradient/cg_global.py, SURVEY.md §8e optional variant):
the communication schedule on a world_size-2 gloo group against the single-process reference recurrence on the
concatenated batch.  The K-loop kernels need CUDA, so the slice operations are restated with torch here (same arithmetic
as K2 / K3, csrc/kloop.cu) and the local product comes from the oracle; `solve_sharded` itself is the product code."""
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


class TorchSliceOps:
    """What NativeSliceOps does with bb_cg_init / bb_cg_dots / bb_cg_update_xr / bb_cg_update_p, in torch (test only)."""

    def __init__(self):
        self.s = torch.zeros(5, dtype=torch.float64)    # rr, php, rr_new, alpha, beta
        self._i = {"rr": 0, "php": 1, "rr_new": 2}

    def scalar(self, name):
        i = self._i[name]
        return self.s[i:i + 1]

    def rr_init(self, r):
        self.s[0] = torch.dot(r.double(), r.double())

    def dots(self, r, hp, p, cg_alpha):
        self.s[1] = torch.dot((cg_alpha * hp).double(), p.double())
        self.s[3] = self.s[0] / self.s[1]               # the kernel's (partial) quotient: must be overwritten

    def set_alpha(self):
        self.s[3] = self.s[0] / self.s[1]

    def update_xr(self, x, r, p, hp):
        rr_old = self.s[0:1].clone()
        a = float(self.s[3])
        x.add_(p, alpha=a)
        r.sub_(hp, alpha=a)
        self.s[2] = torch.dot(r.double(), r.double())
        self.s[4] = self.s[2] / self.s[0]               # partial, as the kernel leaves it
        self.s[0] = self.s[2]
        return rr_old

    def set_beta(self, rr_old):
        self.s[4] = self.s[2] / rr_old[0]
        self.s[0] = self.s[2]

    def update_p(self, p, r):
        p.mul_(float(self.s[4])).add_(r)


def _problem(rank_batches, seed=0):
    """mlp_reweight workload whose lower batch is the given (x, y)."""
    from betty_b200 import workloads as W

    wl = W.mlp_reweight(device="cpu", method="cg", K=4, seed=seed)
    wl.lower.cur_batch = rank_batches
    return wl


def _batches(world):
    out = []
    for rank in range(world):
        g = torch.Generator().manual_seed(100 + rank)
        out.append((torch.randn(48, 32, generator=g), torch.randint(0, 10, (48,), generator=g)))
    return out


def _worker(rank, world, port, out_dir, K, cg_alpha):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from betty_b200.arena import ArenaLayout
    from betty_b200.hypergradient.cg_global import solve_sharded
    from oracle import ref_port

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    wl = _problem(_batches(world)[rank])
    params = wl.lower.trainable_parameters()
    lay = ArenaLayout.like(params)
    in_grad = ref_port.lower_gradient(wl.lower)
    hvp = ref_port.make_hvp(in_grad, wl.lower.parameters())
    # rank-specific right-hand side (each rank's own upper batch would give its own v)
    g = torch.Generator().manual_seed(7 + rank)
    v = [torch.randn(p.shape, generator=g) for p in params]
    flat = lay.new("cpu")
    for view, t in zip(lay.views(flat), v):
        view.copy_(t)
    dist.all_reduce(flat)
    flat.div_(world)

    def local_hvp(p_full, out_full):
        hp = hvp([t.clone() for t in lay.views(p_full.clone())])
        tmp = lay.new("cpu")
        for view, t in zip(lay.views(tmp), hp):
            view.copy_(t)
        out_full.copy_(tmp)

    x = solve_sharded(flat, lay.total, K, cg_alpha, local_hvp, TorchSliceOps())
    torch.save({"x": [t.clone() for t in lay.views(x.clone())], "v": v}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("cg_alpha", [1.0, 0.3])
def test_sharded_schedule_matches_single_process_reference_on_the_concatenated_batch(tmp_path, cg_alpha):
    from oracle import ref_port

    world, K = 2, 4
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), K, cg_alpha), nprocs=world, join=True)
    recs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    # every rank holds the same solution
    for a, b in zip(recs[0]["x"], recs[1]["x"]):
        assert torch.equal(a, b)
    # single process: mean of the per-rank losses = loss of the concatenated batch (equal batch sizes)
    bs = _batches(world)
    wl = _problem((torch.cat([b[0] for b in bs]), torch.cat([b[1] for b in bs])))
    in_grad = ref_port.lower_gradient(wl.lower)
    hvp = ref_port.make_hvp(in_grad, wl.lower.parameters())
    v_mean = [(a + b) / world for a, b in zip(recs[0]["v"], recs[1]["v"])]
    want = ref_port.cg_solve(v_mean, hvp, K, cg_alpha)
    num = torch.sqrt(sum(((a.double() - b.double()) ** 2).sum() for a, b in zip(recs[0]["x"], want)))
    den = torch.sqrt(sum((b.double() ** 2).sum() for b in want))
    assert float(num / den) < 2e-5, float(num / den)


def test_table_has_the_additional_key():
    from betty_b200 import hypergradient as H

    assert callable(H.jvp_fn_mapping["cg_global"])
    assert set(H.jvp_fn_mapping) >= {"neumann", "cg", "darts", "finite_diff", "sama", "cg_global"}
