"""``cg_global`` -- global-batch conjugate gradient with rank-sharded vectors (optional variant;
BASELINE.json north_star: "NCCL over NVLink reducing only the two CG scalars and the final hypergradient").

This is NOT what the reference computes: ``betty.hypergradient.cg`` solves each rank's LOCAL system ``H_rank x = v_rank``
and only averages the resulting hypergradients (DDP).  Here the system is the global one,

    H = mean_rank H_rank        v = mean_rank v_rank        x ~= H^-1 v   (K steps, the reference's recurrence, cg.py:34-56)

so the result is the hypergradient of the *global* batch, independent of how it is split over the ranks (oracle:
the single-process reference on the concatenated batch, tests/test_cg_global_cpu.py).  Select it with
``Config(type="cg_global")`` after ``betty_b200.install()``.

Layout (ZeRO-style): rank ``k`` owns slice ``k`` of the flat arenas ``x, r, p`` (slice length a multiple
of 4 floats, the arena padded to ``world`` slices).  Per iteration

    all-gather p            (P floats)            -> every rank's direction arena
    local H_rank . p        (the native plan, csrc/plan.cu: tangent forward + tangent backward on the rank's batch)
    reduce-scatter (H p)    (P floats, sum / world) -> the rank's slice of the global H p
    K2 on the slice         -> partial (cg_alpha H p) . p        ALL-REDUCE OF ONE SCALAR
    K3 on the slice         -> partial r' . r'                   ALL-REDUCE OF ONE SCALAR
    p-update on the slice

The slice kernels are the unchanged K2 / K3 entry points of the C ABI (``bb_cg_dots``, ``bb_cg_update_xr``,
``bb_cg_update_p``) called on slice pointers; their scalars live in the device workspace (``bb_kloop_scalars``) and are
completed across ranks there -- no value ever visits the host.

The communication schedule (`solve_sharded`) is device independent and takes the vector operations and the local
product as arguments; the plugin passes the native ones.  Only the tests pass anything else.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def _avg_(t: torch.Tensor, group, world: int):
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    if world > 1:
        t.div_(world)
    return t


def solve_sharded(v_mean: torch.Tensor, total: int, K: int, cg_alpha: float,
                  local_hvp: Callable[[torch.Tensor, torch.Tensor], None], ops, group=None) -> torch.Tensor:
    """K global-batch CG steps.  ``v_mean`` = flat fp32 arena (``total`` floats, multiple of 4) holding the rank-averaged
    right-hand side; ``local_hvp(p_full, out_full)`` writes ``H_rank . p_full`` into ``out_full`` (both ``total`` floats);
    ``ops`` = slice kernels + the device scalars they share (``NativeSliceOps``).  Returns ``cg_alpha * x`` as a full
    arena, identical on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = v_mean.device
    shard = -(-total // (4 * world)) * 4          # slice length: multiple of 4 floats (128-bit kernels)
    padded = shard * world
    lo = rank * shard
    full = torch.zeros(padded, dtype=torch.float32, device=dev)
    full[:total].copy_(v_mean[:total])
    r = full[lo:lo + shard].clone()               # cg.py:36-37: r = v, p = r, x = 0
    p = r.clone()
    x = torch.zeros(shard, dtype=torch.float32, device=dev)
    hp = torch.empty(shard, dtype=torch.float32, device=dev)
    p_full = torch.zeros(padded, dtype=torch.float32, device=dev)
    h_full = torch.zeros(padded, dtype=torch.float32, device=dev)   # padding stays zero

    ops.rr_init(r)                                                   # partial r.r
    dist.all_reduce(ops.scalar("rr"), op=dist.ReduceOp.SUM, group=group)
    for _ in range(K):
        dist.all_gather_into_tensor(p_full, p, group=group)
        local_hvp(p_full[:total], h_full[:total])                    # cg.py:39-41 on this rank's batch
        dist.reduce_scatter_tensor(hp, h_full, op=dist.ReduceOp.SUM, group=group)
        if world > 1:
            hp.div_(world)                                           # H = mean over ranks
        ops.dots(r, hp, p, cg_alpha)                                 # partial (cg_alpha H p).p      cg.py:42-46
        dist.all_reduce(ops.scalar("php"), op=dist.ReduceOp.SUM, group=group)
        ops.set_alpha()                                              # alpha = rr / php              cg.py:47
        rr_old = ops.update_xr(x, r, p, hp)                          # x += a p; r -= a Hp; partial r'.r'   cg.py:49-52
        dist.all_reduce(ops.scalar("rr_new"), op=dist.ReduceOp.SUM, group=group)
        ops.set_beta(rr_old)                                         # beta = rr' / rr; rr = rr'     cg.py:52
        ops.update_p(p, r)                                           # p = r + beta p                cg.py:53
    x_full = torch.empty(padded, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(x_full, x, group=group)
    return x_full[:total] * cg_alpha                                  # cg.py:56


class NativeSliceOps:
    """K2 / K3 of the C ABI on one rank's slice.  The kernels leave the slice's partial sums in the shared scalar block
    (``bb_kloop_scalars``: rr, php, rr_new, alpha, beta); `solve_sharded` all-reduces them in place and the two
    quotients are recomputed on the device from the completed sums."""
    _INDEX = {"rr": 0, "php": 1, "rr_new": 2, "alpha": 3, "beta": 4}

    def __init__(self, device):
        from .. import _native as N
        from ..engine import Workspace

        N.require_cuda()
        self.N = N
        self.dev = device
        self.ws = Workspace.get(device)

    def _s(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def scalar(self, name: str) -> torch.Tensor:
        i = self._INDEX[name]
        return self.ws.scalars[i:i + 1]

    def rr_init(self, r):
        self.N.call("bb_cg_init", r.data_ptr(), r.numel(), self.ws.ptr, self._s())

    def dots(self, r, hp, p, cg_alpha: float):
        self.N.call("bb_cg_dots", r.data_ptr(), hp.data_ptr(), p.data_ptr(), cg_alpha, 0.0, 0, r.numel(), self.ws.ptr,
                    self._s())

    def set_alpha(self):
        s = self.ws.scalars
        torch.div(s[0:1], s[1:2], out=s[3:4])

    def update_xr(self, x, r, p, hp):
        rr_old = self.ws.scalars[0:1].clone()        # the kernel overwrites rr with its partial r'.r'
        self.N.call("bb_cg_update_xr", x.data_ptr(), r.data_ptr(), p.data_ptr(), hp.data_ptr(), 0.0, r.numel(),
                    self.ws.ptr, self._s())
        return rr_old

    def set_beta(self, rr_old):
        s = self.ws.scalars
        torch.div(s[2:3], rr_old, out=s[4:5])
        s[0:1].copy_(s[2:3])

    def update_p(self, p, r):
        self.N.call("bb_cg_update_p", p.data_ptr(), r.data_ptr(), r.numel(), self.ws.ptr, self._s())


def cg_global(vector: Sequence[torch.Tensor], curr, prev, sync: bool, group=None) -> List[torch.Tensor] | None:
    """Plugin ``fn(vector, curr, prev, sync)``: global-batch CG over the ranks of ``group`` (default: the world).  With one
    process it equals ``cg``.  ``sync=True`` accumulates the rank-averaged hypergradient into ``.grad`` through
    ``autograd.backward`` (the upper module's DDP reducer averages it, as in the reference); ``sync=False`` returns the
    rank-averaged list."""
    from .. import engine as E
    from ..arena import pack

    assert len(curr.paths) == 0, "cg method is not supported for higher order MLO!"
    if not (dist.is_available() and dist.is_initialized()):
        return _single(vector, curr, prev, sync)          # one process: the global batch is the local one
    world = dist.get_world_size(group)
    call = E.HypergradientCall(curr, "cg")
    lay = call.layout
    with torch.cuda.device(call.dev):
        pack(lay, vector, call.d)                         # this rank's v in the direction arena
        v_mean = _avg_(call.d.clone(), group, world)

        def local_hvp(p_full, out_full):
            call.d.copy_(p_full)
            call.hvp()                                    # H_rank . d  ->  call.hd   (includes a folded c*I term)
            out_full.copy_(call.hd)

        x = solve_sharded(v_mean, lay.total, call.K, call.alpha, local_hvp, NativeSliceOps(call.dev), group)
        call.out.copy_(x)
        res = call.finish(prev, lay.views(call.out), sync)
        if res is None:
            return None
        return [_avg_(g.contiguous(), group, world) for g in res]


def _single(vector, curr, prev, sync):
    from .cg import cg

    return cg(vector, curr, prev, sync)
