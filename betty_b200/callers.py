"""Caller-side overheads around the hypergradient path -- optional drop-ins for three methods of
the reference's ``Problem`` classes that run once per hypergradient step, each per-tensor in the reference:

  * ``Problem.synchronize_params`` (reference problems/problem.py:599-609): one ``dist.broadcast`` /
    ``div_ + all_reduce`` **per parameter tensor** (201 collectives for RoBERTa-base, 1 399 for the DARTS
    supernet).  Here: one multi-tensor pack (``bb_mt_copy``, K4) into a flat fp32 arena, ONE collective, one
    multi-tensor unpack.  Same arithmetic (divide by world size, then sum).
  * ``ImplicitProblem.cache_states`` / ``recover_states`` (reference problems/implicit_problem.py:67-78):
    ``copy.deepcopy`` of the module and optimizer state dicts, then ``load_state_dict`` (a Python loop of
    per-tensor clones and ``copy_``).  Here: ONE arena snapshot of every fp32 tensor of both (one ``bb_mt_copy``
    each way, restored in place); the handful of non-fp32 entries (``num_batches_tracked``, Adam ``step``) are
    cloned individually; optimizer hyper-parameters (``param_groups`` minus the tensors) are deep-copied as before.

``install_callers()`` rebinds the three methods on the reference's classes; nothing else of ``Problem`` changes.
CUDA tensors go through the native multi-tensor kernels (and raise ``NativeError`` when the library is missing);
CPU tensors (the reference's ``strategy="cpu"`` / gloo case, PR1 plumbing) are flattened with torch.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Sequence

import torch

from . import _native as N
from .arena import ArenaLayout, stream_ptr


# --------------------------------------------------------------------------------------------------------------------
# flat pack / unpack of a tensor list
# --------------------------------------------------------------------------------------------------------------------
class FlatPack:
    """A fixed list of fp32 contiguous tensors on one device and a flat arena over them.  ``gather()`` copies
    tensors -> arena, ``scatter()`` arena -> tensors (in place).  The chunk table is built once and reused while the
    tensors keep their storage addresses (parameters do)."""

    def __init__(self, tensors: Sequence[torch.Tensor]):
        self.tensors = list(tensors)
        for t in self.tensors:
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("FlatPack takes contiguous fp32 tensors")
        self.device = self.tensors[0].device if self.tensors else torch.device("cpu")
        self.layout = ArenaLayout.like(self.tensors)
        self.flat = torch.zeros(self.layout.total, dtype=torch.float32, device=self.device)
        self._ptrs = tuple(t.data_ptr() for t in self.tensors)
        self._table = None
        if self.device.type == "cuda":
            N.require_cuda()
            self._table = self.layout.chunk_table(self.tensors, self.flat)

    def matches(self, tensors: Sequence[torch.Tensor]) -> bool:
        return len(tensors) == len(self._ptrs) and all(t.data_ptr() == p for t, p in zip(tensors, self._ptrs))

    def gather(self) -> torch.Tensor:
        if self._table is not None:
            with torch.cuda.device(self.device):
                N.call("bb_mt_copy", self._table.ptr, self._table.n, 0, stream_ptr())
        else:
            for v, t in zip(self.layout.views(self.flat), self.tensors):
                v.copy_(t)
        return self.flat

    def scatter(self):
        if self._table is not None:
            with torch.cuda.device(self.device):
                N.call("bb_mt_copy", self._table.ptr, self._table.n, 1, stream_ptr())
        else:
            for v, t in zip(self.layout.views(self.flat), self.tensors):
                t.copy_(v)


def _flat_ok(t: torch.Tensor) -> bool:
    return t.dtype == torch.float32 and t.is_contiguous() and t.numel() > 0


# --------------------------------------------------------------------------------------------------------------------
# Problem.synchronize_params
# --------------------------------------------------------------------------------------------------------------------
def synchronize_params(self, params, all_reduce: bool = False):
    """Replacement for reference ``Problem.synchronize_params`` (problems/problem.py:599-609): identical result, one
    collective for all fp32 parameters of a device instead of one per tensor."""
    import torch.distributed as dist

    if not (self._world_size > 1 and self._strategy not in ["fsdp", "accelerate"]):
        return
    params = list(params)
    datas = [p.data for p in params]
    flat_idx = [i for i, d in enumerate(datas) if _flat_ok(d)]
    rest = [i for i in range(len(datas)) if i not in set(flat_idx)]
    by_dev: Dict[torch.device, List[int]] = {}
    for i in flat_idx:
        by_dev.setdefault(datas[i].device, []).append(i)
    cache = self.__dict__.setdefault("_bb200_sync_packs", {})
    for dev, idx in by_dev.items():
        ts = [datas[i] for i in idx]
        pk = cache.get(dev)
        if pk is None or not pk.matches(ts):
            pk = cache[dev] = FlatPack(ts)
        flat = pk.gather()
        if not all_reduce:
            dist.broadcast(flat, 0)
        else:
            flat.div_(self._world_size)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        pk.scatter()
    for i in rest:   # non-fp32 / strided parameters: the reference's own per-tensor form
        if not all_reduce:
            dist.broadcast(datas[i], 0)
        else:
            datas[i].div_(self._world_size)
            dist.all_reduce(datas[i], op=dist.ReduceOp.SUM)


# --------------------------------------------------------------------------------------------------------------------
# ImplicitProblem.cache_states / recover_states
# --------------------------------------------------------------------------------------------------------------------
class StateSnapshot:
    """Values of a module's state dict and an optimizer's per-parameter state at one instant."""

    def __init__(self, module: torch.nn.Module, optimizer=None):
        self.module_items = list(module.state_dict(keep_vars=True).items())
        self.opt = optimizer
        self.opt_entries = []     # (param, key, tensor)
        self.opt_plain = []       # (param, {key: non-tensor value})
        self.opt_groups = None
        tensors = [t.detach() for _, t in self.module_items]
        if optimizer is not None:
            for group in optimizer.param_groups:
                for p in group["params"]:
                    st = optimizer.state.get(p, {})
                    plain = {}
                    for k, v in st.items():
                        if isinstance(v, torch.Tensor):
                            self.opt_entries.append((p, k, v))
                            tensors.append(v.detach())
                        else:
                            plain[k] = copy.deepcopy(v)
                    self.opt_plain.append((p, plain))
            self.opt_groups = [{k: copy.deepcopy(v) for k, v in g.items() if k != "params"}
                               for g in optimizer.param_groups]
        self.tensors = tensors
        by_dev: Dict[torch.device, List[torch.Tensor]] = {}
        self.others = []
        for t in tensors:
            if _flat_ok(t):
                by_dev.setdefault(t.device, []).append(t)
            else:
                self.others.append((t, t.clone()))
        self.packs = []
        for dev, ts in by_dev.items():
            pk = FlatPack(ts)
            pk.gather()
            self.packs.append(pk)

    def restore(self):
        """Write the snapshot back in place; optimizer state created after the snapshot is dropped, exactly as the
        reference's ``optimizer.load_state_dict`` of the cached dict does."""
        for pk in self.packs:
            pk.scatter()
        with torch.no_grad():
            for t, saved in self.others:
                t.copy_(saved)
        if self.opt is not None:
            for p, plain in self.opt_plain:
                keep = {k for (q, k, _) in self.opt_entries if q is p} | set(plain)
                st = self.opt.state.get(p)
                if st is None:
                    continue
                for k in list(st.keys()):
                    if k not in keep:
                        del st[k]
                st.update({k: copy.deepcopy(v) for k, v in plain.items()})
                if not st and p in self.opt.state:
                    del self.opt.state[p]
            for p, k, v in self.opt_entries:
                self.opt.state[p][k] = v       # same tensor object, values restored in place above
            for g, saved in zip(self.opt.param_groups, self.opt_groups):
                g.update(copy.deepcopy(saved))


def cache_states(self):
    """Replacement for reference ``ImplicitProblem.cache_states`` (problems/implicit_problem.py:67-70)."""
    self._bb200_snapshot = StateSnapshot(self.module, self.optimizer)
    # the reference's attributes stay meaningful for user code that only tests them for None
    self.module_state_dict_cache = self._bb200_snapshot
    if self.optimizer is not None:
        self.opitmizer_state_dict_cache = self._bb200_snapshot


def recover_states(self, clean: bool = True):
    """Replacement for reference ``ImplicitProblem.recover_states`` (problems/implicit_problem.py:72-78)."""
    self._bb200_snapshot.restore()
    if clean:
        self._bb200_snapshot = None
        self.module_state_dict_cache = None
        self.opitmizer_state_dict_cache = None


def install_callers(betty_module=None):
    """Rebind the three caller-side methods on the reference's classes.  Returns the patched classes."""
    if betty_module is None:
        import betty as betty_module
    import importlib

    problem = importlib.import_module(betty_module.__name__ + ".problems.problem")
    implicit = importlib.import_module(betty_module.__name__ + ".problems.implicit_problem")
    problem.Problem.synchronize_params = synchronize_params
    implicit.ImplicitProblem.cache_states = cache_states
    implicit.ImplicitProblem.recover_states = recover_states
    return problem.Problem, implicit.ImplicitProblem
