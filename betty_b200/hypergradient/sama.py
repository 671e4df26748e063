"""``sama`` plugin -- drop-in for reference ``betty/hypergradient/sama.py:7-61``.

SAMA = the finite-difference hypergradient of ``darts`` taken along the direction preconditioned by the lower
optimizer's adaptive state (reference ``betty/hypergradient/utils.py:24-97``): identity for SGD, for Adam

    v <- v * lr * ((1-b1) b2 s_old - b1 (1-b2) g m_old) / (sqrt(s) + eps)^3

with ``m_old, s_old`` the moment estimates before the last step.  Here the preconditioner is ONE multi-tensor K4
kernel over every parameter (``bb_mt_adam_precondition``) instead of ~12 element-wise launches and 8 temporaries per
parameter tensor, and it feeds the same device-resident-eps finite-difference body as ``darts``.
"""
import ctypes as C

import numpy as np
import torch

from .. import _native as N
from ..arena import MT_CHUNK, as_f32_contig, stream_ptr
from .darts import finite_difference

_ADAM_CHUNK = np.dtype([("v", np.uint64), ("g", np.uint64), ("m", np.uint64), ("s", np.uint64), ("out", np.uint64),
                        ("n", np.int32), ("beta1", np.float32), ("beta2", np.float32), ("eps", np.float32),
                        ("lr", np.float32), ("pad", np.int32)], align=True)


def _optimizer_kind(optimizer) -> str:
    # reference hypergradient/utils.py:24-30
    name = type(optimizer).__name__.lower()
    if "adam" in name:
        return "adam"
    if "rmsprop" in name:
        return "rmsprop"
    return "sgd"


def _state_tensor(state, key, like):
    t = state.get(key)
    if t is None:
        return None, 0                                   # the reference substitutes zeros (utils.py:48-50)
    if t.dtype != torch.float32 or not t.is_contiguous() or t.shape != like.shape:
        t = t.detach().to(torch.float32).contiguous()
    return t, t.data_ptr()


def precondition(vector, problem):
    """reference hypergradient/utils.py:90-97 (dispatch) and :37-63 (Adam)."""
    kind = _optimizer_kind(problem.optimizer)
    if kind == "sgd":
        return list(vector)
    if kind != "adam":
        raise NotImplementedError(f"SAMA preconditioning for {kind} is not implemented!")     # as the reference
    params = problem.meta_trainable_parameters()
    dev = params[0].device
    rows, keep, outs = [], [], []
    for v, p in zip(vector, params):
        group = problem.get_opt_param_group_for_param(p)
        state = problem.get_opt_state_for_param(p)
        b1, b2 = group["betas"]
        out = torch.empty_like(v)
        g, gp = _state_tensor(state, "last_grad", v)
        m, mp = _state_tensor(state, "exp_avg", v)
        q, qp = _state_tensor(state, "exp_avg_sq", v)
        keep += [g, m, q]
        outs.append(out)
        n, done = v.numel(), 0
        while done < n:
            k = min(MT_CHUNK, n - done)
            off = 4 * done
            rows.append((v.data_ptr() + off, gp + off if gp else 0, mp + off if mp else 0, qp + off if qp else 0,
                         out.data_ptr() + off, k, float(b1), float(b2), float(group["eps"]), float(group["lr"]), 0))
            done += k
    arr = np.array(rows, dtype=_ADAM_CHUNK)
    tab = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(dev)
    with torch.cuda.device(dev):
        N.call("bb_mt_adam_precondition", tab.data_ptr(), len(rows), stream_ptr())
    del keep
    return outs


def sama(vector, curr, prev, sync):
    N.require_cuda()
    cfg = curr.config
    v = precondition(as_f32_contig(vector), curr)                         # sama.py:24
    restore = None
    if cfg.sama_multitask:                                                # sama.py:52-55
        restore = lambda: curr.synchronize_params(curr.meta_trainable_parameters(), all_reduce=True)
    return finite_difference(v, curr, prev, sync, float(cfg.sama_adam_alpha), restore)
