"""``darts`` / ``finite_diff`` plugin -- drop-in for reference ``betty/hypergradient/darts.py:8-69`` -- and the
finite-difference body it shares with ``sama`` (reference ``betty/hypergradient/sama.py:7-61``).

Central finite difference of grad_lambda L_in along v.  The two lower forward/backward passes stay on
PyTorch (they go through user code and the *upper* module); what moves to CUDA kernels (K4) is the
norm, the three parameter sweeps ``w += eps v``, ``w -= 2 eps v``, ``w += eps v`` and the final
``(g- - g+)/(2 eps)`` -- with eps kept on the device, so the reference's ``.item()`` host sync
(darts.py:35) disappears and the whole call is stream-ordered.
"""
import torch

from .. import _native as N
from ..arena import ChunkTable, as_f32_contig, stream_ptr
from ..engine import Workspace


def _table(a_list, b_list, device):
    return ChunkTable([t.data_ptr() for t in a_list], [t.data_ptr() for t in b_list],
                      [t.numel() for t in a_list], device, keep=(a_list, b_list))


def _grad_lambda(loss, lam):
    g = torch.autograd.grad(loss, lam, allow_unused=True)
    # reference betty/utils.py:132-137
    out = [torch.zeros_like(p) if gi is None else gi.contiguous() for gi, p in zip(g, lam)]
    for gi in out:
        if gi.dtype != torch.float32:
            raise N.NativeError("betty_b200.darts: upper gradient of dtype %s (fp32 expected)" % gi.dtype)
    return out


def finite_difference(v, curr, prev, sync, radius: float, restore):
    """``(grad_lambda L(w - eps v) - grad_lambda L(w + eps v)) / (2 eps)``, eps = radius / ||v||  (darts.py:27-67 =
    sama.py:26-59).  ``v``: contiguous fp32 tensors; ``restore``: callable run at the end instead of ``w += eps v``
    (multitask modes), or None."""
    w = curr.meta_trainable_parameters()
    lam = prev.trainable_parameters()
    dev = w[0].device
    for p in w:
        if p.dtype != torch.float32 or not p.is_contiguous():
            raise N.NativeError("betty_b200 finite difference needs contiguous fp32 lower parameters")
    for p in lam:
        # the K4 chunk tables address the upper gradients as fp32 words
        if p.dtype != torch.float32:
            raise N.NativeError("betty_b200 finite difference needs fp32 upper parameters (got %s)" % p.dtype)
    with torch.cuda.device(dev):
        ws = Workspace.get(dev)
        s = stream_ptr()
        tab = _table(v, [p.data for p in w], dev)          # a = v_i, b = w_i
        N.call("bb_mt_sumsq", tab.ptr, tab.n, ws.ptr, s)   # ||v||^2          (darts.py:30)
        N.call("bb_fd_eps", ws.ptr, float(radius), s)      # eps, 1/(2 eps)   (darts.py:35)
        eps_ptr = ws.scalar_ptr(6)
        inv_2eps = ws.scalars[7]                           # 0-dim device tensor view, no host sync

        N.call("bb_mt_axpby", tab.ptr, tab.n, 1.0, eps_ptr, 1.0, s)       # w += eps v      (darts.py:37-38)
        g_plus = _grad_lambda(curr.training_step_exec(curr.cur_batch), lam)
        if sync:
            tp = _table(g_plus, g_plus, dev)
            N.call("bb_mt_axpby", tp.ptr, tp.n, -1.0, ws.scalar_ptr(7), 0.0, s)  # -g+/(2 eps)  (darts.py:44-45)
            prev.set_grads(lam, g_plus)
        N.call("bb_mt_axpby", tab.ptr, tab.n, -2.0, eps_ptr, 1.0, s)      # w -= 2 eps v    (darts.py:49-50)
        loss_n = curr.training_step_exec(curr.cur_batch)
        out = None
        if sync:
            torch.autograd.backward(loss_n * inv_2eps.to(loss_n.dtype), inputs=lam)   # (darts.py:52-53)
        else:
            g_minus = _grad_lambda(loss_n, lam)
            g_minus = [g.clone() if g.data_ptr() == gp.data_ptr() else g for g, gp in zip(g_minus, g_plus)]
            tc = _table(g_minus, g_plus, dev)
            N.call("bb_mt_fd_combine", tc.ptr, tc.n, ws.ptr, s)          # (g- - g+)/(2 eps) (darts.py:65-67)
            out = g_minus
        if restore is None:
            N.call("bb_mt_axpby", tab.ptr, tab.n, 1.0, eps_ptr, 1.0, s)  # restore w        (darts.py:61-63)
        else:
            restore()
    return out


def darts(vector, curr, prev, sync):
    N.require_cuda()
    cfg = curr.config
    if getattr(curr, "_strategy", None) == "fsdp":
        raise NotImplementedError("betty_b200.darts: FSDP-sharded parameters are out of scope (SURVEY.md §2b)")
    restore = (lambda: None) if cfg.darts_multitask else None            # darts.py:60: parameters stay perturbed
    return finite_difference(as_f32_contig(vector), curr, prev, sync, float(cfg.darts_alpha), restore)
