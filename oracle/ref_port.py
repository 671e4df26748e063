"""TEST INFRASTRUCTURE -- CPU/GPU restatement ("port") of the reference hypergradient algorithms.

This file is the *oracle*: a torch-autograd restatement of ``betty.hypergradient.{neumann,cg,darts}``.
It is imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (``cpu_baseline`` /
``--impl reference`` legs).  The product path (``betty_b200/``) never imports it.

Pinning: the reference's own tests hold no golden vectors for this path, so the
pin is created by ``oracle/make_golden.py``, which runs the *real* reference functions imported from
``/root/reference`` on seeded workloads and commits their outputs under ``tests/golden/``;
``tests/test_oracle.py`` checks this port against those vectors (and against the fp64 dense
restatement in ``oracle/dense.py``).

Arithmetic lives in a third-party dependency of the reference: PyTorch autograd (``torch>=1.8.0``,
reference ``requirements/requirements.txt:1``; installed here 2.11.0+cu128).
"""
from __future__ import annotations

import warnings
from typing import Callable, List, Sequence

import torch


def _flat(tensors: Sequence[torch.Tensor], scale: float = 1.0) -> torch.Tensor:
    # reference betty/utils.py:117-118 (scale each tensor, then concatenate)
    return torch.cat([scale * t.reshape(-1) for t in tensors])


def lower_gradient(curr):
    """Prologue shared by neumann/cg: lower loss on the last batch and its gradient with a graph
    (reference neumann.py:31-36, cg.py:27-32)."""
    loss = curr.training_step_exec(curr.cur_batch)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.autograd.grad(loss, curr.trainable_parameters(), create_graph=True)


def make_hvp(in_grad, params) -> Callable[[Sequence[torch.Tensor]], List[torch.Tensor]]:
    """H.v by double backward through the retained graph (reference neumann.py:62, cg.py:39-41)."""

    def hvp(direction):
        return list(torch.autograd.grad(in_grad, params, grad_outputs=direction, retain_graph=True))

    return hvp


def neumann_series(v, hvp, iterations: int, alpha: float):
    """alpha * sum_{j<=K} (I - alpha H)^j v  (reference neumann.py:59-66)."""
    acc = v
    for _ in range(iterations):
        hv = hvp(v)
        v = [vi - alpha * hi for vi, hi in zip(v, hv)]
        acc = [vi + ai for vi, ai in zip(v, acc)]
    return [alpha * ai for ai in acc]


def cg_solve(v, hvp, iterations: int, cg_alpha: float):
    """K conjugate-gradient steps exactly as the reference writes them (cg.py:34-56), including the
    ``cg_alpha`` asymmetry: the step-size denominator uses the *scaled* H.p while the residual update
    uses the unscaled one."""
    x = [torch.zeros_like(t) for t in v]
    r = [t.clone() for t in v]
    p = [t.clone() for t in r]
    for _ in range(iterations):
        hp = hvp(p)
        rr = torch.dot(_flat(r), _flat(r))
        step = rr / torch.dot(_flat(hp, cg_alpha), _flat(p))
        x = [xi + step * pi for xi, pi in zip(x, p)]
        r = [ri - step * hi for ri, hi in zip(r, hp)]
        rr_new = torch.dot(_flat(r), _flat(r))
        p = [ri + (rr_new / rr) * pi for ri, pi in zip(r, p)]
    return [cg_alpha * xi for xi in x]


def mixed_product(in_grad, prev, x, sync: bool):
    """-(d^2 L_in / d lambda d w)^T x  (reference neumann.py:44-54, cg.py:58-68)."""
    if sync:
        torch.autograd.backward(in_grad, inputs=prev.trainable_parameters(), grad_tensors=[-xi for xi in x])
        return None
    out = torch.autograd.grad(in_grad, prev.trainable_parameters(), grad_outputs=x)
    return [-g for g in out]


def neumann(vector, curr, prev, sync):
    """Reference ``betty/hypergradient/neumann.py:8-56``."""
    assert len(curr.paths) == 0, "neumann method is not supported for higher order MLO!"
    in_grad = lower_gradient(curr)
    x = neumann_series(list(vector), make_hvp(in_grad, curr.trainable_parameters()),
                       curr.config.neumann_iterations, curr.config.neumann_alpha)
    return mixed_product(in_grad, prev, x, sync)


def cg(vector, curr, prev, sync):
    """Reference ``betty/hypergradient/cg.py:8-70``."""
    assert len(curr.paths) == 0, "cg method is not supported for higher order MLO!"
    in_grad = lower_gradient(curr)
    x = cg_solve(list(vector), make_hvp(in_grad, curr.parameters()), curr.config.cg_iterations,
                 curr.config.cg_alpha)
    return mixed_product(in_grad, prev, x, sync)


def _finite_difference(vector, curr, prev, sync, radius, multitask):
    """Shared body of darts.py:27-67 and sama.py:26-59 (non-FSDP branch): central difference of grad_lambda L_in
    along ``vector``, eps = radius / ||vector||."""
    lam = prev.trainable_parameters()
    w = curr.meta_trainable_parameters()
    eps = radius / (_flat(vector).norm() + 1e-15).item()

    def grad_lambda(loss):
        g = torch.autograd.grad(loss, lam, allow_unused=True)
        return [torch.zeros_like(p) if gi is None else gi for gi, p in zip(g, lam)]

    with torch.no_grad():
        for p, vi in zip(w, vector):
            p.add_(vi, alpha=eps)
    g_plus = grad_lambda(curr.training_step_exec(curr.cur_batch))
    if sync:
        prev.set_grads(lam, [-(g / (2 * eps)) for g in g_plus])
    with torch.no_grad():
        for p, vi in zip(w, vector):
            p.sub_(vi, alpha=2 * eps)
    loss_minus = curr.training_step_exec(curr.cur_batch)
    if sync:
        torch.autograd.backward(loss_minus / (2 * eps), inputs=lam)
        g_minus = None
    else:
        g_minus = grad_lambda(loss_minus)
    if not multitask:
        with torch.no_grad():
            for p, vi in zip(w, vector):
                p.add_(vi, alpha=eps)
    if sync:
        return None
    return [(gm - gp) / (2 * eps) for gm, gp in zip(g_minus, g_plus)]


def darts(vector, curr, prev, sync):
    """Central finite difference of grad_lambda L_in along v, eps = darts_alpha/||v||
    (reference ``betty/hypergradient/darts.py:8-69``, non-FSDP branch)."""
    cfg = curr.config
    return _finite_difference(vector, curr, prev, sync, cfg.darts_alpha, cfg.darts_multitask)


def precondition(vector, problem):
    """reference betty/hypergradient/utils.py:24-97: identity for SGD; for Adam the derivative of the update w.r.t.
    the last gradient, from the state before the last step."""
    name = type(problem.optimizer).__name__.lower()
    if "adam" not in name:
        if "rmsprop" in name:
            raise NotImplementedError("SAMA preconditioning for RMSProp is not implemented!")
        return list(vector)
    out = []
    for v, p in zip(vector, problem.meta_trainable_parameters()):
        group = problem.get_opt_param_group_for_param(p)
        state = problem.get_opt_state_for_param(p)
        b1, b2 = group["betas"]
        zeros = torch.zeros_like(v)
        g, m, s = state.get("last_grad", zeros), state.get("exp_avg", zeros), state.get("exp_avg_sq", zeros)
        m_old = (m - (1 - b1) * g) / b1 if b1 != 0 else 0                  # utils.py:51-53
        s_old = (s - (1 - b2) * g * g) / b2                                # utils.py:54
        scale = ((1 - b1) * b2 * s_old - b1 * (1 - b2) * g * m_old) / (torch.sqrt(s) + group["eps"]) ** 3
        out.append(v * scale * group["lr"])
    return out


def sama(vector, curr, prev, sync):
    """reference betty/hypergradient/sama.py:7-61: the finite difference of `darts` along the preconditioned
    direction, radius ``sama_adam_alpha``."""
    cfg = curr.config
    v = precondition(vector, curr)                                         # sama.py:24
    return _finite_difference(v, curr, prev, sync, cfg.sama_adam_alpha, cfg.sama_multitask)


METHODS = {"neumann": neumann, "cg": cg, "darts": darts, "finite_diff": darts, "sama": sama}


def k_loop_only(method: str, vector, curr):
    """Just the K-loop (what the HVP-iters/s metric times): returns the approximate H^-1 v."""
    in_grad = lower_gradient(curr)
    hvp = make_hvp(in_grad, curr.trainable_parameters())
    if method == "neumann":
        return neumann_series(list(vector), hvp, curr.config.neumann_iterations, curr.config.neumann_alpha)
    return cg_solve(list(vector), hvp, curr.config.cg_iterations, curr.config.cg_alpha)
