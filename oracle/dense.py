"""TEST INFRASTRUCTURE -- fp64 dense restatement of the K-loop recurrences for small problems.

Second, independent oracle: with the lower Hessian ``H`` and the mixed
block ``M = d g / d lambda`` formed explicitly (numpy, float64) the Neumann and CG recurrences are
written out literally (reference neumann.py:59-66, cg.py:34-56) and the hypergradient is
``-M^T x``.  Only imported from ``tests/``.
"""
from __future__ import annotations

import numpy as np
import torch


def dense_blocks(lower, upper):
    """Explicit H = d^2 L/dw^2 and M = d(grad_w L)/d lambda in float64 for a *small* problem.
    The modules are deep-copied to double first so the caller's problem is untouched."""
    import copy

    lo = copy.copy(lower)
    up = copy.copy(upper)
    lo.module = copy.deepcopy(lower.module).double()
    up.module = copy.deepcopy(upper.module).double()
    lo.peers = {"upper": up}
    up.peers = {"lower": lo}
    lo.cur_batch = tuple(b.double() if torch.is_tensor(b) and b.is_floating_point() else b for b in lower.cur_batch)
    w = lo.trainable_parameters()
    lam = up.trainable_parameters()
    loss = lo.training_step_exec(lo.cur_batch)
    g = torch.autograd.grad(loss, w, create_graph=True)
    gflat = torch.cat([t.reshape(-1) for t in g])
    P = gflat.numel()
    H = np.zeros((P, P))
    PU = sum(p.numel() for p in lam)
    M = np.zeros((P, PU))
    for i in range(P):
        row_w = torch.autograd.grad(gflat[i], w, retain_graph=True, allow_unused=True)
        H[i] = torch.cat([(torch.zeros_like(p) if r is None else r).reshape(-1) for r, p in zip(row_w, w)]).numpy()
        row_l = torch.autograd.grad(gflat[i], lam, retain_graph=True, allow_unused=True)
        M[i] = torch.cat([(torch.zeros_like(p) if r is None else r).reshape(-1) for r, p in zip(row_l, lam)]).numpy()
    return H, M


def neumann_dense(H, v, K, alpha):
    v = v.astype(np.float64).copy()
    acc = v.copy()
    for _ in range(K):
        v = v - alpha * (H @ v)
        acc = v + acc
    return alpha * acc


def cg_dense(H, v, K, cg_alpha):
    x = np.zeros_like(v, dtype=np.float64)
    r = v.astype(np.float64).copy()
    p = r.copy()
    for _ in range(K):
        hp = H @ p
        rr = r @ r
        step = rr / ((cg_alpha * hp) @ p)
        x = x + step * p
        r = r - step * hp
        p = r + ((r @ r) / rr) * p
    return cg_alpha * x


def hypergradient_dense(M, x):
    return -(M.T @ x)
