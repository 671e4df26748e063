"""Host-side mirror of the slice of the reference ``Problem`` interface that the
hypergradient plugins touch.

The plugins in ``betty_b200.hypergradient`` are duck-typed exactly like the reference's
(``fn(vector, curr, prev, sync)``, reference ``betty/hypergradient/__init__.py:33-37``), so a real
``betty.problems.ImplicitProblem`` works unchanged.  On a machine without the reference installed
(the GPU box), tests and ``bench.py`` drive the plugins through these stand-ins, which expose the
attributes listed in SURVEY.md §8(a) row a9:

  ``.config``                      reference ``betty/configs/problem_dataclass.py:4-48``
  ``.cur_batch``                   reference ``betty/problems/problem.py:334-345``
  ``.training_step_exec(batch)``   reference ``betty/problems/problem.py:327-332`` (autocast wrapper)
  ``.trainable_parameters()`` / ``.parameters()`` / ``.meta_trainable_parameters()``
                                   reference ``betty/problems/implicit_problem.py:80-84``,
                                   ``betty/problems/problem.py:850-854``
  ``.paths`` / ``._strategy``      reference ``betty/problems/problem.py`` (asserted in neumann.py:29)
  ``.set_grads(params, grads)``    reference ``betty/problems/problem.py:583-597``
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Sequence

import torch


@dataclass
class ShimConfig:
    """Same field names and defaults as the reference ``Config`` dataclass (only the knobs the
    hot path reads; reference ``betty/configs/problem_dataclass.py:10-48``)."""

    type: str = "darts"
    precision: str = "fp32"
    darts_alpha: float = 0.01
    darts_multitask: bool = False
    sama_adam_alpha: float = 1.0
    sama_multitask: bool = False
    neumann_iterations: int = 1
    neumann_alpha: float = 1.0
    cg_iterations: int = 1
    cg_alpha: float = 1.0
    retain_graph: bool = False
    allow_unused: bool = True


_DTYPES = {"fp16": torch.float16, "bf16": torch.bfloat16}


class ShimProblem:
    """Minimal problem: a module, a user-supplied ``training_step`` closure, and the bookkeeping
    attributes the plugins read.  ``training_step(problem, batch)`` may reach the *other* problem
    through ``problem.peers[name]`` the way reference problems reach each other by attribute
    (``engine.py:303-328``)."""

    def __init__(
        self,
        name: str,
        module: torch.nn.Module,
        training_step: Callable[["ShimProblem", Any], torch.Tensor],
        config: Optional[ShimConfig] = None,
    ):
        self.name = name
        self.module = module
        self._training_step = training_step
        self.config = config or ShimConfig()
        self.cur_batch: Any = None
        self.paths: List[Any] = []
        self._strategy = "default"
        self.peers: dict = {}
        self.optimizer = None           # `sama` reads the lower optimizer's state (reference utils.py:37-63)

    # -- reference problem.py:320-332 -------------------------------------------------------
    def training_step(self, batch):
        return self._training_step(self, batch)

    def training_step_exec(self, batch):
        prec = self.config.precision
        if prec in _DTYPES and torch.cuda.is_available():
            with torch.autocast("cuda", dtype=_DTYPES[prec]):
                return self.training_step(batch)
        return self.training_step(batch)

    def __call__(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # -- reference implicit_problem.py:80-84, problem.py:850-854 ----------------------------
    def parameters(self):
        return list(self.module.parameters())

    def trainable_parameters(self):
        return list(self.module.parameters())

    def meta_trainable_parameters(self):
        return self.trainable_parameters()

    # -- reference problem.py:583-597 -------------------------------------------------------
    def set_grads(self, params: Sequence[torch.Tensor], grads: Sequence[Optional[torch.Tensor]]):
        for p, g in zip(params, grads):
            if g is None:
                continue
            p.grad = g if p.grad is None else p.grad + g

    # -- reference problem.py:697-723 (used by the SAMA preconditioner) ----------------------
    def get_opt_param_group_for_param(self, param):
        for group in self.optimizer.param_groups:
            for p in group["params"]:
                if param is p:
                    return group

    def get_opt_state_for_param(self, param):
        return self.optimizer.state[param]

    def synchronize_params(self, params, all_reduce=False):
        # reference problem.py:599-610: a no-op on one process
        return None

    def zero_grad(self):
        for p in self.trainable_parameters():
            p.grad = None


def upper_direct_gradient(upper_loss: torch.Tensor, lower: ShimProblem):
    """``v = dL_upper/dw`` with ``None`` replaced by zeros -- what the reference computes before it
    dispatches to the plugin (``betty/hypergradient/__init__.py:24-31``)."""
    params = lower.meta_trainable_parameters()
    grads = torch.autograd.grad(upper_loss, params, retain_graph=True, allow_unused=True)
    return tuple(torch.zeros_like(p) if g is None else g for g, p in zip(grads, params))
