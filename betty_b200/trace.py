"""Record the lower problem's forward as an aten-level op tape.

The drop-in boundary hands the engine an opaque Python closure (``curr.training_step_exec``), not a
layer list (SURVEY.md §7 hard part 1).  To run the Hessian-vector products with hand-written kernels
the engine needs the structure, so the prologue's forward -- which the reference also runs once per
call (neumann.py:31, cg.py:27) -- is executed under a ``TorchDispatchMode`` that logs every aten op
with its actual input/output tensors.  PyTorch still executes the forward itself (it goes through
user code and the upper module); ``lower.py`` then turns the tape into the second-order plan.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, List, Sequence, Tuple

import torch
from torch.utils._python_dispatch import TorchDispatchMode


_NAMES: dict = {}   # OpOverload -> "aten.addmm.default" (str() of an overload costs ~3 us; asked ~1.3 times per op)


@dataclass
class TapeOp:
    func: Any           # torch._ops.OpOverload
    args: tuple
    kwargs: dict
    out: Any            # Tensor | tuple | list

    @property
    def name(self) -> str:
        n = _NAMES.get(self.func)
        if n is None:
            n = _NAMES[self.func] = str(self.func)  # e.g. "aten.addmm.default"
        return n


@dataclass
class Tape:
    ops: List[TapeOp] = field(default_factory=list)
    params: List[torch.Tensor] = field(default_factory=list)
    loss: torch.Tensor = None


class _Recorder(TorchDispatchMode):
    def __init__(self, tape: Tape):
        super().__init__()
        self.tape = tape

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        out = func(*args, **kwargs)
        # keeping args/out alive keeps id() stable and the base activations resident for the K-loop
        self.tape.ops.append(TapeOp(func, args, kwargs, out))
        return out


def record_tape(fn: Callable[[], Any], params: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, Tape]:
    """Run ``fn`` (the lower ``training_step_exec``) and return ``(loss, tape)``.  The autograd graph
    above the dispatch layer is built as usual, so the caller can still differentiate the loss."""
    tape = Tape(params=list(params))
    with _Recorder(tape):
        out = fn()
    loss = out["loss"] if isinstance(out, dict) else (out[0] if isinstance(out, (tuple, list)) else out)
    tape.loss = loss
    return loss, tape
