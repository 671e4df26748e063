"""TEST INFRASTRUCTURE -- the *unmodified* reference, importable wherever ``oracle/_ref`` exists.

``oracle/fetch_ref.sh`` mirrors ``/root/reference/{betty, test/test_regression.py,
examples/neural_architecture_search}`` into the git-ignored ``oracle/_ref/`` (the reference is pure Python, so
the mirror *is* the build); the directory travels to the GPU box with the snapshot.  This module

  * imports the real ``betty`` from there (``load()``),
  * wraps a synthetic workload (``betty_b200.workloads``) into real ``betty.problems.ImplicitProblem`` objects
    wired by a real ``betty.engine.Engine`` (``real_problems``), following the recipe of SURVEY.md Appendix A,
  * exposes the reference's own K-loops for the CPU baseline of ``bench.py`` (``kloop_seconds``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s reference / cpu_baseline legs may import it.
"""
from __future__ import annotations

import importlib
import os
import sys
import time
import warnings
from typing import Callable, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(_HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "betty", "hypergradient"))


def load():
    """The real ``betty`` package (v0.2.1 mirror).  Raises when ``oracle/_ref`` has not been fetched."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `bash oracle/fetch_ref.sh` where /root/reference exists "
                           "(python -c 'import __graft_entry__ as g; g.build()' does it)")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        betty = importlib.import_module("betty")
        importlib.import_module("betty.hypergradient")
        importlib.import_module("betty.engine")
    assert os.path.abspath(betty.__file__).startswith(REF_ROOT), betty.__file__
    return betty


def reference_table():
    """A *copy* of the reference's own plugin table (so a test can compare against it after ``install()``)."""
    load()
    import betty.hypergradient as RH

    return dict(RH.jvp_fn_mapping)


def real_problems(wl, upper_step: Optional[Callable] = None, strategy: str = "default", lower_config=None,
                  upper_batch=None, lower_optimizer=None, upper_optimizer=None):
    """Real ``ImplicitProblem`` pair + ``Engine`` around a workload's modules, data and loss closure.

    The workload's lower ``training_step`` closure reaches the upper module through ``p.peers['upper'].module``;
    the real problems get the same ``peers`` attribute, so the *same closure* runs inside the reference.
    Returns ``(engine, upper_problem, lower_problem)``."""
    load()
    from betty.configs import Config, EngineConfig
    from betty.engine import Engine
    from betty.problems import ImplicitProblem

    lower_step = wl.lower._training_step
    ucfg, lcfg = Config(), lower_config
    if lcfg is None:
        sc = wl.lower.config
        lcfg = Config(type=sc.type, precision=sc.precision, darts_alpha=sc.darts_alpha,
                      darts_multitask=sc.darts_multitask, neumann_iterations=sc.neumann_iterations,
                      neumann_alpha=sc.neumann_alpha, cg_iterations=sc.cg_iterations, cg_alpha=sc.cg_alpha,
                      unroll_steps=1)
        for k in ("sama_adam_alpha", "sama_multitask"):
            if hasattr(sc, k):
                setattr(lcfg, k, getattr(sc, k))

    class Lower(ImplicitProblem):
        def training_step(self, batch):
            return lower_step(self, batch)

    class Upper(ImplicitProblem):
        def training_step(self, batch):
            return upper_step(self, batch)

    lmod, umod = wl.lower.module, wl.upper.module
    lopt = lower_optimizer or torch.optim.SGD(lmod.parameters(), lr=0.1)
    uopt = upper_optimizer or torch.optim.SGD(umod.parameters(), lr=0.1)
    lbatch = tuple(b.cpu() if torch.is_tensor(b) else b for b in wl.lower.cur_batch)
    ubatch = tuple(b.cpu() if torch.is_tensor(b) else b for b in (upper_batch or wl.lower.cur_batch))
    lower = Lower(name="lower", module=lmod, optimizer=lopt, train_data_loader=[lbatch], config=lcfg)
    upper = Upper(name="upper", module=umod, optimizer=uopt, train_data_loader=[ubatch], config=ucfg)
    lower.peers, upper.peers = {"upper": upper}, {"lower": lower}
    deps = {"l2u": {lower: [upper]}, "u2l": {upper: [lower]}}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        engine = Engine(config=EngineConfig(train_iters=1, strategy=strategy), problems=[upper, lower],
                        dependencies=deps)
    lower.cur_batch = lower.get_batch()
    return engine, upper, lower


def hypergradient_through_reference(upper, lower, retain_graph: bool = False, do_sync: bool = False):
    """``betty.hypergradient.get_grads`` on the upper problem's loss -- whatever table is bound right now."""
    import betty.hypergradient as RH

    loss = upper.training_step_exec(upper.get_batch())
    return RH.get_grads(loss, upper.paths[0], retain_graph, do_sync)


# ---- the reference's own K-loops, for bench.py's reference arm / cpu_baseline ------------------------------
def kloop_seconds(wl, method: str, iterations: int) -> dict:
    """Wall time of ``iterations`` K-loop iterations of the real reference on ``wl`` (CPU tensors).

    neumann: ``betty.hypergradient.neumann.approx_inverse_hvp`` is the K-loop itself (neumann.py:59-66).
    cg:      the loop is inline in ``cg()`` (cg.py:34-56), so the whole call is timed at ``cg_iterations`` = 1 and
             = 1 + ``iterations``; the difference is ``iterations`` K-loop iterations (prologue/epilogue cancel).
    darts:   one call = one iter-equivalent."""
    load()
    # the functions themselves, not the (possibly rebound) plugin table
    approx_inverse_hvp = importlib.import_module("betty.hypergradient.neumann").approx_inverse_hvp
    fns = {"cg": importlib.import_module("betty.hypergradient.cg").cg,
           "darts": importlib.import_module("betty.hypergradient.darts").darts,
           "sama": importlib.import_module("betty.hypergradient.sama").sama}

    lower, upper = wl.lower, wl.upper
    vec = list(wl.vector)
    if method == "neumann":
        in_loss = lower.training_step_exec(lower.cur_batch)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            in_grad = torch.autograd.grad(in_loss, lower.trainable_parameters(), create_graph=True)
        t0 = time.perf_counter()
        approx_inverse_hvp(vec, in_grad, lower.trainable_parameters(), iterations=iterations,
                           alpha=lower.config.neumann_alpha)
        return {"kloop_s": time.perf_counter() - t0, "call_s": None}
    if method == "cg":
        keep = lower.config.cg_iterations

        def call(k):
            lower.config.cg_iterations = k
            t0 = time.perf_counter()
            fns["cg"](vec, lower, upper, False)
            return time.perf_counter() - t0

        try:
            t_one = call(1)                    # prologue + 1 iteration + epilogue
            t_full = call(1 + iterations)      # ... + `iterations` more K-loop iterations
        finally:
            lower.config.cg_iterations = keep
        return {"kloop_s": max(t_full - t_one, 1e-9), "call_s": t_full}
    t0 = time.perf_counter()
    fns[method](vec, lower, upper, False)
    dt = time.perf_counter() - t0
    return {"kloop_s": dt, "call_s": dt}
