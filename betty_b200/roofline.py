"""Algorithmic bytes of one K-loop iteration, by the formula of SURVEY.md 8(d) -- the denominator-free part of
the roofline the bench reports (``roofline.iteration``).  Independent of how the plan executes:

    per parametric layer l:  2 * s_w * P_l                      (W read in the R-forward and the R-backward;
                                                                 0 for an Embedding, which is linear in W)
                             + s_a * (5 * A_in_l + 3 * A_out_l) (R-fwd reads x, x' writes y'; R-bwd reads
                                                                 x, x', d, d' writes d'_x)
    vectors:                 20 * P (Neumann)  |  52 * P (CG)    fp32 v / p / x / r / H.p passes

``s_w`` = bytes per weight element as streamed, ``s_a`` = bytes per activation element: 2 for a graph recorded
under bf16 / fp16 autocast, 4 for fp32.  Parametric layers = Linear (gemm with a parameter operand), Conv2d,
BatchNorm, LayerNorm, Embedding; non-parametric ops count 0 (they are expected to be fused).
"""
from __future__ import annotations

import torch

_PARAMETRIC = ("gemm", "conv2d", "batchnorm", "layernorm", "embedding", "convblock", "convblock2")


def _is_param(v) -> bool:
    return v is not None and v.root.param_index is not None


def survey_bytes(graph, method: str) -> dict:
    """``graph``: the lowered IR (ir.Graph, *before or after* fusion -- fused nodes carry their members in
    ``attrs['members']``).  Returns the byte count and its terms."""
    reduced = False
    w_bytes = a_in = a_out = 0
    P = sum(p.base.numel() for p in graph.params)

    def visit(n):
        nonlocal reduced, w_bytes, a_in, a_out
        for m in n.attrs.get("members", ()):
            visit(m)
        if n.op not in _PARAMETRIC or n.op in ("convblock", "convblock2"):
            return
        pl = [v for v in n.ins if _is_param(v)]
        if not pl:
            return      # attention bmm etc.: no parameter operand
        acts = [v for v in n.ins if v is not None and not _is_param(v)]
        for key in ("A", "B", "X", "W"):
            t = n.attrs.get(key)
            if isinstance(t, torch.Tensor) and t.dtype in (torch.bfloat16, torch.float16):
                reduced = True
        if n.op != "embedding":
            w_bytes += sum(v.base.numel() for v in pl)
        a_in += sum(v.base.numel() for v in acts)
        if n.op in ("gemm", "conv2d") and not acts:
            # data-input layer: its input is a constant of the K-loop but is still read (x in both sweeps)
            x = n.attrs.get("X") if n.op == "conv2d" else (n.attrs.get("A") if not _is_param(n.ins[0]) else n.attrs.get("B"))
            a_in += x.numel() if isinstance(x, torch.Tensor) else 0
        a_out += n.out.base.numel()

    for n in graph.nodes:
        visit(n)
    s = 2 if reduced else 4
    vec = (20 if method == "neumann" else 52) * P
    total = 2 * s * w_bytes + s * (5 * a_in + 3 * a_out) + vec
    return {"bytes": int(total), "P": int(P), "A_in": int(a_in), "A_out": int(a_out), "s_a": s, "s_w": s,
            "vector_bytes": int(vec), "formula": "sum_l[2 s_w P_l + s_a (5 A_in + 3 A_out)] + (20|52) P  (SURVEY.md 8d)"}
