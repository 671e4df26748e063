"""TEST INFRASTRUCTURE -- torch interpreter of the second-order tape IR (betty_b200/ir.py).

Executes the same node list the CUDA executor runs, with plain torch ops (any device, float64 by
default), so that (i) the tape lowering and the rule maths can be checked on CPU against autograd's
double backward -- which is what the reference evaluates (neumann.py:62, cg.py:39-41) -- and (ii) each
CUDA kernel can be unit-tested against the rule it implements.  Never imported by the product.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from betty_b200.ir import Graph, Node, Val


class Interp:
    def __init__(self, graph: Graph, dtype=torch.float64):
        self.g, self.dtype = graph, dtype
        for v in graph.values:
            if v.parent is None and (v.needed or v.param_index is not None or v.boundary or getattr(v, "interp_only", False)):
                shape, stride = v.base.shape, v.base.stride()
                mk = lambda: torch.zeros_like(v.base, dtype=dtype)  # preserves dense strides
                v.t, v.a, v.at = mk(), mk(), mk()
                assert v.t.stride() == stride or v.base.numel() <= 1, (v.t.stride(), stride)
        self.saved: Dict[int, dict] = {}

    # ---- buffer access through aliases --------------------------------------------------------
    def buf(self, v: Val, kind: str) -> torch.Tensor:
        if v.parent is None:
            return getattr(v, kind)
        return v.viewfn(self.buf(v.parent, kind))

    def base(self, v_or_t):
        t = v_or_t.base if isinstance(v_or_t, Val) else v_or_t
        return t.detach().to(self.dtype)

    @staticmethod
    def put(dst: torch.Tensor, val: torch.Tensor, beta: int):
        if beta:
            dst.add_(val.reshape(dst.shape) if val.shape != dst.shape else val)
        else:
            dst.copy_(val.reshape(dst.shape) if val.shape != dst.shape else val)

    # ---- passes ---------------------------------------------------------------------------------
    def set_direction(self, vec: List[torch.Tensor]):
        for p, d in zip(self.g.params, vec):
            p.t.copy_(d.to(self.dtype))

    def zero(self, kind: str):
        for v in self.g.values:
            if v.parent is None and v.zero_init and getattr(v, kind) is not None:
                getattr(v, kind).zero_()

    def base_backward(self):
        self.zero("a")
        self.g.loss.root.a.zero_()
        self.buf(self.g.loss, "a").fill_(1.0)   # only the loss's own element of a larger root
        for n in reversed(self.g.nodes):
            getattr(self, "bb_" + n.op)(n)

    def tangent_forward(self):
        for n in self.g.nodes:
            getattr(self, "tf_" + n.op)(n)

    def tangent_backward(self):
        self.zero("at")
        self.g.loss.root.at.zero_()
        for n in reversed(self.g.nodes):
            getattr(self, "tb_" + n.op)(n)

    def hvp(self, vec: List[torch.Tensor]) -> List[torch.Tensor]:
        self.set_direction(vec)
        self.tangent_forward()
        self.tangent_backward()
        return [p.at.clone() for p in self.g.params]

    # ---- convblock: fused conv3x3(data) -> batch_norm -> [relu] -> max_pool2d (ir._fuse_data_conv_block) -------------
    # The executable specification of the fused CUDA node is the composition of its three member rules.
    def tf_convblock(self, n):
        for m in n.attrs["members"]:
            getattr(self, "tf_" + m.op)(m)

    def _sync_member_modes(self, n):
        if n.op == "convblock2":            # the conv member writes x_in's adjoint the way the fused node is told to
            n.attrs["members"][0].beta[0] = n.beta[0]

    def bb_convblock(self, n):
        self._sync_member_modes(n)
        for m in reversed(n.attrs["members"]):
            getattr(self, "bb_" + m.op)(m)

    def tb_convblock(self, n):
        self._sync_member_modes(n)
        for m in reversed(n.attrs["members"]):
            getattr(self, "tb_" + m.op)(m)

    tf_convblock2, bb_convblock2, tb_convblock2 = tf_convblock, bb_convblock, tb_convblock

    # ---- diagshift: folded c*sum((w-const)^2) terms (ir._fold_quadratic_regularisers) -----------------
    def tf_diagshift(self, n):
        pass

    def bb_diagshift(self, n):
        for tgt, x in zip(n.attrs["targets"], n.attrs["xs"]):
            if x is not None:
                tgt.a.add_(n.attrs["coef"] * self.base(x))   # gradient of (coef/2) * sum(x^2) w.r.t. the parameter

    def tb_diagshift(self, n):
        for src, tgt in zip(n.ins, n.attrs["targets"]):
            tgt.at.add_(n.attrs["coef"] * src.t)             # curvature coef*I (target = parameter) / mixed term (theta)

    def mixed_seeds(self, x: List[torch.Tensor]):
        """d(g.x)/dB for every boundary value B: one more tangent forward/backward along x."""
        for b in self.g.boundaries:
            if b.at is not None:
                b.at.zero_()
        self.hvp(x)
        return [(b.base, b.at) for b in self.g.boundaries if b.at is not None]

    # ---- unary ------------------------------------------------------------------------------------
    def _d12(self, n: Node, x: torch.Tensor):
        k = n.attrs["kind"]
        if k == "relu":
            return (x > 0).to(x.dtype), torch.zeros_like(x)
        if k == "tanh":
            y = torch.tanh(x)
            return 1 - y * y, -2 * y * (1 - y * y)
        if k == "sigmoid":
            s = torch.sigmoid(x)
            return s * (1 - s), s * (1 - s) * (1 - 2 * s)
        if k == "gelu":
            if n.attrs.get("approximate", "none") != "none":
                raise NotImplementedError("tanh-approximated gelu")
            pdf = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
            cdf = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
            return cdf + x * pdf, pdf * (2 - x * x)
        if k == "pow":
            e = n.attrs["scalar"]
            d1 = e * x.pow(e - 1)
            d2 = e * (e - 1) * x.pow(e - 2) if e != 1 else torch.zeros_like(x)
            return d1, d2
        if k in ("scale", "neg"):
            c = -1.0 if k == "neg" else n.attrs["scalar"]
            return torch.full_like(x, c), torch.zeros_like(x)
        raise NotImplementedError(k)

    def tf_unary(self, n):
        x = n.ins[0]
        d1, _ = self._d12(n, self.base(x))
        self.buf(n.out, "t").copy_(d1 * self.buf(x, "t"))

    def bb_unary(self, n):
        x = n.ins[0]
        d1, _ = self._d12(n, self.base(x))
        self.put(self.buf(x, "a"), d1 * self.buf(n.out, "a"), n.beta[0])

    def tb_unary(self, n):
        x = n.ins[0]
        d1, d2 = self._d12(n, self.base(x))
        self.put(self.buf(x, "at"), d1 * self.buf(n.out, "at") + d2 * self.buf(x, "t") * self.buf(n.out, "a"), n.beta[0])

    # ---- copy ---------------------------------------------------------------------------------------
    def tf_copy(self, n):
        self.buf(n.out, "t").copy_(self.buf(n.ins[0], "t"))

    def bb_copy(self, n):
        self.put(self.buf(n.ins[0], "a"), self.buf(n.out, "a"), n.beta[0])

    def tb_copy(self, n):
        self.put(self.buf(n.ins[0], "at"), self.buf(n.out, "at"), n.beta[0])

    # ---- add2 / mulc / mul2 -------------------------------------------------------------------------
    def tf_add2(self, n):
        a, b = n.ins
        self.buf(n.out, "t").copy_(n.attrs["sa"] * self.buf(a, "t") + n.attrs["sb"] * self.buf(b, "t"))

    def _bw_add2(self, n, kind):
        a, b = n.ins
        g = self.buf(n.out, kind)
        self.put(self.buf(a, kind), n.attrs["sa"] * g, n.beta[0])
        self.put(self.buf(b, kind), n.attrs["sb"] * g, n.beta[1])

    def bb_add2(self, n):
        self._bw_add2(n, "a")

    def tb_add2(self, n):
        self._bw_add2(n, "at")

    def tf_mulc(self, n):
        self.buf(n.out, "t").copy_(self.base(n.attrs["const"]) * self.buf(n.ins[0], "t"))

    def bb_mulc(self, n):
        self.put(self.buf(n.ins[0], "a"), self.base(n.attrs["const"]) * self.buf(n.out, "a"), n.beta[0])

    def tb_mulc(self, n):
        self.put(self.buf(n.ins[0], "at"), self.base(n.attrs["const"]) * self.buf(n.out, "at"), n.beta[0])

    def tf_mul2(self, n):
        a, b = n.ins
        self.buf(n.out, "t").copy_(self.buf(a, "t") * self.base(b) + self.base(a) * self.buf(b, "t"))

    def bb_mul2(self, n):
        a, b = n.ins
        g = self.buf(n.out, "a")
        va, vb = g * self.base(b), g * self.base(a)
        self.put(self.buf(a, "a"), va, n.beta[0])
        self.put(self.buf(b, "a"), vb, n.beta[1])

    def tb_mul2(self, n):
        a, b = n.ins
        g, gt = self.buf(n.out, "a"), self.buf(n.out, "at")
        va = gt * self.base(b) + g * self.buf(b, "t")
        vb = gt * self.base(a) + g * self.buf(a, "t")
        self.put(self.buf(a, "at"), va, n.beta[0])
        self.put(self.buf(b, "at"), vb, n.beta[1])

    # ---- sumall -------------------------------------------------------------------------------------
    def tf_sumall(self, n):
        self.buf(n.out, "t").copy_((n.attrs["scale"] * self.buf(n.ins[0], "t").sum()).reshape(n.out.shape))

    def _bw_sumall(self, n, kind):
        x = n.ins[0]
        g = self.buf(n.out, kind).reshape(())
        self.put(self.buf(x, kind), (n.attrs["scale"] * g).expand(x.shape), n.beta[0])

    def bb_sumall(self, n):
        self._bw_sumall(n, "a")

    def tb_sumall(self, n):
        self._bw_sumall(n, "at")

    # ---- gemm: C = A.B (+ bias), each of A, B, bias active or constant --------------------------------
    def _ab(self, n, which, kind):
        v = n.ins[0 if which == "A" else 1]
        if kind == "base":
            t = self.base(n.attrs[which])
        else:
            t = self.buf(v, kind) if v is not None else None
        if t is not None and n.attrs["mv"] and which == "B":
            t = t.unsqueeze(-1)
        return t

    def _c(self, n, kind):
        t = self.buf(n.out, kind)
        return t.unsqueeze(-1) if n.attrs["mv"] else t

    def tf_gemm(self, n):
        A, B = self._ab(n, "A", "base"), self._ab(n, "B", "base")
        tA, tB = self._ab(n, "A", "t"), self._ab(n, "B", "t")
        out = torch.zeros_like(self._c(n, "t"))
        if tA is not None:
            out = out + tA @ B
        if tB is not None:
            out = out + A @ tB
        if n.ins[2] is not None:
            out = out + self.buf(n.ins[2], "t")
        self._c(n, "t").copy_(out)

    def bb_gemm(self, n):
        A, B = self._ab(n, "A", "base"), self._ab(n, "B", "base")
        g = self._c(n, "a")
        if n.ins[0] is not None:
            self.put(self._ab(n, "A", "a"), g @ B.transpose(-1, -2), n.beta[0])
        if n.ins[1] is not None:
            self.put(self._ab(n, "B", "a"), A.transpose(-1, -2) @ g, n.beta[1])
        if n.ins[2] is not None:
            self.put(self.buf(n.ins[2], "a"), g.reshape(-1, g.shape[-1]).sum(0), n.beta[2])

    def tb_gemm(self, n):
        A, B = self._ab(n, "A", "base"), self._ab(n, "B", "base")
        tA, tB = self._ab(n, "A", "t"), self._ab(n, "B", "t")
        g, gt = self._c(n, "a"), self._c(n, "at")
        if n.ins[0] is not None:
            val = gt @ B.transpose(-1, -2)
            if tB is not None:
                val = val + g @ tB.transpose(-1, -2)
            self.put(self._ab(n, "A", "at"), val, n.beta[0])
        if n.ins[1] is not None:
            val = A.transpose(-1, -2) @ gt
            if tA is not None:
                val = val + tA.transpose(-1, -2) @ g
            self.put(self._ab(n, "B", "at"), val, n.beta[1])
        if n.ins[2] is not None:
            self.put(self.buf(n.ins[2], "at"), gt.reshape(-1, gt.shape[-1]).sum(0), n.beta[2])

    # ---- conv2d ----------------------------------------------------------------------------------------
    def _conv(self, n, x, w):
        at = n.attrs
        return F.conv2d(x, w, None, at["stride"], at["padding"], at["dilation"], at["groups"])

    def _dgrad(self, n, g, w):
        at = n.attrs
        return torch.nn.grad.conv2d_input(at["X"].shape, w, g, at["stride"], at["padding"], at["dilation"], at["groups"])

    def _wgrad(self, n, g, x):
        at = n.attrs
        return torch.nn.grad.conv2d_weight(x, at["W"].shape, g, at["stride"], at["padding"], at["dilation"], at["groups"])

    def tf_conv2d(self, n):
        x, w, b = n.ins
        X, W = self.base(n.attrs["X"]), self.base(n.attrs["W"])
        out = torch.zeros_like(self.buf(n.out, "t"))
        if x is not None:
            out = out + self._conv(n, self.buf(x, "t"), W)
        if w is not None:
            out = out + self._conv(n, X, self.buf(w, "t"))
        if b is not None:
            out = out + self.buf(b, "t").view(1, -1, 1, 1)
        self.buf(n.out, "t").copy_(out)

    def bb_conv2d(self, n):
        x, w, b = n.ins
        X, W = self.base(n.attrs["X"]), self.base(n.attrs["W"])
        g = self.buf(n.out, "a")
        if x is not None:
            self.put(self.buf(x, "a"), self._dgrad(n, g, W), n.beta[0])
        if w is not None:
            self.put(self.buf(w, "a"), self._wgrad(n, g, X), n.beta[1])
        if b is not None:
            self.put(self.buf(b, "a"), g.sum((0, 2, 3)), n.beta[2])

    def tb_conv2d(self, n):
        x, w, b = n.ins
        X, W = self.base(n.attrs["X"]), self.base(n.attrs["W"])
        g, gt = self.buf(n.out, "a"), self.buf(n.out, "at")
        if x is not None:
            val = self._dgrad(n, gt, W)
            if w is not None:
                val = val + self._dgrad(n, g, self.buf(w, "t"))
            self.put(self.buf(x, "at"), val, n.beta[0])
        if w is not None:
            val = self._wgrad(n, gt, X)
            if x is not None:
                val = val + self._wgrad(n, g, self.buf(x, "t"))
            self.put(self.buf(w, "at"), val, n.beta[1])
        if b is not None:
            self.put(self.buf(b, "at"), gt.sum((0, 2, 3)), n.beta[2])

    # ---- maxpool2d (gather / scatter with the base argmax) ------------------------------------------------
    def tf_maxpool2d(self, n):
        x = n.ins[0]
        idx = n.attrs["indices"]
        t = self.buf(x, "t")
        N, C = t.shape[:2]
        ty = t.reshape(N, C, -1).gather(2, idx.reshape(N, C, -1)).reshape(idx.shape)
        if n.attrs.get("relu"):      # fused ReLU (ir._fuse_relu_maxpool): relu'(x[argmax]) = [pooled output > 0]
            ty = ty * (n.out.base > 0).to(ty.dtype)
        self.buf(n.out, "t").copy_(ty)

    def _bw_pool(self, n, kind):
        x = n.ins[0]
        idx = n.attrs["indices"]
        g = self.buf(n.out, kind)
        if n.attrs.get("relu"):
            g = g * (n.out.base > 0).to(g.dtype)
        N, C = g.shape[:2]
        z = torch.zeros(x.shape, dtype=self.dtype, device=g.device).reshape(N, C, -1)
        z.scatter_add_(2, idx.reshape(N, C, -1), g.reshape(N, C, -1))
        self.put(self.buf(x, kind), z.reshape(x.shape), n.beta[0])

    def bb_maxpool2d(self, n):
        self._bw_pool(n, "a")

    def tb_maxpool2d(self, n):
        self._bw_pool(n, "at")

    # ---- avgpool2d (linear) -------------------------------------------------------------------------------
    def _avg(self, n, t):
        at = n.attrs
        return F.avg_pool2d(t, at["kernel"], at["stride"], at["padding"], False, True, None) * (
            at["kernel"][0] * at["kernel"][1] / at["divisor"])

    def _avg_bw(self, n, g):
        at = n.attrs
        x = n.ins[0]
        ref = torch.zeros(x.shape, dtype=self.dtype, device=g.device)
        out = torch.ops.aten.avg_pool2d_backward(g.contiguous(), ref, list(at["kernel"]), list(at["stride"]),
                                                 list(at["padding"]), False, True, None)
        return out * (at["kernel"][0] * at["kernel"][1] / at["divisor"])

    def tf_avgpool2d(self, n):
        self.buf(n.out, "t").copy_(self._avg(n, self.buf(n.ins[0], "t")))

    def bb_avgpool2d(self, n):
        self.put(self.buf(n.ins[0], "a"), self._avg_bw(n, self.buf(n.out, "a")), n.beta[0])

    def tb_avgpool2d(self, n):
        self.put(self.buf(n.ins[0], "at"), self._avg_bw(n, self.buf(n.out, "at")), n.beta[0])

    # ---- batchnorm (batch statistics) / layernorm: y = gamma * xhat + beta --------------------------------
    def _norm_dims(self, n, x):
        if n.op == "batchnorm":
            return (0, 2, 3), (1, -1, 1, 1)
        return (x.dim() - 1,), None

    def _norm_common(self, n):
        X = self.base(n.attrs["X"])
        dims, pshape = self._norm_dims(n, X)
        mu = X.mean(dims, keepdim=True)
        var = ((X - mu) ** 2).mean(dims, keepdim=True)
        rstd = (var + n.attrs["eps"]).rsqrt()
        xhat = (X - mu) * rstd
        gam = n.attrs["gamma"]
        gamma = self.base(gam) if gam is not None else torch.ones(X.shape[1] if n.op == "batchnorm" else X.shape[-1],
                                                                   dtype=self.dtype, device=X.device)
        gb = gamma.view(pshape) if pshape else gamma
        red = dims if n.op == "batchnorm" else tuple(range(X.dim() - 1))  # dims the parameter grads reduce over
        return X, dims, red, rstd, xhat, gb

    def _tf_norm(self, n):
        x, gv, bv = n.ins
        X, dims, red, rstd, xhat, gb = self._norm_common(n)
        tx = self.buf(x, "t")
        txc = tx - tx.mean(dims, keepdim=True)
        sdot = (xhat * txc).mean(dims, keepdim=True)
        dxhat = (txc - xhat * sdot) * rstd
        out = gb * dxhat
        if gv is not None:
            tg = self.buf(gv, "t")
            out = out + (tg.view(gb.shape) if gb.dim() == X.dim() else tg) * xhat
        if bv is not None:
            tb = self.buf(bv, "t")
            out = out + (tb.view(gb.shape) if gb.dim() == X.dim() else tb)
        self.buf(n.out, "t").copy_(out)

    def _bb_norm(self, n):
        x, gv, bv = n.ins
        X, dims, red, rstd, xhat, gb = self._norm_common(n)
        g = self.buf(n.out, "a")
        gh = g * gb                                   # dL/dxhat
        u = gh - gh.mean(dims, keepdim=True) - xhat * (gh * xhat).mean(dims, keepdim=True)
        self.put(self.buf(x, "a"), rstd * u, n.beta[0])
        if gv is not None:
            self.put(self.buf(gv, "a"), (g * xhat).sum(red), n.beta[1])
        if bv is not None:
            self.put(self.buf(bv, "a"), g.sum(red), n.beta[2])

    def _tb_norm(self, n):
        x, gv, bv = n.ins
        X, dims, red, rstd, xhat, gb = self._norm_common(n)
        g, gt = self.buf(n.out, "a"), self.buf(n.out, "at")
        tx = self.buf(x, "t")
        txc = tx - tx.mean(dims, keepdim=True)
        sdot = (xhat * txc).mean(dims, keepdim=True)          # = sigma_dot
        dxhat = (txc - xhat * sdot) * rstd
        drstd = -rstd * rstd * sdot                            # d(1/sigma) = -sigma_dot / sigma^2
        gh = g * gb
        tg = None
        if gv is not None:
            tg = self.buf(gv, "t")
            tg = tg.view(gb.shape) if gb.dim() == X.dim() else tg
        ght = gt * gb + (g * tg if tg is not None else 0)     # tangent of dL/dxhat
        m1, m2 = gh.mean(dims, keepdim=True), (gh * xhat).mean(dims, keepdim=True)
        u = gh - m1 - xhat * m2
        ut = ght - ght.mean(dims, keepdim=True) - dxhat * m2 - xhat * (ght * xhat + gh * dxhat).mean(dims, keepdim=True)
        self.put(self.buf(x, "at"), drstd * u + rstd * ut, n.beta[0])
        if gv is not None:
            self.put(self.buf(gv, "at"), (gt * xhat + g * dxhat).sum(red), n.beta[1])
        if bv is not None:
            self.put(self.buf(bv, "at"), gt.sum(red), n.beta[2])

    tf_batchnorm = tf_layernorm = _tf_norm
    bb_batchnorm = bb_layernorm = _bb_norm
    tb_batchnorm = tb_layernorm = _tb_norm

    # ---- softmax / log-softmax over the last dim -----------------------------------------------------
    def tf_softmax(self, n):
        x = n.ins[0]
        p = torch.softmax(self.base(x), -1)
        tz = self.buf(x, "t")
        self.buf(n.out, "t").copy_(p * (tz - (p * tz).sum(-1, keepdim=True)))

    def bb_softmax(self, n):
        x = n.ins[0]
        p = torch.softmax(self.base(x), -1)
        g = self.buf(n.out, "a")
        self.put(self.buf(x, "a"), p * (g - (p * g).sum(-1, keepdim=True)), n.beta[0])

    def tb_softmax(self, n):
        x = n.ins[0]
        p = torch.softmax(self.base(x), -1)
        tz = self.buf(x, "t")
        pt = p * (tz - (p * tz).sum(-1, keepdim=True))
        g, gt = self.buf(n.out, "a"), self.buf(n.out, "at")
        val = pt * (g - (p * g).sum(-1, keepdim=True)) + p * (gt - (pt * g + p * gt).sum(-1, keepdim=True))
        self.put(self.buf(x, "at"), val, n.beta[0])

    def tf_logsoftmax(self, n):
        x = n.ins[0]
        p = torch.softmax(self.base(x), -1)
        tz = self.buf(x, "t")
        self.buf(n.out, "t").copy_(tz - (p * tz).sum(-1, keepdim=True))

    def bb_logsoftmax(self, n):
        x = n.ins[0]
        p = torch.softmax(self.base(x), -1)
        g = self.buf(n.out, "a")
        self.put(self.buf(x, "a"), g - p * g.sum(-1, keepdim=True), n.beta[0])

    def tb_logsoftmax(self, n):
        x = n.ins[0]
        p = torch.softmax(self.base(x), -1)
        tz = self.buf(x, "t")
        pt = p * (tz - (p * tz).sum(-1, keepdim=True))
        g, gt = self.buf(n.out, "a"), self.buf(n.out, "at")
        self.put(self.buf(x, "at"), gt - p * gt.sum(-1, keepdim=True) - pt * g.sum(-1, keepdim=True), n.beta[0])

    # ---- nll (linear in its input) ---------------------------------------------------------------------
    def tf_nll(self, n):
        x = n.ins[0]
        tz = self.buf(x, "t")
        picked = -tz.gather(1, n.attrs["target"].view(-1, 1)).squeeze(1)
        if n.attrs["reduction"] == 0:
            self.buf(n.out, "t").copy_(picked)
        else:
            self.buf(n.out, "t").copy_((n.attrs["scale"] * picked.sum()).reshape(n.out.shape))

    def _bw_nll(self, n, kind):
        x = n.ins[0]
        g = self.buf(n.out, kind)
        z = torch.zeros(x.shape, dtype=self.dtype, device=g.device)
        gv = g if n.attrs["reduction"] == 0 else (n.attrs["scale"] * g.reshape(())).expand(x.shape[0])
        z.scatter_(1, n.attrs["target"].view(-1, 1), -gv.reshape(-1, 1))
        self.put(self.buf(x, kind), z, n.beta[0])

    def bb_nll(self, n):
        self._bw_nll(n, "a")

    def tb_nll(self, n):
        self._bw_nll(n, "at")

    # ---- bce with logits, mean reduction -----------------------------------------------------------------
    def tf_bce_logits(self, n):
        x = n.ins[0]
        z, y = self.base(x), self.base(n.attrs["target"])
        self.buf(n.out, "t").copy_((((torch.sigmoid(z) - y) * self.buf(x, "t")).sum() / z.numel()).reshape(n.out.shape))

    def bb_bce_logits(self, n):
        x = n.ins[0]
        z, y = self.base(x), self.base(n.attrs["target"])
        g = self.buf(n.out, "a").reshape(())
        self.put(self.buf(x, "a"), g * (torch.sigmoid(z) - y) / z.numel(), n.beta[0])

    def tb_bce_logits(self, n):
        x = n.ins[0]
        z, y = self.base(x), self.base(n.attrs["target"])
        s = torch.sigmoid(z)
        g, gt = self.buf(n.out, "a").reshape(()), self.buf(n.out, "at").reshape(())
        self.put(self.buf(x, "at"), (gt * (s - y) + g * s * (1 - s) * self.buf(x, "t")) / z.numel(), n.beta[0])

    # ---- embedding (linear in the table) -------------------------------------------------------------------
    def tf_embedding(self, n):
        w = n.ins[0]
        self.buf(n.out, "t").copy_(self.buf(w, "t")[n.attrs["indices"]])

    def _bw_embedding(self, n, kind):
        w = n.ins[0]
        idx = n.attrs["indices"].reshape(-1)
        g = self.buf(n.out, kind).reshape(idx.numel(), -1)
        if n.attrs["padding_idx"] is not None and n.attrs["padding_idx"] >= 0:
            g = g * (idx != n.attrs["padding_idx"]).to(g.dtype).unsqueeze(1)
        dst = self.buf(w, kind)
        assert n.beta[0] == 1
        dst.index_add_(0, idx, g)

    def bb_embedding(self, n):
        self._bw_embedding(n, "a")

    def tb_embedding(self, n):
        self._bw_embedding(n, "at")
