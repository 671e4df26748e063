"""Dev probe: bf16 parity of a fourconv_mini case with and without the native prologue BatchNorm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betty_b200 import hypergradient as H, workloads as W, trace as T, engine as E
from oracle import ref_port

def rel(a, b):
    a = torch.cat([t.reshape(-1).double() for t in a]); b = torch.cat([t.reshape(-1).double() for t in b])
    return float((a - b).norm() / b.norm())

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for case, kw in {"small_alpha": dict(method="neumann", n=6, hidden=64, image="miniimagenet", K=20, alpha=1e-3, reg=2.0),
                 "cg": dict(method="cg", n=6, hidden=64, image="miniimagenet", K=3, alpha=1.0),
                 "omni": dict(method="neumann", n=25, hidden=64, K=20, alpha=1e-3, reg=2.0)}.items():
    m = kw["method"]
    wl = W.FACTORIES["implicit_maml"](device="cuda", precision="bf16", **kw)
    want = ref_port.METHODS[m](wl.vector, wl.lower, wl.upper, False)
    for thr in (1 << 20, 0, 1 << 20):
        T.native_bn_min_numel = thr
        E.plan_cache.clear()
        n0 = T.native_bn_calls
        got = H.jvp_fn_mapping[m](wl.vector, wl.lower, wl.upper, False)
        print(f"{case}: threshold {thr}: native BN calls {T.native_bn_calls - n0}, engine-vs-reference {rel(got, want):.3e}", flush=True)
    # the BN outputs themselves
    T.native_bn_min_numel = 1 << 20
    params = wl.lower.trainable_parameters()
    _, tape_a = T.record_tape(lambda: wl.lower.training_step_exec(wl.lower.cur_batch), params)
    T.native_bn_min_numel = 0
    _, tape_b = T.record_tape(lambda: wl.lower.training_step_exec(wl.lower.cur_batch), params)
    for oa, ob in zip(tape_a.ops, tape_b.ops):
        assert oa.name == ob.name
        outs_a = oa.out if isinstance(oa.out, (tuple, list)) else [oa.out]
        outs_b = ob.out if isinstance(ob.out, (tuple, list)) else [ob.out]
        for k, (a, b) in enumerate(zip(outs_a, outs_b)):
            if torch.is_tensor(a) and a.is_floating_point() and a.numel() and a.shape == b.shape:
                ne = float((a != b).float().mean())
                if ne > 0:
                    d = float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
                    print(f"   {oa.name}[{k}] {tuple(a.shape)} {a.dtype}: differing elements {ne:.3e}, rel-L2 {d:.3e}")
    T.native_bn_min_numel = 1 << 20
