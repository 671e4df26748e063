"""Dump the aten-level op stream (as seen by TorchDispatchMode) of each workload's lower training_step on
the current device.  Development aid for betty_b200/trace.py; writes gpurun_out/ops_<name>.txt."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from torch.utils._python_dispatch import TorchDispatchMode

from betty_b200 import workloads as W


class Rec(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.log, self.keep = [], []

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        self.keep.append((args, out))

        def d(a):
            if isinstance(a, torch.Tensor):
                return f"T{tuple(a.shape)}:{str(a.dtype)[6:]}:{'R' if a.requires_grad else ''}{'P' if isinstance(a, nn.Parameter) else ''}:s{a.stride()}"
            if isinstance(a, (list, tuple)):
                return "[" + ",".join(d(x) for x in a) + "]"
            return repr(a)

        self.log.append(f"{func} ({', '.join(d(a) for a in args)}) {kwargs if kwargs else ''} -> {d(out)}")
        return out


def main():
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    os.makedirs("gpurun_out", exist_ok=True)
    cases = {
        "lenet": ("learning_to_reweight", dict(batch=4)),
        "fourconv_fp32": ("implicit_maml", dict(n=4, hidden=8)),
        "fourconv_bf16": ("implicit_maml", dict(n=4, hidden=8, precision="bf16")),
        "roberta_fp32": ("bert_data_reweighting", dict(batch=2, seq=6, tiny=True)),
        "roberta_bf16": ("bert_data_reweighting", dict(batch=2, seq=6, tiny=True, precision="bf16")),
        "logistic": ("logistic_regression_hpo", dict()),
        "darts": ("neural_architecture_search", dict(batch=2, c=4, cells=1)),
    }
    for name, (fac, kw) in cases.items():
        wl = W.FACTORIES[fac](device=dev, **kw)
        with Rec() as r:
            loss = wl.lower.training_step_exec(wl.lower.cur_batch)
        cnt = collections.Counter(l.split(" ")[0] for l in r.log)
        with open(f"gpurun_out/ops_{name}_{dev}.txt", "w") as f:
            f.write(f"# {name} on {dev}: loss={float(loss):.6f}\n# {dict(cnt)}\n")
            f.write("\n".join(r.log) + "\n")
        print(name, dict(cnt))


if __name__ == "__main__":
    main()
