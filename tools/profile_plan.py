"""Per-node device time of the two K-loop passes of a workload's native plan (CUDA events around every
node, best of 3) with algorithmic bytes -> gpurun_out/plan_profile_<workload>.md"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from betty_b200 import engine as E
from betty_b200.plan import PASS_TB, PASS_TF


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    wl, kw, desc = bench.build_workload(name, torch.device("cuda", 0))
    call = E.HypergradientCall(wl.lower, wl.lower.config.type)
    call.solve(wl.vector)
    plan = call.hvp
    hbm, _, which = bench.measured_peaks()
    rows = []
    for pas, pn in ((PASS_TF, "TF"), (PASS_TB, "TB")):
        ms = np.min(np.stack([plan.profile(pas) for _ in range(3)]), axis=0)
        for i, t in enumerate(ms):
            n = plan.g.nodes[i]
            b = plan.node_bytes(i, pas)
            rows.append((float(t), pn, i, n.op, tuple(n.out.base.shape) if n.out is not None else ('params',), b, n.src))
    tot = sum(r[0] for r in rows)
    totb = sum(r[5] for r in rows)
    out = [f"# plan profile: {name} ({desc})", "",
           f"nodes={len(plan.g.nodes)} launches/iter={plan.launches_per_iter} sum-of-nodes={tot:.3f} ms "
           f"alg-bytes/iter={totb/1e6:.1f} MB -> {totb/tot/1e6:.1f} GB/s = {100*totb/tot/1e6/hbm:.2f}% of {which} HBM peak {hbm} GB/s",
           "", "| ms | share | pass | node | op | out shape | alg MB | GB/s | % HBM |", "|---|---|---|---|---|---|---|---|---|"]
    for t, pn, i, op, shp, b, src in sorted(rows, reverse=True)[:40]:
        gbs = b / t / 1e6 if t > 0 else 0
        out.append(f"| {t:.4f} | {100*t/tot:.1f}% | {pn} | {i} | {op} | {shp} | {b/1e6:.2f} | {gbs:.0f} | {100*gbs/hbm:.1f} |")
    os.makedirs("gpurun_out", exist_ok=True)
    open(f"gpurun_out/plan_profile_{name}.md", "w").write("\n".join(out) + "\n")
    print("\n".join(out[:30]))


if __name__ == "__main__":
    main()
