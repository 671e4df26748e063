"""Per golden case on the GPU: engine vs the oracle in fp64 (deterministic), engine vs the oracle in fp32 (same device, not
run-to-run reproducible), and the oracle's own fp32-vs-fp64 gap in the same run."""
import glob, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betty_b200 import hypergradient as H, workloads as W
from oracle import ref_port
from tests.helpers import GOLDEN, load_golden, rel_l2, to_double

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
only = sys.argv[1] if len(sys.argv) > 1 else ""
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for case in sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, "*.pt"))):
    if only and only not in case:
        continue
    rec = load_golden(case)
    if rec["method"] not in ("neumann", "cg"):
        continue
    fn = ref_port.METHODS[rec["method"]]
    w64 = to_double(W.FACTORIES[rec["factory"]](device="cuda", **rec["kwargs"]))
    want64 = fn(w64.vector, w64.lower, w64.upper, False)
    for _ in range(reps):
        wl = W.FACTORIES[rec["factory"]](device="cuda", **rec["kwargs"])
        want32 = fn(wl.vector, wl.lower, wl.upper, False)
        got = H.jvp_fn_mapping[rec["method"]](wl.vector, wl.lower, wl.upper, False)
        print(f"{case:28s} engine-vs-fp64 {rel_l2(got, want64):.3e}   engine-vs-oracle32 {rel_l2(got, want32):.3e}   "
              f"oracle32-vs-fp64 {rel_l2(want32, want64):.3e}   engine-vs-golden {rel_l2(got, rec['hypergrad']):.3e}", flush=True)
