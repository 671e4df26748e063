"""Print, for every golden case, the engine's rel-L2 error against the reference's golden vector next to the test
tolerance (max(1e-4, 5 x the reference's own fp32-vs-fp64 gap)).  Usage: python tools/parity_margin.py [case ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from betty_b200 import hypergradient as H
from betty_b200 import workloads as W
from oracle import ref_port
from tests.helpers import load_golden, rel_l2, to_double
from tests.test_parity_gpu import CASES


def main():
    cases = sys.argv[1:] or CASES
    for case in cases:
        rec = load_golden(case)
        wl = W.FACTORIES[rec["factory"]](device="cuda", **rec["kwargs"])
        got = H.jvp_fn_mapping[rec["method"]](wl.vector, wl.lower, wl.upper, False)
        err = rel_l2(got, rec["hypergrad"])
        w32 = W.FACTORIES[rec["factory"]](device="cuda", **rec["kwargs"])
        w64 = to_double(W.FACTORIES[rec["factory"]](device="cuda", **rec["kwargs"]))
        fn = ref_port.METHODS[rec["method"]]
        floor = rel_l2(fn(w32.vector, w32.lower, w32.upper, False), fn(w64.vector, w64.lower, w64.upper, False))
        port = rel_l2(fn(w32.vector, w32.lower, w32.upper, False), rec["hypergrad"])
        print(f"{case:28s} engine-vs-golden {err:.3e}  ref-port(fp32, this GPU)-vs-golden {port:.3e}  "
              f"ref fp32-vs-fp64 {floor:.3e}  tol {max(1e-4, 5 * floor):.3e}  {'OK' if err <= max(1e-4, 5 * floor) else 'OVER'}",
              flush=True)


if __name__ == "__main__":
    main()
