"""Wall-clock breakdown of one hypergradient call (prologue / plan / K-loop / epilogue), both epilogue modes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from betty_b200 import engine as E
from betty_b200.plan import HvpPlan
from betty_b200.trace import record_tape
from betty_b200.ir import lower_tape


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT
    wl, kw, desc = bench.build_workload(name, torch.device("cuda", 0))
    method = wl.lower.config.type
    for mode in ("native", "autograd", "native"):
        E.settings.native_epilogue = mode == "native"
        for rep in range(3):
            t0 = sync()
            call = E.HypergradientCall(wl.lower, method)
            t1 = sync()
            x = call.solve(wl.vector)
            t2 = sync()
            out = call.finish(wl.upper, x, False)
            t3 = sync()
            del call
        # finer split of the prologue
        params = wl.lower.trainable_parameters()
        t4 = sync()
        loss, tape = record_tape(lambda: wl.lower.training_step_exec(wl.lower.cur_batch), params)
        t5 = sync()
        g = lower_tape(tape)
        t6 = sync()
        print(f"{name} epilogue={mode:8s} init {1e3*(t1-t0):7.2f} ms (trace {1e3*(t5-t4):6.2f} + lower {1e3*(t6-t5):6.2f} + plan/BB rest) "
              f"solve {1e3*(t2-t1):7.2f} ms  finish {1e3*(t3-t2):7.2f} ms  total {1e3*(t3-t0):7.2f} ms  nodes={len(g.nodes)}")


if __name__ == "__main__":
    main()
