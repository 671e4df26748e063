"""Time the halo convolution / weight-gradient kernels alone at the implicit-MAML shapes (CUDA events on the launch
stream, L2 flushed between launches by rotating over operand sets larger than L2).  Environment switches of
csrc/conv_halo.cu (BB200_HALO_PIECES, BB200_HALO_STREAM_W) are read once per process: run one process per variant.
    python tools/halo_bench.py [N H W]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from betty_b200 import _native as N


def main():
    n, h, w = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (800, 42, 42)
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    nset = 3
    acts = [[torch.randn(n, h + 2, w + 2, 64, generator=g, device=dev).bfloat16() for _ in range(2)] for _ in range(nset)]
    wms = [(0.1 * torch.randn(64, 9, 64, generator=g, device=dev)).bfloat16() for _ in range(2)]
    outp = [torch.empty(n, h + 2, w + 2, 64, dtype=torch.bfloat16, device=dev) for _ in range(nset)]
    outf = torch.zeros(n, 64, h, w, device=dev)
    wg = torch.zeros(64, 64, 3, 3, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def timed(fn, reps=12):
        for i in range(3):
            fn(i % nset)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn(i % nset)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
        return ts[len(ts) // 2] * 1e3

    def conv_nhwc(np_):
        return lambda i: N.call("bb_conv_halo_bf16_nhwc", n, h, w, np_, acts[i][0].data_ptr(), acts[i][1].data_ptr() if np_ > 1 else 0,
                                wms[0].data_ptr(), wms[1].data_ptr() if np_ > 1 else 0, 0, outp[i].data_ptr(), 0, st)

    def conv_f32(np_):
        return lambda i: N.call("bb_conv_halo_bf16", n, h, w, np_, acts[i][0].data_ptr(), acts[i][1].data_ptr() if np_ > 1 else 0,
                                wms[0].data_ptr(), wms[1].data_ptr() if np_ > 1 else 0, 1, outf.data_ptr(), 0, 0, st)

    def wgrad(np_):
        return lambda i: N.call("bb_wgrad_halo_bf16", n, h, w, np_, acts[i][0].data_ptr(), acts[i][1].data_ptr() if np_ > 1 else 0,
                                acts[(i + 1) % nset][0].data_ptr(), acts[(i + 1) % nset][1].data_ptr() if np_ > 1 else 0,
                                wg.data_ptr(), st)

    flops = 2.0 * n * h * w * 64 * 64 * 9
    env = {k: v for k, v in os.environ.items() if k.startswith("BB200_")}
    print(f"shape N={n} H={h} W={w} env={env}")
    runs = [("conv->nchw f32,  1 pair", conv_f32(1), 1), ("conv->nchw f32,  2 pairs", conv_f32(2), 2), ("wgrad, 2 pairs", wgrad(2), 2)]
    if "BB200_LIB" not in os.environ:
        runs += [("conv->nhwc bf16, 1 pair", conv_nhwc(1), 1), ("conv->nhwc bf16, 2 pairs", conv_nhwc(2), 2)]
    for name, fn, np_ in runs:
        us = timed(fn)
        print(f"  {name:28s} {us:8.1f} us   {np_ * flops / us * 1e-6:7.1f} TF/s")


if __name__ == "__main__":
    main()
