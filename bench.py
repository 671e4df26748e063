#!/usr/bin/env python
"""HVP-iters/sec of the hypergradient K-loop (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W] [--workload NAME] [--impl reference]

One "step" = one full K-loop (K Hessian-vector products + the Neumann/CG vector updates) over one
synthetic batch, inputs resident in HBM; prologue (lower forward + tape) is built once, outside the
timed region, exactly as the metric is defined in SURVEY.md §8(d).  ``e2e`` times the whole plugin
call ``fn(vector, curr, prev, sync)`` with the batch and direction coming from pinned HOST memory and
the hypergradient read back to the host every step.  Under torchrun every rank solves its own local
system on its own batch (the reference's DDP semantics): no collective in the
K-loop, one all-reduce of the hypergradient in ``e2e``.

``--impl reference`` times the oracle port of the reference's CPU autograd path on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (factory, kwargs, method, describe)
    "learning_to_reweight": ("learning_to_reweight", dict(method="cg", batch=4096, K=20), "LeNet-5 3x32x32 B=4096, MWN-weighted CE + 0.05|w|^2, CG K=20, fp32"),
    "learning_to_reweight_b100": ("learning_to_reweight", dict(method="cg", batch=100, K=20), "LeNet-5 B=100 (reference batch), CG K=20, fp32"),
    "implicit_maml": ("implicit_maml", dict(method="neumann", n=800, image="miniimagenet", K=20, alpha=0.01, precision="bf16"), "4-conv mini-ImageNet N=800, CE + 0.5|w-theta|^2, Neumann K=20, bf16 autocast"),
    "implicit_maml_n25": ("implicit_maml", dict(method="neumann", n=25, image="miniimagenet", K=20, alpha=0.01, precision="bf16"), "4-conv mini-ImageNet N=25, Neumann K=20, bf16 autocast"),
    "bert_data_reweighting": ("bert_data_reweighting", dict(method="cg", batch=16, seq=50, K=10, precision="bf16"), "RoBERTa-base B=16xL=50, MWN-weighted CE + 5e-3|w|^2, CG K=10, bf16 autocast"),
    "logistic_regression_hpo": ("logistic_regression_hpo", dict(method="neumann", K=5), "20-dim logistic HPO, Neumann K=5, fp32"),
    "neural_architecture_search": ("neural_architecture_search", dict(batch=64, c=16, cells=4), "DARTS-style supernet c16 x 4 cells B=64, finite-difference hypergradient (1 call = 1 iter-equivalent), fp32"),
}
DEFAULT = "learning_to_reweight"
L2_BYTES = 126 * 1024 * 1024


# -------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.power = [], set(), []
        self.stop_flag = threading.Event()
        self.max_mhz = None
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            self.err = repr(e)

    _NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
              0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
              0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self._NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": max(self.power) if self.power else None}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


def flush_l2(buf):
    buf.add_(1.0)


# -------------------------------------------------------------------------------------------------
def build_workload(name, device, seed=0):
    from betty_b200 import workloads as W

    factory, kw, desc = WORKLOADS[name]
    wl = W.FACTORIES[factory](device=device, seed=seed, **kw)
    return wl, kw, desc


def run_reference(args):
    """--impl reference: the oracle port of the reference's CPU autograd path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref_port

    wl, kw, desc = build_workload(args.workload, "cpu")
    method = wl.lower.config.type
    K = kw.get("K", 1)
    if method == "darts":
        return run_reference_fd(args, wl, desc)
    # bounded sample: cap K so that one step stays within a few seconds-to-tens-of-seconds of CPU work
    k_sample = min(K, args.ref_iters)
    if method == "neumann":
        wl.lower.config.neumann_iterations = k_sample
    else:
        wl.lower.config.cg_iterations = k_sample
    in_grad = ref_port.lower_gradient(wl.lower)
    hvp = ref_port.make_hvp(in_grad, wl.lower.trainable_parameters())
    cores = best_thread_count(hvp, list(wl.vector))

    def step():
        if method == "neumann":
            return ref_port.neumann_series(list(wl.vector), hvp, k_sample, wl.lower.config.neumann_alpha)
        return ref_port.cg_solve(list(wl.vector), hvp, k_sample, wl.lower.config.cg_alpha)

    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        step()
        n += 1
        if time.perf_counter() - t0 > args.ref_budget_s:
            break
    dt = time.perf_counter() - t0
    value = n * k_sample / dt
    line = {
        "impl": "reference", "metric": "HVP-iters/sec", "value": value, "unit": "HVP-iters/s", "n_gpus": args.gpus,
        "steps": n, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / n, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "describe": desc, "K": K},
        "cpu_baseline": {"value": value, "unit": "HVP-iters/s", "cores": cores, "kind": "port",
                         "sample": f"{n} step(s) x {k_sample} of K={K} iterations, full batch, torch {torch.__version__} CPU autograd double backward"},
        "e2e": {"value": value, "unit": "HVP-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_reference_fd(args, wl, desc):
    from oracle import ref_port

    cores = os.cpu_count() or 1
    torch.set_num_threads(min(cores, 32))
    ref_port.darts(wl.vector, wl.lower, wl.upper, False)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        ref_port.darts(wl.vector, wl.lower, wl.upper, False)
        n += 1
        if time.perf_counter() - t0 > args.ref_budget_s:
            break
    dt = time.perf_counter() - t0
    value = n / dt
    print(json.dumps({
        "impl": "reference", "metric": "HVP-iters/sec", "value": value, "unit": "HVP-iters/s", "n_gpus": args.gpus,
        "steps": n, "warmup": 1, "ms_per_step": 1e3 * dt / n, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": args.workload, "describe": desc, "K": 1},
        "cpu_baseline": {"value": value, "unit": "HVP-iters/s", "cores": min(cores, 32), "kind": "port",
                         "sample": f"{n} finite-difference call(s), torch CPU autograd"},
        "e2e": {"value": value, "unit": "HVP-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


# -------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hvp", default="native", choices=["native", "autograd"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-iters", type=int, default=4, help="K-loop iterations per reference step (bounded sample)")
    ap.add_argument("--ref-budget-s", type=float, default=60.0)
    ap.add_argument("--e2e-steps", type=int, default=5)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    from betty_b200 import _native as N
    from betty_b200 import engine as E
    from betty_b200 import hypergradient as H

    E.settings.hvp = args.hvp
    E.settings.cuda_graph = not args.no_graph
    wl, kw, desc = build_workload(args.workload, dev, seed=rank)
    kw = dict(kw)
    method = wl.lower.config.type
    K = kw.get("K", 1)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- prologue once (outside the timed region) -------------------------------------------
    if method == "darts":
        # finite difference: no K-loop to isolate -- one step = one plugin call (two lower fwd/bwd passes on
        # PyTorch + the K4 kernels), reported as calls/s ("1 call = 1 iter-equivalent")
        class _FdCall:
            hvp = None
            layout = type("L", (), {"n_logical": sum(p.numel() for p in wl.lower.parameters())})()

            def solve(self, vec):
                return H.darts(vec, wl.lower, wl.upper, False)

        call = _FdCall()
        K = kw["K"] = 1
    else:
        call = E.HypergradientCall(wl.lower, method)
    flush = torch.zeros(L2_BYTES // 4 * 2, device=dev)  # 252 MB > L2, written between steps

    for _ in range(args.warmup):
        call.solve(wl.vector)
        flush_l2(flush)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = N.launch_counter
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ev[i][0].record()
        call.solve(wl.vector)
        ev[i][1].record()
        flush_l2(flush)
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag.set()
    sampler.join()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = sum(step_ms)  # device time of the K steps, L2 flushes excluded
    launches = N.launch_counter - launches0
    plan_launches = getattr(call.hvp, "launches_per_iter", 0) * K * args.steps
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    ms_per_step = dev_ms_max / args.steps
    value = world * K * args.steps / (dev_ms_max / 1e3)

    # ---- e2e: whole plugin call from pinned host buffers -------------------------------------
    host_batch = [b.cpu().pin_memory() if torch.is_tensor(b) else b for b in wl.lower.cur_batch]
    host_vec = [v.cpu().pin_memory() for v in wl.vector]
    h2d = sum(b.numel() * b.element_size() for b in host_batch if torch.is_tensor(b)) + sum(v.numel() * 4 for v in host_vec)
    d2h = 0
    fn = H.jvp_fn_mapping[method]

    def e2e_step():
        nonlocal d2h
        wl.lower.cur_batch = tuple(b.to(dev, non_blocking=True) if torch.is_tensor(b) else b for b in host_batch)
        vec = [v.to(dev, non_blocking=True) for v in host_vec]
        hg = fn(vec, wl.lower, wl.upper, False)
        if dist is not None:
            flat = torch.cat([g.reshape(-1) for g in hg])
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)  # what the upper module's DDP reducer does
            hg = [flat]
        out = [g.cpu() for g in hg]
        d2h = sum(o.numel() * o.element_size() for o in out)
        return out

    import gc

    for _ in range(2):      # warm-up: allocator pools, cuBLAS/cuDNN handles of the lower forward, graph instantiation
        e2e_step()
    gc.collect()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_wall = time.perf_counter() - t1
    t = torch.tensor([e2e_wall], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * K * args.e2e_steps / float(t.item())

    # ---- roofline of the dominant kernel -------------------------------------------------------
    hbm, tf, which = measured_peaks()
    roof = None
    if call.hvp is not None and hasattr(call.hvp, "roofline"):
        roof = call.hvp.roofline(hbm_gbs=hbm, which=which)
    if roof is None:
        # development mode: the only kernels of ours in the loop are the flat-arena updates
        n = 128 * 1024 * 1024
        a, b, c, d = (torch.randn(n, device=dev) for _ in range(4))
        ws = E.Workspace.get(dev)
        s = torch.cuda.current_stream().cuda_stream
        N.call("bb_cg_dots", b.data_ptr(), d.data_ptr(), c.data_ptr(), 1.0, 1, n, ws.ptr, s)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            N.call("bb_cg_update_xr", a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr(), n, ws.ptr, s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        ach = 24.0 * n / (ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "cg_update_xr_kernel (128Mi-element arena, isolated)", "achieved": ach,
                "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None, "peak_source": which}
        del a, b, c, d

    # DRAM traffic of the dominant node from a committed `ncu --set full` capture of the same workload (sum of
    # dram__bytes_read.sum + dram__bytes_write.sum over the node's kernels, per launch), when one exists
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")) as f:
            tr = json.load(f).get(args.workload, {})
        for key, rec in tr.items():
            if roof.get("kernel", "").startswith(key):
                roof["traffic"] = rec["bytes"]
                roof["traffic_source"] = rec["source"]
    except (OSError, ValueError, KeyError):
        pass

    line = {
        "metric": "HVP-iters/sec", "value": value, "unit": "HVP-iters/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if wl.lower.config.precision == "fp32" else "bf16+f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "describe": desc, "K": K, "method": method, "hvp": args.hvp,
                   "cuda_graph": E.settings.cuda_graph, "l2": "252 MB buffer rewritten between timed steps",
                   "parallelism": f"replicas x{world} (local solve per rank, SURVEY 8e)",
                   "n_params": call.layout.n_logical, "wall_s": wall},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "HVP-iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": args.e2e_steps, "includes": "H2D batch+v, prologue, K-loop, epilogue, D2H hypergradient"},
        "gpu_launches": launches + plan_launches,
        "roofline": roof,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, kw)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def best_thread_count(hvp, v, budget_s: float = 40.0):
    """torch's CPU autograd is not monotone in thread count on many-core hosts (128 threads were 4x slower than 16
    on the pool's boxes): time one H.v at 8 / 16 / 32 threads (capped by the core count), stop as soon as more
    threads get slower or the calibration budget is spent, keep the fastest."""
    cores = os.cpu_count() or 1
    cands = sorted({min(cores, 8), min(cores, 16), min(cores, 32)})
    t_start = time.perf_counter()
    best, best_t = cands[0], None
    for i, nt in enumerate(cands):
        torch.set_num_threads(nt)
        if i == 0:
            hvp(v)          # warm-up (allocator, oneDNN primitive caches)
        t0 = time.perf_counter()
        hvp(v)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = nt, dt
        elif dt > 1.1 * best_t:
            break
        if time.perf_counter() - t_start > budget_s:
            break
    torch.set_num_threads(best)
    return best


def cpu_baseline(args, kw):
    """Oracle port (reference algorithm, torch CPU autograd) on this box's host cores, bounded sample."""
    from oracle import ref_port

    wl, kw, desc = build_workload(args.workload, "cpu")
    method = wl.lower.config.type
    if method == "darts":
        cores = min(os.cpu_count() or 1, 32)
        torch.set_num_threads(cores)
        ref_port.darts(wl.vector, wl.lower, wl.upper, False)
        t0 = time.perf_counter()
        n = 0
        while n < 3 and time.perf_counter() - t0 < 12.0:
            ref_port.darts(wl.vector, wl.lower, wl.upper, False)
            n += 1
        dt = time.perf_counter() - t0
        return {"value": n / dt, "unit": "HVP-iters/s", "cores": cores, "kind": "port",
                "sample": f"{n} finite-difference call(s) ({dt:.1f} s), fp32, torch CPU autograd"}
    in_grad = ref_port.lower_gradient(wl.lower)
    hvp = ref_port.make_hvp(in_grad, wl.lower.trainable_parameters())
    v = list(wl.vector)
    cores = best_thread_count(hvp, v)
    iters = 0
    t0 = time.perf_counter()
    while iters < kw.get("K", 1) and (time.perf_counter() - t0 < 12.0 or iters < 1):
        if method == "neumann":
            ref_port.neumann_series(v, hvp, 1, wl.lower.config.neumann_alpha)
        else:
            ref_port.cg_solve(v, hvp, 1, wl.lower.config.cg_alpha)
        iters += 1
    dt = time.perf_counter() - t0
    return {"value": iters / dt, "unit": "HVP-iters/s", "cores": cores, "kind": "port",
            "sample": f"{iters} K-loop iteration(s) of the full-size workload ({dt:.1f} s), fp32, torch CPU autograd"}


if __name__ == "__main__":
    sys.exit(main())
