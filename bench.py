#!/usr/bin/env python
"""bench.py — headline benchmark: probe-batched GGN-vector products/s (+ SLQ logdet time) on the MNIST-MLP
configuration C3b of BASELINE.md (LargeClassifier 784-1024-512-256-128-10, D=1,494,154, M=512 inducing points,
N=60,000, alpha=1e-3), synthetic weights / points / Rademacher probes.

  python bench.py --gpus N --steps K --warmup W            # B200 path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on the host CPU cores

One "step" = one Hutchinson trace-estimator pass (src/stochtrace.py:22-34) over this rank's B probes:
curvature_vp(V) (src/lla.py:19-23 -> src/ggn.py:133-144) for all probes in ONE lip_ggn_vp call, the per-probe
quadratic forms v.(Gv) (lip_dot), and — for N>1 — one NCCL all-reduce of the trace accumulator.  Probes are
sharded over ranks (weak scaling: B per GPU fixed), weights/points/activation cache replicated.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DIMS = [784, 1024, 512, 256, 128, 10]
SIGMA_ALL = sum(DIMS[i] * DIMS[i + 1] for i in range(len(DIMS) - 1))
SIGMA_GE2 = sum(DIMS[i] * DIMS[i + 1] for i in range(1, len(DIMS) - 1))

# workload -> shape of the synthetic problem.  "mlp" (C3b) is the headline the metric is quoted on;
# "lenet5" (C4) is a secondary measurement of the conv path.
WORKLOADS = {
    "mlp": dict(name="C3b MNIST-MLP 784-1024-512-256-128-10 (D=1494154), M=512, N=60000, alpha=1e-3",
                M=512, N=60_000, alpha=1e-3, K=10,
                flop_per_product=512 * (4 * SIGMA_ALL + 4 * SIGMA_GE2)),          # 4.468 GFLOP
    "lenet5": dict(name="C4 LeNet5 (D=61706), M=200, N=60000, alpha=5e-3", M=200, N=60_000, alpha=5e-3, K=10,
                   flop_per_product=200 * (8 * 416_520 - 4 * 117_600)),           # 2,861,760 FLOP / point
    # C5 shape at the reference's own largest M (config/scale/resnet1-2_cifar10.yml:15, m=100): the conv path
    # materialises im2col patches, so BASELINE's synthetic M=4096 needs the implicit-GEMM kernel of a later round
    "resnet1m": dict(name="C5 ResNet1M CIFAR-shaped 32x32x3 (D=1084586), M=100, N=49000, alpha=5e-3", M=100, N=49_000,
                     alpha=5e-3, K=10, flop_per_product=100 * 1_300_000_000),     # ~1.30 GFLOP / point
}
M_POINTS, N_FULL, ALPHA = 512, 60_000, 1e-3
FLOP_PER_PRODUCT = WORKLOADS["mlp"]["flop_per_product"]
WORKLOAD = WORKLOADS["mlp"]["name"]


def set_workload(name):
    global M_POINTS, N_FULL, ALPHA, FLOP_PER_PRODUCT, WORKLOAD
    w = WORKLOADS[name]
    M_POINTS, N_FULL, ALPHA, FLOP_PER_PRODUCT, WORKLOAD = w["M"], w["N"], w["alpha"], w["flop_per_product"], w["name"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--probes", type=int, default=256, help="probes per GPU per step")
    ap.add_argument("--slq-k", type=int, default=409, help="GKL depth; int(0.8*M) as train_inducing.py:148")
    ap.add_argument("--slq-probes", type=int, default=4)
    ap.add_argument("--cpu-probes", type=int, default=2, help="products in the CPU-baseline sample")
    ap.add_argument("--no-slq", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the Z-gradient / inducing-point training-step leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the C4 (LeNet5) / C5 (ResNet1M) sub-records")
    ap.add_argument("--tensor-path", type=int, default=-1, help="-1 auto, 0 SIMT fp32, 1 tcgen05 3xTF32")
    ap.add_argument("--workload", default="mlp", choices=sorted(WORKLOADS), help="mlp = headline C3b; lenet5 = C4 conv path")
    ap.add_argument("--points", type=int, default=0, help="override the workload's number of inducing points M")
    return ap.parse_args()


def build_states(seed=1003, workload="mlp"):
    import numpy as np
    from helpers import make_pair
    rng = np.random.default_rng(seed + 1)
    if workload == "resnet1m":
        ost, lst = make_pair("resnet1m", n_out=10, seed=seed, in_shape=(32, 32, 3))
        Z = rng.random((M_POINTS, 32, 32, 3), dtype=np.float32)
        return ost, lst, Z
    if workload == "lenet5":
        ost, lst = make_pair("lenet5", seed=seed)
        Z = rng.random((M_POINTS, 28, 28, 1), dtype=np.float32)
        return ost, lst, Z
    ost, lst = make_pair("large", hidden=DIMS[1:-1], n_out=DIMS[-1], in_dim=DIMS[0], seed=seed, in_shape=(28, 28, 1))
    Z = rng.random((M_POINTS, DIMS[0]), dtype=np.float32)
    return ost, lst, Z


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons DURING the timed region (NVML; nvidia-smi as the fallback)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.active = True          # samples are kept only while this is set (the timed region)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        bits = [getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)]
        return [str(mhz), str(self.max_mhz)] + ["Active" if (r & b) else "Not Active" for b in bits]

    def run(self):
        while not self.stop_flag:
            try:
                if not self.active:
                    time.sleep(0.002)
                    continue
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    time.sleep(0.01)
                    continue
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                if self.nvml is not None:
                    self.nvml = None       # fall back to nvidia-smi
                    continue
                return
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return None
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = sorted({self.NAMES[i] for s in self.samples for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measure_tf32_peak(torch):
    """cuBLAS TF32 8192^3, best of 10 (the way MEASURED_PEAKS.json measures bf16) — the tensor roofline denominator."""
    n = 8192
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    best = float("inf")
    for _ in range(3):
        torch.matmul(a, b)
    for _ in range(10):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return 2 * n ** 3 / (best * 1e-3) / 1e12


def cpu_reference_products_per_s(ost, Z, n_products=1, steps=2, warmup=1):
    """The reference's algorithm (src/ggn.py:133-144: sequential loop over the M points, per-point jvp, redundant
    forward, closed-form Hessian, per-point vjp) restated in torch fp32 on the host cores (oracle/, kind 'port':
    JAX is not installed in this image so /root/reference cannot run).  ONE function for both arms (`cpu_baseline` of the GPU
    line and `--impl reference`): `warmup` untimed passes (the first pass pays torch's thread-pool / autograd warm-up, which
    made round 1's two numbers differ by 2x), then `steps` timed passes of `n_products` products each."""
    import numpy as np
    import torch
    from oracle import lip_oracle as O
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    vp = O.compute_curvature_approx(ost, Z, "classifier", ALPHA, full_set_size=N_FULL, sequential=True,
                                    dtype=torch.float32)
    D = ost.flat()[0].size
    rng = np.random.default_rng(7)
    V = rng.choice([-1.0, 1.0], size=(n_products, D)).astype(np.float32)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for v in V:
            vp(v)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return n_products * len(times) / total, cores, total / len(times)


def bench_config(args, world):
    """The `config` object of BOTH arms (the driver compares them): the workload and the per-GPU probe batch the metric is quoted on."""
    import numpy as np
    D = {"mlp": 1_494_154, "lenet5": 61_706, "resnet1m": 1_084_586}[args.workload]
    return {"workload": WORKLOAD, "probes_per_gpu": args.probes,
            "l2": "inputs larger than L2 (V and out are %.2f GB each per GPU)" % (args.probes * D * 4 / 1e9),
            "parallelism": f"probe-sharded x{world}"}


def reference_import_status():
    """The unmodified reference (pip-installed from /root/reference into baseline/_ref) is pure Python on top
    of JAX / flax / matfree.  Returns None when it can be imported, else the reason it cannot (then the oracle port runs)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "src")):
        return "baseline/_ref/src missing (reference not installed in this tree)"
    try:
        import importlib
        for mod in ("jax", "flax", "matfree"):
            importlib.import_module(mod)
    except Exception as e:          # noqa: BLE001
        return f"{type(e).__name__}: {e}"
    return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ost, _, Z = build_states(workload=args.workload)
    # bounded sample: ONE product (all M points) per step so that --steps 20 --warmup 5 ends within a minute; the metric is per product,
    # so it is comparable with the GPU arm's batched steps (config.probes_per_gpu is the batch the GPU arm is quoted on)
    n_prod = 1
    pps, cores, sec = cpu_reference_products_per_s(ost, Z, n_prod, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    why = reference_import_status()
    sample = (f"{n_prod} product (all {M_POINTS} points) per step, torch fp32 restatement of src/ggn.py:133-144 "
              f"(oracle port; the installed reference itself is not runnable here: {why})")
    line = {"impl": "reference", "metric": "ggn_vec_products_per_s", "value": pps, "unit": "products/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, args.gpus),
            "cpu_baseline": {"value": pps, "unit": "products/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": pps, "unit": "products/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def measure_extra(name, points, probes, steps, warmup, shard_points, torch, dist, world, rank, dev):
    """A secondary workload of BASELINE.json's configs (C4 LeNet5 probe-sharded, C5 ResNet1M point-sharded) measured like the headline:
    `steps` curvature_vp calls over `probes` Rademacher probes, CUDA events, max over ranks.  Returns a sub-record."""
    import numpy as np
    from lip_b200 import lla
    saved = (M_POINTS, N_FULL, ALPHA, FLOP_PER_PRODUCT, WORKLOAD)
    try:
        set_workload(name)
        g = globals()
        if points:
            g["FLOP_PER_PRODUCT"] = FLOP_PER_PRODUCT // M_POINTS * points
            g["WORKLOAD"] = WORKLOAD.replace(f"M={M_POINTS}", f"M={points}")
            g["M_POINTS"] = points
        ost, lst, Z = build_states(workload=name)
        D = ost.flat()[0].size
        Zd = torch.as_tensor(Z, device=dev)
        t0 = time.perf_counter()
        cvp = lla.compute_curvature_approx(lst, Zd, "classifier", ALPHA, full_set_size=N_FULL, shard_points=shard_points and world > 1)
        torch.cuda.synchronize()
        bind_s = time.perf_counter() - t0
        seed = 3000 if shard_points else 3000 + rank           # point sharding: every rank pushes the SAME probes
        V = torch.from_numpy(np.random.default_rng(seed).integers(0, 2, size=(probes, D), dtype=np.int8) * 2 - 1).to(dev).float()
        for _ in range(warmup):
            cvp(V)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(steps):
            Y = cvp(V)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_step = float(ms.item()) / steps
        total_products = probes * (1 if shard_points else world)
        bm = getattr(cvp, "_lip_model", None)
        rec = {"workload": WORKLOAD, "value": total_products / (ms_step * 1e-3), "unit": "products/s", "ms_per_step": ms_step,
               "probes_per_step": total_products, "steps": steps, "warmup": warmup, "bind_seconds": bind_s,
               "parallelism": (f"point-sharded x{world}: M/{world} points per GPU, one NCCL all-reduce of the [B, D] block per call"
                               if shard_points and world > 1 else f"probe-sharded x{world}"),
               "path": bm.path_name() if bm is not None and hasattr(bm, "path_name") else None,
               "algorithmic_tflops": FLOP_PER_PRODUCT * total_products / (ms_step * 1e-3) / 1e12 / world,
               "finite": bool(torch.isfinite(Y).all())}
        del cvp, V, Y
        torch.cuda.empty_cache()
        return rec
    finally:
        g = globals()
        g["M_POINTS"], g["N_FULL"], g["ALPHA"], g["FLOP_PER_PRODUCT"], g["WORKLOAD"] = saved


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as entry
    entry.build()
    import lip_b200  # noqa: F401
    from lip_b200 import _cabi, lla, matfree, ggn
    from lip_b200._runtime import ptr, scratch, stream
    L = _cabi.lib()

    ost, lst, Z = build_states(workload=args.workload)
    D = ost.flat()[0].size
    B = args.probes
    dev = torch.device("cuda", local)
    Zd = torch.as_tensor(Z, device=dev)
    tp = None if args.tensor_path < 0 else bool(args.tensor_path)
    cvp = lla.compute_curvature_approx(lst, Zd, "classifier", ALPHA, full_set_size=N_FULL, tensor_path=tp)
    bm = cvp._lip_model
    path = bm.path_name() if hasattr(bm, "path_name") else "simt-fp32"

    rng = np.random.default_rng(2000 + rank)
    host_i8 = torch.from_numpy(rng.integers(0, 2, size=(B, D), dtype=np.int8) * 2 - 1).pin_memory()
    # +-1 probes in padded rows, marked exactly-TF32 (what stochtrace._rademacher / unpack_rademacher hand out): lip_ggn_vp_ex then
    # reads the probe block in place instead of running the TF32 split pass
    from lip_b200._runtime import exact_tf32_block
    V = exact_tf32_block(B, D, dev)
    V.copy_(host_i8.to(dev).float())
    q = torch.empty(B, device=dev)
    acc = torch.zeros(1, device=dev)
    sc, _ = scratch(L.lip_dot_scratch_bytes(D, B))

    def step(Vin):
        Y = cvp(Vin)                                                  # [B, D]: one lip_ggn_vp call
        _cabi.check(L.lip_dot(ptr(Vin), ptr(Y), ptr(q), D, B, Vin.stride(0), D, sc, stream()))
        part = q.sum(0, keepdim=True)
        if world > 1:
            dist.all_reduce(part)                                     # trace accumulator (probe-sharded Hutchinson)
        acc.add_(part)
        return Y

    # tensor roofline denominator, first reading (cool GPU); a second one is taken after the runs and the LARGER of the two
    # is used, so a throttled reading can only lower the reported fraction
    tf32_peak_pre = measure_tf32_peak(torch) if rank == 0 else 0.0
    # NVML is initialised and the sampler thread started BEFORE the warm-up and the barrier: its start-up cost differs from rank to
    # rank (10 - 100 ms), and paid between the barrier and the timed loop it de-synchronised the ranks, so that the first all-reduce of
    # the loop absorbed the skew - up to 3 ms per step over a 10-step run on 8 GPUs.  It only keeps samples inside the timed region.
    sampler = ClockSampler(local)
    sampler.active = False
    sampler.start()
    for _ in range(args.warmup):
        step(V)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = L.lip_launch_count()
    sampler.active = True
    e0.record()
    for _ in range(args.steps):
        step(V)
    e1.record()
    torch.cuda.synchronize()
    sampler.active = False
    launches = L.lip_launch_count() - l0
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total * 1e-3)
    trace_est = float(acc.item()) / (B * world * (args.steps + args.warmup))
    # ---- end to end through the public API with HOST probes, copies inside the timed region ----
    # Rademacher probes live in pinned host memory in their packed wire format (1 bit / element, numpy.packbits); every step
    # copies its probes H2D, unpacks them on the device (lip_unpack_rademacher), runs the public
    # lla.compute_curvature_approx(...)(V) + quadratic forms, and copies the step's B results back to pinned host memory.
    e2e = None
    if not args.no_e2e:
        from lip_b200 import stochtrace as st_mod
        host_bits = torch.from_numpy(st_mod.pack_rademacher(host_i8.numpy())).pin_memory()
        nbytes_row = host_bits.shape[1]
        copy_stream = torch.cuda.Stream()
        bufs = [torch.empty(B, nbytes_row, dtype=torch.uint8, device=dev) for _ in range(2)]
        vbuf = exact_tf32_block(B, D, dev)
        evs = [torch.cuda.Event() for _ in range(2)]
        host_q = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(2)]
        qev = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                bufs[i % 2].copy_(host_bits, non_blocking=True)
                evs[i % 2].record(copy_stream)

        def e2e_run(nsteps):
            res = []
            prefetch(0)
            for i in range(nsteps):
                if i + 1 < nsteps:
                    copy_stream.wait_stream(torch.cuda.current_stream())   # buffer (i+1)%2 is free once step i-1 is done
                    prefetch(i + 1)
                torch.cuda.current_stream().wait_event(evs[i % 2])
                st_mod.unpack_rademacher(bufs[i % 2], D, out=vbuf)
                step(vbuf)
                if i >= 2:
                    qev[i % 2].synchronize()                               # result of step i-2 has landed: consume it
                    res.append(float(host_q[i % 2].sum()))
                host_q[i % 2].copy_(q, non_blocking=True)                  # D2H of the step's result (B floats)
                qev[i % 2].record()
            torch.cuda.synchronize()
            return res

        e2e_run(2)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world * args.steps / float(dt.item()), "unit": "products/s",
               "h2d_bytes_per_step": B * nbytes_row, "d2h_bytes_per_step": B * 4,
               "note": "pinned bit-packed +-1 probes -> H2D -> lip_unpack_rademacher -> lla.compute_curvature_approx(...)(V) -> "
                       "v.(Gv) -> pinned host; H2D double-buffered, D2H pipelined two steps deep"}

    # second end-to-end leg: GENERAL fp32 vectors (what a CG / Lanczos caller on the host would hand over): pinned [B, D] fp32 in,
    # the full [B, D] fp32 result out, both inside the timed region (3.06 GB each way per step at B = 256: PCIe-bound by construction)
    e2e_fp32 = None
    if not args.no_e2e:
        hV = torch.empty(B, D, dtype=torch.float32).pin_memory()
        hV.copy_(V)
        hYs = [torch.empty(B, D, dtype=torch.float32).pin_memory() for _ in range(2)]
        dVs = [torch.empty(B, D, device=dev) for _ in range(2)]
        n_e = max(2, min(args.steps, 6))
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def e2e_fp32_run(n):
            # the host link is full duplex: step i + 1 is copied in and step i - 1 copied out while step i computes (two device input
            # buffers, two pinned output buffers, one copy stream per direction)
            main_s = torch.cuda.current_stream(dev)
            ev_in = [torch.cuda.Event() for _ in range(2)]
            ev_done = [torch.cuda.Event() for _ in range(2)]
            ev_out = [torch.cuda.Event() for _ in range(2)]
            keep = [None, None]
            for i in range(n):
                j = i % 2
                with torch.cuda.stream(s_in):
                    if i >= 2:
                        s_in.wait_event(ev_done[j])              # the product of step i - 2 has consumed this input buffer
                    dVs[j].copy_(hV, non_blocking=True)
                    ev_in[j].record(s_in)
                main_s.wait_event(ev_in[j])
                Y = cvp(dVs[j])
                ev_done[j].record(main_s)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[j])
                    if i >= 2:
                        s_out.wait_event(ev_out[j])              # (stream order already guarantees it)
                    hYs[j].copy_(Y, non_blocking=True)
                    ev_out[j].record(s_out)
                Y.record_stream(s_out)
                keep[j] = Y
            torch.cuda.synchronize()

        e2e_fp32_run(2)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e2e_fp32_run(n_e)
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_fp32 = {"value": B * world * n_e / float(dt.item()), "unit": "products/s", "steps": n_e,
                    "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": B * D * 4,
                    "note": "pinned fp32 [B, D] vectors -> H2D -> lla.compute_curvature_approx(...)(V) -> [B, D] products -> D2H to pinned "
                            "host; copies double-buffered on one stream per direction (full-duplex link): bounded by the host link, not by "
                            "the kernels"}
        del hV, hYs, dVs

    # the dominant kernel group on its own: ONE lip_ggn_vp call (all JVP / VJP GEMM launches) under CUDA events on the
    # launching stream, same inputs, no quadratic form / all-reduce around it -> roofline.achieved
    k0, k1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(args.steps):
        cvp(V)
    k1.record()
    torch.cuda.synchronize()
    ms_ggn = torch.tensor([k0.elapsed_time(k1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms_ggn, op=dist.ReduceOp.MAX)
    ms_ggn = float(ms_ggn.item())
    # the same call on GENERAL (Gaussian) vectors - CG / Lanczos iterates are not exactly TF32, so the zero-lo shortcut
    # that +-1 probes enjoy does not apply: reported beside the headline
    Vg = torch.randn(B, D, device=dev)
    cvp(Vg)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(args.steps):
        cvp(Vg)
    k1.record()
    torch.cuda.synchronize()
    ms_gauss = k0.elapsed_time(k1) / args.steps
    del Vg

    # ---- SLQ logdet (GKL form, src/train_inducing.py:148-171), probes sharded over ranks ----
    slq = None
    if not args.no_slq:
        from lip_b200 import _dist
        k, ns = args.slq_k, args.slq_probes
        Wz, WzT = ggn.compute_W_vps(lst, Zd, "classifier", full_set_size=None, tensor_path=tp)
        sa = math.sqrt(ALPHA)
        d = M_POINTS * DIMS[-1]

        # bidiag_target v -> [sqrt(alpha) v ; Wz^T v] (train_inducing.py:166-169) and its transpose (jax.vjp in the reference)
        Av = matfree.gkl_target(WzT, Wz, ALPHA)
        vA = Av._lip_transpose

        # identical probe matrix on every rank (probes are inputs).  Layout: probes are split over the ranks first; with more GPUs
        # than probes the Krylov bases of a probe are cut column-wise over the ranks of its group (_dist.slq_logdet_hybrid ->
        # lip_slq_quadrature_sharded).  ONE native call per rank runs the whole recurrence.
        gp = torch.Generator(device=dev)
        gp.manual_seed(4242)
        slq_probes = torch.randint(0, 2, (ns, D), generator=gp, device=dev, dtype=torch.int8).float() * 2 - 1
        layout = _dist.group_layout(world, ns)
        # warm-up at the measured size: communicators, kernel attributes and the Krylov workspace (the 2.4 GB-per-probe basis is
        # allocated once and kept, as in a training loop that evaluates the logdet every step; timing the first call charged
        # 0.15 - 0.2 s of cudaMalloc to some runs and not to others)
        _dist.slq_logdet_hybrid(Av, slq_probes, k, form="gkl")
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = L.lip_launch_count()
        t0 = time.perf_counter()
        est = _dist.slq_logdet_hybrid(Av, slq_probes, k, form="gkl")
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        slq_launches = L.lip_launch_count() - l0
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # the Lanczos form (integrand_funm_sym_logdet of src/matfree_monkeypatch.py:25-41, eigenvalues clipped to >= 1;
        # train_inducing.py:152-153) on the curvature operator itself, same probes and depth
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        est_lz = _dist.slq_logdet_hybrid(cvp, slq_probes, k, form="lanczos", clip_min=1.0)
        torch.cuda.synchronize()
        dt_lz = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt_lz, op=dist.ReduceOp.MAX)
        Pg, Sg = layout
        n_loc = (ns + Pg - 1) // Pg                  # probes of the busiest group
        # algorithmic re-orthogonalisation traffic of the slowest rank.  The reference's recurrence (matfree bidiag) reads, at step i,
        # i rows of U (n = D + d) and i + 1 rows of V (n = D) twice each (project, subtract): SURVEY 8d, `reference_bytes` below.
        # The native recurrence carries U in reduced coordinates (k + d floats per row, L2-resident; lip_krylov.cu gkl_run), so what it
        # has to move through HBM is the V side, of which a rank owns 1 / Sg of every row - and the u rows, replicated.
        reduced = os.environ.get("LIP_GKL_REDUCED", "1") != "0"
        ref_bytes = n_loc * 4 * sum(2 * i * (D + d) + 2 * (i + 1) * D for i in range(k)) // Sg
        kp = (k + 3) // 4 * 4
        reorth_bytes = (n_loc * 4 * sum(2 * (i + 1) * D for i in range(k)) // Sg + n_loc * 4 * sum(2 * i * (kp + d) for i in range(k))
                        if reduced else ref_bytes)
        secs = float(dt.item())
        hbm = None
        try:
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
        except Exception:
            pass
        # three-term pass (2 rows) + ONE full pass (project, subtract) over i + 1 rows of Q; two full passes (CGS2) read twice as much
        lz_bytes = n_loc * 4 * sum(2 * (i + 1) * D + 4 * D for i in range(k)) // Sg
        slq = {"seconds": secs, "k": k, "probes": ns, "probes_per_group": n_loc, "logdet_estimate": float(est.item()),
               "layout": {"probe_groups": Pg, "basis_shards_per_group": Sg,
                          "note": "probes over groups of GPUs (no communication); inside a group the Krylov bases are cut column-wise "
                                  "(one all-gather of the new vector + one all-reduce per norm / coefficient vector per step)"},
               "launches_per_logdet": int(slq_launches),
               "lanczos_form": {"seconds": float(dt_lz.item()), "logdet_estimate_clip1": float(est_lz.item()),
                                "form": "Lanczos tridiag_sym(k) on curvature_vp, full re-orthogonalisation (three-term pass + one full "
                                        "pass, a second full pass per column where the first removed more than half), "
                                        "log(clip(eig, 1)) quadrature",
                                "reorth_GBps": lz_bytes / float(dt_lz.item()) / 1e9},
               "form": "GKL bidiag on [sqrt(alpha) I; Wz^T], full re-orthogonalisation; one native call (lip_slq_quadrature_sharded)"
                       + ("; u basis in reduced coordinates [coefficients over V ; output-space part]" if reduced else ""),
               "roofline": {"bound": "hbm", "achieved": reorth_bytes / secs / 1e9, "peak": hbm, "unit": "GB/s",
                            "frac": (reorth_bytes / secs / 1e9 / hbm) if hbm else None,
                            "algorithmic_bytes": reorth_bytes, "reference_recurrence_bytes": ref_bytes,
                            "note": "re-orthogonalisation bytes only, divided by the WHOLE logdet wall time (mat-vecs, "
                                    "eigensolve and host orchestration included)"}}

    # ---- gradient with respect to Z (lip_zgrad) and one inducing-point training step ----
    train = None
    if not args.no_train_step and world == 1 and args.workload == "mlp":
        from lip_b200 import train_inducing, utils as lip_utils
        nb = 64
        Pz = (torch.randint(0, 2, (nb, D), device=dev, generator=torch.Generator(device=dev).manual_seed(5)).float() * 2 - 1)
        cvp_z = lla.compute_curvature_approx(lst, Zd, "classifier", ALPHA, full_set_size=N_FULL, tensor_path=tp)
        cvp_z.zgrad(Pz, Pz)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            cvp_z.zgrad(Pz, Pz)
        e1.record()
        torch.cuda.synchronize()
        ms_z = e0.elapsed_time(e1) / 3
        # the reference's own training configuration (config/scale/mlp_mnist.yml: m = 50 inducing points, batch 256, 64 probes)
        m_ref, k_ref = 50, 40
        Zs = Zd[:m_ref].contiguous()
        Xb = torch.rand(256, *Zd.shape[1:], device=dev, generator=torch.Generator(device=dev).manual_seed(6))
        opt = lip_utils.adam(1e-3)
        ost_z = opt.init(Zs)
        times = []
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            Zs, ost_z, loss_z = train_inducing.optimize_step(Zs, Xb, lst, ALPHA, ost_z, it, opt, None, "classifier", full_set_size=N_FULL,
                                                             scalable=True, st_samples=nb, slq_samples=2, slq_num_matvecs=k_ref)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        train = {"zgrad_ms": ms_z, "zgrad_probe_pairs": nb, "zgrad_points": M_POINTS,
                 "zgrad_note": "lip_zgrad GGN mode, cotangent = vector = 64 Rademacher probes: d/dZ of their quadratic forms (tcgen05 3xTF32 GEMMs for the wide layers)",
                 "optimize_step_seconds": min(times[1:]), "optimize_step_loss": float(loss_z),
                 "optimize_step_config": f"train_inducing.optimize_step, scalable objective + exact dZ + Adam: m={m_ref} inducing points, "
                                         f"|X|=256, st_samples={nb}, slq k={k_ref} x 2 probes (config/scale/mlp_mnist.yml sizes)"}

    # ---- BASELINE configs 4 and 5 under the same clock: C4 LeNet5 probe-sharded, C5 ResNet1M (M = 4096) point-sharded ----
    extras = None
    if not args.no_extra and args.workload == "mlp":
        del V
        torch.cuda.empty_cache()
        extras = {}
        try:
            extras["c4_lenet5"] = measure_extra("lenet5", 0, 256, 5, 3, False, torch, dist, world, rank, dev)
            extras["c5_resnet1m_m4096"] = measure_extra("resnet1m", 4096, 4, 3, 3, True, torch, dist, world, rank, dev)
        except Exception as e:              # noqa: BLE001  (the headline line must still be printed)
            extras["error"] = f"{type(e).__name__}: {e}"
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    tf32_peak_post = measure_tf32_peak(torch)
    tf32_peak = max(tf32_peak_pre, tf32_peak_post)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    achieved = FLOP_PER_PRODUCT * B / (ms_ggn * 1e-3) / 1e12
    # DRAM bytes of one lip_ggn_vp call from the committed ncu pass (dram__bytes_read.sum + dram__bytes_write.sum summed
    # over the call's launches); only valid for the configuration it was captured on
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic_ggn_vp.json")))
        if args.workload == "mlp" and B == 256:
            traffic = tj["traffic_bytes"]
    except Exception:
        pass
    peak = tf32_peak / 3.0
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic,
                "kernel": "one lip_ggn_vp call = the JVP + VJP GEMM sweeps over all probes (tcgen05 gemm_tc*_kernel launches "
                          "+ split / head / bias kernels), timed alone with CUDA events",
                "ms_per_call": ms_ggn,
                "ms_per_call_gaussian_probes": ms_gauss,
                "frac_gaussian_probes": FLOP_PER_PRODUCT * B / (ms_gauss * 1e-3) / 1e12 / peak,
                "peak_source": f"cuBLAS TF32 8192^3 best-of-10 measured in this run (before / after the timed loops: "
                               f"{tf32_peak_pre:.1f} / {tf32_peak_post:.1f}, larger used) = {tf32_peak:.1f} TFLOP/s, divided by 3 "
                               f"(3xTF32 emulated fp32); bf16 of MEASURED_PEAKS.json = {peaks.get('bf16_tflops')}",
                "algorithmic_flop_per_launch_group": FLOP_PER_PRODUCT * B}
    cpu = None
    if not args.no_cpu and world == 1:
        pps, cores, sec = cpu_reference_products_per_s(ost, Z, 1, steps=3, warmup=1)
        cpu = {"value": pps, "unit": "products/s", "cores": cores, "kind": "port",
               "sample": f"3 timed passes of 1 product (all {M_POINTS} points) after 1 warm-up pass, torch fp32 restatement of "
                         f"src/ggn.py:133-144 (JAX not installed), {sec:.1f} s per product; same function as --impl reference"}
    line = {"metric": "ggn_vec_products_per_s", "value": value, "unit": "products/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, world), "path": path,
            "clocks": sampler.summary(), "e2e": e2e, "e2e_fp32_vectors": e2e_fp32, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "slq_logdet": slq, "train_step": train, "hutchinson_trace_estimate": trace_est,
            "extra_workloads": extras}
    if extras and tf32_peak > 0:
        hbm = peaks.get("hbm_gbs")
        c4, c5 = extras.get("c4_lenet5"), extras.get("c5_resnet1m_m4096")
        if c5:
            c5["roofline"] = {"bound": "tensor", "achieved": c5["algorithmic_tflops"], "peak": peak, "unit": "TFLOP/s",
                              "frac": c5["algorithmic_tflops"] / peak, "note": "per GPU; peak = the headline's TF32 / 3"}
        if c4:
            # LeNet5 (channels 1 / 6 / 16) cannot feed the tensor cores; its yardsticks are the fp32 FMA pipes and HBM.  Bytes: the tangents and
            # cotangents of every stage written and read once per (probe, point) — 2 x 8,094 floats x 8 B — plus 8 D per product.
            # With the fused stage kernels (lip_cnn_fused.cu) only the POOLED tangents / cotangents cross HBM.
            per_pair = 2 * (1176 + 400 + 120 + 84 + 10) * 8
            gbps = (per_pair * 200 + 8 * 61_706) * c4["probes_per_step"] / (c4["ms_per_step"] * 1e-3) / 1e9 / world
            clk = (line["clocks"] or {}).get("sm_max_mhz") or 1965.0
            fma_peak = 148 * 128 * 2 * clk * 1e6 / 1e12
            c4["roofline"] = {"bound": "fp32-fma", "achieved": c4["algorithmic_tflops"], "peak": fma_peak, "unit": "TFLOP/s",
                              "frac": c4["algorithmic_tflops"] / fma_peak,
                              "hbm": {"achieved": gbps, "peak": hbm, "unit": "GB/s", "frac": (gbps / hbm) if hbm else None,
                                      "bytes_per_probe_point": per_pair},
                              "note": "per GPU; peak = 148 SMs x 128 FMA lanes x 2 x max SM clock (exact fp32: 1 / 6 / 16 channels are below any "
                                      "tensor-core tile; the 400-120-84 dense tail, 17 % of the FLOP, runs as tcgen05 3xTF32 GEMMs); hbm = pooled tangent + "
                                      "cotangent of each stage written once and read once + 8 D per product"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of this run, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout must carry exactly one JSON line, but NCCL prints its version banner to fd 1 at communicator creation (and other
    # native libraries may chat too): point fd 1 at stderr for the whole run and keep the original stdout for emit().
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    set_workload(args.workload)
    if args.points > 0:
        global M_POINTS, FLOP_PER_PRODUCT, WORKLOAD
        FLOP_PER_PRODUCT = FLOP_PER_PRODUCT // M_POINTS * args.points
        WORKLOAD = WORKLOAD.replace(f"M={M_POINTS}", f"M={args.points}")
        M_POINTS = args.points
    if args.workload != "mlp":
        args.no_slq = True          # the SLQ leg is defined on the headline MLP
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
