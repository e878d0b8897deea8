#!/usr/bin/env python
"""bench.py -- the hot path's headline number on B200, one JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                       # the CPU golden (oracle port) on the host cores
    python bench.py --all-shapes                                 # extra: every README shape, N=1 latency + N=256

Workload (BASELINE.json configs[1] shape at the configs[3] batch): fused 3x3 conv (Winograd F(2x2,3x3)) + BN + ReLU,
Cin = Cout = 256, 14x14 maps (16x16 frames), 256 images per GPU per step, FP32 in/out, TF32 tensor-core arithmetic,
synthetic seeded data in the reference's data_generator.py distributions. A step = ONE kernel launch over the batch.
Batch-sharded across ranks with no collective on the hot path (weak scaling: 256 images per GPU).

  value      images/s, inputs resident in HBM, CUDA events on the launch stream, max over ranks
  e2e        images/s through the C-ABI call with HOST buffers (wg_run_host: pinned H2D + kernel + D2H every step)
  roofline   direct-conv-equivalent FLOPs per launch / measured launch time vs the tensor peak
  cpu_baseline  the NumPy FP32 golden (9 shifted sgemms + BN + ReLU, oracle/golden.py) timed on this box's host cores
Extra keys (reported, outside the timed region): all_shapes (six README shapes, N=1 and N=256), strong_scaling
(BASELINE configs[3]: global N=256 sharded over the ranks, 128/128 and 256/256, TF32 and bf16), bottleneck_block
(configs[4]: 1x1 -> 3x3 -> 1x1 + residual + ReLU), cudnn_baseline (cuDNN on the same B200, reported baseline only),
tensor_peak (the tcgen05 dense peaks measured on this device in this run: the roofline denominator).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C_IN = C_OUT = 256
IMAGES_PER_GPU = 256
FLOP_PER_IMAGE = 2 * 196 * C_IN * C_OUT * 9            # direct-conv equivalent, SURVEY.md section 8d: 231.21 MFLOP
BYTES_PER_IMAGE = (256 * C_IN + 196 * C_OUT) * 4       # compulsory fp32 in + out
WEIGHT_BYTES = C_OUT * C_IN * 9 * 4 + 2 * C_OUT * 4
METRIC = "fused_conv3x3_bn_relu_256x256_images_per_s"
UNIT = "images/s"
N_SETS = 4                                             # rotating buffer sets: 4 x 118 MB > 126 MB L2


def workload_name():
    return ("fused conv3x3+BN+ReLU (the reference's Winograd layer kernel_256) Cin=Cout=256 14x14 (16x16 frames), "
            "N=256 images/GPU/step, fp32 I/O, tf32 MMA [BASELINE configs[1] shape at configs[3] batch]; this repo "
            "computes it as a direct convolution on tcgen05 at this batch size, F(2x2,3x3) Winograd below 6 images")


def make_params(seed=0):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    rs = np.random.RandomState(seed)
    w = (rs.rand(C_OUT, C_IN, 3, 3) - 0.5).astype(np.float32)
    gamma, beta, mean = ((rs.rand(C_OUT) - 0.5).astype(np.float32) for _ in range(3))
    var = (rs.rand(C_OUT) * 3 + 5).astype(np.float32)
    sd = np.sqrt(var + np.float32(1e-5))
    return w, (gamma / sd).astype(np.float32), (beta - gamma * mean / sd).astype(np.float32)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def measure_tensor_peaks(wg, dev_index):
    """Roofline denominators measured on this device in this run: back-to-back tcgen05.mma on
    every SM (wg_measure_tensor_peak; ms-long launches, i.e. the same burst regime the kernel launches are timed in),
    next to cuBLAS TF32 / bf16 8192^3 matmuls through torch as a cross-check (reported only)."""
    import torch
    out = {}
    for name, dt in (("tf32", wg.WG_TF32), ("bf16", wg.WG_BF16)):
        tf, clk = wg.measure_tensor_peak(dt, dev_index)
        out[f"{name}_tcgen05_tflops"] = tf
        out[f"{name}_clk_per_mma_m128_n256"] = clk
    try:
        dev = torch.device("cuda", dev_index)
        a = torch.randn(8192, 8192, device=dev)
        b = torch.randn(8192, 8192, device=dev)
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        best = {}
        for name, (x, y) in (("tf32", (a, b)), ("bf16", (a.bfloat16(), b.bfloat16()))):
            torch.matmul(x, y)
            torch.cuda.synchronize(dev)
            t_best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(x, y)
                e1.record()
                torch.cuda.synchronize(dev)
                t_best = min(t_best, e0.elapsed_time(e1))
            best[name] = 2 * 8192.0 ** 3 / (t_best * 1e-3) / 1e12
        torch.backends.cuda.matmul.allow_tf32 = old
        out["tf32_cublas_8192_tflops"] = best["tf32"]
        out["bf16_cublas_8192_tflops"] = best["bf16"]
        del a, b
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["cublas_note"] = str(e)[:120]
    out["how"] = ("tcgen05: wg_measure_tensor_peak -- 20000 back-to-back M=128 N=256 MMAs per SM, operands in shared "
                  "memory, best of 3 launches, CUDA events; cublas: torch.matmul 8192^3 best of 5 (allow_tf32)")
    return out


def traffic_from_profiles():
    """Per-launch memory traffic of the headline kernel from THIS round's ncu capture (profiles/traffic_r02.json, written
    by tools/ncu_summary.py from the committed raw CSV of an `ncu --set full` run with rotating buffers)."""
    tpath = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if not os.path.exists(tpath):
        return None, None
    d = json.load(open(tpath))
    k = d.get("conv3x3_256_n256", {})
    return k.get("dram_bytes_per_launch"), k


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU while the timed region runs (NVML in-process, ~2 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(s))


# ----------------------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import numpy as np
    import torch
    import wg_loader
    wg = wg_loader.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback for the product arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    w, scale, shift = make_params()
    layer = wg.Conv3x3BnRelu(w, scale, shift, relu=True, device=local)
    n = IMAGES_PER_GPU
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    xs = [torch.rand((n, 16, 16, C_IN), device=dev, generator=g) - 0.5 for _ in range(N_SETS)]
    ys = [torch.empty((n, 14, 14, C_OUT), device=dev) for _ in range(N_SETS)]
    stream = torch.cuda.current_stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        layer(xs[i % N_SETS], out=ys[i % N_SETS])
    barrier()

    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = wg.launch_count()
    sampler.start()
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        layer(xs[i % N_SETS], out=ys[i % N_SETS])
    e1.record(stream)
    barrier()
    launches = wg.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    if clocks["samples"] == 0 and sampler.ok:
        # timed region shorter than one sampling period: sample the same loop again, untimed, to report clocks
        sampler = ClockSampler(local)
        sampler.start()
        t_end = time.time() + 0.3
        while time.time() < t_end:
            for i in range(50):
                layer(xs[i % N_SETS], out=ys[i % N_SETS])
            torch.cuda.synchronize(dev)
        clocks = sampler.stop()
        clocks["note"] = "timed region shorter than the sampling period; sampled over an identical untimed loop after it"

    # ---- e2e: host buffers through wg_run_host (pinned H2D + kernel + D2H each step)
    xh = [torch.empty((n, 16, 16, C_IN), pin_memory=True).copy_(xs[i].cpu()) for i in range(2)]
    yh = torch.empty((n, 14, 14, C_OUT), pin_memory=True)
    e2e_steps = max(3, min(args.steps, 20))
    for i in range(2):
        layer.run_host_ptr(xh[i % 2].data_ptr(), yh.data_ptr(), n)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        layer.run_host_ptr(xh[i % 2].data_ptr(), yh.data_ptr(), n)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    # the step's result is read on the host: checksum of the whole host output, next to the device arm's on the same input
    last = (e2e_steps - 1) % 2
    host_sum = float(yh.double().sum())
    host_abs = float(yh.double().abs().sum())
    dev_y = layer(xs[last], out=ys[0])
    dev_sum = float(dev_y.double().sum())
    dev_maxdiff = float((dev_y.cpu() - yh).abs().max())
    e2e_check = {"host_sum": host_sum, "device_sum": dev_sum, "sum_abs": host_abs,
                 "rel_diff": abs(host_sum - dev_sum) / max(host_abs, 1e-30), "max_abs_diff_vs_device_arm": dev_maxdiff,
                 "note": "wg_run_host output (chunks of 16,32,64,64,40,20,20 images, each through the kernel its size "
                         "selects) vs the one-launch device arm on the same input; differences are fp32 summation order"}
    # PCIe ceiling of this box for the same bytes: pinned H2D and D2H copies running concurrently on two streams
    pcie = None
    try:
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        xd_tmp, yd_tmp = torch.empty_like(xs[0]), ys[0]
        for _ in range(2):
            with torch.cuda.stream(s_in):
                xd_tmp.copy_(xh[0], non_blocking=True)
            with torch.cuda.stream(s_out):
                yh.copy_(yd_tmp, non_blocking=True)
        barrier()                                   # all ranks copy at the same time: the ceiling under contention
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s_in):
                xd_tmp.copy_(xh[0], non_blocking=True)
            with torch.cuda.stream(s_out):
                yh.copy_(yd_tmp, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / reps
        pcie = {"ms_per_step_copies_only": dt * 1e3, "images_per_s_ceiling": n / dt,
                "h2d_gbs": n * 256 * C_IN * 4 / dt / 1e9, "d2h_gbs": n * 196 * C_OUT * 4 / dt / 1e9,
                "how": "one 67 MB pinned H2D and one 51 MB pinned D2H copy per step, concurrently on two streams, "
                       "no kernel, all ranks at the same time: the ceiling of any host-buffer path on this box"}
        del xd_tmp
    except Exception as e:  # noqa: BLE001
        pcie = {"unavailable": str(e)[:120]}

    # ---- optional: the one exchange step, an all-gather of the output shards (north_star), timed separately
    gather_ms = None
    bare_gather_ms = None
    fused_ms, fused_note = None, None
    if dist is not None:
        out_all = torch.empty((world * n, 14, 14, C_OUT), device=dev)
        for _ in range(2):
            dist.all_gather_into_tensor(out_all, ys[0])
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for i in range(5):
            layer(xs[i % N_SETS], out=ys[i % N_SETS])
            dist.all_gather_into_tensor(out_all, ys[i % N_SETS])
        g1.record(stream)
        barrier()
        gather_ms = g0.elapsed_time(g1) / 5
        # the exchange alone (no kernel): what NCCL's all-gather of the same 51 MB shards costs on this box's NVLink
        barrier()
        g0.record(stream)
        for i in range(10):
            dist.all_gather_into_tensor(out_all, ys[i % N_SETS])
        g1.record(stream)
        barrier()
        bare_gather_ms = g0.elapsed_time(g1) / 10
        # the same exchange fused into the kernel: output stores go to an NVLS multicast address (multimem.st), the
        # switch replicates each shard into every GPU's copy of the gathered tensor; one cross-rank barrier per step
        try:
            fused = wg.FusedGatherConv3x3(layer, n)
            for i in range(3):
                fused(xs[i % N_SETS])
            barrier()
            g0.record(stream)
            for i in range(20):
                fused(xs[i % N_SETS])
            g1.record(stream)
            barrier()
            fused_ms = g0.elapsed_time(g1) / 20
        except Exception as e:  # noqa: BLE001  (no multicast support on this system)
            fused_note = str(e)

    if dist is not None:
        t = torch.tensor([ms, e2e_s, gather_ms or 0.0, fused_ms or 0.0, bare_gather_ms or 0.0], device=dev,
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, gmax, fmax, bare_gather_ms = (float(v) for v in t.tolist())
        gather_ms = gmax if gather_ms is not None else None
        fused_ms = fmax if fused_ms is not None else None

    # ---- BASELINE configs[3] (strong scaling: global N=256 sharded over the ranks) and configs[4] (bottleneck block)
    del xh
    strong = strong_scaling(wg, dev, dist, world, rank)
    block = bottleneck_block_bench(wg, dev, dist, world)

    if rank == 0:
        pk = peaks()
        tpk = measure_tensor_peaks(wg, local)
        ms_per_step = ms / args.steps
        value = world * n * args.steps / (ms * 1e-3)
        tflops = FLOP_PER_IMAGE * n / (ms_per_step * 1e-3) / 1e12
        tf32_peak = tpk["tf32_tcgen05_tflops"]
        traffic, traffic_detail = traffic_from_profiles()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": workload_name(), "images_per_gpu": n, "global_batch": world * n,
                       "parallelism": f"batch-sharded x{world}, no collective on the hot path",
                       "l2": f"rotating {N_SETS} input/output sets ({N_SETS * n * BYTES_PER_IMAGE / 1e6:.0f} MB) > 126 MB L2",
                       "peaks": pk["source"]},
            "clocks": clocks,
            "e2e": {"value": world * n * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * 256 * C_IN * 4,
                    "d2h_bytes_per_step": n * 196 * C_OUT * 4, "steps": e2e_steps, "api": "wg_run_host (C-ABI)",
                    "result_check": e2e_check, "pcie_ceiling": pcie,
                    "frac_of_pcie_ceiling": (n * e2e_steps / e2e_s) / pcie["images_per_s_ceiling"]
                    if pcie and "images_per_s_ceiling" in pcie else None},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "tensor", "achieved": tflops, "peak": tf32_peak, "unit": "TFLOP/s",
                         "frac": tflops / tf32_peak, "traffic": traffic,
                         "kernel": "conv3x3_direct_kernel (direct convolution on tcgen05: couts on M, one image's 14 frame rows on "
                                   "N = 224, nine taps = nine shifted shared-memory descriptors of one activation box)",
                         "algorithmic": "2*196*256*256*9 = 231.21 MFLOP/image x 256 images/launch (the kernel executes 224/196 "
                                        "of that: the frame's two border columns ride along)",
                         "peak_note": "dense TF32 tcgen05 peak MEASURED in this run on this device (tensor_peak.tf32_tcgen05_tflops: "
                                      "back-to-back kind::tf32 MMAs on every SM, ms-long launch = the burst regime the kernel "
                                      "launches are timed in); MEASURED_PEAKS.json holds no TF32 figure",
                         "frac_of_half_bf16_burst": tflops / (pk["bf16_burst"] / 2.0),
                         "frac_of_half_bf16_sustained": tflops / (pk["bf16_sustained"] / 2.0),
                         "traffic_detail": traffic_detail,
                         "hbm": {"achieved_gbs": (BYTES_PER_IMAGE * n + WEIGHT_BYTES) / (ms_per_step * 1e-3) / 1e9,
                                 "peak_gbs": pk["hbm"],
                                 "frac": (BYTES_PER_IMAGE * n + WEIGHT_BYTES) / (ms_per_step * 1e-3) / 1e9 / pk["hbm"]}},
            "tensor_peak": tpk,
            "strong_scaling": strong,
            "bottleneck_block": block,
        }
        if gather_ms is not None:
            shard_mb = n * 196 * C_OUT * 4 / 1e6
            line["with_output_allgather"] = {
                "ms_per_step": gather_ms, "value": world * n / (gather_ms * 1e-3), "unit": UNIT,
                "collective": "nccl all_gather_into_tensor of fp32 output",
                "nvlink_ceiling": {"bare_allgather_ms": bare_gather_ms,
                                   "ingress_gbs_per_gpu": (world - 1) * shard_mb / bare_gather_ms,
                                   "how": f"the same all-gather alone, no kernel: each GPU receives {world - 1} x "
                                          f"{shard_mb:.0f} MB per step; kernel + gather cannot beat max(kernel, this)"},
                "frac_of_ceiling": max(bare_gather_ms, ms / args.steps) / gather_ms}
            if fused_ms is not None:
                line["with_output_allgather"]["fused_multicast"] = {
                    "ms_per_step": fused_ms, "value": world * n / (fused_ms * 1e-3), "unit": UNIT,
                    "frac_of_ceiling": max(bare_gather_ms, ms / args.steps) / fused_ms,
                    "how": "same kernel, output stores as multimem.st to an NVLS multicast address (WG_OUT_MULTICAST) "
                           "+ one symmetric-memory barrier per step; no NCCL call"}
            elif fused_note:
                line["with_output_allgather"]["fused_multicast"] = {"unavailable": fused_note[:200]}
        if world == 1:
            # BASELINE.json's metric has two halves: us/layer at N=1 and throughput at N=256, for the six README
            # shapes. The headline above is one of them; the rest ride along (outside the timed region, 60 launches each).
            del xs, ys, yh
            torch.cuda.empty_cache()
            line["all_shapes"] = [
                {"layer": f"{r['kind']} {r['cin']}->{r['cout']}" + ("+relu" if r["relu"] else ""), "n": r["n"],
                 "us_per_layer": round(r["us_per_layer"], 2), "images_per_s": round(r["images_per_s"]),
                 "frac_tf32_peak": round(r["frac_tf32_peak"], 4), "frac_hbm_peak": round(r["frac_hbm_peak"], 4)}
                for r in measure_shapes(wg, dev, iters=60, sets_n256=3, verbose=False, tf32_peak=tf32_peak)]
            line["cudnn_baseline"] = cudnn_baseline(dev)
            line["cpu_baseline"] = cpu_baseline(budget_s=12.0)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------- BASELINE configs[3] / configs[4] / cuDNN
def _time_loop(torch, dev, fn, sets, iters, dist=None, graph=False):
    """CUDA events around `iters` back-to-back calls (rotating `sets` buffers), barrier on both sides, max over ranks.
    graph=True: the `iters` calls are captured once into a CUDA graph and the replay is timed -- for steps of a few
    microseconds, where issuing the launches from Python (10-20 us per call, more with eight ranks on one host) would be
    what is measured. Kernels only (no collectives inside `fn`); falls back to the eager loop if capture fails."""
    for i in range(4):
        fn(i % sets)
    torch.cuda.synchronize(dev)
    g = None
    if graph:
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for i in range(iters):
                        fn(i % sets)
            torch.cuda.current_stream(dev).wait_stream(side)
            g.replay()                      # one untimed replay (graph upload)
            torch.cuda.synchronize(dev)
        except Exception:  # noqa: BLE001
            g = None
            torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if g is not None:
        g.replay()
    else:
        for i in range(iters):
            fn(i % sets)
    e1.record()
    torch.cuda.synchronize(dev)
    us = e0.elapsed_time(e1) * 1e3 / iters
    if dist is not None:
        t = torch.tensor([us], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        us = float(t.item())
    return us


def strong_scaling(wg, dev, dist, world, rank):
    """BASELINE.json configs[3]: 3x3 128/128 and 256/256 at GLOBAL batch N=256 sharded over the ranks (256 / world images
    per GPU), TF32 and bf16 operands, compute only (no collective on the hot path; 60 launches replayed from one CUDA graph
    so that the few-microsecond shards are not timed through Python's launch overhead) and -- 256/256 TF32, world > 1 -- with
    the output all-gather fused into the kernel. At world == 1 the shard sizes of 2 / 4 / 8 GPUs are also timed on this
    one GPU: shards are independent, so that IS the per-GPU time of the sharded run (efficiency_projected); the driver's
    multi-GPU runs report the measured time at their own world size (efficiency vs the same line's n=256 single shard)."""
    import numpy as np
    import torch
    out = {"global_batch": 256, "world": world, "rows": []}
    shards = [256, 128, 64, 32] if world == 1 else [256, 256 // world]
    for c in (128, 256):
        rs = np.random.RandomState(c)
        w = (rs.rand(c, c, 3, 3) - 0.5).astype(np.float32)
        sc, sh = (rs.rand(c) + 0.5).astype(np.float32), (rs.rand(c) - 0.5).astype(np.float32)
        for name, dt in (("tf32", wg.WG_TF32), ("bf16", wg.WG_BF16)):
            layer = wg.Conv3x3BnRelu(w, sc, sh, relu=True, device=dev.index, dtype=dt)
            base_us = None
            for nl in shards:
                sets = max(2, min(8, int(300e6 // (nl * (256 * c + 196 * c) * 4)) + 1))
                xs = [torch.rand((nl, 16, 16, c), device=dev) - 0.5 for _ in range(sets)]
                ys = [torch.empty((nl, 14, 14, c), device=dev) for _ in range(sets)]
                us = _time_loop(torch, dev, lambda i: layer(xs[i], out=ys[i]), sets, 60, dist, graph=True)
                if nl == 256:
                    base_us = us
                row = {"layer": f"3x3 {c}->{c}", "dtype": name, "images_per_gpu": nl, "gpus": 256 // nl,
                       "us_per_step": round(us, 2), "images_per_s_global": round(256 / (us * 1e-6)),
                       "speedup_vs_1gpu": round(base_us / us, 3) if base_us else None,
                       "efficiency": round(base_us / us / (256 // nl), 3) if base_us else None,
                       "measured_on": f"{world} GPU(s)" + ("" if world > 1 or nl == 256 else
                                                             " (one shard of the sharded run, timed on this GPU)")}
                if world > 1 and nl != 256 and c == 256 and name == "tf32":
                    try:
                        fused = wg.FusedGatherConv3x3(layer, nl)
                        row["with_fused_gather_us"] = round(_time_loop(torch, dev, lambda i: fused(xs[i]), sets, 30, dist), 2)
                        gathered = torch.empty((world * nl, 14, 14, c), device=dev)

                        def nccl_step(i):
                            layer(xs[i], out=ys[i])
                            dist.all_gather_into_tensor(gathered, ys[i])
                        row["with_nccl_allgather_us"] = round(_time_loop(torch, dev, nccl_step, sets, 30, dist), 2)
                        del fused, gathered
                    except Exception as e:  # noqa: BLE001
                        row["fused_gather_note"] = str(e)[:160]
                out["rows"].append(row)
                del xs, ys
            layer.close()
    torch.cuda.empty_cache()
    return out


def bottleneck_block_bench(wg, dev, dist, world):
    """BASELINE.json configs[4]: the full ResNet-50 bottleneck residual block 1x1 -> 3x3 -> 1x1 (+ x, ReLU), three fused
    launches (cuda_winograd_b200.Bottleneck(residual=True)), 256 images per GPU per step (weak) and global N=256 sharded
    (strong), max over ranks; at world == 1 next to cuDNN's fused conv+bias+ReLU chain (+ add + ReLU) through torch."""
    import numpy as np
    import torch
    rows = []
    for ch, c in ((512, 128), (1024, 256)):
        rs = np.random.RandomState(ch)
        w1 = ((rs.rand(ch, c) - 0.5) * 0.2).astype(np.float32)
        w3 = ((rs.rand(c, c, 3, 3) - 0.5) * 0.2).astype(np.float32)
        w2 = ((rs.rand(c, ch) - 0.5) * 0.2).astype(np.float32)
        bn = [((rs.rand(k) + 0.5).astype(np.float32), (rs.rand(k) - 0.3).astype(np.float32)) for k in (c, c, ch)]
        block = wg.Bottleneck(w1, *bn[0], w3, *bn[1], w2, *bn[2], device=dev.index, residual=True)
        flop_img = 2.0 * 196 * (ch * c + 9 * c * c + c * ch)
        for nl in sorted({256, 256 // world}, reverse=True):
            sets = 2 if nl == 256 else 4
            xs = [torch.rand((nl, 196, ch), device=dev) - 0.5 for _ in range(sets)]
            outs = [torch.empty((nl, 196, ch), device=dev) for _ in range(sets)]
            us = _time_loop(torch, dev, lambda i: block(xs[i], out=outs[i]), sets, 30, dist, graph=True)
            row = {"block": f"{ch}->{c}->{c}->{ch} +residual", "images_per_gpu": nl, "gpus": world,
                   "scaling": "weak" if nl == 256 else "strong (global N=256)", "us_per_step": round(us, 1),
                   "images_per_s": round(world * nl / (us * 1e-6)),
                   "tflops_direct_equiv_per_gpu": round(flop_img * nl / us * 1e-6, 1), "launches_per_step": 3}
            if world == 1:
                try:
                    row["cudnn_tf32_us"] = round(_cudnn_block_us(torch, dev, xs, w1, w3, w2, bn, nl, ch, c), 1)
                except Exception as e:  # noqa: BLE001
                    row["cudnn_note"] = str(e)[:160]
            rows.append(row)
            del xs, outs
        del block
    torch.cuda.empty_cache()
    return rows


def _cudnn_block_us(torch, dev, xs, w1, w3, w2, bn, n, ch, c):
    """cuDNN on the same GPU (reported baseline only): fused conv+bias+ReLU x2, conv+bias, add, ReLU; channels_last."""
    cl = torch.channels_last
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    k1 = torch.from_numpy((w1.T * bn[0][0][:, None])[:, :, None, None].copy()).to(dev).contiguous(memory_format=cl)
    k3 = torch.from_numpy(w3 * bn[1][0][:, None, None, None]).to(dev).contiguous(memory_format=cl)
    k2 = torch.from_numpy((w2.T * bn[2][0][:, None])[:, :, None, None].copy()).to(dev).contiguous(memory_format=cl)
    b1, b3, b2 = (torch.from_numpy(b[1]).to(dev) for b in bn)
    x_cl = [t.view(n, 14, 14, ch).permute(0, 3, 1, 2) for t in xs]

    def chain(i):
        a = torch.cudnn_convolution_relu(x_cl[i], k1, b1, (1, 1), (0, 0), (1, 1), 1)
        a = torch.cudnn_convolution_relu(a, k3, b3, (1, 1), (1, 1), (1, 1), 1)
        return torch.relu_(torch.nn.functional.conv2d(a, k2, b2).add_(x_cl[i]))
    return _time_loop(torch, dev, chain, len(xs), 30)


def cudnn_baseline(dev):
    """Reported baseline only (north_star: "cuDNN's Winograd/implicit-GEMM fused conv+BN+ReLU on the same B200"): cuDNN
    through torch (cudnn_convolution_relu = cudnnConvolutionBiasActivationForward, BN scale folded into the weights,
    shift as bias, channels_last), the six README shapes at N=256 (rotating buffers) and N=1 (CUDA graph of back-to-back
    launches), TF32 allowed; 3x3 also with bf16 tensors. Never on the product path: nothing under cuda-winograd_b200/
    links or calls cuDNN."""
    import numpy as np
    import torch
    rows = []
    try:
        torch.backends.cudnn.benchmark = True
        for kind, cin, cout, relu in [("3x3", 128, 128, True), ("3x3", 256, 256, True), ("1x1", 512, 128, True),
                                      ("1x1", 128, 512, False), ("1x1", 1024, 256, True), ("1x1", 256, 1024, False)]:
            rs = np.random.RandomState(0)
            ks = 3 if kind == "3x3" else 1
            wt = torch.from_numpy((rs.rand(cout, cin, ks, ks) - 0.5).astype(np.float32)).to(dev)
            wt = wt.contiguous(memory_format=torch.channels_last)
            bias = torch.from_numpy((rs.rand(cout) - 0.5).astype(np.float32)).to(dev)
            hw = 16 if kind == "3x3" else 14
            row = {"layer": f"{kind} {cin}->{cout}" + ("+relu" if relu else "")}
            for n in (256, 1):
                sets = 3 if n == 256 else 1
                xs = [(torch.rand((n, hw, hw, cin), device=dev) - 0.5).permute(0, 3, 1, 2) for _ in range(sets)]
                for name, dt in (("tf32", torch.float32),) + ((("bf16", torch.bfloat16),) if kind == "3x3" else ()):
                    torch.backends.cudnn.allow_tf32 = True
                    xq = [t.to(dt) for t in xs] if dt != torch.float32 else xs
                    wq, bq = wt.to(dt), bias.to(dt)
                    if relu:
                        f = lambda i: torch.cudnn_convolution_relu(xq[i], wq, bq, (1, 1), (0, 0), (1, 1), 1)  # noqa: E731
                    else:
                        f = lambda i: torch.nn.functional.conv2d(xq[i], wq, bq)  # noqa: E731
                    if n == 256:
                        us = _time_loop(torch, dev, f, sets, 40)
                    else:
                        for _ in range(3):
                            f(0)
                        torch.cuda.synchronize(dev)
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            for _ in range(100):
                                f(0)
                        us = _time_loop(torch, dev, lambda i: g.replay(), 1, 5) / 100
                        del g
                    row[f"cudnn_{name}_n{n}_us"] = round(us, 2)
                del xs
            rows.append(row)
        torch.cuda.empty_cache()
        return {"cudnn_version": torch.backends.cudnn.version(), "rows": rows,
                "note": "reported baseline only; same process, same GPU, CUDA events; N=256 rotating buffers, N=1 CUDA graph"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:200], "rows": rows}


# ----------------------------------------------------------------------------------------------------------- CPU arm
def _cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        return int(n)
    except Exception:
        return os.cpu_count() or 1


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is asked to use every host thread it can."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def cpu_step(x, w, scale, shift):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import golden
    return golden.conv3x3_bn_relu_fp32(x, w, scale, shift, relu=True)


def cpu_baseline(budget_s=12.0, sample_images=32):
    """The NumPy FP32 golden on the host cores over a bounded sample of the same workload."""
    import numpy as np
    _use_all_host_threads()
    w, scale, shift = make_params()
    rs = np.random.RandomState(1)
    x = (rs.rand(sample_images, 16, 16, C_IN) - 0.5).astype(np.float32)
    cpu_step(x[:2], w, scale, shift)
    t0 = time.perf_counter()
    reps = 0
    while True:
        cpu_step(x, w, scale, shift)
        reps += 1
        el = time.perf_counter() - t0
        if el > budget_s or reps >= 200:
            break
    return {"value": reps * sample_images / el, "unit": UNIT, "cores": _cpu_threads(), "kind": "port",
            "sample": f"{reps} x {sample_images} images of the same 256->256 3x3 layer, NumPy fp32 (9 shifted sgemms + "
                      f"BN + ReLU, oracle/golden.py), {el:.1f} s", "host_cpus": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference has no CPU implementation and its CUDA/cuDNN program is not the CPU arm asked
    for here; the arm is the oracle port (NumPy FP32 golden) with all host threads, same metric/config. Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; BLAS reads it when numpy is first imported (nothing imported it yet)
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(os.cpu_count() or 1)
    import numpy as np
    _use_all_host_threads()
    w, scale, shift = make_params()
    sample = 64
    rs = np.random.RandomState(1)
    x = (rs.rand(sample, 16, 16, C_IN) - 0.5).astype(np.float32)
    for _ in range(max(1, min(args.warmup, 3))):
        cpu_step(x, w, scale, shift)
    steps = max(1, min(args.steps, 50))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(x, w, scale, shift)
    el = time.perf_counter() - t0
    value = steps * sample / el
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(1, min(args.warmup, 3)), "ms_per_step": el / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "sample_images_per_step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": _cpu_threads(), "kind": "port",
                             "sample": f"{steps} steps x {sample} images, NumPy fp32 golden (oracle/golden.py)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- all README shapes
def measure_shapes(wg, dev, iters=200, sets_n256=N_SETS, verbose=True, tf32_peak=None):
    """Every README shape: N=1 latency (L2-warm, like the reference's loop) and N=256 throughput (rotating buffers),
    CUDA events. Returns one row per (shape, N)."""
    import numpy as np
    import torch
    pk = peaks()
    if tf32_peak is None:
        tf32_peak = wg.measure_tensor_peak(wg.WG_TF32, dev.index)[0]
    rows = []
    shapes = [("3x3", 128, 128, True), ("3x3", 256, 256, True), ("1x1", 512, 128, True), ("1x1", 128, 512, False),
              ("1x1", 1024, 256, True), ("1x1", 256, 1024, False)]
    for kind, cin, cout, relu in shapes:
        rs = np.random.RandomState(0)
        if kind == "3x3":
            layer = wg.Conv3x3BnRelu((rs.rand(cout, cin, 3, 3) - 0.5).astype(np.float32),
                                     rs.rand(cout).astype(np.float32), rs.rand(cout).astype(np.float32), relu,
                                     device=dev.index)
            in_shape, out_shape, taps = (16, 16, cin), (14, 14, cout), 9
        else:
            layer = wg.Conv1x1Bn((rs.rand(cin, cout) - 0.5).astype(np.float32), rs.rand(cout).astype(np.float32),
                                 rs.rand(cout).astype(np.float32), relu, device=dev.index)
            in_shape, out_shape, taps = (196, cin), (196, cout), 1
        for n in (1, 256):
            sets = 1 if n == 1 else sets_n256
            xs = [torch.rand((n,) + in_shape, device=dev) - 0.5 for _ in range(sets)]
            ys = [torch.empty((n,) + out_shape, device=dev) for _ in range(sets)]
            for i in range(5):
                layer(xs[i % sets], out=ys[i % sets])
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(iters):
                layer(xs[i % sets], out=ys[i % sets])
            e1.record()
            torch.cuda.synchronize(dev)
            us = e0.elapsed_time(e1) * 1e3 / iters
            us_hostloop = us
            if n == 1:
                # a ~6 us layer is shorter than one Python/ctypes call: time the same back-to-back launches captured
                # once into a CUDA graph (what a C caller's loop, like the reference's Test.c, sees); the Python-loop
                # figure is kept as us_hostloop
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(iters):
                        layer(xs[0], out=ys[0])
                g.replay()
                torch.cuda.synchronize(dev)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize(dev)
                us = e0.elapsed_time(e1) * 1e3 / iters
                del g
            flops = 2.0 * 196 * cin * cout * taps * n
            byts = float(np.prod(in_shape) + np.prod(out_shape)) * 4.0 * n + cin * cout * taps * 4.0
            row = dict(kind=kind, cin=cin, cout=cout, relu=relu, n=n, us_per_layer=us, us_hostloop=us_hostloop,
                       timing="cuda graph of back-to-back launches" if n == 1 else "python loop, rotating buffers",
                       images_per_s=n / (us * 1e-6), tflops_direct_equiv=flops / us * 1e-6,
                       hbm_gbs_algorithmic=byts / us * 1e-3,
                       frac_tf32_peak=flops / us * 1e-6 / tf32_peak,
                       frac_hbm_peak=byts / us * 1e-3 / pk["hbm"])
            rows.append(row)
            if verbose:
                print(f"{kind} {cin:>4}->{cout:<4} N={n:<3} {us:9.2f} us  {row['tflops_direct_equiv']:7.1f} TF/s "
                      f"({100 * row['frac_tf32_peak']:.1f}% tf32)  {row['hbm_gbs_algorithmic']:7.0f} GB/s "
                      f"({100 * row['frac_hbm_peak']:.1f}% hbm)", file=sys.stderr)
            del xs, ys
        layer.close()
    return rows


def run_all_shapes(args):
    """`--all-shapes`: the table above with 200 launches per cell; writes gpurun_out/all_shapes_latest.json."""
    import torch
    import wg_loader
    rows = measure_shapes(wg_loader.load(), torch.device("cuda", 0))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "all_shapes_latest.json"), "w") as f:
        json.dump(dict(peaks=peaks(), rows=rows), f, indent=1)
    print(json.dumps({"all_shapes": rows}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--all-shapes", action="store_true")
    args = ap.parse_args()
    # keep stdout to the one JSON line: native libraries (NCCL prints its version banner there) get fd 2 instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        return run_reference(args)
    if args.all_shapes:
        return run_all_shapes(args)
    run_gpu(args)


if __name__ == "__main__":
    main()
