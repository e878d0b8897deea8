#!/usr/bin/env python
"""Reported baseline only (never on the product path): cuDNN's fused conv + bias + ReLU on the same B200, through
torch (cudnn_convolution_relu -> cudnnConvolutionBiasActivationForward), BN scale folded into the weights and shift as
bias, channels_last, TF32 allowed and not allowed, bf16 too; N=1 latency (L2-warm back-to-back launches captured in a CUDA graph so
that neither side is bound by Python dispatch; the Python-loop figures are kept as *_hostloop_us) and N=256 throughput (rotating buffers > L2). Also times this repo's kernels in the same process for a side-by-side table.
BASELINE.md section 2 asks for exactly this. Writes gpurun_out/cudnn_baseline.json.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wg_loader  # noqa: E402

SHAPES = [("3x3", 128, 128, True), ("3x3", 256, 256, True), ("1x1", 512, 128, True), ("1x1", 128, 512, False),
          ("1x1", 1024, 256, True), ("1x1", 256, 1024, False)]


def timeit(fn, sets, iters):
    for i in range(5):
        fn(i % sets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % sets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def timeit_graph(fn, iters):
    """N=1: a layer is shorter than one Python call, so time `iters` back-to-back calls captured once into a CUDA graph
    (both for this repo's kernels and for cuDNN): the GPU-side latency a C caller's loop would see."""
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn(0)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    wg = wg_loader.load()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    rows = []
    print(f"torch {torch.__version__}, cuDNN {torch.backends.cudnn.version()}", file=sys.stderr)
    for kind, cin, cout, relu in SHAPES:
        rs = np.random.RandomState(0)
        ks = 3 if kind == "3x3" else 1
        w_kcrs = (rs.rand(cout, cin, ks, ks) - 0.5).astype(np.float32)
        scale = (rs.rand(cout) + 0.5).astype(np.float32)
        shift = (rs.rand(cout) - 0.5).astype(np.float32)
        if kind == "3x3":
            ours = wg.Conv3x3BnRelu(w_kcrs, scale, shift, relu)
        else:
            ours = wg.Conv1x1Bn(np.ascontiguousarray(w_kcrs[:, :, 0, 0].T), scale, shift, relu)
        w_fold = torch.from_numpy(w_kcrs * scale[:, None, None, None]).to(dev).contiguous(memory_format=torch.channels_last)
        bias = torch.from_numpy(shift).to(dev)
        for n in (1, 256):
            sets = 1 if n == 1 else 4
            hw = 16 if kind == "3x3" else 14
            x_nhwc = [torch.rand((n, hw, hw, cin), device=dev) - 0.5 for _ in range(sets)]
            x_cl = [t.permute(0, 3, 1, 2) for t in x_nhwc]  # NCHW view of NHWC storage = channels_last
            x_ours = x_nhwc if kind == "3x3" else [t.view(n, 196, cin) for t in x_nhwc]
            y_ours = [torch.empty((n,) + ours.out_shape(), device=dev) for _ in range(sets)]
            iters = 300 if n == 1 else 100
            row = dict(kind=kind, cin=cin, cout=cout, relu=relu, n=n)
            row["ours_tf32_us"] = timeit(lambda i: ours(x_ours[i], out=y_ours[i]), sets, iters)
            if n == 1:
                row["ours_tf32_hostloop_us"] = row["ours_tf32_us"]
                row["ours_tf32_us"] = timeit_graph(lambda i: ours(x_ours[i], out=y_ours[i]), 100)
            for name, dt, tf32 in (("cudnn_fp32_us", torch.float32, False), ("cudnn_tf32_us", torch.float32, True),
                                   ("cudnn_bf16_us", torch.bfloat16, True)):
                torch.backends.cudnn.allow_tf32 = tf32
                xs = [t.to(dt) for t in x_cl] if dt != torch.float32 else x_cl
                wt, bt = w_fold.to(dt), bias.to(dt)
                if relu:
                    f = lambda i: torch.cudnn_convolution_relu(xs[i], wt, bt, (1, 1), (0, 0), (1, 1), 1)
                else:
                    f = lambda i: torch.nn.functional.conv2d(xs[i], wt, bt)
                try:
                    row[name] = timeit(f, sets, iters)
                    if n == 1:
                        row[name.replace("_us", "_hostloop_us")] = row[name]
                        row[name] = timeit_graph(f, 100)
                except Exception as e:  # noqa: BLE001
                    row[name] = None
                    print("cudnn failed", name, kind, cin, cout, n, e, file=sys.stderr)
            # sanity: cuDNN fp32 result vs ours
            torch.backends.cudnn.allow_tf32 = False
            ref = (torch.cudnn_convolution_relu(x_cl[0], w_fold, bias, (1, 1), (0, 0), (1, 1), 1) if relu
                   else torch.nn.functional.conv2d(x_cl[0], w_fold, bias)).permute(0, 2, 3, 1).reshape(y_ours[0].shape)
            got = ours(x_ours[0])
            row["ours_vs_cudnn_fp32_rel"] = float((got - ref).abs().max() / ref.abs().max())
            rows.append(row)
            print(f"{kind} {cin:>4}->{cout:<4} N={n:<3} ours {row['ours_tf32_us']:8.2f} us | cuDNN fp32 "
                  f"{row['cudnn_fp32_us'] or -1:8.2f}  tf32 {row['cudnn_tf32_us'] or -1:8.2f}  bf16 "
                  f"{row['cudnn_bf16_us'] or -1:8.2f} us | rel diff {row['ours_vs_cudnn_fp32_rel']:.1e}", file=sys.stderr)
            del x_nhwc, x_cl, y_ours
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "cudnn_baseline.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, cudnn=torch.backends.cudnn.version(), rows=rows), f, indent=1)


if __name__ == "__main__":
    main()
