#!/usr/bin/env python
"""BASELINE.json configs[4] -- the ResNet-50 bottleneck chain 1x1 -> 3x3 -> 1x1 (+BN, ReLU on the first two), N=256
per GPU, this repo's three fused launches (cuda_winograd_b200.Bottleneck) next to cuDNN's fused conv+bias+ReLU chain on
the same B200 (through torch; reported baseline only, never on the product path). CUDA events, rotating input sets.
Writes gpurun_out/chain_bench.json.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import golden  # noqa: E402  (checker only)
import wg_loader  # noqa: E402


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


def main():
    wg = wg_loader.load()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    rows = []
    for cin, c in ((512, 128), (1024, 256)):
        cout = cin
        rs = np.random.RandomState(0)
        w1 = ((rs.rand(cin, c) - 0.5) * 0.2).astype(np.float32)
        w3 = ((rs.rand(c, c, 3, 3) - 0.5) * 0.2).astype(np.float32)
        w2 = ((rs.rand(c, cout) - 0.5) * 0.2).astype(np.float32)
        bn = [((rs.rand(k) + 0.5).astype(np.float32), (rs.rand(k) - 0.3).astype(np.float32)) for k in (c, c, cout)]
        block = wg.Bottleneck(w1, *bn[0], w3, *bn[1], w2, *bn[2])
        block_bf16 = wg.Bottleneck(w1, *bn[0], w3, *bn[1], w2, *bn[2], dtype=wg.WG_BF16)
        # cuDNN: BN scale folded into the weights, shift as bias, channels_last
        cl = torch.channels_last
        k1 = torch.from_numpy((w1.T * bn[0][0][:, None])[:, :, None, None].copy()).to(dev).contiguous(memory_format=cl)
        k3 = torch.from_numpy(w3 * bn[1][0][:, None, None, None]).to(dev).contiguous(memory_format=cl)
        k2 = torch.from_numpy((w2.T * bn[2][0][:, None])[:, :, None, None].copy()).to(dev).contiguous(memory_format=cl)
        b1, b3, b2 = (torch.from_numpy(b[1]).to(dev) for b in bn)
        for n in (1, 256):
            sets = 1 if n == 1 else 3
            xs = [torch.rand((n, 14, 14, cin), device=dev) - 0.5 for _ in range(sets)]
            x_ours = [t.view(n, 196, cin) for t in xs]
            x_cl = [t.permute(0, 3, 1, 2) for t in xs]
            outs = [torch.empty((n, 196, cout), device=dev) for _ in range(sets)]
            iters = 200 if n == 1 else 50
            row = dict(cin=cin, c=c, cout=cout, n=n)
            row["ours_tf32_us"] = timeit(lambda i: block(x_ours[i], out=outs[i]), sets, iters)
            row["ours_bf16_3x3_us"] = timeit(lambda i: block_bf16(x_ours[i], out=outs[i]), sets, iters)
            replay, _ = block.capture(x_ours[0], out=outs[0])       # one CUDA-graph launch for the three kernels
            row["ours_tf32_graph_us"] = timeit(lambda i: replay(), 1, iters)

            def cudnn_chain(i):
                a = torch.cudnn_convolution_relu(x_cl[i], k1, b1, (1, 1), (0, 0), (1, 1), 1)
                a = torch.cudnn_convolution_relu(a, k3, b3, (1, 1), (1, 1), (1, 1), 1)
                return torch.nn.functional.conv2d(a, k2, b2)

            for name, tf32 in (("cudnn_tf32_us", True), ("cudnn_fp32_us", False)):
                torch.backends.cudnn.allow_tf32 = tf32
                row[name] = timeit(cudnn_chain, sets, iters)
                if n == 1:  # cuDNN's chain in a CUDA graph too, so that neither side pays Python dispatch
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        cudnn_chain(0)
                    row[name.replace("_us", "_graph_us")] = timeit(lambda i: g.replay(), 1, iters)
            torch.backends.cudnn.allow_tf32 = False
            ref = cudnn_chain(0).permute(0, 2, 3, 1).reshape(n, 196, cout)
            got = block(x_ours[0])
            row["ours_vs_cudnn_fp32_rel"] = float((got - ref).abs().max() / ref.abs().max())
            if n == 1:
                gold = golden.bottleneck_chain(xs[0].view(n, 196, cin).cpu().numpy(), w1, *bn[0], w3, *bn[1], w2, *bn[2])
                row["ours_vs_oracle_rel"] = golden.rel_err(got.cpu().numpy(), gold)
            flops = 2.0 * n * 196 * (cin * c + 9 * c * c + c * cout)
            row["ours_tflops_direct_equiv"] = flops / row["ours_tf32_us"] * 1e-6
            rows.append(row)
            print(f"chain {cin}->{c}->{c}->{cout} N={n:<3} ours {row['ours_tf32_us']:8.1f} us (graph {row['ours_tf32_graph_us']:.1f}; bf16 3x3 "
                  f"{row['ours_bf16_3x3_us']:8.1f}) | cuDNN tf32 {row['cudnn_tf32_us']:8.1f} (graph "
                  f"{row.get('cudnn_tf32_graph_us', float('nan')):.1f}) fp32 {row['cudnn_fp32_us']:8.1f} us"
                  f" | rel diff vs cuDNN fp32 {row['ours_vs_cudnn_fp32_rel']:.1e}", file=sys.stderr)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "chain_bench.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, cudnn=torch.backends.cudnn.version(), rows=rows), f, indent=1)


if __name__ == "__main__":
    main()
