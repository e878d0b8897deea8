#!/usr/bin/env python
"""Launch time of the fused 3x3 layer on the other ResNet stages (wg_conv3x3_create_hw: 56x56x64, 28x28x128, 14x14x256,
7x7x512) at equal pixel counts, this repo vs cuDNN's fused conv+bias+ReLU (TF32, via torch; reported baseline only).
CUDA events over back-to-back launches, rotating buffer sets; writes gpurun_out/mapsize_bench.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import wg_loader
    wg = wg_loader.load()
    torch.backends.cudnn.allow_tf32 = True
    rows = []

    def timed(fn, iters=30):
        for i in range(4):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters

    for (h, c, n) in [(56, 64, 16), (56, 64, 64), (28, 128, 64), (28, 128, 256), (14, 256, 256), (7, 512, 256),
                      (7, 512, 1024)]:
        rs = np.random.RandomState(c)
        w = (rs.rand(c, c, 3, 3) - 0.5).astype(np.float32)
        sc, sh = rs.rand(c).astype(np.float32), rs.rand(c).astype(np.float32)
        layer = wg.Conv3x3BnRelu(w, sc, sh, relu=True, hw=(h, h))
        hf, wf = wg.frame_dims(h, h)
        wt = torch.from_numpy(w * sc[:, None, None, None]).cuda().contiguous(memory_format=torch.channels_last)
        bt = torch.from_numpy(sh).cuda()
        sets = 3
        xs = [torch.zeros((n, hf, wf, c), device="cuda") for _ in range(sets)]
        for x in xs:
            x[:, 1:h + 1, 1:h + 1] = torch.rand((n, h, h, c), device="cuda") - 0.5
        ys = [torch.empty((n, h, h, c), device="cuda") for _ in range(sets)]
        us = timed(lambda i: layer(xs[i % sets], out=ys[i % sets]))
        us16 = None
        if c % 16 == 0 and c % 64 == 0:
            l16 = wg.Conv3x3BnRelu(w, sc, sh, relu=True, hw=(h, h), dtype=wg.WG_BF16)
            us16 = timed(lambda i: l16(xs[i % sets], out=ys[i % sets]))
            l16.close()
        xcs = [x[:, :h + 2, :h + 2].permute(0, 3, 1, 2) for x in xs]  # NCHW view of the NHWC frame (1-px border)
        us_cudnn = timed(lambda i: torch.cudnn_convolution_relu(xcs[i % sets], wt, bt, (1, 1), (0, 0), (1, 1), 1))
        flop = 2.0 * h * h * c * c * 9 * n
        rows.append(dict(h=h, c=c, n=n, ours_us=round(us, 2), ours_bf16_operands_us=us16 and round(us16, 2),
                         cudnn_tf32_us=round(us_cudnn, 2),
                         ours_tflops=round(flop / us / 1e6, 1)))
        print(rows[-1], flush=True)
        layer.close()
        del xs, ys, xcs
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "mapsize_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
