#!/usr/bin/env python
"""Launch time of the fused 3x3 layer over the batch size (the per-GPU batch of a 256-image job on 1..8 GPUs and the
range between the latency and the throughput kernels), this repo vs cuDNN's fused conv+bias+ReLU (TF32, via torch; reported
baseline only). CUDA events over back-to-back launches; writes gpurun_out/batch_sweep.json."""
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
    torch.backends.cuda.matmul.allow_tf32 = True
    rows = []

    def timed(fn, iters=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters

    for c in (256, 128):
        rs = np.random.RandomState(c)
        w = (rs.rand(c, c, 3, 3) - 0.5).astype(np.float32)
        sc, sh = rs.rand(c).astype(np.float32), rs.rand(c).astype(np.float32)
        layer = wg.Conv3x3BnRelu(w, sc, sh, relu=True)
        wt = torch.from_numpy(w * sc[:, None, None, None]).cuda().contiguous(memory_format=torch.channels_last)
        bt = torch.from_numpy(sh).cuda()
        for n in (2, 4, 8, 12, 16, 24, 32, 48, 64, 96, 128, 192, 256):
            x = torch.rand((n, 16, 16, c), device="cuda") - 0.5
            y = torch.empty((n, 14, 14, c), device="cuda")
            us = timed(lambda: layer(x, out=y))
            xc = x.permute(0, 3, 1, 2)  # NCHW view of NHWC storage = channels_last
            us_cudnn = timed(lambda: torch.cudnn_convolution_relu(xc, wt, bt, (1, 1), (0, 0), (1, 1), 1))
            rows.append(dict(c=c, n=n, ours_us=us, cudnn_tf32_us=us_cudnn,
                             ours_tflops_direct=2 * 196 * c * c * 9 * n / us / 1e6))
            print(rows[-1], flush=True)
        layer.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "batch_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
