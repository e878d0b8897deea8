#!/usr/bin/env python
"""Launch time of the fused 1x1 layers over the batch size, this repo vs cuDNN's fused conv+bias(+ReLU) (TF32, via torch;
reported baseline only). CUDA events over back-to-back launches, rotating buffer sets from N = 64 on; writes
gpurun_out/batch_sweep_1x1.json."""
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

    def timed(fn, iters=40):
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

    for (cin, cout, relu) in [(512, 128, True), (128, 512, False), (1024, 256, True), (256, 1024, False)]:
        rs = np.random.RandomState(cin)
        w = (rs.rand(cin, cout) - 0.5).astype(np.float32)
        sc, sh = rs.rand(cout).astype(np.float32), rs.rand(cout).astype(np.float32)
        layer = wg.Conv1x1Bn(w, sc, sh, relu=relu)
        wt = torch.from_numpy((w * sc[None, :]).T.copy().reshape(cout, cin, 1, 1)).cuda().contiguous(
            memory_format=torch.channels_last)
        bt = torch.from_numpy(sh).cuda()
        for n in (4, 8, 16, 32, 64, 96, 128, 192, 256):
            sets = 3 if n >= 64 else 1
            xs = [torch.rand((n, 196, cin), device="cuda") - 0.5 for _ in range(sets)]
            ys = [torch.empty((n, 196, cout), device="cuda") for _ in range(sets)]
            us = timed(lambda i: layer(xs[i % sets], out=ys[i % sets]))
            xcs = [x.view(n, 14, 14, cin).permute(0, 3, 1, 2) for x in xs]
            if relu:
                us_cudnn = timed(lambda i: torch.cudnn_convolution_relu(xcs[i % sets], wt, bt, (1, 1), (0, 0), (1, 1), 1))
            else:
                us_cudnn = timed(lambda i: torch.nn.functional.conv2d(xcs[i % sets], wt, bt))
            rows.append(dict(cin=cin, cout=cout, n=n, ours_us=round(us, 2), cudnn_tf32_us=round(us_cudnn, 2)))
            print(rows[-1], flush=True)
            del xs, ys, xcs
        layer.close()
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "batch_sweep_1x1.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
