#!/usr/bin/env python
"""Run one layer shape a few times (for ncu captures) and print its CUDA-event time.
    python tools/one_shape.py 1x1 256 1024 [--n 256] [--iters 20] [--relu 0]
    python tools/one_shape.py 3x3 128 128
Also prints the time of a plain device fill and copy of the same output / total bytes (HBM write and copy ceilings)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kind")
    ap.add_argument("cin", type=int)
    ap.add_argument("cout", type=int)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--relu", type=int, default=1)
    ap.add_argument("--sets", type=int, default=3)
    args = ap.parse_args()
    import numpy as np
    import torch
    import wg_loader
    wg = wg_loader.load()
    rs = np.random.RandomState(0)
    n, cin, cout = args.n, args.cin, args.cout
    sc, sh = rs.rand(cout).astype(np.float32), rs.rand(cout).astype(np.float32)
    if args.kind == "1x1":
        layer = wg.Conv1x1Bn((rs.rand(cin, cout).astype(np.float32) - 0.5), sc, sh, relu=bool(args.relu))
        xs = [torch.rand((n, 196, cin), device="cuda") - 0.5 for _ in range(args.sets)]
        ys = [torch.empty((n, 196, cout), device="cuda") for _ in range(args.sets)]
    else:
        layer = wg.Conv3x3BnRelu((rs.rand(cout, cin, 3, 3).astype(np.float32) - 0.5), sc, sh, relu=bool(args.relu))
        xs = [torch.rand((n, 16, 16, cin), device="cuda") - 0.5 for _ in range(args.sets)]
        ys = [torch.empty((n, 14, 14, cout), device="cuda") for _ in range(args.sets)]

    def timed(fn, iters):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters

    us = timed(lambda i: layer(xs[i % args.sets], out=ys[i % args.sets]), args.iters)
    in_b, out_b = xs[0].numel() * 4, ys[0].numel() * 4
    print(f"{args.kind} {cin}->{cout} N={n}: {us:.2f} us  ({(in_b + out_b) / us / 1e3:.0f} GB/s algorithmic)")
    us_fill = timed(lambda i: ys[i % args.sets].fill_(1.0), args.iters)
    print(f"fill of the output ({out_b / 1e6:.0f} MB): {us_fill:.2f} us = {out_b / us_fill / 1e3:.0f} GB/s")
    big = [torch.empty((in_b + out_b) // 8, device="cuda") for _ in range(2 * args.sets)]
    us_copy = timed(lambda i: big[2 * (i % args.sets)].copy_(big[2 * (i % args.sets) + 1]), args.iters)
    print(f"copy moving the same total bytes ({(in_b + out_b) / 1e6:.0f} MB read+write): {us_copy:.2f} us = "
          f"{(in_b + out_b) / us_copy / 1e3:.0f} GB/s")


if __name__ == "__main__":
    main()
