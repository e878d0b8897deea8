#!/usr/bin/env python
"""Quick A/B timing table (developer aid): 3x3 128/128 and 256/256 at N = 256, 128, 64, 32 (TF32, optionally bf16) and the
four 1x1 shapes at N = 256; CUDA events, rotating buffers > L2. One line per case.  python tools/quick.py [--bf16] [--one]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--one", action="store_true", help="also the 1x1 shapes")
    ap.add_argument("--ns", default="256,128,64,32")
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import numpy as np
    import torch
    import wg_loader
    wg = wg_loader.load()
    dev = torch.device("cuda", 0)

    def timeit(fn, sets):
        for i in range(4):
            fn(i % sets)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.iters):
                fn(i % sets)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / args.iters)
        return best

    out = []
    for c in (256, 128):
        rs = np.random.RandomState(c)
        w = (rs.rand(c, c, 3, 3) - 0.5).astype(np.float32)
        sc, sh = (rs.rand(c) + 0.5).astype(np.float32), (rs.rand(c) - 0.5).astype(np.float32)
        for name, dt in (("tf32", wg.WG_TF32),) + ((("bf16", wg.WG_BF16),) if args.bf16 else ()):
            layer = wg.Conv3x3BnRelu(w, sc, sh, relu=True, dtype=dt)
            for n in [int(v) for v in args.ns.split(",")]:
                sets = max(2, min(8, int(300e6 // (n * (256 * c + 196 * c) * 4)) + 1))
                xs = [torch.rand((n, 16, 16, c), device=dev) - 0.5 for _ in range(sets)]
                ys = [torch.empty((n, 14, 14, c), device=dev) for _ in range(sets)]
                us = timeit(lambda i: layer(xs[i], out=ys[i]), sets)
                out.append(f"3x3 {c}->{c} {name} N={n:<3} {us:7.2f} us")
                del xs, ys
            layer.close()
    if args.one:
        for cin, cout, relu in ((512, 128, True), (128, 512, False), (1024, 256, True), (256, 1024, False)):
            rs = np.random.RandomState(cin)
            layer = wg.Conv1x1Bn((rs.rand(cin, cout) - 0.5).astype(np.float32), rs.rand(cout).astype(np.float32),
                                 rs.rand(cout).astype(np.float32), relu)
            n, sets = 256, 3
            xs = [torch.rand((n, 196, cin), device=dev) - 0.5 for _ in range(sets)]
            ys = [torch.empty((n, 196, cout), device=dev) for _ in range(sets)]
            us = timeit(lambda i: layer(xs[i], out=ys[i]), sets)
            gbs = (n * 196 * (cin + cout) + cin * cout) * 4 / us * 1e-3
            out.append(f"1x1 {cin}->{cout} N=256 {us:7.2f} us  {gbs:6.0f} GB/s")
            rr = torch.rand((n, 196, cout), device=dev)
            us = timeit(lambda i: layer(xs[i], out=ys[i], residual=rr, relu_after_add=True), sets)
            out.append(f"1x1 {cin}->{cout} N=256 +residual {us:7.2f} us")
            del xs, ys, rr
            layer.close()
    print(f"== quick {args.tag}")
    print("\n".join(out))


if __name__ == "__main__":
    main()
