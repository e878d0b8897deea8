#!/usr/bin/env python
"""Developer check of the 3x3 throughput kernels on a B200: parity against the oracle on sampled images and launch
time at N=256 (rotating buffer sets, CUDA events) for every (kernel variant, operand type, shape).

    python tools/ff_check.py [--iters 40] [--out gpurun_out/ff_check.json]

kn = 96: full-fold kernel (wino_ff_kernel.cu), kn = 48: half-fold V-in-TMEM kernel (wino_tm_kernel.cu).
The oracle (oracle/golden.py) is used as the checker only.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ff_check.json"))
    ap.add_argument("--kns", default="96")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--time-only", action="store_true", help="skip the parity section, never fail (ablation runs)")
    ap.add_argument("--dtypes", default="tf32,bf16,fp16")
    ap.add_argument("--sets", type=int, default=4, help="rotating input/output sets (1 = L2-warm input)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import golden
    import wg_loader
    wg = wg_loader.load()
    dev = torch.device("cuda", 0)
    rows = []

    def set_kn(kn):
        """kn != 96 (superseded kernel generations) exists in the developer build only: WG_B200_DEV_LIB=1."""
        if wg.IS_DEV_LIB:
            wg.lib().wg_dev_set_wino_kn(kn)
        else:
            assert kn == 96, "kn != 96 needs the developer build: make dev && WG_B200_DEV_LIB=1 python tools/ff_check.py"

    def rand(rs, n, c, k):
        x = (rs.rand(n, 16, 16, c) - 0.5).astype(np.float32)
        w = (rs.rand(k, c, 3, 3) - 0.5).astype(np.float32)
        sc, sh = golden.fold_bn(rs.rand(k) - 0.5, rs.rand(k) - 0.5, rs.rand(k) - 0.5, rs.rand(k) * 3 + 5)
        return x, w, sc, sh

    dts = {"tf32": wg.WG_TF32, "bf16": wg.WG_BF16, "fp16": wg.WG_FP16}
    dts = {k: v for k, v in dts.items() if k in args.dtypes.split(",")}
    # ---- parity on awkward shapes (every slice width, ragged last M-block, padded frame)
    shapes = [(40, 64, 256), (33, 32, 96), (50, 24, 160), (20, 16, 512), (64, 128, 128), (131, 48, 192), (37, 8, 32),
              (29, 40, 64), (300, 16, 32), (24, 256, 128), (37, 64, 160)]
    if args.quick:
        shapes = shapes[:3] + shapes[-2:]
    if args.time_only:
        shapes = []
    for kn in [int(v) for v in args.kns.split(",")]:
        set_kn(kn)
        for (n, c, k) in shapes:
            rs = np.random.RandomState(n + c + k)
            x, w, sc, sh = rand(rs, n, c, k)
            gold = golden.conv3x3_bn_relu(x, w, sc, sh, True)
            for name, dt in dts.items():
                if dt != wg.WG_TF32 and (c % 16 or k % 64):
                    continue
                layer = wg.Conv3x3BnRelu(w, sc, sh, relu=True, dtype=dt)
                xd = torch.from_numpy(x).cuda()
                y = layer(xd).cpu().numpy()
                yp = layer(xd, out_padded=True).cpu().numpy()
                err = float(golden.rel_err(y, gold))
                border0 = bool(np.all(yp[:, 0] == 0) and np.all(yp[:, 15] == 0) and np.all(yp[:, :, 0] == 0)
                               and np.all(yp[:, :, 15] == 0))
                same = bool(np.array_equal(yp[:, 1:15, 1:15], y))
                tol = 1e-2 if name == "bf16" else 1e-3
                ok = err <= tol and border0 and same
                rows.append(dict(check="parity", kn=kn, dtype=name, n=n, c=c, k=k, rel_err=err, border_zero=border0,
                                 padded_equals_dense=same, ok=ok))
                print(rows[-1], flush=True)
                layer.close()

    # ---- timing + sampled parity at N=256
    n = 256
    for (c, k) in [(256, 256), (128, 128)]:
        rs = np.random.RandomState(7)
        w = (rs.rand(k, c, 3, 3) - 0.5).astype(np.float32)
        sc, sh = golden.fold_bn(rs.rand(k) - 0.5, rs.rand(k) - 0.5, rs.rand(k) - 0.5, rs.rand(k) * 3 + 5)
        sets = args.sets
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        xs = [torch.rand((n, 16, 16, c), device=dev, generator=g) - 0.5 for _ in range(sets)]
        ys = [torch.empty((n, 14, 14, k), device=dev) for _ in range(sets)]
        sample = [0, 1, 127, 255]
        x_s = xs[0][sample].cpu().numpy()
        gold = golden.conv3x3_bn_relu(x_s, w, sc, sh, True)
        for kn in [int(v) for v in args.kns.split(",")]:
            set_kn(kn)
            for name, dt in dts.items():
                layer = wg.Conv3x3BnRelu(w, sc, sh, relu=True, dtype=dt)
                for i in range(5):
                    layer(xs[i % sets], out=ys[i % sets])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(args.iters):
                    layer(xs[i % sets], out=ys[i % sets])
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / args.iters
                layer(xs[0], out=ys[0])
                err = float(golden.rel_err(ys[0][sample].cpu().numpy(), gold))
                tol = 1e-2 if name == "bf16" else 1e-3
                rows.append(dict(check="time", kn=kn, dtype=name, n=n, c=c, k=k, us=us, rel_err=err, ok=err <= tol,
                                 tflops_direct=2 * 196 * c * k * 9 * n / us / 1e6))
                print(rows[-1], flush=True)
                layer.close()
        del xs, ys
        torch.cuda.empty_cache()
    set_kn(96)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)
    bad = [r for r in rows if not r["ok"]]
    print(f"{len(rows)} rows, {len(bad)} failures")
    sys.exit(1 if bad and not args.time_only else 0)


if __name__ == "__main__":
    main()
