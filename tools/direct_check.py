#!/usr/bin/env python
"""Developer check of the direct-convolution 3x3 kernel (make dev): parity against the oracle and launch time.

    python tools/direct_check.py [--ns 4,32,256] [--bo 0,1]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ns", default="3,32,256")
    ap.add_argument("--bo", default="1,2,4", help="mode: cluster size 1 / 2 / 4, + 8 = no mixed schedule")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--shapes", default="256x256,128x128")
    ap.add_argument("--op16", default="0", help="operand types: 0 = TF32, 1 = bf16, 2 = fp16")
    args = ap.parse_args()
    import numpy as np
    import torch
    import golden
    L = ctypes.CDLL(os.path.join(ROOT, "tools", "libwinograd_b200_dev.so"))
    vp = ctypes.c_void_p
    L.wg_dev_direct_pack.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.wg_dev_direct_run.argtypes = [vp] * 5 + [ctypes.c_int] * 7 + [vp]
    dev = torch.device("cuda", 0)
    bad = 0
    for shape in args.shapes.split(","):
        c, k = [int(v) for v in shape.split("x")]
        rs = np.random.RandomState(c + k)
        w = (rs.rand(k, c, 3, 3) - 0.5).astype(np.float32)
        sc, sh = golden.fold_bn(rs.rand(k) - 0.5, rs.rand(k) - 0.5, rs.rand(k) - 0.5, rs.rand(k) * 3 + 5)
        wd = torch.from_numpy(w).to(dev)
        scd = torch.from_numpy(np.asarray(sc, np.float32)).to(dev)
        shd = torch.from_numpy(np.asarray(sh, np.float32)).to(dev)
        for op16 in [int(v) for v in args.op16.split(",")]:
            if op16 and c % 64:
                continue
            img = torch.zeros(9 * c * k, device=dev)
            assert L.wg_dev_direct_pack(wd.data_ptr(), img.data_ptr(), c, k, op16) == 0
            for n in [int(v) for v in args.ns.split(",")]:
                x = np.zeros((n, 16, 16, c), np.float32)
                x[:, 1:15, 1:15] = rs.rand(n, 14, 14, c) - 0.5
                sample = sorted(set([0, n // 2, n - 1]))
                gold = golden.conv3x3_bn_relu(x[sample], w, sc, sh, True)
                sets = 4 if n >= 64 else 1
                xs = [torch.from_numpy(x).to(dev) for _ in range(sets)]
                for bo in [int(v) | (op16 << 8) for v in args.bo.split(",")]:
                    if op16 and (bo & 7) != 1:
                        continue
                    for padded in (0, 1):
                        ys = [torch.full((n, 16, 16, k) if padded else (n, 14, 14, k), float("nan"), device=dev)
                              for _ in range(sets)]

                        def run(i):
                            rc = L.wg_dev_direct_run(xs[i % sets].data_ptr(), img.data_ptr(), scd.data_ptr(),
                                                     shd.data_ptr(), ys[i % sets].data_ptr(), n, c, k, 1, padded,
                                                     148, bo, None)
                            assert rc == 0, rc
                        run(0)
                        torch.cuda.synchronize()
                        got = ys[0][sample].cpu().numpy()
                        if padded:
                            border = bool(np.all(got[:, 0] == 0) and np.all(got[:, 15] == 0) and
                                          np.all(got[:, :, 0] == 0) and np.all(got[:, :, 15] == 0))
                            got = got[:, 1:15, 1:15]
                        else:
                            border = True
                        err = float(golden.rel_err(got, gold)) if np.isfinite(got).all() else float("nan")
                        for i in range(3):
                            run(i)
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for i in range(args.iters):
                            run(i)
                        e1.record()
                        torch.cuda.synchronize()
                        us = e0.elapsed_time(e1) * 1e3 / args.iters
                        tol = 1e-2 if op16 == 1 else 1e-3
                        ok = bool(err <= tol) and border
                        bad += 0 if ok else 1
                        print(f"direct {c}->{k} N={n} mode={bo} padded={padded}: rel_err {err:.2e} "
                              f"border_zero {border}  {us:8.2f} us{'' if ok else '  FAILED'}", flush=True)
    print(f"{bad} failures")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
