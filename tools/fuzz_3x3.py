#!/usr/bin/env python
"""Randomised parity check of the 3x3 path on a B200: random (N, C, K), operand type and output frame against the oracle
(checker only), plus run-to-run bit-identity. Seeds are printed; a failure exits 1.
    python tools/fuzz_3x3.py [--cases 60] [--seed 1]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    import numpy as np
    import torch
    import golden
    import wg_loader
    wg = wg_loader.load()
    rs = np.random.RandomState(args.seed)
    bad = 0
    for case in range(args.cases):
        c = int(rs.choice([8, 16, 24, 32, 48, 64, 96, 128, 256]))
        k = int(rs.choice([32, 64, 96, 128, 160, 192, 256, 320]))
        n = int(rs.choice([1, 2, 3, 5, 9, 12, 13, 17, 24, 31, 40, 53, 64, 77]))
        if c * k * n > 256 * 256 * 40:
            n = max(1, (256 * 256 * 40) // (c * k))
        x = (rs.rand(n, 16, 16, c) - 0.5).astype(np.float32)
        w = (rs.rand(k, c, 3, 3) - 0.5).astype(np.float32)
        sc, sh = golden.fold_bn(rs.rand(k) - 0.5, rs.rand(k) - 0.5, rs.rand(k) - 0.5, rs.rand(k) * 3 + 5)
        relu = bool(rs.randint(2))
        gold = golden.conv3x3_bn_relu(x, w, sc, sh, relu)
        xd = torch.from_numpy(x).cuda()
        for name, dt, tol in (("tf32", wg.WG_TF32, 1e-3), ("bf16", wg.WG_BF16, 1e-2), ("fp16", wg.WG_FP16, 1e-3)):
            if dt != wg.WG_TF32 and (c % 16 or k % 64):
                continue
            layer = wg.Conv3x3BnRelu(w, sc, sh, relu=relu, dtype=dt)
            y = layer(xd)
            yp = layer(xd, out_padded=True)
            again = layer(xd)
            err = float(golden.rel_err(y.cpu().numpy(), gold))
            ok = (err <= tol and torch.equal(yp[:, 1:15, 1:15], y) and torch.equal(again, y)
                  and float(yp[:, 0].abs().max()) == 0 and float(yp[:, :, 15].abs().max()) == 0)
            if not ok:
                bad += 1
                print(f"FAIL case {case}: n={n} c={c} k={k} relu={relu} {name} rel_err={err:.3e}", flush=True)
            layer.close()
        if case % 10 == 9:
            print(f"{case + 1} cases, {bad} failures", flush=True)
    print(f"done: {args.cases} cases, {bad} failures (seed {args.seed})")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
