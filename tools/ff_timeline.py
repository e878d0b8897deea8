#!/usr/bin/env python
"""Stage timeline of the full-fold 3x3 kernel (developer build): run with WG_FF_DEBUG=48, CTA 0 dumps clock64() stamps of
stages 8..15 of its first item into y. Prints clocks relative to the first stamp.

worker warp 0: 0 before wait raw_full | 1 raw there | 2 loads + column pass done (before wait v_empty[0]) | 3 V half 0 free
               | 4 half 0 stored + arrived | 5 before wait v_empty[1] | 6 V half 1 free | 7 half 1 stored + arrived
MMA thread   : 0 U chunk 0 there | 1 V half 0 there | 2 18 MMAs issued | 4 U chunk 1 there | 5 V half 1 there | 6 issued
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("WG_B200_DEV_LIB", "1")   # the ablation build lives in the developer library
os.environ.setdefault("WG_FF_DEBUG", "112")


def main():
    import numpy as np
    import torch
    import wg_loader
    wg = wg_loader.load()
    c_arg = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dts = (("tf32", wg.WG_TF32), ("bf16", wg.WG_BF16)) if len(sys.argv) <= 2 else (("tf32", wg.WG_TF32),)
    for name, dt in dts:
        c = k = c_arg
        n = 256
        rs = np.random.RandomState(0)
        w = (rs.rand(k, c, 3, 3) - 0.5).astype(np.float32)
        sc, sh = np.ones(k, np.float32), np.zeros(k, np.float32)
        layer = wg.Conv3x3BnRelu(w, sc, sh, relu=True, dtype=dt)
        x = torch.rand((n, 16, 16, c), device="cuda") - 0.5
        y = torch.zeros((n, 14, 14, k), device="cuda")
        y2 = torch.zeros((n, 14, 14, k), device="cuda")
        for _ in range(3):
            layer(x, out=y2)
            layer(x, out=y)
        torch.cuda.synchronize()
        ga = y2.view(-1)[8192:8192 + 148 * 32].cpu().numpy().view(np.int64).reshape(148, 16)[:, 12:16]
        gb = y.view(-1)[8192:8192 + 148 * 32].cpu().numpy().view(np.int64).reshape(148, 16)[:, 12:16]
        z = ga[:, 0].min()
        def pct(v):
            return [int(q) for q in np.percentile(v, [0, 10, 50, 90, 100])]
        print("   globaltimer ns since the previous launch's first CTA start, percentiles 0/10/50/90/100 over CTAs")
        print("   prev launch: CTA start", pct(ga[:, 0] - z), " CTA done", pct(ga[:, 3] - z))
        print("   last launch: CTA start", pct(gb[:, 0] - z), " past griddepcontrol.wait", pct(gb[:, 1] - z))
        print("                first raw stage landed", pct(gb[:, 2] - z), " CTA done", pct(gb[:, 3] - z))
        print("   launch period (first CTA start to first CTA start):", int(gb[:, 0].min() - z), "ns; last CTA done to next "
              "launch past the wait:", int(np.median(gb[:, 1]) - ga[:, 3].max()), "ns")
        raw = y.view(-1)[:512].cpu().numpy().view(np.int64)
        tw, tm = raw[:64].reshape(8, 8), raw[128:192].reshape(8, 8)
        t0 = tw[0, 0]
        print(f"== {name}: worker warp 0 (rows = stages 8..15; clocks since the first stamp)")
        for r in tw:
            print("  ", " ".join(f"{int(v - t0):7d}" for v in r))
        print(f"== {name}: MMA thread")
        for r in tm:
            print("  ", " ".join(f"{int(v - t0):7d}" if v else "      -" for v in r))
        per = (tw[7, 0] - tw[0, 0]) / 7.0
        print(f"   stage period {per:.0f} clk")
        it = y.view(-1)[8192:8192 + 148 * 32].cpu().numpy().view(np.int64).reshape(148, 4, 4)
        n_stage = c // (8 if name == "tf32" else 16)
        for idx in range(2):
            main = (it[:, idx, 1] - it[:, idx, 0]).astype(np.float64)
            epi = (it[:, idx, 2] - it[:, idx, 1]).astype(np.float64)
            ok = it[:, idx, 2] > 0
            n_sl = (k + 95) // 96
            sl = it[:, idx, 3] % n_sl
            for w in range(n_sl):
                m = ok & (sl == w)
                if m.any():
                    print(f"   item #{idx} slice {w}: {int(m.sum()):3d} CTAs  main loop {main[m].mean():8.0f} clk "
                          f"(min {main[m].min():.0f} max {main[m].max():.0f}; {main[m].mean() / n_stage:.0f}/stage)  "
                          f"epilogue {epi[m].mean():6.0f} (max {epi[m].max():.0f})")
        main0 = (it[:, 0, 1] - it[:, 0, 0])
        order = np.argsort(main0)
        print("   item #0 main loop, sorted (clk:block):", " ".join(f"{int(main0[b])}:{b}" for b in order[-24:]))
        print("   item #0 main loop percentiles 10/50/90:", [int(v) for v in np.percentile(main0, [10, 50, 90])])
        two = it[:, 1, 2] > 0
        tot2 = (it[two, 1, 2] - it[two, 0, 0])
        print(f"   CTAs with two items: total mean {tot2.mean():.0f} max {tot2.max():.0f} clk")
        tot = (it[:, 1, 2] - it[:, 0, 0]).astype(np.float64)
        gap = (it[:, 1, 0] - it[:, 0, 2]).astype(np.float64)
        print(f"   both items: mean {tot.mean():.0f} max {tot.max():.0f} clk; gap between items {gap.mean():.0f}")
        layer.close()


if __name__ == "__main__":
    main()
