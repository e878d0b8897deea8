#!/usr/bin/env python
"""Item timeline of the 1x1 throughput kernel (developer build, WG_ONE_ABLATE includes 16): CTA 0's MMA thread and
epilogue warp 2 stamp clock64() per item. usage: WG_B200_DEV_LIB=1 WG_ONE_ABLATE=<16|20|31> python tools/one_timeline.py cin cout"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import wg_loader
    wg = wg_loader.load()
    cin, cout = int(sys.argv[1]), int(sys.argv[2])
    rs = np.random.RandomState(0)
    layer = wg.Conv1x1Bn((rs.rand(cin, cout) - 0.5).astype(np.float32), np.ones(cout, np.float32),
                         np.zeros(cout, np.float32), False)
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    x = torch.rand((n, 196, cin), device="cuda") - 0.5
    y = torch.empty((n, 196, cout), device="cuda")
    dbg = torch.zeros(8192, dtype=torch.int64, device="cuda")
    wg.lib().wg_dev_set_dbg_ptr(ctypes.c_void_p(dbg.data_ptr()))
    for _ in range(3):
        layer(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        layer(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    print(f"  (this single-buffer loop: {e0.elapsed_time(e1) * 50:.1f} us per launch)")
    g = dbg.cpu().numpy()[1024:1024 + 4 * 148].reshape(148, 4)
    z = g[:, 1].min()
    import numpy as _np
    for k in sorted(set(g[:, 3].tolist())):
        m = g[:, 3] == k
        print(f"  CTAs with {k} items ({int(m.sum())}): entry {int((g[m, 0] - z).min())}..{int((g[m, 0] - z).max())} ns, past the wait "
              f"{int((g[m, 1] - z).min())}..{int((g[m, 1] - z).max())}, exit {int((g[m, 2] - z).min())} .. median {int(_np.median(g[m, 2] - z))} .. {int((g[m, 2] - z).max())}")
    c = dbg.cpu().numpy()[2048:2048 + 64].reshape(8, 8)
    if c[0, 0]:
        print("  epilogue warp 2, third item, per 32-cout chunk: start | store-read wait done | tcgen05.ld done | BN + staging done | "
              "fence + warp sync done | TMA store issued (clk since the chunk start; last column: start of the next chunk)")
        for i in range(8):
            if c[i, 0] == 0:
                break
            nxt = int(c[i + 1, 0] - c[i, 0]) if i + 1 < 8 and c[i + 1, 0] else -1
            print("   chunk", i, " ".join(f"{int(v - c[i, 0]):6d}" for v in c[i, 1:6]), f"{nxt:7d}")
    t = dbg.cpu().numpy()[:128].reshape(16, 8)
    t0 = t[0, 0]
    print(f"== N={n} {cin}->{cout} WG_ONE_ABLATE={os.environ.get('WG_ONE_ABLATE')}: per item: MMA start | acc free | last commit || epi start | epi end")
    print(f"  producer past griddepcontrol.wait (previous launch complete) at {int(t[0, 3] - t0)}")
    for i in range(16):
        if t[i, 0] == 0:
            break
        r = [int(v - t0) if v else -1 for v in t[i]]
        print(f"  item {i:2d}: {r[0]:7d} {r[1]:7d} {r[2]:7d}   || {r[4]:7d} {r[5]:7d}   (mainloop {r[2]-r[1]:6d}, epilogue {r[5]-r[4]:6d})")


if __name__ == "__main__":
    main()
