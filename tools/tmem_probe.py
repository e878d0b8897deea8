#!/usr/bin/env python
"""Developer probe (make dev): clocks per tcgen05.ld.32x32b.x32 (4 KB per warp) for 1..8 warps of one CTA, 1..4 loads
in flight per tcgen05.wait::ld, with and without the shared-memory write + fence.proxy.async the 1x1 epilogue pays per
chunk. Prints bytes per clock per SM.   WG_B200_DEV_LIB=1 python tools/tmem_probe.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = ctypes.CDLL(os.path.join(ROOT, "tools", "libwinograd_b200_dev.so"))
L.wg_dev_tmem_ld_probe.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_longlong)]
iters = 2000
for fence in (0, 1):
    for warps in (1, 2, 4, 8):
        for depth in (1, 2, 4):
            c = ctypes.c_longlong(0)
            rc = L.wg_dev_tmem_ld_probe(warps, iters, depth, fence, ctypes.byref(c))
            assert rc == 0, rc
            per = c.value / iters
            print(f"fence={fence} warps={warps} loads/wait={depth}: {per:7.1f} clk per iteration, "
                  f"{per / depth:6.1f} clk per x32 load, {warps * depth * 4096 / per:6.1f} B/clk")
