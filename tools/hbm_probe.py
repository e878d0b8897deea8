#!/usr/bin/env python
"""Read-only, write-only and copy HBM bandwidth of this GPU (torch kernels, 1 GiB buffers, CUDA events, best of 10):
the 1x1 layers are HBM streams of very different read / write mixes, and write-only traffic does not reach the copy
figure that MEASURED_PEAKS.json holds."""
import json
import torch


def best(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = min(t, e0.elapsed_time(e1))
    return t * 1e-3


def main():
    n = 1 << 28  # 1 GiB of fp32
    a = torch.empty(n, device="cuda")
    b = torch.empty(n, device="cuda")
    a.normal_()
    out = {
        "write_only_gbs (fill_)": n * 4 / best(lambda: b.fill_(1.0)) / 1e9,
        "read_only_gbs (sum)": n * 4 / best(lambda: a.sum()) / 1e9,
        "copy_gbs (read+write, copy_)": 2 * n * 4 / best(lambda: b.copy_(a)) / 1e9,
        "read3_write1_gbs (add of 3 -> 1)": None,
    }
    c = torch.empty(n // 4, device="cuda")
    a4 = a.view(4, n // 4)
    out["read4_write1_gbs (sum over dim 0 of [4, n/4])"] = (n * 4 + n) / best(lambda: torch.sum(a4, 0, out=c)) / 1e9
    del out["read3_write1_gbs (add of 3 -> 1)"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
