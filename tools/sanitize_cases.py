#!/usr/bin/env python
"""Every kernel family of the product library once, at sizes small enough to run under compute-sanitizer
(tools/sanitize.sh: memcheck, racecheck, synccheck). compute-sanitizer is CLOSED on this GPU pool (its wrapper prints
"compute-sanitizer is closed on this pool and stays closed ... Find a bad access with bounds checks and asserts of your
own, small cases, and a comparison with the CPU reference"), so this script carries its own checks:
  * parity: every result against the oracle (whole tensor);
  * out-of-bounds writes: every output lives inside a larger allocation whose guard zones (64 KB on either side) hold a
    sentinel that must survive the launch bit for bit;
  * races / missing hand-offs: every launch is repeated and must be bit-identical run to run (fixed accumulation
    order, no atomics: a lost or early mbarrier hand-off shows up as a differing or stale element), with the output
    pre-filled with NaN so that an element nobody wrote is caught."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import numpy as np
    import torch
    import golden
    import wg_loader
    wg = wg_loader.load()
    rs = np.random.RandomState(0)
    fails = 0

    def r3(n, c, k, hw=(14, 14)):
        hf, wf = golden.frame_dims(*hw)
        x = (rs.rand(n, hf, wf, c) - 0.5).astype(np.float32)
        w = (rs.rand(k, c, 3, 3) - 0.5).astype(np.float32)
        sc, sh = golden.fold_bn(rs.rand(k) - 0.5, rs.rand(k) - 0.5, rs.rand(k) - 0.5, rs.rand(k) * 3 + 5)
        return x, w, sc, sh

    GUARD = 16384  # floats on either side of every output

    def run_guarded(fn, shape, reps=4):
        """fn(out) launches into `out`; returns the result after checking guard zones, full coverage and determinism."""
        nonlocal fails
        numel = int(np.prod(shape))
        buf = torch.full((numel + 2 * GUARD,), -123456.0, device="cuda")
        out = buf[GUARD:GUARD + numel].view(shape)
        first = None
        for _ in range(reps):
            out.fill_(float("nan"))
            fn(out)
            torch.cuda.synchronize()
            if first is None:
                first = out.clone()
            elif not torch.equal(out, first):
                fails += 1
                print("FAIL run-to-run difference", flush=True)
        if not (bool((buf[:GUARD] == -123456.0).all()) and bool((buf[GUARD + numel:] == -123456.0).all())):
            fails += 1
            print("FAIL guard zone overwritten", flush=True)
        if not bool(torch.isfinite(first).all()):
            fails += 1
            print("FAIL output element never written", flush=True)
        return first.cpu().numpy()

    def check(name, got, gold, tol):
        nonlocal fails
        e = golden.rel_err(got, gold)
        ok = e <= tol
        fails += not ok
        print(f"{'ok  ' if ok else 'FAIL'} {name:60s} rel_err {e:.2e}", flush=True)

    # 3x3: latency kernel (clusters, DSMEM reduction), throughput kernel (96 / 64 / 32 wide slices, ragged M-block),
    # 16-warp sibling (split-C clusters, narrow image), 16-bit operands, padded frame, other map sizes
    for (n, c, k, dt, tol, hw, padded, tag) in [
            (1, 128, 128, wg.WG_TF32, 1e-3, (14, 14), True, "wino3x3_small_kernel (cluster split-C)"),
            (2, 64, 64, wg.WG_TF32, 1e-3, (14, 14), False, "wino3x3_small_kernel"),
            (40, 24, 160, wg.WG_TF32, 1e-3, (14, 14), True, "wino3x3_ff_kernel 96+64 slices (C % 32 != 0)"),
            (131, 16, 96, wg.WG_TF32, 1e-3, (14, 14), False, "wino3x3_ff_kernel ragged, several items per CTA"),
            (12, 256, 192, wg.WG_BF16, 1e-2, (14, 14), False, "wino3x3_ffw_kernel bf16, split-C (K % 128 != 0)"),
            (33, 256, 64, wg.WG_FP16, 1e-3, (14, 14), False, "wino3x3_ffw_kernel fp16 operands (C >= 256)"),
            (40, 32, 64, wg.WG_BF16, 1e-2, (14, 14), True, "wino3x3_ff_kernel bf16 operands"),
            (2, 64, 64, wg.WG_BF16, 1e-2, (14, 14), False, "wino3x3_bn_relu_kernel bf16 small batch (split-C)"),
            (9, 24, 64, wg.WG_TF32, 1e-3, (28, 28), True, "wino3x3_ff_kernel 28x28 (C % 32 != 0)"),
            (21, 32, 96, wg.WG_TF32, 1e-3, (7, 7), True, "wino3x3_ff_kernel 7x7 (masked edge tiles)"),
            (12, 256, 256, wg.WG_TF32, 1e-3, (14, 14), False, "conv3x3_direct_kernel half-image items"),
            (75, 128, 128, wg.WG_TF32, 1e-3, (14, 14), True, "conv3x3_direct_kernel mixed schedule, frame"),
            (150, 32, 192, wg.WG_TF32, 1e-3, (14, 14), True, "conv3x3_direct_kernel K % 128 == 64"),
            (40, 64, 128, wg.WG_BF16, 1e-2, (14, 14), True, "conv3x3_direct16_kernel bf16"),
            (9, 128, 256, wg.WG_FP16, 1e-3, (14, 14), False, "conv3x3_direct16_kernel fp16"),
            (9, 32, 64, wg.WG_TF32, 1e-3, (28, 28), True, "conv3x3_direct_gen_kernel 28x28 row bands"),
            (21, 64, 128, wg.WG_TF32, 1e-3, (7, 7), True, "conv3x3_direct_gen_kernel 7x7, two images per item"),
            (5, 64, 128, wg.WG_BF16, 1e-2, (28, 28), True, "conv3x3_direct16_gen_kernel 28x28")]:
        x, w, sc, sh = r3(n, c, k, hw)
        layer = wg.Conv3x3BnRelu(w, sc, sh, relu=True, dtype=dt, hw=hw)
        xd = torch.from_numpy(x).cuda()
        y = run_guarded(lambda out: layer(xd, out=out, out_padded=padded), (n,) + layer.out_shape(padded))
        gold = golden.conv3x3_bn_relu(x, w, sc, sh, True, hw=hw)
        check(tag + f" N={n} {c}->{k}", y[:, 1:hw[0] + 1, 1:hw[1] + 1] if padded else y, gold, tol)
        layer.close()

    # 1x1: latency kernel (split-K clusters), throughput kernel (plain, weight-stationary), residual, bf16, padded frame
    for (n, cin, cout, dt, tol, res, padded, tag) in [
            (1, 512, 128, wg.WG_TF32, 1e-3, False, False, "conv1x1_small_kernel split-K"),
            (1, 256, 1024, wg.WG_TF32, 1e-3, True, False, "conv1x1_small_kernel + residual"),
            (37, 96, 256, wg.WG_TF32, 1e-3, False, True, "conv1x1_bn_act_kernel<256> padded frame, ragged"),
            (40, 128, 512, wg.WG_TF32, 1e-3, False, False, "conv1x1_bn_act_kernel weight-stationary"),
            (40, 128, 512, wg.WG_TF32, 1e-3, True, False, "conv1x1_bn_act_kernel weight-stationary + residual"),
            (40, 512, 128, wg.WG_TF32, 1e-3, True, False, "conv1x1_bn_act_kernel<128> + residual"),
            (37, 64, 384, wg.WG_BF16, 1e-2, False, True, "conv1x1_bn_act_kernel bf16 operands, padded"),
            (37, 64, 384, wg.WG_BF16, 1e-2, True, False, "conv1x1_bn_act_kernel bf16 operands + residual"),
            (150, 128, 512, wg.WG_TF32, 1e-3, False, False, "conv1x1_t_kernel<128> pairs, ragged pixel tile"),
            (150, 128, 512, wg.WG_TF32, 1e-3, True, False, "conv1x1_t_kernel<128> pairs + residual"),
            (75, 256, 1024, wg.WG_TF32, 1e-3, False, False, "conv1x1_t_kernel<256> pairs, resident slab"),
            (75, 256, 1024, wg.WG_TF32, 1e-3, True, False, "conv1x1_t_kernel<256> pairs + residual (streamed weights)"),
            (199, 96, 384, wg.WG_TF32, 1e-3, True, False, "conv1x1_t_kernel<128> no pairs + residual"),
            (200, 512, 128, wg.WG_TF32, 1e-3, False, True, "conv1x1_tf_kernel frame output, no pairs, ragged item"),
            (90, 1024, 256, wg.WG_TF32, 1e-3, False, True, "conv1x1_tf_kernel frame output, pairs")]:
        x = ((rs.rand(n, 196, cin) - 0.5) * 4).astype(np.float32)
        w = (rs.rand(cin, cout) - 0.5).astype(np.float32)
        sc, sh = (rs.rand(cout) + 0.5).astype(np.float32), (rs.rand(cout) - 0.5).astype(np.float32)
        r = (rs.rand(n, 196, cout) - 0.5).astype(np.float32)
        layer = wg.Conv1x1Bn(w, sc, sh, relu=not res, dtype=dt)
        xd = torch.from_numpy(x).cuda()
        if res:
            rd = torch.from_numpy(r).cuda()
            y = run_guarded(lambda out: layer(xd, out=out, residual=rd, relu_after_add=True), (n, 196, cout))
            gold = golden.conv1x1_bn_residual(x.reshape(-1, cin), w, sc, sh, False, r, True).reshape(n, 196, cout)
        else:
            y = run_guarded(lambda out: layer(xd, out=out, out_padded=padded), (n,) + layer.out_shape(padded))
            gold = golden.conv1x1_bn(x, w, sc, sh, True)
            if padded:
                y = y[:, 1:15, 1:15].reshape(n, 196, cout)
        check(tag + f" N={n} {cin}->{cout}", y, gold, tol)
        layer.close()

    # host-buffer pipeline (chunks, three streams) and the packed blob
    x, w, sc, sh = r3(100, 32, 64)
    layer = wg.Conv3x3BnRelu(w, sc, sh)
    check("wg_run_host N=100 (chunks 50,25,25)", layer.run_host(x), golden.conv3x3_bn_relu(x, w, sc, sh), 1e-3)
    again = wg._Layer.deserialize(layer.serialize())
    check("wg_layer_deserialize", again(torch.from_numpy(x).cuda()).cpu().numpy(), golden.conv3x3_bn_relu(x, w, sc, sh), 1e-3)
    torch.cuda.synchronize()
    print(f"{fails} failures")
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
