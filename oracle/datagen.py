#!/usr/bin/env python
"""Seeded restatement of the reference's data_generator.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

Writes the same raw little-endian float32 files, names, layouts and distributions as
/root/reference/data_generator.py:20-113, with three differences the reference cannot offer:
  * a seed (the reference draws from the unseeded global numpy RNG, data_generator.py:13,21);
  * any channel count / both 3x3 sets in one run (its __main__ only writes the 128 set, :116-127; Test 1 needs a
    hand-edited call, README.md:17);
  * golden_test{0..5}.bin: the FP64-accumulated golden of oracle/golden.py for each ./Test mode, dense [196][Cout].

Draw order inside each generator follows the reference so the files relate to one another the same way.

    python oracle/datagen.py --out data --seed 0            # everything ./Test 0..5 needs (+ goldens)
    python oracle/datagen.py --out data --no-f4x4           # skip weight_winograd_*.bin (reference-kernel input only)
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden  # noqa: E402


def _write(path, arr):
    np.ascontiguousarray(arr, dtype='<f4').tofile(path)


def bias_generator(rs, out, output_channel=128):
    """data_generator.py:20-47 -- conv bias (dead), gamma, beta, running mean U(-.5,.5), running var U(5,8), folded."""
    k = output_channel
    bias = (rs.rand(k) - 0.5).astype(np.float32)
    bn_scale = (rs.rand(k) - 0.5).astype(np.float32)
    bn_bias = (rs.rand(k) - 0.5).astype(np.float32)
    e_mean = (rs.rand(k) - 0.5).astype(np.float32)
    e_var = (rs.rand(k) * 3 + 5).astype(np.float32)
    sc, sh = golden.fold_bn(bn_scale, bn_bias, e_mean, e_var)
    for name, a in (("bias", bias), ("bnScale", bn_scale), ("bnBias", bn_bias), ("eMean", e_mean), ("eVar", e_var),
                    ("bnScale_winograd", sc), ("bnBias_winograd", sh)):
        _write(os.path.join(out, f"{name}_{k}.bin"), a)
    return dict(gamma=bn_scale, beta=bn_bias, mean=e_mean, var=e_var, scale=sc, shift=sh)


def input_generator(rs, out, input_channel=128, feature_map_size=14, padding=1, batch=1):
    """data_generator.py:49-53 -- (14+2)^2 * C values U(-.5,.5), HWC, the border is random too. batch>1 appends
    further images; image 0 is what the reference-named file holds."""
    side = feature_map_size + 2 * padding
    x = (rs.rand(batch * side * side * input_channel) - 0.5).astype(np.float32)
    x = x.reshape(batch, side, side, input_channel)
    _write(os.path.join(out, f"input_{feature_map_size}_{padding}_{input_channel}.bin"), x[0])
    return x


def weight_generator(rs, out, input_channel=128, output_channel=128, f4x4=True):
    """data_generator.py:55-78 -- KCRS weights U(-.5,.5) and (reference-only) their F(4x4,3x3) transform
    [36][Cin][Cout]."""
    c, k = input_channel, output_channel
    w = (rs.rand(k * c * 9) - 0.5).astype(np.float32).reshape(k, c, 3, 3)
    _write(os.path.join(out, f"weight_NCHW_{c}_{k}.bin"), w)
    if f4x4:
        _write(os.path.join(out, f"weight_winograd_{c}_{k}.bin"), golden.filter_transform(w, golden.G_4))
    return w


def onebyone_generator(rs, out, input_channel=256, output_channel=1024, feature_map_size=14):
    """data_generator.py:80-113 -- the shared 1x1 set: everything U(-20,20), var U(5,25); each ./Test 2..5 reads a
    prefix of these files."""
    p = feature_map_size * feature_map_size
    x = ((rs.rand(p * output_channel) - 0.5) * 40).astype(np.float32)
    w = ((rs.rand(input_channel * output_channel) - 0.5) * 40).astype(np.float32)
    bn_scale = ((rs.rand(output_channel) - 0.5) * 40).astype(np.float32)
    bn_bias = ((rs.rand(output_channel) - 0.5) * 40).astype(np.float32)
    e_mean = ((rs.rand(output_channel) - 0.5) * 40).astype(np.float32)
    e_var = (rs.rand(output_channel) * 20 + 5).astype(np.float32)
    sc, sh = golden.fold_bn(bn_scale, bn_bias, e_mean, e_var)
    _write(os.path.join(out, f"input_one_{feature_map_size}_{output_channel}.bin"), x)
    _write(os.path.join(out, f"weight_one_{output_channel}.bin"), w)
    for name, a in (("bnScale_one", bn_scale), ("bnBias_one", bn_bias), ("eMean_one", e_mean), ("eVar_one", e_var),
                    ("bnScale_myKernel_one", sc), ("bnBias_myKernel_one", sh)):
        _write(os.path.join(out, f"{name}_{output_channel}.bin"), a)
    return dict(x=x, w=w, gamma=bn_scale, beta=bn_bias, mean=e_mean, var=e_var, scale=sc, shift=sh)


# (mode, Cin, Cout, relu) of ./Test 2..5 (Test.c:31-42; Kernel128_one.cu:58-59,277-278; Kernel256_one.cu:60-61,278-279)
ONE_CASES = ((2, 512, 128, True), (3, 128, 512, False), (4, 1024, 256, True), (5, 256, 1024, False))


def one_case_tensors(one, cin, cout):
    """The prefixes each 1x1 entry point reads from the shared files."""
    return (one["x"][:196 * cin].reshape(196, cin), one["w"][:cin * cout].reshape(cin, cout),
            one["scale"][:cout], one["shift"][:cout])


def generate_all(out, seed=0, f4x4=True, goldens=True):
    os.makedirs(out, exist_ok=True)
    rs = np.random.RandomState(seed)
    result = {}
    for mode, ch in ((0, 128), (1, 256)):
        bn = bias_generator(rs, out, ch)
        x = input_generator(rs, out, ch)
        w = weight_generator(rs, out, ch, ch, f4x4=f4x4)
        result[mode] = dict(x=x, w=w, **bn)
        if goldens:
            g = golden.conv3x3_bn_relu(x, w, bn["scale"], bn["shift"], relu=True)[0]
            _write(os.path.join(out, f"golden_test{mode}.bin"), g)
            result[mode]["golden"] = g
    one = onebyone_generator(rs, out)
    for mode, cin, cout, relu in ONE_CASES:
        x, w, sc, sh = one_case_tensors(one, cin, cout)
        result[mode] = dict(x=x, w=w, scale=sc, shift=sh)
        if goldens:
            g = golden.conv1x1_bn(x, w, sc, sh, relu)
            _write(os.path.join(out, f"golden_test{mode}.bin"), g)
            result[mode]["golden"] = g
    return result


def generate_reference_main(out, seed=0):
    """Exactly the call sequence of the reference's __main__ (data_generator.py:116-127): the 128 set, then the 1x1
    set. With np.random.seed(seed) the reference script writes byte-identical files (tests/golden/ref_datagen_seed0.json)."""
    os.makedirs(out, exist_ok=True)
    rs = np.random.RandomState(seed)
    bias_generator(rs, out, 128)
    input_generator(rs, out, 128)
    weight_generator(rs, out, 128, 128, f4x4=True)
    onebyone_generator(rs, out)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--out", default="data")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-f4x4", action="store_true")
    ap.add_argument("--no-golden", action="store_true")
    a = ap.parse_args()
    generate_all(a.out, a.seed, f4x4=not a.no_f4x4, goldens=not a.no_golden)
    print(f"wrote {len(os.listdir(a.out))} files to {a.out}/ (seed {a.seed})")


if __name__ == "__main__":
    main()
