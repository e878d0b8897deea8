#!/usr/bin/env python
"""Generates tests/golden/*.json|*.npz by IMPORTING THE REFERENCE'S OWN data_generator.py from /root/reference
(read-only; only available in the build container, never on the GPU box -- which is why the results are committed).

The reference checks in no golden vectors, no tests and no seed, and its only Python is the data
generator, so what can be pinned from reference-authored code is:
  (1) the byte-exact content of every data/*.bin its __main__ writes once numpy's global RNG is seeded
      (np.random.seed(0); the script draws with `from numpy.random import *` + rand(), data_generator.py:13,21) --
      committed as SHA-256 + size + leading values in ref_datagen_seed0.json; oracle/datagen.py must reproduce it;
  (2) its offline F(4x4,3x3) filter transform (weight_generator, :63-78) and BN folding (bias_generator, :40-47) on a
      tiny 8-channel case, committed in full in ref_tiny.npz;
  (3) outputs of the oracle on those reference-written files (regression vectors for the golden itself and for the
      emulation of the reference's three CUDA kernels), committed as leading slices + checksums in
      oracle_on_ref_seed0.json.

    python tests/golden/make_fixtures.py          # rewrites the fixtures; needs /root/reference
"""
import hashlib
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def import_reference_generator():
    # data_generator.py imports matplotlib / scipy.misc / requests at module scope and never uses them (:8-17)
    for name in ("matplotlib", "matplotlib.pyplot", "requests"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import scipy
    if not hasattr(scipy, "misc"):
        scipy.misc = types.ModuleType("scipy.misc")
        sys.modules["scipy.misc"] = scipy.misc
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_data_generator", os.path.join(REF, "data_generator.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def main():
    import golden
    ref = import_reference_generator()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        os.mkdir("data")
        # ---- (1) the reference's __main__ sequence (data_generator.py:116-127) under a seed
        np.random.seed(0)
        ref.bias_generator(output_channel=128)
        ref.input_generator(input_channel=128)
        ref.weight_generator(128, 128)
        ref.onebyone_generator()
        files = {}
        for name in sorted(os.listdir("data")):
            p = os.path.join("data", name)
            a = np.fromfile(p, "<f4")
            files[name] = dict(bytes=os.path.getsize(p), sha256=sha(p), head=[float(v) for v in a[:8]])
        with open(os.path.join(HERE, "ref_datagen_seed0.json"), "w") as f:
            json.dump(dict(seed=0, sequence="bias_generator(128); input_generator(128); weight_generator(128,128); "
                                            "onebyone_generator()", files=files), f, indent=1, sort_keys=True)

        # ---- (3) oracle outputs on the reference-written files
        ld = lambda n, shape: np.fromfile(os.path.join("data", n), "<f4").reshape(shape)
        x = ld("input_14_1_128.bin", (1, 16, 16, 128))
        w = ld("weight_NCHW_128_128.bin", (128, 128, 3, 3))
        u36 = ld("weight_winograd_128_128.bin", (36, 128, 128))
        sc, sh = ld("bnScale_winograd_128.bin", (128,)), ld("bnBias_winograd_128.bin", (128,))
        gold = golden.conv3x3_bn_relu(x, w, sc, sh, True)[0]
        emu = golden.reference_pipeline_f4x4(x[0], u36, sc, sh)
        out = dict(test0=dict(golden_head=[float(v) for v in gold.reshape(-1)[:32]],
                              golden_sum=float(gold.astype(np.float64).sum()),
                              golden_max=float(gold.max()),
                              ref_pipeline_max_abs_diff=float(np.abs(emu[1:15, 1:15] - gold).max()),
                              ref_pipeline_cnt_over_1e5=int((np.abs(emu[1:15, 1:15] - gold) > 1e-5).sum())))
        xo = np.fromfile("data/input_one_14_1024.bin", "<f4")
        wo = np.fromfile("data/weight_one_1024.bin", "<f4")
        so = np.fromfile("data/bnScale_myKernel_one_1024.bin", "<f4")
        bo = np.fromfile("data/bnBias_myKernel_one_1024.bin", "<f4")
        for mode, cin, cout, relu in ((2, 512, 128, True), (3, 128, 512, False), (4, 1024, 256, True),
                                      (5, 256, 1024, False)):
            g = golden.conv1x1_bn(xo[:196 * cin].reshape(196, cin), wo[:cin * cout].reshape(cin, cout), so[:cout],
                                  bo[:cout], relu)
            out[f"test{mode}"] = dict(golden_head=[float(v) for v in g.reshape(-1)[:32]],
                                      golden_sum=float(g.astype(np.float64).sum()), golden_max=float(g.max()))
        with open(os.path.join(HERE, "oracle_on_ref_seed0.json"), "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)

        # ---- (2) tiny case, committed in full
        for n in os.listdir("data"):
            os.remove(os.path.join("data", n))
        np.random.seed(7)
        ref.bias_generator(output_channel=8)
        ref.input_generator(input_channel=8)
        ref.weight_generator(8, 8)
        tiny = {n[:-4]: np.fromfile(os.path.join("data", n), "<f4") for n in sorted(os.listdir("data"))}
        np.savez_compressed(os.path.join(HERE, "ref_tiny.npz"), **tiny)
        os.chdir(cwd)
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
