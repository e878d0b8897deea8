"""CPU tests (-m "not gpu"): the oracle against reference-authored fixtures, against brute force, and its C port."""
import ctypes
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import datagen
import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


# ---------------------------------------------------------------------------------------------- reference-authored pins
def test_datagen_reproduces_reference_script_bytes(tmp_path, golden_dir):
    """oracle/datagen.py == the reference's data_generator.py run under np.random.seed(0), file by file."""
    fix = json.load(open(os.path.join(golden_dir, "ref_datagen_seed0.json")))
    out = tmp_path / "data"
    datagen.generate_reference_main(str(out), seed=0)
    assert len(fix["files"]) == 18  # 10 for the 3x3 set + 8 for the 1x1 set
    for name, meta in fix["files"].items():
        p = out / name
        assert p.exists(), name
        assert os.path.getsize(p) == meta["bytes"], name
        if name.startswith("weight_winograd_"):
            # float64 G g G^T rounded to float32: allow the last bit (different summation order)
            a = np.fromfile(p, "<f4")
            np.testing.assert_allclose(a[:8], np.array(meta["head"], np.float32), rtol=2e-7, atol=1e-9)
        else:
            assert _sha(p) == meta["sha256"], name


def test_filter_transform_and_bn_fold_match_reference_tiny(golden_dir):
    """weight_generator's F(4x4) transform (data_generator.py:63-78) and bias_generator's folding (:40-47), 8 channels."""
    t = np.load(os.path.join(golden_dir, "ref_tiny.npz"))
    w = t["weight_NCHW_8_8"].reshape(8, 8, 3, 3)
    u = golden.filter_transform(w, golden.G_4).astype(np.float32)
    np.testing.assert_allclose(u.reshape(-1), t["weight_winograd_8_8"], rtol=3e-7, atol=1e-9)
    sc, sh = golden.fold_bn(t["bnScale_8"], t["bnBias_8"], t["eMean_8"], t["eVar_8"])
    np.testing.assert_array_equal(sc, t["bnScale_winograd_8"])
    np.testing.assert_array_equal(sh, t["bnBias_winograd_8"])


def test_oracle_regression_on_reference_written_files(tmp_path, golden_dir):
    fix = json.load(open(os.path.join(golden_dir, "oracle_on_ref_seed0.json")))
    out = tmp_path / "data"
    datagen.generate_reference_main(str(out), seed=0)
    ld = lambda n: np.fromfile(out / n, "<f4")
    x = ld("input_14_1_128.bin").reshape(1, 16, 16, 128)
    w = ld("weight_NCHW_128_128.bin").reshape(128, 128, 3, 3)
    sc, sh = ld("bnScale_winograd_128.bin"), ld("bnBias_winograd_128.bin")
    g = golden.conv3x3_bn_relu(x, w, sc, sh)[0]
    np.testing.assert_allclose(g.reshape(-1)[:32], fix["test0"]["golden_head"], rtol=1e-6, atol=1e-7)
    assert abs(float(g.astype(np.float64).sum()) - fix["test0"]["golden_sum"]) < 1e-3
    xo, wo = ld("input_one_14_1024.bin"), ld("weight_one_1024.bin")
    so, bo = ld("bnScale_myKernel_one_1024.bin"), ld("bnBias_myKernel_one_1024.bin")
    for mode, cin, cout, relu in datagen.ONE_CASES:
        y = golden.conv1x1_bn(xo[:196 * cin].reshape(196, cin), wo[:cin * cout].reshape(cin, cout), so[:cout],
                              bo[:cout], relu)
        np.testing.assert_allclose(y.reshape(-1)[:32], fix[f"test{mode}"]["golden_head"], rtol=1e-6)
        assert (y.min() >= 0) == relu


def test_emulated_reference_kernels_agree_with_golden(seeded_data):
    """Index-for-index emulation of kernel_128_winograd_BtdB / OuterProduct / AtIA (Kernel128_winograd.cu:28-213) on the
    F(4x4) file vs the direct golden: the report's acceptance rule (max abs ~1e-5, <0.1 % over 1e-5; report.pdf sec. 5)."""
    out_dir, t = seeded_data
    u36 = np.fromfile(os.path.join(out_dir, "weight_winograd_128_128.bin"), "<f4").reshape(36, 128, 128)
    d = t[0]
    frame = golden.reference_pipeline_f4x4(d["x"][0], u36, d["scale"], d["shift"])
    assert frame[0].max() == 0 and frame[15].max() == 0 and frame[:, 0].max() == 0 and frame[:, 15].max() == 0
    max_err, cnt = golden.output_checker(frame, d["golden"], 14, 128, 1)
    assert max_err < 5e-5
    assert cnt < 0.001 * 196 * 128


# ------------------------------------------------------------------------------------------------- internal consistency
def test_golden_vs_bruteforce_loops():
    rs = np.random.RandomState(1)
    x = (rs.rand(1, 16, 16, 4) - 0.5).astype(np.float32)
    w = (rs.rand(3, 4, 3, 3) - 0.5).astype(np.float32)
    sc = (rs.rand(3) - 0.5).astype(np.float32)
    sh = (rs.rand(3) - 0.5).astype(np.float32)
    for relu in (True, False):
        a = golden.conv3x3_bn_relu(x, w, sc, sh, relu)
        b = golden.conv3x3_bn_relu_loops(x, w, sc, sh, relu)
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-7)


def test_winograd_f2x2_restatement_and_tf32_prediction(seeded_data):
    _, t = seeded_data
    d = t[0]
    y = golden.winograd_f2x2(d["x"], d["w"], d["scale"], d["shift"])[0]
    assert golden.rel_err(y, d["golden"]) < 1e-6
    y_tf32 = golden.winograd_f2x2(d["x"], d["w"], d["scale"], d["shift"], operand_dtype="tf32")[0]
    e = golden.rel_err(y_tf32, d["golden"])
    assert 1e-5 < e < 1e-3  # the tolerance north_star states for TF32, with room (measured ~5e-4)


def test_unfolded_bn_equals_folded(seeded_data):
    _, t = seeded_data
    d = t[0]
    y = golden.conv3x3_bn_relu_unfolded(d["x"], d["w"], d["gamma"], d["beta"], d["mean"], d["var"])[0]
    assert golden.rel_err(y, d["golden"]) < 1e-6


def test_golden_files_and_prefix_semantics(seeded_data):
    out_dir, t = seeded_data
    for mode, cin, cout, relu in datagen.ONE_CASES:
        g = np.fromfile(os.path.join(out_dir, f"golden_test{mode}.bin"), "<f4").reshape(196, cout)
        np.testing.assert_array_equal(g, t[mode]["golden"])
        assert t[mode]["x"].shape == (196, cin) and t[mode]["w"].shape == (cin, cout)
    # all four cases read prefixes of the same two files (Kernel128_one.h:8-9)
    assert np.shares_memory(t[2]["x"], t[5]["x"]) or np.array_equal(t[2]["x"].reshape(-1)[:196 * 128],
                                                                    t[3]["x"].reshape(-1))


def test_output_checker_shift_semantics():
    a = np.zeros((16, 16, 3), np.float32)
    b = np.ones((14, 14, 3), np.float32)
    a[1:15, 1:15] = b
    a[2, 3, 1] += 1e-3
    mx, cnt = golden.output_checker(a, b, 14, 3, 1)
    assert cnt == 1 and abs(mx - 1e-3) < 1e-6


def test_round_tf32():
    v = np.array([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -10, -3.14159], np.float32)
    r = golden.round_tf32(v)
    assert r[0] == 1.0 and r[2] == 1.0 + 2 ** -10
    assert r[1] == 1.0 + 2 ** -10  # tie rounds away from zero (cvt.rna)
    assert abs(r[3] + 3.14159) < 2 ** -10


# ---------------------------------------------------------------------------------------------------------- C port
@pytest.fixture(scope="module")
def liboracle():
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    L = ctypes.CDLL(path)
    fp = ctypes.POINTER(ctypes.c_float)
    L.oracle_conv3x3_bn_relu.argtypes = [fp, fp, fp, fp, fp] + [ctypes.c_int] * 5
    L.oracle_conv1x1_bn.argtypes = [fp, fp, fp, fp, fp, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.oracle_fold_bn.argtypes = [ctypes.c_int, fp, fp, fp, fp, ctypes.c_float, fp, fp]
    L.oracle_output_checker.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
    return L


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def test_c_port_matches_numpy_golden(liboracle, seeded_data):
    _, t = seeded_data
    d = t[0]
    x = np.ascontiguousarray(d["x"])
    for padded in (0, 1):
        y = np.full((1, 16, 16, 128) if padded else (1, 14, 14, 128), 7.0, np.float32)
        liboracle.oracle_conv3x3_bn_relu(_p(x), _p(np.ascontiguousarray(d["w"])), _p(d["scale"]), _p(d["shift"]), _p(y),
                                         1, 128, 128, 1, padded)
        ref = golden.pad_frame(d["golden"][None]) if padded else d["golden"][None]
        np.testing.assert_allclose(y, ref, rtol=1e-6, atol=1e-7)
    for mode, cin, cout, relu in datagen.ONE_CASES[:2]:
        d = t[mode]
        y = np.empty((196, cout), np.float32)
        liboracle.oracle_conv1x1_bn(_p(np.ascontiguousarray(d["x"])), _p(np.ascontiguousarray(d["w"])),
                                    _p(np.ascontiguousarray(d["scale"])), _p(np.ascontiguousarray(d["shift"])), _p(y),
                                    196, cin, cout, int(relu))
        np.testing.assert_allclose(y, d["golden"], rtol=1e-6, atol=1e-3)


def test_c_port_fold_bn_and_checker(liboracle, golden_dir):
    t = np.load(os.path.join(golden_dir, "ref_tiny.npz"))
    sc, sh = np.empty(8, np.float32), np.empty(8, np.float32)
    liboracle.oracle_fold_bn(8, _p(t["bnScale_8"]), _p(t["bnBias_8"]), _p(t["eMean_8"]), _p(t["eVar_8"]),
                             ctypes.c_float(1e-5), _p(sc), _p(sh))
    np.testing.assert_allclose(sc, t["bnScale_winograd_8"], rtol=2e-7)
    np.testing.assert_allclose(sh, t["bnBias_winograd_8"], rtol=2e-7, atol=1e-8)
    a = np.zeros((16, 16, 2), np.float32)
    b = np.zeros((14, 14, 2), np.float32)
    a[5, 5, 0] = 0.5
    mx = ctypes.c_float()
    cnt = liboracle.oracle_output_checker(_p(a), _p(b), 14, 2, 1, ctypes.byref(mx))
    assert cnt == 1 and mx.value == 0.5
