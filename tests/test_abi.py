"""CPU tests (-m "not gpu"): the C-ABI library builds, loads and exports what include/*.h declares; host-side helpers;
loud failure without a GPU; batch sharding + output gather over gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built(wg):
    wg.build()
    return wg


def _declared_symbols():
    names = set()
    for h in ("winograd_b200.h", "wg_legacy.h", "util.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"^[\w\s\*]+?\b(\w+)\s*\([^;{]*\)\s*;", src, flags=re.M):
            if "typedef" not in m.group(0):
                names.add(m.group(1))
    return names


def test_library_exports_every_declared_symbol(built):
    L = built.lib()
    declared = _declared_symbols()
    assert {"wg_run", "wg_conv3x3_create", "kernel_128", "kernel_256_1_out", "get_parameter"} <= declared
    assert declared == set(built.ABI_SYMBOLS), declared ^ set(built.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"


def test_compat_headers_have_the_reference_names():
    for h in ("Kernel128_winograd.h", "Kernel256_winograd.h", "Kernel128_one.h", "Kernel256_one.h", "util.h"):
        assert os.path.exists(os.path.join(ROOT, "include", h))
    assert os.path.exists(os.path.join(ROOT, "Test")), "make builds the ./Test harness"


def test_reference_test_c_compiles_and_links_unchanged(built, tmp_path):
    """Drop-in at the source level: the reference's own Test.c, untouched, against include/ + libwinograd_b200.so.
    Only where /root/reference exists (the build container); the GPU box runs the prebuilt oracle/_ref/Test_dropin."""
    ref = "/root/reference/Test.c"
    if not os.path.exists(ref):
        pytest.skip("reference sources not present on this machine")
    exe = tmp_path / "Test_ref"
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-w", "-I", os.path.join(ROOT, "include"),
                        "-o", str(exe), ref, "-L", os.path.dirname(built.LIB_PATH), "-lwinograd_b200",
                        "-Xlinker", "-rpath", "-Xlinker", os.path.dirname(built.LIB_PATH)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    run = subprocess.run([str(exe), "0"], cwd=tmp_path, capture_output=True, text=True)   # no data/ here
    assert "---- Iter: 0 ----" in run.stdout and "Bad file path" in run.stdout and run.returncode == 0


def test_no_cudnn_no_cublas_in_product(built):
    out = subprocess.run(["ldd", built.LIB_PATH], capture_output=True, text=True).stdout
    assert "cudnn" not in out and "cublas" not in out, out


def test_sass_is_blackwell_native(built):
    sass = subprocess.run(["cuobjdump", "-sass", built.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCQMMA" in sass or re.search(r"UTC\w*MMA", sass)   # tcgen05.mma
    assert "UTMALDG" in sass and "UBLKCP" in sass                                     # TMA tensor + bulk copies
    assert "LDTM" in sass                                                             # tcgen05.ld
    assert "HMMA" not in sass.replace("UTCHMMA", "")                                   # no legacy mma.sync path


def test_fold_bn_matches_reference_formula(built, golden_dir):
    t = np.load(os.path.join(golden_dir, "ref_tiny.npz"))
    sc, sh = built.fold_bn(t["bnScale_8"], t["bnBias_8"], t["eMean_8"], t["eVar_8"])
    np.testing.assert_allclose(sc, t["bnScale_winograd_8"], rtol=2e-7)
    np.testing.assert_allclose(sh, t["bnBias_winograd_8"], rtol=2e-7, atol=1e-8)


def test_host_utils_match_reference_semantics(built, tmp_path):
    L = built.lib()
    fp = ctypes.POINTER(ctypes.c_float)
    L.get_parameter.restype = fp
    L.get_parameter.argtypes = [ctypes.c_char_p, ctypes.c_int]
    L.transpose.restype = fp
    L.transpose.argtypes = [fp, ctypes.c_int, ctypes.c_int]
    L.output_checker.restype = ctypes.c_float
    L.output_checker.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.getTimeMicroseconds64.restype = ctypes.c_uint64
    a = np.arange(12, dtype=np.float32)
    p = tmp_path / "a.bin"
    a.tofile(p)
    buf = L.get_parameter(str(p).encode(), 12)
    assert [buf[i] for i in range(12)] == list(a)
    # transpose(weight, h, w): weight is [w][h] -> [h][w] (util.c:15-26)
    t = L.transpose(buf, 3, 4)
    got = np.array([t[i] for i in range(12)], np.float32).reshape(3, 4)
    np.testing.assert_array_equal(got, a.reshape(4, 3).T)
    t0, t1 = L.getTimeMicroseconds64(), L.getTimeMicroseconds64()
    assert 0 <= t1 - t0 < 10_000_000
    A = np.zeros((16, 16, 2), np.float32)
    B = np.zeros((14, 14, 2), np.float32)
    A[3, 4, 1] = 0.25
    assert L.output_checker(A.ctypes.data_as(fp), B.ctypes.data_as(fp), 14, 2, 1) == 0.25


def test_missing_file_exits_like_the_reference(built, tmp_path):
    """util.c:36-39: a missing data file prints 'Bad file path' and exit(0)s the process."""
    code = ("import sys; sys.path.insert(0, %r); import wg_loader; m = wg_loader.load(); m.kernel_128(); print('survived')"
            % ROOT)
    r = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, capture_output=True, text=True)
    assert "Bad file path" in r.stdout and "survived" not in r.stdout and r.returncode == 0


def test_fails_loudly_without_a_gpu(built):
    if built.device_count() > 0:
        pytest.skip("a B200 is visible")
    w = np.zeros((32, 32, 3, 3), np.float32)
    s = np.ones(32, np.float32)
    with pytest.raises(built.WinogradB200Error, match="no sm_100"):
        built.Conv3x3BnRelu(w, s, s)
    with pytest.raises(built.WinogradB200Error, match="no sm_100"):
        built.Conv1x1Bn(np.zeros((32, 128), np.float32), np.ones(128, np.float32), np.ones(128, np.float32), True)


def test_bad_arguments_are_rejected(built):
    L = built.lib()
    h = ctypes.c_void_p()
    fp = ctypes.POINTER(ctypes.c_float)
    z = np.zeros(8, np.float32).ctypes.data_as(fp)
    assert L.wg_conv3x3_create(ctypes.byref(h), 7, 32, z, z, z, 1, 0, 0) == -1      # C not a multiple of 8
    assert L.wg_conv1x1_create(ctypes.byref(h), 32, 100, z, z, z, 1, 0, 0) == -1    # Cout not a multiple of 128
    assert L.wg_run(None, None, None, 1, 0, None) == -1
    assert L.wg_destroy(None) == -1
    assert b"argument" in L.wg_strerror(-1)


def test_shard_range_partitions_the_batch(wg):
    for n in (1, 7, 255, 256, 257):
        for world in (1, 2, 3, 4, 8):
            r = [wg.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, torch, torch.distributed as dist
import wg_loader, golden
wg = wg_loader.load()
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[3])
dist.init_process_group("gloo", rank=rank, world_size=world)
N = 5
rs = np.random.RandomState(3)
x = (rs.rand(N, 196, 32) - 0.5).astype(np.float32)
w = (rs.rand(32, 128) - 0.5).astype(np.float32)
sc = rs.rand(128).astype(np.float32); sh = rs.rand(128).astype(np.float32)
lo, hi = wg.shard_range(N, rank, world)
# the per-rank compute is stood in for by the ORACLE here (CPU test of the host-side shard/gather logic only)
y_local = torch.from_numpy(golden.conv1x1_bn(x[lo:hi], w, sc, sh, True))
y = wg.gather_output(y_local, N)
full = golden.conv1x1_bn(x, w, sc, sh, True)
assert y.shape == full.shape, (y.shape, full.shape)
assert np.array_equal(y.numpy(), full)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_shard_and_gather_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT})
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", port], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r} ok" in o, o


def test_product_library_has_no_developer_kernels(built):
    """The product .so ships no ablation instantiation (wino3x3_ff_kernel<..., DBG = true, ...>), no CTA-pair or
    superseded-generation kernel, reads no environment variable (no getenv import) and exports no developer knob."""
    syms = subprocess.run(["nm", "-D", "-C", built.LIB_PATH], capture_output=True, text=True).stdout
    kernels = [l for l in syms.splitlines() if "wg::wino3x3" in l or "wg::conv1x1" in l]
    assert kernels, "kernel symbols expected in the dynamic table"
    for l in kernels:
        m = re.search(r"wino3x3_ff_kernel<(\w+), (\w+), (\w+), (\w+)>", l)
        if m:
            assert m.group(2) == "false", l        # DBG
            assert m.group(3) == "true", l         # parity-plane raw layout only
            assert m.group(4) == "false", l        # no CTA pairs
    assert "wino3x3_tm_kernel" not in syms
    assert "wg_dev_set_wino_kn" not in syms and "wg_set_wino_kn" not in syms
    # (getenv itself is still imported -- the statically linked CUDA runtime reads CUDA_* variables -- but none of the
    #  library's own knob names survives in the product binary)
    strs = subprocess.run(["strings", built.LIB_PATH], capture_output=True, text=True).stdout.splitlines()
    assert not [t for t in strs if re.fullmatch(r"WG_[A-Z0-9_]+", t)], "the product library must not read WG_* knobs"


@pytest.mark.parametrize("n", [1, 15, 16, 17, 63, 64, 100, 255, 256, 257, 300, 2560, 2561, 3584, 100000, 300000,
                               458000, 500000, 1000000, 2 ** 31 - 1])
def test_host_chunk_schedule_is_bounded(built, n):
    """wg_run_host's chunk schedule (65+ chunks overflowed fixed 64-entry arrays for N >~ 458k): at most
    64 chunks for any N, positive sizes summing to N, small chunks at both ends (the first copy-in and the last copy-out
    are what the pipeline exposes); 256 images -> 16, 32, 64, 64, 40, 20, 20."""
    s = built.host_chunk_schedule(n)
    assert 1 <= len(s) <= 64 and all(c > 0 for c in s) and sum(s) == n
    if n == 256:
        assert s == [16, 32, 64, 64, 40, 20, 20]
    if n >= 128:
        assert s[-1] < 32 or 2 * s[-1] <= max(s)                   # the exposed last chunk is never the big one
        assert s[0] <= 16 or 2 * s[0] <= max(s) or len(s) == 64    # nor is the first (nothing is copied out before it)
    # cap smaller than the schedule: still returns the count, writes only `cap` entries
    buf = (ctypes.c_int * 2)(-1, -1)
    assert built.lib().wg_host_chunk_schedule(n, buf, 2) == len(s)
    assert [buf[i] for i in range(min(2, len(s)))] == s[:2]
    assert built.lib().wg_host_chunk_schedule(0, buf, 2) == 0


def test_direct_kernel_geometry_invariants(built):
    """wg_direct_geometry (host-only): how the direct-convolution 3x3 kernels cut a map into work items. For every map
    size the frame supports: the MMA N is a multiple of 16 and at most 256, the halo covers one frame row + 1 pixel and
    keeps the TMA boxes 1024-byte aligned, the boxes cover item + halo on both sides and fit the kernel's stage (384
    rows TF32, 304 rows 16-bit operands), and the items of an image cover all of its output rows."""
    assert built.direct_geometry(14, 14) == dict(R=14, bands=1, G=1, n_pad=224, halo=24, n_boxes=2, box_rows=136, Hf=16, Wf=16)
    g28 = built.direct_geometry(28, 28)
    assert (g28["R"], g28["bands"], g28["n_pad"]) == (8, 4, 240)
    g7 = built.direct_geometry(7, 7)
    assert (g7["bands"], g7["G"], g7["n_pad"]) == (1, 2, 176)
    assert built.direct_geometry(56, 56)["R"] == 4 and built.direct_geometry(56, 56, built.WG_BF16)["R"] == 3
    seen = 0
    for dtype, max_rows in ((built.WG_TF32, 384), (built.WG_BF16, 304)):
        for h in range(3, 105, 1 if dtype == built.WG_TF32 else 7):
            for w in range(3, 105, 3):
                g = built.direct_geometry(h, w, dtype)
                if g is None:
                    continue
                seen += 1
                if (h, w) == (14, 14):
                    continue
                hf, wf = built.frame_dims(h, w)
                assert (g["Hf"], g["Wf"]) == (hf, wf)
                assert g["n_pad"] % 16 == 0 and 16 <= g["n_pad"] <= 256
                assert g["halo"] % 8 == 0 and g["halo"] >= wf + 1
                rows = g["n_boxes"] * g["box_rows"]
                assert g["box_rows"] % 8 == 0 and g["box_rows"] <= 256 and g["n_pad"] + 2 * g["halo"] <= rows <= max_rows
                if g["bands"] == 1 and g["R"] == 0:      # whole images per item
                    assert g["G"] >= 2 and (g["G"] - 1) * hf * wf + h * wf <= g["n_pad"] < (g["G"] - 1) * hf * wf + h * wf + 16
                else:                                     # row bands of one image
                    assert g["G"] == 1 and 1 <= g["R"] <= h and g["bands"] * g["R"] >= h > (g["bands"] - 1) * g["R"]
                    assert g["R"] * wf <= g["n_pad"] < g["R"] * wf + 16
    assert seen > 1000
    buf = (ctypes.c_int * 9)()
    assert built.lib().wg_direct_geometry(2, 14, 0, buf) < 0 and built.lib().wg_direct_geometry(14, 14, 7, buf) < 0


def test_blob_and_residual_calls_fail_loudly_without_gpu(built):
    """No GPU here: deserialising a well-formed header must end in WG_ERR_NODEVICE or WG_ERR_IO, never in a CPU path."""
    L = built.lib()
    h = ctypes.c_void_p()
    assert L.wg_layer_deserialize(ctypes.byref(h), b"x" * 16, 16, 0) == -7          # WG_ERR_IO: shorter than a header
    assert L.wg_layer_load(ctypes.byref(h), b"/nonexistent/blob.wgb", 0) == -7
    assert L.wg_run_residual(None, None, None, None, 1, 0, None) == -1               # WG_ERR_ARG
    tf = ctypes.c_double()
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if not have_gpu:
        assert L.wg_measure_tensor_peak(0, 0, ctypes.byref(tf), None) == -6          # WG_ERR_NODEVICE


def test_frame_dims_match_the_oracle(built):
    """f4: frame of an H x W map = one border pixel all round, rows / columns rounded up to even (host-only logic)."""
    import golden
    for h, w in [(14, 14), (28, 28), (56, 56), (7, 7), (3, 3), (9, 13), (104, 80), (8, 5)]:
        assert built.frame_dims(h, w) == golden.frame_dims(h, w)
    assert built.frame_dims(14, 14) == (16, 16) and built.frame_dims(7, 7) == (10, 10)
    for bad in ((2, 14), (14, 2), (112, 112)):     # too small to tile / wider than one raw-tile plane holds (W <= 104)
        with pytest.raises(built.WinogradB200Error):
            built.frame_dims(*bad)
