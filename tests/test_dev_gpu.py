"""Developer-build checks (-m gpu): tools/libwinograd_b200_dev.so keeps the superseded 3x3 kernel generations, the
CTA-pair experiments and the WG_* A/B knobs that the profiles under profiles/ were taken with. None of that is in the
product library (tests/test_abi.py::test_product_library_has_no_developer_kernels); these tests only keep the
developer build honest: every variant still agrees with the oracle. Knobs are read once per process, hence subprocesses.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV_LIB = os.path.join(ROOT, "tools", "libwinograd_b200_dev.so")


def _run_ff_check(env_extra, kns="96"):
    assert os.path.exists(DEV_LIB), "developer library must be prebuilt in-tree (make dev / __graft_entry__.build())"
    env = dict(os.environ, WG_B200_DEV_LIB="1", **env_extra)
    tag = "_".join(f"{k}{v}" for k, v in env_extra.items()).lower() or "plain"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ff_check.py"), "--quick", "--kns", kns,
                        "--iters", "3", "--out", os.path.join(ROOT, "gpurun_out", f"ff_check_dev_{tag}_{kns}.json")],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("knob", ["WG_FF_CG2=1", "WG_FF_W16=1", "WG_FF_SPLIT=2", "WG_FF_P9=0"])
def test_full_fold_kernel_experiment_knobs(knob):
    """WG_FF_CG2=1 = CTA pairs (tcgen05 cta_group::2), WG_FF_W16=1 = sixteen transform warps for every layer,
    WG_FF_SPLIT=2 = split-C for every channel count, WG_FF_P9=0 = single-box raw layout."""
    k, v = knob.split("=")
    _run_ff_check({k: v})


@pytest.mark.parametrize("kn", [48, 64])
def test_superseded_kernel_generations(kn):
    """kn = 48: V in TMEM, half fold (wino_tm_kernel.cu); kn = 64: both operands in shared memory (winograd_kernels.cu)."""
    _run_ff_check({}, kns=str(kn))


def test_1x1_cta_pair_variant():
    """WG_ONE_PAIR=1: the 1x1 throughput kernel as tcgen05 cta_group::2 pairs (experiment that did not pay)."""
    env = dict(os.environ, WG_B200_DEV_LIB="1", WG_ONE_PAIR="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_parity_gpu.py"), "-q", "-x",
                        "-m", "gpu", "-k", "1x1_ragged or 1x1_padded"],
                       env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_winograd_throughput_kernel_behind_the_direct_engine():
    """WG_3X3_DIRECT_MIN=1000000 keeps every TF32 batch on the fused Winograd kernels (what the product ran before the
    direct-convolution engine): the full-fold kernel still agrees with the oracle on the shapes that now default to the
    direct engine, N = 256 included (tools/ff_check.py: parity on awkward shapes + sampled images at N = 256)."""
    _run_ff_check({"WG_3X3_DIRECT_MIN": "1000000"})


def test_direct_kernel_cluster_variants():
    """conv3x3_direct_kernel<2> / <4>: clusters of 2 / 4 images sharing each weight block by multicast (measured no faster
    than CL = 1, developer build only): still agree with the oracle, odd batches included (the last group repeats the
    last image)."""
    assert os.path.exists(DEV_LIB)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "direct_check.py"), "--ns", "3,37", "--bo", "2,4,10",
                        "--shapes", "128x128,64x256", "--iters", "3"],
                       env=dict(os.environ), capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
