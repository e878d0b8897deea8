"""Host-side mirror of the reference's surface over libwinograd_b200.so (ctypes; no torch types cross the ABI).

The reference's whole public API is six zero-argument C functions that read data/<name>.bin (Test.c:25-43,
Kernel128_winograd.h:20, Kernel256_winograd.h:20, Kernel128_one.h:18-19, Kernel256_one.h:18-19). They are re-exported
here under the same names (`kernel_128()` ... `kernel_256_1_out()`), next to the tensor-level calls they wrap
(`Conv3x3BnRelu`, `Conv1x1Bn` -> wg_conv3x3_create / wg_conv1x1_create / wg_run, include/winograd_b200.h).

PyTorch is used only as plumbing (device buffers, streams, torch.distributed); the arithmetic is the hand-written
sm_100a kernels in csrc/. There is NO CPU fallback: if the shared library is missing, or no B200 is visible, the
calls raise.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libwinograd_b200.so")

WG_TF32, WG_BF16, WG_FP16 = 0, 1, 2

# every symbol include/winograd_b200.h, include/wg_legacy.h and include/util.h declare
ABI_SYMBOLS = (
    "wg_conv3x3_create", "wg_conv1x1_create", "wg_run", "wg_run_host", "wg_destroy", "wg_layer_info",
    "wg_launch_count", "wg_strerror", "wg_last_cuda_error", "wg_device_count", "wg_fold_bn", "wg_set_max_ctas",
    "wg_set_wino_kn",
    "kernel_128", "kernel_256", "kernel_128_1_in", "kernel_128_1_out", "kernel_256_1_in", "kernel_256_1_out",
    "wg_set_baseline_hook", "wg_legacy_last_output",
    "get_parameter", "transpose", "getTimeMicroseconds64", "output_checker",
)


class WinogradB200Error(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU). Returns its path."""
    r = subprocess.run(["make", "-C", _ROOT, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise WinogradB200Error("building libwinograd_b200.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """The loaded C-ABI library. Raises if it has not been built -- there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WinogradB200Error(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()); "
                                    "this package has no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        c_fp = ctypes.POINTER(ctypes.c_float)
        L.wg_conv3x3_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, c_fp, c_fp, c_fp,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.wg_conv1x1_create.argtypes = L.wg_conv3x3_create.argtypes
        L.wg_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                             ctypes.c_void_p]
        L.wg_run_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.wg_destroy.argtypes = [ctypes.c_void_p]
        L.wg_layer_info.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_int)] * 4
        L.wg_launch_count.restype = ctypes.c_longlong
        L.wg_strerror.restype = ctypes.c_char_p
        L.wg_strerror.argtypes = [ctypes.c_int]
        L.wg_last_cuda_error.restype = ctypes.c_char_p
        L.wg_fold_bn.argtypes = [ctypes.c_int, c_fp, c_fp, c_fp, c_fp, ctypes.c_float, c_fp, c_fp]
        L.wg_fold_bn.restype = None
        L.wg_set_max_ctas.argtypes = [ctypes.c_int]
        L.wg_set_wino_kn.argtypes = [ctypes.c_int]
        L.wg_legacy_last_output.argtypes = [c_fp, ctypes.c_int]
        _lib = L
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        L = lib()
        raise WinogradB200Error(f"{what}: {L.wg_strerror(rc).decode()} [{L.wg_last_cuda_error().decode()}]")


def _fptr(a):
    import numpy as np
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def device_count() -> int:
    return int(lib().wg_device_count())


def launch_count() -> int:
    return int(lib().wg_launch_count())


def fold_bn(gamma, beta, mean, var, eps=1e-5):
    """Folded BN exactly as data_generator.py:41-47 (float32)."""
    import numpy as np
    g, b, m, v = (np.ascontiguousarray(a, np.float32) for a in (gamma, beta, mean, var))
    sc, sh = np.empty_like(g), np.empty_like(g)
    lib().wg_fold_bn(len(g), _fptr(g), _fptr(b), _fptr(m), _fptr(v), ctypes.c_float(eps), _fptr(sc), _fptr(sh))
    return sc, sh


class _Layer:
    """A fused layer living on one GPU: packed filter + folded BN on the device, one kernel launch per call."""
    kind = -1

    def __init__(self, cin, cout, w, scale, shift, relu, device=0, dtype=WG_TF32):
        import numpy as np
        self.cin, self.cout, self.relu, self.device = int(cin), int(cout), bool(relu), int(device)
        w = np.ascontiguousarray(w, np.float32)
        scale = np.ascontiguousarray(scale, np.float32)
        shift = np.ascontiguousarray(shift, np.float32)
        assert scale.shape == (cout,) and shift.shape == (cout,)
        self._h = ctypes.c_void_p()
        create = lib().wg_conv3x3_create if self.kind == 0 else lib().wg_conv1x1_create
        _check(create(ctypes.byref(self._h), cin, cout, _fptr(w), _fptr(scale), _fptr(shift), int(relu), dtype,
                      device), "create")

    # -- device tensors (torch is only the allocator / stream provider here)
    def __call__(self, x, out=None, out_padded=False):
        import torch
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.device.index == self.device
        n = x.shape[0]
        assert tuple(x.shape[1:]) == self.in_shape(), f"expected [N,{self.in_shape()}], got {tuple(x.shape)}"
        oshape = (n,) + self.out_shape(out_padded)
        if out is None:
            out = torch.empty(oshape, device=x.device, dtype=torch.float32)
        assert tuple(out.shape) == oshape and out.is_contiguous() and out.dtype == torch.float32
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _check(lib().wg_run(self._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), n,
                            int(bool(out_padded)), ctypes.c_void_p(stream)), "wg_run")
        return out

    # -- host buffers, end to end (H2D + kernel + D2H inside)
    def run_host(self, x_host, y_host=None, out_padded=False):
        import numpy as np
        x_host = np.ascontiguousarray(x_host, np.float32)
        n = x_host.shape[0]
        assert tuple(x_host.shape[1:]) == self.in_shape()
        if y_host is None:
            y_host = np.empty((n,) + self.out_shape(out_padded), np.float32)
        _check(lib().wg_run_host(self._h, x_host.ctypes.data, y_host.ctypes.data, n, int(bool(out_padded))),
               "wg_run_host")
        return y_host

    def run_host_ptr(self, x_ptr: int, y_ptr: int, n: int, out_padded=False):
        """Raw host pointers (e.g. pinned torch tensors' data_ptr())."""
        _check(lib().wg_run_host(self._h, ctypes.c_void_p(x_ptr), ctypes.c_void_p(y_ptr), n, int(bool(out_padded))),
               "wg_run_host")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().wg_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Conv3x3BnRelu(_Layer):
    """3x3 conv (Winograd F(2x2,3x3) on tcgen05) + folded BN + ReLU. x [N,16,16,C] NHWC incl. border, w [K,C,3,3]
    (the layouts of input_14_1_C.bin / weight_NCHW_C_K.bin) -> [N,14,14,K] or the zero-bordered [N,16,16,K] frame.
    Replaces kernel_128()/kernel_256()'s three launches (Kernel128_winograd.cu:263-265)."""
    kind = 0

    def __init__(self, w_kcrs, scale, shift, relu=True, device=0, dtype=WG_TF32):
        """dtype = WG_TF32 (default; tolerance 1e-3) or WG_BF16 (bf16 V/U operands, fp32 I/O and accumulation;
        tolerance 1e-2; needs C % 16 == 0 and K % 64 == 0)."""
        k, c = w_kcrs.shape[0], w_kcrs.shape[1]
        assert tuple(w_kcrs.shape) == (k, c, 3, 3)
        super().__init__(c, k, w_kcrs, scale, shift, relu, device, dtype)

    def in_shape(self):
        return (16, 16, self.cin)

    def out_shape(self, out_padded=False):
        return (16, 16, self.cout) if out_padded else (14, 14, self.cout)


class Conv1x1Bn(_Layer):
    """1x1 conv (GEMM on tcgen05) + folded BN (+ ReLU). x [N,196,Cin], w [Cin,Cout] (input_one_14_1024.bin /
    weight_one_1024.bin prefixes) -> [N,196,Cout]. Replaces kernel_512_one_128 etc. (Kernel128_one.cu:98,316;
    Kernel256_one.cu:100,318)."""
    kind = 1

    def __init__(self, w_cin_cout, scale, shift, relu, device=0):
        cin, cout = w_cin_cout.shape
        super().__init__(cin, cout, w_cin_cout, scale, shift, relu, device)

    def in_shape(self):
        return (196, self.cin)

    def out_shape(self, out_padded=False):
        # out_padded: the zero-bordered frame a following 3x3 layer reads (chain mode)
        return (16, 16, self.cout) if out_padded else (196, self.cout)


class Bottleneck:
    """ResNet bottleneck chain 1x1 (Cin->C, +BN+ReLU) -> 3x3 (C->C, +BN+ReLU) -> 1x1 (C->Cout, +BN, no ReLU), the
    three layer kinds of the reference chained the way its layouts suggest (SURVEY.md section 8f rank 1;
    BASELINE.json configs[4]): the first 1x1 writes the zero-bordered 16x16 frame (Kernel128_winograd.cu:163,243
    layout) that the 3x3 consumes, the 3x3 writes dense [N,14,14,C] = [N,196,C] for the last 1x1. Three kernel
    launches, intermediates stay in L2/HBM, no padding or layout pass in between. The residual add is not part of
    the reference (its `_out` kernels stop before it, Kernel128_one.cu:272) and is not done here."""

    def __init__(self, w1, s1, b1, w3, s3, b3, w2, s2, b2, device=0, dtype=WG_TF32):
        self.l1 = Conv1x1Bn(w1, s1, b1, relu=True, device=device)
        self.l3 = Conv3x3BnRelu(w3, s3, b3, relu=True, device=device, dtype=dtype)
        self.l2 = Conv1x1Bn(w2, s2, b2, relu=False, device=device)
        assert self.l1.cout == self.l3.cin and self.l3.cout == self.l2.cin
        self._bufs = {}

    def __call__(self, x, out=None):
        import torch
        n = x.shape[0]
        key = (n, x.device.index)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty((n, 16, 16, self.l1.cout), device=x.device),
                               torch.empty((n, 14, 14, self.l3.cout), device=x.device))
        frame, mid = self._bufs[key]
        self.l1(x, out=frame, out_padded=True)
        self.l3(frame, out=mid)
        return self.l2(mid.view(n, 196, self.l3.cout), out=out)

    def capture(self, x, out=None):
        """Record the three launches on `x` into a CUDA graph (the launches are plain stream work: no allocation, no
        sync inside wg_run) and return (replay, out): `replay()` re-runs the chain on the current contents of `x`
        with ONE graph launch -- the launch-bound small-batch case. `x` and `out` must stay alive and in place."""
        import torch
        n = x.shape[0]
        if out is None:
            out = torch.empty((n, 196, self.l2.cout), device=x.device)
        self(x, out=out)                      # warm-up: tensor maps cached, kernels configured, buffers allocated
        torch.cuda.synchronize(x.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self(x, out=out)
        return graph.replay, out


# ---------------------------------------------------------------------------------------------------------------------
# The reference's six entry points, same names, same behaviour (read data/*.bin from the CWD, print the same lines,
# return (mine_us << 16) | baseline_us). Thin pass-throughs to the C symbols.
def _legacy(name):
    def f():
        return int(getattr(lib(), name)())
    f.__name__ = name
    f.__doc__ = f"C entry point `{name}` of libwinograd_b200.so (include/wg_legacy.h)."
    return f


kernel_128 = _legacy("kernel_128")
kernel_256 = _legacy("kernel_256")
kernel_128_1_in = _legacy("kernel_128_1_in")
kernel_128_1_out = _legacy("kernel_128_1_out")
kernel_256_1_in = _legacy("kernel_256_1_in")
kernel_256_1_out = _legacy("kernel_256_1_out")
LEGACY_ENTRIES = (kernel_128, kernel_256, kernel_128_1_in, kernel_128_1_out, kernel_256_1_in, kernel_256_1_out)


def legacy_last_output(cout: int):
    """Dense [196][cout] copy of what the last legacy entry point computed."""
    import numpy as np
    out = np.empty(196 * cout, np.float32)
    n = lib().wg_legacy_last_output(_fptr(out), out.size)
    assert n == out.size, (n, out.size)
    return out.reshape(196, cout)


# ---------------------------------------------------------------------------------------------------------------------
# Batch sharding across the GPUs of one box (one process per GPU, torch.distributed for the plumbing). Images are
# independent and inference BN has no cross-sample statistics, so the hot path needs no collective; the only exchange
# is one gather of the output.
def shard_range(n_total: int, rank: int, world: int):
    """Contiguous image range [lo, hi) of `rank`: the first n_total % world ranks get one extra image."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_output(y_local, n_total: int, group=None):
    """all_gather of the per-rank output shards ([n_r, ...]) into [n_total, ...] on every rank (NCCL on GPUs,
    gloo on CPU tensors in the tests). Uneven shards are padded to the largest one for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    assert y_local.shape[0] == hi - lo
    pad = torch.zeros((nmax,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    pad[:hi - lo] = y_local
    out = torch.empty((world * nmax,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(h - l == nmax for l, h in sizes):
        return out
    return torch.cat([out[r * nmax:r * nmax + (h - l)] for r, (l, h) in enumerate(sizes)], dim=0)


WG_OUT_PADDED, WG_OUT_MULTICAST = 1, 2


class FusedGatherConv3x3:
    """conv3x3 + BN + ReLU fused with the all-gather of its output over NVSwitch (SURVEY.md section 8e names one gather
    of the fp32 output as the path's only exchange). Every rank owns `n_local` images; the gathered
    [world * n_local, 14, 14, K] tensor lives in torch symmetric memory bound to an NVLS multicast object, and each rank's
    kernel writes its shard through the MULTICAST address with multimem.st (wg_run flag WG_OUT_MULTICAST): the switch
    replicates every 16-byte store into all GPUs' copies, so the gather costs no extra kernel, no extra HBM read and
    overlaps the convolution. `__call__` returns the gathered tensor after a cross-rank barrier on the stream.
    Needs multicast support (NVSwitch + fabric manager); otherwise construction raises and gather_output() (NCCL) is
    the fallback."""

    def __init__(self, layer, n_local: int, group=None, out_padded=False):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        assert layer.kind == 0, "3x3 layers only"
        self.layer, self.n_local, self.out_padded = layer, n_local, bool(out_padded)
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        shape = (self.world * n_local,) + layer.out_shape(out_padded)
        dev = torch.device("cuda", layer.device)
        self.full = symm_mem.empty(shape, dtype=torch.float32, device=dev)
        self.hdl = symm_mem.rendezvous(self.full, self.group)
        if not self.hdl.multicast_ptr:
            raise WinogradB200Error("no NVLS multicast support on this system; use gather_output() (NCCL) instead")
        self.shard_bytes = n_local * int(np_prod(layer.out_shape(out_padded))) * 4
        self.y_mc = self.hdl.multicast_ptr + self.rank * self.shard_bytes

    def __call__(self, x):
        import torch
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[0] == self.n_local
        stream = torch.cuda.current_stream(x.device).cuda_stream
        flags = WG_OUT_MULTICAST | (WG_OUT_PADDED if self.out_padded else 0)
        _check(lib().wg_run(self.layer._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(self.y_mc), self.n_local,
                            flags, ctypes.c_void_p(stream)), "wg_run (multicast)")
        self.hdl.barrier()          # all ranks' kernels are complete: every shard has landed everywhere
        return self.full


def np_prod(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n
