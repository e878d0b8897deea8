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
# developer build (make dev): superseded kernel generations, ablation instantiations, WG_* environment knobs. Only the
# tools/ scripts and the developer tests select it, with WG_B200_DEV_LIB=1 in the environment.
DEV_LIB_PATH = os.path.join(_ROOT, "tools", "libwinograd_b200_dev.so")
IS_DEV_LIB = os.environ.get("WG_B200_DEV_LIB", "0") == "1"
if IS_DEV_LIB:
    LIB_PATH = DEV_LIB_PATH

WG_TF32, WG_BF16, WG_FP16 = 0, 1, 2
WG_OUT_PADDED, WG_OUT_MULTICAST, WG_OUT_RELU_AFTER_ADD, WG_OUT_INTERIOR_ONLY = 1, 2, 4, 8

# every symbol include/winograd_b200.h, include/wg_legacy.h and include/util.h declare
ABI_SYMBOLS = (
    "wg_conv3x3_create", "wg_conv1x1_create", "wg_conv3x3_create_hw", "wg_conv1x1_create_hw", "wg_layer_geometry",
    "wg_frame_dims", "wg_direct_geometry", "wg_run", "wg_run_residual", "wg_run_host", "wg_host_chunk_schedule",
    "wg_destroy", "wg_layer_info", "wg_layer_serialize", "wg_layer_deserialize", "wg_layer_save", "wg_layer_load",
    "wg_launch_count", "wg_strerror", "wg_last_cuda_error", "wg_device_count", "wg_fold_bn", "wg_set_max_ctas",
    "wg_measure_tensor_peak",
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
        L.wg_conv3x3_create_hw.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, c_fp, c_fp, c_fp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.wg_conv1x1_create_hw.argtypes = L.wg_conv3x3_create_hw.argtypes
        L.wg_layer_geometry.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_int)] * 4
        L.wg_frame_dims.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 2
        L.wg_direct_geometry.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.wg_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                             ctypes.c_void_p]
        L.wg_run_residual.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_void_p]
        L.wg_run_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.wg_host_chunk_schedule.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        L.wg_layer_serialize.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.POINTER(ctypes.c_size_t)]
        L.wg_layer_deserialize.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_size_t,
                                           ctypes.c_int]
        L.wg_layer_save.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.wg_layer_load.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_int]
        L.wg_measure_tensor_peak.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                             ctypes.POINTER(ctypes.c_double)]
        L.wg_destroy.argtypes = [ctypes.c_void_p]
        L.wg_layer_info.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_int)] * 4
        L.wg_launch_count.restype = ctypes.c_longlong
        L.wg_strerror.restype = ctypes.c_char_p
        L.wg_strerror.argtypes = [ctypes.c_int]
        L.wg_last_cuda_error.restype = ctypes.c_char_p
        L.wg_fold_bn.argtypes = [ctypes.c_int, c_fp, c_fp, c_fp, c_fp, ctypes.c_float, c_fp, c_fp]
        L.wg_fold_bn.restype = None
        L.wg_set_max_ctas.argtypes = [ctypes.c_int]
        if IS_DEV_LIB:
            L.wg_dev_set_wino_kn.argtypes = [ctypes.c_int]
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


def frame_dims(h: int, w: int):
    """(Hf, Wf) of the input frame of an h x w map (host-only; wg_frame_dims)."""
    hf, wf = ctypes.c_int(), ctypes.c_int()
    _check(lib().wg_frame_dims(int(h), int(w), ctypes.byref(hf), ctypes.byref(wf)), "wg_frame_dims")
    return hf.value, wf.value


def direct_geometry(h: int, w: int, dtype: int = WG_TF32):
    """Work-item geometry of the direct-convolution 3x3 kernels for an h x w map (host-only; wg_direct_geometry):
    dict(R, bands, G, n_pad, halo, n_boxes, box_rows, Hf, Wf), or None if the map does not fit those kernels."""
    out = (ctypes.c_int * 9)()
    rc = int(lib().wg_direct_geometry(int(h), int(w), int(dtype), out))
    if rc < 0:
        _check(rc, "wg_direct_geometry")
    if rc == 0:
        return None
    return dict(zip(("R", "bands", "G", "n_pad", "halo", "n_boxes", "box_rows", "Hf", "Wf"), [int(v) for v in out]))


def host_chunk_schedule(n: int):
    """Chunk sizes wg_run_host uses for a batch of n images (host-only logic; no GPU needed)."""
    buf = (ctypes.c_int * 64)()
    k = int(lib().wg_host_chunk_schedule(int(n), buf, 64))
    assert 0 <= k <= 64, k
    return [int(buf[i]) for i in range(k)]


def measure_tensor_peak(dtype=WG_TF32, device=0):
    """(TFLOP/s, clocks per M=128 N=256 MMA) of back-to-back tcgen05.mma on every SM of `device` (wg_measure_tensor_peak)."""
    tf, clk = ctypes.c_double(), ctypes.c_double()
    _check(lib().wg_measure_tensor_peak(device, dtype, ctypes.byref(tf), ctypes.byref(clk)), "wg_measure_tensor_peak")
    return tf.value, clk.value


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

    def __init__(self, cin, cout, w, scale, shift, relu, device=0, dtype=WG_TF32, hw=(14, 14)):
        import numpy as np
        self.cin, self.cout, self.relu, self.device = int(cin), int(cout), bool(relu), int(device)
        self.dtype = int(dtype)
        self.h, self.w = int(hw[0]), int(hw[1])
        w = np.ascontiguousarray(w, np.float32)
        scale = np.ascontiguousarray(scale, np.float32)
        shift = np.ascontiguousarray(shift, np.float32)
        assert scale.shape == (cout,) and shift.shape == (cout,)
        self._h = ctypes.c_void_p()
        create = lib().wg_conv3x3_create_hw if self.kind == 0 else lib().wg_conv1x1_create_hw
        _check(create(ctypes.byref(self._h), cin, cout, self.h, self.w, _fptr(w), _fptr(scale), _fptr(shift), int(relu),
                      dtype, device), "create")
        self._read_geometry()

    def _read_geometry(self):
        h, w, hf, wf = (ctypes.c_int() for _ in range(4))
        _check(lib().wg_layer_geometry(self._h, ctypes.byref(h), ctypes.byref(w), ctypes.byref(hf), ctypes.byref(wf)),
               "wg_layer_geometry")
        self.h, self.w, self.hf, self.wf = h.value, w.value, hf.value, wf.value

    # -- packed blob (wg_layer_save / wg_layer_load): cold start without the filter transform
    def save(self, path):
        _check(lib().wg_layer_save(self._h, os.fsencode(path)), "wg_layer_save")

    def serialize(self) -> bytes:
        need = ctypes.c_size_t()
        _check(lib().wg_layer_serialize(self._h, None, 0, ctypes.byref(need)), "wg_layer_serialize")
        buf = ctypes.create_string_buffer(need.value)
        _check(lib().wg_layer_serialize(self._h, buf, need.value, ctypes.byref(need)), "wg_layer_serialize")
        return buf.raw

    @classmethod
    def _from_handle(cls, h, device):
        kind, cin, cout, relu = (ctypes.c_int() for _ in range(4))
        _check(lib().wg_layer_info(h, ctypes.byref(kind), ctypes.byref(cin), ctypes.byref(cout), ctypes.byref(relu)),
               "wg_layer_info")
        sub = Conv3x3BnRelu if kind.value == 0 else Conv1x1Bn
        self = sub.__new__(sub)
        self.cin, self.cout, self.relu, self.device = cin.value, cout.value, bool(relu.value), int(device)
        self.dtype = WG_TF32
        self._h = h
        self._read_geometry()
        return self

    @staticmethod
    def load(path, device=0):
        """Layer from a blob written by save(): Conv3x3BnRelu or Conv1x1Bn, whichever the blob holds."""
        h = ctypes.c_void_p()
        _check(lib().wg_layer_load(ctypes.byref(h), os.fsencode(path), device), "wg_layer_load")
        return _Layer._from_handle(h, device)

    @staticmethod
    def deserialize(blob: bytes, device=0):
        h = ctypes.c_void_p()
        _check(lib().wg_layer_deserialize(ctypes.byref(h), blob, len(blob), device), "wg_layer_deserialize")
        return _Layer._from_handle(h, device)

    # -- device tensors (torch is only the allocator / stream provider here)
    def __call__(self, x, out=None, out_padded=False, residual=None, relu_after_add=False, interior_only=False):
        """One fused launch. residual (1x1 layers, dense output): y = act2(act(scale * conv + shift) + residual), the
        add that follows the reference's `_out` layers (Kernel128_one.cu:271-272), act2 = ReLU iff relu_after_add."""
        import torch
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.device.index == self.device
        n = x.shape[0]
        assert tuple(x.shape[1:]) == self.in_shape(), f"expected [N,{self.in_shape()}], got {tuple(x.shape)}"
        oshape = (n,) + self.out_shape(out_padded)
        if out is None:
            out = torch.empty(oshape, device=x.device, dtype=torch.float32)
        assert tuple(out.shape) == oshape and out.is_contiguous() and out.dtype == torch.float32
        assert out.device == x.device
        stream = torch.cuda.current_stream(x.device).cuda_stream
        flags = WG_OUT_PADDED if out_padded else 0
        if interior_only:       # 1x1 chain mode: `out`'s border is already zero (caller's guarantee), write the interior only
            assert out_padded
            flags |= WG_OUT_INTERIOR_ONLY
        if residual is not None:
            assert residual.is_cuda and residual.dtype == torch.float32 and residual.is_contiguous()
            assert residual.device == x.device and residual.numel() == out.numel() and not out_padded
            flags |= WG_OUT_RELU_AFTER_ADD if relu_after_add else 0
            _check(lib().wg_run_residual(self._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(residual.data_ptr()),
                                         ctypes.c_void_p(out.data_ptr()), n, flags, ctypes.c_void_p(stream)),
                   "wg_run_residual")
            return out
        _check(lib().wg_run(self._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), n, flags,
                            ctypes.c_void_p(stream)), "wg_run")
        return out

    # -- host buffers, end to end (H2D + kernel + D2H inside)
    def run_host(self, x_host, y_host=None, out_padded=False):
        import numpy as np
        x_host = np.ascontiguousarray(x_host, np.float32)
        n = x_host.shape[0]
        assert tuple(x_host.shape[1:]) == self.in_shape()
        if y_host is None:
            y_host = np.empty((n,) + self.out_shape(out_padded), np.float32)
        # the C side writes n * prod(out_shape) floats through this pointer: it must be exactly that buffer
        assert isinstance(y_host, np.ndarray) and y_host.dtype == np.float32 and y_host.flags["C_CONTIGUOUS"] \
            and y_host.flags["WRITEABLE"] and tuple(y_host.shape) == (n,) + self.out_shape(out_padded), \
            "y_host must be a writable C-contiguous float32 array of shape (n,) + out_shape(out_padded)"
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

    def __init__(self, w_kcrs, scale, shift, relu=True, device=0, dtype=WG_TF32, hw=(14, 14)):
        """dtype = WG_TF32 (default; tolerance 1e-3) or WG_BF16 (bf16 V/U operands, fp32 I/O and accumulation;
        tolerance 1e-2; needs C % 16 == 0 and K % 64 == 0). hw = (H, W) output map, default the reference's 14 x 14;
        the input frame is [N, hf, wf, C] with hf = H + 2 (H + 3 for odd H), see include/winograd_b200.h."""
        k, c = w_kcrs.shape[0], w_kcrs.shape[1]
        assert tuple(w_kcrs.shape) == (k, c, 3, 3)
        super().__init__(c, k, w_kcrs, scale, shift, relu, device, dtype, hw)

    def in_shape(self):
        return (self.hf, self.wf, self.cin)

    def out_shape(self, out_padded=False):
        return (self.hf, self.wf, self.cout) if out_padded else (self.h, self.w, self.cout)


class Conv1x1Bn(_Layer):
    """1x1 conv (GEMM on tcgen05) + folded BN (+ ReLU). x [N,196,Cin], w [Cin,Cout] (input_one_14_1024.bin /
    weight_one_1024.bin prefixes) -> [N,196,Cout]. Replaces kernel_512_one_128 etc. (Kernel128_one.cu:98,316;
    Kernel256_one.cu:100,318)."""
    kind = 1

    def __init__(self, w_cin_cout, scale, shift, relu, device=0, dtype=WG_TF32, hw=(14, 14)):
        """dtype = WG_TF32 (default; tolerance 1e-3) or WG_BF16 (bf16 operands: the activation stage is converted into
        tensor memory by four extra warps, bf16 weight image; fp32 I/O and accumulation; tolerance 1e-2).
        hw = (H, W) pixels per image (default 14 x 14 = the reference's 196)."""
        cin, cout = w_cin_cout.shape
        super().__init__(cin, cout, w_cin_cout, scale, shift, relu, device, dtype, hw)

    def in_shape(self):
        return (self.h * self.w, self.cin)

    def out_shape(self, out_padded=False):
        # out_padded: the zero-bordered frame a following 3x3 layer reads (chain mode)
        return (self.hf, self.wf, self.cout) if out_padded else (self.h * self.w, self.cout)


class Bottleneck:
    """ResNet bottleneck chain 1x1 (Cin->C, +BN+ReLU) -> 3x3 (C->C, +BN+ReLU) -> 1x1 (C->Cout, +BN, no ReLU), the
    three layer kinds of the reference chained the way its layouts suggest (SURVEY.md section 8f rank 1;
    BASELINE.json configs[4]): the first 1x1 writes the zero-bordered 16x16 frame (Kernel128_winograd.cu:163,243
    layout) that the 3x3 consumes, the 3x3 writes dense [N,14,14,C] = [N,196,C] for the last 1x1. Three kernel
    launches, intermediates stay in L2/HBM, no padding or layout pass in between.
    residual=True (needs Cin == Cout) completes the block: out = relu(chain(x) + x), the add and the final ReLU fused
    into the last 1x1 launch's epilogue (the reference's `_out` kernels stop right before it, Kernel128_one.cu:271-272,
    Kernel256_one.cu:273) -- still three launches."""

    def __init__(self, w1, s1, b1, w3, s3, b3, w2, s2, b2, device=0, dtype=WG_TF32, residual=False, hw=(14, 14)):
        self.l1 = Conv1x1Bn(w1, s1, b1, relu=True, device=device, hw=hw)
        self.l3 = Conv3x3BnRelu(w3, s3, b3, relu=True, device=device, dtype=dtype, hw=hw)
        self.l2 = Conv1x1Bn(w2, s2, b2, relu=False, device=device, hw=hw)
        self.px = self.l1.h * self.l1.w
        assert self.l1.cout == self.l3.cin and self.l3.cout == self.l2.cin
        self.residual = bool(residual)
        assert not self.residual or self.l1.cin == self.l2.cout, "the identity shortcut needs Cin == Cout"
        self._bufs = {}

    def __call__(self, x, out=None):
        import torch
        n = x.shape[0]
        key = (n, x.device.index)
        if key not in self._bufs:
            # the frame is zeroed ONCE here; every call rewrites its interior only (WG_OUT_INTERIOR_ONLY), the border stays
            self._bufs[key] = (torch.zeros((n,) + self.l1.out_shape(True), device=x.device),
                               torch.empty((n,) + self.l3.out_shape(), device=x.device))
        frame, mid = self._bufs[key]
        self.l1(x, out=frame, out_padded=True, interior_only=True)
        self.l3(frame, out=mid)
        if self.residual:
            return self.l2(mid.view(n, self.px, self.l3.cout), out=out, residual=x, relu_after_add=True)
        return self.l2(mid.view(n, self.px, self.l3.cout), out=out)

    def capture(self, x, out=None):
        """Record the three launches on `x` into a CUDA graph (the launches are plain stream work: no allocation, no
        sync inside wg_run) and return (replay, out): `replay()` re-runs the chain on the current contents of `x`
        with ONE graph launch -- the launch-bound small-batch case. `x` and `out` must stay alive and in place."""
        import torch
        n = x.shape[0]
        if out is None:
            out = torch.empty((n, self.px, self.l2.cout), device=x.device)
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
        """Returns the gathered tensor (symmetric memory, REUSED by the next call). Two cross-rank barriers on the
        stream per call: one BEFORE the kernel -- a fast rank's multimem.st must not land in a slower rank's copy while
        that rank's consumers of the previous result are still reading it (work enqueued on this stream before this
        call is covered; readers on other streams are the caller's to order) -- and one after it (all shards landed)."""
        import torch
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[0] == self.n_local
        assert x.device.index == self.layer.device
        stream = torch.cuda.current_stream(x.device).cuda_stream
        flags = WG_OUT_MULTICAST | (WG_OUT_PADDED if self.out_padded else 0)
        self.hdl.barrier()          # every rank has finished reading the previous gathered result
        _check(lib().wg_run(self.layer._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(self.y_mc), self.n_local,
                            flags, ctypes.c_void_p(stream)), "wg_run (multicast)")
        self.hdl.barrier()          # all ranks' kernels are complete: every shard has landed everywhere
        return self.full


def np_prod(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n
