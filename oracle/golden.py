"""CPU oracle for the fused conv + inference-BatchNorm + ReLU layer -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
the product (cuda-winograd_b200/) never does and fails loudly when its CUDA library is missing.

What it restates (the reference has no CPU implementation of the path -- its only oracle is cuDNN run in the same
process and compared by output_checker, util.c:46-63 -- so the golden follows the cuDNN call sequence the reference
treats as truth, fed by the files data_generator.py writes):

  3x3:  valid cross-correlation (CUDNN_CROSS_CORRELATION, pad 0, stride 1; Kernel128_winograd.cu:352) of the NHWC frame
        x[16][16][C] (:335) with the KCRS filter w[K][C][3][3] (:343) -> [14][14][K] (:339); BN inference
        gamma*(y-mean)/sqrt(var+eps)+beta (:391-394) with eps = 1e-5 (the value folded by data_generator.py:41-45);
        ReLU (:397-399).  The conv bias file is loaded but never used by any kernel (:217,249) -> ignored.
  1x1:  y[196][Cout] = x[196][Cin] . W[Cin][Cout] (Kernel128_one.cu:46-48), BN as above, ReLU for the two `_in`
        shapes only (Kernel128_one.cu:53 vs :272; Kernel256_one.cu:55 vs :273).
  folded form (what the reference kernels evaluate): act(scale' * conv + shift') with the *_winograd_* /
        *_myKernel_* files (Kernel128_winograd.cu:162-163, Kernel128_one.cu:52-53).

Pinning: the reference checks in no golden vectors and no tests. This oracle is pinned by
(1) an index-for-index emulation of the reference's own three CUDA kernels (`reference_pipeline_f4x4`, below)
agreeing with the direct form, (2) brute-force loops on small shapes (tests/test_oracle.py), and (3) on the GPU box,
the reference's kernels compiled from /root/reference into oracle/_ref/ and run on the same seeded files
(tests/test_reference_gpu.py).
"""
from __future__ import annotations

import numpy as np

EPS = 1e-5  # data_generator.py:40,106


# --------------------------------------------------------------------------------------------------------------- BN
def fold_bn(gamma, beta, mean, var, eps=EPS):
    """scale' = gamma/sqrt(var+eps), shift' = beta - gamma*mean/sqrt(var+eps), float32 like data_generator.py:41-47."""
    gamma, beta, mean, var = (np.asarray(a, np.float32) for a in (gamma, beta, mean, var))
    sd = np.sqrt(var + np.float32(eps))
    return (gamma / sd).astype(np.float32), (beta - gamma * mean / sd).astype(np.float32)


# ------------------------------------------------------------------------------------------------------ direct golden
def frame_dims(h, w):
    """Input frame of an H x W map (the f4 generalisation of the reference's fixed 14x14-in-16x16, Kernel128_winograd.cu:
    26-31): one border pixel all round, plus one extra trailing row / column for odd sizes so that frame rows pair up
    (F(2x2,3x3) tiles cover 2x2 outputs): Hf = 2*ceil(H/2) + 2."""
    return 2 * ((h + 1) // 2) + 2, 2 * ((w + 1) // 2) + 2


def conv3x3_raw(x, w, acc=np.float64, hw=None):
    """x [N,Hf,Wf,C] NHWC (border included), w [K,C,3,3] -> [N,H,W,K] valid cross-correlation, `acc` accumulation.
    hw = (H, W) output map; default: the frame minus its 1-pixel border (the reference: [N,16,16,C] -> [N,14,14,K])."""
    x = np.asarray(x)
    w = np.asarray(w)
    n, hf, wf, c = x.shape
    k = w.shape[0]
    h, wd = hw if hw is not None else (hf - 2, wf - 2)
    assert w.shape == (k, c, 3, 3) and (hf, wf) == frame_dims(h, wd), (x.shape, hw)
    out = np.zeros((n, h, wd, k), acc)
    xa = x.astype(acc, copy=False)
    for r in range(3):
        for s in range(3):
            patch = xa[:, r:r + h, s:s + wd, :].reshape(n * h * wd, c)
            out += (patch @ w[:, :, r, s].T.astype(acc)).reshape(n, h, wd, k)
    return out


def conv3x3_bn_relu(x, w, scale, shift, relu=True, acc=np.float64, hw=None):
    """Golden for ./Test 0 and 1 (folded form). Returns float32 [N,14,14,K] ([N,H,W,K] for other map sizes)."""
    y = conv3x3_raw(x, w, acc, hw) * np.asarray(scale, acc) + np.asarray(shift, acc)
    if relu:
        y = np.maximum(y, 0)
    return y.astype(np.float32)


def conv3x3_bn_relu_unfolded(x, w, gamma, beta, mean, var, eps=EPS, relu=True):
    """The cuDNN call sequence itself: conv -> BN inference -> ReLU (Kernel128_winograd.cu:382-399), float64."""
    y = conv3x3_raw(x, w, np.float64)
    g, b, m, v = (np.asarray(a, np.float64) for a in (gamma, beta, mean, var))
    y = g * (y - m) / np.sqrt(v + eps) + b
    if relu:
        y = np.maximum(y, 0)
    return y.astype(np.float32)


def conv1x1_bn(x, w, scale, shift, relu, acc=np.float64):
    """x [N,196,Cin] (or [N*196,Cin]), w [Cin,Cout] -> float32 [..., Cout]; golden for ./Test 2..5."""
    x = np.asarray(x)
    y = x.astype(acc, copy=False) @ np.asarray(w).astype(acc, copy=False)
    y = y * np.asarray(scale, acc) + np.asarray(shift, acc)
    if relu:
        y = np.maximum(y, 0)
    return y.astype(np.float32)


def pad_frame(y):
    """[N,14,14,K] -> the reference's zero-bordered [N,16,16,K] output frame (Kernel128_winograd.cu:163,243); for other
    map sizes [N,H,W,K] -> [N,Hf,Wf,K] (frame_dims)."""
    n, h, w, k = y.shape
    hf, wf = frame_dims(h, w)
    out = np.zeros((n, hf, wf, k), y.dtype)
    out[:, 1:h + 1, 1:w + 1, :] = y
    return out


def bottleneck_chain(x, w1, s1, b1, w3, s3, b3, w2, s2, b2, hw=(14, 14)):
    """1x1 (+BN+ReLU) -> zero-pad 1 -> 3x3 (+BN+ReLU) -> 1x1 (+BN, no ReLU): the three reference layer kinds chained
    (Kernel128_one.cu:24-54 -> Kernel128_winograd.cu:28-213 -> Kernel128_one.cu:244-273). x [N,196,Cin] -> [N,196,Cout].
    Unlike the reference's stand-alone 3x3 test input (random border, data_generator.py:49-53) the frame between the
    layers has a ZERO border -- that is what the reference's own padded output frame provides (Kernel128_winograd.cu:243)."""
    n = x.shape[0]
    h, w = hw
    a = conv1x1_bn(x, w1, s1, b1, True)
    frame = pad_frame(a.reshape(n, h, w, -1))
    m = conv3x3_bn_relu(frame, w3, s3, b3, True, hw=hw)
    return conv1x1_bn(m.reshape(n, h * w, -1), w2, s2, b2, False)


def conv1x1_bn_residual(x, w, scale, shift, relu, residual, relu_after_add, acc=np.float64):
    """The step after the reference's `_out` 1x1 layers, which stop right before it (Kernel128_one.cu:271-272 and
    Kernel256_one.cu:273 store scale*acc + shift without a ReLU because the residual add comes next):
    y = act2(act(scale * (x w) + shift) + residual), act = ReLU iff `relu`, act2 = ReLU iff `relu_after_add`."""
    y = np.asarray(x).astype(acc, copy=False) @ np.asarray(w).astype(acc, copy=False)
    y = y * np.asarray(scale, acc) + np.asarray(shift, acc)
    if relu:
        y = np.maximum(y, 0)
    y = y + np.asarray(residual).astype(acc, copy=False).reshape(y.shape)
    if relu_after_add:
        y = np.maximum(y, 0)
    return y.astype(np.float32)


def bottleneck_block(x, w1, s1, b1, w3, s3, b3, w2, s2, b2, hw=(14, 14)):
    """Full ResNet bottleneck block with the identity shortcut: relu(bottleneck_chain(x) + x); needs Cin == Cout
    (BASELINE.json configs[4]; the add + final ReLU are what the reference leaves out, see conv1x1_bn_residual)."""
    n = x.shape[0]
    h, w = hw
    a = conv1x1_bn(x, w1, s1, b1, True)
    frame = pad_frame(a.reshape(n, h, w, -1))
    m = conv3x3_bn_relu(frame, w3, s3, b3, True, hw=hw)
    return conv1x1_bn_residual(m.reshape(n, h * w, -1), w2, s2, b2, False, x, True)


# ------------------------------------------------------------------------------------------ brute force (tiny shapes)
def conv3x3_bn_relu_loops(x, w, scale, shift, relu=True):
    """Pure-Python loops, float64; only for tiny C, K in tests."""
    n, _, _, c = x.shape
    k = w.shape[0]
    out = np.zeros((n, 14, 14, k), np.float32)
    for b in range(n):
        for oy in range(14):
            for ox in range(14):
                for kk in range(k):
                    acc = 0.0
                    for r in range(3):
                        for s in range(3):
                            for cc in range(c):
                                acc += float(x[b, oy + r, ox + s, cc]) * float(w[kk, cc, r, s])
                    v = float(scale[kk]) * acc + float(shift[kk])
                    out[b, oy, ox, kk] = max(v, 0.0) if relu else v
    return out


# ----------------------------------------------------------------------------------------------- Winograd restatements
# F(2x2,3x3) (Lavin & Gray), the algorithm the B200 kernel runs:
BT_2 = np.array([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], np.float64)
G_2 = np.array([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], np.float64)
AT_2 = np.array([[1, 1, 1, 0], [0, 1, -1, -1]], np.float64)

# F(4x4,3x3), the algorithm the reference runs (Kernel128_winograd.cu:44-71 B^T, :138-147 A^T, data_generator.py:65 G)
BT_4 = np.array([[4, 0, -5, 0, 1, 0], [0, -4, -4, 1, 1, 0], [0, 4, -4, -1, 1, 0],
                 [0, -2, -1, 2, 1, 0], [0, 2, -1, -2, 1, 0], [0, 4, 0, -5, 0, 1]], np.float64)
G_4 = np.array([[0.25, 0, 0], [-1.0 / 6, -1.0 / 6, -1.0 / 6], [-1.0 / 6, 1.0 / 6, -1.0 / 6],
                [1.0 / 24, 1.0 / 12, 1.0 / 6], [1.0 / 24, -1.0 / 12, 1.0 / 6], [0, 0, 1]], np.float64)
AT_4 = np.array([[1, 1, 1, 1, 1, 0], [0, 1, -1, 2, -2, 0], [0, 1, 1, 4, 4, 0], [0, 1, -1, 8, -8, 1]], np.float64)


def filter_transform(w, G):
    """U[xi][c][k] = (G g G^T)[xi] for g = w[k][c]; layout [P*P][C][K] like weight_winograd_*.bin (data_generator.py:66-75)."""
    u = np.einsum('ir,kcrs,js->ijck', G, np.asarray(w, np.float64), G)
    p = G.shape[0]
    return u.reshape(p * p, w.shape[1], w.shape[0])


def winograd_f2x2(x, w, scale, shift, relu=True, operand_dtype=None):
    """Stage-by-stage F(2x2,3x3) restatement of the B200 kernel's arithmetic: V = B^T d B on 7x7 tiles of 4x4 at stride 2,
    16 point-GEMMs M = V.U, Y = A^T M A, folded BN, ReLU. `operand_dtype` = 'tf32' rounds V and U like the kernel does
    (round-to-nearest, ties away, 10 explicit mantissa bits) to predict its error; None keeps float64."""
    x = np.asarray(x, np.float64)
    n, _, _, c = x.shape
    k = w.shape[0]
    u = filter_transform(w, G_2)                                     # [16][C][K]
    tiles = np.empty((n, 7, 7, 4, 4, c))
    for dy in range(4):
        for dx in range(4):
            tiles[:, :, :, dy, dx, :] = x[:, dy:dy + 13:2, dx:dx + 13:2, :]
    v = np.einsum('iy,ntuyxc,jx->ntuijc', BT_2, tiles, BT_2).reshape(n * 49, 16, c)
    if operand_dtype == 'tf32':
        v, u = round_tf32(v), round_tf32(u)
    m = np.einsum('tpc,pck->tpk', v, u).reshape(n, 7, 7, 4, 4, k)
    yt = np.einsum('ai,ntuijk,bj->ntuabk', AT_2, m, AT_2)            # [n][ty][tx][a][b][k]
    y = yt.transpose(0, 1, 3, 2, 4, 5).reshape(n, 14, 14, k)
    y = y * np.asarray(scale, np.float64) + np.asarray(shift, np.float64)
    if relu:
        y = np.maximum(y, 0)
    return y.astype(np.float32)


def round_tf32(a):
    """Round float values to TF32 precision (cvt.rna.tf32.f32: nearest, ties away from zero)."""
    a32 = np.ascontiguousarray(a, np.float32)
    bits = a32.view(np.uint32).astype(np.uint64)
    bits = (bits + 0x1000) & 0xFFFFE000
    return bits.astype(np.uint32).view(np.float32).reshape(a32.shape).astype(np.float64)


def reference_pipeline_f4x4(x_frame, u36, scale, shift):
    """Index-for-index float32 emulation of the reference's three kernels for ONE image:
      kernel_*_winograd_BtdB   (Kernel128_winograd.cu:28-120): 4x4 grid of 6x6 tiles at stride 4 over a frame that is
                                allocated twice as large and zeroed (:236,242) so tile rows 16,17 read zeros and tile
                                columns 16,17 wrap into the next row;
      kernel_*_OuterProduct_*  (:186-213): ip[xi][tile][k] = sum_c t_input[xi][tile][c] * U[xi][c][k];
      kernel_*_winograd_AtIA   (:123-183): Y = A^T M A, relu(scale*Y + bias), crop to 14x14, write at (+1,+1) of a
                                zeroed 16x16 frame.
    x_frame [16,16,C] float32, u36 [36,C,K] float32 (weight_winograd_*.bin). Returns the padded [16,16,K] frame."""
    x_frame = np.asarray(x_frame, np.float32)
    c = x_frame.shape[2]
    k = u36.shape[2]
    flat = np.zeros(2 * 16 * 16 * c, np.float32)
    flat[:16 * 16 * c] = x_frame.reshape(-1)
    bt = BT_4.astype(np.float32)
    at = AT_4.astype(np.float32)
    t_input = np.zeros((36, 16, c), np.float32)
    for bx in range(4):
        for by in range(4):
            d = np.empty((6, 6, c), np.float32)
            for i in range(6):
                for j in range(6):
                    start = ((bx * 4 + i) * 16 + (by * 4 + j)) * c      # c_glb_start + i*stride_r, :31,38
                    d[i, j] = flat[start:start + c]
            btd = np.einsum('iy,yxc->ixc', bt, d).astype(np.float32)
            v = np.einsum('ixc,jx->ijc', btd, bt).astype(np.float32)
            # pOutputs[(Iny1 + i*6)*2048 + tile*128 + c] = BTdB[i] with BTdB[i] = row i, column Iny1 (:118)
            t_input[:, bx * 4 + by, :] = v.reshape(36, c)
    ip = np.einsum('ptc,pck->ptk', t_input, np.asarray(u36, np.float32)).astype(np.float32)
    out = np.zeros((16, 16, k), np.float32)
    for tx in range(4):
        for ty in range(4):
            m = ip[:, tx * 4 + ty, :].reshape(6, 6, k)                  # input[c_input] with c_input = Inx*6+Iny (:125,131)
            yy = np.einsum('ai,ijk,bj->abk', at, m, at).astype(np.float32)
            for a in range(4):
                for b in range(4):
                    oy, ox = tx * 4 + a, ty * 4 + b
                    if oy < 14 and ox < 14:                             # crop (:155,171,177)
                        o = scale * yy[a, b] + shift
                        out[oy + 1, ox + 1] = np.maximum(o, 0)
    return out


# ------------------------------------------------------------------------------------------------------------ checkers
def output_checker(a, b, length, channel, shift):
    """util.c:46-63: A is a (len+2*shift)^2 frame, B dense; returns (max_error, error_cnt) for |a-b| > 1e-5."""
    a = np.asarray(a, np.float32).reshape(length + 2 * shift, length + 2 * shift, channel)
    b = np.asarray(b, np.float32).reshape(length, length, channel)
    diff = np.abs(a[shift:shift + length, shift:shift + length, :] - b)
    return float(diff.max()), int((diff > 1e-5).sum())


def rel_err(out, gold):
    """The parity metric of BASELINE.md section 4: max|out - gold| / max|gold| over the layer."""
    out = np.asarray(out, np.float64)
    gold = np.asarray(gold, np.float64)
    return float(np.abs(out - gold).max() / np.abs(gold).max())


# ------------------------------------------------------------------------------------------- fp32 BLAS form (timed leg)
def conv3x3_bn_relu_fp32(x, w, scale, shift, relu=True):
    """The 'NumPy FP32 golden' BASELINE.md names as CPU baseline: 9 shifted sgemms + folded BN + ReLU, all float32."""
    return conv3x3_bn_relu(np.asarray(x, np.float32), np.asarray(w, np.float32), np.asarray(scale, np.float32),
                           np.asarray(shift, np.float32), relu, acc=np.float32)


def conv1x1_bn_fp32(x, w, scale, shift, relu):
    return conv1x1_bn(np.asarray(x, np.float32), np.asarray(w, np.float32), np.asarray(scale, np.float32),
                      np.asarray(shift, np.float32), relu, acc=np.float32)
