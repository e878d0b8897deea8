/*
 * winograd_b200.h -- tensor-level C ABI of libwinograd_b200.so (B200 / sm_100a).
 *
 * The reference (bssrdf/CUDA-Winograd) exposes its one hot path -- fused conv + inference-BatchNorm + ReLU -- only
 * as six zero-argument entry points that read data/<name>.bin files (Kernel128_winograd.h:20, Kernel256_winograd.h:20,
 * Kernel128_one.h:18-19, Kernel256_one.h:18-19; re-declared for this library in include/Kernel*.h). Those symbols
 * cannot express a batch, a device or caller-owned buffers, so they are thin wrappers over the functions below,
 * which are what a host-language binding (ctypes / cgo / JNI) should bind. Plain pointers and sizes only.
 *
 * Layouts (identical to the files written by the reference's data_generator.py):
 *   3x3 input   x  [N][16][16][C]   fp32 NHWC, the 1-pixel border is part of the data (data_generator.py:49-53,
 *                                   Kernel128_winograd.cu:26-31)
 *   3x3 filter  w  [K][C][3][3]     fp32 "NCHW"/KCRS (data_generator.py:55-61, Kernel128_winograd.cu:343)
 *   3x3 output  y  [N][14][14][K]   fp32 NHWC, or -- out_padded != 0 -- the reference's own [N][16][16][K] frame with
 *                                   a zero border and the result at (+1,+1) (Kernel128_winograd.cu:163,243)
 *   1x1 input   x  [N][196][Cin]    fp32 (Kernel128_one.cu:33-34)
 *   1x1 weight  w  [Cin][Cout]      fp32 (Kernel128_one.cu:41-48, :259)
 *   1x1 output  y  [N][196][Cout]   fp32
 *   scale/shift    [K]              folded BN: scale = gamma/sqrt(var+eps), shift = beta - gamma*mean/sqrt(var+eps)
 *                                   (data_generator.py:41-47,107-113); y = act(scale*conv + shift)
 *
 * Every function returns WG_OK (0) or a negative wg_status; nothing here calls exit().
 */
#ifndef WINOGRAD_B200_H_
#define WINOGRAD_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wg_layer wg_layer_t; /* opaque: device copies of the packed filter, folded BN, cached tensor maps */

/* operand type of the tensor-core GEMMs (inputs, outputs and accumulation are always fp32):
 *   WG_TF32  10-bit mantissa, fp32 range                      tolerance 1e-3   all layers (default)
 *   WG_BF16   7-bit mantissa, fp32 range                      tolerance 1e-2   3x3 layers with C % 16 == 0 and
 *                                                                              K % 64 == 0; every 1x1 layer
 *   WG_FP16  10-bit mantissa, |V| = |sum of 4 inputs| < 65504  tolerance 1e-3   3x3 layers, same constraints;
 *            TF32-class accuracy at bf16 speed for bounded activations (e.g. post-BN/ReLU feature maps) */
typedef enum { WG_TF32 = 0, WG_BF16 = 1, WG_FP16 = 2 } wg_dtype_t;

typedef enum {
  WG_OK = 0,
  WG_ERR_ARG = -1,    /* bad shape / null pointer / unsupported channel count */
  WG_ERR_CUDA = -2,   /* a CUDA runtime call or kernel launch failed (see wg_last_cuda_error) */
  WG_ERR_DRIVER = -3, /* cuTensorMapEncodeTiled could not be resolved */
  WG_ERR_TMAP = -4,   /* tensor-map encoding rejected */
  WG_ERR_NOMEM = -5,
  WG_ERR_NODEVICE = -6, /* no sm_100 device: the library never falls back to a CPU path */
  WG_ERR_IO = -7        /* a layer blob is missing, short, corrupt or of another version */
} wg_status;

/* One-time per layer: copies weights / BN to `device`, runs the filter transform U = G g G^T (F(2x2,3x3)) or the
 * 1x1 weight pack on the GPU, rounds the filter operand to TF32 (round-to-nearest).
 * Replaces the offline weight_generator (data_generator.py:63-78) and the per-call cudaMalloc/cudaMemcpy block of
 * kernel_128() (Kernel128_winograd.cu:236-256). All pointers are HOST pointers. Shape constraints (else WG_ERR_ARG):
 * 3x3: C % 8 == 0 and K % 32 == 0 (16-bit operand types: C % 16 == 0 and K % 64 == 0); 1x1: Cin % 32 == 0 and
 * Cout % 128 == 0 (WG_FP16 does not exist for 1x1 layers). */
int wg_conv3x3_create(wg_layer_t** out, int C, int K, const float* w_kcrs, const float* scale, const float* shift,
                      int relu, wg_dtype_t dtype, int device);
int wg_conv1x1_create(wg_layer_t** out, int Cin, int Cout, const float* w_cin_cout, const float* scale,
                      const float* shift, int relu, wg_dtype_t dtype, int device);

/* The same layers on other feature-map sizes (the reference hard-codes 14x14 outputs in 16x16 frames,
 * Kernel128_winograd.cu:26-31,263-265; wg_conv3x3_create == wg_conv3x3_create_hw with H = W = 14). H x W is the OUTPUT
 * map, 3 <= H, W and W <= 104 (e.g. the ResNet-50 stages 56x56x64, 28x28x128, 14x14x256, 7x7x512). Layouts:
 *   3x3 input   x [N][Hf][Wf][C]   the frame INCLUDING its border, Hf = H + 2 and Wf = W + 2 for even sizes; an odd H (W)
 *                                  has Hf = H + 3 (Wf = W + 3): one extra trailing row (column) that is never read into a
 *                                  valid output, so that frame rows pair up (F(2x2,3x3) tiles cover 2x2 outputs). The data
 *                                  sit at rows 1..H, columns 1..W; wg_layer_geometry() reports Hf, Wf.
 *   3x3 output  y [N][H][W][K], or with WG_OUT_PADDED the [N][Hf][Wf][K] frame (zero border, result at (+1,+1))
 *   1x1 input / output   [N][H*W][Cin] / [N][H*W][Cout]; WG_OUT_PADDED writes the [N][Hf][Wf][Cout] frame a following
 *                        3x3 layer of the same map size reads.
 * Stride-2 / downsampling layers are out of scope (no Winograd form). */
int wg_conv3x3_create_hw(wg_layer_t** out, int C, int K, int H, int W, const float* w_kcrs, const float* scale,
                         const float* shift, int relu, wg_dtype_t dtype, int device);
int wg_conv1x1_create_hw(wg_layer_t** out, int Cin, int Cout, int H, int W, const float* w_cin_cout,
                         const float* scale, const float* shift, int relu, wg_dtype_t dtype, int device);
int wg_layer_geometry(const wg_layer_t* layer, int* H, int* W, int* frame_h, int* frame_w);
/* Frame size of an H x W map (host only, no GPU needed); WG_ERR_ARG for maps the kernels cannot tile (H or W < 3). */
int wg_frame_dims(int H, int W, int* frame_h, int* frame_w);
/* How the direct-convolution 3x3 kernels cut an H x W map into work items (host only, no GPU needed; what a layer of
 * that map size and operand type runs when its channel counts fit the kernel's blocking). Returns 1 and fills
 * out[0..8] = {rows per item R (0: whole images), bands per image, images per item G, pixels per item rounded up to 16
 * (the MMA N, <= 256), halo rows, TMA boxes per stage, rows per box, frame_h, frame_w}; 0 if the map does not fit the
 * kernel (the Winograd kernels run); WG_ERR_ARG (< 0) for a bad argument. */
int wg_direct_geometry(int H, int W, wg_dtype_t dtype, int* out9);

/* The hot path: ONE kernel launch (conv + BN + optional ReLU) on `cuda_stream` (a cudaStream_t, may be NULL),
 * asynchronous. x and y are DEVICE pointers on the layer's device, 16-byte aligned.
 * Replaces the timed region of kernel_128()/kernel_256() (three launches, Kernel128_winograd.cu:263-265) and of
 * kernel_*_1_in/out() (one launch, Kernel128_one.cu:98,316). For a 1x1 layer out_padded != 0 writes the same
 * zero-bordered [N][16][16][Cout] frame instead of [N][196][Cout], i.e. exactly the input a following 3x3 layer
 * reads, so bottleneck chains 1x1 -> 3x3 -> 1x1 need no padding pass in between.
 * out_flags is a flag word: WG_OUT_PADDED (1) as above; WG_OUT_MULTICAST (2, 3x3 TF32 layers): y_dev is an NVLS
 * multicast address (cuMulticast* / torch symmetric memory) and every output store is issued as multimem.st, i.e. the
 * NVSwitch replicates this GPU's shard into the buffers of all GPUs bound to the multicast object -- conv + BN + ReLU
 * fused with the all-gather of the output, no NCCL call (the caller still has to barrier across ranks before reading
 * the gathered tensor AND before the next call overwrites it). Other values / layer kinds with the multicast flag:
 * WG_ERR_ARG. Thread safety: concurrent wg_run calls on one layer from several host threads are safe (the tensor-map
 * cache is locked, the launch takes copies); the caller's current device is restored on return. */
#define WG_OUT_PADDED 1
#define WG_OUT_MULTICAST 2
#define WG_OUT_RELU_AFTER_ADD 4
#define WG_OUT_INTERIOR_ONLY 8 /* 1x1 layers with WG_OUT_PADDED: write the H x W interior of the frame only -- the caller
                                  guarantees that the border is already zero (e.g. a frame buffer zeroed once and reused
                                  by every call of a bottleneck block); saves the ~30 % extra stores of the edge pixels */
int wg_run(wg_layer_t* layer, const float* x_dev, float* y_dev, int N, int out_flags, void* cuda_stream);

/* The step that follows the reference's `_out` 1x1 layers in a ResNet bottleneck block (they stop before it:
 * Kernel128_one.cu:271-272, Kernel256_one.cu:273 write scale*acc + shift with no ReLU): the residual add, fused into
 * the same launch. 1x1 layers, dense output only: y = act2(act(scale * (x W) + shift) + residual), act = the layer's own
 * ReLU flag, act2 = ReLU iff out_flags has WG_OUT_RELU_AFTER_ADD. residual_dev: [N][196][Cout] fp32 on the layer's
 * device, 16-byte aligned; it may alias y_dev (each element is read before it is written by the same thread). */
int wg_run_residual(wg_layer_t* layer, const float* x_dev, const float* residual_dev, float* y_dev, int N,
                    int out_flags, void* cuda_stream);

/* Same, end to end with HOST buffers: H2D copy of x, wg_run, D2H copy of y, stream-synchronised on return.
 * Device staging buffers are cached in the layer (grown on demand). This is the call bench.py's `e2e` times. */
int wg_run_host(wg_layer_t* layer, const float* x_host, float* y_host, int N, int out_padded);

/* The chunk schedule wg_run_host uses for a batch of N images (host-only, no GPU needed): writes up to `cap` chunk
 * sizes and returns the number of chunks (<= 64 for every N; the sizes sum to N). */
int wg_host_chunk_schedule(int N, int* sizes, int cap);

/* Packed per-layer blob: everything create() produced -- the raw weights, the folded BN vectors and the kernels'
 * packed filter images (TF32 / bf16 U = G g G^T in shared-memory order, swizzled 1x1 weight tiles) -- behind a
 * versioned header with the layer's shape and a checksum. Replaces the 18 loose files of data_generator.py:55-78,
 * 116-127 as the deployment format (data/<name>.bin stay valid inputs to create()); loading skips the filter transform.
 * A blob whose images were packed for other kernel choices is re-packed from its raw weights. serialize(): buf == NULL
 * queries the size into *need. load/deserialize: WG_ERR_IO for a missing, short, corrupt or other-version blob. */
int wg_layer_serialize(const wg_layer_t* layer, void* buf, size_t cap, size_t* need);
int wg_layer_deserialize(wg_layer_t** out, const void* buf, size_t bytes, int device);
int wg_layer_save(const wg_layer_t* layer, const char* path);
int wg_layer_load(wg_layer_t** out, const char* path, int device);

int wg_destroy(wg_layer_t* layer);

/* Introspection used by the tests and the bench. */
int wg_layer_info(const wg_layer_t* layer, int* kind /*0 = 3x3, 1 = 1x1*/, int* cin, int* cout, int* relu);
long long wg_launch_count(void);        /* kernels launched by this library so far (this process) */
const char* wg_strerror(int status);
const char* wg_last_cuda_error(void);   /* cudaGetErrorString of the last failing CUDA call, "" if none */
int wg_device_count(void);              /* number of sm_100 devices visible; 0 => every create() fails loudly */

/* Folded BN exactly as data_generator.py:41-47 does it (fp32 arithmetic, eps = 1e-5 there). Host only. */
void wg_fold_bn(int K, const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                float* scale_out, float* shift_out);

/* Roofline denominator measured on the device itself: runs nothing but back-to-back tcgen05.mma (M=128, N=256, both
 * operands in shared memory) on every SM for ~1-3 ms and reports the dense tensor-pipe throughput in TFLOP/s for
 * dtype = WG_TF32 (kind::tf32, K=8) or WG_BF16 (kind::f16, K=16); *clk_per_mma (optional) = SM clocks per MMA at the
 * device's nominal clock. bench.py divides the kernels' achieved TFLOP/s by this number. */
int wg_measure_tensor_peak(int device, int dtype, double* tflops, double* clk_per_mma);

/* Benchmarking knob, process-wide (atomic): max CTAs per launch, 0 = the device's SM count (default). The product
 * library reads no environment variables; A/B knobs and superseded kernel generations live in the developer build
 * (make dev -> tools/libwinograd_b200_dev.so). */
void wg_set_max_ctas(int max_ctas);

#ifdef __cplusplus
}
#endif
#endif /* WINOGRAD_B200_H_ */
