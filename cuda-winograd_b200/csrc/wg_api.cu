// C-ABI of libwinograd_b200.so (include/winograd_b200.h): layer handles, one-time filter packing, the single-launch
// hot path, and the host-buffer end-to-end call. No cuDNN, no cuBLAS, no CPU fallback: without an sm_100 device every
// create() returns WG_ERR_NODEVICE.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <cmath>

#include "wg_internal.h"

namespace wg {

static thread_local char g_last_cuda_error[256] = "";
static std::atomic<long long> g_launches{0};
static int g_max_ctas = 0;  // 0 = number of SMs
// 96 = V-in-TMEM kernel with the whole inverse transform folded into the MMAs (default); 48 = V-in-TMEM, half fold;
// 64 / 32 = the shared-memory-operand kernels (half fold / one accumulator per Winograd point)
static int g_wino_kn = -1;
static int wino_kn() {
  if (g_wino_kn < 0) {
    const char* e = getenv("WG_WINO_KN");  // same meaning as wg_set_wino_kn()
    const int v = e ? atoi(e) : 96;
    g_wino_kn = (v == 32 || v == 64 || v == 48) ? v : 96;
  }
  return g_wino_kn;
}
static bool kn_tm(int kn) { return kn == 48 || kn == 96; }  // V-in-TMEM kernels

static int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s", what, cudaGetErrorString(e));
  return WG_ERR_CUDA;
}
#define WG_CUDA(call)                                 \
  do {                                                \
    cudaError_t e_ = (call);                          \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

CUtensorMapL2promotion l2_promotion() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WG_L2_PROMO");
    v = e ? atoi(e) : 128;
  }
  switch (v) {
    case 0: return CU_TENSOR_MAP_L2_PROMOTION_NONE;
    case 64: return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
    case 256: return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    default: return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  }
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WG_PDL");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

}  // namespace wg

struct wg_layer {
  int kind;  // 0 = 3x3 Winograd, 1 = 1x1 GEMM
  int cin, cout, relu, dtype, device;
  int tile_n;  // 3x3: cout slice KN; 1x1: BN
  int tm_db;   // 3x3 TM kernel: 1 = double-buffered V, 32-wide slices
  int tm16_ff; // 3x3 bf16/fp16: d_filter_tm16 is the full-fold kernel's image (96-wide slices)
  int ff_cg2;  // 3x3 full-fold kernel: CTA-pair variant (filter image split in cout halves)
  float* d_filter_n64;       // 3x3 full-fold kernel: second image with all slices 64 wide (one-wave launches), or null
  float* d_filter_tm16_n64;  // the same for the 16-bit operand image
  int num_sms;
  float* d_filter;  // packed filter image (U or swizzled W^T)
  float* d_filter_tm16;   // 3x3 bf16/fp16 only: U in the 16-bit image of the V-in-TMEM throughput kernel (48/32 slices)
  float* d_filter_small;  // 3x3 TF32 only: U in the plain KN=32 image the small-batch kernel reads (may alias d_filter)
  float* d_scale;
  float* d_shift;
  // tensor-map cache for the last (x, N) seen
  const float* tmap_x;
  int tmap_n;
  CUtensorMap tmap;
  const float* tmap_tm_x;     // 3x3 bf16/fp16: tensor map of the V-in-TMEM kernel (32-byte swizzle)
  int tmap_tm_n;
  CUtensorMap tmap_tm;
  const float* tmap_small_x;  // 3x3 small-batch kernel: same view, 26-row box
  int tmap_small_n;
  CUtensorMap tmap_small;
  const float* tmap_y_ptr;  // 1x1 only: output tensor map for the TMA-store epilogue
  int tmap_y_n;
  CUtensorMap tmap_out;
  // staging for wg_run_host
  float* d_x;
  float* d_y;
  size_t d_x_bytes, d_y_bytes;
  cudaStream_t stream;
  // wg_run_host pipeline: copy-in / compute / copy-out streams and per-chunk events (created on first use)
  cudaStream_t s_h2d, s_d2h;
  cudaEvent_t ev_in[64], ev_done[64];
  int n_events;
};

using namespace wg;

static int check_device(int device, int* num_sms) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return WG_ERR_NODEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return WG_ERR_NODEVICE;
  if (prop.major != 10) return WG_ERR_NODEVICE;  // tcgen05 / TMEM: sm_100 family only
  *num_sms = prop.multiProcessorCount;
  return WG_OK;
}

static int create_common(wg_layer_t** out, int kind, int cin, int cout, const float* w, size_t w_elems,
                         const float* scale, const float* shift, int relu, wg_dtype_t dtype, int device) {
  if (!out || !w || !scale || !shift) return WG_ERR_ARG;
  if (dtype != WG_TF32 && dtype != WG_BF16 && dtype != WG_FP16) return WG_ERR_ARG;
  // bf16 operands: 3x3 only (the 1x1 activation operand goes HBM -> TMA -> MMA untouched, there is nothing to convert it)
  if (dtype != WG_TF32 && (kind != 0 || cin % 16 != 0 || cout % 64 != 0)) return WG_ERR_ARG;
  int num_sms = 0;
  int rc = check_device(device, &num_sms);
  if (rc != WG_OK) return rc;
  WG_CUDA(cudaSetDevice(device));

  wg_layer* L = static_cast<wg_layer*>(calloc(1, sizeof(wg_layer)));
  if (!L) return WG_ERR_NOMEM;
  L->kind = kind;
  L->cin = cin;
  L->cout = cout;
  L->relu = relu ? 1 : 0;
  L->dtype = dtype;
  L->device = device;
  L->num_sms = num_sms;

  float* d_w = nullptr;
  size_t filter_elems = 0;
  if (kind == 0) {
    if (dtype != WG_TF32) L->tile_n = 64;
    else if (kn_tm(wino_kn())) L->tile_n = wino_kn();
    else L->tile_n = (wino_kn() == 32 || cout % 64 != 0) ? 32 : 64;
    filter_elems = (size_t)16 * cin * cout;
  } else {
    L->tile_n = (cout % 256 == 0) ? 256 : 128;
    filter_elems = (size_t)cin * cout;
  }
  cudaError_t e;
#define WG_TRY(call)                \
  if ((e = (call)) != cudaSuccess) { \
    cuda_fail(e, #call);            \
    wg_destroy(L);                  \
    if (d_w) cudaFree(d_w);         \
    return WG_ERR_CUDA;             \
  }
  WG_TRY(cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking));
  WG_TRY(cudaMalloc(&d_w, w_elems * sizeof(float)));
  WG_TRY(cudaMalloc(&L->d_filter, filter_elems * sizeof(float)));
  WG_TRY(cudaMalloc(&L->d_scale, cout * sizeof(float)));
  WG_TRY(cudaMalloc(&L->d_shift, cout * sizeof(float)));
  WG_TRY(cudaMemcpyAsync(d_w, w, w_elems * sizeof(float), cudaMemcpyHostToDevice, L->stream));
  WG_TRY(cudaMemcpyAsync(L->d_scale, scale, cout * sizeof(float), cudaMemcpyHostToDevice, L->stream));
  WG_TRY(cudaMemcpyAsync(L->d_shift, shift, cout * sizeof(float), cudaMemcpyHostToDevice, L->stream));
  if (kind == 0) L->ff_cg2 = wino_ff_cg2();
  if (kind == 0 && L->tile_n == 96) {
    rc = filter_transform_ff_launch(d_w, L->d_filter, cin, cout, 0, L->ff_cg2, 0, L->stream);
    if (rc == WG_OK && !L->ff_cg2 && wino_ff_has_narrow(cout)) {
      WG_TRY(cudaMalloc(&L->d_filter_n64, filter_elems * sizeof(float)));
      rc = filter_transform_ff_launch(d_w, L->d_filter_n64, cin, cout, 0, 0, 1, L->stream);
      g_launches++;
    }
  } else if (kind == 0 && L->tile_n == 48) {
    L->tm_db = wino_tm_choose_db(cin, cout);
    rc = filter_transform_tm_launch(d_w, L->d_filter, cin, cout, L->tm_db, 0, L->stream);
  }
  else if (kind == 0) rc = filter_transform_launch(d_w, L->d_filter, cin, cout, L->tile_n, (int)dtype, L->stream);
  else rc = weight_pack_launch(d_w, L->d_filter, cin, cout, L->tile_n, L->stream);
  g_launches++;
  if (rc == WG_OK && kind == 0 && dtype != WG_TF32 && kn_tm(wino_kn())) {
    // 16-bit operands: the throughput kernel keeps V packed in TMEM (its own filter image); small batches stay on the
    // split-C variant of the shared-memory-operand kernel (d_filter)
    WG_TRY(cudaMalloc(&L->d_filter_tm16, filter_elems * sizeof(uint16_t)));
    L->tm16_ff = wino_kn() == 96;
    rc = L->tm16_ff ? filter_transform_ff_launch(d_w, L->d_filter_tm16, cin, cout, dtype == WG_FP16 ? 2 : 1, L->ff_cg2,
                                                 0, L->stream)
                    : filter_transform_tm_launch(d_w, L->d_filter_tm16, cin, cout, 0, dtype == WG_FP16 ? 2 : 1,
                                                 L->stream);
    g_launches++;
    if (rc == WG_OK && L->tm16_ff && !L->ff_cg2 && wino_ff_has_narrow(cout)) {
      WG_TRY(cudaMalloc(&L->d_filter_tm16_n64, filter_elems * sizeof(uint16_t)));
      rc = filter_transform_ff_launch(d_w, L->d_filter_tm16_n64, cin, cout, dtype == WG_FP16 ? 2 : 1, 0, 1, L->stream);
      g_launches++;
    }
  }
  if (rc == WG_OK && kind == 0 && dtype == WG_TF32) {
    if (L->tile_n == 32) {
      L->d_filter_small = L->d_filter;
    } else {
      WG_TRY(cudaMalloc(&L->d_filter_small, filter_elems * sizeof(float)));
      rc = filter_transform_launch(d_w, L->d_filter_small, cin, cout, 32, 0, L->stream);
      g_launches++;
    }
  }
  if (rc != WG_OK) {
    cuda_fail(cudaGetLastError(), "filter pack launch");
    wg_destroy(L);
    cudaFree(d_w);
    return rc;
  }
  WG_TRY(cudaStreamSynchronize(L->stream));
#undef WG_TRY
  cudaFree(d_w);
  *out = L;
  return WG_OK;
}

extern "C" {

int wg_conv3x3_create(wg_layer_t** out, int C, int K, const float* w_kcrs, const float* scale, const float* shift,
                      int relu, wg_dtype_t dtype, int device) {
  if (C <= 0 || K <= 0 || C % 8 != 0 || K % 32 != 0) return WG_ERR_ARG;
  return create_common(out, 0, C, K, w_kcrs, (size_t)K * C * 9, scale, shift, relu, dtype, device);
}

int wg_conv1x1_create(wg_layer_t** out, int Cin, int Cout, const float* w_cin_cout, const float* scale,
                      const float* shift, int relu, wg_dtype_t dtype, int device) {
  if (Cin <= 0 || Cout <= 0 || Cin % 32 != 0 || Cout % 128 != 0) return WG_ERR_ARG;
  return create_common(out, 1, Cin, Cout, w_cin_cout, (size_t)Cin * Cout, scale, shift, relu, dtype, device);
}

int wg_run(wg_layer_t* L, const float* x, float* y, int N, int out_padded, void* cuda_stream) {
  if (!L || !x || !y || N <= 0) return WG_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return WG_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  // out_padded is a flag word: WG_OUT_PADDED (1) = zero-bordered frame, WG_OUT_MULTICAST (2) = y is an NVLS multicast
  // address (stores go out as multimem.st and land in every GPU's buffer: fused conv + all-gather of the output).
  if (out_padded & ~3) return WG_ERR_ARG;
  const int out_flags = out_padded & 3;
  if ((out_flags & 2) && !(L->kind == 0 && L->dtype == WG_TF32 && kn_tm(L->tile_n))) return WG_ERR_ARG;
  out_padded &= 1;
  int cur = -1;
  WG_CUDA(cudaGetDevice(&cur));
  if (cur != L->device) WG_CUDA(cudaSetDevice(L->device));
  const int max_ctas = g_max_ctas > 0 ? g_max_ctas : L->num_sms;
  if (L->kind == 0 && L->d_filter_small != nullptr) {
    // small batches: the latency variant (one 64-tile x 32-cout item per cluster, split-C), see wino_small_kernel.cu
    const int cs = wino_small_cs(N, L->cin, L->cout, max_ctas);
    if (cs > 0) {
      if (L->tmap_small_x != x || L->tmap_small_n != N) {
        int rc = wino_small_make_tmap(&L->tmap_small, x, N, L->cin);
        if (rc != WG_OK) return rc;
        L->tmap_small_x = x;
        L->tmap_small_n = N;
      }
      int rc = wino_small_launch(L->tmap_small, L->d_filter_small, L->d_scale, L->d_shift, y, N, L->cin, L->cout,
                                 L->relu, out_flags, cs, stream);
      g_launches++;
      if (rc == WG_ERR_CUDA) cuda_fail(cudaGetLastError(), "kernel launch");
      return rc;
    }
  }
  if (L->kind == 0 && L->d_filter_tm16 != nullptr && !(out_flags & 2)) {
    // bf16 / fp16 operands, throughput-sized batch (the split-C latency mode of the other kernel takes the small ones)
    const long long items64 = (long long)((N * 49 + 63) / 64) * (L->cout / 64);
    if (items64 * 4 > max_ctas) {
      if (L->tmap_tm_x != x || L->tmap_tm_n != N) {
        int rc = L->tm16_ff ? wino_ff_make_tmap(&L->tmap_tm, x, N, L->cin) : wino_tm_make_tmap(&L->tmap_tm, x, N, L->cin, 1);
        if (rc != WG_OK) return rc;
        L->tmap_tm_x = x;
        L->tmap_tm_n = N;
      }
      const int op16 = L->dtype == WG_FP16 ? 2 : 1;
      int rc = L->tm16_ff ? wino_ff_launch(L->tmap_tm, x, L->d_filter_tm16, L->d_filter_tm16_n64, L->d_scale, L->d_shift, y, N, L->cin, L->cout,
                                           op16, L->ff_cg2, L->relu, out_flags, max_ctas, stream)
                          : wino_tm_launch(L->tmap_tm, L->d_filter_tm16, L->d_scale, L->d_shift, y, N, L->cin, L->cout,
                                           0, op16, L->relu, out_flags, max_ctas, stream);
      g_launches++;
      if (rc == WG_ERR_CUDA) cuda_fail(cudaGetLastError(), "kernel launch");
      return rc;
    }
  }
  if (L->tmap_x != x || L->tmap_n != N) {
    int rc = L->kind == 1 ? one_make_tmap(&L->tmap, x, (long long)N * 196, L->cin)
             : L->tile_n == 96 ? wino_ff_make_tmap(&L->tmap, x, N, L->cin)
             : L->tile_n == 48 ? wino_tm_make_tmap(&L->tmap, x, N, L->cin, wino_tm_cls(L->cout, L->tm_db))
                               : wino_make_tmap(&L->tmap, x, N, L->cin);
    if (rc != WG_OK) return rc;
    L->tmap_x = x;
    L->tmap_n = N;
  }
  if (L->kind == 1 && (L->tmap_y_ptr != y || L->tmap_y_n != N)) {
    int rc = one_make_tmap_out(&L->tmap_out, y, (long long)N * 196, L->cout);
    if (rc != WG_OK) return rc;
    L->tmap_y_ptr = y;
    L->tmap_y_n = N;
  }
  int rc;
  if (L->kind == 0 && L->tile_n == 96)
    rc = wino_ff_launch(L->tmap, x, L->d_filter, L->d_filter_n64, L->d_scale, L->d_shift, y, N, L->cin, L->cout, 0, L->ff_cg2, L->relu,
                        out_flags, max_ctas, stream);
  else if (L->kind == 0 && L->tile_n == 48)
    rc = wino_tm_launch(L->tmap, L->d_filter, L->d_scale, L->d_shift, y, N, L->cin, L->cout, L->tm_db, 0, L->relu,
                        out_flags, max_ctas, stream);
  else if (L->kind == 0)
    rc = wino_launch(L->tmap, L->d_filter, L->d_scale, L->d_shift, y, N, L->cin, L->cout, L->tile_n,
                     L->dtype, L->relu, out_padded ? 1 : 0, max_ctas, stream);
  else
    rc = one_launch(L->tmap, L->tmap_out, L->d_filter, L->d_scale, L->d_shift, y, out_padded ? 1 : 0,
                    (long long)N * 196, L->cin, L->cout, L->tile_n, L->relu, max_ctas, stream);
  g_launches++;
  if (rc == WG_ERR_CUDA) cuda_fail(cudaGetLastError(), "kernel launch");
  return rc;
}

int wg_run_host(wg_layer_t* L, const float* x_host, float* y_host, int N, int out_padded) {
  if (!L || !x_host || !y_host || N <= 0) return WG_ERR_ARG;
  int cur = -1;
  WG_CUDA(cudaGetDevice(&cur));
  if (cur != L->device) WG_CUDA(cudaSetDevice(L->device));
  const size_t in_px = L->kind == 0 ? 256 : 196;
  const size_t out_px = out_padded ? 256 : 196;
  const size_t xb = (size_t)N * in_px * L->cin * sizeof(float);
  const size_t yb = (size_t)N * out_px * L->cout * sizeof(float);
  if (L->d_x_bytes < xb) {
    if (L->d_x) cudaFree(L->d_x);
    L->d_x = nullptr;
    L->d_x_bytes = 0;
    L->tmap_x = nullptr;
    WG_CUDA(cudaMalloc(&L->d_x, xb));
    L->d_x_bytes = xb;
  }
  if (L->d_y_bytes < yb) {
    if (L->d_y) cudaFree(L->d_y);
    L->d_y = nullptr;
    L->d_y_bytes = 0;
    L->tmap_y_ptr = nullptr;
    WG_CUDA(cudaMalloc(&L->d_y, yb));
    L->d_y_bytes = yb;
  }
  // Chunked three-stream pipeline: H2D of chunk c+1, the kernel on chunk c and D2H of chunk c-1 overlap (PCIe is
  // full duplex), so a large batch costs ~max(copy-in, copy-out) instead of their sum plus the kernel. Host buffers
  // should be pinned for the copies to be asynchronous; pageable memory still works, just without the overlap.
  if (!L->s_h2d) {
    WG_CUDA(cudaStreamCreateWithFlags(&L->s_h2d, cudaStreamNonBlocking));
    WG_CUDA(cudaStreamCreateWithFlags(&L->s_d2h, cudaStreamNonBlocking));
  }
  static int chunk_env = -1;  // WG_HOST_CHUNK=<images per chunk> (experiments); default 64 (measured: 64 > 32 > 16 on PCIe Gen5; per-chunk event and launch costs outweigh the shorter pipeline tail)
  if (chunk_env < 0) {
    const char* e = getenv("WG_HOST_CHUNK");
    chunk_env = e ? atoi(e) : 0;
  }
  int chunk = chunk_env > 0 ? chunk_env : 64;
  if ((N + chunk - 1) / chunk > 56) chunk = (N + 55) / 56;
  // Chunk schedule: full chunks, then a tapering tail (halving down to 16 images). The copy-in stream is the critical
  // path from t = 0 whatever the chunking; what is exposed at the end is the LAST chunk's kernel + copy-out, so the
  // last chunks are small (256 images: 64, 64, 64, 32, 16, 16). WG_HOST_TAPER=0 keeps equal chunks. Measured (3x3
  // 256->256, 256 images per call, PCIe Gen5): 152-159 -> 162-164 k images/s; with the taper, chunk = 32 / 48 / 64 / 96 /
  // 128 give 157 / 161 / 163 / 157 / 151 k.
  static int taper_env = -1;
  if (taper_env < 0) {
    const char* e = getenv("WG_HOST_TAPER");
    taper_env = e ? atoi(e) : 1;
  }
  int sizes[64];
  int n_chunks = 0;
  for (int rem = N; rem > 0;) {
    int c = rem < chunk ? rem : chunk;
    if (taper_env && rem <= 2 * chunk && rem > 16) {
      c = rem / 2;
      if (c < 16) c = 16;
      if (c > chunk) c = chunk;
    }
    sizes[n_chunks++] = c;
    rem -= c;
  }
  while (L->n_events < n_chunks) {
    WG_CUDA(cudaEventCreateWithFlags(&L->ev_in[L->n_events], cudaEventDisableTiming));
    WG_CUDA(cudaEventCreateWithFlags(&L->ev_done[L->n_events], cudaEventDisableTiming));
    L->n_events++;
  }
  const size_t x_img = in_px * L->cin, y_img = out_px * L->cout;  // floats per image
  int n0 = 0;
  for (int c = 0; c < n_chunks; ++c) {
    const int nc = sizes[c];
    WG_CUDA(cudaMemcpyAsync(L->d_x + (size_t)n0 * x_img, x_host + (size_t)n0 * x_img, (size_t)nc * x_img * sizeof(float),
                            cudaMemcpyHostToDevice, L->s_h2d));
    WG_CUDA(cudaEventRecord(L->ev_in[c], L->s_h2d));
    WG_CUDA(cudaStreamWaitEvent(L->stream, L->ev_in[c], 0));
    int rc = wg_run(L, L->d_x + (size_t)n0 * x_img, L->d_y + (size_t)n0 * y_img, nc, out_padded, L->stream);
    if (rc != WG_OK) return rc;
    WG_CUDA(cudaEventRecord(L->ev_done[c], L->stream));
    WG_CUDA(cudaStreamWaitEvent(L->s_d2h, L->ev_done[c], 0));
    WG_CUDA(cudaMemcpyAsync(y_host + (size_t)n0 * y_img, L->d_y + (size_t)n0 * y_img, (size_t)nc * y_img * sizeof(float),
                            cudaMemcpyDeviceToHost, L->s_d2h));
    n0 += nc;
  }
  WG_CUDA(cudaStreamSynchronize(L->s_d2h));
  WG_CUDA(cudaStreamSynchronize(L->stream));
  return WG_OK;
}

int wg_destroy(wg_layer_t* L) {
  if (!L) return WG_ERR_ARG;
  if (L->d_filter_small && L->d_filter_small != L->d_filter) cudaFree(L->d_filter_small);
  if (L->d_filter_tm16) cudaFree(L->d_filter_tm16);
  if (L->d_filter_n64) cudaFree(L->d_filter_n64);
  if (L->d_filter_tm16_n64) cudaFree(L->d_filter_tm16_n64);
  if (L->d_filter) cudaFree(L->d_filter);
  if (L->d_scale) cudaFree(L->d_scale);
  if (L->d_shift) cudaFree(L->d_shift);
  if (L->d_x) cudaFree(L->d_x);
  if (L->d_y) cudaFree(L->d_y);
  for (int i = 0; i < L->n_events; ++i) {
    cudaEventDestroy(L->ev_in[i]);
    cudaEventDestroy(L->ev_done[i]);
  }
  if (L->s_h2d) cudaStreamDestroy(L->s_h2d);
  if (L->s_d2h) cudaStreamDestroy(L->s_d2h);
  if (L->stream) cudaStreamDestroy(L->stream);
  free(L);
  return WG_OK;
}

int wg_layer_info(const wg_layer_t* L, int* kind, int* cin, int* cout, int* relu) {
  if (!L) return WG_ERR_ARG;
  if (kind) *kind = L->kind;
  if (cin) *cin = L->cin;
  if (cout) *cout = L->cout;
  if (relu) *relu = L->relu;
  return WG_OK;
}

long long wg_launch_count(void) { return g_launches.load(); }

const char* wg_strerror(int status) {
  switch (status) {
    case WG_OK: return "ok";
    case WG_ERR_ARG: return "invalid argument";
    case WG_ERR_CUDA: return "CUDA error";
    case WG_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable";
    case WG_ERR_TMAP: return "tensor map encoding failed";
    case WG_ERR_NOMEM: return "out of memory";
    case WG_ERR_NODEVICE: return "no sm_100 (B200) device; this library has no CPU fallback";
    case WG_ERR_IO: return "data file missing or short";
    default: return "unknown status";
  }
}

const char* wg_last_cuda_error(void) { return g_last_cuda_error; }

int wg_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int n = 0;
  for (int i = 0; i < count; ++i) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) ++n;
  }
  return n;
}

void wg_fold_bn(int K, const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                float* scale_out, float* shift_out) {
  for (int k = 0; k < K; ++k) {
    const float sd = sqrtf(var[k] + eps);  // same operation order as the numpy expressions
    scale_out[k] = gamma[k] / sd;
    shift_out[k] = beta[k] - (gamma[k] * mean[k]) / sd;
  }
}

void wg_set_max_ctas(int max_ctas) { g_max_ctas = max_ctas; }
void wg_set_wino_kn(int kn) { g_wino_kn = (kn == 32 || kn == 64 || kn == 48) ? kn : 96; }

}  // extern "C"
