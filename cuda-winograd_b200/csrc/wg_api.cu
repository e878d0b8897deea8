// C-ABI of libwinograd_b200.so (include/winograd_b200.h): layer handles, one-time filter packing, the single-launch
// hot path, the host-buffer end-to-end call and the packed per-layer blob. No cuDNN, no cuBLAS, no CPU fallback: without
// an sm_100 device every create() returns WG_ERR_NODEVICE.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <cmath>
#include <mutex>
#include <new>
#include <vector>

#include "wg_internal.h"

namespace wg {

static thread_local char g_last_cuda_error[256] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_max_ctas{0};  // 0 = number of SMs; process-wide benchmarking knob (wg_set_max_ctas)

// 3x3 TF32 kernel generation for layers created from now on. The product build has exactly one: 96 = V in tensor
// memory, whole inverse transform folded into the MMAs. The developer build also keeps 48 (V in TMEM, half fold) and
// 64 / 32 (both operands in shared memory), selectable with wg_dev_set_wino_kn() / WG_WINO_KN.
#ifdef WG_DEV_BUILD
static std::atomic<int> g_wino_kn{-1};
static int wino_kn() {
  int v = g_wino_kn.load();
  if (v < 0) {
    const char* e = dev_env("WG_WINO_KN");
    v = e ? atoi(e) : 96;
    v = (v == 32 || v == 64 || v == 48) ? v : 96;
    g_wino_kn.store(v);
  }
  return v;
}
#else
static int wino_kn() { return 96; }
#endif
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
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

CUtensorMapL2promotion l2_promotion() {
  static int v = -1;
  if (v < 0) {
    const char* e = dev_env("WG_L2_PROMO");
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
    const char* e = dev_env("WG_PDL");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

// Restores the caller's current device on every exit path (a wg_* call on a layer that lives on another GPU must not
// change the calling thread's device for its later CUDA calls).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  int enter(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) return WG_ERR_CUDA;
    if (prev != device) {
      if (cudaSetDevice(device) != cudaSuccess) return WG_ERR_CUDA;
      switched = true;
    }
    return WG_OK;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// Tensor maps are keyed on (base pointer, rows): a few entries per kind so that callers rotating buffers (bench.py,
// ping-pong chains) do not re-encode on every call. Looked up and filled under the layer's mutex; the launch uses a
// by-value copy (the kernels take the map as a __grid_constant__ parameter), so concurrent wg_run calls on one layer
// from several host threads are safe.
struct TmapCache {
  static constexpr int kWays = 4;
  struct Entry {
    const void* ptr = nullptr;
    long long n = -1;
    CUtensorMap map;
  } e[kWays];
  int next = 0;
  void clear() {
    for (auto& x : e) x.ptr = nullptr, x.n = -1;
  }
  template <class Make>
  int get(const void* ptr, long long n, CUtensorMap* out, Make make) {
    for (auto& x : e)
      if (x.ptr == ptr && x.n == n) {
        *out = x.map;
        return WG_OK;
      }
    Entry& slot = e[next];
    next = (next + 1) % kWays;
    slot.ptr = nullptr;
    int rc = make(&slot.map);
    if (rc != WG_OK) return rc;
    slot.ptr = ptr;
    slot.n = n;
    *out = slot.map;
    return WG_OK;
  }
};

}  // namespace wg

using namespace wg;

static constexpr int kNumImgs = 6;  // packed images per layer (device buffers and blob sections)

struct wg_layer {
  int kind = 0;  // 0 = 3x3 Winograd, 1 = 1x1 GEMM
  int cin = 0, cout = 0, relu = 0, dtype = 0, device = 0;
  int tile_n = 0;   // 3x3: cout slice KN; 1x1: BN
  int tm_db = 0;    // 3x3 TM kernel (developer build): 1 = double-buffered V, 32-wide slices
  int tm16_ff = 0;  // 3x3 bf16/fp16: d_filter_tm16 is the full-fold kernel's image (96-wide slices)
  int ff_cg2 = 0;   // 3x3 full-fold kernel: CTA-pair variant (developer build)
  int num_sms = 0;
  int H = 14, W = 14;      // output map (the reference: 14 x 14 everywhere)
  ff::Geo geo{};           // 3x3: tiles / frame / raw-plane geometry of the full-fold kernel
  OneGeo one_geo{14, 14, 16, 16, 0};  // 1x1: pixels per image and the padded frame of chain mode
  DirGeo dgeo{};           // 3x3 direct-convolution kernel on a map size other than 14x14
  bool dgen = false;
  // packed images (device). Sizes in bytes in img_bytes[], same order as the blob sections.
  float* d_filter = nullptr;           // packed filter image (U or swizzled W^T)
  float* d_filter_n64 = nullptr;       // 3x3 full-fold kernel: second image with all slices 64 wide, or null
  float* d_filter_tm16 = nullptr;      // 3x3 bf16/fp16: U in the 16-bit image of the V-in-TMEM throughput kernel
  float* d_filter_tm16_n64 = nullptr;  // the same with all slices 64 wide
  float* d_filter_small = nullptr;     // 3x3 TF32: U in the plain KN=32 image of the small-batch kernel (may alias d_filter)
  float* d_filter_direct = nullptr;    // 3x3, 14x14: per-tap weight blocks of the direct-convolution kernel (TF32 or 16-bit), or null
  float* d_scale = nullptr;
  float* d_shift = nullptr;
  size_t img_bytes[kNumImgs] = {0, 0, 0, 0, 0, 0};
  std::vector<float> w_host;  // the raw weights as given to create() (kept for wg_layer_serialize)
  // tensor-map caches
  std::mutex mu;
  // tm_x / tm_x16 / tm_small / tm_y / tm_res: the Winograd and pixels-on-M 1x1 kernels' maps. tm_xd / tm_yd / tm_ydp: the
  // direct 3x3 kernels' input, dense-output and frame-output maps; a 1x1 layer uses the same three for the transposed
  // kernel's input, output and residual maps (a layer is one kind, the boxes never mix).
  TmapCache tm_x, tm_x16, tm_small, tm_y, tm_res, tm_xd, tm_yd, tm_ydp, tm_xf, tm_yf;  // tm_xf / tm_yf: 1x1 frame output
  size_t in_px() const { return kind == 0 ? (size_t)geo.Hf * geo.Wf : (size_t)H * W; }
  size_t out_px(int padded) const {
    return padded ? (kind == 0 ? (size_t)geo.Hf * geo.Wf : (size_t)one_geo.Hf * one_geo.Wf) : (size_t)H * W;
  }
  // staging for wg_run_host
  float* d_x = nullptr;
  float* d_y = nullptr;
  size_t d_x_bytes = 0, d_y_bytes = 0;
  cudaStream_t stream = nullptr;
  // wg_run_host pipeline: copy-in / compute / copy-out streams and per-chunk events (created on first use)
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_done;
};

static int check_device(int device, int* num_sms) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return WG_ERR_NODEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return WG_ERR_NODEVICE;
  // the library embeds sm_100a SASS only (tcgen05 / TMEM, arch-specific target): exactly compute capability 10.0
  if (prop.major != 10 || prop.minor != 0) return WG_ERR_NODEVICE;
  *num_sms = prop.multiProcessorCount;
  return WG_OK;
}

// ---- layer construction: (1) shape-dependent kernel choices, (2) device buffers, (3) images packed on the GPU from
// the raw weights (create) or uploaded as they are (deserialize)
static void layer_plan(wg_layer* L) {
  if (L->kind == 0) {
    if (L->dtype != WG_TF32) L->tile_n = 64;
    else if (kn_tm(wino_kn())) L->tile_n = wino_kn();
    else L->tile_n = (wino_kn() == 32 || L->cout % 64 != 0) ? 32 : 64;
    L->ff_cg2 = wino_ff_cg2();
#ifdef WG_DEV_BUILD
    if (L->tile_n == 48) L->tm_db = wino_tm_choose_db(L->cin, L->cout);
#endif
    L->tm16_ff = (L->dtype != WG_TF32 && wino_kn() == 96) ? 1 : 0;
    const size_t fe = (size_t)16 * L->cin * L->cout;
    L->img_bytes[0] = fe * 4;
    L->img_bytes[1] = (L->tile_n == 96 && !L->ff_cg2 && wino_ff_has_narrow(L->cout)) ? fe * 4 : 0;
    const bool t16 = L->dtype != WG_TF32 && kn_tm(wino_kn());
    L->img_bytes[2] = t16 ? fe * 2 : 0;
    L->img_bytes[3] = (t16 && L->tm16_ff && !L->ff_cg2 && wino_ff_has_narrow(L->cout)) ? fe * 2 : 0;
    // the small-batch kernel (and its plain KN=32 image) exists for the reference's 14x14 geometry only
    L->img_bytes[4] = (L->dtype == WG_TF32 && L->tile_n != 32 && ff::geo_is_ref(L->geo)) ? fe * 4 : 0;  // tile_n == 32: aliases d_filter
    // the direct-convolution kernel: TF32, reference geometry, 32-channel chunks, 128-cout blocks
    // (16-bit operands: 64-channel chunks, 2-byte elements)
    // other map sizes: geometry as launch parameters (direct_gen_geo decides whether the map fits the kernel's stages)
    L->dgen = !ff::geo_is_ref(L->geo) &&
              direct_gen_geo(L->geo.H, L->geo.W, L->geo.Hf, L->geo.Wf,
                             L->dtype == WG_TF32 ? kDirectGenMaxRows : kDirect16GenMaxRows, &L->dgeo);
    // (TF32: a last block of 64 couts is zero-padded to 128 in the image; 16-bit operands: whole 128-cout blocks only)
    const bool dir_ok = (ff::geo_is_ref(L->geo) || L->dgen) && L->cout % (L->dtype == WG_TF32 ? 64 : 128) == 0 &&
                        L->cin % (L->dtype == WG_TF32 ? 32 : 64) == 0;
    const size_t cout_pad = (size_t)(L->cout + 127) / 128 * 128;
    L->img_bytes[5] = dir_ok ? (size_t)9 * L->cin * (L->dtype == WG_TF32 ? cout_pad * 4 : (size_t)L->cout * 2) : 0;
  } else {
    L->tile_n = (L->cout % 256 == 0 && L->dtype == WG_TF32) ? 256 : 128;  // bf16 operands: 128-wide N-tiles
    L->img_bytes[0] = (size_t)L->cin * L->cout * (L->dtype == WG_TF32 ? 4 : 2);
  }
}

static float** layer_img_slot(wg_layer* L, int i) {
  switch (i) {
    case 0: return &L->d_filter;
    case 1: return &L->d_filter_n64;
    case 2: return &L->d_filter_tm16;
    case 3: return &L->d_filter_tm16_n64;
    case 4: return &L->d_filter_small;
    default: return &L->d_filter_direct;
  }
}

static int layer_alloc(wg_layer* L) {
  WG_CUDA(cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking));
  for (int i = 0; i < kNumImgs; ++i)
    if (L->img_bytes[i]) WG_CUDA(cudaMalloc(layer_img_slot(L, i), L->img_bytes[i]));
  if (L->kind == 0 && L->dtype == WG_TF32 && L->tile_n == 32) L->d_filter_small = L->d_filter;
  WG_CUDA(cudaMalloc(&L->d_scale, L->cout * sizeof(float)));
  WG_CUDA(cudaMalloc(&L->d_shift, L->cout * sizeof(float)));
  return WG_OK;
}

static int layer_pack(wg_layer* L, const float* d_w) {
  const int cin = L->cin, cout = L->cout;
  const int op16 = L->dtype == WG_FP16 ? 2 : (L->dtype == WG_BF16 ? 1 : 0);
  int rc = WG_OK;
  auto count = [&] { g_launches++; };
  if (L->kind == 1) {
    rc = weight_pack_launch(d_w, L->d_filter, cin, cout, L->tile_n, op16, L->stream);
    count();
    return rc;
  }
  if (L->tile_n == 96) {
    rc = filter_transform_ff_launch(d_w, L->d_filter, cin, cout, 0, L->ff_cg2, 0, L->stream);
    count();
    if (rc == WG_OK && L->d_filter_n64) {
      rc = filter_transform_ff_launch(d_w, L->d_filter_n64, cin, cout, 0, 0, 1, L->stream);
      count();
    }
  }
#ifdef WG_DEV_BUILD
  else if (L->tile_n == 48) {
    rc = filter_transform_tm_launch(d_w, L->d_filter, cin, cout, L->tm_db, 0, L->stream);
    count();
  }
#endif
  else {
    rc = filter_transform_launch(d_w, L->d_filter, cin, cout, L->tile_n, (int)L->dtype, L->stream);
    count();
  }
  if (rc == WG_OK && L->d_filter_tm16) {
    // 16-bit operands: the throughput kernel keeps V packed in TMEM (its own filter image); small batches stay on the
    // split-C variant of the shared-memory-operand kernel (d_filter)
    if (L->tm16_ff) rc = filter_transform_ff_launch(d_w, L->d_filter_tm16, cin, cout, op16, L->ff_cg2, 0, L->stream);
#ifdef WG_DEV_BUILD
    else rc = filter_transform_tm_launch(d_w, L->d_filter_tm16, cin, cout, 0, op16, L->stream);
#endif
    count();
    if (rc == WG_OK && L->d_filter_tm16_n64) {
      rc = filter_transform_ff_launch(d_w, L->d_filter_tm16_n64, cin, cout, op16, 0, 1, L->stream);
      count();
    }
  }
  if (rc == WG_OK && L->img_bytes[4]) {
    rc = filter_transform_launch(d_w, L->d_filter_small, cin, cout, 32, 0, L->stream);
    count();
  }
  if (rc == WG_OK && L->img_bytes[5]) {
    rc = direct_pack_launch(d_w, L->d_filter_direct, cin, cout, op16, L->stream);
    count();
  }
  return rc;
}

static int layer_set_geometry(wg_layer* L, int H, int W) {
  L->H = H, L->W = W;
  ff::Geo g{};
  const int rc = wino_ff_geo(H, W, &g);  // also defines the frame a 1x1 layer writes in chain mode
  if (rc != WG_OK) return rc;
  L->geo = g;
  L->one_geo = OneGeo{H, W, g.Hf, g.Wf, 0};
  return WG_OK;
}

static int create_common(wg_layer_t** out, int kind, int cin, int cout, int H, int W, const float* w, size_t w_elems,
                         const float* scale, const float* shift, int relu, wg_dtype_t dtype, int device) {
  if (!out || !w || !scale || !shift) return WG_ERR_ARG;
  if (dtype != WG_TF32 && dtype != WG_BF16 && dtype != WG_FP16) return WG_ERR_ARG;
  // 16-bit operands: 3x3 needs C % 16 == 0 and K % 64 == 0; 1x1 has a bf16 variant (no fp16 one: the reference's 1x1
  // activations are U(-20,20) pre-BN sums, outside what fp16 operands are stated for)
  if (dtype != WG_TF32 && kind == 0 && (cin % 16 != 0 || cout % 64 != 0)) return WG_ERR_ARG;
  if (dtype == WG_FP16 && kind == 1) return WG_ERR_ARG;
  int num_sms = 0;
  int rc = check_device(device, &num_sms);
  if (rc != WG_OK) return rc;
  DeviceGuard guard;
  if (guard.enter(device) != WG_OK) return cuda_fail(cudaGetLastError(), "cudaSetDevice");

  wg_layer* L = new (std::nothrow) wg_layer();
  if (!L) return WG_ERR_NOMEM;
  L->kind = kind;
  L->cin = cin;
  L->cout = cout;
  L->relu = relu ? 1 : 0;
  L->dtype = dtype;
  L->device = device;
  L->num_sms = num_sms;
  if ((rc = layer_set_geometry(L, H, W)) != WG_OK) {
    delete L;
    return rc;
  }
  layer_plan(L);
  float* d_w = nullptr;
  auto fail = [&](int code) {
    wg_destroy(L);
    if (d_w) cudaFree(d_w);
    return code;
  };
  try {
    L->w_host.assign(w, w + w_elems);
  } catch (...) {
    return fail(WG_ERR_NOMEM);
  }
  if ((rc = layer_alloc(L)) != WG_OK) return fail(rc);
  cudaError_t e;
#define WG_TRY(call)                 \
  if ((e = (call)) != cudaSuccess) { \
    cuda_fail(e, #call);             \
    return fail(WG_ERR_CUDA);        \
  }
  WG_TRY(cudaMalloc(&d_w, w_elems * sizeof(float)));
  WG_TRY(cudaMemcpyAsync(d_w, w, w_elems * sizeof(float), cudaMemcpyHostToDevice, L->stream));
  WG_TRY(cudaMemcpyAsync(L->d_scale, scale, cout * sizeof(float), cudaMemcpyHostToDevice, L->stream));
  WG_TRY(cudaMemcpyAsync(L->d_shift, shift, cout * sizeof(float), cudaMemcpyHostToDevice, L->stream));
  rc = layer_pack(L, d_w);
  if (rc != WG_OK) {
    cuda_fail(cudaGetLastError(), "filter pack launch");
    return fail(rc);
  }
  WG_TRY(cudaStreamSynchronize(L->stream));
#undef WG_TRY
  cudaFree(d_w);
  *out = L;
  return WG_OK;
}

// ---- wg_run_host chunk schedule (host-only logic, exported for the unit tests)
// A ramp-up head (16, 32), full chunks of `chunk` images, then a tapering tail (halving down to 16 images): the copy-in
// and copy-out streams are about equally long, so both the FIRST chunk's copy-in (nothing can be copied out before it)
// and the LAST chunk's kernel + copy-out are exposed: both ends are small (256 images: 16, 32, 64, 64, 40, 20, 20). At most kMaxChunks chunks: the chunk size grows with N so that
// the schedule (and the per-chunk events) stay bounded for any N.
static constexpr int kMaxChunks = 64;
static int chunk_schedule(int N, int chunk, int taper, int* sizes, int cap) {
  if (N <= 0 || chunk <= 0) return 0;
  // equal chunks are capped well below kMaxChunks: the taper adds about log2(chunk / 8) more
  const int max_equal = kMaxChunks - 24;
  if ((N + chunk - 1) / chunk > max_equal) chunk = (N + max_equal - 1) / max_equal;
  int n = 0;
  int rem = N;
  // ramp-up head (16, 32 images): the copy-out stream cannot start before the first chunk has been copied in and
  // computed, and with output bytes ~ input bytes it is as long as the copy-in stream -- what is exposed at the START is
  // the first chunk's copy-in, so the first chunks are small too
  if (taper && N >= 128 && chunk >= 64) {
    for (int c = 16; c <= 32; c *= 2) {
      if (sizes && n < cap) sizes[n] = c;
      ++n;
      rem -= c;
    }
  }
  while (rem > 0) {
    int c = rem < chunk ? rem : chunk;
    if (taper && rem <= 2 * chunk && rem > 16) {
      c = rem / 2;
      if (c < 16) c = 16;
      if (c > chunk) c = chunk;
    }
    if (rem - c > 0 && rem - c < 16) c = rem;  // no crumbs: a remainder below 16 images rides with this chunk
    if (n == kMaxChunks - 1) c = rem;  // hard bound: whatever is left goes out as one last chunk
    if (sizes && n < cap) sizes[n] = c;
    ++n;
    rem -= c;
  }
  return n;
}

// ---- packed per-layer blob (wg_layer_serialize / wg_layer_deserialize). Little-endian; every field is 4 or 8 bytes
// and the header is a multiple of 8 bytes.
struct BlobHeader {
  char magic[8];          // "WGB200L\0"
  uint32_t version;       // kBlobVersion
  uint32_t header_bytes;  // sizeof(BlobHeader)
  int32_t kind, cin, cout, relu, dtype, tile_n, tm_db, tm16_ff, ff_cg2, dev_build;
  int32_t height, width;   // output map (14 x 14 for every reference shape)
  uint64_t w_elems;        // raw weights (fp32): first payload section, then scale[cout], shift[cout]
  uint64_t img_bytes[kNumImgs];   // packed images, in layer_img_slot() order
  uint64_t payload_bytes;  // everything after the header
  uint64_t checksum;       // FNV-1a 64 over the payload
};
static constexpr uint32_t kBlobVersion = 3;  // 3: + the direct-convolution kernel's weight image
static_assert(sizeof(BlobHeader) % 8 == 0, "header layout");

static uint64_t fnv1a(const uint8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

// 3x3, TF32, reference geometry: batch size from which the direct-convolution kernel is used (below it the split-C
// Winograd latency kernels win: their channel loop is divided over a cluster). WG_3X3_DIRECT_MIN (developer build)
// overrides; a huge value keeps every batch on the Winograd kernels.
static int direct_min_batch(int cin, int dtype) {
  static int env = -2;
  if (env == -2) {
    const char* e = dev_env("WG_3X3_DIRECT_MIN");
    env = e ? atoi(e) : -1;
  }
  if (env >= 0) return env;
  // measured cross-overs with the split-C latency kernels (profiles/direct3x3_r02.md); 16-bit operands: the direct
  // kernel ties at N = 1..2 and wins from there on
  if (dtype != WG_TF32) return 3;
  return cin >= 256 ? 6 : 11;
}

// developer build: WG_ONE_T=0 keeps the wide-Cout 1x1 shapes on conv1x1_bn_act_kernel (A/B runs)
static int onet_min_rows() {
  static int v = -2;
  if (v == -2) {
    const char* e = dev_env("WG_ONE_T");
    v = (e && atoi(e) == 0) ? -1 : 0;
  }
  return v;
}

static int run_impl(wg_layer_t* L, const float* x, const float* residual, float* y, int N, int flags,
                    cudaStream_t stream) {
  if (!L || !x || !y || N <= 0) return WG_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) ||
      (reinterpret_cast<uintptr_t>(residual) & 15))
    return WG_ERR_ARG;
  // flags: WG_OUT_PADDED (1) = zero-bordered frame, WG_OUT_MULTICAST (2) = y is an NVLS multicast address (stores go
  // out as multimem.st and land in every GPU's buffer: fused conv + all-gather of the output), WG_OUT_RELU_AFTER_ADD
  // (4, with a residual) = ReLU applied to the sum.
  if (flags & ~15) return WG_ERR_ARG;
  // WG_OUT_INTERIOR_ONLY (8, 1x1 layers, with WG_OUT_PADDED): the caller guarantees that the frame's border is already zero
  if ((flags & WG_OUT_INTERIOR_ONLY) && !(L->kind == 1 && (flags & WG_OUT_PADDED))) return WG_ERR_ARG;
  if ((flags & WG_OUT_MULTICAST) && !(L->kind == 0 && L->dtype == WG_TF32 && kn_tm(L->tile_n))) return WG_ERR_ARG;
  if ((flags & WG_OUT_RELU_AFTER_ADD) && !residual) return WG_ERR_ARG;
  // the residual add exists where the reference's block structure puts it: after the 1x1 `_out` layers, dense output
  if (residual && (L->kind != 1 || (flags & (WG_OUT_PADDED | WG_OUT_MULTICAST)))) return WG_ERR_ARG;
  const int out_flags = flags & 3;
  const int out_padded = flags & 1;
  DeviceGuard guard;
  if (guard.enter(L->device) != WG_OK) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  const int mc = g_max_ctas.load();
  const int max_ctas = mc > 0 ? mc : L->num_sms;
  CUtensorMap tmap, tmap_y, tmap_res;
  int rc = WG_OK;
  auto launched = [&](int r) {
    g_launches++;
    if (r == WG_ERR_CUDA) cuda_fail(cudaGetLastError(), "kernel launch");
    return r;
  };

  const bool ref_geo = L->H == 14 && L->W == 14;
  const long long px = (long long)L->H * L->W;  // 1x1: GEMM rows per image
  // map sizes other than the reference's run the full-fold kernel only (developer build: not with a superseded generation)
  if (L->kind == 0 && !ref_geo && !(L->dtype == WG_TF32 ? L->tile_n == 96 : L->tm16_ff != 0)) return WG_ERR_ARG;
  if (L->kind == 0 && L->d_filter_direct != nullptr && L->dgen && !(out_flags & 2)) {
    // other map sizes: the direct kernels with runtime geometry, every batch size
    {
      std::lock_guard<std::mutex> lk(L->mu);
      rc = L->tm_xd.get(x, N, &tmap, [&](CUtensorMap* m) { return direct_gen_make_tmap_in(m, x, N, L->cin, L->dgeo); });
    }
    if (rc != WG_OK) return rc;
    if (L->dtype != WG_TF32)
      return launched(direct16_gen_launch(tmap, L->d_filter_direct, L->d_scale, L->d_shift, y, N, L->cin, L->cout,
                                          L->dtype == WG_FP16 ? 2 : 1, L->relu, out_padded, max_ctas, L->dgeo, stream));
    return launched(direct_gen_launch(tmap, L->d_filter_direct, L->d_scale, L->d_shift, y, N, L->cin, L->cout, L->relu,
                                      out_padded, max_ctas, L->dgeo, stream));
  }
  if (L->kind == 0 && L->d_filter_direct != nullptr && !L->dgen && !(out_flags & 2) &&
      N >= direct_min_batch(L->cin, L->dtype)) {
    // throughput-sized batches of the reference geometry: direct convolution on the tensor core (no CUDA-core transform;
    // measured faster than the fused Winograd pipeline from a handful of images on, see conv3x3_direct_kernel.cu)
    {
      std::lock_guard<std::mutex> lk(L->mu);
      rc = L->tm_xd.get(x, N, &tmap, [&](CUtensorMap* m) { return direct_make_tmap_in(m, x, N, L->cin); });
      if (rc == WG_OK)
        rc = (out_padded ? L->tm_ydp : L->tm_yd).get(y, N, &tmap_y, [&](CUtensorMap* m) {
          return direct_make_tmap_out(m, y, N, L->cout, out_padded);
        });
    }
    if (rc != WG_OK) return rc;
    if (L->dtype != WG_TF32)
      return launched(direct16_launch(tmap, tmap_y, L->d_filter_direct, L->d_scale, L->d_shift, N, L->cin, L->cout,
                                      L->dtype == WG_FP16 ? 2 : 1, L->relu, out_padded, max_ctas, 1, stream));
    return launched(direct_launch(tmap, tmap_y, L->d_filter_direct, L->d_scale, L->d_shift, N, L->cin, L->cout, 1,
                                  L->relu, out_padded, max_ctas, 1, stream));
  }
  if (L->kind == 0 && ref_geo && L->d_filter_small != nullptr) {
    // small batches: the latency variant (one 64-tile x 32-cout item per cluster, split-C), see wino_small_kernel.cu
    const int cs = wino_small_cs(N, L->cin, L->cout, max_ctas);
    if (cs > 0) {
      {
        std::lock_guard<std::mutex> lk(L->mu);
        rc = L->tm_small.get(x, N, &tmap, [&](CUtensorMap* m) { return wino_small_make_tmap(m, x, N, L->cin); });
      }
      if (rc != WG_OK) return rc;
      return launched(wino_small_launch(tmap, L->d_filter_small, L->d_scale, L->d_shift, y, N, L->cin, L->cout, L->relu,
                                        out_flags, cs, stream));
    }
  }
  if (L->kind == 0 && L->d_filter_tm16 != nullptr && !(out_flags & 2)) {
    // bf16 / fp16 operands, throughput-sized batch (the split-C latency mode of the other kernel takes the small ones
    // of the reference geometry; other map sizes always run the full-fold kernel)
    const long long items64 = (long long)((N * 49 + 63) / 64) * (L->cout / 64);
    if (items64 * 4 > max_ctas || !ref_geo) {
      {
        std::lock_guard<std::mutex> lk(L->mu);
        rc = L->tm_x16.get(x, N, &tmap, [&](CUtensorMap* m) {
#ifdef WG_DEV_BUILD
          if (!L->tm16_ff) return wino_tm_make_tmap(m, x, N, L->cin, 1);
#endif
          return wino_ff_make_tmap(m, x, N, L->cin, L->geo);
        });
      }
      if (rc != WG_OK) return rc;
      const int op16 = L->dtype == WG_FP16 ? 2 : 1;
#ifdef WG_DEV_BUILD
      if (!L->tm16_ff)
        return launched(wino_tm_launch(tmap, L->d_filter_tm16, L->d_scale, L->d_shift, y, N, L->cin, L->cout, 0, op16,
                                       L->relu, out_flags, max_ctas, stream));
#endif
      return launched(wino_ff_launch(tmap, x, L->d_filter_tm16, L->d_filter_tm16_n64, L->d_scale, L->d_shift, y, N,
                                     L->cin, L->cout, op16, L->ff_cg2, L->relu, out_flags, max_ctas, L->geo, stream));
    }
  }
  if (L->kind == 1 && L->dtype == WG_TF32 && out_padded && !residual && ref_geo &&
      onetf_eligible(N, L->cin, L->cout, max_ctas)) {
    // chain mode at throughput sizes: the transposed kernel writing the next 3x3 layer's frame, one TMA store per row
    {
      std::lock_guard<std::mutex> lk(L->mu);
      rc = L->tm_xf.get(x, N, &tmap, [&](CUtensorMap* m) { return onetf_make_tmap_in(m, x, (long long)N * px, L->cin, L->cout); });
      if (rc == WG_OK) rc = L->tm_yf.get(y, N, &tmap_y, [&](CUtensorMap* m) { return onetf_make_tmap_out(m, y, N, L->cout); });
    }
    if (rc != WG_OK) return rc;
    return launched(onetf_launch(tmap, tmap_y, L->d_filter, L->d_scale, L->d_shift, N, L->cin, L->cout, L->relu,
                                 L->tile_n, (flags & WG_OUT_INTERIOR_ONLY) ? 1 : 0, max_ctas, stream));
  }
  if (L->kind == 1 && L->dtype == WG_TF32 && !out_padded && onet_min_rows() >= 0 &&
      onet_eligible((long long)N * px, L->cin, L->cout, max_ctas)) {
    // wide-Cout shapes at throughput sizes: couts on M, pixels on N, weight slab resident (conv1x1_t_kernel.cu)
    const long long rows = (long long)N * px;
    {
      std::lock_guard<std::mutex> lk(L->mu);
      rc = L->tm_xd.get(x, N, &tmap, [&](CUtensorMap* m) { return onet_make_tmap_in(m, x, rows, L->cin, L->cout); });
      if (rc == WG_OK) rc = L->tm_yd.get(y, N, &tmap_y, [&](CUtensorMap* m) { return onet_make_tmap_out(m, y, rows, L->cout); });
      if (rc == WG_OK && residual)
        rc = L->tm_ydp.get(residual, N, &tmap_res,
                           [&](CUtensorMap* m) { return onet_make_tmap_out(m, residual, rows, L->cout); });
    }
    if (rc != WG_OK) return rc;
    return launched(onet_launch(tmap, tmap_y, residual ? &tmap_res : nullptr, L->d_filter, L->d_scale, L->d_shift, rows,
                                L->cin, L->cout, L->relu, L->tile_n, (flags & WG_OUT_RELU_AFTER_ADD) ? 1 : 0, max_ctas,
                                stream));
  }
  {
    std::lock_guard<std::mutex> lk(L->mu);
    rc = L->tm_x.get(x, N, &tmap, [&](CUtensorMap* m) {
      if (L->kind == 1) return one_make_tmap(m, x, (long long)N * px, L->cin);
      if (L->tile_n == 96) return wino_ff_make_tmap(m, x, N, L->cin, L->geo);
#ifdef WG_DEV_BUILD
      if (L->tile_n == 48) return wino_tm_make_tmap(m, x, N, L->cin, wino_tm_cls(L->cout, L->tm_db));
#endif
      return wino_make_tmap(m, x, N, L->cin);
    });
    if (rc == WG_OK && L->kind == 1)
      rc = L->tm_y.get(y, N, &tmap_y,
                       [&](CUtensorMap* m) { return one_make_tmap_out(m, y, (long long)N * px, L->cout); });
    if (rc == WG_OK && L->kind == 1 && residual)
      rc = L->tm_res.get(residual, N, &tmap_res,
                         [&](CUtensorMap* m) { return one_make_tmap_out(m, residual, (long long)N * px, L->cout); });
  }
  if (rc != WG_OK) return rc;
  if (L->kind == 1) {
    if (!residual) tmap_res = tmap_y;
    return launched(one_launch(tmap, tmap_y, tmap_res, L->d_filter, L->d_scale, L->d_shift, y,
                               out_padded | ((flags & WG_OUT_INTERIOR_ONLY) ? 2 : 0),
                               (long long)N * px, L->cin, L->cout, L->tile_n, L->dtype != WG_TF32, L->relu, residual,
                               (flags & WG_OUT_RELU_AFTER_ADD) ? 1 : 0, max_ctas, L->one_geo, stream));
  }
  if (L->tile_n == 96)
    return launched(wino_ff_launch(tmap, x, L->d_filter, L->d_filter_n64, L->d_scale, L->d_shift, y, N, L->cin, L->cout,
                                   0, L->ff_cg2, L->relu, out_flags, max_ctas, L->geo, stream));
#ifdef WG_DEV_BUILD
  if (L->tile_n == 48)
    return launched(wino_tm_launch(tmap, L->d_filter, L->d_scale, L->d_shift, y, N, L->cin, L->cout, L->tm_db, 0,
                                   L->relu, out_flags, max_ctas, stream));
#endif
  return launched(wino_launch(tmap, L->d_filter, L->d_scale, L->d_shift, y, N, L->cin, L->cout, L->tile_n, L->dtype,
                              L->relu, out_padded, max_ctas, stream));
}

extern "C" {

int wg_conv3x3_create_hw(wg_layer_t** out, int C, int K, int H, int W, const float* w_kcrs, const float* scale,
                         const float* shift, int relu, wg_dtype_t dtype, int device) {
  if (C <= 0 || K <= 0 || C % 8 != 0 || K % 32 != 0 || H < 3 || W < 3) return WG_ERR_ARG;
  return create_common(out, 0, C, K, H, W, w_kcrs, (size_t)K * C * 9, scale, shift, relu, dtype, device);
}

int wg_conv1x1_create_hw(wg_layer_t** out, int Cin, int Cout, int H, int W, const float* w_cin_cout, const float* scale,
                         const float* shift, int relu, wg_dtype_t dtype, int device) {
  if (Cin <= 0 || Cout <= 0 || Cin % 32 != 0 || Cout % 128 != 0 || H < 3 || W < 3) return WG_ERR_ARG;
  return create_common(out, 1, Cin, Cout, H, W, w_cin_cout, (size_t)Cin * Cout, scale, shift, relu, dtype, device);
}

int wg_conv3x3_create(wg_layer_t** out, int C, int K, const float* w_kcrs, const float* scale, const float* shift,
                      int relu, wg_dtype_t dtype, int device) {
  return wg_conv3x3_create_hw(out, C, K, 14, 14, w_kcrs, scale, shift, relu, dtype, device);
}

int wg_conv1x1_create(wg_layer_t** out, int Cin, int Cout, const float* w_cin_cout, const float* scale,
                      const float* shift, int relu, wg_dtype_t dtype, int device) {
  return wg_conv1x1_create_hw(out, Cin, Cout, 14, 14, w_cin_cout, scale, shift, relu, dtype, device);
}

int wg_frame_dims(int H, int W, int* frame_h, int* frame_w) {
  ff::Geo g{};
  const int rc = wino_ff_geo(H, W, &g);
  if (rc != WG_OK) return rc;
  if (frame_h) *frame_h = g.Hf;
  if (frame_w) *frame_w = g.Wf;
  return WG_OK;
}

int wg_direct_geometry(int H, int W, wg_dtype_t dtype, int* out9) {
  if (!out9 || (dtype != WG_TF32 && dtype != WG_BF16 && dtype != WG_FP16)) return WG_ERR_ARG;
  ff::Geo g{};
  const int rc = wino_ff_geo(H, W, &g);
  if (rc != WG_OK) return rc;
  DirGeo d{};
  if (ff::geo_is_ref(g)) {  // the reference geometry has its own kernels: one image (or half of one) per item
    d.R = 14, d.bands = 1, d.G = 1, d.n_pad = 224, d.halo = 24, d.n_boxes = 2, d.box_rows = 136;
  } else if (!direct_gen_geo(g.H, g.W, g.Hf, g.Wf, dtype == WG_TF32 ? kDirectGenMaxRows : kDirect16GenMaxRows, &d)) {
    return 0;
  }
  const int v[9] = {d.R, d.bands, d.G, d.n_pad, d.halo, d.n_boxes, d.box_rows, g.Hf, g.Wf};
  for (int i = 0; i < 9; ++i) out9[i] = v[i];
  return 1;
}

int wg_layer_geometry(const wg_layer_t* L, int* H, int* W, int* frame_h, int* frame_w) {
  if (!L) return WG_ERR_ARG;
  if (H) *H = L->H;
  if (W) *W = L->W;
  if (frame_h) *frame_h = L->geo.Hf;
  if (frame_w) *frame_w = L->geo.Wf;
  return WG_OK;
}

int wg_run(wg_layer_t* L, const float* x, float* y, int N, int out_flags, void* cuda_stream) {
  if (out_flags & WG_OUT_RELU_AFTER_ADD) return WG_ERR_ARG;
  return run_impl(L, x, nullptr, y, N, out_flags, static_cast<cudaStream_t>(cuda_stream));
}

int wg_run_residual(wg_layer_t* L, const float* x, const float* residual, float* y, int N, int out_flags,
                    void* cuda_stream) {
  if (!residual) return WG_ERR_ARG;
  return run_impl(L, x, residual, y, N, out_flags, static_cast<cudaStream_t>(cuda_stream));
}

int wg_host_chunk_schedule(int N, int* sizes, int cap) { return chunk_schedule(N, 64, 1, sizes, cap); }

int wg_run_host(wg_layer_t* L, const float* x_host, float* y_host, int N, int out_padded) {
  if (!L || !x_host || !y_host || N <= 0 || (out_padded & ~1)) return WG_ERR_ARG;
  DeviceGuard guard;
  if (guard.enter(L->device) != WG_OK) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  const size_t in_px = L->in_px();
  const size_t out_px = L->out_px(out_padded);
  const size_t xb = (size_t)N * in_px * L->cin * sizeof(float);
  const size_t yb = (size_t)N * out_px * L->cout * sizeof(float);
  if (L->d_x_bytes < xb) {
    if (L->d_x) cudaFree(L->d_x);
    L->d_x = nullptr;
    L->d_x_bytes = 0;
    {
      std::lock_guard<std::mutex> lk(L->mu);  // the freed address may come back with other contents / sizes
      L->tm_x.clear(), L->tm_x16.clear(), L->tm_small.clear(), L->tm_xd.clear(), L->tm_xf.clear();
    }
    WG_CUDA(cudaMalloc(&L->d_x, xb));
    L->d_x_bytes = xb;
  }
  if (L->d_y_bytes < yb) {
    if (L->d_y) cudaFree(L->d_y);
    L->d_y = nullptr;
    L->d_y_bytes = 0;
    {
      std::lock_guard<std::mutex> lk(L->mu);
      L->tm_y.clear(), L->tm_yd.clear(), L->tm_ydp.clear(), L->tm_yf.clear();
    }
    WG_CUDA(cudaMalloc(&L->d_y, yb));
    L->d_y_bytes = yb;
  }
  // Chunked three-stream pipeline: H2D of chunk c+1, the kernel on chunk c and D2H of chunk c-1 overlap (PCIe is
  // full duplex), so a large batch costs ~max(copy-in, copy-out) instead of their sum plus the kernel. Host buffers
  // should be pinned for the copies to be asynchronous; pageable memory still works, just without the overlap.
  if (!L->s_h2d) WG_CUDA(cudaStreamCreateWithFlags(&L->s_h2d, cudaStreamNonBlocking));
  if (!L->s_d2h) WG_CUDA(cudaStreamCreateWithFlags(&L->s_d2h, cudaStreamNonBlocking));
  // images per chunk: 64 (measured: 64 > 32 > 16 on PCIe Gen5; per-chunk event and launch costs outweigh the shorter
  // pipeline tail); developer build: WG_HOST_CHUNK / WG_HOST_TAPER=0 for the A/B measurements
  static int chunk_env = -1, taper_env = -1;
  if (chunk_env < 0) {
    const char* t = dev_env("WG_HOST_TAPER");
    taper_env = t ? atoi(t) : 1;
    const char* e = dev_env("WG_HOST_CHUNK");
    chunk_env = e ? atoi(e) : 0;
  }
  int sizes[kMaxChunks];
  const int n_chunks = chunk_schedule(N, chunk_env > 0 ? chunk_env : 64, taper_env, sizes, kMaxChunks);
  if (n_chunks <= 0 || n_chunks > kMaxChunks) return WG_ERR_ARG;
  while ((int)L->ev_in.size() < n_chunks) {
    cudaEvent_t a = nullptr;
    WG_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    L->ev_in.push_back(a);
  }
  while ((int)L->ev_done.size() < n_chunks) {
    cudaEvent_t b = nullptr;
    WG_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    L->ev_done.push_back(b);
  }
  const size_t x_img = in_px * L->cin, y_img = out_px * L->cout;  // floats per image
  int n0 = 0, rc = WG_OK;
  cudaError_t e = cudaSuccess;
  const char* what = "";
#define WG_STEP(call)                               \
  if (rc == WG_OK && (e = (call)) != cudaSuccess) { \
    rc = WG_ERR_CUDA;                               \
    what = #call;                                   \
  }
  for (int c = 0; c < n_chunks && rc == WG_OK; ++c) {
    const int nc = sizes[c];
    WG_STEP(cudaMemcpyAsync(L->d_x + (size_t)n0 * x_img, x_host + (size_t)n0 * x_img, (size_t)nc * x_img * sizeof(float),
                            cudaMemcpyHostToDevice, L->s_h2d));
    WG_STEP(cudaEventRecord(L->ev_in[c], L->s_h2d));
    WG_STEP(cudaStreamWaitEvent(L->stream, L->ev_in[c], 0));
    if (rc == WG_OK)
      rc = run_impl(L, L->d_x + (size_t)n0 * x_img, nullptr, L->d_y + (size_t)n0 * y_img, nc, out_padded, L->stream);
    WG_STEP(cudaEventRecord(L->ev_done[c], L->stream));
    WG_STEP(cudaStreamWaitEvent(L->s_d2h, L->ev_done[c], 0));
    WG_STEP(cudaMemcpyAsync(y_host + (size_t)n0 * y_img, L->d_y + (size_t)n0 * y_img, (size_t)nc * y_img * sizeof(float),
                            cudaMemcpyDeviceToHost, L->s_d2h));
    n0 += nc;
  }
#undef WG_STEP
  // on success AND on failure: nothing may still read x_host / write y_host when this function returns
  const cudaError_t s0 = cudaStreamSynchronize(L->s_h2d);
  const cudaError_t s1 = cudaStreamSynchronize(L->stream);
  const cudaError_t s2 = cudaStreamSynchronize(L->s_d2h);
  if (rc == WG_ERR_CUDA && e != cudaSuccess) return cuda_fail(e, what);
  if (rc != WG_OK) return rc;
  if (s0 != cudaSuccess) return cuda_fail(s0, "cudaStreamSynchronize(h2d)");
  if (s1 != cudaSuccess) return cuda_fail(s1, "cudaStreamSynchronize(compute)");
  if (s2 != cudaSuccess) return cuda_fail(s2, "cudaStreamSynchronize(d2h)");
  return WG_OK;
}

int wg_destroy(wg_layer_t* L) {
  if (!L) return WG_ERR_ARG;
  DeviceGuard guard;
  guard.enter(L->device);
  if (L->d_filter_small && L->d_filter_small != L->d_filter) cudaFree(L->d_filter_small);
  if (L->d_filter_direct) cudaFree(L->d_filter_direct);
  if (L->d_filter_tm16) cudaFree(L->d_filter_tm16);
  if (L->d_filter_n64) cudaFree(L->d_filter_n64);
  if (L->d_filter_tm16_n64) cudaFree(L->d_filter_tm16_n64);
  if (L->d_filter) cudaFree(L->d_filter);
  if (L->d_scale) cudaFree(L->d_scale);
  if (L->d_shift) cudaFree(L->d_shift);
  if (L->d_x) cudaFree(L->d_x);
  if (L->d_y) cudaFree(L->d_y);
  for (cudaEvent_t ev : L->ev_in) cudaEventDestroy(ev);
  for (cudaEvent_t ev : L->ev_done) cudaEventDestroy(ev);
  if (L->s_h2d) cudaStreamDestroy(L->s_h2d);
  if (L->s_d2h) cudaStreamDestroy(L->s_d2h);
  if (L->stream) cudaStreamDestroy(L->stream);
  delete L;
  return WG_OK;
}

// ------------------------------------------------------------------------------------------------ packed blob
int wg_layer_serialize(const wg_layer_t* Lc, void* buf, size_t cap, size_t* need) {
  wg_layer* L = const_cast<wg_layer*>(Lc);
  if (!L || !need) return WG_ERR_ARG;
  size_t payload = L->w_host.size() * 4 + (size_t)L->cout * 8;
  for (int i = 0; i < kNumImgs; ++i) payload += L->img_bytes[i];
  *need = sizeof(BlobHeader) + payload;
  if (!buf) return WG_OK;  // size query
  if (cap < *need) return WG_ERR_ARG;
  DeviceGuard guard;
  if (guard.enter(L->device) != WG_OK) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  uint8_t* p = static_cast<uint8_t*>(buf) + sizeof(BlobHeader);
  memcpy(p, L->w_host.data(), L->w_host.size() * 4);
  p += L->w_host.size() * 4;
  WG_CUDA(cudaMemcpy(p, L->d_scale, (size_t)L->cout * 4, cudaMemcpyDeviceToHost));
  p += (size_t)L->cout * 4;
  WG_CUDA(cudaMemcpy(p, L->d_shift, (size_t)L->cout * 4, cudaMemcpyDeviceToHost));
  p += (size_t)L->cout * 4;
  for (int i = 0; i < kNumImgs; ++i)
    if (L->img_bytes[i]) {
      WG_CUDA(cudaMemcpy(p, *layer_img_slot(L, i), L->img_bytes[i], cudaMemcpyDeviceToHost));
      p += L->img_bytes[i];
    }
  BlobHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "WGB200L", 8);
  h.version = kBlobVersion;
  h.header_bytes = (uint32_t)sizeof(BlobHeader);
  h.kind = L->kind, h.cin = L->cin, h.cout = L->cout, h.relu = L->relu, h.dtype = L->dtype, h.tile_n = L->tile_n;
  h.tm_db = L->tm_db, h.tm16_ff = L->tm16_ff, h.ff_cg2 = L->ff_cg2, h.dev_build = kDev ? 1 : 0;
  h.height = L->H, h.width = L->W;
  h.w_elems = L->w_host.size();
  for (int i = 0; i < kNumImgs; ++i) h.img_bytes[i] = L->img_bytes[i];
  h.payload_bytes = payload;
  h.checksum = fnv1a(static_cast<const uint8_t*>(buf) + sizeof(BlobHeader), payload);
  memcpy(buf, &h, sizeof(h));
  return WG_OK;
}

int wg_layer_deserialize(wg_layer_t** out, const void* buf, size_t bytes, int device) {
  if (!out || !buf) return WG_ERR_ARG;
  if (bytes < sizeof(BlobHeader)) return WG_ERR_IO;
  BlobHeader h;
  memcpy(&h, buf, sizeof(h));
  if (memcmp(h.magic, "WGB200L", 8) != 0 || h.version != kBlobVersion || h.header_bytes != sizeof(BlobHeader))
    return WG_ERR_IO;
  if (h.payload_bytes != bytes - sizeof(BlobHeader)) return WG_ERR_IO;
  const uint8_t* payload = static_cast<const uint8_t*>(buf) + sizeof(BlobHeader);
  if (fnv1a(payload, h.payload_bytes) != h.checksum) return WG_ERR_IO;
  if (h.kind < 0 || h.kind > 1 || h.cin <= 0 || h.cout <= 0 || h.height < 3 || h.width < 3) return WG_ERR_IO;
  const size_t w_expect = h.kind == 0 ? (size_t)h.cout * h.cin * 9 : (size_t)h.cin * h.cout;
  if (h.w_elems != w_expect) return WG_ERR_IO;
  size_t total = h.w_elems * 4 + (size_t)h.cout * 8;
  for (int i = 0; i < kNumImgs; ++i) total += h.img_bytes[i];
  if (total != h.payload_bytes) return WG_ERR_IO;
  const float* w = reinterpret_cast<const float*>(payload);
  const float* scale = w + h.w_elems;
  const float* shift = scale + h.cout;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(shift + h.cout);

  int num_sms = 0;
  int rc = check_device(device, &num_sms);
  if (rc != WG_OK) return rc;
  DeviceGuard guard;
  if (guard.enter(device) != WG_OK) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  wg_layer* L = new (std::nothrow) wg_layer();
  if (!L) return WG_ERR_NOMEM;
  L->kind = h.kind, L->cin = h.cin, L->cout = h.cout, L->relu = h.relu, L->dtype = h.dtype, L->device = device;
  L->num_sms = num_sms;
  if (layer_set_geometry(L, h.height, h.width) != WG_OK) {
    delete L;
    return WG_ERR_IO;
  }
  layer_plan(L);
  // The images are the shared-memory images of THIS build's kernels. A blob written by a build with other kernel
  // choices (developer knobs, an older version) still carries the raw weights: re-pack from those instead of failing.
  bool same = L->tile_n == h.tile_n && L->tm_db == h.tm_db && L->tm16_ff == h.tm16_ff && L->ff_cg2 == h.ff_cg2;
  for (int i = 0; i < kNumImgs; ++i) same = same && L->img_bytes[i] == h.img_bytes[i];
  if (!same) {
    delete L;
    return create_common(out, h.kind, h.cin, h.cout, h.height, h.width, w, h.w_elems, scale, shift, h.relu,
                         (wg_dtype_t)h.dtype, device);
  }
  auto fail = [&](int code) {
    wg_destroy(L);
    return code;
  };
  try {
    L->w_host.assign(w, w + h.w_elems);
  } catch (...) {
    return fail(WG_ERR_NOMEM);
  }
  if ((rc = layer_alloc(L)) != WG_OK) return fail(rc);
  cudaError_t e;
#define WG_TRY(call)                 \
  if ((e = (call)) != cudaSuccess) { \
    cuda_fail(e, #call);             \
    return fail(WG_ERR_CUDA);        \
  }
  WG_TRY(cudaMemcpyAsync(L->d_scale, scale, (size_t)h.cout * 4, cudaMemcpyHostToDevice, L->stream));
  WG_TRY(cudaMemcpyAsync(L->d_shift, shift, (size_t)h.cout * 4, cudaMemcpyHostToDevice, L->stream));
  for (int i = 0; i < kNumImgs; ++i)
    if (L->img_bytes[i]) {
      WG_TRY(cudaMemcpyAsync(*layer_img_slot(L, i), img, L->img_bytes[i], cudaMemcpyHostToDevice, L->stream));
      img += L->img_bytes[i];
    }
  WG_TRY(cudaStreamSynchronize(L->stream));
#undef WG_TRY
  *out = L;
  return WG_OK;
}

int wg_layer_save(const wg_layer_t* L, const char* path) {
  if (!L || !path) return WG_ERR_ARG;
  size_t need = 0;
  int rc = wg_layer_serialize(L, nullptr, 0, &need);
  if (rc != WG_OK) return rc;
  std::vector<uint8_t> buf;
  try {
    buf.resize(need);
  } catch (...) {
    return WG_ERR_NOMEM;
  }
  rc = wg_layer_serialize(L, buf.data(), buf.size(), &need);
  if (rc != WG_OK) return rc;
  FILE* f = fopen(path, "wb");
  if (!f) return WG_ERR_IO;
  const size_t wr = fwrite(buf.data(), 1, buf.size(), f);
  const int cl = fclose(f);
  return (wr == buf.size() && cl == 0) ? WG_OK : WG_ERR_IO;
}

int wg_layer_load(wg_layer_t** out, const char* path, int device) {
  if (!out || !path) return WG_ERR_ARG;
  FILE* f = fopen(path, "rb");
  if (!f) return WG_ERR_IO;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (sz < (long)sizeof(BlobHeader)) {
    fclose(f);
    return WG_ERR_IO;
  }
  std::vector<uint8_t> buf;
  try {
    buf.resize((size_t)sz);
  } catch (...) {
    fclose(f);
    return WG_ERR_NOMEM;
  }
  const size_t rd = fread(buf.data(), 1, buf.size(), f);
  fclose(f);
  if (rd != buf.size()) return WG_ERR_IO;
  return wg_layer_deserialize(out, buf.data(), buf.size(), device);
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
    case WG_ERR_IO: return "file missing, short or not a layer blob of this version";
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
    if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10 && prop.minor == 0) ++n;
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

void wg_set_max_ctas(int max_ctas) { g_max_ctas.store(max_ctas); }

#ifdef WG_DEV_BUILD
// developer build only (tools/libwinograd_b200_dev.so): choose a superseded 3x3 kernel generation
void wg_dev_set_wino_kn(int kn) { g_wino_kn.store((kn == 32 || kn == 64 || kn == 48) ? kn : 96); }
#endif

}  // extern "C"
