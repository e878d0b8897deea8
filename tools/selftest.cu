// Developer self-test (not part of the product or the test-suite): exercises libwinograd_b200.so through its C ABI
// against an in-program FP64 direct convolution, plus a one-instruction UMMA descriptor probe that tells which
// (LBO, SBO) reading of the no-swizzle K-major layout the hardware implements. Run on a B200:
//   tools/selftest [quick]
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../cuda-winograd_b200/csrc/ptx.cuh"
#include "../include/winograd_b200.h"
extern "C" void wg_dev_set_wino_kn(int kn);  // developer build only (tools/libwinograd_b200_dev.so)

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

static uint32_t rng_state = 12345;
static wg_dtype_t g_dtype = WG_TF32;  // operand type used for the 3x3 layers
static float frand() {  // U(-0.5, 0.5)
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xffffff) / 16777216.0f - 0.5f;
}

// ------------------------------------------------------------------------------------------------ UMMA probe
// A: 128 x 8 (K-major), B: 32 x 8 (K-major), both in the no-swizzle canonical layout with element (r, k) at byte
// (k/4)*kstride + (r/8)*gstride + (r%8)*16 + (k%4)*4. One tcgen05.mma, D dumped to global.
__global__ void umma_probe_kernel(const float* a, const float* b, float* d, uint32_t a_kstride, uint32_t a_gstride,
                                  uint32_t b_kstride, uint32_t b_gstride, uint32_t lbo_a, uint32_t sbo_a,
                                  uint32_t lbo_b, uint32_t sbo_b, int m64) {
  using namespace wg;
  __shared__ __align__(1024) uint8_t sa[8192];
  __shared__ __align__(1024) uint8_t sz[8192];  // zeros: an M=128 MMA with it clears the accumulator (m64 probe)
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) reinterpret_cast<uint32_t*>(sz)[i] = 0;
  __shared__ __align__(1024) uint8_t sb[4096];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x;
  for (int i = tid; i < 128 * 8; i += blockDim.x) {
    const int r = i / 8, k = i % 8;
    *reinterpret_cast<float*>(sa + (k / 4) * a_kstride + (r / 8) * a_gstride + (r % 8) * 16 + (k % 4) * 4) = a[i];
  }
  for (int i = tid; i < 32 * 8; i += blockDim.x) {
    const int r = i / 8, k = i % 8;
    *reinterpret_cast<float*>(sb + (k / 4) * b_kstride + (r / 8) * b_gstride + (r % 8) * 16 + (k % 4) * 4) = b[i];
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc<32>(&tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tptr;
  if (tid == 0) {
    const uint64_t ad = make_smem_desc(smem_u32(sa), lbo_a, sbo_a, kLayoutNone);
    const uint64_t bd = make_smem_desc(smem_u32(sb), lbo_b, sbo_b, kLayoutNone);
    if (m64) {  // where do the 64 rows of an M=64 accumulator land in TMEM? clear all 128 lanes, then M=64
      const uint64_t zd = make_smem_desc(smem_u32(sz), lbo_a, sbo_a, kLayoutNone);
      umma_tf32_ss(tb, zd, bd, make_idesc(kFmtTF32, 128, 32), 0);
      umma_tf32_ss(tb, ad, bd, make_idesc(kFmtTF32, 64, 32), 0);
    } else {
      umma_tf32_ss(tb, ad, bd, make_idesc(kFmtTF32, 128, 32), 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  float v[32];
  tmem_ld_x16(tb + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_x16(tb + ((uint32_t)(warp * 32) << 16) + 16, v + 16);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) d[(warp * 32 + lane) * 32 + j] = v[j];
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<32>(tb);
}

static int run_probe() {
  std::vector<float> a(128 * 8), b(32 * 8), d(128 * 32);
  for (auto& v : a) v = roundf(frand() * 16.f);  // small integers: exact in tf32
  for (auto& v : b) v = roundf(frand() * 16.f);
  float *da, *db, *dd;
  CK(cudaMalloc(&da, a.size() * 4));
  CK(cudaMalloc(&db, b.size() * 4));
  CK(cudaMalloc(&dd, d.size() * 4));
  CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  int ok_any = 0;
  // layout in smem: k-chunk stride 2112 (A) / 512 (B), 8-row-group stride 128
  for (int variant = 0; variant < 1; ++variant) {  // the swapped reading faults (read past shared memory): confirmed on B200
    const uint32_t aks = 2112, ags = 128, bks = 512, bgs = 128;
    const uint32_t lbo_a = variant == 0 ? aks : ags, sbo_a = variant == 0 ? ags : aks;
    const uint32_t lbo_b = variant == 0 ? bks : bgs, sbo_b = variant == 0 ? bgs : bks;
    CK(cudaMemset(dd, 0, d.size() * 4));
    umma_probe_kernel<<<1, 128>>>(da, db, dd, aks, ags, bks, bgs, lbo_a, sbo_a, lbo_b, sbo_b, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("probe variant %d: kernel failed: %s\n", variant, cudaGetErrorString(e));
      return -1;
    }
    CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 32; ++n) {
        double ref = 0;
        for (int k = 0; k < 8; ++k) ref += (double)a[m * 8 + k] * b[n * 8 + k];
        maxerr = fmax(maxerr, fabs(ref - d[m * 32 + n]));
      }
    printf("probe variant %d (%s): max abs err %.3g -> %s\n", variant,
           variant == 0 ? "LBO = K-chunk stride, SBO = 8-row-group stride" : "swapped", maxerr,
           maxerr == 0 ? "MATCH" : "mismatch");
    if (maxerr == 0) ok_any |= 1 << variant;
  }
  {  // M=64: print the TMEM lane each accumulator row landed in
    CK(cudaMemset(dd, 0, d.size() * 4));
    umma_probe_kernel<<<1, 128>>>(da, db, dd, 2112, 128, 512, 128, 2112, 128, 512, 128, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("probe M=64: kernel failed: %s\n", cudaGetErrorString(e));
      return -1;
    }
    CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
    printf("probe M=64 row->lane:");
    for (int m = 0; m < 64; ++m) {
      int found = -1;
      for (int l = 0; l < 128 && found < 0; ++l) {
        bool same = true, nz = false;
        for (int n = 0; n < 32; ++n) {
          double ref = 0;
          for (int k = 0; k < 8; ++k) ref += (double)a[m * 8 + k] * b[n * 8 + k];
          same = same && ref == d[l * 32 + n];
          nz = nz || ref != 0;
        }
        if (same && nz) found = l;
      }
      printf(" %d", found);
    }
    int nzl = 0;
    for (int l = 0; l < 128; ++l) {
      bool nz = false;
      for (int n = 0; n < 32; ++n) nz = nz || d[l * 32 + n] != 0;
      nzl += nz;
    }
    printf("\nprobe M=64: %d non-zero lanes\n", nzl);
  }
  cudaFree(da);
  cudaFree(db);
  cudaFree(dd);
  return ok_any;
}

// ------------------------------------------------------------------------------------------------ conv checks
static double check3x3(int N, int C, int K, int relu, int padded, const std::vector<int>& imgs) {
  std::vector<float> x((size_t)N * 256 * C), w((size_t)K * C * 9), sc(K), sh(K);
  for (auto& v : x) v = frand();
  for (auto& v : w) v = frand();
  for (auto& v : sc) v = frand() * 0.8f;
  for (auto& v : sh) v = frand();
  const int W = padded ? 16 : 14, o = padded ? 1 : 0;
  std::vector<float> y((size_t)N * W * W * K, 123.f);
  wg_layer_t* L = nullptr;
  int rc = wg_conv3x3_create(&L, C, K, w.data(), sc.data(), sh.data(), relu, g_dtype, 0);
  if (rc) {
    printf("create3x3 failed: %s (%s)\n", wg_strerror(rc), wg_last_cuda_error());
    return 1e30;
  }
  float *dx, *dy;
  CK(cudaMalloc(&dx, x.size() * 4));
  CK(cudaMalloc(&dy, y.size() * 4));
  CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy, y.data(), y.size() * 4, cudaMemcpyHostToDevice));
  rc = wg_run(L, dx, dy, N, padded, nullptr);
  if (rc) {
    printf("wg_run failed: %s (%s)\n", wg_strerror(rc), wg_last_cuda_error());
    return 1e30;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("3x3 kernel failed: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  CK(cudaMemcpy(y.data(), dy, y.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  long long border_bad = 0;
  for (int n : imgs) {
    if (n >= N) continue;
    for (int oy = 0; oy < 14; ++oy)
      for (int ox = 0; ox < 14; ++ox)
        for (int k = 0; k < K; ++k) {
          double acc = 0;
          for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) {
              const float* xp = &x[(((size_t)n * 16 + oy + r) * 16 + ox + s) * C];
              const float* wp = &w[((size_t)k * C) * 9 + r * 3 + s];
              for (int c = 0; c < C; ++c) acc += (double)xp[c] * wp[(size_t)c * 9];
            }
          double ref = sc[k] * acc + sh[k];
          if (relu && ref < 0) ref = 0;
          const double got = y[(((size_t)n * W + oy + o) * W + ox + o) * K + k];
          maxerr = fmax(maxerr, fabs(ref - got));
          maxref = fmax(maxref, fabs(ref));
        }
    if (padded)
      for (int py = 0; py < 16; ++py)
        for (int px = 0; px < 16; ++px)
          if (py == 0 || py == 15 || px == 0 || px == 15)
            for (int k = 0; k < K; ++k)
              if (y[(((size_t)n * 16 + py) * 16 + px) * K + k] != 0.f) ++border_bad;
  }
  wg_destroy(L);
  cudaFree(dx);
  cudaFree(dy);
  printf("3x3 N=%d C=%d K=%d relu=%d padded=%d: max|err|=%.3g max|ref|=%.3g rel=%.3g border_bad=%lld\n", N, C, K, relu,
         padded, maxerr, maxref, maxerr / maxref, border_bad);
  return border_bad ? 1e30 : maxerr / maxref;
}

static double check1x1(int N, int Cin, int Cout, int relu, const std::vector<int>& imgs) {
  std::vector<float> x((size_t)N * 196 * Cin), w((size_t)Cin * Cout), sc(Cout), sh(Cout);
  for (auto& v : x) v = frand() * 40.f;
  for (auto& v : w) v = frand() * 40.f;
  for (auto& v : sc) v = frand() * 8.f;
  for (auto& v : sh) v = frand() * 40.f;
  std::vector<float> y((size_t)N * 196 * Cout, 123.f);
  wg_layer_t* L = nullptr;
  int rc = wg_conv1x1_create(&L, Cin, Cout, w.data(), sc.data(), sh.data(), relu, WG_TF32, 0);
  if (rc) {
    printf("create1x1 failed: %s (%s)\n", wg_strerror(rc), wg_last_cuda_error());
    return 1e30;
  }
  rc = wg_run_host(L, x.data(), y.data(), N, 0);
  if (rc) {
    printf("wg_run_host failed: %s (%s)\n", wg_strerror(rc), wg_last_cuda_error());
    exit(3);
  }
  double maxerr = 0, maxref = 0;
  for (int n : imgs) {
    if (n >= N) continue;
    for (int p = 0; p < 196; ++p)
      for (int k = 0; k < Cout; ++k) {
        double acc = 0;
        const float* xp = &x[((size_t)n * 196 + p) * Cin];
        for (int c = 0; c < Cin; ++c) acc += (double)xp[c] * w[(size_t)c * Cout + k];
        double ref = sc[k] * acc + sh[k];
        if (relu && ref < 0) ref = 0;
        const double got = y[((size_t)n * 196 + p) * Cout + k];
        maxerr = fmax(maxerr, fabs(ref - got));
        maxref = fmax(maxref, fabs(ref));
      }
  }
  wg_destroy(L);
  printf("1x1 N=%d %d->%d relu=%d: max|err|=%.3g max|ref|=%.3g rel=%.3g\n", N, Cin, Cout, relu, maxerr, maxref,
         maxerr / maxref);
  return maxerr / maxref;
}

static void time_layer(int kind, int N, int C, int K, int relu) {
  std::vector<float> w((size_t)K * C * (kind == 0 ? 9 : 1)), sc(K, 1.f), sh(K, 0.f);
  for (auto& v : w) v = frand();
  wg_layer_t* L = nullptr;
  int rc = kind == 0 ? wg_conv3x3_create(&L, C, K, w.data(), sc.data(), sh.data(), relu, g_dtype, 0)
                     : wg_conv1x1_create(&L, C, K, w.data(), sc.data(), sh.data(), relu, WG_TF32, 0);
  if (rc) return;
  const size_t xe = (size_t)N * (kind == 0 ? 256 : 196) * C, ye = (size_t)N * 196 * K;
  float *dx, *dy;
  CK(cudaMalloc(&dx, xe * 4));
  CK(cudaMalloc(&dy, ye * 4));
  CK(cudaMemset(dx, 0, xe * 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) wg_run(L, dx, dy, N, 0, nullptr);
  CK(cudaDeviceSynchronize());
  const int iters = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) wg_run(L, dx, dy, N, 0, nullptr);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1000.0 / iters;
  const double flops = 2.0 * N * 196 * C * K * (kind == 0 ? 9 : 1);
  printf("time %s N=%d %d->%d: %.2f us/launch (L2-warm, back-to-back), %.1f TFLOP/s direct-equivalent\n",
         kind == 0 ? "3x3" : "1x1", N, C, K, us, flops / us * 1e-6);
  wg_destroy(L);
  cudaFree(dx);
  cudaFree(dy);
}

// ------------------------------------------------------------------------------------------------ TMA latency probe
// What bounds a small-batch layer: how long after a kernel starts do `total` bytes per CTA, requested as `pieces`
// concurrent 1-D bulk copies from an L2-warm buffer, take to land in shared memory? Thread 0 of every CTA issues the
// copies at once and times issue -> mbarrier completion with clock64; the host prints min / median / max over CTAs.
__global__ void tma_probe_kernel(const uint8_t* src, long long* out, int total, int pieces, int distinct) {
  using namespace wg;
  extern __shared__ __align__(1024) uint8_t dyn[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* s = src + (distinct ? (size_t)blockIdx.x * total : 0);
    const int piece = total / pieces;
    const long long t0 = clock64();
    mbar_arrive_expect_tx(&bar, total);
    for (int p = 0; p < pieces; ++p) tma_bulk_g2s(dyn + p * piece, s + (size_t)p * piece, piece, &bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
}

static void time_tma(int grid, int total, int pieces, int distinct) {
  static uint8_t* src = nullptr;
  static long long* out = nullptr;
  const size_t bytes = (size_t)148 * 192 * 1024;
  if (!src) {
    CK(cudaMalloc(&src, bytes));
    CK(cudaMemset(src, 1, bytes));
    CK(cudaMalloc(&out, 148 * sizeof(long long)));
    CK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  std::vector<long long> h(grid), best(grid, 1ll << 60);
  for (int rep = 0; rep < 5; ++rep) {  // first rep warms L2 / TLBs; keep the per-CTA minimum of the rest
    tma_probe_kernel<<<grid, 32, total, 0>>>(src, out, total, pieces, distinct);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    if (rep > 0)
      for (int i = 0; i < grid; ++i) best[i] = h[i] < best[i] ? h[i] : best[i];
  }
  std::vector<long long> s = best;
  for (size_t i = 0; i < s.size(); ++i)
    for (size_t j = i + 1; j < s.size(); ++j)
      if (s[j] < s[i]) std::swap(s[i], s[j]);
  printf("tma grid=%3d total=%6d B in %2d pieces %s: clk min %lld median %lld max %lld  -> %.1f B/clk/SM at the median\n",
         grid, total, pieces, distinct ? "distinct" : "shared  ", s[0], s[grid / 2], s[grid - 1],
         (double)total / s[grid / 2]);
}

// ------------------------------------------------------------------------------------------------ TMA tensor-load probe
// How fast does one SM's TMA unit deliver the 3x3 kernels' raw-tile boxes? x[N][16][16][C] viewed as
// (c, x/2, x&1, n*16+y); box = (cb channels, 8, 2, rows): the inner run is only cb*4 bytes. `count` loads, `depth` in
// flight, every CTA walks its own images; clk per load from thread 0 of each CTA.
#include <cuda.h>
__global__ void tma_tensor_probe_kernel(const __grid_constant__ CUtensorMap tmap, long long* out, int count, int depth,
                                        int box_bytes, int n_img, int C, int cb) {
  using namespace wg;
  extern __shared__ __align__(1024) uint8_t dyn[];
  __shared__ uint64_t bar[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x < 32 && elect_one()) {
    const long long t0 = clock64();
    int issued = 0, done = 0;
    uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    while (done < count) {
      while (issued < count && issued - done < depth) {
        const int b = issued % depth;
        const int img = (blockIdx.x * 7 + issued / (C / cb)) % (n_img - 3);
        mbar_arrive_expect_tx(&bar[b], box_bytes);
        tma_tensor_4d_g2s(dyn + b * 32768, &tmap, (issued % (C / cb)) * cb, 0, 0, img * 16, &bar[b]);
        ++issued;
      }
      const int b = done % depth;
      mbar_wait(&bar[b], ph[b]);
      ph[b] ^= 1;
      ++done;
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

static void time_tma_tensor(int grid, int cb, int rows, int swz, int depth) {
  const int n_img = 256, C = 256, count = 64;
  static float* x = nullptr;
  static long long* out = nullptr;
  if (!x) {
    CK(cudaMalloc(&x, (size_t)n_img * 256 * C * 4));
    CK(cudaMemset(x, 0, (size_t)n_img * 256 * C * 4));
    CK(cudaMalloc(&out, 148 * sizeof(long long)));
    CK(cudaFuncSetAttribute(tma_tensor_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768));
  }
  typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  CUtensorMap tmap;
  cuuint64_t dims[4] = {(cuuint64_t)C, 8, 2, (cuuint64_t)n_img * 16};
  cuuint64_t strides[3] = {(cuuint64_t)2 * C * 4, (cuuint64_t)C * 4, (cuuint64_t)16 * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)cb, 8, 2, (cuuint32_t)rows};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                : swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
  if (reinterpret_cast<PFN>(fp)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("tensor map encode failed (cb=%d swz=%d)\n", cb, swz);
    return;
  }
  const int box_bytes = cb * 4 * 16 * rows;
  std::vector<long long> h(grid);
  long long best = 1ll << 60, worst = 0;
  for (int rep = 0; rep < 3; ++rep) {
    tma_tensor_probe_kernel<<<grid, 32, 6 * 32768>>>(tmap, out, count, depth, box_bytes, n_img, C, cb);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  for (long long v : h) {
    best = v < best ? v : best;
    worst = v > worst ? v : worst;
  }
  printf("tma tensor grid=%3d box=(%2d ch x 16 px x %2d rows = %5d B) swizzle=%3d depth=%d: %.0f .. %.0f clk/load -> "
         "%.1f B/clk/SM\n",
         grid, cb, rows, box_bytes, swz, depth, (double)best / count, (double)worst / count,
         (double)box_bytes * count / worst);
}

// ------------------------------------------------------------------------------------------------ MMA issue probe
// Time `count` back-to-back tcgen05.mma kind::tf32 (K=8, operands from shared memory, no-swizzle K-major layout as in
// the 3x3 kernels) from first issue to commit completion. acc_stride = TMEM column distance between consecutive MMAs'
// accumulators (0 = all into the same accumulator).
// MODE 0: `if (threadIdx.x == 0)` around the issue loop; 1: `if (elect_one())`; 2: the whole warp runs the loop
// (uniform control flow and operands) and only the tcgen05.mma itself is predicated on an elected lane.
template <int M, int N, int MODE>
__global__ void mma_probe_kernel(long long* out, int count, int acc_stride, int distinct_ops) {
  using namespace wg;
  extern __shared__ __align__(1024) uint8_t dyn[];  // A images then B images, zero-filled
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(dyn)[i] = 0;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tptr;
  const uint32_t a0 = smem_u32(dyn), b0 = a0 + 96 * 1024;
  constexpr uint32_t kALbo = M * 16 + 64, kAPer = kALbo + M * 16, kBLbo = N * 16, kBPer = 2 * N * 16;
  constexpr uint32_t idesc = make_idesc(kFmtTF32, M, N);
  if (threadIdx.x < 32) {
    long long t0 = 0, t1 = 0;
    if constexpr (MODE == 2) {
      t0 = clock64();
      for (int i = 0; i < count; ++i) {
        const int p = distinct_ops ? (i & 15) : 0;
        const uint64_t ad = make_smem_desc(a0 + p * kAPer, kALbo, 128, kLayoutNone);
        const uint64_t bd = make_smem_desc(b0 + p * kBPer, kBLbo, 128, kLayoutNone);
        if (elect_one()) umma_tf32_ss(tb + ((i * acc_stride) & 511 & ~(N - 1)), ad, bd, idesc, 1u);
      }
      t1 = clock64();
      if (elect_one()) umma_commit(&bar);
    } else if (MODE == 1 ? elect_one() : threadIdx.x == 0) {
      t0 = clock64();
      for (int i = 0; i < count; ++i) {
        const int p = distinct_ops ? (i & 15) : 0;
        const uint64_t ad = make_smem_desc(a0 + p * kAPer, kALbo, 128, kLayoutNone);
        const uint64_t bd = make_smem_desc(b0 + p * kBPer, kBLbo, 128, kLayoutNone);
        umma_tf32_ss(tb + ((i * acc_stride) & 511 & ~(N - 1)), ad, bd, idesc, 1u);
      }
      t1 = clock64();
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (threadIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tb);
}

template <int M, int N, int MODE>
static void time_mma(int count, int acc_stride, int distinct_ops) {
  static long long* out = nullptr;
  if (!out) CK(cudaMalloc(&out, 2 * sizeof(long long)));
  CK(cudaFuncSetAttribute(mma_probe_kernel<M, N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  long long h[2] = {0, 0};
  for (int rep = 0; rep < 2; ++rep) {
    mma_probe_kernel<M, N, MODE><<<1, 128, 160 * 1024>>>(out, count, acc_stride, distinct_ops);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
  printf("mma mode %d M=%3d N=%3d K=8 tf32 x%3d acc_stride=%3d %s operands: issue %5lld clk, issue->complete %5lld clk = %.1f clk/MMA\n",
         MODE, M, N, count, acc_stride, distinct_ops ? "16 distinct" : "same       ", h[0], h[1], (double)h[1] / count);
}

// ------------------------------------------------------------------------------------------------ A-from-TMEM probe
// tcgen05.mma kind::tf32 with the A operand in TMEM: A[128][8] written with tcgen05.st (thread = row, 8 columns),
// B[48][8] in shared memory (no-swizzle K-major), D = A*B into columns 0..47, D2 = (-A)*B (negate-A bit) into columns
// 64..111. Then `count` such MMAs back to back for the issue -> complete rate.
__global__ void umma_ts_probe_kernel(const float* a, const float* b, float* d, long long* clk, int count) {
  using namespace wg;
  __shared__ __align__(1024) uint8_t sb[2 * 48 * 16];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 8; i += blockDim.x) {
    const int r = i / 8, k = i % 8;
    *reinterpret_cast<float*>(sb + (k / 4) * (48 * 16) + r * 16 + (k % 4) * 4) = b[i];
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc<512>(&tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tptr;
  const uint32_t a_col = 256;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  tmem_st_x4(tb + lane_base + a_col, a[tid * 8 + 0], a[tid * 8 + 1], a[tid * 8 + 2], a[tid * 8 + 3]);
  tmem_st_x4(tb + lane_base + a_col + 4, a[tid * 8 + 4], a[tid * 8 + 5], a[tid * 8 + 6], a[tid * 8 + 7]);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long t0 = 0, t1 = 0;
  if (tid < 32) {
    if (elect_one()) {
      const uint64_t bd = make_smem_desc(smem_u32(sb), 48 * 16, 128, kLayoutNone);
      umma_tf32_ts(tb, tb + a_col, bd, make_idesc(kFmtTF32, 128, 48), 0);
      umma_tf32_ts(tb + 64, tb + a_col, bd, make_idesc(kFmtTF32, 128, 48, 1), 0);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[16];
  for (int c0 = 0; c0 < 112; c0 += 16) {
    tmem_ld_x16(tb + lane_base + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d[tid * 112 + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid < 32) {
    if (elect_one()) {
      const uint64_t bd = make_smem_desc(smem_u32(sb), 48 * 16, 128, kLayoutNone);
      t0 = clock64();
      for (int i = 0; i < count; ++i)
        umma_tf32_ts(tb + (i & 3) * 48, tb + a_col, bd, make_idesc(kFmtTF32, 128, 48), 1u);
      t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 1);
      clk[0] = t1 - t0;
      clk[1] = clock64() - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tb);
}

// 16-bit operands (kind::f16, bf16 here): A[128][16] packed two per TMEM column (k = 2c in the low half), B[48][16] in
// shared memory (K-major no-swizzle: 16-byte chunks of 8 elements), D = A*B.
__global__ void umma_ts16_probe_kernel(const float* a, const float* b, float* d) {
  using namespace wg;
  __shared__ __align__(1024) uint8_t sb[2 * 48 * 16];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 16; i += blockDim.x) {
    const int r = i / 16, k = i % 16;
    *reinterpret_cast<__nv_bfloat16*>(sb + (k / 8) * (48 * 16) + r * 16 + (k % 8) * 2) = __float2bfloat16_rn(b[i]);
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc<512>(&tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tptr, a_col = 256;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  float pk[8];
  for (int c = 0; c < 8; ++c) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a[tid * 16 + 2 * c], a[tid * 16 + 2 * c + 1]);  // .x = low half
    pk[c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h2));
  }
  tmem_st_x4(tb + lane_base + a_col, pk[0], pk[1], pk[2], pk[3]);
  tmem_st_x4(tb + lane_base + a_col + 4, pk[4], pk[5], pk[6], pk[7]);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid < 32) {
    if (elect_one()) {
      const uint64_t bd = make_smem_desc(smem_u32(sb), 48 * 16, 128, kLayoutNone);
      umma_f16_ts(tb, tb + a_col, bd, make_idesc(kFmtBF16, 128, 48), 0);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[16];
  for (int c0 = 0; c0 < 48; c0 += 16) {
    tmem_ld_x16(tb + lane_base + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d[tid * 48 + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tb);
}

static void run_ts16_probe() {
  std::vector<float> a(128 * 16), b(48 * 16), d(128 * 48);
  for (auto& v : a) v = roundf(frand() * 16.f);  // small integers: exact in bf16
  for (auto& v : b) v = roundf(frand() * 16.f);
  float *da, *db, *dd;
  CK(cudaMalloc(&da, a.size() * 4));
  CK(cudaMalloc(&db, b.size() * 4));
  CK(cudaMalloc(&dd, d.size() * 4));
  CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  umma_ts16_probe_kernel<<<1, 128>>>(da, db, dd);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("TS16 probe: kernel failed: %s\n", cudaGetErrorString(e));
    return;
  }
  CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
  double e1 = 0, e2 = 0;  // e2: hypothesis "k = 2c in the HIGH half"
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 48; ++n) {
      double r1 = 0, r2 = 0;
      for (int k = 0; k < 16; ++k) {
        r1 += (double)a[m * 16 + k] * b[n * 16 + k];
        r2 += (double)a[m * 16 + (k ^ 1)] * b[n * 16 + k];
      }
      e1 = fmax(e1, fabs(r1 - d[m * 48 + n]));
      e2 = fmax(e2, fabs(r2 - d[m * 48 + n]));
    }
  printf("TS16 probe (bf16 A from TMEM, M=128 N=48 K=16): max abs err %.3g with k=2c in the low half, %.3g with the halves "
         "swapped -> %s\n", e1, e2, e1 == 0 ? "low-half-first MATCH" : (e2 == 0 ? "high-half-first MATCH" : "mismatch"));
}

static void run_ts_probe() {
  std::vector<float> a(128 * 8), b(48 * 8), d(128 * 112);
  for (auto& v : a) v = roundf(frand() * 16.f);
  for (auto& v : b) v = roundf(frand() * 16.f);
  float *da, *db, *dd;
  long long* dc;
  CK(cudaMalloc(&da, a.size() * 4));
  CK(cudaMalloc(&db, b.size() * 4));
  CK(cudaMalloc(&dd, d.size() * 4));
  CK(cudaMalloc(&dc, 16));
  CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  umma_ts_probe_kernel<<<1, 128>>>(da, db, dd, dc, 96);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("TS probe: kernel failed: %s\n", cudaGetErrorString(e));
    return;
  }
  long long clk[2];
  CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(clk, dc, 16, cudaMemcpyDeviceToHost));
  double e1 = 0, e2 = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 48; ++n) {
      double ref = 0;
      for (int k = 0; k < 8; ++k) ref += (double)a[m * 8 + k] * b[n * 8 + k];
      e1 = fmax(e1, fabs(ref - d[m * 112 + n]));
      e2 = fmax(e2, fabs(-ref - d[m * 112 + 64 + n]));
    }
  printf("TS probe (A from TMEM, M=128 N=48 K=8 tf32): max abs err %.3g, negate-A %.3g -> %s; 96 MMAs: issue %lld clk, "
         "issue->complete %lld clk = %.1f clk/MMA\n",
         e1, e2, (e1 == 0 && e2 == 0) ? "MATCH" : "mismatch", clk[0], clk[1], clk[1] / 96.0);
}


// ------------------------------------------------------------------------------------------------ A-from-TMEM MMA rate
// Issue -> complete rate of `count` back-to-back tcgen05.mma with A in TMEM (M=128, K=8 tf32 or K=16 bf16) for a range of
// N, rotating over as many accumulators as fit next to the A columns, B = 16 distinct K-major no-swizzle operands in
// shared memory ([2 k-chunks][N][16 B] each), as the full-fold 3x3 kernel issues them.
template <int N, bool H16>
__global__ void umma_ts_rate_kernel(long long* clk, int count) {
  using namespace wg;
  extern __shared__ __align__(1024) uint8_t sb[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16 * 2 * N * 4; i += blockDim.x) reinterpret_cast<float*>(sb)[i] = 0.f;
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc<512>(&tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tptr;
  const uint32_t a_col = 448;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < 64; c += 4) tmem_st_x4(tb + lane_base + a_col + c, 0.f, 0.f, 0.f, 0.f);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr int kAcc = (448 / N) < 4 ? (448 / N) : 4;
  if (tid < 32) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(H16 ? kFmtBF16 : kFmtTF32, 128, N);
      const long long t0 = clock64();
      for (int i = 0; i < count; ++i) {
        const uint64_t bd = make_smem_desc(smem_u32(sb) + (i & 15) * (2 * N * 16), N * 16, 128, kLayoutNone);
        if constexpr (H16) umma_f16_ts(tb + (i % kAcc) * N, tb + a_col + (i & 7) * 8, bd, idesc, 1u);
        else umma_tf32_ts(tb + (i % kAcc) * N, tb + a_col + (i & 7) * 8, bd, idesc, 1u);
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      clk[0] = t1 - t0;
      clk[1] = clock64() - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(tb);
}

template <int N, bool H16>
static void time_ts_rate(int count) {
  long long* dc;
  CK(cudaMalloc(&dc, 16));
  const int smem = 16 * 2 * N * 16;
  CK(cudaFuncSetAttribute(umma_ts_rate_kernel<N, H16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) {
    umma_ts_rate_kernel<N, H16><<<1, 128, smem>>>(dc, count);
    CK(cudaDeviceSynchronize());
  }
  long long clk[2];
  CK(cudaMemcpy(clk, dc, 16, cudaMemcpyDeviceToHost));
  printf("TS rate M=128 N=%3d K=%2d %s x%3d: issue %5lld clk, issue->complete %5lld clk = %.1f clk/MMA (tensor floor %.1f)\n",
         N, H16 ? 16 : 8, H16 ? "bf16" : "tf32", count, clk[0], clk[1], (double)clk[1] / count, 128.0 * N * 8 / 2048.0 / 8);
  cudaFree(dc);
}

// ------------------------------------------------------------------------------------------------ patch-load probe
// Shared-memory cost of the transform warps' 16 LDS.128 per stage under three address patterns: 0 = one 16-byte chunk
// per lane, consecutive (the conflict-free reference), 1 = the full-fold kernel's parity-plane layout (9-slot pitch,
// tiles right-to-left), 2 = the single-box layout of the TM kernel (8-slot pitch). 8 warps as in the kernels.
__global__ void lds_probe_kernel(long long* out, int pattern, int iters) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 28 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = (float)i;
  __syncthreads();
  const int quad = warp & 3, cq = warp >> 2, row = quad * 32 + lane;
  const int t = row % 49, n = row / 49, ty = t / 7;
  uint32_t ad[16];
  for (int dy = 0; dy < 4; ++dy)
    for (int dx = 0; dx < 4; ++dx) {
      uint32_t a;
      if (pattern == 0) {
        a = (uint32_t)((dy * 4 + dx) * 1024 + tid * 16) % (24 * 1024);
      } else if (pattern == 1) {
        const int tx = 6 - t % 7;
        const uint32_t sl = (uint32_t)(((n * 16 + 2 * ty) >> 1) * 9 + tx + 1 + 9 * (dy >> 1) + (dx >> 1));
        a = ((dy & 1) * 2 + (dx & 1)) * 6912 + sl * 32 + ((cq ^ ((sl >> 2) & 1)) * 16);
      } else {
        const int tx = t % 7;
        const int x2 = tx + (dx >> 1);
        a = (uint32_t)((n * 16 + 2 * ty + dy) * 512 + (dx & 1) * 256 + x2 * 32 + ((cq ^ ((x2 >> 2) & 1)) * 16));
      }
      ad[dy * 4 + dx] = wg::smem_u32(sm) + a;
    }
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float4 v;  // "memory" clobber: keeps the loop-invariant loads inside the loop
      asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ad[k]) : "memory");
      acc += v.x + v.y + v.z + v.w;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) out[0] = t1 - t0;
  if (acc == 123.456f) out[1] = 1;
}

static void time_lds(int pattern) {
  long long* dc;
  CK(cudaMalloc(&dc, 16));
  CK(cudaFuncSetAttribute(lds_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 28 * 1024));
  const int iters = 200;
  for (int rep = 0; rep < 2; ++rep) {
    lds_probe_kernel<<<1, 256, 28 * 1024>>>(dc, pattern, iters);
    CK(cudaDeviceSynchronize());
  }
  long long clk = 0;
  CK(cudaMemcpy(&clk, dc, 8, cudaMemcpyDeviceToHost));
  printf("LDS.128 probe pattern %d: %lld clk for %d x 16 loads x 8 warps = %.2f clk per warp-level LDS.128 (4.00 = "
         "conflict-free)\n", pattern, clk, iters, (double)clk / (iters * 16 * 8));
  cudaFree(dc);
}

// ------------------------------------------------------------------------------------------------ launch floor
// What an N=1 layer cannot go below: empty kernels launched back to back the way the product launches its own
// (dynamic smem opt-in, 128-byte __grid_constant__ parameter, optional cluster), timed with the same event loop.
struct Blob128 {
  uint8_t b[128];
};
// spin_clk: busy body of that many clocks (stands in for a layer); pdl: 1 = griddepcontrol.launch_dependents at entry
// and griddepcontrol.wait before the body (programmatic dependent launch), 2 = wait first, trigger after the body.
__global__ void floor_kernel(const __grid_constant__ Blob128 blob, float* out, int touch_tmem, int spin_clk, int pdl) {
  extern __shared__ __align__(16) uint8_t dyn[];
  __shared__ uint32_t slot;
  if (pdl == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (touch_tmem) {
    if (threadIdx.x < 32) wg::tmem_alloc<512>(&slot);
    __syncthreads();
  }
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long t0 = clock64();
  while (clock64() - t0 < spin_clk) {
  }
  if (pdl == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (touch_tmem) {
    __syncthreads();
    if (threadIdx.x < 32) wg::tmem_dealloc<512>(slot);
  }
  if (out != nullptr && threadIdx.x == 0 && blob.b[0] == 77) out[blockIdx.x] = dyn[0];
}

static void time_floor(int grid, int threads, int smem, int cluster, int tmem, int spin_clk = 0, int pdl = 0,
                       int graph = 0) {
  CK(cudaFuncSetAttribute(floor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(floor_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  Blob128 blob;
  memset(&blob, 0, sizeof(blob));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (cluster > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = cluster;
    at[na].val.clusterDim.y = at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cfg.stream = st;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20;
  cudaGraphExec_t gx = nullptr;
  if (graph) {
    cudaGraph_t g;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < iters; ++i) CK(cudaLaunchKernelEx(&cfg, floor_kernel, blob, (float*)nullptr, tmem, spin_clk, pdl));
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&gx, g, 0));
    CK(cudaGraphLaunch(gx, st));
  } else {
    for (int i = 0; i < 3; ++i) CK(cudaLaunchKernelEx(&cfg, floor_kernel, blob, (float*)nullptr, tmem, spin_clk, pdl));
  }
  CK(cudaStreamSynchronize(st));
  cudaEventRecord(e0, st);
  if (graph) {
    CK(cudaGraphLaunch(gx, st));
  } else {
    for (int i = 0; i < iters; ++i) cudaLaunchKernelEx(&cfg, floor_kernel, blob, (float*)nullptr, tmem, spin_clk, pdl);
  }
  cudaEventRecord(e1, st);
  CK(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("floor grid=%3d threads=%d smem=%6d cluster=%d tmem=%d spin=%5d clk pdl=%d graph=%d: %.2f us/launch\n", grid,
         threads, smem, cluster, tmem, spin_clk, pdl, graph, ms * 1000.0 / iters);
  cudaStreamDestroy(st);
}

int main(int argc, char** argv) {
  const bool quick = argc > 1 && !strcmp(argv[1], "quick");
  printf("sm_100 devices: %d\n", wg_device_count());
  if (wg_device_count() == 0) return 1;
  if (argc > 1 && !strcmp(argv[1], "ts")) {
    run_ts_probe();
    run_ts16_probe();
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "lds")) {
    for (int p = 0; p < 3; ++p) time_lds(p);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "tsrate")) {
    time_ts_rate<32, false>(144);
    time_ts_rate<48, false>(144);
    time_ts_rate<64, false>(144);
    time_ts_rate<96, false>(144);
    time_ts_rate<128, false>(144);
    time_ts_rate<192, false>(144);
    time_ts_rate<256, false>(144);
    time_ts_rate<48, true>(144);
    time_ts_rate<64, true>(144);
    time_ts_rate<96, true>(144);
    time_ts_rate<128, true>(144);
    time_ts_rate<256, true>(144);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "mma")) {
    for (int count : {1, 16, 96}) {
      time_mma<64, 32, 0>(count, 32, 1);
      time_mma<64, 32, 1>(count, 32, 1);
      time_mma<64, 32, 2>(count, 32, 1);
      time_mma<128, 64, 0>(count, 64, 1);
      time_mma<128, 64, 1>(count, 64, 1);
      time_mma<128, 64, 2>(count, 64, 1);
      time_mma<128, 128, 2>(count, 128, 1);
    }
    time_mma<128, 64, 2>(96, 0, 1);
    time_mma<128, 64, 2>(96, 64, 0);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "tmat")) {
    for (int grid : {1, 148}) {
      for (int depth : {1, 3, 6}) {
        time_tma_tensor(grid, 8, 48, 0, depth);
        time_tma_tensor(grid, 8, 48, 32, depth);
        time_tma_tensor(grid, 16, 24, 0, depth);
        time_tma_tensor(grid, 16, 24, 64, depth);
        time_tma_tensor(grid, 32, 12, 0, depth);
        time_tma_tensor(grid, 32, 12, 128, depth);
        time_tma_tensor(grid, 64, 6, 0, depth);
      }
    }
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "tma")) {
    for (int grid : {1, 16, 64, 128})
      for (int total : {1024, 16384, 65536, 163840}) {
        time_tma(grid, total, 1, 1);
        if (total >= 16384) time_tma(grid, total, total / 4096, 1);
      }
    time_tma(128, 16384, 1, 0);
    time_tma(128, 65536, 4, 0);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "floor")) {
    time_floor(1, 32, 0, 1, 0);
    time_floor(148, 320, 0, 1, 0);
    time_floor(148, 320, 230000, 1, 0);
    time_floor(148, 320, 230000, 1, 1);
    time_floor(144, 320, 230000, 8, 0);
    time_floor(144, 320, 230000, 8, 1);
    time_floor(32, 192, 200000, 8, 1);
    time_floor(16, 192, 200000, 1, 1);
    for (int graph = 0; graph < 2; ++graph)
      for (int pdl = 0; pdl < 3; ++pdl) {
        time_floor(32, 320, 230000, 8, 1, 0, pdl, graph);
        time_floor(32, 320, 230000, 8, 1, 10000, pdl, graph);
        time_floor(148, 320, 230000, 1, 1, 10000, pdl, graph);
      }
    run_probe();
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "timen1")) {  // small-batch latency only
    for (int n : {1, 2, 4, 8}) {
      time_layer(0, n, 128, 128, 1);
      time_layer(0, n, 256, 256, 1);
    }
    time_layer(1, 1, 512, 128, 1);
    time_layer(1, 1, 128, 512, 0);
    time_layer(1, 1, 1024, 256, 1);
    time_layer(1, 1, 256, 1024, 0);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "time3x3")) {  // timing only (used with WG_DEBUG_ABLATE experiments)
    time_layer(0, 256, 128, 128, 1);
    time_layer(0, 256, 256, 256, 1);
    return 0;
  }
  int probe = run_probe();
  printf("probe result mask: %d\n", probe);
  int fails = 0;
  auto bad = [&](double rel, double tol) {
    if (!(rel <= tol)) ++fails;
  };
  for (int kn : {48, 64, 32}) {
    wg_dev_set_wino_kn(kn);
    printf("-- 3x3 variant KN=%d\n", kn);
    bad(check3x3(1, 128, 128, 1, 0, {0}), 1e-3);
    bad(check3x3(1, 128, 128, 1, 1, {0}), 1e-3);
    bad(check3x3(3, 64, 64, 0, 0, {0, 1, 2}), 1e-3);
    bad(check3x3(7, 32, 32, 1, 1, {0, 3, 6}), 1e-3);
    if (!quick) {
      bad(check3x3(256, 128, 128, 1, 0, {0, 131, 255}), 1e-3);
      time_layer(0, 256, 128, 128, 1);
      time_layer(0, 256, 256, 256, 1);
    }
  }
  wg_dev_set_wino_kn(96);
  g_dtype = WG_BF16;
  printf("-- 3x3 bf16 operand variant (tolerance 1e-2)\n");
  bad(check3x3(1, 128, 128, 1, 1, {0}), 1e-2);
  bad(check3x3(3, 64, 64, 0, 0, {0, 1, 2}), 1e-2);
  bad(check3x3(7, 32, 64, 1, 1, {0, 3, 6}), 1e-2);
  if (!quick) {
    bad(check3x3(256, 128, 128, 1, 0, {0, 131, 255}), 1e-2);
    bad(check3x3(64, 256, 256, 1, 0, {0, 63}), 1e-2);
    time_layer(0, 256, 128, 128, 1);
    time_layer(0, 256, 256, 256, 1);
  }
  g_dtype = WG_TF32;
  bad(check1x1(1, 512, 128, 1, {0}), 1e-3);
  bad(check1x1(1, 128, 512, 0, {0}), 1e-3);
  bad(check1x1(3, 64, 256, 0, {0, 2}), 1e-3);
  if (!quick) {
    bad(check3x3(1, 256, 256, 1, 1, {0}), 1e-3);
    bad(check3x3(256, 128, 128, 1, 0, {0, 131, 255}), 1e-3);
    bad(check3x3(64, 256, 256, 1, 0, {0, 63}), 1e-3);
    bad(check1x1(1, 1024, 256, 1, {0}), 1e-3);
    bad(check1x1(1, 256, 1024, 0, {0}), 1e-3);
    bad(check1x1(256, 512, 128, 1, {0, 255}), 1e-3);
    bad(check1x1(256, 256, 1024, 0, {0, 255}), 1e-3);
    time_layer(0, 1, 128, 128, 1);
    time_layer(0, 256, 128, 128, 1);
    time_layer(0, 1, 256, 256, 1);
    time_layer(0, 256, 256, 256, 1);
    time_layer(1, 1, 512, 128, 1);
    time_layer(1, 256, 512, 128, 1);
    time_layer(1, 256, 128, 512, 0);
    time_layer(1, 256, 1024, 256, 1);
    time_layer(1, 256, 256, 1024, 0);
  }
  printf("selftest: %d failing checks, %lld launches\n", fails, wg_launch_count());
  return fails ? 1 : 0;
}
