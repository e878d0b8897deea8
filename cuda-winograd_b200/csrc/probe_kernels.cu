// Measured denominators for the roofline: the tensor pipe's dense peak on THIS device, by running nothing but
// back-to-back tcgen05.mma on every SM ("measure the TF32 peak on the box").
//
// One CTA per SM, one elected thread issues `iters` MMAs of M=128, N=256 (K = 8 for kind::tf32, 16 for kind::f16) with
// both operands in shared memory, alternating between two 256-column accumulators so that consecutive MMAs do not
// serialise on the accumulator, one tcgen05.commit at the end. Operands are all-zero tiles: the probe measures issue
// rate and pipe throughput, not arithmetic. FLOPs = grid * iters * 2 * 128 * 256 * K, time = CUDA events around the
// launch (a few milliseconds per launch, i.e. the burst regime the per-launch kernel times are taken in).
#include "ptx.cuh"
#include "wg_internal.h"

namespace wg {

constexpr int kProbeN = 256;
constexpr uint32_t kProbeABytes = 128 * 32;      // 128 rows x 32 B (one K slice), K-major, no swizzle
constexpr uint32_t kProbeBBytes = kProbeN * 32;  // 256 rows x 32 B
constexpr uint32_t kProbeSmem = kProbeABytes + kProbeBBytes + 64;

template <bool F16>
__global__ void __launch_bounds__(128, 1) tensor_peak_probe_kernel(int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kProbeABytes + kProbeBBytes);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kProbeABytes + kProbeBBytes + 16);
  for (uint32_t i = threadIdx.x; i < (kProbeABytes + kProbeBBytes) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(F16 ? kFmtBF16 : kFmtTF32, 128, kProbeN);
      // canonical K-major no-swizzle layout: core matrix = 8 rows x 16 B; the two 16-byte K chunks of a row group are
      // LBO apart, 8-row groups SBO apart
      const uint64_t a_desc = make_smem_desc(smem_u32(smem), 128 * 16, 128, kLayoutNone);
      const uint64_t b_desc = make_smem_desc(smem_u32(smem + kProbeABytes), kProbeN * 16, 128, kLayoutNone);
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tmem_base + (uint32_t)(i & 1) * kProbeN;
        if constexpr (F16) umma_bf16_ss(d, a_desc, b_desc, idesc, i > 1 ? 1u : 0u);
        else umma_tf32_ss(d, a_desc, b_desc, idesc, i > 1 ? 1u : 0u);
      }
      umma_commit(bar);
      mbar_wait(bar, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace wg

extern "C" int wg_measure_tensor_peak(int device, int dtype, double* tflops_out, double* clk_per_mma_out) {
  using namespace wg;
  if (!tflops_out || (dtype != WG_TF32 && dtype != WG_BF16)) return WG_ERR_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return WG_ERR_NODEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10 || prop.minor != 0) return WG_ERR_NODEVICE;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  const bool f16 = dtype == WG_BF16;
  const int iters = 20000;  // ~1.3 ms at 128 clk per MMA
  const int grid = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto launch = [&](int n) {
    if (f16) tensor_peak_probe_kernel<true><<<grid, 128, kProbeSmem>>>(n);
    else tensor_peak_probe_kernel<false><<<grid, 128, kProbeSmem>>>(n);
  };
  launch(2000);  // warm-up
  float best_ms = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    launch(iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best_ms) best_ms = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const cudaError_t err = cudaGetLastError();
  cudaSetDevice(prev);
  if (err != cudaSuccess || best_ms > 1e29f) return WG_ERR_CUDA;
  const double flops = (double)grid * iters * 2.0 * 128 * kProbeN * (f16 ? 16 : 8);
  *tflops_out = flops / (best_ms * 1e-3) / 1e12;
  if (clk_per_mma_out) *clk_per_mma_out = best_ms * 1e-3 * (prop.clockRate * 1e3) / iters;
  return WG_OK;
}

#ifdef WG_DEV_BUILD
// Developer probe: how fast do `warps` warps (one per 32-lane TMEM quadrant, two per quadrant from 5 warps on) read TMEM
// with tcgen05.ld.32x32b.x32 -- `depth` loads in flight per tcgen05.wait::ld? Answers whether the 1x1 epilogue's
// ~350-450 clk per 32-column chunk is the TMEM read port (per SM), a per-warp latency, or something around it.
namespace wg {
__global__ void __launch_bounds__(256, 1) tmem_ld_probe_kernel(int iters, int depth, int with_fence, long long* out) {
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(16) float sink[256 * 4];
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float v[4][32];
    for (int d = 0; d < 4; ++d)
      if (d < depth) tmem_ld_x32(tbase + (uint32_t)(((i * 4 + d) * 32) & 511), v[d]);
    tmem_ld_wait();
    for (int d = 0; d < 4; ++d)
      if (d < depth)
#pragma unroll
        for (int j = 0; j < 32; ++j) asm volatile("" ::"f"(v[d][j]));  // consumed, no dependent arithmetic
    if (with_fence) {
      sink[threadIdx.x * 4] = v[0][0];
      fence_proxy_async_smem();
      __syncwarp();
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
  if (sink[(threadIdx.x * 4 + 4) & 1023] == 12345.678f) out[100] = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_ptr);
}
}  // namespace wg

extern "C" int wg_dev_tmem_ld_probe(int warps, int iters, int depth, int with_fence, long long* clk_out) {
  long long* d = nullptr;
  if (warps < 1 || warps > 8 || depth < 1 || depth > 4) return WG_ERR_ARG;
  if (cudaMalloc(&d, 128 * sizeof(long long)) != cudaSuccess) return WG_ERR_CUDA;
  cudaMemset(d, 0, 128 * sizeof(long long));
  wg::tmem_ld_probe_kernel<<<1, warps * 32>>>(iters, depth, with_fence, d);
  long long h[8] = {0};
  const cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return WG_ERR_CUDA;
  long long mx = 0;
  for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
  *clk_out = mx;
  return WG_OK;
}
#endif
