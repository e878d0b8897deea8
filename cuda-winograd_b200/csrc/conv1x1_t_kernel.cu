// 1x1 conv (GEMM) + folded BatchNorm (+ ReLU) (+ residual add) for the WIDE-COUT shapes, sm_100a: the GEMM laid out
// transposed -- M = output channels, N = pixels -- with the weight block resident in shared memory.
//
// Replaces kernel_128_one_512 (/root/reference/Kernel128_one.cu:244-273) and kernel_256_one_1024 (Kernel256_one.cu:
// 246-274) at throughput batch sizes; same operator and data layouts as conv1x1_bn_act_kernel (one_kernels.cu):
// y[M x Cout] = act(scale * (x[M x Cin] * W[Cin x Cout]) + shift (+ residual)), fp32 in / out, TF32 operands.
//
// Why another layout. The wide-Cout shapes are store-heavy (4x more output than input bytes) and were bound by the
// epilogue's instruction chain (profiles/one_ablation_r02.md: the kernel with all memory traffic and MMAs switched off
// still took half the time). With pixels on M every thread owns one pixel and 32 couts per chunk: 16 folded-BN vector
// loads per chunk, and four epilogue warps per 128 x 256 tile. With COUTS on M (TMEM lanes) a thread owns one cout --
// its scale / shift are two registers for the whole item -- and the 3x3 direct kernel's epilogue applies unchanged:
// eight warps, per 16 pixels one tcgen05.ld.x16, 16 FMAs, a transposing [16 px][32 couts] staging tile written one
// conflict-free 128-byte row per instruction, one 2 KB TMA tensor store.
//
// Work item = (256 pixels, 128 couts): D[128 couts][256 px] += W[128 couts][32 ch] . X[256 px][32 ch]^T, Cin/32 k-blocks
// of 4 MMAs (M = 128, N = 256, K = 8), two TMEM accumulator buffers. A CTA keeps ONE 128-cout block for its whole life:
// the [Cin/32][128][32] weight slab (<= 128 KB, Cin <= 256) is loaded once; the ring carries activations only.
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue (quad = 32 couts, half of the 256 pixels each).
// RES: the residual's [16 px][32 couts] sub-tile is TMA-loaded into the staging buffer one chunk ahead (plain layout, same
// box as the store); each thread adds its own column and the same buffer is stored.
#include <stdlib.h>

#include "ptx.cuh"
#include "wg_internal.h"

namespace wg {

constexpr int kOneTThreads = 32 * 10;

// CIN_MAX = 128: the [Cin/32][128][32] weight slab (<= 64 KB) is resident, the ring carries activations only.
// CIN_MAX = 256: a resident slab (128 KB) would leave two activation stages = 64 KB in flight, a third of what the
// latency x bandwidth product of the tiled-TMA path needs (measured: 84 us instead of 64); the weight block of a k-block
// then rides in the same ring stage as its activations (4 stages of 32 + 16 KB) and is re-read from L2 per item.
// PAIR: the two CTAs of a cluster form a tcgen05 cta_group::2 pair on one (256 pixels, 256 couts) item: M = 256 couts (128
// per CTA, each with its own weight blocks) and -- the point -- each CTA loads only HALF of the pixels (128 x 32 channels
// per k-block); the tensor core reads each half of the B operand from the shared memory it lives in. The activations are
// the stream every cout block re-reads (Cout / 128 times per layer): halving it per SM takes the kernel off the tiled-TMA
// ceiling (~37 B/clk per SM). Hand-offs as in the other pair kernels: the peer's MMA warp relays "landed" to the leader,
// commits are multicast, the peer's epilogue warps arrive on the leader's acc_empty.
template <int CIN_MAX, bool PAIR = false, bool RES = false>
struct OneTSmem {
  // resident slab: Cin <= 128 always; Cin = 256 with CTA pairs and no residual (16 KB activation stages: five of them +
  // the 128 KB slab fit when every epilogue warp has one staging tile instead of the two the residual prefetch needs)
  static constexpr bool kBigResident = CIN_MAX > 128 && PAIR && !RES;
  static constexpr bool kResident = CIN_MAX <= 128 || kBigResident;
  static constexpr int kOutBufs = kBigResident ? 1 : 2;
  static constexpr int kSX = PAIR ? (kBigResident ? 5 : 6) : 4;  // k-blocks in flight
  static constexpr uint32_t kXBytes = (PAIR ? 128 : 256) * 128;  // this CTA's pixels x 32 channels
  static constexpr uint32_t kWBytes = 128 * 128;         // one k-block of the slab: 128 couts x 32 channels
  static constexpr uint32_t kSlabBytes = (kResident ? CIN_MAX / 32 : kSX) * kWBytes;
  static constexpr uint32_t kStageOutBytes = 16 * 128;   // [16 px][32 couts]
  static constexpr uint32_t kOffX = 0;
  static constexpr uint32_t kOffW = kOffX + kSX * kXBytes;
  static constexpr uint32_t kOffOut = kOffW + kSlabBytes;  // [8 warps][kOutBufs buffers]
  static constexpr uint32_t kOffBar = kOffOut + 8 * kOutBufs * kStageOutBytes;
  static constexpr uint32_t kNumBars = 2 * kSX + 4 + 1 + 16;
  static constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr uint32_t kTotal = kOffTmemPtr + 16 + 1024;
  static_assert(kOffW % 1024 == 0, "swizzled buffers must be 1024-byte aligned");
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

// the residual / output sub-tile of (pixel chunk j of 16) is the same [16 px][32 couts] box
template <int CIN_MAX, bool RES, bool PAIR>
__global__ void __launch_bounds__(kOneTThreads, 1)
conv1x1_t_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
                 const __grid_constant__ CUtensorMap tmap_r, const float* __restrict__ w_img,
                 const float* __restrict__ scale, const float* __restrict__ shift, long long m_rows, int Cin, int Cout,
                 int relu, int bn_packed, int relu_after) {
  using S = OneTSmem<CIN_MAX, PAIR, RES>;
  constexpr uint16_t kPairMask = 0x3;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* x_full = bars;
  uint64_t* x_empty = x_full + S::kSX;
  uint64_t* acc_full = x_empty + S::kSX;  // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint64_t* w_full = acc_empty + 2;       // the resident weight slab has landed
  uint64_t* res_full = w_full + 1;        // RES: [8 warps][2 buffers]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::kOffTmemPtr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_y);
    const uint32_t n_full = (PAIR && crank == 0) ? 2 : 1;  // leader: own TMA bytes + the peer's relay
    for (int i = 0; i < S::kSX; ++i) mbar_init(&x_full[i], n_full), mbar_init(&x_empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], PAIR ? 16 : 8);
    mbar_init(w_full, n_full);
    for (int i = 0; i < 16; ++i) mbar_init(&res_full[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_cg2<512>(tmem_ptr);
    else tmem_alloc<512>(tmem_ptr);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_kb = Cin / 32;
  const int n_pt = (int)((m_rows + 255) / 256);  // 256-pixel tiles
  // Cluster (PAIR) / CTA b keeps cout group b % n_cg -- 256 couts per pair, 128 per CTA -- and walks the pixel tiles
  // b / n_cg, + grid / n_cg, ... (the host makes the grid a multiple of n_cg)
  constexpr int kCL = PAIR ? 2 : 1;
  const int n_cg = Cout / (128 * kCL);
  const int cl_id = (int)blockIdx.x / kCL, n_cl = (int)gridDim.x / kCL;
  const int cb = (cl_id % n_cg) * kCL + (int)crank;  // this CTA's 128-cout block
  const int first_pt = cl_id / n_cg, pt_stride = n_cl / n_cg;

  if (warp == 0) {
    if (elect_one()) {
      // the weight slab: 128 couts of the packed image [Cout/bn_packed][Cin/32][bn_packed rows][128 B] -- requested before
      // waiting for the previous kernel in the stream
      const int col0 = cb * 128;
      const uint8_t* w_src = reinterpret_cast<const uint8_t*>(w_img) +
                             ((size_t)(col0 / bn_packed) * n_kb * bn_packed + col0 % bn_packed) * 128;
      if constexpr (S::kResident) {
        mbar_arrive_expect_tx(w_full, (uint32_t)n_kb * S::kWBytes);
        for (int kb = 0; kb < n_kb; ++kb)
          tma_bulk_g2s(smem + S::kOffW + kb * S::kWBytes, w_src + (size_t)kb * bn_packed * 128, S::kWBytes, w_full);
      }
      pdl_wait();  // activations come from the previous kernel in the stream
      uint32_t sx = 0, px = 0;
      for (int pt = first_pt; pt < n_pt; pt += pt_stride)
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&x_empty[sx], px ^ 1);
          mbar_arrive_expect_tx(&x_full[sx], S::kXBytes + (S::kResident ? 0u : S::kWBytes));
          tma_tensor_2d_g2s(smem + S::kOffX + sx * S::kXBytes, &tmap_x, kb * 32, pt * 256 + (PAIR ? (int)crank * 128 : 0),
                            &x_full[sx]);
          if constexpr (!S::kResident)
            tma_bulk_g2s(smem + S::kOffW + sx * S::kWBytes, w_src + (size_t)kb * bn_packed * 128, S::kWBytes, &x_full[sx]);
          if (++sx == S::kSX) { sx = 0; px ^= 1; }
        }
    }
  } else if (warp == 1 && PAIR && crank != 0) {
    // peer of a pair: no MMAs to issue; relay "landed here" to the leader's barriers, in consumption order
    if (elect_one()) {
      uint32_t sx = 0, px = 0;
      if constexpr (S::kResident) {
        mbar_wait(w_full, 0);
        mbar_arrive_remote_plain(w_full, 0);
      }
      for (int pt = first_pt; pt < n_pt; pt += pt_stride)
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&x_full[sx], px);
          mbar_arrive_remote_plain(&x_full[sx], 0);
          if (++sx == S::kSX) { sx = 0; px ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(kFmtTF32, PAIR ? 256 : 128, 256);
      const uint32_t x_base = smem_u32(smem + S::kOffX);
      const uint32_t w_base = smem_u32(smem + S::kOffW);
      uint32_t sx = 0, px = 0, it = 0;
      if constexpr (S::kResident) mbar_wait(w_full, 0);
      for (int pt = first_pt; pt < n_pt; pt += pt_stride, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&x_full[sx], px);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a_desc =
                make_smem_desc(w_base + (S::kResident ? (uint32_t)kb : sx) * S::kWBytes + k * 32, 0, 1024, kLayoutSW128);
            const uint64_t b_desc = make_smem_desc(x_base + sx * S::kXBytes + k * 32, 0, 1024, kLayoutSW128);
            if constexpr (PAIR) umma_tf32_ss_cg2(tmem_base + buf * 256, a_desc, b_desc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else umma_tf32_ss(tmem_base + buf * 256, a_desc, b_desc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if constexpr (PAIR) umma_commit_mcast_cg2(&x_empty[sx], kPairMask);
          else umma_commit(&x_empty[sx]);
          if (++sx == S::kSX) { sx = 0; px ^= 1; }
        }
        if constexpr (PAIR) umma_commit_mcast_cg2(&acc_full[buf], kPairMask);
        else umma_commit(&acc_full[buf]);
      }
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3, hsel = ew >> 2;
    uint8_t* stage_out = smem + S::kOffOut + ew * S::kOutBufs * S::kStageOutBytes;
    const uint32_t stage_u32 = smem_u32(stage_out);
    uint64_t* rbar = res_full + ew * 2;
    const int cout0 = cb * 128 + quad * 32;  // this warp's 32 couts
    const float sc = __ldg(scale + cout0 + lane), sh = __ldg(shift + cout0 + lane);
    uint32_t it = 0, chunk = 0;
    // RES: request the residual sub-tile of (pixel tile pt, chunk j) into staging buffer b (rows beyond M: clipped)
    auto res_request = [&](int pt, int j, uint32_t b) {
      mbar_arrive_expect_tx(&rbar[b], S::kStageOutBytes);
      tma_tensor_2d_g2s(stage_out + b * S::kStageOutBytes, &tmap_r, cout0, pt * 256 + hsel * 128 + j * 16, &rbar[b]);
    };
    if constexpr (RES) {
      pdl_wait();  // the residual may come from the previous kernel in the stream
      if (lane == 0 && first_pt < n_pt) res_request(first_pt, 0, 0);
    }
    for (int pt = first_pt; pt < n_pt; pt += pt_stride, ++it) {
      const uint32_t buf = it & 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256 + hsel * 128;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        float v[16];
        tmem_ld_x16(taddr + j * 16, v);
        const uint32_t sbuf = (chunk % S::kOutBufs) * S::kStageOutBytes;
        if constexpr (RES) {
          // the OTHER buffer's store (previous chunk) must have read it before the next residual lands there
          if (lane == 0) {
            tma_store_wait_read<0>();
            const int nj = j + 1 < 8 ? j + 1 : 0, npt = j + 1 < 8 ? pt : pt + pt_stride;
            if (npt < n_pt) res_request(npt, nj, (chunk + 1) & 1);
          }
          __syncwarp();
          mbar_wait(&rbar[chunk & 1], (chunk >> 1) & 1);
        } else {
          if (lane == 0) tma_store_wait_read<S::kOutBufs - 1>();  // the staging buffer about to be written has been read
          __syncwarp();
        }
        tmem_ld_wait();
        const uint32_t dst = stage_u32 + sbuf + lane * 4;
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          float o = fmaf(sc, v[x], sh);
          if (relu) o = fmaxf(o, 0.f);
          if constexpr (RES) {
            o += ld_shared_f32(dst + x * 128);  // this thread's own column of the residual sub-tile
            if (relu_after) o = fmaxf(o, 0.f);
          }
          st_shared_f32(dst + x * 128, o);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_tensor_2d_s2g(&tmap_y, stage_out + sbuf, cout0, pt * 256 + hsel * 128 + j * 16);
          tma_store_commit();
        }
        ++chunk;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && crank != 0) mbar_arrive_remote_plain(&acc_empty[buf], 0);  // the leader issues the MMAs of both CTAs
        else mbar_arrive(&acc_empty[buf]);
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();  // the peer's shared memory and barriers stay alive
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc_cg2<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

// ---- Padded-frame output (chain mode, WG_OUT_PADDED; the reference's 14x14 maps): the 1x1 layer in front of a 3x3 layer
// writes the zero-bordered [N][16][16][Cout] frame that layer reads (Kernel128_winograd.cu:163,243 layout). Same
// transposed GEMM; a work item is 16 IMAGE ROWS (N = 224 pixels = 16 x 14, any alignment to image boundaries), so that
// the epilogue's unit -- 14 accumulator columns of one image row -- becomes ONE 16-pixel frame row (x = 0 / 15 zeroed) and
// goes out as one TMA tensor store, like the 3x3 direct kernel's frame output; the rows y = 0 / 15 of an image are stored
// from an all-zero tile by the warp that handles its first / last image row (not with WG_OUT_INTERIOR_ONLY). Weights are
// streamed with the activations (any Cin), CTA pairs when the couts come in groups of 256.
template <bool PAIR>
struct OneTFSmem {
  static constexpr int kSX = PAIR ? 6 : 4;
  static constexpr int kRows = PAIR ? 112 : 224;          // pixels of the item this CTA loads
  static constexpr uint32_t kXBytes = kRows * 128;
  static constexpr uint32_t kWBytes = 128 * 128;
  static constexpr uint32_t kStageOutBytes = 16 * 128;
  static constexpr uint32_t kOffX = 0;
  static constexpr uint32_t kOffW = kOffX + kSX * kXBytes;
  static constexpr uint32_t kOffOut = kOffW + kSX * kWBytes;  // [8 warps][2 buffers]
  static constexpr uint32_t kOffZero = kOffOut + 8 * 2 * kStageOutBytes;
  static constexpr uint32_t kOffBar = kOffZero + kStageOutBytes;
  static constexpr uint32_t kNumBars = 2 * kSX + 4;
  static constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr uint32_t kTotal = kOffTmemPtr + 16 + 1024;
  static_assert(kOffW % 1024 == 0 && kXBytes % 1024 == 0, "swizzled buffers must be 1024-byte aligned");
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

template <bool PAIR>
__global__ void __launch_bounds__(kOneTThreads, 1)
conv1x1_tf_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
                  const float* __restrict__ w_img, const float* __restrict__ scale, const float* __restrict__ shift,
                  int n_img, int Cin, int Cout, int relu, int bn_packed, int interior_only) {
  using S = OneTFSmem<PAIR>;
  constexpr uint16_t kPairMask = 0x3;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* x_full = bars;
  uint64_t* x_empty = x_full + S::kSX;
  uint64_t* acc_full = x_empty + S::kSX;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::kOffTmemPtr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_y);
    const uint32_t n_full = (PAIR && crank == 0) ? 2 : 1;
    for (int i = 0; i < S::kSX; ++i) mbar_init(&x_full[i], n_full), mbar_init(&x_empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], PAIR ? 16 : 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_cg2<512>(tmem_ptr);
    else tmem_alloc<512>(tmem_ptr);
  }
  if (warp >= 2) {  // the all-zero frame row
    for (uint32_t i = threadIdx.x - 64; i < S::kStageOutBytes / 16; i += kOneTThreads - 64)
      reinterpret_cast<uint4*>(smem + S::kOffZero)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_kb = Cin / 32;
  const int n_rows = n_img * 14;               // image rows of the batch
  const int n_pt = (n_rows + 15) / 16;         // items: 16 image rows = 224 pixels
  constexpr int kCL = PAIR ? 2 : 1;
  const int n_cg = Cout / (128 * kCL);
  const int cl_id = (int)blockIdx.x / kCL, n_cl = (int)gridDim.x / kCL;
  const int cb = (cl_id % n_cg) * kCL + (int)crank;
  const int first_pt = cl_id / n_cg, pt_stride = n_cl / n_cg;

  if (warp == 0) {
    if (elect_one()) {
      const int col0 = cb * 128;
      const uint8_t* w_src = reinterpret_cast<const uint8_t*>(w_img) +
                             ((size_t)(col0 / bn_packed) * n_kb * bn_packed + col0 % bn_packed) * 128;
      pdl_wait();
      uint32_t sx = 0, px = 0;
      for (int pt = first_pt; pt < n_pt; pt += pt_stride)
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&x_empty[sx], px ^ 1);
          mbar_arrive_expect_tx(&x_full[sx], S::kXBytes + S::kWBytes);
          tma_tensor_2d_g2s(smem + S::kOffX + sx * S::kXBytes, &tmap_x, kb * 32, pt * 224 + (PAIR ? (int)crank * 112 : 0),
                            &x_full[sx]);
          tma_bulk_g2s(smem + S::kOffW + sx * S::kWBytes, w_src + (size_t)kb * bn_packed * 128, S::kWBytes, &x_full[sx]);
          if (++sx == S::kSX) { sx = 0; px ^= 1; }
        }
    }
  } else if (warp == 1 && PAIR && crank != 0) {
    if (elect_one()) {
      uint32_t sx = 0, px = 0;
      for (int pt = first_pt; pt < n_pt; pt += pt_stride)
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&x_full[sx], px);
          mbar_arrive_remote_plain(&x_full[sx], 0);
          if (++sx == S::kSX) { sx = 0; px ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(kFmtTF32, PAIR ? 256 : 128, 224);
      const uint32_t x_base = smem_u32(smem + S::kOffX);
      const uint32_t w_base = smem_u32(smem + S::kOffW);
      uint32_t sx = 0, px = 0, it = 0;
      for (int pt = first_pt; pt < n_pt; pt += pt_stride, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&x_full[sx], px);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a_desc = make_smem_desc(w_base + sx * S::kWBytes + k * 32, 0, 1024, kLayoutSW128);
            const uint64_t b_desc = make_smem_desc(x_base + sx * S::kXBytes + k * 32, 0, 1024, kLayoutSW128);
            if constexpr (PAIR) umma_tf32_ss_cg2(tmem_base + buf * 256, a_desc, b_desc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else umma_tf32_ss(tmem_base + buf * 256, a_desc, b_desc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if constexpr (PAIR) umma_commit_mcast_cg2(&x_empty[sx], kPairMask);
          else umma_commit(&x_empty[sx]);
          if (++sx == S::kSX) { sx = 0; px ^= 1; }
        }
        if constexpr (PAIR) umma_commit_mcast_cg2(&acc_full[buf], kPairMask);
        else umma_commit(&acc_full[buf]);
      }
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3, hsel = ew >> 2;
    uint8_t* stage_out = smem + S::kOffOut + ew * 2 * S::kStageOutBytes;
    const uint32_t stage_u32 = smem_u32(stage_out);
    const int cout0 = cb * 128 + quad * 32;
    const float sc = __ldg(scale + cout0 + lane), sh = __ldg(shift + cout0 + lane);
    uint32_t it = 0, chunk = 0;
    for (int pt = first_pt; pt < n_pt; pt += pt_stride, ++it) {
      const uint32_t buf = it & 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        const int jr = hsel * 8 + j;      // image row inside the item
        const int r = pt * 16 + jr;       // image row of the batch
        if (r >= n_rows) break;           // warp-uniform
        const int img = r / 14, yy = r - img * 14;
        float v[16];
        tmem_ld_x16(taddr + jr * 14, v);  // 14 pixels of this image row (+ two of the next, unused)
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        tmem_ld_wait();
        const uint32_t sbuf = (chunk & 1) * S::kStageOutBytes;
        const uint32_t dst = stage_u32 + sbuf + lane * 4;
        st_shared_f32(dst, 0.f);             // frame column 0
        st_shared_f32(dst + 15 * 128, 0.f);  // frame column 15
#pragma unroll
        for (int x = 0; x < 14; ++x) {
          float o = fmaf(sc, v[x], sh);
          if (relu) o = fmaxf(o, 0.f);
          st_shared_f32(dst + (x + 1) * 128, o);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_tensor_2d_s2g(&tmap_y, stage_out + sbuf, cout0, img * 256 + (yy + 1) * 16);
          tma_store_commit();
          if (!interior_only && yy == 0) {
            tma_tensor_2d_s2g(&tmap_y, smem + S::kOffZero, cout0, img * 256);
            tma_store_commit();
          }
          if (!interior_only && yy == 13) {
            tma_tensor_2d_s2g(&tmap_y, smem + S::kOffZero, cout0, img * 256 + 15 * 16);
            tma_store_commit();
          }
        }
        ++chunk;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && crank != 0) mbar_arrive_remote_plain(&acc_empty[buf], 0);
        else mbar_arrive(&acc_empty[buf]);
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc_cg2<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

static int onet_encode(CUtensorMap* tmap, const float* base, int inner, long long rows, int box_rows,
                       CUtensorMapSwizzle swz) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return WG_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)inner * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, wg::l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WG_OK : WG_ERR_TMAP;
}
// activations: 256 (CTA pairs: 128) pixels x 32 channels per box, 128-byte swizzle
int onet_make_tmap_in(CUtensorMap* tmap, const float* x, long long m_rows, int Cin, int Cout) {
  return onet_encode(tmap, x, Cin, m_rows, onet_pair(Cout) ? 128 : 256, CU_TENSOR_MAP_SWIZZLE_128B);
}
// output / residual: 16 pixels x 32 couts per box, plain layout
int onet_make_tmap_out(CUtensorMap* tmap, const float* y, long long m_rows, int Cout) {
  return onet_encode(tmap, y, Cout, m_rows, 16, CU_TENSOR_MAP_SWIZZLE_NONE);
}

// CTA pairs whenever the couts come in groups of 256 (developer build: WG_ONE_T_PAIR=0 switches them off)
bool onet_pair(int Cout) {
  static int v = -1;
  if (v < 0) {
    const char* e = dev_env("WG_ONE_T_PAIR");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v && Cout % 256 == 0;
}

bool onet_eligible(long long m_rows, int Cin, int Cout, int max_ctas) {
  if (Cin % 32 != 0 || Cin > 256 || Cout % 128 != 0 || Cout < 2 * Cin) return false;
  const long long n_cb = Cout / 128, n_pt = (m_rows + 255) / 256;
  // from one pixel tile per CTA on (measured against conv1x1_bn_act_kernel, profiles/one_ablation_r02.md: ahead from
  // there, level below). Developer build: WG_ONE_T_MIN = tiles per CTA from which the kernel is used, in quarters.
  static int q = -1;
  if (q < 0) {
    const char* e = dev_env("WG_ONE_T_MIN");
    q = e ? atoi(e) : 4;
  }
  return max_ctas >= n_cb && 4 * n_pt >= q * (max_ctas / n_cb);
}

template <int CIN_MAX, bool RES, bool PAIR>
static int launch_onet(const CUtensorMap& tmap_x, const CUtensorMap& tmap_y, const CUtensorMap& tmap_r,
                       const float* w_img, const float* scale, const float* shift, long long m_rows, int Cin, int Cout,
                       int relu, int bn_packed, int relu_after, int max_ctas, cudaStream_t stream) {
  using S = OneTSmem<CIN_MAX, PAIR, RES>;
  static unsigned long long configured = 0;
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long bit = 1ull << (dev_ & 63);
  if (!(configured & bit)) {
    if (cudaFuncSetAttribute(conv1x1_t_kernel<CIN_MAX, RES, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)S::kTotal) != cudaSuccess)
      return WG_ERR_CUDA;
    configured |= bit;
  }
  constexpr int CL = PAIR ? 2 : 1;
  const long long n_cg = Cout / (128 * CL), n_pt = (m_rows + 255) / 256;
  long long n_cl = (max_ctas / CL / n_cg) * n_cg;  // clusters: a multiple of the cout groups
  if (n_cl > n_pt * n_cg) n_cl = n_pt * n_cg;
  if (n_cl < n_cg) n_cl = n_cg;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_cl * CL));
  cfg.blockDim = dim3(kOneTThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv1x1_t_kernel<CIN_MAX, RES, PAIR>, tmap_x, tmap_y, tmap_r, w_img, scale, shift,
                                     m_rows, Cin, Cout, relu, bn_packed, relu_after);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

// bn_packed: row count of a packed weight tile of the layer's image (128 or 256)
int onet_launch(const CUtensorMap& tmap_x, const CUtensorMap& tmap_y, const CUtensorMap* tmap_r, const float* w_img,
                const float* scale, const float* shift, long long m_rows, int Cin, int Cout, int relu, int bn_packed,
                int relu_after, int max_ctas, cudaStream_t stream) {
#define WG_ONET(CM_, R_, P_)                                                                                  \
  return launch_onet<CM_, R_, P_>(tmap_x, tmap_y, tmap_r ? *tmap_r : tmap_y, w_img, scale, shift, m_rows, Cin, Cout, \
                                  relu, bn_packed, relu_after, max_ctas, stream)
  const bool pair = onet_pair(Cout);
  if (Cin <= 128) {
    if (tmap_r && pair) WG_ONET(128, true, true);
    if (tmap_r) WG_ONET(128, true, false);
    if (pair) WG_ONET(128, false, true);
    WG_ONET(128, false, false);
  }
  if (tmap_r && pair) WG_ONET(256, true, true);
  if (tmap_r) WG_ONET(256, true, false);
  if (pair) WG_ONET(256, false, true);
  WG_ONET(256, false, false);
#undef WG_ONET
}

// ---- frame-output variant
bool onetf_eligible(int n_img, int Cin, int Cout, int max_ctas) {
  static int on = -1;  // developer build: WG_ONE_TF=0 keeps the frame output on conv1x1_bn_act_kernel
  if (on < 0) {
    const char* e = dev_env("WG_ONE_TF");
    on = e ? atoi(e) : 1;
  }
  if (!on || Cin % 32 != 0 || Cout % 128 != 0) return false;
  const long long n_cb = Cout / 128, n_pt = ((long long)n_img * 14 + 15) / 16;
  return max_ctas >= n_cb && n_pt >= max_ctas / n_cb;  // at least one item per CTA
}
int onetf_make_tmap_in(CUtensorMap* tmap, const float* x, long long m_rows, int Cin, int Cout) {
  return onet_encode(tmap, x, Cin, m_rows, onet_pair(Cout) ? 112 : 224, CU_TENSOR_MAP_SWIZZLE_128B);
}
int onetf_make_tmap_out(CUtensorMap* tmap, const float* y_frame, int n_img, int Cout) {
  return onet_encode(tmap, y_frame, Cout, (long long)n_img * 256, 16, CU_TENSOR_MAP_SWIZZLE_NONE);
}
template <bool PAIR>
static int launch_onetf(const CUtensorMap& tmap_x, const CUtensorMap& tmap_y, const float* w_img, const float* scale,
                        const float* shift, int n_img, int Cin, int Cout, int relu, int bn_packed, int interior_only,
                        int max_ctas, cudaStream_t stream) {
  using S = OneTFSmem<PAIR>;
  static unsigned long long configured = 0;
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long bit = 1ull << (dev_ & 63);
  if (!(configured & bit)) {
    if (cudaFuncSetAttribute(conv1x1_tf_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kTotal) !=
        cudaSuccess)
      return WG_ERR_CUDA;
    configured |= bit;
  }
  constexpr int CL = PAIR ? 2 : 1;
  const long long n_cg = Cout / (128 * CL), n_pt = ((long long)n_img * 14 + 15) / 16;
  long long n_cl = (max_ctas / CL / n_cg) * n_cg;
  if (n_cl > n_pt * n_cg) n_cl = n_pt * n_cg;
  if (n_cl < n_cg) n_cl = n_cg;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_cl * CL));
  cfg.blockDim = dim3(kOneTThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv1x1_tf_kernel<PAIR>, tmap_x, tmap_y, w_img, scale, shift, n_img, Cin, Cout,
                                     relu, bn_packed, interior_only);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}
int onetf_launch(const CUtensorMap& tmap_x, const CUtensorMap& tmap_y, const float* w_img, const float* scale,
                 const float* shift, int n_img, int Cin, int Cout, int relu, int bn_packed, int interior_only,
                 int max_ctas, cudaStream_t stream) {
  if (onet_pair(Cout))
    return launch_onetf<true>(tmap_x, tmap_y, w_img, scale, shift, n_img, Cin, Cout, relu, bn_packed, interior_only,
                              max_ctas, stream);
  return launch_onetf<false>(tmap_x, tmap_y, w_img, scale, shift, n_img, Cin, Cout, relu, bn_packed, interior_only,
                             max_ctas, stream);
}

}  // namespace wg
