// 3x3 conv + folded BatchNorm (+ ReLU) as a DIRECT convolution on the tensor core (implicit GEMM without an im2col
// buffer and without any CUDA-core transform), sm_100a.
//
// Same operator as the fused Winograd kernels (/root/reference/Kernel128_winograd.cu:153-262, Kernel256_winograd.cu:
// 176-296: input = zero-bordered [N][16][16][C] frame, weights [K][C][3][3], output [N][14][14][K] or the next layer's
// frame), different decomposition. With the frame flattened to rows p = (n*16 + y)*16 + x of C channels, the output at
// row p is  sum over the 9 taps (dy, dx) of  W[dy][dx] . X[p + (dy-1)*16 + (dx-1)]  -- nine GEMMs whose activation
// operands are the SAME rows moved by a constant. The GEMM is laid out TRANSPOSED: M = output channels (128 per CTA),
// N = pixels, K = input channels; N = 224 is exactly the 14 valid rows (y = 1..14, all 16 x) of one image, so only the
// two border columns are computed for nothing (196 of 224 outputs useful; with pixels on M, tiles of 128 rows would
// waste the border rows too: 196 of 256).
//   activations (B operand): one TMA box per 32-channel chunk brings the image's rows p0-24 .. into shared memory ONCE;
//     each tap's tcgen05.mma reads them through a descriptor whose start address is moved by (dy-1)*16 + (dx-1) rows;
//   weights (A operand): pre-swizzled [128 couts][32 channels] block per (chunk, tap), one bulk copy each;
//   accumulator: TMEM lanes = couts, columns = pixels; the epilogue warps scale / shift / ReLU with their lane's own
//     folded-BN pair, transpose 16 pixels x 32 couts (= one frame row) through shared memory and write it with one TMA
//     tensor store -- to the dense [N][14][14][K] map (14 pixels) or to the next layer's frame (16 pixels, x = 0 / 15
//     zeroed; the rows y = 0 / 15 are written as zeros too).
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = one thread issuing 4 MMAs (M = 128, N = 224, K = 8) per
// (chunk, tap) into one of two TMEM accumulator buffers, warps 2..9 = epilogue of the previous item (two warps per TMEM
// lane quadrant).
// CL = thread-block cluster size: the CL CTAs of a cluster work on CL consecutive IMAGES and the same 128 couts, in step;
// each loads 1/CL of every weight block and multicasts it to all (the weight stream out of L2 -- 16 KB per (chunk, tap)
// and CTA, 34 B/clk and SM at full MMA rate, 68 B/clk for the half items -- is what this kernel is bound by when every
// CTA fetches its own copy; the L2 delivers ~12 TB/s = 43 B/clk per SM to 148 SMs).
// (A cta_group::2 variant -- M = 256 couts over a CTA pair, each CTA holding half of the image's pixels -- was measured
// too: same time per item as CL = 1, it shares the activations, which are the small stream, not the weights.)
#include <stdlib.h>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"
#include "wg_internal.h"

namespace wg {

constexpr int kDirThreads = 32 * 10;  // producer, MMA, 8 epilogue warps
constexpr int kDirHalo = 24;          // rows in front of the first pixel: >= 17 (one frame row + 1), multiple of 8
constexpr int kDirN = 224;            // pixels of a full item: frame rows 1..14
constexpr int kDirAccCols = 256;      // TMEM columns per accumulator buffer

struct DirSmem {
  static constexpr int kSX = 2;  // activation chunks in flight
  static constexpr int kSW = 7;  // weight blocks in flight
  // activation rows per chunk: the image's 224 pixels + halo on both sides, as two TMA boxes (<= 256 rows each)
  static constexpr int kXRows = kDirN + 2 * kDirHalo;  // 272
  static constexpr int kXBoxRows = 136;
  static constexpr int kXBoxes = kXRows / kXBoxRows;
  static constexpr uint32_t kXBytes = kXRows * 128;
  static constexpr uint32_t kWBytes = 128 * 128;        // [128 couts][32 channels] fp32
  static constexpr uint32_t kStageOutBytes = 16 * 128;  // one frame row: [16 px][32 couts]
  static constexpr uint32_t kOffX = 0;
  static constexpr uint32_t kOffW = kOffX + kSX * kXBytes;
  static constexpr uint32_t kOffOut = kOffW + kSW * kWBytes;              // [8 warps][2 buffers]
  static constexpr uint32_t kOffZero = kOffOut + 8 * 2 * kStageOutBytes;  // one all-zero frame row (padded output)
  static constexpr uint32_t kOffBar = kOffZero + kStageOutBytes;
  static constexpr uint32_t kNumBars = 2 * kSX + 2 * kSW + 4;
  static constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr uint32_t kTotal = kOffTmemPtr + 16 + 1024;
  static_assert(kXRows % kXBoxRows == 0 && (kXBoxRows * 128) % 1024 == 0, "activation boxes");
  static_assert(kOffW % 1024 == 0 && kOffOut % 1024 == 0, "swizzled buffers must be 1024-byte aligned");
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

// Work items of a cluster. The first n_big items are (group of CL images, 128-cout block), cout block fastest, whole
// images (N = 224 pixels); the rest are half images (N = 112: frame rows 1..7 or 8..14) -- the host sizes n_big so that
// the whole-image items fill whole rounds of the grid and what is left of the batch is spread over all clusters as half
// items (a partial last round then costs half an item).
struct DirItem {
  int grp, cb, half;  // half: -1 = whole image, 0 / 1 = rows 1..7 / 8..14
};
__device__ __forceinline__ DirItem dir_item(int i, int n_big, int n_cb) {
  if (i < n_big) return DirItem{i / n_cb, i % n_cb, -1};
  const int j = i - n_big;
  return DirItem{n_big / n_cb + j / (2 * n_cb), (j % (2 * n_cb)) >> 1, j & 1};
}

template <int CL>
__global__ void __launch_bounds__(kDirThreads, 1)
conv3x3_direct_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
                      const float* __restrict__ w_img, const float* __restrict__ scale, const float* __restrict__ shift,
                      int n_img, int Cin, int Cout, int relu, int out_padded, int n_big, int n_items) {
  using S = DirSmem;
  constexpr uint16_t kClusterMask = (uint16_t)((1u << CL) - 1u);
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* x_full = bars;
  uint64_t* x_empty = x_full + S::kSX;
  uint64_t* w_full = x_empty + S::kSX;
  uint64_t* w_empty = w_full + S::kSW;
  uint64_t* acc_full = w_empty + S::kSW;  // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::kOffTmemPtr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_y);
    for (int i = 0; i < S::kSX; ++i) mbar_init(&x_full[i], 1), mbar_init(&x_empty[i], 1);
    // a weight stage is refilled (by all CTAs of the cluster) once every CTA's MMAs have read it
    for (int i = 0; i < S::kSW; ++i) mbar_init(&w_full[i], 1), mbar_init(&w_empty[i], CL);
    for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  if (warp >= 2) {  // the all-zero frame row
    for (uint32_t i = threadIdx.x - 64; i < S::kStageOutBytes / 16; i += kDirThreads - 64)
      reinterpret_cast<uint4*>(smem + S::kOffZero)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_c = Cin / 32;  // 32-channel chunks
  // 128-cout blocks; a layer with Cout % 128 == 64 has a last block whose upper 64 weight rows are zeros in the packed
  // image (half of those MMAs' rows are idle) and whose upper two epilogue quadrants store nothing
  const int n_cb = (Cout + 127) / 128;
  const int first_item = (int)blockIdx.x / CL;
  const int item_stride = (int)gridDim.x / CL;
  // this CTA's image of an item; the last group of an odd batch repeats the last image (same values stored twice)
#define WG_DIR_IMG(w) min((w).grp * CL + (int)crank, n_img - 1)

  if (warp == 0) {
    if (elect_one()) {
      uint32_t sx = 0, px = 0, sw = 0, pw = 0;
      pdl_wait();  // the frame comes from the previous kernel in the stream
      for (int item = first_item; item < n_items; item += item_stride) {
        const DirItem w = dir_item(item, n_big, n_cb);
        const int p0 = WG_DIR_IMG(w) * 256 + 16 + (w.half > 0 ? kDirN / 2 : 0);  // first pixel (frame row index)
        const uint8_t* w_src = reinterpret_cast<const uint8_t*>(w_img) + (size_t)w.cb * n_c * 9 * S::kWBytes;
        for (int c = 0; c < n_c; ++c) {
          mbar_wait(&x_empty[sx], px ^ 1);
          mbar_arrive_expect_tx(&x_full[sx], S::kXBytes);
#pragma unroll
          for (int b = 0; b < S::kXBoxes; ++b)
            tma_tensor_2d_g2s(smem + S::kOffX + sx * S::kXBytes + b * (S::kXBoxRows * 128), &tmap_x, c * 32,
                              p0 - kDirHalo + b * S::kXBoxRows, &x_full[sx]);
          if (++sx == S::kSX) { sx = 0; px ^= 1; }
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&w_empty[sw], pw ^ 1);
            mbar_arrive_expect_tx(&w_full[sw], S::kWBytes);
            const uint8_t* blk = w_src + (size_t)(c * 9 + t) * S::kWBytes;
            if constexpr (CL == 1) {
              tma_bulk_g2s(smem + S::kOffW + sw * S::kWBytes, blk, S::kWBytes, &w_full[sw]);
            } else {
              constexpr uint32_t part = S::kWBytes / CL;  // couts [crank*128/CL, +128/CL) of the swizzled block
              tma_bulk_g2s_mcast(smem + S::kOffW + sw * S::kWBytes + crank * part, blk + crank * part, part,
                                 &w_full[sw], kClusterMask);
            }
            if (++sw == S::kSW) { sw = 0; pw ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t x_base = smem_u32(smem + S::kOffX);
      const uint32_t w_base = smem_u32(smem + S::kOffW);
      uint32_t sx = 0, px = 0, sw = 0, pw = 0, it = 0;
      for (int item = first_item; item < n_items; item += item_stride, ++it) {
        const uint32_t idesc = make_idesc(kFmtTF32, 128, item < n_big ? kDirN : kDirN / 2);
        const uint32_t buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kDirAccCols;
        for (int c = 0; c < n_c; ++c) {
          mbar_wait(&x_full[sx], px);
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&w_full[sw], pw);
            tc_fence_after();
            // pixels of tap (dy, dx): the item's rows moved by (dy-1)*16 + (dx-1) frame pixels. The start address is then
            // 128-byte but not 1024-byte aligned; the 128-byte swizzle is a function of the ADDRESS bits (7..9 into
            // 4..6) for the TMA write and the MMA read alike, so the shifted descriptor reads the rows as written
            // (measured: same results as an aligned copy; the descriptor's base-offset field stays 0).
            const int rshift = (t / 3 - 1) * 16 + (t % 3 - 1);
            const uint32_t x_tap = x_base + sx * S::kXBytes + (uint32_t)(kDirHalo + rshift) * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a_desc = make_smem_desc(w_base + sw * S::kWBytes + k * 32, 0, 1024, kLayoutSW128);
              const uint64_t b_desc = make_smem_desc(x_tap + k * 32, 0, 1024, kLayoutSW128);
              umma_tf32_ss(d_tmem, a_desc, b_desc, idesc, (c > 0 || t > 0 || k > 0) ? 1u : 0u);
            }
            if constexpr (CL == 1) umma_commit(&w_empty[sw]);
            else umma_commit_mcast(&w_empty[sw], kClusterMask);
            if (++sw == S::kSW) { sw = 0; pw ^= 1; }
          }
          umma_commit(&x_empty[sx]);
          if (++sx == S::kSX) { sx = 0; px ^= 1; }
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // epilogue: warps 2..9; warp & 3 = TMEM lane quadrant (32 of the item's couts), (warp - 2) / 4 = which half of the
    // item's frame rows. One chunk = 16 accumulator columns = one frame row of 16 pixels.
    const int ew = warp - 2;
    const int quad = warp & 3, hsel = ew >> 2;
    uint8_t* stage_out = smem + S::kOffOut + ew * 2 * S::kStageOutBytes;
    const uint32_t stage_u32 = smem_u32(stage_out);
    uint32_t it = 0, chunk = 0;
    for (int item = first_item; item < n_items; item += item_stride, ++it) {
      const DirItem w = dir_item(item, n_big, n_cb);
      const int img = WG_DIR_IMG(w);
      const uint32_t buf = it & 1;
      const int cout0 = w.cb * 128 + quad * 32;  // this warp's 32 couts
      const bool couts_here = cout0 < Cout;      // false: padding rows of the last block (warp-uniform)
      const float sc = couts_here ? __ldg(scale + cout0 + lane) : 0.f, sh = couts_here ? __ldg(shift + cout0 + lane) : 0.f;
      const int rows = w.half < 0 ? 14 : 7;          // frame rows of the item
      const int y_first = 1 + (w.half > 0 ? 7 : 0);  // frame row of accumulator columns 0..15
      const int j0 = hsel ? (rows + 1) / 2 : 0, j1 = (!couts_here) ? j0 : (hsel ? rows : (rows + 1) / 2);
      if (out_padded == 1 && lane == 0 && couts_here) {
        // the frame's zero rows y = 0 / y = 15 (not with WG_OUT_INTERIOR_ONLY, out_padded == 3)
        if (hsel == 0 && w.half <= 0) {
          tma_tensor_2d_s2g(&tmap_y, smem + S::kOffZero, cout0, img * 256);
          tma_store_commit();
        }
        if (hsel == 1 && w.half != 0) {
          tma_tensor_2d_s2g(&tmap_y, smem + S::kOffZero, cout0, img * 256 + 15 * 16);
          tma_store_commit();
        }
      }
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * kDirAccCols;
#pragma unroll 1
      for (int j = j0; j < j1; ++j) {
        float v[16];
        tmem_ld_x16(taddr + j * 16, v);
        if (lane == 0) tma_store_wait_read<1>();  // the staging buffer written two chunks ago has been read
        __syncwarp();
        tmem_ld_wait();
        const uint32_t dst = stage_u32 + (chunk & 1) * S::kStageOutBytes + lane * 4;
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          float o = fmaf(sc, v[x], sh);
          if (relu) o = fmaxf(o, 0.f);
          if (x == 0 || x == 15) o = 0.f;  // the frame's border columns (never stored to the dense map)
          st_shared_f32(dst + x * 128, o);  // [px][cout]: a warp writes one 128-byte row per instruction
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int yy = y_first + j;
          if (out_padded)  // next layer's frame: 16 pixels of frame row yy
            tma_tensor_2d_s2g(&tmap_y, stage_out + (chunk & 1) * S::kStageOutBytes, cout0, img * 256 + yy * 16);
          else  // dense map: pixels x = 1..14 of the row
            tma_tensor_2d_s2g(&tmap_y, stage_out + (chunk & 1) * S::kStageOutBytes + 128, cout0,
                              img * 196 + (yy - 1) * 14);
          tma_store_commit();
        }
        ++chunk;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if (lane == 0) tma_store_wait_read<0>();  // the staging buffers have been read; the writes complete with the grid
  }
#undef WG_DIR_IMG

  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all(); else __syncthreads();  // no CTA may leave while peers still multicast to it
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---- Other map sizes (wg_conv3x3_create_hw: 56x56, 28x28, 7x7 ...; TF32): the same direct convolution with the
// geometry as launch parameters. The frame [N][Hf][Wf][C] is still one flat sequence of pixel rows, tap (dy, dx) is a
// shift by (dy-1)*Wf + (dx-1), and a work item is a contiguous range of that sequence: R frame rows of one image
// (R*Wf <= 256: 28x28 -> 8 rows, 56x56 -> 4) or, for small maps, G whole images (7x7: two 10x10 frames per item). Pixels of
// the range that are border / padding / beyond the batch are computed and not stored (dense output) or stored as the
// zeros the next layer's frame needs. Output pixels are not affine in the range index here, so the epilogue writes with
// plain 128-byte row stores (8 lanes x 16 B per pixel) through a per-warp table of destinations.
// DirGeo (wg_internal.h): R, bands, G: item = R frame rows of one image (bands per image), or G images (bands == 1);
// n_pad = MMA N (pixels of an item rounded up to 16); halo = rows loaded in front of / behind the range (>= Wf + 1,
// multiple of 8); the activation stage is n_boxes TMA boxes of box_rows rows.
struct DirGenSmem {
  static constexpr int kSX = 2, kSW = 5;
  static constexpr int kXRowsMax = 384;
  static constexpr uint32_t kXBytes = kXRowsMax * 128;
  static constexpr uint32_t kWBytes = 128 * 128;
  static constexpr uint32_t kStageOutBytes = 16 * 128;
  static constexpr uint32_t kOffX = 0;
  static constexpr uint32_t kOffW = kOffX + kSX * kXBytes;
  static constexpr uint32_t kOffOut = kOffW + kSW * kWBytes;   // [8 warps] one staging tile each
  static constexpr uint32_t kOffTab = kOffOut + 8 * kStageOutBytes;  // [8 warps][128] destination per column
  static constexpr uint32_t kOffBar = kOffTab + 8 * 128 * 4;
  static constexpr uint32_t kNumBars = 2 * kSX + 2 * kSW + 4;
  static constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr uint32_t kTotal = kOffTmemPtr + 16 + 1024;
  static_assert(kOffW % 1024 == 0 && kOffOut % 1024 == 0, "swizzled buffers must be 1024-byte aligned");
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

__global__ void __launch_bounds__(kDirThreads, 1)
conv3x3_direct_gen_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ w_img,
                          const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ y,
                          int n_img, int Cin, int Cout, int relu, int out_padded, int n_items, const DirGeo g) {
  using S = DirGenSmem;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* x_full = bars;
  uint64_t* x_empty = x_full + S::kSX;
  uint64_t* w_full = x_empty + S::kSX;
  uint64_t* w_empty = w_full + S::kSW;
  uint64_t* acc_full = w_empty + S::kSW;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::kOffTmemPtr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < S::kSX; ++i) mbar_init(&x_full[i], 1), mbar_init(&x_empty[i], 1);
    for (int i = 0; i < S::kSW; ++i) mbar_init(&w_full[i], 1), mbar_init(&w_empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_c = Cin / 32;
  const int n_cb = (Cout + 127) / 128;  // the last block may be half padding (Cout % 128 == 64), see the 14x14 kernel
  const int frame = g.Hf * g.Wf;
  const int first_item = (int)blockIdx.x, item_stride = (int)gridDim.x;
  // item -> (cout block, first image, first frame row of the range)
#define WG_GEN_ITEM(item, cb_, img0_, y0_)            \
  const int cb_ = (item) % n_cb;                      \
  const int u_##cb_ = (item) / n_cb;                  \
  const int img0_ = (u_##cb_ / g.bands) * g.G;        \
  const int y0_ = 1 + (u_##cb_ % g.bands) * g.R

  if (warp == 0) {
    if (elect_one()) {
      uint32_t sx = 0, px = 0, sw = 0, pw = 0;
      const uint32_t x_bytes = (uint32_t)(g.n_boxes * g.box_rows) * 128;
      pdl_wait();
      for (int item = first_item; item < n_items; item += item_stride) {
        WG_GEN_ITEM(item, cb, img0, y0);
        const int p0 = img0 * frame + y0 * g.Wf;
        const uint8_t* w_src = reinterpret_cast<const uint8_t*>(w_img) + (size_t)cb * n_c * 9 * S::kWBytes;
        for (int c = 0; c < n_c; ++c) {
          mbar_wait(&x_empty[sx], px ^ 1);
          mbar_arrive_expect_tx(&x_full[sx], x_bytes);
          for (int b = 0; b < g.n_boxes; ++b)
            tma_tensor_2d_g2s(smem + S::kOffX + sx * S::kXBytes + b * (g.box_rows * 128), &tmap_x, c * 32,
                              p0 - g.halo + b * g.box_rows, &x_full[sx]);
          if (++sx == S::kSX) { sx = 0; px ^= 1; }
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&w_empty[sw], pw ^ 1);
            mbar_arrive_expect_tx(&w_full[sw], S::kWBytes);
            tma_bulk_g2s(smem + S::kOffW + sw * S::kWBytes, w_src + (size_t)(c * 9 + t) * S::kWBytes, S::kWBytes,
                         &w_full[sw]);
            if (++sw == S::kSW) { sw = 0; pw ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t x_base = smem_u32(smem + S::kOffX);
      const uint32_t w_base = smem_u32(smem + S::kOffW);
      const uint32_t idesc = make_idesc(kFmtTF32, 128, (uint32_t)g.n_pad);
      uint32_t sx = 0, px = 0, sw = 0, pw = 0, it = 0;
      for (int item = first_item; item < n_items; item += item_stride, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kDirAccCols;
        for (int c = 0; c < n_c; ++c) {
          mbar_wait(&x_full[sx], px);
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&w_full[sw], pw);
            tc_fence_after();
            const int rshift = (t / 3 - 1) * g.Wf + (t % 3 - 1);
            const uint32_t x_tap = x_base + sx * S::kXBytes + (uint32_t)(g.halo + rshift) * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a_desc = make_smem_desc(w_base + sw * S::kWBytes + k * 32, 0, 1024, kLayoutSW128);
              const uint64_t b_desc = make_smem_desc(x_tap + k * 32, 0, 1024, kLayoutSW128);
              umma_tf32_ss(d_tmem, a_desc, b_desc, idesc, (c > 0 || t > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&w_empty[sw]);
            if (++sw == S::kSW) { sw = 0; pw ^= 1; }
          }
          umma_commit(&x_empty[sx]);
          if (++sx == S::kSX) { sx = 0; px ^= 1; }
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3, hsel = ew >> 2;
    const uint32_t stage_u32 = smem_u32(smem + S::kOffOut + ew * S::kStageOutBytes);
    int* tab = reinterpret_cast<int*>(smem + S::kOffTab) + ew * 128;
    const int n_chunks = g.n_pad / 16;
    const int j0 = hsel ? (n_chunks + 1) / 2 : 0, j1 = hsel ? n_chunks : (n_chunks + 1) / 2;
    const int sub = lane & 7, pxs = lane >> 3;  // store phase: 16-byte piece of a pixel's 128 bytes, pixel of a group of 4
    uint32_t it = 0;
    for (int item = first_item; item < n_items; item += item_stride, ++it) {
      WG_GEN_ITEM(item, cb, img0, y0);
      const uint32_t buf = it & 1;
      const int cout0 = cb * 128 + quad * 32;
      const bool couts_here = cout0 < Cout;  // false: padding rows of the last block (warp-uniform): nothing to store
      const float sc = couts_here ? __ldg(scale + cout0 + lane) : 0.f, sh = couts_here ? __ldg(shift + cout0 + lane) : 0.f;
      // destinations of this warp's columns: entry = (output pixel index << 1) | store-zero, or -1 = no store
      for (int j = j0 * 16 + lane; j < j1 * 16; j += 32) {
        const int q = y0 * g.Wf + j;
        const int img = img0 + q / frame, rem = q % frame, fy = rem / g.Wf, fx = rem % g.Wf;
        const bool in_batch = img < n_img;
        const bool interior = fy >= 1 && fy <= g.H && fx >= 1 && fx <= g.W;
        int e = -1;
        if (in_batch && out_padded) e = ((img * frame + rem) << 1) | (interior ? 0 : 1);
        else if (in_batch && interior) e = ((img * g.H + fy - 1) * g.W + fx - 1) << 1;
        tab[j - j0 * 16] = couts_here ? e : -1;
      }
      if (out_padded && couts_here) {
        // frame rows no item range covers: row 0 of the first image of the range (hsel 0) and what lies behind the last
        // covered row of its last image (hsel 1); zeros, this warp's 32 couts
        const int band = ((item / n_cb) % g.bands);
        const int rows_cov = (g.bands == 1) ? g.n_pad / g.Wf : g.R;  // whole frame rows of the range, from its first row
        int zimg = -1, zy0 = 0, zy1 = 0;
        if (hsel == 0 && band == 0) zimg = img0, zy0 = 0, zy1 = 1;
        if (hsel == 1 && band == g.bands - 1) {
          const int last = img0 + g.G - 1;
          const int covered_to = (g.bands == 1) ? (1 + rows_cov - (g.G - 1) * g.Hf) : (y0 + rows_cov);  // first row not covered
          zimg = last < n_img ? last : -1, zy0 = covered_to > g.H + 1 ? covered_to : g.H + 1, zy1 = g.Hf;
        }
        if (zimg >= 0 && zimg < n_img)
          for (int pz = zy0 * g.Wf + pxs; pz < zy1 * g.Wf; pz += 4)
            *reinterpret_cast<float4*>(y + ((size_t)zimg * frame + pz) * Cout + cout0 + sub * 4) =
                make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * kDirAccCols;
#pragma unroll 1
      for (int j = j0; j < j1; ++j) {
        float v[16];
        tmem_ld_x16(taddr + j * 16, v);
        tmem_ld_wait();
        const uint32_t dst = stage_u32 + lane * 4;
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          float o = fmaf(sc, v[x], sh);
          if (relu) o = fmaxf(o, 0.f);
          st_shared_f32(dst + x * 128, o);
        }
        __syncwarp();
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int pxl = i4 * 4 + pxs;
          const int e = tab[(j - j0) * 16 + pxl];
          if (e >= 0) {
            float4 val = ld_shared_v4(stage_u32 + pxl * 128 + sub * 16);
            if (e & 1) val = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(y + (size_t)(e >> 1) * Cout + cout0 + sub * 4) = val;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
#undef WG_GEN_ITEM

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// geometry of the direct kernel for an H x W map in its Hf x Wf frame; false = not supported (the Winograd kernel runs)
// (max_rows: rows an activation stage can hold -- 384 for the TF32 kernel, 304 for the 16-bit one)
bool direct_gen_geo(int H, int W, int Hf, int Wf, int max_rows, DirGeo* out) {
  DirGeo g{};
  g.H = H, g.W = W, g.Hf = Hf, g.Wf = Wf;
  g.halo = ((Wf + 1) + 7) / 8 * 8;
  const int frame = Hf * Wf;
  if (Wf > 128) return false;
  int n_px;
  if (frame + H * Wf <= 256) {  // at least two whole images per item
    g.bands = 1;
    g.G = (256 - H * Wf) / frame + 1;
    g.R = 0;
    n_px = (g.G - 1) * frame + H * Wf;
  } else {
    g.G = 1;
    g.R = 256 / Wf;
    if (g.R > H) g.R = H;
    while (g.R >= 1 && (g.R * Wf + 15) / 16 * 16 + 2 * g.halo > max_rows) --g.R;  // a shorter band for a smaller stage
    if (g.R < 1) return false;
    g.bands = (H + g.R - 1) / g.R;
    n_px = g.R * Wf;
  }
  g.n_pad = (n_px + 15) / 16 * 16;
  const int rows = g.n_pad + 2 * g.halo;
  if (g.n_pad > 256 || rows > max_rows) return false;
  g.n_boxes = (rows + 255) / 256;
  g.box_rows = ((rows + g.n_boxes - 1) / g.n_boxes + 7) / 8 * 8;  // multiple of 8 rows: every box 1024-byte aligned
  if (g.n_boxes * g.box_rows > max_rows) return false;
  *out = g;
  return true;
}

// ---- 16-bit operands (WG_BF16 / WG_FP16): the same direct convolution with kind::f16 MMAs (K = 16, twice the rate).
// The frame is still fp32 in HBM (the reference's format), so the operand tile is made on the SM: the TMA producer
// brings a chunk's 64 channels as two fp32 half-chunks (32 channels x 272 rows each) into a two-slot staging ring, four
// converter warps (10..13) round them to bf16 / fp16 (round to nearest) into a [272 rows][64 channels] 16-bit tile in the
// SAME K-major 128-byte-swizzled layout a TMA box load would produce (so that the shifted-descriptor trick applies
// unchanged), and the MMA thread reads tap (dy, dx) from that tile. Weights: [K/128][C/64][9][128 couts][64 ch] 16-bit.
// The producer keeps the activations one chunk ahead of the weights: staging(c+1) is requested before the nine weight
// blocks of chunk c, so that the conversion of chunk c+1 runs under the MMAs of chunk c.
struct Dir16Smem {
  static constexpr int kSF = 2;  // fp32 staging half-chunks in flight
  static constexpr int kSO = 2;  // 16-bit operand tiles
  static constexpr int kSW = 4;  // weight blocks in flight
  static constexpr int kXRows = DirSmem::kXRows, kXBoxRows = DirSmem::kXBoxRows, kXBoxes = DirSmem::kXBoxes;
  static constexpr uint32_t kFBytes = kXRows * 128;  // 32 fp32 channels per row
  static constexpr uint32_t kOBytes = kXRows * 128;  // 64 16-bit channels per row
  static constexpr uint32_t kWBytes = 128 * 128;     // [128 couts][64 channels] 16-bit
  static constexpr uint32_t kStageOutBytes = 16 * 128;
  static constexpr uint32_t kOffF = 0;
  static constexpr uint32_t kOffO = kOffF + kSF * kFBytes;
  static constexpr uint32_t kOffW = kOffO + kSO * kOBytes;
  static constexpr uint32_t kOffOut = kOffW + kSW * kWBytes;  // [8 warps] one staging tile each
  static constexpr uint32_t kOffZero = kOffOut + 8 * kStageOutBytes;
  static constexpr uint32_t kOffBar = kOffZero + kStageOutBytes;
  static constexpr uint32_t kNumBars = 2 * kSF + 2 * kSO + 2 * kSW + 4;
  static constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr uint32_t kTotal = kOffTmemPtr + 16 + 1024;
  static_assert(kOffO % 1024 == 0 && kOffW % 1024 == 0 && kOffOut % 1024 == 0, "swizzled buffers: 1024-byte aligned");
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};
constexpr int kDir16Threads = 32 * 14;  // producer, MMA, 8 epilogue warps, 4 converter warps

template <bool FP16>
__global__ void __launch_bounds__(kDir16Threads, 1)
conv3x3_direct16_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
                        const uint16_t* __restrict__ w_img, const float* __restrict__ scale,
                        const float* __restrict__ shift, int n_img, int Cin, int Cout, int relu, int out_padded,
                        int n_big, int n_items) {
  using S = Dir16Smem;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* f_full = bars;
  uint64_t* f_empty = f_full + S::kSF;
  uint64_t* o_ready = f_empty + S::kSF;
  uint64_t* o_empty = o_ready + S::kSO;
  uint64_t* w_full = o_empty + S::kSO;
  uint64_t* w_empty = w_full + S::kSW;
  uint64_t* acc_full = w_empty + S::kSW;  // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::kOffTmemPtr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_y);
    for (int i = 0; i < S::kSF; ++i) mbar_init(&f_full[i], 1), mbar_init(&f_empty[i], 4);   // 4 converter warps
    for (int i = 0; i < S::kSO; ++i) mbar_init(&o_ready[i], 4), mbar_init(&o_empty[i], 1);
    for (int i = 0; i < S::kSW; ++i) mbar_init(&w_full[i], 1), mbar_init(&w_empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  if (warp >= 2 && warp < 10) {  // the all-zero frame row
    for (uint32_t i = threadIdx.x - 64; i < S::kStageOutBytes / 16; i += 256)
      reinterpret_cast<uint4*>(smem + S::kOffZero)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_c = Cin / 64;  // 64-channel chunks
  const int n_cb = Cout / 128;
  const int first_item = (int)blockIdx.x, item_stride = (int)gridDim.x;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t sf = 0, pf = 0, sw = 0, pw = 0;
      pdl_wait();  // the frame comes from the previous kernel in the stream
      // staging of (item, chunk): two half-chunks of 32 fp32 channels
      auto load_x = [&](int item, int c) {
        const DirItem w = dir_item(item, n_big, n_cb);
        const int p0 = min(w.grp, n_img - 1) * 256 + 16 + (w.half > 0 ? kDirN / 2 : 0);
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&f_empty[sf], pf ^ 1);
          mbar_arrive_expect_tx(&f_full[sf], S::kFBytes);
#pragma unroll
          for (int b = 0; b < S::kXBoxes; ++b)
            tma_tensor_2d_g2s(smem + S::kOffF + sf * S::kFBytes + b * (S::kXBoxRows * 128), &tmap_x, c * 64 + h * 32,
                              p0 - kDirHalo + b * S::kXBoxRows, &f_full[sf]);
          if (++sf == S::kSF) { sf = 0; pf ^= 1; }
        }
      };
      if (first_item < n_items) load_x(first_item, 0);
      for (int item = first_item; item < n_items; item += item_stride) {
        const DirItem w = dir_item(item, n_big, n_cb);
        const uint8_t* w_src = reinterpret_cast<const uint8_t*>(w_img) + (size_t)w.cb * n_c * 9 * S::kWBytes;
        for (int c = 0; c < n_c; ++c) {
          // activations one chunk ahead of the weights
          if (c + 1 < n_c) load_x(item, c + 1);
          else if (item + item_stride < n_items) load_x(item + item_stride, 0);
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&w_empty[sw], pw ^ 1);
            mbar_arrive_expect_tx(&w_full[sw], S::kWBytes);
            tma_bulk_g2s(smem + S::kOffW + sw * S::kWBytes, w_src + (size_t)(c * 9 + t) * S::kWBytes, S::kWBytes,
                         &w_full[sw]);
            if (++sw == S::kSW) { sw = 0; pw ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t o_base = smem_u32(smem + S::kOffO);
      const uint32_t w_base = smem_u32(smem + S::kOffW);
      uint32_t so = 0, po = 0, sw = 0, pw = 0, it = 0;
      for (int item = first_item; item < n_items; item += item_stride, ++it) {
        const uint32_t idesc = make_idesc(FP16 ? kFmtF16 : kFmtBF16, 128, item < n_big ? kDirN : kDirN / 2);
        const uint32_t buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kDirAccCols;
        for (int c = 0; c < n_c; ++c) {
          mbar_wait(&o_ready[so], po);
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&w_full[sw], pw);
            tc_fence_after();
            const int rshift = (t / 3 - 1) * 16 + (t % 3 - 1);
            const uint32_t x_tap = o_base + so * S::kOBytes + (uint32_t)(kDirHalo + rshift) * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // K = 16 per MMA: 32 bytes of a row
              const uint64_t a_desc = make_smem_desc(w_base + sw * S::kWBytes + k * 32, 0, 1024, kLayoutSW128);
              const uint64_t b_desc = make_smem_desc(x_tap + k * 32, 0, 1024, kLayoutSW128);
              umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, (c > 0 || t > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&w_empty[sw]);
            if (++sw == S::kSW) { sw = 0; pw ^= 1; }
          }
          umma_commit(&o_empty[so]);
          if (++so == S::kSO) { so = 0; po ^= 1; }
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp >= 10) {
    // converter warps: fp32 staging half-chunk -> 16-bit operand tile. Unit = (row r, 16-byte output chunk qq): 8 floats
    // from the two swizzled 16-byte chunks 2qq, 2qq+1 of staging row r -> chunk (4h + qq) ^ (r & 7) of tile row r.
    const int ct = threadIdx.x - 320;
    uint32_t sf = 0, pf = 0, so = 0, po = 0;
    const uint32_t f_base = smem_u32(smem + S::kOffF), o_base = smem_u32(smem + S::kOffO);
    for (int item = first_item; item < n_items; item += item_stride)
      for (int c = 0; c < n_c; ++c) {
        mbar_wait(&o_empty[so], po ^ 1);
        tc_fence_after();
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&f_full[sf], pf);
          const uint32_t src = f_base + sf * S::kFBytes, dst = o_base + so * S::kOBytes;
#pragma unroll 2
          for (int u = ct; u < S::kXRows * 4; u += 128) {
            const int r = u >> 2, qq = u & 3, sw7 = r & 7;
            const float4 a = ld_shared_v4(src + r * 128 + (((2 * qq) ^ sw7) << 4));
            const float4 b = ld_shared_v4(src + r * 128 + (((2 * qq + 1) ^ sw7) << 4));
            uint32_t p0, p1, p2, p3;
            if constexpr (FP16) {
              asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(a.y), "f"(a.x));
              asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(a.w), "f"(a.z));
              asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(b.y), "f"(b.x));
              asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(b.w), "f"(b.z));
            } else {
              asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(a.y), "f"(a.x));
              asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(a.w), "f"(a.z));
              asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(b.y), "f"(b.x));
              asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(b.w), "f"(b.z));
            }
            st_shared_v4(dst + r * 128 + (((4 * h + qq) ^ sw7) << 4), __uint_as_float(p0), __uint_as_float(p1),
                         __uint_as_float(p2), __uint_as_float(p3));
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&f_empty[sf]);  // this warp is done reading the staging slot
          if (++sf == S::kSF) { sf = 0; pf ^= 1; }
        }
        fence_proxy_async_smem();  // the tile was written by the generic proxy, the MMA reads it through the async proxy
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_ready[so]);
        if (++so == S::kSO) { so = 0; po ^= 1; }
      }
  } else {
    // epilogue: as in the TF32 kernel, one staging tile per warp
    const int ew = warp - 2;
    const int quad = warp & 3, hsel = ew >> 2;
    uint8_t* stage_out = smem + S::kOffOut + ew * S::kStageOutBytes;
    const uint32_t stage_u32 = smem_u32(stage_out);
    uint32_t it = 0;
    for (int item = first_item; item < n_items; item += item_stride, ++it) {
      const DirItem w = dir_item(item, n_big, n_cb);
      const int img = min(w.grp, n_img - 1);
      const uint32_t buf = it & 1;
      const int cout0 = w.cb * 128 + quad * 32;
      const float sc = __ldg(scale + cout0 + lane), sh = __ldg(shift + cout0 + lane);
      const int rows = w.half < 0 ? 14 : 7;
      const int y_first = 1 + (w.half > 0 ? 7 : 0);
      const int j0 = hsel ? (rows + 1) / 2 : 0, j1 = hsel ? rows : (rows + 1) / 2;
      if (out_padded == 1 && lane == 0) {
        if (hsel == 0 && w.half <= 0) {
          tma_tensor_2d_s2g(&tmap_y, smem + S::kOffZero, cout0, img * 256);
          tma_store_commit();
        }
        if (hsel == 1 && w.half != 0) {
          tma_tensor_2d_s2g(&tmap_y, smem + S::kOffZero, cout0, img * 256 + 15 * 16);
          tma_store_commit();
        }
      }
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * kDirAccCols;
#pragma unroll 1
      for (int j = j0; j < j1; ++j) {
        float v[16];
        tmem_ld_x16(taddr + j * 16, v);
        if (lane == 0) tma_store_wait_read<0>();  // the previous row's store has read the staging tile
        __syncwarp();
        tmem_ld_wait();
        const uint32_t dst = stage_u32 + lane * 4;
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          float o = fmaf(sc, v[x], sh);
          if (relu) o = fmaxf(o, 0.f);
          if (x == 0 || x == 15) o = 0.f;
          st_shared_f32(dst + x * 128, o);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int yy = y_first + j;
          if (out_padded) tma_tensor_2d_s2g(&tmap_y, stage_out, cout0, img * 256 + yy * 16);
          else tma_tensor_2d_s2g(&tmap_y, stage_out + 128, cout0, img * 196 + (yy - 1) * 14);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if (lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---- 16-bit operands on other map sizes: conv3x3_direct16_kernel's pipeline (fp32 staging -> converter warps -> 16-bit
// operand tile) with conv3x3_direct_gen_kernel's geometry (DirGeo) and epilogue. Stages hold up to 304 rows (28x28: 8 frame
// rows + halo; 56x56: 3 rows; 7x7: two images).
struct Dir16GenSmem {
  static constexpr int kSF = 2, kSO = 2, kSW = 3;
  static constexpr int kXRowsMax = 304;
  static constexpr uint32_t kFBytes = kXRowsMax * 128;
  static constexpr uint32_t kOBytes = kXRowsMax * 128;
  static constexpr uint32_t kWBytes = 128 * 128;
  static constexpr uint32_t kStageOutBytes = 16 * 128;
  static constexpr uint32_t kOffF = 0;
  static constexpr uint32_t kOffO = kOffF + kSF * kFBytes;
  static constexpr uint32_t kOffW = kOffO + kSO * kOBytes;
  static constexpr uint32_t kOffOut = kOffW + kSW * kWBytes;
  static constexpr uint32_t kOffTab = kOffOut + 8 * kStageOutBytes;
  static constexpr uint32_t kOffBar = kOffTab + 8 * 128 * 4;
  static constexpr uint32_t kNumBars = 2 * kSF + 2 * kSO + 2 * kSW + 4;
  static constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr uint32_t kTotal = kOffTmemPtr + 16 + 1024;
  static_assert(kOffO % 1024 == 0 && kOffW % 1024 == 0 && kOffOut % 1024 == 0, "swizzled buffers: 1024-byte aligned");
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

template <bool FP16>
__global__ void __launch_bounds__(kDir16Threads, 1)
conv3x3_direct16_gen_kernel(const __grid_constant__ CUtensorMap tmap_x,
                        const uint16_t* __restrict__ w_img, const float* __restrict__ scale,
                        const float* __restrict__ shift, float* __restrict__ y, int n_img, int Cin, int Cout, int relu,
                        int out_padded, int n_items, const DirGeo g) {
  using S = Dir16GenSmem;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* f_full = bars;
  uint64_t* f_empty = f_full + S::kSF;
  uint64_t* o_ready = f_empty + S::kSF;
  uint64_t* o_empty = o_ready + S::kSO;
  uint64_t* w_full = o_empty + S::kSO;
  uint64_t* w_empty = w_full + S::kSW;
  uint64_t* acc_full = w_empty + S::kSW;  // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::kOffTmemPtr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < S::kSF; ++i) mbar_init(&f_full[i], 1), mbar_init(&f_empty[i], 4);   // 4 converter warps
    for (int i = 0; i < S::kSO; ++i) mbar_init(&o_ready[i], 4), mbar_init(&o_empty[i], 1);
    for (int i = 0; i < S::kSW; ++i) mbar_init(&w_full[i], 1), mbar_init(&w_empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_c = Cin / 64;  // 64-channel chunks
  const int n_cb = Cout / 128;
  const int frame = g.Hf * g.Wf;
  const int x_rows = g.n_boxes * g.box_rows;  // rows of a staging half-chunk / an operand tile
  const int first_item = (int)blockIdx.x, item_stride = (int)gridDim.x;
  // item -> (cout block, first image, first frame row of the range), as in conv3x3_direct_gen_kernel
#define WG_GEN_ITEM(item, cb_, img0_, y0_)            \
  const int cb_ = (item) % n_cb;                      \
  const int u_##cb_ = (item) / n_cb;                  \
  const int img0_ = (u_##cb_ / g.bands) * g.G;        \
  const int y0_ = 1 + (u_##cb_ % g.bands) * g.R

  if (warp == 0) {
    if (elect_one()) {
      uint32_t sf = 0, pf = 0, sw = 0, pw = 0;
      pdl_wait();  // the frame comes from the previous kernel in the stream
      // staging of (item, chunk): two half-chunks of 32 fp32 channels
      auto load_x = [&](int item, int c) {
        WG_GEN_ITEM(item, cb, img0, y0);
        (void)cb;
        const int p0 = img0 * frame + y0 * g.Wf;
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&f_empty[sf], pf ^ 1);
          mbar_arrive_expect_tx(&f_full[sf], (uint32_t)x_rows * 128);
          for (int b = 0; b < g.n_boxes; ++b)
            tma_tensor_2d_g2s(smem + S::kOffF + sf * S::kFBytes + b * (g.box_rows * 128), &tmap_x, c * 64 + h * 32,
                              p0 - g.halo + b * g.box_rows, &f_full[sf]);
          if (++sf == S::kSF) { sf = 0; pf ^= 1; }
        }
      };
      if (first_item < n_items) load_x(first_item, 0);
      for (int item = first_item; item < n_items; item += item_stride) {
        const uint8_t* w_src = reinterpret_cast<const uint8_t*>(w_img) + (size_t)(item % n_cb) * n_c * 9 * S::kWBytes;
        for (int c = 0; c < n_c; ++c) {
          // activations one chunk ahead of the weights
          if (c + 1 < n_c) load_x(item, c + 1);
          else if (item + item_stride < n_items) load_x(item + item_stride, 0);
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&w_empty[sw], pw ^ 1);
            mbar_arrive_expect_tx(&w_full[sw], S::kWBytes);
            tma_bulk_g2s(smem + S::kOffW + sw * S::kWBytes, w_src + (size_t)(c * 9 + t) * S::kWBytes, S::kWBytes,
                         &w_full[sw]);
            if (++sw == S::kSW) { sw = 0; pw ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t o_base = smem_u32(smem + S::kOffO);
      const uint32_t w_base = smem_u32(smem + S::kOffW);
      uint32_t so = 0, po = 0, sw = 0, pw = 0, it = 0;
      const uint32_t idesc = make_idesc(FP16 ? kFmtF16 : kFmtBF16, 128, (uint32_t)g.n_pad);
      for (int item = first_item; item < n_items; item += item_stride, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kDirAccCols;
        for (int c = 0; c < n_c; ++c) {
          mbar_wait(&o_ready[so], po);
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&w_full[sw], pw);
            tc_fence_after();
            const int rshift = (t / 3 - 1) * g.Wf + (t % 3 - 1);
            const uint32_t x_tap = o_base + so * S::kOBytes + (uint32_t)(g.halo + rshift) * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // K = 16 per MMA: 32 bytes of a row
              const uint64_t a_desc = make_smem_desc(w_base + sw * S::kWBytes + k * 32, 0, 1024, kLayoutSW128);
              const uint64_t b_desc = make_smem_desc(x_tap + k * 32, 0, 1024, kLayoutSW128);
              umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, (c > 0 || t > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&w_empty[sw]);
            if (++sw == S::kSW) { sw = 0; pw ^= 1; }
          }
          umma_commit(&o_empty[so]);
          if (++so == S::kSO) { so = 0; po ^= 1; }
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp >= 10) {
    // converter warps: fp32 staging half-chunk -> 16-bit operand tile. Unit = (row r, 16-byte output chunk qq): 8 floats
    // from the two swizzled 16-byte chunks 2qq, 2qq+1 of staging row r -> chunk (4h + qq) ^ (r & 7) of tile row r.
    const int ct = threadIdx.x - 320;
    uint32_t sf = 0, pf = 0, so = 0, po = 0;
    const uint32_t f_base = smem_u32(smem + S::kOffF), o_base = smem_u32(smem + S::kOffO);
    for (int item = first_item; item < n_items; item += item_stride)
      for (int c = 0; c < n_c; ++c) {
        mbar_wait(&o_empty[so], po ^ 1);
        tc_fence_after();
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&f_full[sf], pf);
          const uint32_t src = f_base + sf * S::kFBytes, dst = o_base + so * S::kOBytes;
#pragma unroll 2
          for (int u = ct; u < x_rows * 4; u += 128) {
            const int r = u >> 2, qq = u & 3, sw7 = r & 7;
            const float4 a = ld_shared_v4(src + r * 128 + (((2 * qq) ^ sw7) << 4));
            const float4 b = ld_shared_v4(src + r * 128 + (((2 * qq + 1) ^ sw7) << 4));
            uint32_t p0, p1, p2, p3;
            if constexpr (FP16) {
              asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(a.y), "f"(a.x));
              asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(a.w), "f"(a.z));
              asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(b.y), "f"(b.x));
              asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(b.w), "f"(b.z));
            } else {
              asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(a.y), "f"(a.x));
              asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(a.w), "f"(a.z));
              asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(b.y), "f"(b.x));
              asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(b.w), "f"(b.z));
            }
            st_shared_v4(dst + r * 128 + (((4 * h + qq) ^ sw7) << 4), __uint_as_float(p0), __uint_as_float(p1),
                         __uint_as_float(p2), __uint_as_float(p3));
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&f_empty[sf]);  // this warp is done reading the staging slot
          if (++sf == S::kSF) { sf = 0; pf ^= 1; }
        }
        fence_proxy_async_smem();  // the tile was written by the generic proxy, the MMA reads it through the async proxy
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_ready[so]);
        if (++so == S::kSO) { so = 0; po ^= 1; }
      }
  } else {
    // epilogue: as in conv3x3_direct_gen_kernel (per-warp table of destinations, plain 128-byte row stores)
    const int ew = warp - 2;
    const int quad = warp & 3, hsel = ew >> 2;
    const uint32_t stage_u32 = smem_u32(smem + S::kOffOut + ew * S::kStageOutBytes);
    int* tab = reinterpret_cast<int*>(smem + S::kOffTab) + ew * 128;
    const int n_chunks = g.n_pad / 16;
    const int j0 = hsel ? (n_chunks + 1) / 2 : 0, j1 = hsel ? n_chunks : (n_chunks + 1) / 2;
    const int sub = lane & 7, pxs = lane >> 3;
    uint32_t it = 0;
    for (int item = first_item; item < n_items; item += item_stride, ++it) {
      WG_GEN_ITEM(item, cb, img0, y0);
      const uint32_t buf = it & 1;
      const int cout0 = cb * 128 + quad * 32;
      const float sc = __ldg(scale + cout0 + lane), sh = __ldg(shift + cout0 + lane);
      for (int j = j0 * 16 + lane; j < j1 * 16; j += 32) {
        const int q = y0 * g.Wf + j;
        const int img = img0 + q / frame, rem = q % frame, fy = rem / g.Wf, fx = rem % g.Wf;
        const bool in_batch = img < n_img;
        const bool interior = fy >= 1 && fy <= g.H && fx >= 1 && fx <= g.W;
        int e = -1;
        if (in_batch && out_padded) e = ((img * frame + rem) << 1) | (interior ? 0 : 1);
        else if (in_batch && interior) e = ((img * g.H + fy - 1) * g.W + fx - 1) << 1;
        tab[j - j0 * 16] = e;
      }
      if (out_padded) {
        const int band = ((item / n_cb) % g.bands);
        const int rows_cov = (g.bands == 1) ? g.n_pad / g.Wf : g.R;
        int zimg = -1, zy0 = 0, zy1 = 0;
        if (hsel == 0 && band == 0) zimg = img0, zy0 = 0, zy1 = 1;
        if (hsel == 1 && band == g.bands - 1) {
          const int last = img0 + g.G - 1;
          const int covered_to = (g.bands == 1) ? (1 + rows_cov - (g.G - 1) * g.Hf) : (y0 + rows_cov);
          zimg = last < n_img ? last : -1, zy0 = covered_to > g.H + 1 ? covered_to : g.H + 1, zy1 = g.Hf;
        }
        if (zimg >= 0 && zimg < n_img)
          for (int pz = zy0 * g.Wf + pxs; pz < zy1 * g.Wf; pz += 4)
            *reinterpret_cast<float4*>(y + ((size_t)zimg * frame + pz) * Cout + cout0 + sub * 4) =
                make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * kDirAccCols;
#pragma unroll 1
      for (int j = j0; j < j1; ++j) {
        float v[16];
        tmem_ld_x16(taddr + j * 16, v);
        tmem_ld_wait();
        const uint32_t dst = stage_u32 + lane * 4;
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          float o = fmaf(sc, v[x], sh);
          if (relu) o = fmaxf(o, 0.f);
          st_shared_f32(dst + x * 128, o);
        }
        __syncwarp();
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int pxl = i4 * 4 + pxs;
          const int e = tab[(j - j0) * 16 + pxl];
          if (e >= 0) {
            float4 val = ld_shared_v4(stage_u32 + pxl * 128 + sub * 16);
            if (e & 1) val = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(y + (size_t)(e >> 1) * Cout + cout0 + sub * 4) = val;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
#undef WG_GEN_ITEM

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// Once per layer, 16-bit operands: [K/128][C/64][9 taps][128 couts][64 channels], RN, 128-byte swizzle applied
// (16-byte chunk q = 8 channels at position q ^ (cout & 7) of the cout's 128-byte row).
__global__ void direct_pack16_kernel(const float* __restrict__ w, uint16_t* __restrict__ w_img, int Cin, int Cout,
                                     int fp16) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cin * Cout * 9) return;
  const int t = idx % 9;
  const int ci = (idx / 9) % Cin;
  const int co = idx / (9 * Cin);
  const int nt = co / 128, r = co % 128;
  const int c = ci / 64, kk = ci % 64;
  const int q = (kk >> 3) ^ (r & 7);
  const size_t blk = ((size_t)nt * (Cin / 64) + c) * 9 + t;
  uint16_t bits;
  if (fp16) {
    const __half hv = __float2half_rn(w[idx]);
    bits = *reinterpret_cast<const uint16_t*>(&hv);
  } else {
    const __nv_bfloat16 bv = __float2bfloat16_rn(w[idx]);
    bits = *reinterpret_cast<const uint16_t*>(&bv);
  }
  w_img[blk * (size_t)(128 * 64) + (size_t)r * 64 + q * 8 + (kk & 7)] = bits;
}

// Once per layer: w[K][C][3][3] (the reference's filter layout, Kernel128_winograd.cu:274-281 reads it the same way)
// -> per (128-cout block, 32-channel chunk, tap) the K-major 128-byte-swizzled block [128 couts][32 channels], RN-rounded
// to TF32.
__global__ void direct_pack_kernel(const float* __restrict__ w, float* __restrict__ w_img, int Cin, int Cout) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int cout_pad = (Cout + 127) / 128 * 128;  // rows beyond Cout: zeros
  if (idx >= Cin * cout_pad * 9) return;
  const int t = idx % 9;
  const int ci = (idx / 9) % Cin;
  const int co = idx / (9 * Cin);
  const int nt = co / 128, r = co % 128;
  const int c = ci / 32, kk = ci % 32;
  const int q = (kk >> 2) ^ (r & 7);
  const size_t blk = ((size_t)nt * (Cin / 32) + c) * 9 + t;
  w_img[blk * (size_t)(128 * 32) + (size_t)r * 32 + q * 4 + (kk & 3)] = co < Cout ? to_tf32_rn(w[idx]) : 0.f;
}

// op16: 0 = TF32 image (fp32 words), 1 = bf16, 2 = fp16
int direct_pack_launch(const float* w, float* w_img, int Cin, int Cout, int op16, cudaStream_t stream) {
  const int n = Cin * (op16 ? Cout : (Cout + 127) / 128 * 128) * 9;
  if (op16) direct_pack16_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w, reinterpret_cast<uint16_t*>(w_img), Cin, Cout,
                                                                    op16 == 2);
  else direct_pack_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w, w_img, Cin, Cout);
  return cudaGetLastError() == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

static int dir_encode(CUtensorMap* tmap, const float* base, int inner, long long rows, int box_inner, int box_rows,
                      CUtensorMapSwizzle swz) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return WG_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)inner * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  // FLOAT_OOB_FILL_NONE: rows outside the tensor (the halo of the first / last image) read as zeros
  CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, wg::l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WG_OK : WG_ERR_TMAP;
}

// activations: the frames as [N*256 rows][Cin]; box = 32 channels x half of the rows a CTA holds per chunk
int direct_make_tmap_in(CUtensorMap* tmap, const float* x, int n_img, int Cin) {
  return dir_encode(tmap, x, Cin, (long long)n_img * 256, 32, DirSmem::kXBoxRows, CU_TENSOR_MAP_SWIZZLE_128B);
}
// output: one frame row of 32 couts per store -- 14 pixels into the dense [N*196][Cout] map or 16 into the
// [N*256][Cout] frame; the staging tile is plain [px][32 couts]
int direct_make_tmap_out(CUtensorMap* tmap, const float* y, int n_img, int Cout, int out_padded) {
  return dir_encode(tmap, y, Cout, (long long)n_img * (out_padded ? 256 : 196), 32, out_padded ? 16 : 14,
                    CU_TENSOR_MAP_SWIZZLE_NONE);
}

int direct_gen_make_tmap_in(CUtensorMap* tmap, const float* x, int n_img, int Cin, const DirGeo& g) {
  return dir_encode(tmap, x, Cin, (long long)n_img * g.Hf * g.Wf, 32, g.box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

int direct_gen_launch(const CUtensorMap& tmap_x, const float* w_img, const float* scale, const float* shift, float* y,
                      int n_img, int Cin, int Cout, int relu, int out_padded, int max_ctas, const DirGeo& g,
                      cudaStream_t stream) {
  using S = DirGenSmem;
  if (Cin % 32 != 0 || Cout % 64 != 0) return WG_ERR_ARG;
  static unsigned long long configured = 0;
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long bit = 1ull << (dev_ & 63);
  if (!(configured & bit)) {
    if (cudaFuncSetAttribute(conv3x3_direct_gen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kTotal) !=
        cudaSuccess)
      return WG_ERR_CUDA;
    configured |= bit;
  }
  const long long units = ((long long)n_img + g.G - 1) / g.G * g.bands;
  const long long n_items = units * ((Cout + 127) / 128);
  long long grid = n_items < max_ctas ? n_items : max_ctas;
  if (grid < 1) grid = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kDirThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_direct_gen_kernel, tmap_x, w_img, scale, shift, y, n_img, Cin, Cout,
                                     relu, out_padded, (int)n_items, g);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

template <bool FP16>
static int launch_direct16_gen(const CUtensorMap& tmap_x, const float* w_img, const float* scale, const float* shift,
                               float* y, int n_img, int Cin, int Cout, int relu, int out_padded, int max_ctas,
                               const DirGeo& g, cudaStream_t stream) {
  using S = Dir16GenSmem;
  static unsigned long long configured = 0;
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long bit = 1ull << (dev_ & 63);
  if (!(configured & bit)) {
    if (cudaFuncSetAttribute(conv3x3_direct16_gen_kernel<FP16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)S::kTotal) != cudaSuccess)
      return WG_ERR_CUDA;
    configured |= bit;
  }
  const long long units = ((long long)n_img + g.G - 1) / g.G * g.bands;
  const long long n_items = units * (Cout / 128);
  long long grid = n_items < max_ctas ? n_items : max_ctas;
  if (grid < 1) grid = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kDir16Threads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_direct16_gen_kernel<FP16>, tmap_x,
                                     reinterpret_cast<const uint16_t*>(w_img), scale, shift, y, n_img, Cin, Cout, relu,
                                     out_padded, (int)n_items, g);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

// 16-bit operands on other map sizes (op16: 1 = bf16, 2 = fp16): Cin % 64 == 0, Cout % 128 == 0, g from
// direct_gen_geo(..., kDirect16GenMaxRows, ...)
int direct16_gen_launch(const CUtensorMap& tmap_x, const float* w_img, const float* scale, const float* shift, float* y,
                        int n_img, int Cin, int Cout, int op16, int relu, int out_padded, int max_ctas, const DirGeo& g,
                        cudaStream_t stream) {
  if (Cin % 64 != 0 || Cout % 128 != 0 || (op16 != 1 && op16 != 2)) return WG_ERR_ARG;
  if (g.n_boxes * g.box_rows > Dir16GenSmem::kXRowsMax) return WG_ERR_ARG;
  if (op16 == 2)
    return launch_direct16_gen<true>(tmap_x, w_img, scale, shift, y, n_img, Cin, Cout, relu, out_padded, max_ctas, g, stream);
  return launch_direct16_gen<false>(tmap_x, w_img, scale, shift, y, n_img, Cin, Cout, relu, out_padded, max_ctas, g, stream);
}

template <int CL>
static int launch_direct(const CUtensorMap& tmap_x, const CUtensorMap& tmap_y, const float* w_img, const float* scale,
                         const float* shift, int n_img, int Cin, int Cout, int relu, int out_padded, int max_ctas,
                         bool mixed, cudaStream_t stream) {
  using S = DirSmem;
  static unsigned long long configured = 0;
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long bit = 1ull << (dev_ & 63);
  if (!(configured & bit)) {
    if (cudaFuncSetAttribute(conv3x3_direct_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kTotal) !=
        cudaSuccess)
      return WG_ERR_CUDA;
    configured |= bit;
  }
  const int n_cb = (Cout + 127) / 128;
  const long long n_grp = ((long long)n_img + CL - 1) / CL;  // groups of CL images
  long long n_clusters = max_ctas / CL;
  if (n_clusters < 1) n_clusters = 1;
  // whole images for as many full rounds of the grid as the batch has, half images for the rest when they fit one more
  // round
  long long n_big = n_grp * n_cb;
  if (mixed) {
    const long long per_round = n_clusters / n_cb;  // image groups per round
    if (per_round > 0) {
      const long long whole = (n_grp / per_round) * per_round;
      const long long rest = n_grp - whole;
      if (rest > 0 && rest * 2 * n_cb <= n_clusters) n_big = whole * n_cb;
    }
  }
  const long long n_items = n_big + (n_grp * n_cb - n_big) * 2;
  if (n_clusters > n_items) n_clusters = n_items;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_clusters * CL));
  cfg.blockDim = dim3(kDirThreads);
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_direct_kernel<CL>, tmap_x, tmap_y, w_img, scale, shift, n_img, Cin,
                                     Cout, relu, out_padded, (int)n_big, (int)n_items);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

template <bool FP16>
static int launch_direct16(const CUtensorMap& tmap_x, const CUtensorMap& tmap_y, const float* w_img, const float* scale,
                           const float* shift, int n_img, int Cin, int Cout, int relu, int out_padded, int max_ctas,
                           bool mixed, cudaStream_t stream) {
  using S = Dir16Smem;
  static unsigned long long configured = 0;
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long bit = 1ull << (dev_ & 63);
  if (!(configured & bit)) {
    if (cudaFuncSetAttribute(conv3x3_direct16_kernel<FP16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)S::kTotal) != cudaSuccess)
      return WG_ERR_CUDA;
    configured |= bit;
  }
  const int n_cb = Cout / 128;
  long long n_ctas = max_ctas < 1 ? 1 : max_ctas;
  long long n_big = (long long)n_img * n_cb;
  if (mixed) {
    const long long per_round = n_ctas / n_cb;
    if (per_round > 0) {
      const long long whole = ((long long)n_img / per_round) * per_round;
      const long long rest = n_img - whole;
      if (rest > 0 && rest * 2 * n_cb <= n_ctas) n_big = whole * n_cb;
    }
  }
  const long long n_items = n_big + ((long long)n_img * n_cb - n_big) * 2;
  if (n_ctas > n_items) n_ctas = n_items;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)n_ctas);
  cfg.blockDim = dim3(kDir16Threads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_direct16_kernel<FP16>, tmap_x, tmap_y,
                                     reinterpret_cast<const uint16_t*>(w_img), scale, shift, n_img, Cin, Cout, relu,
                                     out_padded, (int)n_big, (int)n_items);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

// 16-bit operands (op16: 1 = bf16, 2 = fp16): Cin % 64 == 0, Cout % 128 == 0
int direct16_launch(const CUtensorMap& tmap_x, const CUtensorMap& tmap_y, const float* w_img, const float* scale,
                    const float* shift, int n_img, int Cin, int Cout, int op16, int relu, int out_padded, int max_ctas,
                    int mixed, cudaStream_t stream) {
  if (Cin % 64 != 0 || Cout % 128 != 0 || (op16 != 1 && op16 != 2)) return WG_ERR_ARG;
  if (op16 == 2)
    return launch_direct16<true>(tmap_x, tmap_y, w_img, scale, shift, n_img, Cin, Cout, relu, out_padded, max_ctas,
                                 mixed != 0, stream);
  return launch_direct16<false>(tmap_x, tmap_y, w_img, scale, shift, n_img, Cin, Cout, relu, out_padded, max_ctas,
                                mixed != 0, stream);
}

// out_padded: 0 = dense map, 1 = frame with its zero border, 3 = frame, interior rows only (WG_OUT_INTERIOR_ONLY)
int direct_launch(const CUtensorMap& tmap_x, const CUtensorMap& tmap_y, const float* w_img, const float* scale,
                  const float* shift, int n_img, int Cin, int Cout, int cl, int relu, int out_padded, int max_ctas,
                  int mixed, cudaStream_t stream) {
  if (Cin % 32 != 0 || Cout % 64 != 0) return WG_ERR_ARG;
#define WG_DIR(CL_)                                                                                                  \
  return launch_direct<CL_>(tmap_x, tmap_y, w_img, scale, shift, n_img, Cin, Cout, relu, out_padded, max_ctas, mixed != 0, \
                            stream)
  if constexpr (kDev) {  // weight-multicast clusters: measured no faster (profiles/direct3x3_r02.md), developer build only
    if (cl == 4) WG_DIR(4);
    if (cl == 2) WG_DIR(2);
  }
  WG_DIR(1);
#undef WG_DIR
}

}  // namespace wg

#ifdef WG_DEV_BUILD
// developer entry points: the direct kernel on caller-owned device buffers (tools/direct_check.py)
extern "C" int wg_dev_direct_pack(const float* w_dev, float* w_img_dev, int cin, int cout, int op16) {
  int rc = wg::direct_pack_launch(w_dev, w_img_dev, cin, cout, op16, nullptr);
  return cudaDeviceSynchronize() == cudaSuccess ? rc : WG_ERR_CUDA;
}
// mode: low 3 bits = cluster size (1, 2, 4), bit 3 = no mixed schedule
extern "C" int wg_dev_direct_run(const float* x_dev, const float* w_img_dev, const float* scale_dev,
                                 const float* shift_dev, float* y_dev, int n_img, int cin, int cout, int relu,
                                 int out_padded, int max_ctas, int mode, void* stream) {
  CUtensorMap tx, ty;
  int rc = wg::direct_make_tmap_in(&tx, x_dev, n_img, cin);
  if (rc != WG_OK) return rc;
  rc = wg::direct_make_tmap_out(&ty, y_dev, n_img, cout, out_padded);
  if (rc != WG_OK) return rc;
  if (mode & 0x300)  // bits 8 / 9: bf16 / fp16 operands
    return wg::direct16_launch(tx, ty, w_img_dev, scale_dev, shift_dev, n_img, cin, cout, (mode & 0x200) ? 2 : 1, relu,
                               out_padded, max_ctas, (mode & 8) ? 0 : 1, static_cast<cudaStream_t>(stream));
  return wg::direct_launch(tx, ty, w_img_dev, scale_dev, shift_dev, n_img, cin, cout, mode & 7, relu, out_padded,
                           max_ctas, (mode & 8) ? 0 : 1, static_cast<cudaStream_t>(stream));
}
#endif
