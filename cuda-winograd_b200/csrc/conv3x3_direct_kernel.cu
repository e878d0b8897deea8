// 3x3 conv + folded BatchNorm (+ ReLU) as a DIRECT convolution on the tensor core (implicit GEMM without an im2col
// buffer and without any CUDA-core transform), sm_100a.
//
// Same operator as the fused Winograd kernels (/root/reference/Kernel128_winograd.cu:153-262, Kernel256_winograd.cu:
// 176-296: input = zero-bordered [N][16][16][C] frame, weights [K][C][3][3], output [N][14][14][K] or the next layer's
// frame), different decomposition: with the frame flattened to rows m = (n*16 + y)*16 + x of C channels, the output at
// row m is  sum over the 9 taps (dy, dx) of  X[m + (dy-1)*16 + (dx-1)] . W[dy][dx]  -- nine GEMMs whose A operands are
// the SAME rows shifted by a constant. One TMA box per 32-channel chunk brings rows m0-24 .. m0+151 of the M-tile
// (128 rows + halo) into shared memory ONCE; each tap's tcgen05.mma reads it through a descriptor whose start address is
// moved by (dy-1)*16 + (dx-1) rows. Rows of the frame's border (y or x in {0, 15}) compute garbage that is never stored
// (dense output) or stored as the zeros the next layer's frame needs (padded output): 196 of every 256 rows are useful.
//
// Persistent, warp-specialised like the 1x1 kernel: warp 0 = TMA producer (A box per chunk, one bulk copy of the
// pre-swizzled [BN couts][32 channels] weight block per (chunk, tap)), warp 1 = one thread issuing 4 MMAs (M = 128,
// N = BN, K = 8) per (chunk, tap) into one of two TMEM accumulator buffers, warps 2..5 = epilogue of the previous item.
#include <stdlib.h>

#include "ptx.cuh"
#include "wg_internal.h"

namespace wg {

constexpr int kDirThreads = 32 * 10;           // producer, MMA, 8 epilogue warps (two per TMEM lane quadrant)
constexpr int kDirHalo = 24;                   // rows in front of the M-tile: >= 17 (one frame row + 1), multiple of 8
constexpr int kDirARows = 128 + 2 * kDirHalo;  // 176 rows of 128 B per chunk

template <int BN, bool PAIR = false>
struct DirSmem {
  static constexpr int kSA = BN == 256 ? 2 : 3;  // activation chunks in flight
  static constexpr int kSB = PAIR ? (BN == 256 ? 7 : 12) : (BN == 256 ? 4 : 7);  // weight blocks in flight
  static constexpr uint32_t kABytes = kDirARows * 128;
  static constexpr uint32_t kBBytes = BN * (PAIR ? 64 : 128);  // PAIR: this CTA's half of the block's cout rows
  static constexpr uint32_t kStageOutBytes = 32 * 128;
  static constexpr uint32_t kOffA = 0;
  static constexpr uint32_t kOffB = kOffA + kSA * kABytes;
  static constexpr uint32_t kOffOut = kOffB + kSB * kBBytes;  // [8 warps] one 32 x 32 fp32 sub-tile each
  static constexpr uint32_t kOffBar = kOffOut + 8 * kStageOutBytes;
  static constexpr uint32_t kNumBars = 2 * kSA + 2 * kSB + 4;
  static constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr uint32_t kOffTab = kOffTmemPtr + 16;  // dense output: [8 warps][32 rows] pixel index (or -1)
  static constexpr uint32_t kTotal = kOffTab + 8 * 32 * 4 + 1024;
  static_assert(kOffB % 1024 == 0 && kOffOut % 1024 == 0, "swizzled buffers must be 1024-byte aligned");
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

// Work items. The first n_big items are full-width (BN couts), N-tile fastest; the rest are half-width (BN/2 couts) --
// the host sizes n_big so that the full-width items fill whole rounds of the grid and what is left of the batch is
// spread over all CTAs as half-width items (a partial last round costs half an item instead of a whole one).
struct DirItem {
  int unit;   // frame (PAIR) or 128-row M-tile
  int col0;   // first cout
  int width;  // couts of this item: BN or BN / 2
};
template <int BN>
__device__ __forceinline__ DirItem dir_item(int i, int n_big, int n_nt) {
  if (i < n_big) return DirItem{i / n_nt, (i % n_nt) * BN, BN};
  const int j = i - n_big;
  return DirItem{n_big / n_nt + j / (2 * n_nt), (j % (2 * n_nt)) * (BN / 2), BN / 2};
}

// PAIR: the two CTAs of a cluster form a tcgen05 cta_group::2 pair on ONE frame (256 rows = two M-tiles): the leader
// issues M = 256 MMAs for both, each CTA loads its own 128 + 48 activation rows and only HALF of every weight block
// (cout rows [rank*width/2, +width/2)); the tensor core reads each half from the shared memory it lives in. Per SM the
// weight stream -- which is what the un-paired kernel is bound by (60 B/clk and SM at full MMA rate, above the L2's
// ~12 TB/s for 148 SMs) -- is halved. Hand-offs as in the 1x1 kernel's pair variant: the peer's MMA warp relays "landed"
// to the leader's full barriers, commits are multicast to both CTAs, the peer's epilogue warps arrive on the leader's
// acc_empty.
template <int BN, bool PAIR>
__global__ void __launch_bounds__(kDirThreads, 1)
conv3x3_direct_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_y,
                      const float* __restrict__ w_img, const float* __restrict__ scale, const float* __restrict__ shift,
                      float* __restrict__ y, int n_img, int Cin, int Cout, int relu, int out_padded, int n_big,
                      int n_items) {
  using S = DirSmem<BN, PAIR>;
  constexpr uint32_t kTmemCols = 2 * BN;
  constexpr uint16_t kPairMask = 0x3;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + S::kSA;
  uint64_t* b_full = a_empty + S::kSA;
  uint64_t* b_empty = b_full + S::kSB;
  uint64_t* acc_full = b_empty + S::kSB;  // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::kOffTmemPtr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_y);
    const uint32_t n_full = (PAIR && crank == 0) ? 2 : 1;  // leader: own TMA bytes + the peer's relay
    for (int i = 0; i < S::kSA; ++i) mbar_init(&a_full[i], n_full), mbar_init(&a_empty[i], 1);
    for (int i = 0; i < S::kSB; ++i) mbar_init(&b_full[i], n_full), mbar_init(&b_empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], PAIR ? 16 : 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_cg2<kTmemCols>(tmem_ptr);
    else tmem_alloc<kTmemCols>(tmem_ptr);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_c = Cin / 32;  // 32-channel chunks
  const int n_nt = Cout / BN;
  const int first_item = PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x;
  const int item_stride = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;
  constexpr uint32_t kRowBytes = PAIR ? 64 : 128;  // bytes of a weight block this CTA loads per cout of the item

  if (warp == 0) {
    if (elect_one()) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      pdl_wait();  // the frame comes from the previous kernel in the stream
      for (int item = first_item; item < n_items; item += item_stride) {
        const DirItem w = dir_item<BN>(item, n_big, n_nt);
        const int mt = PAIR ? w.unit * 2 + (int)crank : w.unit;
        // this CTA's rows of the [BN couts][128 B] block of (N-tile, chunk, tap)
        const int r0 = w.col0 % BN + (PAIR ? (int)crank * (w.width / 2) : 0);
        const uint8_t* b_src =
            reinterpret_cast<const uint8_t*>(w_img) + ((size_t)(w.col0 / BN) * n_c * 9 * BN + r0) * 128;
        const uint32_t b_bytes = (uint32_t)w.width * kRowBytes;
        for (int c = 0; c < n_c; ++c) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          mbar_arrive_expect_tx(&a_full[sa], S::kABytes);
          tma_tensor_2d_g2s(smem + S::kOffA + sa * S::kABytes, &tmap_a, c * 32, mt * 128 - kDirHalo, &a_full[sa]);
          if (++sa == S::kSA) { sa = 0; pa ^= 1; }
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&b_empty[sb], pb ^ 1);
            mbar_arrive_expect_tx(&b_full[sb], b_bytes);
            tma_bulk_g2s(smem + S::kOffB + sb * S::kBBytes, b_src + (size_t)(c * 9 + t) * (BN * 128), b_bytes,
                         &b_full[sb]);
            if (++sb == S::kSB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && PAIR && crank != 0) {
    // peer of a pair: no MMAs to issue; relay "landed here" to the leader's barriers, in consumption order
    if (elect_one()) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      for (int item = first_item; item < n_items; item += item_stride)
        for (int c = 0; c < n_c; ++c) {
          mbar_wait(&a_full[sa], pa);
          mbar_arrive_remote_plain(&a_full[sa], 0);
          if (++sa == S::kSA) { sa = 0; pa ^= 1; }
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&b_full[sb], pb);
            mbar_arrive_remote_plain(&b_full[sb], 0);
            if (++sb == S::kSB) { sb = 0; pb ^= 1; }
          }
        }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t a_base = smem_u32(smem + S::kOffA);
      const uint32_t b_base = smem_u32(smem + S::kOffB);
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
      for (int item = first_item; item < n_items; item += item_stride, ++it) {
        const uint32_t idesc = make_idesc(kFmtTF32, PAIR ? 256 : 128, item < n_big ? BN : BN / 2);
        const uint32_t buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int c = 0; c < n_c; ++c) {
          mbar_wait(&a_full[sa], pa);
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            // rows of tap (dy, dx): the tile's rows moved by (dy-1)*16 + (dx-1) frame pixels. The start address is then
            // 128-byte but not 1024-byte aligned; the 128-byte swizzle is a function of the ADDRESS bits (7..9 into
            // 4..6) for the TMA write and the MMA read alike, so the shifted descriptor reads the rows as written
            // (measured: bit-identical to the aligned case; the descriptor's base-offset field stays 0).
            const int rshift = (t / 3 - 1) * 16 + (t % 3 - 1);
            const uint32_t a_tap = a_base + sa * S::kABytes + (uint32_t)(kDirHalo + rshift) * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a_desc = make_smem_desc(a_tap + k * 32, 0, 1024, kLayoutSW128);
              const uint64_t b_desc = make_smem_desc(b_base + sb * S::kBBytes + k * 32, 0, 1024, kLayoutSW128);
              const uint32_t acc = (c > 0 || t > 0 || k > 0) ? 1u : 0u;
              if constexpr (PAIR) umma_tf32_ss_cg2(tmem_base + buf * BN, a_desc, b_desc, idesc, acc);
              else umma_tf32_ss(tmem_base + buf * BN, a_desc, b_desc, idesc, acc);
            }
            if constexpr (PAIR) umma_commit_mcast_cg2(&b_empty[sb], kPairMask);
            else umma_commit(&b_empty[sb]);
            if (++sb == S::kSB) { sb = 0; pb ^= 1; }
          }
          if constexpr (PAIR) umma_commit_mcast_cg2(&a_empty[sa], kPairMask);
          else umma_commit(&a_empty[sa]);
          if (++sa == S::kSA) { sa = 0; pa ^= 1; }
        }
        if constexpr (PAIR) umma_commit_mcast_cg2(&acc_full[buf], kPairMask);
        else umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // epilogue: warps 2..9; warp & 3 = TMEM lane quadrant, (warp - 2) / 4 = which half of the item's couts
    const int ew = warp - 2;
    const int quad = warp & 3, hsel = ew >> 2;
    uint8_t* stage_out = smem + S::kOffOut + ew * S::kStageOutBytes;
    const uint32_t stage_u32 = smem_u32(stage_out);
    int* tab = reinterpret_cast<int*>(smem + S::kOffTab) + ew * 32;
    uint32_t it = 0;
    for (int item = first_item; item < n_items; item += item_stride, ++it) {
      const DirItem w = dir_item<BN>(item, n_big, n_nt);
      const int mt = PAIR ? w.unit * 2 + (int)crank : w.unit;
      const uint32_t buf = it & 1;
      const int ncol = w.width / 2;                 // this warp's couts of the item
      const int colg = w.col0 + hsel * ncol;        // first of them in the layer
      const float* sc = scale + colg;
      const float* sh = shift + colg;
      // this thread's row: frame pixel m -> interior or border
      const int m = mt * 128 + quad * 32 + lane;
      const int fy = (m >> 4) & 15, fx = m & 15;
      const bool interior = fy >= 1 && fy <= 14 && fx >= 1 && fx <= 14;
      if (!out_padded) {
        tab[lane] = interior ? (((m >> 8) * 14 + fy - 1) * 14 + fx - 1) : -1;
        __syncwarp();
      }
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BN + hsel * ncol;
#pragma unroll 1
      for (int c0 = 0; c0 < ncol; c0 += 32) {
        float4 s4[8], h4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s4[j] = __ldg(reinterpret_cast<const float4*>(sc + c0 + 4 * j));
          h4[j] = __ldg(reinterpret_cast<const float4*>(sh + c0 + 4 * j));
        }
        float v[32];
        tmem_ld_x32(taddr + c0, v);
        if (out_padded) {
          if (lane == 0) tma_store_wait_read<0>();  // the previous chunk's store has read the staging buffer
          __syncwarp();
        }
        tmem_ld_wait();
        const uint32_t dst = stage_u32 + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o;
          o.x = fmaf(s4[j].x, v[4 * j + 0], h4[j].x);
          o.y = fmaf(s4[j].y, v[4 * j + 1], h4[j].y);
          o.z = fmaf(s4[j].z, v[4 * j + 2], h4[j].z);
          o.w = fmaf(s4[j].w, v[4 * j + 3], h4[j].w);
          if (relu) {
            o.x = fmaxf(o.x, 0.f);
            o.y = fmaxf(o.y, 0.f);
            o.z = fmaxf(o.z, 0.f);
            o.w = fmaxf(o.w, 0.f);
          }
          if (!interior) o = make_float4(0.f, 0.f, 0.f, 0.f);  // the frame's zero border (padded output)
          st_shared_v4(dst + ((j ^ (lane & 7)) << 4), o.x, o.y, o.z, o.w);
        }
        if (out_padded) {
          // next layer's frame: same row numbering as the input -> one 32 x 32 TMA tensor store per warp and chunk
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_tensor_2d_s2g(&tmap_y, stage_out, colg + c0, mt * 128 + quad * 32);
            tma_store_commit();
          }
        } else {
          // dense [N][14][14][K]: interior rows only, 8 lanes write one pixel's 128 bytes, 4 pixels per instruction
          __syncwarp();
          const int j = lane & 7, rsub = lane >> 3;
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int r = i8 * 4 + rsub;
            const int px = tab[r];
            if (px >= 0) {
              const float4 val = ld_shared_v4(stage_u32 + r * 128 + ((j ^ (r & 7)) << 4));
              *reinterpret_cast<float4*>(y + (size_t)px * Cout + colg + c0 + j * 4) = val;
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && crank != 0) mbar_arrive_remote_plain(&acc_empty[buf], 0);  // the leader issues the MMAs of both CTAs
        else mbar_arrive(&acc_empty[buf]);
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();  // the peer's shared memory and barriers stay alive
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc_cg2<kTmemCols>(tmem_base);
    else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// Once per layer: w[K][C][3][3] (the reference's filter layout, Kernel128_winograd.cu:274-281 reads it the same way)
// -> per (N-tile, 32-channel chunk, tap) the K-major 128-byte-swizzled block [BN couts][32 channels], RN-rounded to TF32.
__global__ void direct_pack_kernel(const float* __restrict__ w, float* __restrict__ w_img, int Cin, int Cout, int BN) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cin * Cout * 9) return;
  const int t = idx % 9;
  const int ci = (idx / 9) % Cin;
  const int co = idx / (9 * Cin);
  const int nt = co / BN, r = co % BN;
  const int c = ci / 32, kk = ci % 32;
  const int q = (kk >> 2) ^ (r & 7);
  const size_t blk = ((size_t)nt * (Cin / 32) + c) * 9 + t;
  w_img[blk * (size_t)(BN * 32) + (size_t)r * 32 + q * 4 + (kk & 3)] = to_tf32_rn(w[idx]);
}

int direct_pack_launch(const float* w, float* w_img, int Cin, int Cout, int BN, cudaStream_t stream) {
  const int n = Cin * Cout * 9;
  direct_pack_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w, w_img, Cin, Cout, BN);
  return cudaGetLastError() == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

static int dir_encode(CUtensorMap* tmap, const float* base, int inner, long long rows, int box_inner, int box_rows,
                      bool zero_fill) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return WG_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)inner * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  (void)zero_fill;  // FLOAT_OOB_FILL_NONE fills out-of-range rows with zeros
  CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, wg::l2_promotion(),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WG_OK : WG_ERR_TMAP;
}

int direct_make_tmap_in(CUtensorMap* tmap, const float* x, int n_img, int Cin) {
  return dir_encode(tmap, x, Cin, (long long)n_img * 256, 32, kDirARows, true);
}
int direct_make_tmap_out(CUtensorMap* tmap, const float* y, int n_img, int Cout) {
  return dir_encode(tmap, y, Cout, (long long)n_img * 256, 32, 32, false);
}

template <int BN, bool PAIR>
static int launch_direct(const CUtensorMap& tmap_a, const CUtensorMap& tmap_y, const float* w_img, const float* scale,
                         const float* shift, float* y, int n_img, int Cin, int Cout, int relu, int out_padded,
                         int max_ctas, bool mixed, cudaStream_t stream) {
  using S = DirSmem<BN, PAIR>;
  static unsigned long long configured = 0;
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long bit = 1ull << (dev_ & 63);
  if (!(configured & bit)) {
    if (cudaFuncSetAttribute(conv3x3_direct_kernel<BN, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)S::kTotal) != cudaSuccess)
      return WG_ERR_CUDA;
    configured |= bit;
  }
  constexpr int CL = PAIR ? 2 : 1;
  const int n_nt = Cout / BN;
  const long long n_units = (long long)n_img * (PAIR ? 1 : 2);  // frames (PAIR) or 128-row M-tiles
  long long n_clusters = max_ctas / CL;
  if (n_clusters < 1) n_clusters = 1;
  // full-width items for as many whole rounds of the grid as the batch has, half-width items for the rest when they fit
  // one more round (bo_mode bit 2 of the developer entry switches the mixed schedule off)
  long long n_big = n_units * n_nt;
  if (mixed && BN / 2 >= 64) {
    const long long units_per_round = n_clusters / n_nt;
    if (units_per_round > 0) {
      const long long whole = (n_units / units_per_round) * units_per_round;
      const long long rest = n_units - whole;
      if (rest > 0 && whole > 0 && rest * 2 * n_nt <= n_clusters) n_big = whole * n_nt;
    }
  }
  const long long n_items = n_big + (n_units * n_nt - n_big) * 2;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_direct_kernel<BN, PAIR>, tmap_a, tmap_y, w_img, scale, shift, y, n_img, Cin,
                                     Cout, relu, out_padded, (int)n_big, (int)n_items);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

int direct_launch(const CUtensorMap& tmap_a, const CUtensorMap& tmap_y, const float* w_img, const float* scale,
                  const float* shift, float* y, int n_img, int Cin, int Cout, int BN, int relu, int out_padded,
                  int max_ctas, int bo_mode, cudaStream_t stream) {
  if (Cin % 32 != 0 || Cout % BN != 0) return WG_ERR_ARG;
  const bool pair = (bo_mode & 2) != 0;
  const bool mixed = (bo_mode & 4) == 0;
#define WG_DIR(BN_, P_) \
  return launch_direct<BN_, P_>(tmap_a, tmap_y, w_img, scale, shift, y, n_img, Cin, Cout, relu, out_padded, max_ctas, \
                                mixed, stream)
  if (BN == 256 && pair) WG_DIR(256, true);
  if (BN == 256) WG_DIR(256, false);
  if (BN == 128 && pair) WG_DIR(128, true);
  if (BN == 128) WG_DIR(128, false);
#undef WG_DIR
  return WG_ERR_ARG;
}

}  // namespace wg

#ifdef WG_DEV_BUILD
// developer entry points: the direct kernel on caller-owned device buffers (tools/direct_check.py)
extern "C" int wg_dev_direct_pack(const float* w_dev, float* w_img_dev, int cin, int cout, int bn) {
  int rc = wg::direct_pack_launch(w_dev, w_img_dev, cin, cout, bn, nullptr);
  return cudaDeviceSynchronize() == cudaSuccess ? rc : WG_ERR_CUDA;
}
extern "C" int wg_dev_direct_run(const float* x_dev, const float* w_img_dev, const float* scale_dev,
                                 const float* shift_dev, float* y_dev, int n_img, int cin, int cout, int bn, int relu,
                                 int out_padded, int max_ctas, int bo_mode, void* stream) {
  CUtensorMap ta, ty;
  int rc = wg::direct_make_tmap_in(&ta, x_dev, n_img, cin);
  if (rc != WG_OK) return rc;
  rc = wg::direct_make_tmap_out(&ty, y_dev, out_padded ? n_img : 1, cout);  // only used for the padded output
  if (rc != WG_OK) return rc;
  return wg::direct_launch(ta, ty, w_img_dev, scale_dev, shift_dev, y_dev, n_img, cin, cout, bn, relu, out_padded,
                           max_ctas, bo_mode, static_cast<cudaStream_t>(stream));
}
#endif
