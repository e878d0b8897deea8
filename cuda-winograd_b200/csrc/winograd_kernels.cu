// Fused 3x3 conv (Winograd F(2x2,3x3)) + folded BatchNorm + ReLU for sm_100a, one kernel, no intermediate in HBM.
//
// Replaces the reference's three-kernel pipeline kernel_{128,256}_winograd_BtdB -> kernel_*_OuterProduct_* ->
// kernel_*_winograd_AtIA (/root/reference/Kernel128_winograd.cu:28-213, Kernel256_winograd.cu:27-218), which is
// F(4x4,3x3) in FP32 FFMA with two global round trips. Here (see DESIGN.md "3x3 kernel"):
//
//   * an M-block is 128 consecutive Winograd tiles (tile index T = n*49 + ty*7 + tx over the whole batch);
//   * per 8-channel stage the producer warp TMA-loads the 48 input rows the block touches (4-D tensor map with the
//     x axis split by parity so that stride-2 tile reads are mostly bank-conflict free) and bulk-copies the matching
//     slice of the pre-transformed filter U (written once per layer by filter_transform_f2x2_kernel in exactly the
//     shared-memory image the MMA wants);
//   * 8 transform warps compute V = B^T d B in registers, round to TF32 (nearest) and store it in the UMMA K-major
//     no-swizzle canonical layout;
//   * one thread issues the tcgen05.mma (M=128 tiles, N=KN couts, K=8 channels) of the stage into TMEM:
//       FOLD = false: 16 MMAs, one accumulator per Winograd point (16*32 = 512 fp32 columns, KN = 32);
//       FOLD = true : 24 MMAs; the tensor core itself applies the row half of the inverse transform by accumulating
//                     Z[a][j] = sum_i A^T[a][i] M[i][j] (A^T = [[1,1,1,0],[0,1,-1,-1]], the -1 via the descriptor's
//                     negate-A bit), so only 8 accumulators live in TMEM and KN = 64: the CUDA-core input transform and
//                     its shared-memory traffic -- the measured bottleneck -- are repeated K/64 instead of K/32 times;
//   * after the last stage the same 8 warps run the epilogue: tcgen05.ld, the remaining inverse transform,
//     relu(scale*Y + shift), 128-bit stores of NHWC output (optionally into the reference's zero-bordered
//     16x16 frame, Kernel128_winograd.cu:163,243).
#include "ptx.cuh"
#include "wg_internal.h"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace wg {

constexpr int kWorkerWarps = 8;
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;
constexpr int kThreads = 32 * 10;
constexpr int kRawRows = 48;  // input rows (n*16+y) one M-block can touch, see DESIGN.md
constexpr uint32_t kRawBytes = kRawRows * 2 * 8 * 32;  // [ny][x parity][x/2][8 ch] fp32 = 24576

template <bool BF16>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  if constexpr (BF16) umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, accumulate);
  else umma_tf32_ss(d_tmem, a_desc, b_desc, idesc, accumulate);
}

template <bool FOLD>
struct WinoCfg {
  static constexpr int KN = FOLD ? 64 : 32;              // output channels per work item
  static constexpr int kUChunksPerStage = FOLD ? 2 : 1;  // bulk copies of U per 8-channel stage
  static constexpr int kPointsPerUChunk = FOLD ? 8 : 16;
  static constexpr int kRawStages = 2, kVStages = 2, kUBufs = 3;
  // V: per point [2 k-chunks][128 rows][16 B]; the +64 skews chunk 1 by half a bank window so that a quarter warp
  // writing 4 rows x 2 chunks hits 32 distinct banks.
  static constexpr uint32_t kVLbo = 128 * 16 + 64;
  static constexpr uint32_t kVPerXi = kVLbo + 128 * 16;
  static constexpr uint32_t kVBytes = 16 * kVPerXi;
  static constexpr uint32_t kULbo = KN * 16;
  static constexpr uint32_t kUPerPoint = 2 * KN * 16;
  static constexpr uint32_t kUChunkBytes = kPointsPerUChunk * kUPerPoint;  // 16 KB in both modes
  static constexpr uint32_t kOffRaw = 0;
  static constexpr uint32_t kOffV = kOffRaw + kRawStages * kRawBytes;
  static constexpr uint32_t kOffU = kOffV + kVStages * kVBytes;
  static constexpr uint32_t kOffBar = kOffU + kUBufs * kUChunkBytes;
  static constexpr uint32_t kNumBars = 2 * kRawStages + 2 * kVStages + 2 * kUBufs + 2;
  static constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr uint32_t kTotal = kOffTmemPtr + 16;
  static_assert(kOffV % 128 == 0 && kOffU % 128 == 0 && kOffBar % 8 == 0, "alignment");
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

// Round-to-nearest (ties away) of an fp32 to TF32 precision for a tensor-core operand: the MMA ignores the low 13
// mantissa bits, so adding half a TF32 ulp to the bit pattern is all that is needed (1 integer add; cvt.rna.tf32.f32
// compiles to three ALU instructions on sm_100a).
__device__ __forceinline__ float tf32_operand(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

// BF16 = true: V and U are bf16 operands (tcgen05.mma kind::f16, FP32 accumulate, K = 16 per instruction). A V stage
// then covers 16 channels = two 8-channel raw stages (each fills one 16-byte k-chunk of every row); shared-memory
// traffic per channel for V and U halves. Input/output stay fp32. Tolerance 1e-2 (north_star), measured ~3e-3.
// CS > 1: split-C mode for small batches (latency): a cluster of CS CTAs shares one work item, each CTA runs 1/CS of
// the channel loop, the partial outputs are reduced through distributed shared memory (ld.shared::cluster).
template <bool FOLD, bool BF16 = false, int CS = 1>
__global__ void __launch_bounds__(kThreads, 1)
wino3x3_bn_relu_kernel(const __grid_constant__ CUtensorMap tmap_x, const void* __restrict__ u_img,
                       const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ y,
                       int n_img, int C, int K, int relu, int out_padded, int mv, int fp16, int ablate) {
  static_assert(!BF16 || FOLD, "the bf16 variant is built on the folded accumulation");
  static_assert(CS == 1 || FOLD, "split-C is built on the folded accumulation");
  using S = WinoCfg<FOLD>;
  constexpr int KN = S::KN;
  constexpr int kRawPerV = BF16 ? 2 : 1;  // 8-channel raw stages per V stage
  constexpr uint32_t kTmemCols = 512;
  pdl_launch_dependents();  // the next launch in the stream may start its prologue (it waits before touching x / y)
  extern __shared__ __align__(1024) uint8_t smem[];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // debug only (ablate & 64): phase timestamps of thread 0 of CTA 0, printed at exit
  long long ts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define WG_TS(i) do { if ((ablate & 64) && threadIdx.x == 0 && blockIdx.x == 0) ts[i] = clock64(); } while (0)
  WG_TS(0);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = raw_full + S::kRawStages;
  uint64_t* v_full = raw_empty + S::kRawStages;
  uint64_t* v_empty = v_full + S::kVStages;
  uint64_t* u_full = v_empty + S::kVStages;
  uint64_t* u_empty = u_full + S::kUBufs;
  uint64_t* acc_full = u_empty + S::kUBufs;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::kOffTmemPtr);

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < S::kRawStages; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], kWorkerWarps);
    }
    for (int i = 0; i < S::kVStages; ++i) {
      mbar_init(&v_full[i], kWorkerWarps);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < S::kUBufs; ++i) {
      mbar_init(&u_full[i], 1);
      mbar_init(&u_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kWorkerWarps);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before();
  if constexpr (CS > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  WG_TS(1);

  const int n_kb = C / 8;
  const int n_slices = K / KN;
  const int total_tiles = n_img * 49;
  const int n_mblocks = (total_tiles + mv - 1) / mv;  // mv = tiles per M-block (<= 128), chosen by the host to balance waves
  const int n_items = n_mblocks * n_slices;
  // CS > 1 (small batches): the CS CTAs of a cluster share one item and split its channel loop; partial results are
  // summed through distributed shared memory after the loop. One item per cluster (the host sizes the grid so).
  const uint32_t crank = CS > 1 ? cluster_ctarank() : 0u;
  const int item0 = blockIdx.x / CS, item_step = gridDim.x / CS;
  const int kb_per = n_kb / CS;          // 8-channel raw stages this CTA runs
  const int kb0 = (int)crank * kb_per;   // first one

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {  // elect.sync, not lane == 0: lets ptxas keep descriptors in uniform registers (no per-MMA waterfall loop)
      uint32_t rs = 0, rph = 0, us = 0, uph = 0;
      bool u_primed = false;
      if (item0 < n_items) {
        // the filter does not depend on the previous kernel in the stream: request the first stage's U chunks before
        // waiting for that kernel (programmatic dependent launch), the activations after
        const uint8_t* u_src = reinterpret_cast<const uint8_t*>(u_img) +
                               (size_t)(item0 % n_slices) * (n_kb / kRawPerV) * S::kUChunksPerStage * S::kUChunkBytes;
#pragma unroll
        for (int h = 0; h < S::kUChunksPerStage; ++h) {
          mbar_arrive_expect_tx(&u_full[us], S::kUChunkBytes);
          tma_bulk_g2s(smem + S::kOffU + us * S::kUChunkBytes,
                       u_src + (size_t)((kb0 / kRawPerV) * S::kUChunksPerStage + h) * S::kUChunkBytes, S::kUChunkBytes,
                       &u_full[us]);
          if (++us == S::kUBufs) { us = 0; uph ^= 1; }
        }
        u_primed = true;
      }
      pdl_wait();
      for (int item = item0; item < n_items; item += item_step) {
        const int slice = item % n_slices;
        const int mb = item / n_slices;
        const int t0 = mb * mv;
        const int ny0 = (t0 / 49) * 16 + 2 * ((t0 % 49) / 7);
        const uint8_t* u_src = reinterpret_cast<const uint8_t*>(u_img) +
                               (size_t)slice * (n_kb / kRawPerV) * S::kUChunksPerStage * S::kUChunkBytes;
        for (int kb = kb0; kb < kb0 + kb_per; ++kb) {
          mbar_wait(&raw_empty[rs], rph ^ 1);
          mbar_arrive_expect_tx(&raw_full[rs], kRawBytes);
          tma_tensor_4d_g2s(smem + S::kOffRaw + rs * kRawBytes, &tmap_x, kb * 8, 0, 0, ny0, &raw_full[rs]);
          if (++rs == S::kRawStages) { rs = 0; rph ^= 1; }
          if (kb % kRawPerV != 0) continue;  // the U chunks of a V stage go out with its first raw stage
          if (u_primed) {  // already requested above
            u_primed = false;
            continue;
          }
          const int kv = kb / kRawPerV;
#pragma unroll
          for (int h = 0; h < S::kUChunksPerStage; ++h) {
            mbar_wait(&u_empty[us], uph ^ 1);
            mbar_arrive_expect_tx(&u_full[us], S::kUChunkBytes);
            tma_bulk_g2s(smem + S::kOffU + us * S::kUChunkBytes,
                         u_src + (size_t)(kv * S::kUChunksPerStage + h) * S::kUChunkBytes, S::kUChunkBytes,
                         &u_full[us]);
            if (++us == S::kUBufs) { us = 0; uph ^= 1; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (elect_one()) {  // elect.sync, not lane == 0: lets ptxas keep descriptors in uniform registers (no per-MMA waterfall loop)
      const uint32_t fmt = BF16 ? (fp16 ? kFmtF16 : kFmtBF16) : kFmtTF32;
      const uint32_t idesc = make_idesc(fmt, 128, KN);
      const uint32_t idesc_neg = make_idesc(fmt, 128, KN, 1);  // D += (-A) * B
      const uint32_t v_base = smem_u32(smem + S::kOffV);
      const uint32_t u_base = smem_u32(smem + S::kOffU);
      uint32_t vs = 0, vph = 0, us = 0, uph = 0, aph = 0;
      for (int item = item0; item < n_items; item += item_step) {
        mbar_wait(acc_empty, aph ^ 1);  // epilogue of the previous item has drained TMEM
        tc_fence_after();
        for (int kb = 0; kb < kb_per / kRawPerV; ++kb) {  // V stages
          const uint32_t acc = kb > 0 ? 1u : 0u;
          mbar_wait(&v_full[vs], vph);
          const uint32_t va = v_base + vs * S::kVBytes;
          if (ablate & 8) {
            for (int h = 0; h < S::kUChunksPerStage; ++h) {
              mbar_wait(&u_full[us], uph);
              umma_commit(&u_empty[us]);
              if (++us == S::kUBufs) { us = 0; uph ^= 1; }
            }
          } else if constexpr (!FOLD) {
            mbar_wait(&u_full[us], uph);
            tc_fence_after();
            const uint32_t ua = u_base + us * S::kUChunkBytes;
#pragma unroll
            for (int xi = 0; xi < 16; ++xi) {
              const uint64_t a_desc = make_smem_desc(va + xi * S::kVPerXi, S::kVLbo, 128, kLayoutNone);
              const uint64_t b_desc = make_smem_desc(ua + xi * S::kUPerPoint, S::kULbo, 128, kLayoutNone);
              umma_ss<BF16>(tmem_base + xi * KN, a_desc, b_desc, idesc, acc);
            }
            umma_commit(&u_empty[us]);
            if (++us == S::kUBufs) { us = 0; uph ^= 1; }
          } else {
#pragma unroll
            for (int jh = 0; jh < 2; ++jh) {
              mbar_wait(&u_full[us], uph);
              tc_fence_after();
              const uint32_t ua = u_base + us * S::kUChunkBytes;
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                const int j = jh * 2 + jj;
                const uint32_t z0 = tmem_base + (j * 2 + 0) * KN;  // Z[0][j] = M0j + M1j + M2j
                const uint32_t z1 = tmem_base + (j * 2 + 1) * KN;  // Z[1][j] = M1j - M2j - M3j
                uint64_t a_desc[4], b_desc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  a_desc[i] = make_smem_desc(va + (4 * i + j) * S::kVPerXi, S::kVLbo, 128, kLayoutNone);
                  b_desc[i] = make_smem_desc(ua + (i * 2 + jj) * S::kUPerPoint, S::kULbo, 128, kLayoutNone);
                }
                umma_ss<BF16>(z0, a_desc[1], b_desc[1], idesc, acc);  // first writer of both accumulators
                umma_ss<BF16>(z1, a_desc[1], b_desc[1], idesc, acc);
                umma_ss<BF16>(z0, a_desc[0], b_desc[0], idesc, 1u);
                umma_ss<BF16>(z0, a_desc[2], b_desc[2], idesc, 1u);
                umma_ss<BF16>(z1, a_desc[2], b_desc[2], idesc_neg, 1u);
                umma_ss<BF16>(z1, a_desc[3], b_desc[3], idesc_neg, 1u);
              }
              umma_commit(&u_empty[us]);
              if (++us == S::kUBufs) { us = 0; uph ^= 1; }
            }
          }
          umma_commit(&v_empty[vs]);
          if (++vs == S::kVStages) { vs = 0; vph ^= 1; }
        }
        umma_commit(acc_full);
        aph ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ transform + epilogue warps
    // transform task: row = 16*warp + q (tile within the M-block), c = which 4-channel half of the 8-channel stage
    const int c = (lane >> 2) & 1;
    const int q = (lane & 3) + 4 * (lane >> 3);
    const int trow = warp * 16 + q;
    // epilogue task: row = TMEM lane
    const int quad = warp & 3;
    const int half = warp >> 2;
    const int erow = quad * 32 + lane;
    const uint32_t raw_base = smem_u32(smem + S::kOffRaw);
    const uint32_t v_base = smem_u32(smem + S::kOffV);

    uint32_t rs = 0, rph = 0, vs = 0, vph = 0, aph = 0;
    for (int item = item0; item < n_items; item += item_step) {
      const int slice = item % n_slices;
      const int mb = item / n_slices;
      const int t0 = mb * mv;
      const int ny0 = (t0 / 49) * 16 + 2 * ((t0 % 49) / 7);

      // ---- transform: this thread's tile and its offset inside a raw stage
      const int T = t0 + trow;
      const int valid_rows = min(mv, total_tiles - t0);  // rows of this M-block that hold real tiles
      const bool tvalid = trow < valid_rows;
      const bool warp_active = warp * 16 < valid_rows;  // warp-uniform: warps whose 16 rows are all padding only keep the barriers moving
      uint32_t raw_off = 0;
      {
        const int n = T / 49, t = T % 49, ty = t / 7, tx = t % 7;
        if (tvalid) raw_off = (uint32_t)((n * 16 + 2 * ty - ny0) * 512 + tx * 32 + c * 16);
      }
      const uint32_t v_off = (uint32_t)(c * S::kVLbo + trow * 16);

      for (int kb = kb0; kb < kb0 + kb_per; ++kb) {
        float4 d[4][4];
        mbar_wait(&raw_full[rs], rph);
        if (kb == kb0) WG_TS(2);
        const int sub = kb % kRawPerV;               // which 16-byte k-chunk of the V rows this raw stage fills
        const bool v_first = sub == 0, v_last = sub == kRawPerV - 1;
        if (!warp_active) {
          // nothing to transform: release the raw stage and report "V ready" in step with the other warps
          if (lane == 0) mbar_arrive(&raw_empty[rs]);
          if (++rs == S::kRawStages) { rs = 0; rph ^= 1; }
          if (v_first) mbar_wait(&v_empty[vs], vph ^ 1);
          if (v_last) {
            if (lane == 0) mbar_arrive(&v_full[vs]);
            if (++vs == S::kVStages) { vs = 0; vph ^= 1; }
          }
          continue;
        }
        if (tvalid && !(ablate & 1)) {
          const uint32_t a = raw_base + rs * kRawBytes + raw_off;
#pragma unroll
          for (int dy = 0; dy < 4; ++dy)
#pragma unroll
            for (int dx = 0; dx < 4; ++dx)
              d[dy][dx] = ld_shared_v4(a + dy * 512 + (dx & 1) * 256 + (dx >> 1) * 32);
        } else {
#pragma unroll
          for (int dy = 0; dy < 4; ++dy)
#pragma unroll
            for (int dx = 0; dx < 4; ++dx) d[dy][dx] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // column pass t = B^T d, in place over dy
#pragma unroll
        for (int dx = 0; dx < 4; ++dx) {
          const float4 d0 = d[0][dx], d1 = d[1][dx], d2 = d[2][dx], d3 = d[3][dx];
          d[0][dx] = make_float4(d0.x - d2.x, d0.y - d2.y, d0.z - d2.z, d0.w - d2.w);
          d[1][dx] = make_float4(d1.x + d2.x, d1.y + d2.y, d1.z + d2.z, d1.w + d2.w);
          d[2][dx] = make_float4(d2.x - d1.x, d2.y - d1.y, d2.z - d1.z, d2.w - d1.w);
          d[3][dx] = make_float4(d1.x - d3.x, d1.y - d3.y, d1.z - d3.z, d1.w - d3.w);
        }
        // the raw stage is in registers now: hand it back to the producer before the row pass
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[rs]);
        if (++rs == S::kRawStages) { rs = 0; rph ^= 1; }

        if (v_first) mbar_wait(&v_empty[vs], vph ^ 1);  // MMAs that read this V stage have completed
        const uint32_t vdst = v_base + vs * S::kVBytes + v_off;
        // row pass V = t B, round to the operand type, store point (i,j) at xi = 4*i + j
        if constexpr (BF16) {
          // 4 channels -> 8 bytes at [k-chunk sub][row][c*8]
          const uint32_t bdst = v_base + vs * S::kVBytes + sub * S::kVLbo + trow * 16 + c * 8;
          if (ablate & 2) {
          } else if (fp16) {  // warp-uniform: fp16 operands (10-bit mantissa like TF32, range +-65504)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 a0 = d[i][0], a1 = d[i][1], a2 = d[i][2], a3 = d[i][3];
              st_shared_f16x4(bdst + (4 * i + 0) * S::kVPerXi, a0.x - a2.x, a0.y - a2.y, a0.z - a2.z, a0.w - a2.w);
              st_shared_f16x4(bdst + (4 * i + 1) * S::kVPerXi, a1.x + a2.x, a1.y + a2.y, a1.z + a2.z, a1.w + a2.w);
              st_shared_f16x4(bdst + (4 * i + 2) * S::kVPerXi, a2.x - a1.x, a2.y - a1.y, a2.z - a1.z, a2.w - a1.w);
              st_shared_f16x4(bdst + (4 * i + 3) * S::kVPerXi, a1.x - a3.x, a1.y - a3.y, a1.z - a3.z, a1.w - a3.w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 a0 = d[i][0], a1 = d[i][1], a2 = d[i][2], a3 = d[i][3];
              st_shared_bf16x4(bdst + (4 * i + 0) * S::kVPerXi, a0.x - a2.x, a0.y - a2.y, a0.z - a2.z, a0.w - a2.w);
              st_shared_bf16x4(bdst + (4 * i + 1) * S::kVPerXi, a1.x + a2.x, a1.y + a2.y, a1.z + a2.z, a1.w + a2.w);
              st_shared_bf16x4(bdst + (4 * i + 2) * S::kVPerXi, a2.x - a1.x, a2.y - a1.y, a2.z - a1.z, a2.w - a1.w);
              st_shared_bf16x4(bdst + (4 * i + 3) * S::kVPerXi, a1.x - a3.x, a1.y - a3.y, a1.z - a3.z, a1.w - a3.w);
            }
          }
        } else if (!(ablate & 2))
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a0 = d[i][0], a1 = d[i][1], a2 = d[i][2], a3 = d[i][3];
          st_shared_v4(vdst + (4 * i + 0) * S::kVPerXi, tf32_operand(a0.x - a2.x), tf32_operand(a0.y - a2.y),
                       tf32_operand(a0.z - a2.z), tf32_operand(a0.w - a2.w));
          st_shared_v4(vdst + (4 * i + 1) * S::kVPerXi, tf32_operand(a1.x + a2.x), tf32_operand(a1.y + a2.y),
                       tf32_operand(a1.z + a2.z), tf32_operand(a1.w + a2.w));
          st_shared_v4(vdst + (4 * i + 2) * S::kVPerXi, tf32_operand(a2.x - a1.x), tf32_operand(a2.y - a1.y),
                       tf32_operand(a2.z - a1.z), tf32_operand(a2.w - a1.w));
          st_shared_v4(vdst + (4 * i + 3) * S::kVPerXi, tf32_operand(a1.x - a3.x), tf32_operand(a1.y - a3.y),
                       tf32_operand(a1.z - a3.z), tf32_operand(a1.w - a3.w));
        }
        if (v_last) {
          if (!(ablate & 4)) fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&v_full[vs]);
          if (++vs == S::kVStages) { vs = 0; vph ^= 1; }
        }
      }

      // ---- epilogue: (rest of) Y = A^T M A, BN, ReLU, store
      const int TE = t0 + erow;
      const bool evalid = erow < valid_rows;
      const int n = TE / 49, t = TE % 49, ty = t / 7, tx = t % 7;
      const int W = out_padded ? 16 : 14;
      const int o = out_padded ? 1 : 0;
      const int pix0 = evalid ? ((n * W + 2 * ty + o) * W + 2 * tx + o) : -1;  // first output pixel of this tile
      const size_t rstride = (size_t)W * K;
      const uint32_t stg = v_base + warp * (32 * 512);  // this warp's 16 KB staging area inside the V buffers

      mbar_wait(acc_full, aph);
      WG_TS(3);
      aph ^= 1;
      tc_fence_after();
      constexpr int kColsPerWarp = KN / 2;
      // warps whose 32 TMEM lanes hold only padding rows have nothing to drain (warp-uniform)
      const int cc_end = quad * 32 < valid_rows ? kColsPerWarp : 0;
#pragma unroll 1
      for (int cc = 0; cc < cc_end; cc += 8) {
        const int c0 = half * kColsPerWarp + cc;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + c0;
        float o00[8], o01[8], o10[8], o11[8];  // Y[a][b] before BN
        if constexpr (!FOLD) {
          float m[16][8];
#pragma unroll
          for (int xi = 0; xi < 16; ++xi) tmem_ld_x8(taddr + xi * KN, m[xi]);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float s0[4], s1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              s0[j] = m[0 + j][e] + m[4 + j][e] + m[8 + j][e];
              s1[j] = m[4 + j][e] - m[8 + j][e] - m[12 + j][e];
            }
            o00[e] = s0[0] + s0[1] + s0[2];
            o01[e] = s0[1] - s0[2] - s0[3];
            o10[e] = s1[0] + s1[1] + s1[2];
            o11[e] = s1[1] - s1[2] - s1[3];
          }
        } else {
          float z[8][8];  // z[j*2 + a][e]
#pragma unroll
          for (int p = 0; p < 8; ++p) tmem_ld_x8(taddr + p * KN, z[p]);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            o00[e] = z[0][e] + z[2][e] + z[4][e];
            o01[e] = z[2][e] - z[4][e] - z[6][e];
            o10[e] = z[1][e] + z[3][e] + z[5][e];
            o11[e] = z[3][e] - z[5][e] - z[7][e];
          }
        }
        if constexpr (CS > 1) {
          // split-C: these are partial sums over this CTA's channels. Park them (pre-BN) in local shared memory as
          // part[row][pixel][KN couts] (1 KB rows, 16-byte chunks XOR-swizzled by the row) for the cluster reduction.
          if (erow < 64) {
            const uint32_t pdst = v_base + erow * 1024;
            const uint32_t sw = erow & 15;
            const uint32_t k0 = (uint32_t)c0 >> 2;
            st_shared_v4(pdst + 0 * 256 + (((k0 + 0) ^ sw) << 4), o00[0], o00[1], o00[2], o00[3]);
            st_shared_v4(pdst + 0 * 256 + (((k0 + 1) ^ sw) << 4), o00[4], o00[5], o00[6], o00[7]);
            st_shared_v4(pdst + 1 * 256 + (((k0 + 0) ^ sw) << 4), o01[0], o01[1], o01[2], o01[3]);
            st_shared_v4(pdst + 1 * 256 + (((k0 + 1) ^ sw) << 4), o01[4], o01[5], o01[6], o01[7]);
            st_shared_v4(pdst + 2 * 256 + (((k0 + 0) ^ sw) << 4), o10[0], o10[1], o10[2], o10[3]);
            st_shared_v4(pdst + 2 * 256 + (((k0 + 1) ^ sw) << 4), o10[4], o10[5], o10[6], o10[7]);
            st_shared_v4(pdst + 3 * 256 + (((k0 + 0) ^ sw) << 4), o11[0], o11[1], o11[2], o11[3]);
            st_shared_v4(pdst + 3 * 256 + (((k0 + 1) ^ sw) << 4), o11[4], o11[5], o11[6], o11[7]);
          }
          continue;
        }
        const float4 sc0 = __ldg(reinterpret_cast<const float4*>(scale + slice * KN + c0));
        const float4 sc1 = __ldg(reinterpret_cast<const float4*>(scale + slice * KN + c0 + 4));
        const float4 sh0 = __ldg(reinterpret_cast<const float4*>(shift + slice * KN + c0));
        const float4 sh1 = __ldg(reinterpret_cast<const float4*>(shift + slice * KN + c0 + 4));
        const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
        const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o00[e] = fmaf(sc[e], o00[e], sh[e]);
          o01[e] = fmaf(sc[e], o01[e], sh[e]);
          o10[e] = fmaf(sc[e], o10[e], sh[e]);
          o11[e] = fmaf(sc[e], o11[e], sh[e]);
          if (relu) {
            o00[e] = fmaxf(o00[e], 0.f);
            o01[e] = fmaxf(o01[e], 0.f);
            o10[e] = fmaxf(o10[e], 0.f);
            o11[e] = fmaxf(o11[e], 0.f);
          }
        }
        // stage this 8-cout group in the (now idle) V buffers: [tile][pixel][128 B], 16-byte chunks XOR-swizzled by
        // the tile index so that both the per-tile writes here and the per-pixel reads below are conflict free
        {
          const uint32_t sdst = stg + lane * 512;
          const uint32_t sw = lane & 7;
          const uint32_t k0 = (uint32_t)cc >> 2;
          st_shared_v4(sdst + 0 * 128 + (((k0 + 0) ^ sw) << 4), o00[0], o00[1], o00[2], o00[3]);
          st_shared_v4(sdst + 0 * 128 + (((k0 + 1) ^ sw) << 4), o00[4], o00[5], o00[6], o00[7]);
          st_shared_v4(sdst + 1 * 128 + (((k0 + 0) ^ sw) << 4), o01[0], o01[1], o01[2], o01[3]);
          st_shared_v4(sdst + 1 * 128 + (((k0 + 1) ^ sw) << 4), o01[4], o01[5], o01[6], o01[7]);
          st_shared_v4(sdst + 2 * 128 + (((k0 + 0) ^ sw) << 4), o10[0], o10[1], o10[2], o10[3]);
          st_shared_v4(sdst + 2 * 128 + (((k0 + 1) ^ sw) << 4), o10[4], o10[5], o10[6], o10[7]);
          st_shared_v4(sdst + 3 * 128 + (((k0 + 0) ^ sw) << 4), o11[0], o11[1], o11[2], o11[3]);
          st_shared_v4(sdst + 3 * 128 + (((k0 + 1) ^ sw) << 4), o11[4], o11[5], o11[6], o11[7]);
        }
      }
      __syncwarp();
      // write-out: kChunks lanes cover one output pixel's kColsPerWarp couts = one contiguous 64/128-byte run
      if (CS == 1 && cc_end != 0 && !(ablate & 32)) {
        constexpr int kChunks = kColsPerWarp / 4;      // 16-byte chunks per pixel owned by this warp (4 or 8)
        constexpr int kTilesPerInstr = 32 / kChunks;   // 8 or 4
        const int j = lane % kChunks;
        const int tsub = lane / kChunks;
        // with 4 chunks a quarter warp holds two tiles: pick tiles 4 apart so their swizzled chunks do not collide
        const int tperm = kChunks == 8 ? tsub : ((tsub >> 1) | ((tsub & 1) << 2));
        float* gcol = y + slice * KN + half * kColsPerWarp + j * 4;
#pragma unroll
        for (int it = 0; it < 32 / kTilesPerInstr; ++it) {
          const int tl = it * kTilesPerInstr + tperm;
          const int pix = __shfl_sync(0xffffffffu, pix0, tl);
          const uint32_t src = stg + tl * 512 + ((j ^ (tl & 7)) << 4);
          if (pix >= 0) {
            float* g = gcol + (size_t)pix * K;
            const float4 v0 = ld_shared_v4(src), v1 = ld_shared_v4(src + 128);
            const float4 v2 = ld_shared_v4(src + 256), v3 = ld_shared_v4(src + 384);
            *reinterpret_cast<float4*>(g) = v0;
            *reinterpret_cast<float4*>(g + K) = v1;
            *reinterpret_cast<float4*>(g + rstride) = v2;
            *reinterpret_cast<float4*>(g + rstride + K) = v3;
          }
        }
        if (out_padded && evalid && (ty == 0 || ty == 6 || tx == 0 || tx == 6)) {
          // zero border of the reference's 16x16 frame: edge tiles also own their share of the border
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          float* p = y + (size_t)pix0 * K + slice * KN + half * kColsPerWarp;
          const ptrdiff_t dyb = ty == 0 ? -(ptrdiff_t)rstride : (ty == 6 ? 2 * (ptrdiff_t)rstride : 0);
          const ptrdiff_t dxb = tx == 0 ? -(ptrdiff_t)K : (tx == 6 ? 2 * (ptrdiff_t)K : 0);
#pragma unroll 1
          for (int e = 0; e < kColsPerWarp; e += 4) {
            if (dyb != 0) {
              *reinterpret_cast<float4*>(p + dyb + e) = z4;
              *reinterpret_cast<float4*>(p + dyb + K + e) = z4;
            }
            if (dxb != 0) {
              *reinterpret_cast<float4*>(p + dxb + e) = z4;
              *reinterpret_cast<float4*>(p + dxb + rstride + e) = z4;
            }
            if (dyb != 0 && dxb != 0) *reinterpret_cast<float4*>(p + dyb + dxb + e) = z4;
          }
        }
      }
      // every worker warp must be done with its staging area before any of them refills the V buffers
      asm volatile("bar.sync 1, 256;" ::: "memory");
      WG_TS(4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }

  if constexpr (CS > 1) {
    // ---- split-C reduction over the cluster (deterministic: fixed summation order, no atomics)
    __syncwarp();        // the single-lane roles rejoin their warps before the aligned cluster barrier
    cluster_sync_all();  // every CTA's partial Y is parked in its shared memory
    WG_TS(5);
    // Push model (a pull over ld.shared::cluster measured 10.5 k clk for 50 KB per CTA): every CTA forwards each
    // 16-byte piece of its partial Y to the CTA that owns those couts with a fire-and-forget st.shared::cluster into
    // a receive area [source CTA][row][pixel][couts per owner] behind the parking area; one more cluster barrier, then
    // the owner sums its CS inboxes locally in a fixed order.
    constexpr int kChunksPerCta = KN / CS / 4;           // 16-byte cout chunks a CTA owns (16 couts = 4, 8 couts = 2)
    constexpr uint32_t kInboxBytes = 64 * 4 * (KN / CS) * 4;  // one source's [64 rows][4 px][couts per owner] fp32
    const uint32_t part_local = smem_u32(smem + S::kOffV);
    const uint32_t recv_local = part_local + 64 * 1024;  // CS inboxes = 64 KB, the parking area is the 64 KB before
    const bool reducer = warp < kWorkerWarps && item0 < n_items;
    const int slice = item0 % n_slices;
    const int t0 = (item0 / n_slices) * mv;
    const int valid_rows = reducer ? min(min(mv, total_tiles - t0), 64) : 0;
    if (reducer) {
      uint32_t inbox_remote[CS];  // my inbox inside each owner CTA
#pragma unroll
      for (int p = 0; p < CS; ++p)
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                     : "=r"(inbox_remote[p])
                     : "r"(recv_local + crank * kInboxBytes), "r"(p));
      const int units = valid_rows * 4 * (KN / 4);  // (row, pixel, 16-byte chunk) over all KN couts
      for (int u = threadIdx.x; u < units; u += kWorkerWarps * 32) {
        const int chunk = u % (KN / 4);
        const int rp = u / (KN / 4);  // row * 4 + pixel
        const int row = rp >> 2;
        const float4 v = ld_shared_v4(part_local + rp * 256 + ((chunk ^ (row & 15)) << 4));
        const int owner = chunk / kChunksPerCta;
        uint32_t dst = inbox_remote[0];
#pragma unroll
        for (int p = 1; p < CS; ++p) dst = owner == p ? inbox_remote[p] : dst;
        dst += rp * (kChunksPerCta * 16) + (chunk % kChunksPerCta) * 16;
        asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v.x), "f"(v.y), "f"(v.z),
                     "f"(v.w)
                     : "memory");
      }
    }
    __syncwarp();
    cluster_sync_all();  // all inboxes are filled (release/acquire at cluster scope)
    WG_TS(6);
    if (reducer) {
      const int W = out_padded ? 16 : 14;
      const int o = out_padded ? 1 : 0;
      const int units = valid_rows * 4 * kChunksPerCta;  // (row, pixel, owned chunk)
      for (int u = threadIdx.x; u < units; u += kWorkerWarps * 32) {
        const int lc = u % kChunksPerCta;
        const int rp = u / kChunksPerCta;
        const int row = rp >> 2, px = rp & 3;
        const int chunk = (int)crank * kChunksPerCta + lc;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < CS; ++p) {
          const float4 v = ld_shared_v4(recv_local + p * kInboxBytes + rp * (kChunksPerCta * 16) + lc * 16);
          acc.x += v.x;
          acc.y += v.y;
          acc.z += v.z;
          acc.w += v.w;
        }
        const int cout0 = slice * KN + chunk * 4;
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + cout0));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + cout0));
        acc.x = fmaf(sc.x, acc.x, sh.x);
        acc.y = fmaf(sc.y, acc.y, sh.y);
        acc.z = fmaf(sc.z, acc.z, sh.z);
        acc.w = fmaf(sc.w, acc.w, sh.w);
        if (relu) {
          acc.x = fmaxf(acc.x, 0.f);
          acc.y = fmaxf(acc.y, 0.f);
          acc.z = fmaxf(acc.z, 0.f);
          acc.w = fmaxf(acc.w, 0.f);
        }
        const int T = t0 + row;
        const int n = T / 49, t = T % 49, ty = t / 7, tx = t % 7;
        const int pix = (n * W + 2 * ty + o + (px >> 1)) * W + 2 * tx + o + (px & 1);
        float* g = y + (size_t)pix * K + cout0;
        *reinterpret_cast<float4*>(g) = acc;
        if (out_padded) {
          // zero border of the reference's 16x16 frame, written by whichever corner pixel of an edge tile is nearest
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const int a = px >> 1, b = px & 1;
          const ptrdiff_t dyb = (ty == 0 && a == 0) ? -(ptrdiff_t)W * K : ((ty == 6 && a == 1) ? (ptrdiff_t)W * K : 0);
          const ptrdiff_t dxb = (tx == 0 && b == 0) ? -(ptrdiff_t)K : ((tx == 6 && b == 1) ? (ptrdiff_t)K : 0);
          if (dyb != 0) *reinterpret_cast<float4*>(g + dyb) = z4;
          if (dxb != 0) *reinterpret_cast<float4*>(g + dxb) = z4;
          if (dyb != 0 && dxb != 0) *reinterpret_cast<float4*>(g + dyb + dxb) = z4;
        }
      }
    }
    WG_TS(7);  // no remote access after the second barrier: CTAs may retire independently
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<kTmemCols>(tmem_base);
  if ((ablate & 64) && threadIdx.x == 0 && blockIdx.x == 0)
    printf("wg ts (clk from entry): prologue %lld raw0 %lld acc_full %lld epi %lld csync1 %lld reduce %lld csync2 %lld exit %lld\n",
           ts[1] - ts[0], ts[2] - ts[0], ts[3] - ts[0], ts[4] - ts[0], ts[5] - ts[0], ts[6] - ts[0], ts[7] - ts[0],
           clock64() - ts[0]);
#undef WG_TS
}

// ---------------------------------------------------------------------------------------------------------------
// Once-per-layer filter transform U = G g G^T (F(2x2,3x3)), RN-rounded to TF32, written as the exact shared-memory
// image of the pipeline's bulk copies:
//   plain: [K/32 slice][C/8 k-block][16 points (i,j)]        [2 k-chunks][32 couts][4 channels]
//   fold : [K/64 slice][C/8 k-block][2 j-halves][4 i][2 jj]  [2 k-chunks][64 couts][4 channels]   (j = 2*jh + jj)
// Replaces the offline weight_generator loop (/root/reference/data_generator.py:63-78; that one is F(4x4), 36 points).
//   bf16 : [K/64 slice][C/16 k-block][2 j-halves][4 i][2 jj] [2 k-chunks][64 couts][8 channels] as __nv_bfloat16 (RN)
__global__ void filter_transform_f2x2_kernel(const float* __restrict__ w_kcrs, void* __restrict__ u_img_v, int C, int K,
                                             int KN, int fold, int bf16) {
  float* u_img = static_cast<float*>(u_img_v);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * K) return;
  const int ch = idx % C;
  const int k = idx / C;
  const float* g = w_kcrs + ((size_t)k * C + ch) * 9;
  float gg[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s) gg[r][s] = g[r * 3 + s];
  // t = G g  (4x3)
  float tt[4][3];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    tt[0][s] = gg[0][s];
    tt[1][s] = 0.5f * (gg[0][s] + gg[1][s] + gg[2][s]);
    tt[2][s] = 0.5f * (gg[0][s] - gg[1][s] + gg[2][s]);
    tt[3][s] = gg[2][s];
  }
  const int slice = k / KN, kn = k % KN;
  const int kb = ch / 8, chunk = (ch % 8) / 4, e = ch % 4;
  const size_t point_floats = (size_t)2 * KN * 4;
  const size_t stage = ((size_t)slice * (C / 8) + kb) * 16 * point_floats;
  // bf16 image: 16-channel k-blocks, 8 channels per 16-byte chunk, same bytes per point (2 * KN * 16)
  __nv_bfloat16* u_bf = static_cast<__nv_bfloat16*>(u_img_v);
  const size_t point_halfs = (size_t)2 * KN * 8;
  const size_t stage_bf = ((size_t)slice * (C / 16) + ch / 16) * 16 * point_halfs;
  const int chunk_bf = (ch % 16) / 8, e_bf = ch % 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float u[4];
    u[0] = tt[i][0];
    u[1] = 0.5f * (tt[i][0] + tt[i][1] + tt[i][2]);
    u[2] = 0.5f * (tt[i][0] - tt[i][1] + tt[i][2]);
    u[3] = tt[i][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = fold ? ((j >> 1) * 8 + i * 2 + (j & 1)) : (4 * i + j);
      if (bf16 == 2)
        static_cast<__half*>(u_img_v)[stage_bf + p * point_halfs + ((size_t)chunk_bf * KN + kn) * 8 + e_bf] =
            __float2half_rn(u[j]);
      else if (bf16)
        u_bf[stage_bf + p * point_halfs + ((size_t)chunk_bf * KN + kn) * 8 + e_bf] = __float2bfloat16_rn(u[j]);
      else
        u_img[stage + p * point_floats + ((size_t)chunk * KN + kn) * 4 + e] = to_tf32_rn(u[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side

int wino_make_tmap(CUtensorMap* tmap, const float* x, int n_img, int C) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return WG_ERR_DRIVER;
  // view x[N][16][16][C] as (c, x/2, x&1, n*16+y): the parity split makes stride-2 tile reads conflict free
  cuuint64_t dims[4] = {(cuuint64_t)C, 8, 2, (cuuint64_t)n_img * 16};
  cuuint64_t strides[3] = {(cuuint64_t)2 * C * 4, (cuuint64_t)C * 4, (cuuint64_t)16 * C * 4};
  cuuint32_t box[4] = {8, 8, 2, (cuuint32_t)kRawRows};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, wg::l2_promotion(),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WG_OK : WG_ERR_TMAP;
}

// Small batches: one item per cluster of CS CTAs (64-tile M-blocks), channel loop split CS ways.
template <bool BF16, int CS>
static int launch_wino_split(const CUtensorMap& tmap, const void* u_img, const float* scale, const float* shift,
                             float* y, int n_img, int C, int K, int relu, int out_padded, int fp16, cudaStream_t stream) {
  using S = WinoCfg<true>;
  static unsigned long long configured = 0;  // per device: the attribute is a property of the function on ONE device
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long dev_bit_ = 1ull << (dev_ & 63);
  if (!(configured & dev_bit_)) {
    cudaError_t e = cudaFuncSetAttribute(wino3x3_bn_relu_kernel<true, BF16, CS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kTotal);
    if (e != cudaSuccess) return WG_ERR_CUDA;
    configured |= dev_bit_;
  }
  const int mv = 64;
  const int n_items = ((n_img * 49 + mv - 1) / mv) * (K / S::KN);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_items * CS));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  static int ablate = -1;  // debug only, see launch_wino
  if (ablate < 0) {
    const char* env = dev_env("WG_DEBUG_ABLATE");
    ablate = env ? atoi(env) : 0;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, wino3x3_bn_relu_kernel<true, BF16, CS>, tmap, u_img, scale, shift, y, n_img,
                                     C, K, relu, out_padded, mv, fp16, ablate);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

template <bool FOLD, bool BF16>
static int launch_wino(const CUtensorMap& tmap, const void* u_img, const float* scale, const float* shift, float* y,
                       int n_img, int C, int K, int relu, int out_padded, int fp16, int max_ctas, cudaStream_t stream) {
  using S = WinoCfg<FOLD>;
  if constexpr (FOLD) {
    // latency mode when the whole batch is a handful of items: split the channel loop over a cluster
    static int cs_env = -1;  // WG_WINO_CS=1 disables, 4|8 forces (when legal); default auto
    if (cs_env < 0) {
      const char* e = dev_env("WG_WINO_CS");
      cs_env = e ? atoi(e) : 0;
    }
    const int n_kv = C / (BF16 ? 16 : 8);
    const int items64 = ((n_img * 49 + 63) / 64) * (K / S::KN);
    int cs = 1;
    if (cs_env != 1) {
      if (n_kv % 8 == 0 && (cs_env == 8 || (cs_env == 0 && items64 * 8 <= max_ctas))) cs = 8;
      else if (n_kv % 4 == 0 && (cs_env == 4 || (cs_env == 0 && items64 * 4 <= max_ctas))) cs = 4;
    }
    if (cs == 8)
      return launch_wino_split<BF16, 8>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded, fp16, stream);
    if (cs == 4)
      return launch_wino_split<BF16, 4>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded, fp16, stream);
  }
  static unsigned long long configured = 0;  // per device: the attribute is a property of the function on ONE device
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long dev_bit_ = 1ull << (dev_ & 63);
  if (!(configured & dev_bit_)) {
    cudaError_t e = cudaFuncSetAttribute(wino3x3_bn_relu_kernel<FOLD, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)S::kTotal);
    if (e != cudaSuccess) return WG_ERR_CUDA;
    configured |= dev_bit_;
  }
  // Tiles per M-block: the MMA is always M=128, but only `mv` rows carry tiles. A smaller mv costs MMA efficiency
  // (the tensor pipe has slack) and buys an item count that fills the last wave: per item, ~55 % of the cost (MMA
  // operand reads, TMA) is fixed and ~45 % (the input transform) scales with the active 16-row warps.
  const int total_tiles = n_img * 49;
  const int n_slices = K / S::KN;
  int mv = 128;
  static int mv_env = -1;  // WG_WINO_MV=<16..128> pins it (experiments)
  if (mv_env < 0) {
    const char* e = dev_env("WG_WINO_MV");
    mv_env = e ? atoi(e) : 0;
  }
  if (mv_env >= 16 && mv_env <= 128) {
    mv = mv_env;
  } else {
    double best = 1e30;
    for (int cand = 128; cand >= 64; cand -= 16) {
      const long long items = (long long)((total_tiles + cand - 1) / cand) * n_slices;
      const long long waves = (items + max_ctas - 1) / max_ctas;
      const double cost = (double)waves * (0.55 + 0.45 * cand / 128.0);
      if (cost < best - 1e-9) {
        best = cost;
        mv = cand;
      }
    }
  }
  const int n_items = ((total_tiles + mv - 1) / mv) * n_slices;
  int grid = n_items < max_ctas ? n_items : max_ctas;
  if (grid < 1) grid = 1;
  static int ablate = -1;  // debug only: WG_DEBUG_ABLATE=<bitmask> switches pipeline pieces off for timing experiments
  if (ablate < 0) {
    const char* e = dev_env("WG_DEBUG_ABLATE");
    ablate = e ? atoi(e) : 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, wino3x3_bn_relu_kernel<FOLD, BF16>, tmap, u_img, scale, shift, y, n_img, C,
                                     K, relu, out_padded, mv, fp16, ablate);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

int wino_launch(const CUtensorMap& tmap, const void* u_img, const float* scale, const float* shift, float* y,
                int n_img, int C, int K, int KN, int bf16, int relu, int out_padded, int max_ctas,
                cudaStream_t stream) {
  // bf16: 0 = TF32 operands, 1 = bf16 operands, 2 = fp16 operands (same 16-bit kernel, other format code)
  if (bf16) {
    if (KN != 64 || C % 16 != 0) return WG_ERR_ARG;
    return launch_wino<true, true>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded, bf16 == 2, max_ctas,
                                   stream);
  }
  if constexpr (kDev) {  // the TF32 instantiations of this first-generation kernel: developer build only
    if (KN == 64)
      return launch_wino<true, false>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded, 0, max_ctas, stream);
    if (KN == 32)
      return launch_wino<false, false>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded, 0, max_ctas, stream);
  }
  return WG_ERR_ARG;
}

int filter_transform_launch(const float* w_kcrs, void* u_img, int C, int K, int KN, int bf16, cudaStream_t stream) {
  const int n = C * K;
  filter_transform_f2x2_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w_kcrs, u_img, C, K, KN, KN == 64 ? 1 : 0, bf16);
  return cudaGetLastError() == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

}  // namespace wg
