// Fused 3x3 conv (Winograd F(2x2,3x3)) + folded BatchNorm + ReLU for sm_100a -- throughput kernel, TF32 operands,
// transformed input V kept in TENSOR MEMORY as the MMA's A operand.
//
// Replaces kernel_{128,256}_winograd_BtdB -> kernel_*_OuterProduct_* -> kernel_*_winograd_AtIA
// (/root/reference/Kernel128_winograd.cu:28-213, Kernel256_winograd.cu:27-218). Same decomposition as
// wino3x3_bn_relu_kernel (winograd_kernels.cu: 128-tile M-blocks x cout slices, 8-channel stages, folded accumulation
// Z[a][j] = sum_i A^T[a][i] M[i][j] with the negate-A bit), but that kernel was bound by the shared-memory data path:
// per stage it moved 64 KB of V stores + 96 KB of A-operand reads through it, on top of the raw-tile loads. Here
//
//   * the transform warps write V = B^T d B straight from registers into TMEM (tcgen05.st, thread = tile row,
//     4 channels = 4 columns) and the MMA reads A from there (tcgen05.mma ... [d], [a_tmem], b_desc): no V in shared
//     memory at all, and the MMA runs at the tensor-core floor (measured 27.7 clk for M=128 N=48 K=8 vs 53 clk with
//     A and B in shared memory, tools/selftest ts|mma);
//   * TMEM budget: 8 folded accumulators x 48 couts = 384 columns + two 64-column V halves (8 points x 8 channels
//     each; half jh holds the points with j in {2jh, 2jh+1}) = 512. So cout slices are 48 wide (then 32 for the
//     remainder: 256 = 4x48 + 2x32, 128 = 2x48 + 32), and V is single-buffered per half: the stores of stage s+1,
//     half jh wait for the 12 MMAs of stage s, half jh (tcgen05.commit -> v_empty[jh]) while its loads and FADDs overlap
//     them;
//   * raw tiles land through a SWIZZLE_32B tensor map, which spreads the 16-byte channel halves of the 7 tiles of an
//     image row over 7 different bank groups. A quarter warp is 8 consecutive tiles, though, so it always wraps into
//     the next image row and pays a second wavefront (ncu: 58 % of the load wavefronts are conflicts). Padding the
//     rows to 8 slots removes them but costs 1/8 of the MMA rows and a fifth wave at N=256 -- measured slower;
//   * the epilogue stages relu(scale*Y+shift) in its own shared-memory area ([tile][pixel][couts], rows padded by 16 B)
//     and writes full runs per output pixel.
#include "ptx.cuh"
#include "wg_internal.h"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace wg {

namespace tm {
// WW worker warps (8 or 16) + the TMA producer warp + the MMA warp. A worker thread owns one MMA row (= TMEM lane; warp
// w may only touch lanes 32*(w%4)..+31) and 32/WW of the 8 channels of a stage: 4 channels with WW=8, 2 with WW=16.
constexpr int kRawRows = 48;  // input rows (n*16+y) one 128-tile M-block can touch
constexpr uint32_t kRawBytes = kRawRows * 2 * 8 * 32;  // [ny][x parity][x/2][8 ch] fp32 = 24576
constexpr int kRawStages = 3, kUBufs = 4;
constexpr int kKNmax = 48;
constexpr uint32_t kUChunkMax = 8 * 2 * kKNmax * 16;   // 8 points x [2 k-chunks][KN couts][4 ch] = 12288
constexpr uint32_t kStgBytes = 128 * (16 * kKNmax + 16);  // [128 tiles][4 px][KN couts] fp32, tile rows padded by 16 B
// TMEM map. DB = false: 8 accumulators x 48 columns, then ONE V stage (two 64-column halves).
//           DB = true : 8 accumulators x 32 columns (all cout slices 32 wide), then TWO V stages (double-buffered).
// V stage vb, half jh at kVCol0 + 128*vb + 64*jh, point p8 at +8*p8.
template <bool DB> struct Tmem {
  static constexpr uint32_t kAccStride = DB ? 32 : 48;
  static constexpr uint32_t kVCol0 = 8 * kAccStride;
  static constexpr int kVBufs = DB ? 2 : 1;
};
constexpr uint32_t kOffRaw = 0;
constexpr uint32_t kOffU = kOffRaw + kRawStages * kRawBytes;
constexpr uint32_t kOffStg = kOffU + kUBufs * kUChunkMax;
constexpr uint32_t kOffPix = kOffStg + kStgBytes;      // first output pixel of each tile row (int[128])
constexpr uint32_t kOffBar = kOffPix + 128 * 4;
constexpr uint32_t kNumBars = 3 * kRawStages + 2 * kUBufs + 8 + 2;
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kTotal = kOffTmemPtr + 16;
static_assert(kOffU % 1024 == 0 && kOffStg % 128 == 0 && kOffBar % 8 == 0, "alignment");
static_assert(kTotal <= 227 * 1024, "shared memory budget");
}  // namespace tm

__device__ __forceinline__ float tm_tf32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

template <bool H16>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  if constexpr (H16) umma_f16_ts(d_tmem, a_tmem, b_desc, idesc, accumulate);
  else umma_tf32_ts(d_tmem, a_tmem, b_desc, idesc, accumulate);
}

// two fp32 -> one 32-bit TMEM column of 16-bit operands (first value in the low half), round to nearest
__device__ __forceinline__ float pack16(float lo, float hi, int fp16) {
  if (fp16) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
  }
  const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return __uint_as_float(*reinterpret_cast<const uint32_t*>(&b));
}

// DB: V double-buffered in TMEM at the price of 32-wide cout slices (see Tmem). With one V stage the transform of stage
// s+1 cannot store before the MMAs of stage s have completed, and those cannot start before the slowest of the 8 warps
// has stored stage s: every stage ends in an implicit barrier. With two stages the warps run up to one stage ahead.
// CLS > 1: thread-block clusters of CLS CTAs work on the SAME M-block, each on its own cout slice, and share the raw
// tiles: CTA r loads rows [r*48/CLS, (r+1)*48/CLS) of every stage's box and multicasts them to all CLS CTAs. The kernel
// is bound by what one SM's TMA unit can deliver (tiled boxes with 32-byte runs: ~31 B/clk, profiles/
// tma_tensor_probe_r01.txt; with everything but TMA and barriers switched off it still needs 62 % of its time), and
// the raw tile is the same for all cout slices of an M-block. EXPERIMENT, default off: measured 1.4x (CLS=2) to 3x
// (CLS=3) slower than CLS=1, see wino_tm_cls().
// H16: 16-bit operands (bf16, or fp16 with `fp16` set): V is stored in TMEM as packed pairs (column c = channels 2c
// and 2c+1), tcgen05.mma kind::f16 with K = 16, so a V stage covers 16 channels = TWO 8-channel raw stages and the
// per-stage hand-offs (and the MMA count) per channel halve. I/O stays fp32.
template <bool DB, int WW, int CLS, bool H16 = false>
__global__ void __launch_bounds__(32 * (WW + 2), 1)
wino3x3_tm_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ u_img,
                  const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ y, int n_img,
                  int C, int K, int n48, int relu, int out_padded, int mv, int fp16) {
  using namespace tm;
  constexpr uint32_t kAccStride = Tmem<DB>::kAccStride, kVCol0 = Tmem<DB>::kVCol0;
  constexpr int kWorkerWarps = WW, kProducerWarp = WW, kMmaWarp = WW + 1;
  constexpr int NC = 32 / WW;       // channels per worker thread (4 or 2)
  constexpr int EC = WW == 8 ? 8 : 4;  // couts per epilogue step
  constexpr int kSub = H16 ? 2 : 1;    // 8-channel raw stages per V stage
  static_assert(!H16 || (WW == 8 && !DB && CLS == 1), "16-bit operands: 8 worker warps, one V stage, no clusters");
  const bool mc = (out_padded & 2) != 0;  // y is an NVLS multicast address: stores go out as multimem.st
  out_padded &= 1;
  pdl_launch_dependents();  // the next launch in the stream may start its prologue (it waits before touching x / y)
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = raw_full + kRawStages;
  uint64_t* u_full = raw_empty + kRawStages;
  uint64_t* u_empty = u_full + kUBufs;
  uint64_t* v_full = u_empty + kUBufs;  // [V stage][half]
  uint64_t* v_empty = v_full + 4;       // [V stage][half]
  uint64_t* acc_full = v_empty + 4;
  uint64_t* acc_empty = acc_full + 1;
  uint64_t* peer_empty = acc_empty + 1;  // [kRawStages] CLS > 1: every CTA of the cluster has released this raw stage
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  int* pixtab = reinterpret_cast<int*>(smem + kOffPix);

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < kRawStages; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], kWorkerWarps);
    }
    for (int i = 0; i < kUBufs; ++i) {
      mbar_init(&u_full[i], 1);
      mbar_init(&u_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&v_full[i], kWorkerWarps);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kWorkerWarps);
    for (int i = 0; i < kRawStages; ++i) mbar_init(&peer_empty[i], CLS);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  if constexpr (CLS > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t crank = CLS > 1 ? cluster_ctarank() : 0u;
  const uint32_t tmem_base = *tmem_ptr;

  const int n_kb = C / (8 * kSub);  // V stages (8 or 16 channels each)
  const int n_slices = n48 + (K - 48 * n48) / 32;
  const int total_tiles = n_img * 49;
  const int n_mblocks = (total_tiles + mv - 1) / mv;  // mv = tiles per M-block (<= 128), chosen by the host to balance waves
  // an item = (M-block, group of CLS consecutive cout slices); the CTAs of a cluster take one slice of the group each
  const int n_groups = n_slices / CLS;
  const int n_items = n_mblocks * n_groups;
  const int item0 = blockIdx.x / CLS, item_step = gridDim.x / CLS;

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t rs = 0, rph = 0, us = 0, uph = 0;
      bool u_primed = false;
      if (item0 < n_items) {
        // the filter does not depend on the previous kernel in the stream: request the first stage's U chunks before
        // waiting for that kernel (programmatic dependent launch), the activations after
        const int s = (item0 % n_groups) * CLS + (int)crank;
        const int kn = s < n48 ? 48 : 32;
        const int c0 = s < n48 ? 48 * s : 48 * n48 + 32 * (s - n48);
        const uint8_t* u_src = reinterpret_cast<const uint8_t*>(u_img) + (size_t)n_kb * 512 * c0;
        for (int h = 0; h < 2; ++h) {
          mbar_arrive_expect_tx(&u_full[us], 256u * kn);
          tma_bulk_g2s(smem + kOffU + us * kUChunkMax, u_src + (size_t)h * 256 * kn, 256u * kn, &u_full[us]);
          ++us;
        }
        u_primed = true;
      }
      pdl_wait();
      for (int item = item0; item < n_items; item += item_step) {
        const int s = (item % n_groups) * CLS + (int)crank;
        const int mb = item / n_groups;
        const int kn = s < n48 ? 48 : 32;
        const int c0 = s < n48 ? 48 * s : 48 * n48 + 32 * (s - n48);
        const int t0 = mb * mv;
        const int ny0 = (t0 / 49) * 16 + 2 * ((t0 % 49) / 7);
        const uint8_t* u_src = reinterpret_cast<const uint8_t*>(u_img) + (size_t)n_kb * 512 * c0;
        for (int kb = 0; kb < n_kb; ++kb) {
          if constexpr (H16) {  // first of the two raw stages of this V stage (the second one follows below)
            mbar_wait(&raw_empty[rs], rph ^ 1);
            mbar_arrive_expect_tx(&raw_full[rs], kRawBytes);
            tma_tensor_4d_g2s(smem + kOffRaw + rs * kRawBytes, &tmap_x, kb * 16, 0, 0, ny0, &raw_full[rs]);
            if (++rs == kRawStages) { rs = 0; rph ^= 1; }
          }
          mbar_wait(&raw_empty[rs], rph ^ 1);
          mbar_arrive_expect_tx(&raw_full[rs], kRawBytes);
          if constexpr (CLS == 1) {
            tma_tensor_4d_g2s(smem + kOffRaw + rs * kRawBytes, &tmap_x, H16 ? kb * 16 + 8 : kb * 8, 0, 0, ny0,
                              &raw_full[rs]);
          } else {
            // this CTA is done with the buffer: tell every CTA of the cluster, wait until all of them are, then send
            // this CTA's share of the rows to everybody (each raw_full collects the CLS shares = kRawBytes)
#pragma unroll
            for (int p = 0; p < CLS; ++p) mbar_arrive_remote(&peer_empty[rs], (uint32_t)p);
            mbar_wait_cluster(&peer_empty[rs], rph);
            constexpr int kRowsPer = kRawRows / CLS;
            tma_tensor_4d_g2s_mcast(smem + kOffRaw + rs * kRawBytes + crank * (kRawBytes / CLS), &tmap_x, kb * 8, 0, 0,
                                    ny0 + (int)crank * kRowsPer, &raw_full[rs], (uint16_t)((1u << CLS) - 1u));
          }
          if (++rs == kRawStages) { rs = 0; rph ^= 1; }
          if (u_primed) {  // already requested above
            u_primed = false;
            continue;
          }
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&u_empty[us], uph ^ 1);
            mbar_arrive_expect_tx(&u_full[us], 256u * kn);
            tma_bulk_g2s(smem + kOffU + us * kUChunkMax, u_src + ((size_t)kb * 2 + h) * 256 * kn, 256u * kn,
                         &u_full[us]);
            if (++us == kUBufs) { us = 0; uph ^= 1; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (one elected thread)
    if (elect_one()) {
      const uint32_t u_base = smem_u32(smem + kOffU);
      uint32_t g = 0, us = 0, uph = 0, aph = 0;  // g = stages issued; V stage g % kVBufs, its phase (g / kVBufs) & 1
      for (int item = item0; item < n_items; item += item_step) {
        const int s = (item % n_groups) * CLS + (int)crank;
        const uint32_t kn = s < n48 ? 48u : 32u;
        const uint32_t fmt = H16 ? (fp16 ? kFmtF16 : kFmtBF16) : kFmtTF32;
        const uint32_t idesc = make_idesc(fmt, 128, kn);
        const uint32_t idesc_neg = make_idesc(fmt, 128, kn, 1);  // D += (-A) * B
        const uint32_t u_per_point = 2 * kn * 16, u_lbo = kn * 16;
        mbar_wait(acc_empty, aph ^ 1);  // epilogue of the previous item has drained TMEM
        tc_fence_after();
        for (int kb = 0; kb < n_kb; ++kb) {
          const uint32_t acc = kb > 0 ? 1u : 0u;
#pragma unroll
          for (int jh = 0; jh < 2; ++jh) {
            mbar_wait(&u_full[us], uph);  // there long before V: checked first, off the V -> MMA critical path
            const uint32_t vb = DB ? (g & 1) : 0u, vph = DB ? ((g >> 1) & 1) : (g & 1);
            mbar_wait(&v_full[vb * 2 + jh], vph);
            tc_fence_after();
            const uint32_t ua = u_base + us * kUChunkMax;
            const uint32_t va = tmem_base + kVCol0 + vb * 128 + jh * 64;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = jh * 2 + jj;
              const uint32_t z0 = tmem_base + (j * 2 + 0) * kAccStride;  // Z[0][j] = M0j + M1j + M2j
              const uint32_t z1 = tmem_base + (j * 2 + 1) * kAccStride;  // Z[1][j] = M1j - M2j - M3j
              uint64_t b_desc[4];
              uint32_t a_tm[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                a_tm[i] = va + (i * 2 + jj) * 8;
                b_desc[i] = make_smem_desc(ua + (i * 2 + jj) * u_per_point, u_lbo, 128, kLayoutNone);
              }
              umma_ts<H16>(z0, a_tm[1], b_desc[1], idesc, acc);  // first writer of both accumulators
              umma_ts<H16>(z1, a_tm[1], b_desc[1], idesc, acc);
              umma_ts<H16>(z0, a_tm[0], b_desc[0], idesc, 1u);
              umma_ts<H16>(z0, a_tm[2], b_desc[2], idesc, 1u);
              umma_ts<H16>(z1, a_tm[2], b_desc[2], idesc_neg, 1u);
              umma_ts<H16>(z1, a_tm[3], b_desc[3], idesc_neg, 1u);
            }
            umma_commit(&u_empty[us]);
            umma_commit(&v_empty[vb * 2 + jh]);  // this V half may be overwritten
            if (++us == kUBufs) { us = 0; uph ^= 1; }
          }
          ++g;
        }
        umma_commit(acc_full);
        aph ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ transform + epilogue warps
    // thread = (MMA row = TMEM lane, channel group cq of NC channels): warp w owns TMEM lanes 32*(w&3)..+31
    const int quad = warp & 3;
    const int cq = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t raw_base = smem_u32(smem + kOffRaw);
    const uint32_t stg_base = smem_u32(smem + kOffStg);
    const int half = (cq * NC) >> 2;                       // which 16-byte half of a pixel's 32 bytes
    const uint32_t sub = (uint32_t)((cq * NC) & 3) * 4;    // byte offset inside that half

    uint32_t rs = 0, rph = 0, g = 0, aph = 0;  // g = stages transformed (same counting as the MMA thread)
    for (int item = item0; item < n_items; item += item_step) {
      const int s = (item % n_groups) * CLS + (int)crank;
      const int mb = item / n_groups;
      const int kn = s < n48 ? 48 : 32;
      const int c0s = s < n48 ? 48 * s : 48 * n48 + 32 * (s - n48);
      const int t0 = mb * mv;
      const int ny0 = (t0 / 49) * 16 + 2 * ((t0 % 49) / 7);
      const int T = t0 + row;
      const int valid_rows = min(mv, total_tiles - t0);  // rows of this M-block that hold real tiles
      const bool tvalid = row < valid_rows;
      const bool warp_active = quad * 32 < valid_rows;  // warp-uniform
      const int n = T / 49, t = T % 49, ty = t / 7, tx = t % 7;
      const uint32_t raw_off = tvalid ? (uint32_t)((n * 16 + 2 * ty - ny0) * 512 + tx * 32) : 0u;
      // SWIZZLE_32B: the 16-byte half of a pixel's 32 bytes is XORed with bit 2 of its x/2 index (address bit 7)
      const uint32_t h0 = (uint32_t)((half ^ ((tx >> 2) & 1)) * 16) + sub;        // pixels with x/2 = tx
      const uint32_t h1 = (uint32_t)((half ^ (((tx + 1) >> 2) & 1)) * 16) + sub;  // pixels with x/2 = tx + 1

      for (int kb = 0; kb < n_kb; ++kb) {
        if (!warp_active) {
          // nothing to transform: release the raw stage(s) and report "V ready" in step with the other warps
#pragma unroll
          for (int sb = 0; sb < kSub; ++sb) {
            mbar_wait(&raw_full[rs], rph);
            if (lane == 0) mbar_arrive(&raw_empty[rs]);
            if (++rs == kRawStages) { rs = 0; rph ^= 1; }
          }
          const uint32_t vb = DB ? (g & 1) : 0u, vph = DB ? ((g >> 1) & 1) : (g & 1);
#pragma unroll
          for (int jh = 0; jh < 2; ++jh) {
            mbar_wait(&v_empty[vb * 2 + jh], vph ^ 1);
            if (lane == 0) mbar_arrive(&v_full[vb * 2 + jh]);
          }
          ++g;
          continue;
        }
        const uint32_t vb = DB ? (g & 1) : 0u, vph = DB ? ((g >> 1) & 1) : (g & 1);
#pragma unroll
        for (int sb = 0; sb < kSub; ++sb) {  // H16: two 8-channel raw stages fill one 16-channel V stage
        mbar_wait(&raw_full[rs], rph);
        float d[4][4][NC];
        if (tvalid) {
          const uint32_t a = raw_base + rs * kRawBytes + raw_off;
#pragma unroll
          for (int dy = 0; dy < 4; ++dy)
#pragma unroll
            for (int dx = 0; dx < 4; ++dx) {
              const uint32_t ad = a + dy * 512 + (dx & 1) * 256 + (dx >> 1) * 32 + ((dx >> 1) ? h1 : h0);
              if constexpr (NC == 4) {
                const float4 v = ld_shared_v4(ad);
                d[dy][dx][0] = v.x, d[dy][dx][1] = v.y, d[dy][dx][2] = v.z, d[dy][dx][3] = v.w;
              } else {
                const float2 v = ld_shared_v2(ad);
                d[dy][dx][0] = v.x, d[dy][dx][1] = v.y;
              }
            }
        } else {
#pragma unroll
          for (int dy = 0; dy < 4; ++dy)
#pragma unroll
            for (int dx = 0; dx < 4; ++dx)
#pragma unroll
              for (int c = 0; c < NC; ++c) d[dy][dx][c] = 0.f;
        }
        // column pass t = B^T d, in place over dy
#pragma unroll
        for (int dx = 0; dx < 4; ++dx)
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float d0 = d[0][dx][c], d1 = d[1][dx][c], d2 = d[2][dx][c], d3 = d[3][dx][c];
            d[0][dx][c] = d0 - d2;
            d[1][dx][c] = d1 + d2;
            d[2][dx][c] = d2 - d1;
            d[3][dx][c] = d1 - d3;
          }
        // the raw stage is in registers now: hand it back to the producer before the row pass
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[rs]);
        if (++rs == kRawStages) { rs = 0; rph ^= 1; }

        // row pass V = t B by halves (half jh = points with j in {2jh, 2jh+1}), round to the operand type, store into
        // TMEM (tf32: one column per channel; 16-bit: one column per channel pair, raw stage sb fills columns 4*sb..)
        const uint32_t vcol = tmem_base + lane_base + kVCol0 + vb * 128 +
                              (uint32_t)(H16 ? sb * 4 + cq * 2 : cq * NC);
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
          if (sb == 0) {
            mbar_wait(&v_empty[vb * 2 + jh], vph ^ 1);  // the MMAs that last read this V half have completed
            tc_fence_after();
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float v0[NC], v1[NC];  // points (i, 2jh) and (i, 2jh+1)
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              const float a0 = d[i][0][c], a1 = d[i][1][c], a2 = d[i][2][c], a3 = d[i][3][c];
              v0[c] = jh == 0 ? a0 - a2 : a2 - a1;
              v1[c] = jh == 0 ? a1 + a2 : a1 - a3;
              if constexpr (!H16) v0[c] = tm_tf32(v0[c]), v1[c] = tm_tf32(v1[c]);
            }
            const uint32_t dst = vcol + jh * 64 + (i * 2) * 8;
            if constexpr (H16) {
              tmem_st_x2(dst, pack16(v0[0], v0[1], fp16), pack16(v0[2], v0[3], fp16));
              tmem_st_x2(dst + 8, pack16(v1[0], v1[1], fp16), pack16(v1[2], v1[3], fp16));
            } else if constexpr (NC == 4) {
              tmem_st_x4(dst, v0[0], v0[1], v0[2], v0[3]);
              tmem_st_x4(dst + 8, v1[0], v1[1], v1[2], v1[3]);
            } else {
              tmem_st_x2(dst, v0[0], v0[1]);
              tmem_st_x2(dst + 8, v1[0], v1[1]);
            }
          }
          if (sb == kSub - 1) {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&v_full[vb * 2 + jh]);
          }
        }
        }  // sb
        ++g;
      }

      // ---- epilogue: rest of Y = A^T M A, BN, ReLU, staged, written out as full runs per pixel
      const bool evalid = tvalid;
      const int W = out_padded ? 16 : 14;
      const int o = out_padded ? 1 : 0;
      const int pix0 = evalid ? ((n * W + 2 * ty + o) * W + 2 * tx + o) : -1;  // first output pixel of this tile
      if (cq == 0) pixtab[row] = pix0;
      const uint32_t tstride = (uint32_t)(16 * kn + 16);
      const int ncc = kn / (WW / 4);  // couts drained by this warp: [cq*ncc, (cq+1)*ncc)

      mbar_wait(acc_full, aph);
      aph ^= 1;
      tc_fence_after();
      if (warp_active) {
#pragma unroll 1
        for (int cc = 0; cc < ncc; cc += EC) {
          const int c0 = cq * ncc + cc;
          const uint32_t taddr = tmem_base + lane_base + c0;
          float z[8][EC];  // z[j*2 + a][e]
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            if constexpr (EC == 8) tmem_ld_x8(taddr + p * kAccStride, z[p]);
            else tmem_ld_x4(taddr + p * kAccStride, z[p]);
          }
          tmem_ld_wait();
          float sc[EC], sh[EC];
#pragma unroll
          for (int q4 = 0; q4 < EC / 4; ++q4) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale + c0s + c0 + 4 * q4));
            const float4 h4 = __ldg(reinterpret_cast<const float4*>(shift + c0s + c0 + 4 * q4));
            sc[4 * q4] = s4.x, sc[4 * q4 + 1] = s4.y, sc[4 * q4 + 2] = s4.z, sc[4 * q4 + 3] = s4.w;
            sh[4 * q4] = h4.x, sh[4 * q4 + 1] = h4.y, sh[4 * q4 + 2] = h4.z, sh[4 * q4 + 3] = h4.w;
          }
          float ov[4][EC];  // Y[a][b] at ov[2*a + b]
#pragma unroll
          for (int e = 0; e < EC; ++e) {
            ov[0][e] = fmaf(sc[e], z[0][e] + z[2][e] + z[4][e], sh[e]);
            ov[1][e] = fmaf(sc[e], z[2][e] - z[4][e] - z[6][e], sh[e]);
            ov[2][e] = fmaf(sc[e], z[1][e] + z[3][e] + z[5][e], sh[e]);
            ov[3][e] = fmaf(sc[e], z[3][e] - z[5][e] - z[7][e], sh[e]);
            if (relu) {
#pragma unroll
              for (int p = 0; p < 4; ++p) ov[p][e] = fmaxf(ov[p][e], 0.f);
            }
          }
          const uint32_t sdst = stg_base + (uint32_t)row * tstride + (uint32_t)c0 * 4;
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q4 = 0; q4 < EC / 4; ++q4)
              st_shared_v4(sdst + p * (4 * kn) + 16 * q4, ov[p][4 * q4], ov[p][4 * q4 + 1], ov[p][4 * q4 + 2],
                           ov[p][4 * q4 + 3]);
        }
      }
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * WW) : "memory");  // staging + pixel table complete, TMEM drained
      if (lane == 0) mbar_arrive(acc_empty);
      {
        const int tid = threadIdx.x;  // 0 .. 32*WW-1
        const int chunks = kn / 4;    // 16-byte chunks per pixel
        const int units = valid_rows * kn;  // (tile, pixel, chunk)
        for (int u = tid; u < units; u += kWorkerWarps * 32) {
          const int tile = u / kn;
          const int r = u - tile * kn;
          const int px = r / chunks;
          const int ch = r - px * chunks;
          const float4 v = ld_shared_v4(stg_base + (uint32_t)tile * tstride + (uint32_t)(px * 4 * kn + ch * 16));
          const int pix = pixtab[tile] + (px >> 1) * W + (px & 1);
          st_out_v4(y + (size_t)pix * K + c0s + ch * 4, v, mc);
        }
        if (out_padded && evalid && (ty == 0 || ty == 6 || tx == 0 || tx == 6)) {
          // zero border of the reference's 16x16 frame (Kernel128_winograd.cu:163,243): edge tiles own their share
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const size_t rstride = (size_t)W * K;
          float* p = y + (size_t)pix0 * K + c0s + cq * ncc;
          const ptrdiff_t dyb = ty == 0 ? -(ptrdiff_t)rstride : (ty == 6 ? 2 * (ptrdiff_t)rstride : 0);
          const ptrdiff_t dxb = tx == 0 ? -(ptrdiff_t)K : (tx == 6 ? 2 * (ptrdiff_t)K : 0);
#pragma unroll 1
          for (int e = 0; e < ncc; e += 4) {
            if (dyb != 0) {
              st_out_v4(p + dyb + e, z4, mc);
              st_out_v4(p + dyb + K + e, z4, mc);
            }
            if (dxb != 0) {
              st_out_v4(p + dxb + e, z4, mc);
              st_out_v4(p + dxb + rstride + e, z4, mc);
            }
            if (dyb != 0 && dxb != 0) st_out_v4(p + dyb + dxb + e, z4, mc);
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * WW) : "memory");  // staging area and pixel table free for the next item
    }
  }

  if constexpr (CLS > 1) {  // no CTA leaves while a peer may still multicast into it or arrive on its barriers
    __syncwarp();
    cluster_sync_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// Once per layer: U = G g G^T (F(2x2,3x3)), RN-rounded to TF32, in the shared-memory image of this kernel's bulk
// copies: cout slices of 48 (n48 of them) then 32; per slice [C/8 k-block][2 j-halves][4 i][2 jj][2 k-chunks][KN couts]
// [4 channels] (j = 2*jh + jj). 512 bytes per (k-block, cout), so slice s starts at byte (C/8)*512*c0(s).
// Replaces the offline weight_generator loop (/root/reference/data_generator.py:63-78; that one is F(4x4), 36 points).
// op16 = 1 (bf16) / 2 (fp16): 16-channel k-blocks, 8 channels per 16-byte chunk, same 512 bytes per (k-block, cout).
__global__ void filter_transform_tm_kernel(const float* __restrict__ w_kcrs, float* __restrict__ u_img, int C, int K,
                                           int n48, int op16) {
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
  float tt[4][3];  // t = G g
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    tt[0][s] = gg[0][s];
    tt[1][s] = 0.5f * (gg[0][s] + gg[1][s] + gg[2][s]);
    tt[2][s] = 0.5f * (gg[0][s] - gg[1][s] + gg[2][s]);
    tt[3][s] = gg[2][s];
  }
  int kn, c0;
  if (k < 48 * n48) {
    kn = 48;
    c0 = (k / 48) * 48;
  } else {
    kn = 32;
    c0 = 48 * n48 + ((k - 48 * n48) / 32) * 32;
  }
  const int kl = k - c0;
  const int kb = ch / 8, chunk = (ch % 8) / 4, e = ch % 4;
  const size_t base = (size_t)(C / 8) * 128 * c0 + (size_t)kb * 128 * kn;  // floats
  // 16-bit image, in 2-byte elements: 256 per (16-channel block, cout)
  const int kb16 = ch / 16, chunk16 = (ch % 16) / 8, e16 = ch % 8;
  const size_t base16 = (size_t)(C / 16) * 256 * c0 + (size_t)kb16 * 256 * kn;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float u[4];
    u[0] = tt[i][0];
    u[1] = 0.5f * (tt[i][0] + tt[i][1] + tt[i][2]);
    u[2] = 0.5f * (tt[i][0] - tt[i][1] + tt[i][2]);
    u[3] = tt[i][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = (j >> 1) * 8 + i * 2 + (j & 1);
      const size_t o16 = base16 + (size_t)p * (2 * kn * 8) + ((size_t)chunk16 * kn + kl) * 8 + e16;
      if (op16 == 2) reinterpret_cast<__half*>(u_img)[o16] = __float2half_rn(u[j]);
      else if (op16 == 1) reinterpret_cast<__nv_bfloat16*>(u_img)[o16] = __float2bfloat16_rn(u[j]);
      else u_img[base + (size_t)p * (2 * kn * 4) + ((size_t)chunk * kn + kl) * 4 + e] = to_tf32_rn(u[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side

int wino_tm_n48(int K) {  // 48*a + 32*b = K with a as large as possible
  const int m = K / 16;
  int a = m / 3;
  while (a > 0 && ((m - 3 * a) & 1)) --a;
  return a;
}

// Cluster size of the throughput kernel for K output channels: the cout slices of an M-block are spread over the CTAs
// of a cluster, so it must divide their number. Default 1 (no clusters); WG_WINO_CLS=2|3 enables the experiment.
int wino_tm_cls(int K, int db) {
  static int env = -1;
  if (env < 0) {
    const char* e = dev_env("WG_WINO_CLS");
    env = e ? atoi(e) : 0;
  }
  static int ww16 = -1;  // the 16-worker-warp experiment has no cluster variant
  if (ww16 < 0) {
    const char* e = dev_env("WG_WINO_WW");
    ww16 = (e && atoi(e) == 16) ? 1 : 0;
  }
  // Measured (256->256, N=256): clusters of 2 -> 216 us, of 3 -> 478 us, against 151 us without: the per-stage
  // cross-CTA release/acquire handshake and the lock-step of the CTAs cost far more than the TMA requests they save.
  // Off unless asked for.
  if (db || ww16 || env < 2) return 1;
  const int n48 = wino_tm_n48(K);
  const int n_slices = n48 + (K - 48 * n48) / 32;
  return (env == 2 || env == 3) && n_slices % env == 0 ? env : 1;
}

int wino_tm_make_tmap(CUtensorMap* tmap, const float* x, int n_img, int C, int cls) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return WG_ERR_DRIVER;
  // x[N][16][16][C] viewed as (c, x/2, x&1, n*16+y), 32-byte swizzle on the 8-channel inner box
  cuuint64_t dims[4] = {(cuuint64_t)C, 8, 2, (cuuint64_t)n_img * 16};
  cuuint64_t strides[3] = {(cuuint64_t)2 * C * 4, (cuuint64_t)C * 4, (cuuint64_t)16 * C * 4};
  cuuint32_t box[4] = {8, 8, 2, (cuuint32_t)(tm::kRawRows / cls)};  // a cluster of cls CTAs loads the box in cls row shares
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, wg::l2_promotion(),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WG_OK : WG_ERR_TMAP;
}

// Which TM variant a layer uses (decided once, at create: the filter image differs). WG_WINO_TM=48|32 pins it.
int wino_tm_choose_db(int C, int K) {
  static int env = -1;
  if (env < 0) {
    const char* e = dev_env("WG_WINO_TM");
    env = e ? atoi(e) : 0;
  }
  if (env == 48) return 0;
  if (env == 32) return 1;
  (void)C;
  (void)K;
  return 0;
}

int filter_transform_tm_launch(const float* w_kcrs, float* u_img, int C, int K, int db, int op16,
                               cudaStream_t stream) {
  const int n = C * K;
  filter_transform_tm_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w_kcrs, u_img, C, K, db ? 0 : wino_tm_n48(K), op16);
  return cudaGetLastError() == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

template <bool DB, int WW, int CLS, bool H16 = false>
static int launch_tm(const CUtensorMap& tmap, const float* u_img, const float* scale, const float* shift, float* y,
                     int n_img, int C, int K, int relu, int out_padded, int max_ctas, cudaStream_t stream,
                     int fp16 = 0) {
  static unsigned long long configured = 0;  // per device: the attribute is a property of the function on ONE device
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long dev_bit_ = 1ull << (dev_ & 63);
  if (!(configured & dev_bit_)) {
    if (cudaFuncSetAttribute(wino3x3_tm_kernel<DB, WW, CLS, H16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tm::kTotal) !=
        cudaSuccess)
      return WG_ERR_CUDA;
    configured |= dev_bit_;
  }
  const int n48 = DB ? 0 : wino_tm_n48(K);
  const int n_slices = n48 + (K - 48 * n48) / 32;
  const int total_tiles = n_img * 49;
  // Tiles per M-block: the MMA is always M=128 but only `mv` rows carry tiles; transform warps own 32 rows each, so
  // the per-item cost scales with ceil(mv/32) quarters. Pick the mv that minimises waves x cost (WG_WINO_MV pins it).
  int mv = 128;
  static int mv_env = -1;
  if (mv_env < 0) {
    const char* e = dev_env("WG_WINO_MV");
    mv_env = e ? atoi(e) : 0;
  }
  if (mv_env >= 16 && mv_env <= 128) {
    mv = mv_env;
  } else {
    double best = 1e30;
    for (int cand = 128; cand >= 64; cand -= 32) {
      const long long items = (long long)((total_tiles + cand - 1) / cand) * (n_slices / CLS);
      const long long slots = max_ctas / CLS > 0 ? max_ctas / CLS : 1;
      const long long waves = (items + slots - 1) / slots;
      const double cost = (double)waves * (0.35 + 0.65 * cand / 128.0);
      if (cost < best - 1e-9) {
        best = cost;
        mv = cand;
      }
    }
  }
  const int n_items = ((total_tiles + mv - 1) / mv) * (n_slices / CLS);  // items per cluster
  int n_cl = max_ctas / CLS;
  if (n_cl > n_items) n_cl = n_items;
  if (n_cl < 1) n_cl = 1;
  const int grid = n_cl * CLS;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(32 * (WW + 2));
  cfg.dynamicSmemBytes = tm::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CLS > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CLS;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  // (A build of this kernel with switches for "no patch loads / TMEM stores" and "no MMAs" gave, 256->256 N=256:
  //  153 us full, 111 us without the transform, 130 us without the MMAs, 95 us with neither; profiles/README.md.)
  cudaError_t e = cudaLaunchKernelEx(&cfg, wino3x3_tm_kernel<DB, WW, CLS, H16>, tmap, u_img, scale, shift, y, n_img, C,
                                     K, n48, relu, out_padded, mv, fp16);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

int wino_tm_launch(const CUtensorMap& tmap, const float* u_img, const float* scale, const float* shift, float* y,
                   int n_img, int C, int K, int db, int op16, int relu, int out_padded, int max_ctas,
                   cudaStream_t stream) {
  // op16: 0 = TF32 operands, 1 = bf16, 2 = fp16 (V packed in TMEM, 16-channel stages)
  if (op16)
    return launch_tm<false, 8, 1, true>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded, max_ctas, stream,
                                        op16 == 2);
  static int ww = -1;  // WG_WINO_WW=8|16: worker warps (4 or 2 channels per thread); 16 and db are experiments
  if (ww < 0) {
    const char* e = dev_env("WG_WINO_WW");
    ww = e ? atoi(e) : 8;
    if (ww != 16) ww = 8;
  }
#define WG_TM(DB_, WW_, CLS_) \
  return launch_tm<DB_, WW_, CLS_>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded, max_ctas, stream)
  if (db) {
    if (ww == 16) WG_TM(true, 16, 1);
    WG_TM(true, 8, 1);
  }
  if (ww == 16) WG_TM(false, 16, 1);
  const int cls = wino_tm_cls(K, db);
  if (cls == 3) WG_TM(false, 8, 3);
  if (cls == 2) WG_TM(false, 8, 2);
  WG_TM(false, 8, 1);
#undef WG_TM
}

}  // namespace wg
