// Full-fold 3x3 Winograd kernel (see wino_ff_kernel.cu for the algorithm, the TMEM / shared-memory layout and the
// barrier scheme) with SIXTEEN transform warps: one group of 8 warps per V half.
//
// In wino3x3_ff_kernel a transform thread owns a tile and 4 channels and produces BOTH V halves of a stage one after the
// other (16 patch loads -> column pass -> row pass + TMEM stores of half 0 -> ... of half 1); that serial chain
// (~2.1 k clk per 8-channel stage) is what the bf16 / fp16 kernel waits for, and what the TF32 kernel's MMA chain only
// partially overlaps. Here warps 0-7 produce half 0 (points with j in {0,1}: patch columns 0..2) and warps 8-15 half 1
// (j in {2,3}: patch columns 1..3), each thread loading 12 instead of 16 pixels: the two halves are transformed
// concurrently, the per-stage chain of a thread shrinks to 12 loads + 48 + 32 FADDs + 8 TMEM stores, at the price of 1.5x
// the patch-load traffic and a duplicated column pass for the two shared patch columns.
//
// Same filter images, tensor map, raw layout (parity planes, 9-slot pitch), accumulators and epilogue staging as the
// 8-warp kernel; the epilogue splits the 32-cout chunk over the four warps that share a TMEM lane quadrant.
// Measured (N=256): bf16 / fp16 operands 256->256 82 -> 77 us, TF32 +-0 (the longer patch-load phase -- 864 instead of 585
// shared-memory wavefronts per stage -- eats what the parallel row passes save). This kernel also carries the modes for
// launches that do not fill the SMs (wino_ff_launch decides): the layer's second filter image with 64-wide slices
// (`narrow`), and split-C (SPLIT).
// Replaces kernel_{128,256}_winograd_BtdB -> kernel_*_OuterProduct_* -> kernel_*_winograd_AtIA
// (/root/reference/Kernel128_winograd.cu:28-213, Kernel256_winograd.cu:27-218).
#include "wino_ff_common.cuh"
#include "wg_internal.h"

#include <cuda.h>
#include <stdlib.h>

namespace wg {

namespace ffw {
constexpr int kWorkerWarps = 16, kProducerWarp = 16, kMmaWarp = 17;
constexpr int kThreads = 32 * (kWorkerWarps + 2);
}  // namespace ffw

// Row pass V = t B for one V half. d[i][c] = column-pass output of patch column c + JH (c = 0..2):
// JH = 0: points (i,0) = t0 - t2, (i,1) = t1 + t2; JH = 1: points (i,2) = t2 - t1, (i,3) = t1 - t3.
template <int JH, bool H16>
__device__ __forceinline__ void ffw_row_pass(const float (&d)[4][3][4], uint32_t vcol, int fp16) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v0[4], v1[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x0 = d[i][0][e], x1 = d[i][1][e], x2 = d[i][2][e];
      v0[e] = JH ? x1 - x0 : x0 - x2;
      v1[e] = JH ? x0 - x2 : x1 + x2;
      if constexpr (!H16) v0[e] = ff_tf32(v0[e]), v1[e] = ff_tf32(v1[e]);
    }
    const uint32_t dst = vcol + (i * 2) * 8;
    if constexpr (H16) {
      tmem_st_x2(dst, ff_pack16(v0[0], v0[1], fp16), ff_pack16(v0[2], v0[3], fp16));
      tmem_st_x2(dst + 8, ff_pack16(v1[0], v1[1], fp16), ff_pack16(v1[2], v1[3], fp16));
    } else {
      tmem_st_x4(dst, v0[0], v0[1], v0[2], v0[3]);
      tmem_st_x4(dst + 8, v1[0], v1[1], v1[2], v1[3]);
    }
  }
}

// CG2: CTA pairs (clusters of 2, tcgen05 cta_group::2), exactly as in wino3x3_ff_kernel: neighbouring M-blocks, same cout
// slice, MMAs with M = 256 issued by the leader CTA, each CTA supplying half of the couts of every filter chunk.
// SPLIT: split-C for batches whose work items do not fill the SMs (one wave, each cluster exactly one item). The two
// CTAs of a cluster work on the SAME item, each on half of the channel loop; afterwards CTA r owns tile rows
// [64r, 64r + 64): the warps of the other two TMEM lane quadrants push their partial accumulators into the owner's shared
// memory (st.shared::cluster into a bank-swizzled inbox laid over the pipeline buffers, which are free by then -- the
// owner says so with an mbarrier arrive), the owner's warps add them in the epilogue. Halves an item's stage count at
// the price of 98 KB over DSMEM each way: 256->256 at N=12..64 ~43 -> ~25 us.
template <bool H16, bool CG2, bool SPLIT>
__global__ void __launch_bounds__(ffw::kThreads, 1)
wino3x3_ffw_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ u_img,
                   const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ y, int n_img,
                   int C, int K, int relu, int out_padded, int mv, int fp16, int narrow) {
  using namespace ff;
  using ffw::kMmaWarp;
  using ffw::kProducerWarp;
  using ffw::kWorkerWarps;
  constexpr int kSub = H16 ? 2 : 1;       // 8-channel raw stages per V stage
  const bool mc = (out_padded & 2) != 0;  // y is an NVLS multicast address: stores go out as multimem.st
  out_padded &= 1;
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = raw_full + kRawStages;
  uint64_t* full = raw_empty + kRawStages;  // [2 * (stage & 1) + half]: V half stored (8 warps) + filter chunk landed
  uint64_t* done = full + 4;                // [same]: that half's 18 MMAs have completed
  uint64_t* acc_full = done + 4;
  uint64_t* acc_empty = acc_full + 1;
  uint64_t* u_land = acc_empty + 1;  // [4] CG2, peer CTA only: its half of a filter chunk has landed
  uint64_t* may_push = u_land + 4;   // SPLIT: the other CTA's main loop is done, its pipeline buffers may be overwritten
  uint64_t* inbox_full = may_push + 1;  // SPLIT: the other CTA's 8 pushing warps have delivered their partial sums
  static_assert(!(CG2 && SPLIT), "one cluster role at a time");
  const uint32_t crank = (CG2 || SPLIT) ? cluster_ctarank() : 0u;
  const bool peer = CG2 && crank != 0;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  int* pixtab = reinterpret_cast<int*>(smem + kOffPix);

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < kRawStages; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], kWorkerWarps);
    }
    for (int i = 0; i < 4; ++i) {
      // CG2 (leader): the half's 8 warps of both CTAs + own TMA bytes + the peer's relay
      mbar_init(&full[i], CG2 ? kWorkerWarps + 2 : kWorkerWarps / 2 + 1);
      mbar_init(&done[i], 1);
      mbar_init(&u_land[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, CG2 ? 2 * kWorkerWarps : kWorkerWarps);
    mbar_init(may_push, 1);
    mbar_init(inbox_full, 32 * kWorkerWarps / 2);  // every pushing thread releases its own stores
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    if constexpr (CG2) tmem_alloc_cg2<512>(tmem_ptr);
    else tmem_alloc<512>(tmem_ptr);
  }
  tc_fence_before();
  if constexpr (CG2 || SPLIT) cluster_sync_all();  // the peer's barriers exist before anybody arrives on them remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_kb_all = C / (8 * kSub);
  const int n_kb = SPLIT ? n_kb_all / 2 : n_kb_all;      // V stages this CTA runs per item
  const int kb_off = SPLIT ? (int)crank * n_kb : 0;      // ... starting at this stage of the layer's channel loop
  const int n_sl = n_slices(K, narrow);
  const int total_tiles = n_img * 49;
  const int n_mblocks = (total_tiles + mv - 1) / mv;
  const int n_items = (CG2 ? (n_mblocks + 1) / 2 : n_mblocks) * n_sl;  // CG2: one item per CTA pair
  const int item0 = (CG2 || SPLIT) ? blockIdx.x / 2 : blockIdx.x, item_step = (CG2 || SPLIT) ? gridDim.x / 2 : gridDim.x;
  const uint32_t u_bytes_per_kn = CG2 ? 128u : 256u;  // bytes of a filter chunk this CTA loads, per cout of the slice

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t rs = 0, rph = 0, us = 0, uph = 0;
      int u_primed = 0;
      if (item0 < n_items) {  // the filter does not depend on the previous kernel in the stream
        const Slice sl = slice(K, item0 % n_sl, narrow);
        const uint8_t* u_src = reinterpret_cast<const uint8_t*>(u_img) + (size_t)n_kb_all * 512 * sl.c0 +
                               (size_t)kb_off * 512 * sl.kn;
        for (int h = 0; h < 2; ++h) {
          uint64_t* ubar = peer ? &u_land[us] : &full[us];
          mbar_arrive_expect_tx(ubar, u_bytes_per_kn * sl.kn);
          tma_bulk_g2s(smem + kOffU + us * kUChunkMax, u_src + ((size_t)h * 256 + (CG2 ? crank * 128 : 0)) * sl.kn,
                       u_bytes_per_kn * sl.kn, ubar);
          ++us;
        }
        u_primed = 1;
      }
      pdl_wait();
      for (int item = item0; item < n_items; item += item_step) {
        const Slice sl = slice(K, item % n_sl, narrow);
        const int kn = sl.kn;
        const int t0 = (CG2 ? (item / n_sl) * 2 + (int)crank : item / n_sl) * mv;
        const int ny0 = (t0 / 49) * 16 + 2 * ((t0 % 49) / 7);
        const uint8_t* u_src = reinterpret_cast<const uint8_t*>(u_img) + (size_t)n_kb_all * 512 * sl.c0 +
                               (size_t)kb_off * 512 * sl.kn;
        for (int kb = 0; kb < n_kb; ++kb) {
#pragma unroll
          for (int sb = 0; sb < kSub; ++sb) {
            mbar_wait(&raw_empty[rs], rph ^ 1);
            mbar_arrive_expect_tx(&raw_full[rs], kRawBytesP9);
#pragma unroll
            for (int q = 0; q < 4; ++q)  // plane q = (y parity q>>1, x parity q&1); x/2 starts at -1 (zero-filled)
              tma_tensor_5d_g2s(smem + kOffRaw + rs * kRawStride + q * kPlaneBytes, &tmap_x,
                                ((kb_off + kb) * kSub + sb) * 8, -1, q & 1, q >> 1, ny0 >> 1, &raw_full[rs]);
            if (++rs == kRawStages) { rs = 0; rph ^= 1; }
          }
          if (u_primed) {
            u_primed = 0;
            continue;
          }
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&done[us], uph ^ 1);
            uint64_t* ubar = peer ? &u_land[us] : &full[us];
            mbar_arrive_expect_tx(ubar, u_bytes_per_kn * kn);
            tma_bulk_g2s(smem + kOffU + us * kUChunkMax, u_src + (((size_t)kb * 2 + h) * 256 + (CG2 ? crank * 128 : 0)) * kn,
                         u_bytes_per_kn * kn, ubar);
            if (++us == kUBufs) { us = 0; uph ^= 1; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (one elected thread)
    if (peer) {
      // CG2, peer CTA: no MMAs to issue here; relay "my half of filter chunk us has landed" to the leader's full[us]
      if (elect_one()) {
        uint32_t us = 0, uph = 0;
        for (int item = item0; item < n_items; item += item_step)
          for (int c = 0; c < 2 * n_kb; ++c) {
            mbar_wait(&u_land[us], uph);
            mbar_arrive_remote_plain(&full[us], 0);
            if (++us == kUBufs) { us = 0; uph ^= 1; }
          }
      }
    } else if (elect_one()) {
      const uint32_t u_base = smem_u32(smem + kOffU);
      uint32_t us = 0, uph = 0, aph = 0, gm = 0;  // gm = V stages issued
      for (int item = item0; item < n_items; item += item_step) {
        const uint32_t kn = (uint32_t)slice(K, item % n_sl, narrow).kn;
        const uint32_t fmt = H16 ? (fp16 ? kFmtF16 : kFmtBF16) : kFmtTF32;
        const uint32_t idesc_pos = make_idesc(fmt, CG2 ? 256 : 128, kn);
        const uint32_t idesc_neg = make_idesc(fmt, CG2 ? 256 : 128, kn, 1);  // D += (-A) * B
        // CG2: this CTA's shared memory holds kn/2 couts of every point
        const uint32_t u_per_point = (CG2 ? 1 : 2) * kn * 16, u_lbo = (CG2 ? kn / 2 : kn) * 16;
        // slices of 64 or fewer couts: two V stages in TMEM (accumulator p at 64 p, V stage (g & 1) at 256 + 128 (g & 1)),
        // see wino_ff_kernel.cu
        const bool db = !CG2 && kn <= 64;
        const uint32_t acc_stride = db ? 64u : kAccStride, v_col0 = db ? 256u : kVCol0;
        mbar_wait(acc_empty, aph ^ 1);  // epilogue of the previous item has drained TMEM
        tc_fence_after();
        for (int kb = 0; kb < n_kb; ++kb, ++gm) {
          uint32_t written = kb > 0 ? 0xFu : 0u;  // bit p set = accumulator p has been written in this item
#pragma unroll
          for (int jh = 0; jh < 2; ++jh) {
            mbar_wait(&full[us], uph);  // filter chunk landed and V half stored by its transform warps
            tc_fence_after();
            const uint32_t ua = u_base + us * kUChunkMax;
            const uint32_t va = tmem_base + v_col0 + (db ? (gm & 1) * 128 : 0u) + jh * 64;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = jh * 2 + jj;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t a_tm = va + (i * 2 + jj) * 8;
                const uint64_t b_desc = make_smem_desc(ua + (i * 2 + jj) * u_per_point, u_lbo, 128, kLayoutNone);
                // A^T = [[1,1,1,0],[0,1,-1,-1]]: row a uses point index i with sign sa, column b uses j with sign sb
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                  if ((a == 0 && i == 3) || (a == 1 && i == 0)) continue;
                  const int sa = (a == 1 && i >= 2) ? -1 : 1;
#pragma unroll
                  for (int b = 0; b < 2; ++b) {
                    if ((b == 0 && j == 3) || (b == 1 && j == 0)) continue;
                    const int sb = (b == 1 && j >= 2) ? -1 : 1;
                    const uint32_t p = (uint32_t)(2 * a + b);
                    ff_umma<H16, CG2>(tmem_base + p * acc_stride, a_tm, b_desc, sa * sb > 0 ? idesc_pos : idesc_neg,
                                        (written >> p) & 1u);
                    written |= 1u << p;
                  }
                }
              }
            }
            if constexpr (CG2) umma_commit_mcast_cg2(&done[us], 3);  // frees the filter chunk and this V half, both CTAs
            else umma_commit(&done[us]);
            if (++us == kUBufs) { us = 0; uph ^= 1; }
          }
        }
        if constexpr (CG2) umma_commit_mcast_cg2(acc_full, 3);
        else umma_commit(acc_full);
        aph ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ transform + epilogue warps
    // thread = (tile = TMEM lane, channel half cq, V half jh): warp w owns TMEM lanes 32*(w&3)..+31
    const int quad = warp & 3;
    const int cq = (warp >> 2) & 1;
    const int jh = warp >> 3;
    const int esub = warp >> 2;  // epilogue: which 8 couts of a 32-cout chunk
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t raw_base = smem_u32(smem + kOffRaw);
    const uint32_t stg_base = smem_u32(smem + kOffStg);

    uint32_t rs = 0, rph = 0, g = 0, aph = 0;  // g = V stages transformed (same counting as the MMA thread)
    // CG2: full[] / acc_empty live in the leader CTA (the payload of these hand-offs is TMEM state, ordered by the
    // tcgen05 fences, so a plain remote arrive will do); done[] / acc_full arrive by the leader's multicast commit
    auto arrive_leader = [&](uint64_t* bar) {
      if (peer) mbar_arrive_remote_plain(bar, 0);
      else mbar_arrive(bar);
    };
    for (int item = item0; item < n_items; item += item_step) {
      const Slice sl = slice(K, item % n_sl, narrow);
      const int kn = sl.kn, c0s = sl.c0;
      const int t0 = (CG2 ? (item / n_sl) * 2 + (int)crank : item / n_sl) * mv;
      const int ny0 = (t0 / 49) * 16 + 2 * ((t0 % 49) / 7);
      const int T = t0 + row;
      const int valid_rows = min(mv, total_tiles - t0);  // rows of this M-block that hold real tiles
      const bool tvalid = row < valid_rows;
      const bool warp_active = quad * 32 < valid_rows;  // warp-uniform
      const bool db = !CG2 && kn <= 64;                 // two V stages in TMEM (see the MMA thread)
      const uint32_t acc_stride = db ? 64u : kAccStride, v_col0 = db ? 256u : kVCol0;
      const int n = T / 49, t = T % 49, ty = t / 7;
      const int tx = 6 - t % 7;  // tiles run right-to-left inside a tile row (conflict-free slots, wino_ff_common.cuh)
      // byte offset, inside the raw stage, of patch pixel (dy, c + jh) for c = 0..2: plane (dy&1, dx&1) + slot + half
      uint32_t poff[2][3];  // [dy >> 1][c]; add (dy & 1) * 2 * kPlaneBytes
      {
        const uint32_t s0 = tvalid ? (uint32_t)(((n * 16 + 2 * ty - ny0) >> 1) * 9 + tx + 1) : 1u;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const uint32_t dx = (uint32_t)(c + jh);
            const uint32_t slot = s0 + 9 * a + (dx >> 1);
            poff[a][c] = (dx & 1) * kPlaneBytes + slot * 32 + (uint32_t)((cq ^ ((slot >> 2) & 1)) * 16);
          }
      }

      for (int kb = 0; kb < n_kb; ++kb) {
        // this stage's ring slot for this half; the previous stage's (whose MMAs must have completed before the V half
        // is overwritten) completes for the ((g - 1) >> 1)-th time
        // (two V stages: the MMAs of stage g - 2, which used this very slot; the first stage(s) of an item wait for
        // nothing -- this thread has passed the previous item's acc_full)
        const uint32_t slot = (g & 1) * 2 + jh, pslot = db ? slot : slot ^ 2, pph = ((g - (db ? 2u : 1u)) >> 1) & 1;
        const bool wait_v = kb >= (db ? 2 : 1);
        if (!warp_active) {
#pragma unroll
          for (int sb = 0; sb < kSub; ++sb) {
            mbar_wait(&raw_full[rs], rph);
            if (lane == 0) mbar_arrive(&raw_empty[rs]);
            if (++rs == kRawStages) { rs = 0; rph ^= 1; }
          }
          if (wait_v) mbar_wait(&done[pslot], pph);
          if (lane == 0) arrive_leader(&full[slot]);
          ++g;
          continue;
        }
#pragma unroll
        for (int sb = 0; sb < kSub; ++sb) {  // H16: two 8-channel raw stages fill one 16-channel V stage
          mbar_wait(&raw_full[rs], rph);
          float d[4][3][4];  // patch columns c + jh, c = 0..2
          if (tvalid) {
            const uint32_t a = raw_base + rs * kRawStride;
#pragma unroll
            for (int dy = 0; dy < 4; ++dy)
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const float4 v = ld_shared_v4(a + (dy & 1) * 2 * kPlaneBytes + poff[dy >> 1][c]);
                d[dy][c][0] = v.x, d[dy][c][1] = v.y, d[dy][c][2] = v.z, d[dy][c][3] = v.w;
              }
          } else {
#pragma unroll
            for (int dy = 0; dy < 4; ++dy)
#pragma unroll
              for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int e = 0; e < 4; ++e) d[dy][c][e] = 0.f;
          }
          // column pass t = B^T d, in place over dy
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float d0 = d[0][c][e], d1 = d[1][c][e], d2 = d[2][c][e], d3 = d[3][c][e];
              d[0][c][e] = d0 - d2;
              d[1][c][e] = d1 + d2;
              d[2][c][e] = d2 - d1;
              d[3][c][e] = d1 - d3;
            }
          __syncwarp();
          if (lane == 0) mbar_arrive(&raw_empty[rs]);  // the raw stage is in registers now
          if (++rs == kRawStages) { rs = 0; rph ^= 1; }

          // row pass for this half, rounded to the operand type and stored into TMEM (tf32: one column per channel;
          // 16-bit: one column per channel pair, raw stage sb fills columns 4*sb..)
          if (sb == 0) {
            if (wait_v) mbar_wait(&done[pslot], pph);  // the MMAs that last read this V half have completed
            tc_fence_after();
          }
          const uint32_t vcol = tmem_base + lane_base + v_col0 + (db ? (g & 1) * 128 : 0u) + jh * 64 +
                                (uint32_t)(H16 ? sb * 4 + cq * 2 : cq * 4);
          if (jh == 0) ffw_row_pass<0, H16>(d, vcol, fp16);  // warp-uniform
          else ffw_row_pass<1, H16>(d, vcol, fp16);
          if (sb == kSub - 1) {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader(&full[slot]);
          }
        }  // sb
        ++g;
      }

      // ---- epilogue: the accumulators ARE the output pixels; BN, ReLU, staged per (32 couts, output row a), full runs
      // per pixel. The four warps of a TMEM lane quadrant take 8 couts of the chunk each and share barrier 1 + quad.
      const int W = out_padded ? 16 : 14;
      const int o = out_padded ? 1 : 0;
      const int pix0 = tvalid ? ((n * W + 2 * ty + o) * W + 2 * tx + o) : -1;  // first output pixel of this tile
      if (esub == 0) pixtab[row] = pix0;
      const int n_chunks = kn / kEW;
      const int qrows = min(32, valid_rows - quad * 32);  // real tiles among this quadrant's rows (<= 0: none)
      const int tid128 = esub * 32 + lane;

      mbar_wait(acc_full, aph);
      aph ^= 1;
      tc_fence_after();
      // SPLIT: this CTA finishes tile rows [64 * crank, +64) = TMEM lane quadrants 2 * crank, 2 * crank + 1
      const bool owner = !SPLIT || (uint32_t)(quad >> 1) == crank;
      // inbox (in the OWNER's shared memory, from offset 0): [chunk (ec, a)][64 rows][2 pixels b][32 couts] fp32, the
      // 16-byte pieces of a row XOR-swizzled with the row so that neighbouring lanes do not hit one bank group
      const uint32_t inbox_row = (uint32_t)(row & 63);
      if constexpr (SPLIT) {
        if (threadIdx.x == 0) mbar_arrive_remote_plain(may_push, crank ^ 1u);  // my main loop is done: push
        if (!owner) {
          mbar_wait(may_push, 0);  // the owner's pipeline buffers are free
          uint32_t remote_base, remote_bar;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_base) : "r"(smem_u32(smem)), "r"(crank ^ 1u));
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_bar) : "r"(smem_u32(inbox_full)), "r"(crank ^ 1u));
          if (warp_active) {
            const int n_chunks_p = kn / kEW;
#pragma unroll 1
            for (int ec = 0; ec < n_chunks_p; ++ec)
#pragma unroll
              for (int a = 0; a < 2; ++a) {
                const uint32_t taddr = tmem_base + lane_base + (uint32_t)(2 * a) * acc_stride + (uint32_t)(ec * kEW + esub * 8);
                float z[2][8];
                tmem_ld_x8(taddr, z[0]);
                tmem_ld_x8(taddr + acc_stride, z[1]);
                tmem_ld_wait();
                const uint32_t dst = remote_base + ((uint32_t)(ec * 2 + a) * 64 + inbox_row) * 256;
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                  for (int q4 = 0; q4 < 2; ++q4) {
                    const uint32_t pos = (uint32_t)(b * 8 + esub * 2 + q4) ^ (inbox_row & 7);
                    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + pos * 16),
                                 "f"(z[b][4 * q4]), "f"(z[b][4 * q4 + 1]), "f"(z[b][4 * q4 + 2]), "f"(z[b][4 * q4 + 3])
                                 : "memory");
                  }
              }
          }
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
          continue;  // one item per cluster: nothing else to do for this warp
        }
        mbar_wait_cluster(inbox_full, 0);  // the other CTA's partial sums for my rows have landed
      }
      if (warp_active) {
#pragma unroll 1
        for (int ec = 0; ec < n_chunks; ++ec) {
          const int cl = esub * 8;       // cout inside the chunk
          const int c0 = ec * kEW + cl;  // cout inside the slice
          float sc[8], sh[8];
#pragma unroll
          for (int q4 = 0; q4 < 2; ++q4) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale + c0s + c0 + 4 * q4));
            const float4 h4 = __ldg(reinterpret_cast<const float4*>(shift + c0s + c0 + 4 * q4));
            sc[4 * q4] = s4.x, sc[4 * q4 + 1] = s4.y, sc[4 * q4 + 2] = s4.z, sc[4 * q4 + 3] = s4.w;
            sh[4 * q4] = h4.x, sh[4 * q4 + 1] = h4.y, sh[4 * q4 + 2] = h4.z, sh[4 * q4 + 3] = h4.w;
          }
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const uint32_t taddr = tmem_base + lane_base + (uint32_t)(2 * a) * acc_stride + (uint32_t)c0;
            float z[2][8];  // z[b][e] = Y[a][b]
            tmem_ld_x8(taddr, z[0]);
            tmem_ld_x8(taddr + acc_stride, z[1]);
            tmem_ld_wait();
            if constexpr (SPLIT) {  // add the other CTA's partial sums (the other half of the channel loop)
              const uint32_t src = smem_u32(smem) + ((uint32_t)(ec * 2 + a) * 64 + inbox_row) * 256;
#pragma unroll
              for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int q4 = 0; q4 < 2; ++q4) {
                  const uint32_t pos = (uint32_t)(b * 8 + esub * 2 + q4) ^ (inbox_row & 7);
                  const float4 v = ld_shared_v4(src + pos * 16);
                  z[b][4 * q4] += v.x, z[b][4 * q4 + 1] += v.y, z[b][4 * q4 + 2] += v.z, z[b][4 * q4 + 3] += v.w;
                }
            }
            if (ec == n_chunks - 1 && a == 1) {  // this warp has read its last accumulator columns
              tc_fence_before();
              __syncwarp();
              if (lane == 0) arrive_leader(acc_empty);
            }
            const uint32_t sdst = stg_base + (uint32_t)row * kStgRow + (uint32_t)cl * 4;
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
              for (int q4 = 0; q4 < 2; ++q4) {
                float ov[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  ov[e] = fmaf(sc[4 * q4 + e], z[b][4 * q4 + e], sh[4 * q4 + e]);
                  if (relu) ov[e] = fmaxf(ov[e], 0.f);
                }
                st_shared_v4(sdst + b * (4 * kEW) + 16 * q4, ov[0], ov[1], ov[2], ov[3]);
              }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");  // this quadrant's staging rows (+ pixtab) complete
            {
              const int units = qrows * 16;  // (tile, pixel b, 16-byte chunk)
              float* ybase = y + c0s + ec * kEW + (size_t)a * W * K;
              for (int u = tid128; u < units; u += 128) {
                const int tile = quad * 32 + (u >> 4);
                const int b = (u >> 3) & 1;
                const int ch = u & 7;
                const float4 v = ld_shared_v4(stg_base + (uint32_t)tile * kStgRow + (uint32_t)(b * 128 + ch * 16));
                st_out_v4(ybase + (size_t)(pixtab[tile] + b) * K + ch * 4, v, mc);
              }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");  // staging rows free again
          }
        }
        if (out_padded && tvalid && (ty == 0 || ty == 6 || tx == 0 || tx == 6)) {
          // zero border of the reference's 16x16 frame (Kernel128_winograd.cu:163,243): edge tiles own their share
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const size_t rstride = (size_t)W * K;
          const int ncc = kn / 4;
          float* p = y + (size_t)pix0 * K + c0s + esub * ncc;
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
      } else {
        if (lane == 0) arrive_leader(acc_empty);
      }
    }
  }

  tc_fence_before();
  if constexpr (CG2 || SPLIT) cluster_sync_all();  // the peer's shared memory, TMEM and barriers stay alive until both are done
  else __syncthreads();
  if (warp == kMmaWarp) {
    if constexpr (CG2) tmem_dealloc_cg2<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side

template <bool H16, bool CG2, bool SPLIT>
static int launch_ffw(const CUtensorMap& tmap, const float* u_img, const float* scale, const float* shift, float* y,
                      int n_img, int C, int K, int relu, int out_padded, int mv, int grid, cudaStream_t stream,
                      int fp16, int narrow) {
  static unsigned long long configured = 0;  // per device: the attribute is a property of the function on ONE device
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long dev_bit_ = 1ull << (dev_ & 63);
  if (!(configured & dev_bit_)) {
    if (cudaFuncSetAttribute(wino3x3_ffw_kernel<H16, CG2, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff::kTotal) !=
        cudaSuccess)
      return WG_ERR_CUDA;
    configured |= dev_bit_;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(ffw::kThreads);
  cfg.dynamicSmemBytes = ff::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG2 || SPLIT) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, wino3x3_ffw_kernel<H16, CG2, SPLIT>, tmap, u_img, scale, shift, y, n_img, C, K, relu,
                                     out_padded, mv, fp16, narrow);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

// split: clusters of 2 on one item each, grid = 2 * #items (the caller guarantees that this fits one wave and that the
// layer's stage count is even)
int wino_ffw_launch(const CUtensorMap& tmap, const float* u_img, const float* scale, const float* shift, float* y,
                    int n_img, int C, int K, int op16, int cg2, int split, int narrow, int relu, int out_padded, int mv,
                    int grid, cudaStream_t stream) {
#define WG_FFW(H16_, CG2_, SPLIT_)                                                                                   \
  return launch_ffw<H16_, CG2_, SPLIT_>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded, mv, grid, stream, \
                                        op16 == 2, narrow)
  if (split) {
    if (op16) WG_FFW(true, false, true);
    WG_FFW(false, false, true);
  }
  if constexpr (kDev) {  // CTA pairs: experiment, developer build only
    if (cg2) {
      if (op16) WG_FFW(true, true, false);
      WG_FFW(false, true, false);
    }
  }
  if (op16) WG_FFW(true, false, false);
  WG_FFW(false, false, false);
#undef WG_FFW
}

}  // namespace wg
