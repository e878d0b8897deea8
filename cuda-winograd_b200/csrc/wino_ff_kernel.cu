// Fused 3x3 conv (Winograd F(2x2,3x3)) + folded BatchNorm + ReLU for sm_100a -- throughput kernel with the WHOLE
// inverse transform folded into the tensor core ("full fold"): 4 accumulators x 96 couts.
//
// Replaces kernel_{128,256}_winograd_BtdB -> kernel_*_OuterProduct_* -> kernel_*_winograd_AtIA
// (/root/reference/Kernel128_winograd.cu:28-213, Kernel256_winograd.cu:27-218). Same skeleton as wino3x3_tm_kernel
// (wino_tm_kernel.cu: persistent, 128-tile M-blocks x cout slices, 8-channel stages, V = B^T d B written by the
// transform warps straight into TENSOR MEMORY and read from there as the MMA's A operand), but
//
//   * the accumulators are the four output pixels of a tile themselves: Y[a][b] = sum_{i,j} A^T[a][i] A^T[b][j] M[i][j],
//     so accumulator (a,b) collects the 9 Winograd points with A^T[a][i] != 0 and A^T[b][j] != 0, the sign
//     A^T[a][i] * A^T[b][j] going through the instruction descriptor's negate-A bit. 36 MMAs per stage instead of 24
//     (half fold, 8 accumulators) or 16 (no fold, 16 accumulators) -- the MMA FLOPs of a direct convolution -- but
//     only 4 accumulators: 4 x 96 couts = 384 TMEM columns + the two 64-column V halves = 512. A 96-wide slice means
//     the raw tile is fetched and transformed 3 times per M-block for K = 256 (96 + 96 + 64) instead of 6 times
//     (4 x 48 + 2 x 32): the kernel is bound by the transform warps and the TMA feed, not by the tensor core, so
//     halving the transform passes wins even though the MMA work grows by 1.5x;
//   * the epilogue has no arithmetic left but BN + ReLU: it drains the accumulators in chunks of (32 couts, output row a)
//     through a [tile][2 pixels][32 couts] staging area and writes full 128-byte runs per output pixel; the two warps
//     that own the same 32 TMEM lanes stage and write out their 32 tiles on their own (64-thread named barriers).
//
// Numerically this is still F(2x2,3x3): the operands of every product are the TF32-rounded V = B^T d B and
// U = G g G^T; only the order of the fp32 additions differs from the un-folded form.
#include "wino_ff_common.cuh"
#include "wg_internal.h"

#include <cuda.h>
#include <stdlib.h>

namespace wg {

// H16: 16-bit operands (bf16, or fp16 with `fp16` set): V is stored in TMEM as packed pairs (column c = channels 2c
// and 2c+1), tcgen05.mma kind::f16 with K = 16, so a V stage covers 16 channels = TWO 8-channel raw stages.
// DBG: developer build with ablation switches (WG_FF_DEBUG=<bits>: 1 no MMAs, 2 no patch loads / transform / TMEM
// stores, 4 no filter loads, 8 no raw-tile loads, 16 no output stores); results are garbage, only the time is of interest.
// The product instantiation (DBG = false) contains none of it.
// CG2: CTA pairs (thread-block clusters of 2, tcgen05 cta_group::2). The two CTAs take neighbouring M-blocks and the
// same cout slice; each loads, transforms and drains its own 128 tiles, but the MMAs are issued by the leader CTA
// (cluster rank 0) with M = 256: each CTA supplies HALF of the couts of every filter chunk from its own shared memory,
// so the tensor core's B-operand reads and the filter's TMA traffic per SM halve -- the shared-memory data path is what
// the transform warps' patch loads, the MMAs' operand reads and the TMA writes compete for. Cross-CTA hand-offs: the
// peer's transform warps arrive on the LEADER's full[] barriers (mbarrier.arrive.release.cluster), the peer's otherwise
// idle MMA warp relays "my half of the filter chunk has landed", the leader's tcgen05.commit multicasts done[] /
// acc_full to both CTAs.
// ALT (all cout slices <= 64 wide, i.e. two V stages in TMEM; parity-plane raw layout; no CTA pairs): the eight transform
// warps work as TWO GROUPS OF FOUR ON ALTERNATE V STAGES -- group w>>2 owns the stages with that parity and transforms
// both 4-channel halves of its stage one after the other. A transform warp's stage is a chain of fixed latencies (wait
// for the raw tile, 16 LDS.128, column / row pass, wait for the V buffer, 16 tcgen05.st, tcgen05.wait::st, arrive) of
// ~1.9 k clk against ~1.3 k clk of MMAs per stage (N = 64); with the groups half a period apart one group's loads run
// under the other's stores and hand-offs, and every fixed cost is paid once per TWO stages.
// GEN: runtime map geometry (wg_conv3x3_create_hw). The reference's 14x14 / 16x16 geometry is compiled in (!GEN): tile
// and pixel indices then divide by constants and the edge-tile masks vanish (measured: 46.3 vs 49.4 us at 128->128).
template <bool H16, bool DBG, bool P9, bool CG2, bool ALT = false, bool GEN = false>
__global__ void __launch_bounds__(32 * (ff::kWorkerWarps + 2), 1)
wino3x3_ff_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ u_img,
                  const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ y, int n_img,
                  int C, int K, int relu, int out_padded, int mv, int fp16, int dbg, const ff::Geo geo_arg, int narrow) {
  using namespace ff;
  const Geo geo = GEN ? geo_arg : Geo{14, 14, 16, 16, 7, 7, 49, 9, 8, 24, 128, kRawBytesP9};
  static_assert(!ALT || (P9 && !CG2), "alternating transform groups: parity planes, single CTAs");
  constexpr int kGroupWarps = ALT ? kWorkerWarps / 2 : kWorkerWarps;  // transform warps that share a V stage
  const bool no_mma = DBG && (dbg & 1), no_xf = DBG && (dbg & 2), no_u = DBG && (dbg & 4), no_raw = DBG && (dbg & 8),
             no_out = DBG && (dbg & 16);
  // dbg & 32 (with 16): CTA 0 records clock64() at the hand-off points of stages 8..15 of its first item and dumps
  // them into y: [stage][8] from worker warp 0 at y, [stage][8] from the MMA thread at y + 1024 B
  // dbg & 64 also records globaltimer_ns() per CTA: [12] kernel entry, [13] producer past griddepcontrol.wait,
  // [14] first raw stage landed (worker warp 0), [15] CTA done
  long long* ts_g = reinterpret_cast<long long*>(y) + 4096 + 16 * blockIdx.x;
  if (DBG && (dbg & 64) && threadIdx.x == 0) ts_g[12] = (long long)globaltimer_ns();
  const bool ts_on = DBG && (dbg & 32) && blockIdx.x == 0;
  long long* ts_w = reinterpret_cast<long long*>(y);
  long long* ts_m = reinterpret_cast<long long*>(y) + 128;
#define WG_TS(buf, slot)                                                             \
  if (DBG && ts_on && item == item0 && kb >= 8 && kb < 16) (buf)[(kb - 8) * 8 + (slot)] = clock64()
  constexpr int kSub = H16 ? 2 : 1;  // 8-channel raw stages per V stage
  const bool mc = (out_padded & 2) != 0;  // y is an NVLS multicast address: stores go out as multimem.st
  out_padded &= 1;
  pdl_launch_dependents();  // the next launch in the stream may start its prologue (it waits before touching x / y)
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = raw_full + kRawStages;
  // One "full" / "done" barrier pair per filter-ring slot = (stage parity, V half): full[s] collects the 8 transform
  // warps' "V half stored" AND the filter chunk's TMA bytes, so the MMA thread waits once per half; done[s] is the single
  // tcgen05.commit of that half's 18 MMAs and releases both the filter chunk (producer) and the V half (transform warps).
  static_assert(kUBufs == 4, "ring slot = 2 * (stage & 1) + half");
  uint64_t* full = raw_empty + kRawStages;  // [4]
  uint64_t* done = full + 4;                // [4]
  uint64_t* acc_full = done + 4;
  uint64_t* acc_empty = acc_full + 1;
  uint64_t* u_land = acc_empty + 1;  // [4] CG2, peer CTA only: its half of a filter chunk has landed
  const uint32_t crank = CG2 ? cluster_ctarank() : 0u;
  const bool peer = CG2 && crank != 0;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  int* pixtab = reinterpret_cast<int*>(smem + kOffPix);
  int* masktab = pixtab + 128;  // per tile: bit 2a+b set = output pixel (a, b) of the tile lies inside the H x W map

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < kRawStages; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], kGroupWarps);
    }
    for (int i = 0; i < 4; ++i) {
      // CG2 (leader): both CTAs' transform warps + own TMA bytes + the peer's relay
      mbar_init(&full[i], CG2 ? 2 * kWorkerWarps + 2 : kGroupWarps + 1);
      mbar_init(&done[i], 1);
      mbar_init(&u_land[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, CG2 ? 2 * kWorkerWarps : kWorkerWarps);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    if constexpr (CG2) tmem_alloc_cg2<512>(tmem_ptr);
    else tmem_alloc<512>(tmem_ptr);
  }
  tc_fence_before();
  if constexpr (CG2) cluster_sync_all();  // the peer's barriers exist before anybody arrives on them remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_kb = C / (8 * kSub);  // V stages (8 or 16 channels each)
  const int n_sl = n_slices(K, narrow);
  const int total_tiles = n_img * geo.TT;
  const int n_mblocks = (total_tiles + mv - 1) / mv;  // mv = tiles per M-block (<= 128), chosen by the host
  // item = (M-block, cout slice), slices of an M-block adjacent; CG2: (pair of M-blocks, slice), one item per cluster
  const int n_items = (CG2 ? (n_mblocks + 1) / 2 : n_mblocks) * n_sl;
  const int item0 = CG2 ? blockIdx.x / 2 : blockIdx.x, item_step = CG2 ? gridDim.x / 2 : gridDim.x;
  const uint32_t u_bytes_per_kn = CG2 ? 128u : 256u;  // bytes of a filter chunk this CTA loads, per cout of the slice

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t rs = 0, rph = 0, us = 0, uph = 0;
      int u_primed = 0;
      if (item0 < n_items) {
        // the filter does not depend on the previous kernel in the stream: request the first stage's U chunks before
        // waiting for that kernel (programmatic dependent launch), the activations after
        const Slice sl = slice(K, item0 % n_sl, narrow);
        const int kn = sl.kn, c0 = sl.c0;
        const uint8_t* u_src = reinterpret_cast<const uint8_t*>(u_img) + (size_t)n_kb * 512 * c0;
        for (int h = 0; h < 2; ++h) {
          uint64_t* ubar = peer ? &u_land[us] : &full[us];
          mbar_arrive_expect_tx(ubar, u_bytes_per_kn * kn);
          tma_bulk_g2s(smem + kOffU + us * kUChunkMax, u_src + ((size_t)h * 256 + crank * 128) * kn, u_bytes_per_kn * kn,
                       ubar);
          ++us;
        }
        u_primed = 1;
      }
      pdl_wait();
      if (DBG && (dbg & 64)) ts_g[13] = (long long)globaltimer_ns();
      for (int item = item0; item < n_items; item += item_step) {
        const Slice sl = slice(K, item % n_sl, narrow);
        const int kn = sl.kn, c0 = sl.c0;
        const int mb = CG2 ? (item / n_sl) * 2 + (int)crank : item / n_sl;
        const int t0 = mb * mv;
        const int ny0 = (t0 / geo.TT) * geo.Hf + 2 * ((t0 % geo.TT) / geo.TX);  // first frame row (n*Hf + y), even
        const uint8_t* u_src = reinterpret_cast<const uint8_t*>(u_img) + (size_t)n_kb * 512 * c0;
        for (int kb = 0; kb < n_kb; ++kb) {
#pragma unroll
          for (int sb = 0; sb < kSub; ++sb) {
            mbar_wait(&raw_empty[rs], rph ^ 1);
            if (no_raw) {
              mbar_arrive(&raw_full[rs]);
            } else if constexpr (P9) {
              mbar_arrive_expect_tx(&raw_full[rs], geo.raw_bytes);
#pragma unroll
              for (int q = 0; q < 4; ++q)  // plane q = (y parity q>>1, x parity q&1); x/2 starts at -1 (zero-filled)
                tma_tensor_5d_g2s(smem + kOffRaw + rs * kRawStride + q * kPlaneBytes, &tmap_x, (kb * kSub + sb) * 8, -1,
                                  q & 1, q >> 1, ny0 >> 1, &raw_full[rs]);
            } else {
              mbar_arrive_expect_tx(&raw_full[rs], kRawBytes);
              tma_tensor_4d_g2s(smem + kOffRaw + rs * kRawStride, &tmap_x, (kb * kSub + sb) * 8, 0, 0, ny0,
                                &raw_full[rs]);
            }
            if (++rs == kRawStages) { rs = 0; rph ^= 1; }
          }
          if (u_primed) {  // already requested above
            u_primed = 0;
            continue;
          }
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&done[us], uph ^ 1);
            uint64_t* ubar = peer ? &u_land[us] : &full[us];
            if (no_u) {
              mbar_arrive(ubar);
            } else {
              mbar_arrive_expect_tx(ubar, u_bytes_per_kn * kn);
              tma_bulk_g2s(smem + kOffU + us * kUChunkMax, u_src + (((size_t)kb * 2 + h) * 256 + crank * 128) * kn,
                           u_bytes_per_kn * kn, ubar);
            }
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
      uint32_t g = 0, us = 0, uph = 0, aph = 0;  // g = V stages issued (V phase = g & 1)
      // (Probing the next half's barrier with mbarrier.test_wait while 6 of the 18 MMAs were still to be issued, so
      //  that the ~90 clk of an already-complete try_wait stay off the issue path, was measured: no gain.)
      for (int item = item0; item < n_items; item += item_step) {
        const uint32_t kn = (uint32_t)slice(K, item % n_sl, narrow).kn;
        const uint32_t fmt = H16 ? (fp16 ? kFmtF16 : kFmtBF16) : kFmtTF32;
        const uint32_t idesc_pos = make_idesc(fmt, CG2 ? 256 : 128, kn);
        const uint32_t idesc_neg = make_idesc(fmt, CG2 ? 256 : 128, kn, 1);  // D += (-A) * B
        // CG2: this CTA's shared memory holds kn/2 couts of every point
        const uint32_t u_per_point = (CG2 ? 1 : 2) * kn * 16, u_lbo = (CG2 ? kn / 2 : kn) * 16;
        // slices of 64 or fewer couts leave room for a SECOND V stage in TMEM (4 x 64 accumulator columns + 2 x 128):
        // the transform warps may then run a whole stage ahead of the MMAs. Accumulator p at p * acc_stride, V stage
        // (g & 1) at v_col0 + 128 * (g & 1).
        const bool db = !CG2 && kn <= 64;
        const uint32_t acc_stride = db ? 64u : kAccStride, v_col0 = db ? 256u : kVCol0;
        mbar_wait(acc_empty, aph ^ 1);  // epilogue of the previous item has drained TMEM
        tc_fence_after();
        for (int kb = 0; kb < n_kb; ++kb) {
          // bit p set = accumulator p has been written in this item (first MMA into it overwrites)
          uint32_t written = kb > 0 ? 0xFu : 0u;
#pragma unroll
          for (int jh = 0; jh < 2; ++jh) {
            WG_TS(ts_m, jh * 4 + 0);
            // filter chunk landed and V half stored by all transform warps (of both CTAs)
            mbar_wait(&full[us], uph);
            tc_fence_after();
            WG_TS(ts_m, jh * 4 + 1);
            const uint32_t ua = u_base + us * kUChunkMax;
            const uint32_t va = tmem_base + v_col0 + (db ? (g & 1) * 128 : 0u) + jh * 64;
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
                    if (!no_mma)
                      ff_umma<H16, CG2>(tmem_base + p * acc_stride, a_tm, b_desc, sa * sb > 0 ? idesc_pos : idesc_neg,
                                        (written >> p) & 1u);
                    written |= 1u << p;
                  }
                }
              }
            }
            WG_TS(ts_m, jh * 4 + 2);
            if constexpr (CG2) umma_commit_mcast_cg2(&done[us], 3);  // frees the filter chunk and this V half, both CTAs
            else umma_commit(&done[us]);
            if (++us == kUBufs) { us = 0; uph ^= 1; }
          }
          ++g;
        }
        if constexpr (CG2) umma_commit_mcast_cg2(acc_full, 3);
        else umma_commit(acc_full);
        aph ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ transform + epilogue warps
    // thread = (MMA row = TMEM lane, channel half cq): warp w owns TMEM lanes 32*(w&3)..+31
    const int quad = warp & 3;
    const int cq = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t raw_base = smem_u32(smem + kOffRaw);
    const uint32_t stg_base = smem_u32(smem + kOffStg);

    uint32_t rs = 0, rph = 0, g = 0, aph = 0;  // g = V stages transformed (same counting as the MMA thread)
    // CG2: full[] / acc_empty live in the leader CTA; done[] / acc_full are signalled by the leader's multicast commit
    auto arrive_leader = [&](uint64_t* bar) {
      if (peer) mbar_arrive_remote_plain(bar, 0);
      else mbar_arrive(bar);
    };
    auto wait_done = [&](uint64_t* bar, uint32_t parity) {
      mbar_wait(bar, parity);  // payload is TMEM state, ordered by the tcgen05 fences around the wait
    };
    // dbg & 64 (with 16): every CTA records clock64() at item start / main loop end / epilogue end of its first four
    // items: long long y[4096 + 16 * blockIdx.x + 4 * item_index + {0, 1, 2}], slot 3 = item number
    long long* ts_i = reinterpret_cast<long long*>(y) + 4096 + 16 * blockIdx.x;
    const bool ts_item = DBG && (dbg & 64) && warp == 0 && lane == 0;
    int item_idx = 0;
    for (int item = item0; item < n_items; item += item_step, ++item_idx) {
      if (DBG && ts_item && item_idx < 4) ts_i[4 * item_idx] = clock64(), ts_i[4 * item_idx + 3] = item;
      const Slice sl = slice(K, item % n_sl, narrow);
      const int kn = sl.kn, c0s = sl.c0;
      const int mb = CG2 ? (item / n_sl) * 2 + (int)crank : item / n_sl;
      const int t0 = mb * mv;
      const int ny0 = (t0 / geo.TT) * geo.Hf + 2 * ((t0 % geo.TT) / geo.TX);
      const int T = t0 + row;
      const int valid_rows = min(mv, total_tiles - t0);  // rows of this M-block that hold real tiles
      const bool tvalid = row < valid_rows;
      const bool warp_active = quad * 32 < valid_rows;  // warp-uniform
      const bool db = !CG2 && kn <= 64;                 // two V stages in TMEM (see the MMA thread)
      const uint32_t acc_stride = db ? 64u : kAccStride, v_col0 = db ? 256u : kVCol0;
      const int n = T / geo.TT, t = T % geo.TT, ty = t / geo.TX;
      const int tx = P9 ? geo.TX - 1 - t % geo.TX : t % geo.TX;  // P9: tiles run right-to-left inside a tile row (see kPlaneBytes)
      const uint32_t raw_off = tvalid ? (uint32_t)((n * 16 + 2 * ty - ny0) * 512 + tx * 32) : 0u;  // !P9: 14x14 only
      // SWIZZLE_32B: the 16-byte half of a pixel's 32 bytes is XORed with bit 2 of its x/2 index (address bit 7)
      const uint32_t h0 = (uint32_t)((cq ^ ((tx >> 2) & 1)) * 16);        // pixels with x/2 = tx
      const uint32_t h1 = (uint32_t)((cq ^ (((tx + 1) >> 2) & 1)) * 16);  // pixels with x/2 = tx + 1
      // P9: byte offset inside a plane of the pixel (dy/2, dx/2) of this tile's patch, swizzled half included
      uint32_t p9off[2][2];
      if constexpr (P9) {
        const uint32_t s0 = tvalid ? (uint32_t)(((n * geo.Hf + 2 * ty - ny0) >> 1) * geo.SP + tx + 1) : 1u;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const uint32_t sl = s0 + (uint32_t)geo.SP * a + b;
            p9off[a][b] = sl * 32 + (uint32_t)((cq ^ ((sl >> 2) & 1)) * 16);
          }
      }

      if constexpr (ALT) {
        // ---- alternating groups: group `cq` (= warp >> 2) owns the V stages g with (g & 1) == cq, both channel halves
        for (int kb = 0; kb < n_kb; ++kb) {
          if ((int)(g & 1) != cq) {  // the other group's stage: only keep the ring position in step
#pragma unroll
            for (int sb = 0; sb < kSub; ++sb)
              if (++rs == kRawStages) { rs = 0; rph ^= 1; }
            ++g;
            continue;
          }
          const uint32_t slot = (g & 1) * 2, pph = ((g - 2u) >> 1) & 1;
          const bool wait_v = kb >= 2;  // the MMAs of stage g - 2 (this group's previous stage) read this V buffer
#pragma unroll
          for (int sb = 0; sb < kSub; ++sb) {
            mbar_wait(&raw_full[rs], rph);
            if (warp_active) {
#pragma unroll
              for (int ch = 0; ch < 2; ++ch) {  // the two 4-channel halves of the 8-channel raw stage, one after the other
                f2_t d[4][4][2];
                if (tvalid) {
                  const uint32_t a = raw_base + rs * kRawStride;
#pragma unroll
                  for (int dy = 0; dy < 4; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 4; ++dx) {
                      // p9off holds the address for channel half cq; the other half is the other 16 bytes of the slot
                      const uint32_t off = p9off[dy >> 1][dx >> 1] ^ (ch != cq ? 16u : 0u);
                      ld_shared_f2x2(a + ((dy & 1) * 2 + (dx & 1)) * kPlaneBytes + off, d[dy][dx][0], d[dy][dx][1]);
                    }
                } else {
#pragma unroll
                  for (int dy = 0; dy < 4; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 4; ++dx) d[dy][dx][0] = d[dy][dx][1] = 0ull;
                }
#pragma unroll
                for (int dx = 0; dx < 4; ++dx)
#pragma unroll
                  for (int c = 0; c < 2; ++c) {
                    const f2_t d0 = d[0][dx][c], d1 = d[1][dx][c], d2 = d[2][dx][c], d3 = d[3][dx][c];
                    d[0][dx][c] = f2_sub(d0, d2);
                    d[1][dx][c] = f2_add(d1, d2);
                    d[2][dx][c] = f2_sub(d2, d1);
                    d[3][dx][c] = f2_sub(d1, d3);
                  }
                if (ch == 1) {  // both halves of the raw stage are in registers / consumed: hand it back
                  __syncwarp();
                  if (lane == 0) mbar_arrive(&raw_empty[rs]);
                }
                if (sb == 0 && ch == 0) {
                  if (wait_v) wait_done(&done[slot + 1], pph);  // later commit of stage g - 2: the whole V buffer is free
                  tc_fence_after();
                }
                const uint32_t vcol = tmem_base + lane_base + 256u + (g & 1) * 128 + (uint32_t)(H16 ? sb * 4 + ch * 2 : ch * 4);
#pragma unroll
                for (int jh = 0; jh < 2; ++jh)
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    f2_t v0[2], v1[2];  // points (i, 2jh) and (i, 2jh+1)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                      const f2_t a0 = d[i][0][c], a1 = d[i][1][c], a2 = d[i][2][c], a3 = d[i][3][c];
                      v0[c] = jh == 0 ? f2_sub(a0, a2) : f2_sub(a2, a1);
                      v1[c] = jh == 0 ? f2_add(a1, a2) : f2_sub(a1, a3);
                      if constexpr (!H16) v0[c] = f2_tf32(v0[c]), v1[c] = f2_tf32(v1[c]);
                    }
                    const uint32_t dst = vcol + jh * 64 + (i * 2) * 8;
                    if constexpr (H16) {
                      tmem_st_x2(dst, ff_pack16(f2_lo(v0[0]), f2_hi(v0[0]), fp16), ff_pack16(f2_lo(v0[1]), f2_hi(v0[1]), fp16));
                      tmem_st_x2(dst + 8, ff_pack16(f2_lo(v1[0]), f2_hi(v1[0]), fp16),
                                 ff_pack16(f2_lo(v1[1]), f2_hi(v1[1]), fp16));
                    } else {
                      tmem_st_x4(dst, f2_lo(v0[0]), f2_hi(v0[0]), f2_lo(v0[1]), f2_hi(v0[1]));
                      tmem_st_x4(dst + 8, f2_lo(v1[0]), f2_hi(v1[0]), f2_lo(v1[1]), f2_hi(v1[1]));
                    }
                  }
              }
            } else {  // no tile in this warp's 32 rows: only the hand-offs
              __syncwarp();
              if (lane == 0) mbar_arrive(&raw_empty[rs]);
              if (sb == 0 && wait_v) wait_done(&done[slot + 1], pph);
            }
            if (++rs == kRawStages) { rs = 0; rph ^= 1; }
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&full[slot]);
            mbar_arrive(&full[slot + 1]);
          }
          ++g;
        }
      } else
      for (int kb = 0; kb < n_kb; ++kb) {
        // this stage's ring slots are slot, slot + 1; the previous stage's (whose MMAs must have completed before a V
        // half is overwritten) pslot, pslot + 1, completing for the ((g - 1) >> 1)-th time
        // One V stage (slices wider than 64): the MMAs of stage g - 1 must have completed. Two V stages (db): those of
        // stage g - 2, which used this stage's own ring slots. The first stage(s) of an item wait for nothing: this
        // thread has passed the previous item's acc_full, i.e. every earlier MMA has completed.
        const uint32_t slot = (g & 1) * 2, pslot = db ? slot : slot ^ 2, pph = ((g - (db ? 2u : 1u)) >> 1) & 1;
        const bool wait_v = kb >= (db ? 2 : 1);
        if (!warp_active) {
          // nothing to transform: release the raw stage(s) and report "V ready" in step with the other warps
#pragma unroll
          for (int sb = 0; sb < kSub; ++sb) {
            mbar_wait(&raw_full[rs], rph);
            if (lane == 0) mbar_arrive(&raw_empty[rs]);
            if (++rs == kRawStages) { rs = 0; rph ^= 1; }
          }
#pragma unroll
          for (int jh = 0; jh < 2; ++jh) {
            if (wait_v) wait_done(&done[pslot + jh], pph);
            if (lane == 0) arrive_leader(&full[slot + jh]);
          }
          ++g;
          continue;
        }
#pragma unroll
        for (int sb = 0; sb < kSub; ++sb) {  // H16: two 8-channel raw stages fill one 16-channel V stage
          if (warp == 0 && lane == 0 && sb == 0) { WG_TS(ts_w, 0); }
          mbar_wait(&raw_full[rs], rph);
          if (warp == 0 && lane == 0 && sb == 0) { WG_TS(ts_w, 1); }
          if (DBG && ts_item && item_idx == 0 && kb == 0 && sb == 0) ts_g[14] = (long long)globaltimer_ns();
          f2_t d[4][4][2];  // [dy][dx][channel pair]: packed fp32 pairs, all passes on the 2-wide fp32 pipe
          if (no_xf) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[rs]);
            if (++rs == kRawStages) { rs = 0; rph ^= 1; }
            if (sb == kSub - 1) {
#pragma unroll
              for (int jh = 0; jh < 2; ++jh) {
                if (wait_v) wait_done(&done[pslot + jh], pph);
                if (lane == 0) arrive_leader(&full[slot + jh]);
              }
            }
            continue;
          }
          if (tvalid) {
            const uint32_t a = raw_base + rs * kRawStride + (P9 ? 0u : raw_off);
#pragma unroll
            for (int dy = 0; dy < 4; ++dy)
#pragma unroll
              for (int dx = 0; dx < 4; ++dx) {
                const uint32_t ad = P9 ? a + ((dy & 1) * 2 + (dx & 1)) * kPlaneBytes + p9off[dy >> 1][dx >> 1]
                                       : a + dy * 512 + (dx & 1) * 256 + (dx >> 1) * 32 + ((dx >> 1) ? h1 : h0);
                ld_shared_f2x2(ad, d[dy][dx][0], d[dy][dx][1]);
              }
          } else {
#pragma unroll
            for (int dy = 0; dy < 4; ++dy)
#pragma unroll
              for (int dx = 0; dx < 4; ++dx) d[dy][dx][0] = d[dy][dx][1] = 0ull;
          }
          // column pass t = B^T d, in place over dy
#pragma unroll
          for (int dx = 0; dx < 4; ++dx)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const f2_t d0 = d[0][dx][c], d1 = d[1][dx][c], d2 = d[2][dx][c], d3 = d[3][dx][c];
              d[0][dx][c] = f2_sub(d0, d2);
              d[1][dx][c] = f2_add(d1, d2);
              d[2][dx][c] = f2_sub(d2, d1);
              d[3][dx][c] = f2_sub(d1, d3);
            }
          // the raw stage is in registers now: hand it back to the producer before the row pass
          __syncwarp();
          if (lane == 0) mbar_arrive(&raw_empty[rs]);
          if (++rs == kRawStages) { rs = 0; rph ^= 1; }

          // row pass V = t B by halves (half jh = points with j in {2jh, 2jh+1}), round to the operand type, store into
          // TMEM (tf32: one column per channel; 16-bit: one column per channel pair, raw stage sb fills columns 4*sb..)
          const uint32_t vcol = tmem_base + lane_base + v_col0 + (db ? (g & 1) * 128 : 0u) +
                                (uint32_t)(H16 ? sb * 4 + cq * 2 : cq * 4);
#pragma unroll
          for (int jh = 0; jh < 2; ++jh) {
            if (warp == 0 && lane == 0 && sb == kSub - 1) { WG_TS(ts_w, 2 + jh * 3); }
            if (sb == 0) {
              // the MMAs that last read this V half have completed. Two V stages (db): ONE wait per stage, on the later
              // of the two commits of stage g - 2 (commits complete in order, so the first half is free as well).
              if (wait_v && (!db || jh == 0)) wait_done(&done[pslot + (db ? 1 : jh)], pph);
              tc_fence_after();
            }
            if (warp == 0 && lane == 0 && sb == kSub - 1) { WG_TS(ts_w, 3 + jh * 3); }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              f2_t v0[2], v1[2];  // points (i, 2jh) and (i, 2jh+1)
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                const f2_t a0 = d[i][0][c], a1 = d[i][1][c], a2 = d[i][2][c], a3 = d[i][3][c];
                v0[c] = jh == 0 ? f2_sub(a0, a2) : f2_sub(a2, a1);
                v1[c] = jh == 0 ? f2_add(a1, a2) : f2_sub(a1, a3);
                if constexpr (!H16) v0[c] = f2_tf32(v0[c]), v1[c] = f2_tf32(v1[c]);
              }
              const uint32_t dst = vcol + jh * 64 + (i * 2) * 8;
              if constexpr (H16) {
                tmem_st_x2(dst, ff_pack16(f2_lo(v0[0]), f2_hi(v0[0]), fp16), ff_pack16(f2_lo(v0[1]), f2_hi(v0[1]), fp16));
                tmem_st_x2(dst + 8, ff_pack16(f2_lo(v1[0]), f2_hi(v1[0]), fp16), ff_pack16(f2_lo(v1[1]), f2_hi(v1[1]), fp16));
              } else {
                tmem_st_x4(dst, f2_lo(v0[0]), f2_hi(v0[0]), f2_lo(v0[1]), f2_hi(v0[1]));
                tmem_st_x4(dst + 8, f2_lo(v1[0]), f2_hi(v1[0]), f2_lo(v1[1]), f2_hi(v1[1]));
              }
            }
            // V half stored by this warp: tell the MMA thread. Two V stages (db): one tcgen05.wait::st / fence / warp sync
            // per STAGE, then both halves' arrivals (the MMA thread's second wait is then already satisfied).
            if (sb == kSub - 1 && (!db || jh == 1)) {
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (db) arrive_leader(&full[slot]);
                arrive_leader(&full[slot + jh]);
              }
              if (warp == 0 && lane == 0) { WG_TS(ts_w, 4 + jh * 3); }
            }
          }
        }  // sb
        ++g;
      }

      // ---- epilogue: the accumulators ARE the output pixels; BN, ReLU, staged per (32 couts, output row a), full runs
      // per pixel. Warps (quad, cq = 0/1) own the same 32 tiles: they share staging rows 32*quad..+31 and barrier 1+quad.
      const int W = out_padded ? geo.Wf : geo.W;   // row pitch / image height of the output (frame or dense map)
      const int Hout = out_padded ? geo.Hf : geo.H;
      const int o = out_padded ? 1 : 0;
      const int pix0 = tvalid ? ((n * Hout + 2 * ty + o) * W + 2 * tx + o) : -1;  // first output pixel of this tile
      if (cq == 0) {
        pixtab[row] = pix0;
        // odd H / W: the last tile row / column has one output row / column outside the map
        const int ra = (2 * ty + 1 < geo.H) ? 0xC : 0, cb = (2 * tx + 1 < geo.W) ? 0xA : 0;  // bits 2a+b
        masktab[row] = tvalid ? (0x1 | (cb & 0x2) | (ra & 0x4) | ((ra & cb) & 0x8)) : 0;
      }
      const int n_chunks = kn / kEW;
      const int qrows = min(32, valid_rows - quad * 32);  // real tiles among this quad's rows (<= 0: none)
      const int tid64 = cq * 32 + lane;

      wait_done(acc_full, aph);
      aph ^= 1;
      tc_fence_after();
      if (DBG && ts_item && item_idx < 4) ts_i[4 * item_idx + 1] = clock64();
      if (warp_active) {
#pragma unroll 1
        for (int ec = 0; ec < n_chunks; ++ec) {
          float sc[16], sh[16];  // this thread's 16 couts of the chunk
          const int cl = cq * 16;           // cout inside the chunk
          const int c0 = ec * kEW + cl;     // cout inside the slice
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale + c0s + c0 + 4 * q4));
            const float4 h4 = __ldg(reinterpret_cast<const float4*>(shift + c0s + c0 + 4 * q4));
            sc[4 * q4] = s4.x, sc[4 * q4 + 1] = s4.y, sc[4 * q4 + 2] = s4.z, sc[4 * q4 + 3] = s4.w;
            sh[4 * q4] = h4.x, sh[4 * q4 + 1] = h4.y, sh[4 * q4 + 2] = h4.z, sh[4 * q4 + 3] = h4.w;
          }
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const uint32_t taddr = tmem_base + lane_base + (uint32_t)(2 * a) * acc_stride + (uint32_t)c0;
            float z[2][16];  // z[b][e] = Y[a][b]
            tmem_ld_x16(taddr, z[0]);
            tmem_ld_x16(taddr + acc_stride, z[1]);
            tmem_ld_wait();
            if (ec == n_chunks - 1 && a == 1) {  // this warp has read its last accumulator columns
              tc_fence_before();
              __syncwarp();
              if (lane == 0) arrive_leader(acc_empty);
            }
            const uint32_t sdst = stg_base + (uint32_t)row * kStgRow + (uint32_t)cl * 4;
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                float ov[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  ov[e] = fmaf(sc[4 * q4 + e], z[b][4 * q4 + e], sh[4 * q4 + e]);
                  if (relu) ov[e] = fmaxf(ov[e], 0.f);
                }
                st_shared_v4(sdst + b * (4 * kEW) + 16 * q4, ov[0], ov[1], ov[2], ov[3]);
              }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");  // this quad's staging rows (+ pixtab) complete
            {
              const int units = qrows * 16;  // (tile, pixel b, 16-byte chunk)
              float* ybase = y + c0s + ec * kEW + (size_t)a * W * K;
              for (int u = tid64; u < units; u += 64) {
                const int tile = quad * 32 + (u >> 4);
                const int b = (u >> 3) & 1;
                const int ch = u & 7;
                float4 v = ld_shared_v4(stg_base + (uint32_t)tile * kStgRow + (uint32_t)(b * 128 + ch * 16));
                if (GEN && !((masktab[tile] >> (2 * a + b)) & 1)) {
                  // pixel outside an odd-sized map: not part of the dense output; zero in the padded frame (it lies
                  // in the frame's extra border row / column)
                  if (!out_padded) continue;
                  v = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (!no_out) st_out_v4(ybase + (size_t)(pixtab[tile] + b) * K + ch * 4, v, mc);
              }
            }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");  // staging rows free again
          }
        }
        if (out_padded && tvalid && (ty == 0 || ty == geo.TY - 1 || tx == 0 || tx == geo.TX - 1)) {
          // zero border of the reference's 16x16 frame (Kernel128_winograd.cu:163,243): edge tiles own their share
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const size_t rstride = (size_t)W * K;
          const int ncc = kn / 2;
          float* p = y + (size_t)pix0 * K + c0s + cq * ncc;
          const ptrdiff_t dyb = ty == 0 ? -(ptrdiff_t)rstride : (ty == geo.TY - 1 ? 2 * (ptrdiff_t)rstride : 0);
          const ptrdiff_t dxb = tx == 0 ? -(ptrdiff_t)K : (tx == geo.TX - 1 ? 2 * (ptrdiff_t)K : 0);
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
      if (DBG && ts_item && item_idx < 4) ts_i[4 * item_idx + 2] = clock64();
    }
  }

  tc_fence_before();
  if constexpr (CG2) cluster_sync_all();  // the peer's shared memory, TMEM and barriers stay alive until the pair is done
  else __syncthreads();
  if (DBG && (dbg & 64) && threadIdx.x == 0) ts_g[15] = (long long)globaltimer_ns();
  if (warp == kMmaWarp) {
    if constexpr (CG2) tmem_dealloc_cg2<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Once per layer: U = G g G^T (F(2x2,3x3)), RN-rounded to the operand type, in the shared-memory image of this kernel's
// bulk copies: per cout slice (ff::slice) [C/8 k-block][2 j-halves][4 i][2 jj][2 k-chunks][KN couts][4 channels]
// (j = 2*jh + jj). 512 bytes per (k-block, cout), so slice s starts at byte (C/8)*512*c0(s).
// Replaces the offline weight_generator loop (/root/reference/data_generator.py:63-78; that one is F(4x4), 36 points).
// op16 = 1 (bf16) / 2 (fp16): 16-channel k-blocks, 8 channels per 16-byte chunk, same 512 bytes per (k-block, cout).
// cg2: image of the CTA-pair kernel -- inside every (k-block, j-half) chunk the couts of the slice are split in two
// halves, [rank 2][8 points][2 k-chunks][KN/2 couts][16 B], CTA `rank` of a pair loads its 128*KN contiguous bytes.
// narrow: the all-64-wide slicing (ff::slice(..., narrow)).
__global__ void filter_transform_ff_kernel(const float* __restrict__ w_kcrs, float* __restrict__ u_img, int C, int K,
                                           int op16, int cg2, int narrow) {
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
  const ff::Slice sl = ff::slice(K, ff::slice_of(K, k, narrow), narrow);
  const int kn = sl.kn, c0 = sl.c0;
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
      size_t o16 = base16 + (size_t)p * (2 * kn * 8) + ((size_t)chunk16 * kn + kl) * 8 + e16;
      size_t o32 = base + (size_t)p * (2 * kn * 4) + ((size_t)chunk * kn + kl) * 4 + e;
      if (cg2) {
        const int hk = kn / 2, rank = kl / hk, klr = kl % hk, jh = j >> 1, p8 = i * 2 + (j & 1);
        o16 = base16 + (size_t)jh * (128 * kn) + (size_t)rank * (64 * kn) + (size_t)p8 * (8 * kn) +
              ((size_t)chunk16 * hk + klr) * 8 + e16;
        o32 = base + (size_t)jh * (64 * kn) + (size_t)rank * (32 * kn) + (size_t)p8 * (4 * kn) +
              ((size_t)chunk * hk + klr) * 4 + e;
      }
      if (op16 == 2) reinterpret_cast<__half*>(u_img)[o16] = __float2half_rn(u[j]);
      else if (op16 == 1) reinterpret_cast<__nv_bfloat16*>(u_img)[o16] = __float2bfloat16_rn(u[j]);
      else u_img[o32] = to_tf32_rn(u[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side

// Raw-tile layout: 1 (default) = four parity planes with a 9-slot row pitch, conflict-free patch loads (kPlaneBytes);
// 0 = the TM kernel's single box with an 8-slot pitch (WG_FF_P9=0; A/B measurements).
int wino_ff_p9() {
  static int v = -1;
  if (v < 0) {
    const char* e = dev_env("WG_FF_P9");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v;
}

// CTA pairs (cta_group::2) for layers created from now on: WG_FF_CG2=1. Needs the parity-plane raw layout.
// EXPERIMENT, default off; results are bit-identical to the single-CTA kernel. Measured (N=256, B200, us):
//                          256->256 tf32 | bf16 | 128->128 tf32 | bf16
//   single CTAs                 97.8-98.7 | 83.9 |     47.0      | 37.8
//   pairs, first version            163.4 | 118.3|     84.2      | 61.6   hand-off barriers with .release.cluster /
//                                                                         .acquire.cluster: ~1 k clk per hand-off
//   pairs, plain arrive / wait  96.8-96.9 | 85.1 |     47.6      | 40.4   (the payload is TMEM state, ordered by the
//                                                                         tcgen05 fences on both sides)
//   pairs + 16 transform warps       96.7 | 79.2 |     50.5      | 40.9
// i.e. halving the B-operand reads and the filter traffic per SM buys 1 % on the shape it was built for.
int wino_ff_cg2() {
  static int v = -1;
  if (v < 0) {
    const char* e = dev_env("WG_FF_CG2");
    v = e ? (atoi(e) != 0) : 0;
  }
  return v && wino_ff_p9();
}

// Geometry of an H x W layer: tiles, frame, plane pitch, and the largest M-block whose raw rows fit one parity plane
// (brute force over the M-block's first tile; the schedule is periodic in the tiles of one image).
int wino_ff_geo(int H, int W, ff::Geo* g) {
  if (H < 3 || W < 3 || H > 4096 || W > 4096) return WG_ERR_ARG;
  g->H = H, g->W = W;
  g->TX = (W + 1) / 2, g->TY = (H + 1) / 2;
  g->TT = g->TX * g->TY;
  g->Hf = 2 * g->TY + 2, g->Wf = 2 * g->TX + 2;
  g->SP = g->TX + 2;
  g->RPI = g->Hf / 2;
  g->mv_max = 0;
  for (int mv : {128, 96, 64, 32}) {
    int span = 0;
    for (int t0 = 0; t0 < g->TT; ++t0) {
      const int t1 = t0 + mv - 1;
      const int rp0 = (t0 / g->TT) * g->RPI + (t0 % g->TT) / g->TX;
      const int rp1 = (t1 / g->TT) * g->RPI + (t1 % g->TT) / g->TX;
      if (rp1 - rp0 + 2 > span) span = rp1 - rp0 + 2;  // a tile row needs row pairs ty and ty + 1
    }
    if ((long long)span * g->SP * 32 <= (long long)ff::kPlaneBytes && span <= 256) {
      g->mv_max = mv;
      g->rp_box = span;
      break;
    }
  }
  if (g->mv_max == 0) return WG_ERR_ARG;
  if (ff::geo_is_ref(*g)) g->rp_box = 24;  // the reference geometry keeps its round-1 box (24 row pairs x 9 slots = one full plane)
  g->raw_bytes = 4u * (uint32_t)g->rp_box * (uint32_t)g->SP * 32u;
  return WG_OK;
}

int wino_ff_make_tmap(CUtensorMap* tmap, const float* x, int n_img, int C, const ff::Geo& g) {
#ifdef WG_DEV_BUILD
  if (!wino_ff_p9()) return ff::geo_is_ref(g) ? wino_tm_make_tmap(tmap, x, n_img, C, 1) : WG_ERR_ARG;
#endif
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return WG_ERR_DRIVER;
  // x[N][Hf][Wf][C] viewed as (c, x/2, x&1, y&1, (n*Hf+y)/2); a box is one parity plane: 8 channels x SP column pairs
  // (starting at x/2 = -1) x rp_box row pairs, 32-byte swizzle. (Reference geometry: 16x16 frames, 9 x 24.)
  const cuuint64_t Wf = (cuuint64_t)g.Wf, Hf = (cuuint64_t)g.Hf;
  cuuint64_t dims[5] = {(cuuint64_t)C, Wf / 2, 2, 2, (cuuint64_t)n_img * (Hf / 2)};
  cuuint64_t strides[4] = {(cuuint64_t)2 * C * 4, (cuuint64_t)C * 4, Wf * C * 4, 2 * Wf * C * 4};
  cuuint32_t box[5] = {8, (cuuint32_t)g.SP, 1, 1, (cuuint32_t)g.rp_box};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, wg::l2_promotion(),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WG_OK : WG_ERR_TMAP;
}

int wino_ff_has_narrow(int K) { return ff::has_narrow(K) ? 1 : 0; }

int filter_transform_ff_launch(const float* w_kcrs, float* u_img, int C, int K, int op16, int cg2, int narrow,
                               cudaStream_t stream) {
  const int n = C * K;
  filter_transform_ff_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w_kcrs, u_img, C, K, op16, cg2, narrow);
  return cudaGetLastError() == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

// Tiles per M-block and grid size. The MMA is always M=128 but only `mv` rows carry tiles; transform warps own 32 rows
// each, so the per-item cost scales with ceil(mv/32) quarters: pick the mv that minimises waves x cost (WG_WINO_MV pins
// it). cg2: one item per CTA pair.
static void ff_plan(int n_img, int K, int max_ctas, bool cg2, const ff::Geo& geo, int* mv_out, int* grid_out,
                    int narrow = 0) {
  const int n_sl = ff::n_slices(K, narrow);
  const int total_tiles = n_img * geo.TT;
  int mv = geo.mv_max;
  static int mv_env = -1;
  if (mv_env < 0) {
    const char* e = dev_env("WG_WINO_MV");
    mv_env = e ? atoi(e) : 0;
  }
  if (mv_env >= 16 && mv_env <= geo.mv_max) {
    mv = mv_env;
  } else {
    double best = 1e30;
    for (int cand = geo.mv_max; cand >= 64 || cand == geo.mv_max; cand -= 32) {
      const long long items = (long long)((total_tiles + cand - 1) / cand) * n_sl;
      const long long slots = max_ctas > 0 ? max_ctas : 1;
      const long long waves = (items + slots - 1) / slots;
      const double cost = (double)waves * (0.35 + 0.65 * cand / 128.0);
      if (cost < best - 1e-9) {
        best = cost;
        mv = cand;
      }
    }
  }
  const int n_mblocks = (total_tiles + mv - 1) / mv;
  const int n_items = (cg2 ? (n_mblocks + 1) / 2 : n_mblocks) * n_sl;
  int grid = cg2 ? max_ctas / 2 : max_ctas;
  if (grid > n_items) grid = n_items;
  if (grid < 1) grid = 1;
  if (cg2) grid *= 2;
  *mv_out = mv;
  *grid_out = grid;
}

template <bool H16, bool DBG, bool P9, bool CG2, bool ALT = false, bool GEN = false>
static int launch_ff(const CUtensorMap& tmap, const float* u_img, const float* scale, const float* shift, float* y,
                     int n_img, int C, int K, int relu, int out_padded, int max_ctas, cudaStream_t stream, int fp16,
                     int dbg, const ff::Geo& geo, int narrow = 0) {
  static unsigned long long configured = 0;  // per device: the attribute is a property of the function on ONE device
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long dev_bit_ = 1ull << (dev_ & 63);
  if (!(configured & dev_bit_)) {
    if (cudaFuncSetAttribute(wino3x3_ff_kernel<H16, DBG, P9, CG2, ALT, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)ff::kTotal) != cudaSuccess)
      return WG_ERR_CUDA;
    configured |= dev_bit_;
  }
  int mv = 128, grid = 1;
  ff_plan(n_img, K, max_ctas, CG2, geo, &mv, &grid, narrow);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(32 * (ff::kWorkerWarps + 2));
  cfg.dynamicSmemBytes = ff::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG2) {
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, wino3x3_ff_kernel<H16, DBG, P9, CG2, ALT, GEN>, tmap, u_img, scale, shift, y, n_img, C,
                                     K, relu, out_padded, mv, fp16, dbg, geo, narrow);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

constexpr int kFfAltDefault = 0;  // see the ALT comment in wino_ff_launch

int wino_ff_launch(const CUtensorMap& tmap, const float* x, const float* u_img, const float* u_img_narrow,
                   const float* scale, const float* shift, float* y, int n_img, int C, int K, int op16, int cg2, int relu,
                   int out_padded, int max_ctas, const ff::Geo& geo, cudaStream_t stream) {
  // (An L2 prefetch of the next item's raw rows, cp.async.bulk.prefetch.L2 spread over the stages of the current
  //  item, was measured: no gain -- the kernel is as fast with HBM-cold as with L2-resident input.)
  (void)x;
  // op16: 0 = TF32 operands, 1 = bf16, 2 = fp16 (V packed in TMEM, 16-channel stages)
  static int dbg = -1;  // WG_FF_DEBUG: ablation switches of the developer build (see the kernel)
  if (dbg < 0) {
    const char* e = dev_env("WG_FF_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  const int fp16 = op16 == 2;
  // Sixteen transform warps, one group per V half (wino_ffw_kernel.cu). Measured at N=256: bf16 / fp16 operands
  // 256->256 82.1 -> 77.9 us (there the transform chain is the longer one), 128->128 39.7 -> 40.1; TF32 256->256
  // 98.5 -> 98.4, 128->128 47.1 -> 49.8 (1.5x the patch-load traffic on a shared-memory path that is already ~80 % busy).
  // Default: 16-bit operands with C >= 256 only; WG_FF_W16=0|1 forces it off / on for everything.
  static int w16 = -1;
  if (w16 < 0) {
    const char* e = dev_env("WG_FF_W16");
    w16 = e ? (atoi(e) != 0) : 2;
  }
  // Launches whose work items do not fill the SMs (one wave): the time is that of ONE item, so make the item short.
  //  * narrow slices: with the layer's second filter image (all slices 64 wide) an M-block is K/64 items instead of
  //    ceil(K/96), each with two V stages in TMEM and 2/3 of the MMA time per stage; used when those items still fit
  //    one wave, with the 16-warp kernel;
  //  * split-C (wino_ffw_kernel.cu, SPLIT): when the items, two CTAs each, fit on the chip at once, a cluster of 2
  //    shares one item and each CTA runs half of the channel loop. The exchange costs ~10 us (DSMEM both ways,
  //    release/acquire at cluster scope, cluster launch), so it only pays for long channel loops: C >= 256.
  // WG_FF_SPLIT=0 disables split-C, =2 lifts its C limit; WG_FF_NARROW=0 disables the narrow image.
  static int split_env = -1, narrow_env = -1;
  if (split_env < 0) {
    const char* e = dev_env("WG_FF_SPLIT");
    split_env = e ? atoi(e) : 1;
    const char* n = dev_env("WG_FF_NARROW");
    narrow_env = n ? atoi(n) : 1;
  }
  // (the 16-warp sibling and its split-C / narrow modes are built for the reference's 14x14 geometry only)
  const bool ref_geo = ff::geo_is_ref(geo);
  if (!ref_geo && (cg2 || !wino_ff_p9())) return WG_ERR_ARG;
  if (ref_geo && !cg2 && dbg == 0 && wino_ff_p9()) {
    const int n_kb = C / (op16 ? 16 : 8);
    const int n_mb = (n_img * 49 + 127) / 128;
    // (same choice with and without WG_OUT_MULTICAST: the fused gather must reproduce kernel + all-gather bit for bit)
    const bool split_ok = split_env && (C >= 256 || split_env == 2) && n_kb % 2 == 0 && n_kb >= 4;
    // measured, 256->256, us per launch (Python loop): narrow + split 29-30 (N <= 32), default slices + split 34.5-36
    // (N <= 48), narrow 43.8 (N <= 96), default 46-50
    const bool narrow_ok = narrow_env && u_img_narrow != nullptr;
    const int items_n = n_mb * (K / 64), items_d = n_mb * ff::n_slices(K);
    if (narrow_ok && split_ok && 2 * items_n <= max_ctas)
      return wino_ffw_launch(tmap, u_img_narrow, scale, shift, y, n_img, C, K, op16, 0, 1, 1, relu, out_padded, 128,
                             2 * items_n, stream);
    if (split_ok && 2 * items_d <= max_ctas)
      return wino_ffw_launch(tmap, u_img, scale, shift, y, n_img, C, K, op16, 0, 1, 0, relu, out_padded, 128, 2 * items_d,
                             stream);
    if (narrow_ok && items_n <= max_ctas)
      return wino_ffw_launch(tmap, u_img_narrow, scale, shift, y, n_img, C, K, op16, 0, 0, 1, relu, out_padded, 128,
                             items_n, stream);
  }
  const bool use_w16 = w16 == 1 || (w16 == 2 && op16 != 0 && C >= 256);
  if (ref_geo && use_w16 && dbg == 0 && wino_ff_p9()) {
    int mv = 128, grid = 1;
    ff_plan(n_img, K, max_ctas, cg2 != 0, geo, &mv, &grid);
    return wino_ffw_launch(tmap, u_img, scale, shift, y, n_img, C, K, op16, cg2, 0, 0, relu, out_padded, mv, grid, stream);
  }
#define WG_FF(H16_, DBG_, P9_, CG2_)                                                                                \
  return launch_ff<H16_, DBG_, P9_, CG2_>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded, max_ctas, stream, \
                                          fp16, dbg, geo)
  // Alternating transform groups (ALT, see the kernel): for launches whose cout slices are all <= 64 wide (two V stages
  // in TMEM) -- K <= 64, K = 128 (64 + 64) -- and, mode 2, for wider layers through the all-64-wide filter image.
  static int alt = -1;  // developer build: WG_FF_ALT=0|1|2
  if (alt < 0) {
    const char* e = dev_env("WG_FF_ALT");
    alt = e ? atoi(e) : kFfAltDefault;
  }
  if constexpr (kDev) {  // developer build: ablation flags on the ALT kernel (TF32, default slicing)
    if (ref_geo && alt > 0 && !cg2 && dbg > 0 && wino_ff_p9() && !op16 && K <= 128 && K % 64 == 0)
      return launch_ff<false, true, true, false, true>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded,
                                                       max_ctas, stream, fp16, dbg, geo, 0);
  }
  if constexpr (kDev) {
    // developer build only (tools/libwinograd_b200_dev.so): CTA pairs, the ablation instantiation (WG_FF_DEBUG, results
    // are garbage by design) and the single-box raw layout. None of these exist in the product library.
    if (cg2) {
      if (op16) WG_FF(true, false, true, true);
      WG_FF(false, false, true, true);
    }
    if (wino_ff_p9() && dbg > 0) {
      if (op16) WG_FF(true, true, true, false);
      WG_FF(false, true, true, false);
    }
    if (!wino_ff_p9()) {
      if (op16) WG_FF(true, false, false, false);
      WG_FF(false, false, false, false);
    }
  }
  if constexpr (kDev)  // experiment, measured +-1 us (profiles/wino_r02_notes.md): developer build only
  if (ref_geo && alt > 0 && !cg2 && dbg == 0 && wino_ff_p9()) {
    bool all64 = true;
    for (int sidx = 0; sidx < ff::n_slices(K); ++sidx) all64 = all64 && ff::slice(K, sidx).kn <= 64;
    if (all64) {
      if (op16)
        return launch_ff<true, false, true, false, true>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded,
                                                         max_ctas, stream, fp16, 0, geo, 0);
      return launch_ff<false, false, true, false, true>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded,
                                                        max_ctas, stream, fp16, 0, geo, 0);
    }
    if (alt == 2 && u_img_narrow != nullptr) {
      if (op16)
        return launch_ff<true, false, true, false, true>(tmap, u_img_narrow, scale, shift, y, n_img, C, K, relu, out_padded,
                                                         max_ctas, stream, fp16, 0, geo, 1);
      return launch_ff<false, false, true, false, true>(tmap, u_img_narrow, scale, shift, y, n_img, C, K, relu, out_padded,
                                                        max_ctas, stream, fp16, 0, geo, 1);
    }
  }
  if (!ref_geo) {  // other map sizes: the runtime-geometry instantiation
    if (op16)
      return launch_ff<true, false, true, false, false, true>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded,
                                                              max_ctas, stream, fp16, 0, geo, 0);
    return launch_ff<false, false, true, false, false, true>(tmap, u_img, scale, shift, y, n_img, C, K, relu, out_padded,
                                                             max_ctas, stream, fp16, 0, geo, 0);
  }
  if (op16) WG_FF(true, false, true, false);
  WG_FF(false, false, true, false);
#undef WG_FF
}

}  // namespace wg
