// Small-batch (latency) variant of the fused 3x3 Winograd F(2x2,3x3) conv + BN + ReLU for sm_100a.
//
// Same math and layouts as wino3x3_bn_relu_kernel (winograd_kernels.cu; replaces the reference's three-kernel pipeline,
// /root/reference/Kernel128_winograd.cu:28-213, Kernel256_winograd.cu:27-218), organised for N = 1..8 images, where
// the layer is a handful of 128-tile blocks and what it costs is latency, not throughput:
//
//   * work item = 64 consecutive tiles x 32 output channels, one item per thread-block CLUSTER of CS CTAs; each CTA
//     runs C/CS of the channel loop (at N=1, 256->256: 8 items x 16 CTAs = 128 SMs pull the 4 MB filter image out of
//     L2 in parallel: one SM's TMA unit delivers ~31-37 B/clk of tiled boxes and 65-110 B/clk of 1-D bulk copies,
//     profiles/tma_probe_r01.txt, tma_tensor_probe_r01.txt);
//   * tcgen05.mma with M=64 (half the A-operand shared-memory reads of M=128; TMEM row r lives in lane 32*(r/16)+r%16,
//     probed with tools/selftest), one accumulator per Winograd point (16 x 32 columns), no folding: 16 MMAs per stage;
//   * the CS partial outputs are reduced through distributed shared memory: after the inverse transform every thread
//     pushes its tile's 4 pixels x 8 couts to the CTA that owns that tile row (st.shared::cluster into a dedicated,
//     bank-swizzled inbox), then arrives on the owner's mbarrier (release.cluster); the owner waits for its 128
//     arrivals (acquire.cluster), sums the CS partials in fixed order, applies scale/shift/ReLU and stores.
//     No atomics, deterministic; the only cluster barrier is split (arrive in the prologue, wait before the push);
//   * programmatic dependent launch: the filter slices are requested before griddepcontrol.wait, the activations after.
#include "ptx.cuh"
#include "wg_internal.h"

#include <cuda.h>
#include <stdlib.h>

namespace wg {

namespace small {
constexpr int kWorkerWarps = 8;     // warps 0..3 transform, all 8 drain TMEM / reduce / store
constexpr int kTransformWarps = 4;  // 64 rows x two 4-channel halves
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;
constexpr int kThreads = 32 * 10;
constexpr int kMB = 64;             // tiles per M-block
constexpr int kKN = 32;             // output channels per item
constexpr int kRawRows = 26;        // input rows (n*16+y) 64 consecutive tiles can touch (brute-forced bound)
constexpr uint32_t kRawBytes = kRawRows * 2 * 8 * 32;  // [ny][x parity][x/2][8 ch] fp32 = 13312
constexpr int kRawStages = 2, kVStages = 2, kUBufs = 3;
constexpr uint32_t kVLbo = kMB * 16 + 64;  // k-chunk stride (+64: the two chunks of a row land in different bank halves)
constexpr uint32_t kVPerXi = kVLbo + kMB * 16;
constexpr uint32_t kVBytes = 16 * kVPerXi;  // 33792
constexpr uint32_t kULbo = kKN * 16;
constexpr uint32_t kUPerPoint = 2 * kKN * 16;
constexpr uint32_t kUBytes = 16 * kUPerPoint;  // 16 KB per 8-channel stage
constexpr uint32_t kInboxBytes = kMB * 4 * kKN * 4;  // all sources together: [CS][64/CS rows][4 px][32 couts] fp32 = 32 KB
constexpr uint32_t kOffRaw = 0;
constexpr uint32_t kOffV = kOffRaw + kRawStages * kRawBytes;
constexpr uint32_t kOffU = kOffV + kVStages * kVBytes;
constexpr uint32_t kOffInbox = kOffU + kUBufs * kUBytes;
constexpr uint32_t kOffBar = kOffInbox + kInboxBytes;
constexpr uint32_t kNumBars = 2 * kRawStages + 2 * kVStages + 2 * kUBufs + 2;
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kOffTs = kOffTmemPtr + 16;  // debug timestamps (16 slots)
constexpr uint32_t kTotal = kOffTs + 16 * 8;
static_assert(kOffV % 128 == 0 && kOffU % 128 == 0 && kOffInbox % 128 == 0 && kOffBar % 8 == 0, "alignment");
static_assert(kTotal <= 227 * 1024, "shared memory budget");
}  // namespace small

__device__ __forceinline__ float tf32_rn_operand(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

template <int CS>
__global__ void __launch_bounds__(small::kThreads, 1)
wino3x3_small_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ u_img,
                     const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ y,
                     int n_img, int C, int K, int relu, int out_padded, int debug) {
  using namespace small;
  constexpr int RO = kMB / CS;  // tile rows each CTA finishes
  const bool mc = (out_padded & 2) != 0;  // y is an NVLS multicast address: stores go out as multimem.st
  out_padded &= 1;
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // debug only (debug & 64): phase timestamps of CTA 0, slot i written by whichever single thread reaches WG_TS(i)
  volatile long long* ts = reinterpret_cast<volatile long long*>(smem + kOffTs);
#define WG_TS(i) do { if ((debug & 64) && blockIdx.x == 0) ts[i] = clock64(); } while (0)
  if (threadIdx.x == 0) WG_TS(0);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = raw_full + kRawStages;
  uint64_t* v_full = raw_empty + kRawStages;
  uint64_t* v_empty = v_full + kVStages;
  uint64_t* u_full = v_empty + kVStages;
  uint64_t* u_empty = u_full + kUBufs;
  uint64_t* acc_full = u_empty + kUBufs;
  uint64_t* inbox_full = acc_full + 1;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < kRawStages; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], kTransformWarps);
    }
    for (int i = 0; i < kVStages; ++i) {
      mbar_init(&v_full[i], kTransformWarps);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < kUBufs; ++i) {
      mbar_init(&u_full[i], 1);
      mbar_init(&u_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(inbox_full, 2 * kMB);  // one arrival per (tile row, cout half) thread of all CS sources: 2 * RO * CS
    fence_mbar_init();
  }
  __syncthreads();  // barriers are initialised CTA-wide; TMEM allocation below overlaps the first loads
  if constexpr (CS > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  if (warp == kMmaWarp) {
    tmem_alloc<512>(const_cast<uint32_t*>(tmem_ptr));
    tc_fence_before();
    asm volatile("bar.arrive 2, %0;" ::"n"(32 * (kWorkerWarps + 1)) : "memory");  // publishes the TMEM base to the workers
  }

  const int n_kb = C / 8;
  const int n_slices = K / kKN;
  const int total_tiles = n_img * 49;
  const uint32_t crank = CS > 1 ? cluster_ctarank() : 0u;
  const int item = blockIdx.x / CS;
  const int slice = item % n_slices;
  const int t0 = (item / n_slices) * kMB;
  const int ny0 = (t0 / 49) * 16 + 2 * ((t0 % 49) / 7);
  const int kb_per = n_kb / CS;         // 8-channel stages this CTA runs
  const int kb0 = (int)crank * kb_per;  // first one
  const int valid_rows = min(kMB, total_tiles - t0);

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {  // elect.sync, not lane == 0: lets ptxas keep descriptors in uniform registers (no per-MMA waterfall loop)
      const uint8_t* u_src = reinterpret_cast<const uint8_t*>(u_img) + ((size_t)slice * n_kb + kb0) * kUBytes;
      // the filter does not depend on the previous kernel in the stream: request it before waiting for that kernel
      const int pre = kb_per < kUBufs ? kb_per : kUBufs;
      for (int i = 0; i < pre; ++i) {
        mbar_arrive_expect_tx(&u_full[i], kUBytes);
        tma_bulk_g2s(smem + kOffU + i * kUBytes, u_src + (size_t)i * kUBytes, kUBytes, &u_full[i]);
      }
      pdl_wait();
      uint32_t rs = 0, rph = 0, us = 0, uph = 1;
      for (int i = 0; i < kb_per; ++i) {
        mbar_wait(&raw_empty[rs], rph ^ 1);
        mbar_arrive_expect_tx(&raw_full[rs], kRawBytes);
        tma_tensor_4d_g2s(smem + kOffRaw + rs * kRawBytes, &tmap_x, (kb0 + i) * 8, 0, 0, ny0, &raw_full[rs]);
        if (++rs == kRawStages) { rs = 0; rph ^= 1; }
        if (i >= pre) {
          mbar_wait(&u_empty[us], uph ^ 1);
          mbar_arrive_expect_tx(&u_full[us], kUBytes);
          tma_bulk_g2s(smem + kOffU + us * kUBytes, u_src + (size_t)i * kUBytes, kUBytes, &u_full[us]);
          if (++us == kUBufs) { us = 0; uph ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (elect_one()) {  // elect.sync, not lane == 0: lets ptxas keep descriptors in uniform registers (no per-MMA waterfall loop)
      constexpr uint32_t idesc = make_idesc(kFmtTF32, kMB, kKN);
      const uint32_t v_base = smem_u32(smem + kOffV);
      const uint32_t u_base = smem_u32(smem + kOffU);
      uint32_t vs = 0, vph = 0, us = 0, uph = 0;
      uint32_t tmem_base = 0;
      for (int i = 0; i < kb_per; ++i) {
        mbar_wait(&v_full[vs], vph);
        mbar_wait(&u_full[us], uph);
        if (i == 0) tmem_base = *tmem_ptr;  // written by this warp's tcgen05.alloc long before the first V stage is ready
        tc_fence_after();
        const uint32_t va = v_base + vs * kVBytes;
        const uint32_t ua = u_base + us * kUBytes;
        const uint32_t acc = i > 0 ? 1u : 0u;
#pragma unroll
        for (int xi = 0; xi < 16; ++xi) {
          const uint64_t a_desc = make_smem_desc(va + xi * kVPerXi, kVLbo, 128, kLayoutNone);
          const uint64_t b_desc = make_smem_desc(ua + xi * kUPerPoint, kULbo, 128, kLayoutNone);
          umma_tf32_ss(tmem_base + xi * kKN, a_desc, b_desc, idesc, acc);
        }
        umma_commit(&u_empty[us]);
        umma_commit(&v_empty[vs]);
        if (i < 2) WG_TS(5 + i);  // MMAs of stage i issued
        if (++us == kUBufs) { us = 0; uph ^= 1; }
        if (++vs == kVStages) { vs = 0; vph ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else {
    // ------------------------------------------------------------------ transform (warps 0..3)
    if (warp < kTransformWarps) {
      // task: row = 16*warp + q (tile within the M-block), c = which 4-channel half of the 8-channel stage
      const int c = (lane >> 2) & 1;
      const int q = (lane & 3) + 4 * (lane >> 3);
      const int trow = warp * 16 + q;
      const int T = t0 + trow;
      const bool tvalid = trow < valid_rows;
      const bool warp_active = warp * 16 < valid_rows;  // warp-uniform
      uint32_t raw_off = 0;
      {
        const int n = T / 49, t = T % 49, ty = t / 7, tx = t % 7;
        if (tvalid) raw_off = (uint32_t)((n * 16 + 2 * ty - ny0) * 512 + tx * 32 + c * 16);
      }
      const uint32_t raw_base = smem_u32(smem + kOffRaw);
      const uint32_t v_base = smem_u32(smem + kOffV);
      const uint32_t v_off = (uint32_t)(c * kVLbo + trow * 16);
      uint32_t rs = 0, rph = 0, vs = 0, vph = 0;
      for (int i = 0; i < kb_per; ++i) {
        mbar_wait(&raw_full[rs], rph);
        if (threadIdx.x == 0 && i < 2) WG_TS(1 + i);  // raw stage i has landed
        if (!warp_active) {  // nothing to transform: keep the barriers moving in step with the other warps
          if (lane == 0) mbar_arrive(&raw_empty[rs]);
          if (++rs == kRawStages) { rs = 0; rph ^= 1; }
          mbar_wait(&v_empty[vs], vph ^ 1);
          if (lane == 0) mbar_arrive(&v_full[vs]);
          if (++vs == kVStages) { vs = 0; vph ^= 1; }
          continue;
        }
        float4 d[4][4];
        if (tvalid) {
          const uint32_t a = raw_base + rs * kRawBytes + raw_off;
#pragma unroll
          for (int dy = 0; dy < 4; ++dy)
#pragma unroll
            for (int dx = 0; dx < 4; ++dx) d[dy][dx] = ld_shared_v4(a + dy * 512 + (dx & 1) * 256 + (dx >> 1) * 32);
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
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[rs]);  // the raw stage is in registers
        if (++rs == kRawStages) { rs = 0; rph ^= 1; }
        mbar_wait(&v_empty[vs], vph ^ 1);  // MMAs that read this V stage have completed
        const uint32_t vdst = v_base + vs * kVBytes + v_off;
        // row pass V = t B, round to TF32, store point (i,j) at xi = 4*i + j
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const float4 a0 = d[i4][0], a1 = d[i4][1], a2 = d[i4][2], a3 = d[i4][3];
          st_shared_v4(vdst + (4 * i4 + 0) * kVPerXi, tf32_rn_operand(a0.x - a2.x), tf32_rn_operand(a0.y - a2.y),
                       tf32_rn_operand(a0.z - a2.z), tf32_rn_operand(a0.w - a2.w));
          st_shared_v4(vdst + (4 * i4 + 1) * kVPerXi, tf32_rn_operand(a1.x + a2.x), tf32_rn_operand(a1.y + a2.y),
                       tf32_rn_operand(a1.z + a2.z), tf32_rn_operand(a1.w + a2.w));
          st_shared_v4(vdst + (4 * i4 + 2) * kVPerXi, tf32_rn_operand(a2.x - a1.x), tf32_rn_operand(a2.y - a1.y),
                       tf32_rn_operand(a2.z - a1.z), tf32_rn_operand(a2.w - a1.w));
          st_shared_v4(vdst + (4 * i4 + 3) * kVPerXi, tf32_rn_operand(a1.x - a3.x), tf32_rn_operand(a1.y - a3.y),
                       tf32_rn_operand(a1.z - a3.z), tf32_rn_operand(a1.w - a3.w));
        }
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_full[vs]);
        if (threadIdx.x == 0 && i < 2) WG_TS(3 + i);  // V stage i written
        if (++vs == kVStages) { vs = 0; vph ^= 1; }
      }
    }

    // ------------------------------------------------------------------ inverse transform + push (all 8 warps)
    // M=64 accumulators: tile row r sits in TMEM lane 32*(r/16) + r%16, so warp (quad, half) drains rows
    // 16*quad .. 16*quad+15 with its lanes 0..15, couts [16*half, 16*half+16).
    const int quad = warp & 3;
    const int half = warp >> 2;
    const int erow = quad * 16 + (lane & 15);
    const bool pusher = lane < 16;
    const bool evalid = pusher && erow < valid_rows;
    const int owner = erow / RO, lr = erow % RO;
    const uint32_t inbox_local = smem_u32(smem + kOffInbox);
    uint32_t dst, dst_bar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                 : "=r"(dst)
                 : "r"(inbox_local + (uint32_t)((crank * RO + lr) * 4 * kKN * 4)), "r"(owner));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst_bar) : "r"(smem_u32(inbox_full)), "r"(owner));
    asm volatile("bar.sync 2, %0;" ::"n"(32 * (kWorkerWarps + 1)) : "memory");
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if constexpr (CS > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // peers are resident
    mbar_wait(acc_full, 0);
    if (threadIdx.x == 0) WG_TS(7);
    tc_fence_after();
    if (quad * 16 < valid_rows) {  // warp-uniform: this warp's 16 rows hold at least one real tile
#pragma unroll 1
      for (int cc = 0; cc < 16; cc += 8) {
        const int c0 = half * 16 + cc;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + c0;
        float m[16][8];
#pragma unroll
        for (int xi = 0; xi < 16; ++xi) tmem_ld_x8(taddr + xi * kKN, m[xi]);
        tmem_ld_wait();
        float o[4][8];  // Y[a][b] at o[2*a + b], before BN (partial over this CTA's channels)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float s0[4], s1[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s0[j] = m[0 + j][e] + m[4 + j][e] + m[8 + j][e];
            s1[j] = m[4 + j][e] - m[8 + j][e] - m[12 + j][e];
          }
          o[0][e] = s0[0] + s0[1] + s0[2];
          o[1][e] = s0[1] - s0[2] - s0[3];
          o[2][e] = s1[0] + s1[1] + s1[2];
          o[3][e] = s1[1] - s1[2] - s1[3];
        }
        if (evalid) {
#pragma unroll
          for (int px = 0; px < 4; ++px)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t pos = (uint32_t)((c0 / 4 + h) ^ (lr & 7));  // 16-byte chunk, swizzled by the row
              asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + px * (kKN * 4) + pos * 16),
                           "f"(o[px][4 * h]), "f"(o[px][4 * h + 1]), "f"(o[px][4 * h + 2]), "f"(o[px][4 * h + 3])
                           : "memory");
            }
        }
      }
    }
    tc_fence_before();
    if (threadIdx.x == 0) WG_TS(8);
    // (staging the partial locally and sending it with one cp.async.bulk shared::cta -> shared::cluster per owner was
    // measured too: the copies are faster, but every CTA then has to outwait its peers before it may retire -- a net loss)
    if (pusher) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(dst_bar) : "memory");
    if (threadIdx.x == 0) WG_TS(9);

    // ------------------------------------------------------------------ owner: sum the CS partials, BN, ReLU, store
    mbar_wait_cluster(inbox_full, 0);
    if (threadIdx.x == 0) WG_TS(10);
    const int W = out_padded ? 16 : 14;
    const int o = out_padded ? 1 : 0;
    constexpr int kChunks = kKN / 4;  // 16-byte chunks per pixel
    for (int u = threadIdx.x; u < RO * 4 * kChunks; u += kWorkerWarps * 32) {
      const int ch = u % kChunks;
      const int px = (u / kChunks) & 3;
      const int r = u / (4 * kChunks);  // local row
      const int row = (int)crank * RO + r;
      if (row >= valid_rows) continue;
      const uint32_t pos = (uint32_t)(ch ^ (r & 7));
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < CS; ++p) {
        const float4 v = ld_shared_v4(inbox_local + (uint32_t)((((p * RO + r) * 4 + px) * kChunks + pos) * 16));
        acc.x += v.x;
        acc.y += v.y;
        acc.z += v.z;
        acc.w += v.w;
      }
      const int cout0 = slice * kKN + ch * 4;
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
      st_out_v4(g, acc, mc);
      if (out_padded) {
        // zero border of the reference's 16x16 frame (Kernel128_winograd.cu:163,243), written by whichever corner
        // pixel of an edge tile is nearest
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const int a = px >> 1, b = px & 1;
        const ptrdiff_t dyb = (ty == 0 && a == 0) ? -(ptrdiff_t)W * K : ((ty == 6 && a == 1) ? (ptrdiff_t)W * K : 0);
        const ptrdiff_t dxb = (tx == 0 && b == 0) ? -(ptrdiff_t)K : ((tx == 6 && b == 1) ? (ptrdiff_t)K : 0);
        if (dyb != 0) st_out_v4(g + dyb, z4, mc);
        if (dxb != 0) st_out_v4(g + dxb, z4, mc);
        if (dyb != 0 && dxb != 0) st_out_v4(g + dyb + dxb, z4, mc);
      }
    }
    if (threadIdx.x == 0) WG_TS(11);
  }

  if (warp >= kWorkerWarps) {
    __syncwarp();  // the single-lane roles rejoin their warps
    if constexpr (CS > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // pairs with the arrive above
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<512>(*tmem_ptr);
  if ((debug & 64) && threadIdx.x == 0 && blockIdx.x == 0)
    printf("wg small ts (clk from entry): raw %lld %lld V %lld %lld mma-issued %lld %lld acc_full %lld computed %lld "
           "pushed %lld inbox_full %lld stored %lld exit %lld\n",
           ts[1] - ts[0], ts[2] - ts[0], ts[3] - ts[0], ts[4] - ts[0], ts[5] - ts[0], ts[6] - ts[0], ts[7] - ts[0],
           ts[8] - ts[0], ts[9] - ts[0], ts[10] - ts[0], ts[11] - ts[0], clock64() - ts[0]);
#undef WG_TS
}

// ---------------------------------------------------------------------------------------------------------------
// host side

int wino_small_make_tmap(CUtensorMap* tmap, const float* x, int n_img, int C) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return WG_ERR_DRIVER;
  // same view as wino_make_tmap -- x[N][16][16][C] as (c, x/2, x&1, n*16+y) -- with the 26-row box of a 64-tile block
  cuuint64_t dims[4] = {(cuuint64_t)C, 8, 2, (cuuint64_t)n_img * 16};
  cuuint64_t strides[3] = {(cuuint64_t)2 * C * 4, (cuuint64_t)C * 4, (cuuint64_t)16 * C * 4};
  cuuint32_t box[4] = {8, 8, 2, (cuuint32_t)small::kRawRows};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, wg::l2_promotion(),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WG_OK : WG_ERR_TMAP;
}

// clusters of CS CTAs of this kernel that can be resident at once (0: not launchable); cached per device
template <int CS>
static int small_max_clusters() {
  static int cached[64];
  int dev_ = 0;
  cudaGetDevice(&dev_);
  int& slot = cached[dev_ & 63];
  if (slot == 0) {
    int n = 0;
    bool ok = cudaFuncSetAttribute(wino3x3_small_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)small::kTotal) == cudaSuccess;
    if (ok && CS > 8)
      ok = cudaFuncSetAttribute(wino3x3_small_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) ==
           cudaSuccess;
    if (ok && CS == 1) {
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev_);
    } else if (ok) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(CS);
      cfg.blockDim = dim3(small::kThreads);
      cfg.dynamicSmemBytes = small::kTotal;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = CS;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&n, wino3x3_small_kernel<CS>, &cfg) != cudaSuccess) n = 0;
    }
    cudaGetLastError();
    slot = n > 0 ? n : -1;
  }
  return slot > 0 ? slot : 0;
}

template <int CS>
static int launch_small(const CUtensorMap& tmap, const float* u_img, const float* scale, const float* shift, float* y,
                        int n_img, int C, int K, int relu, int out_padded, int n_items, int debug,
                        cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_items * CS));
  cfg.blockDim = dim3(small::kThreads);
  cfg.dynamicSmemBytes = small::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CS > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CS;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, wino3x3_small_kernel<CS>, tmap, u_img, scale, shift, y, n_img, C, K, relu,
                                     out_padded, debug);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

// Split factor for (n_img, C, K) on `max_ctas` SMs, 0 = the small kernel does not apply (the batch is large enough for
// the persistent kernel). Model in clocks: per 8-channel stage ~700 (ingest of 16 KB U + the raw rows through one SM's TMA unit),
// reduction ~ (CS-1)/CS of a 32 KB partial over DSMEM at ~18 B/clk + a fixed cost. WG_WINO_CS=1 disables the latency
// mode, any other value forces that CS when it is legal.
int wino_small_cs(int n_img, int C, int K, int max_ctas) {
  static int cs_env = -1;
  if (cs_env < 0) {
    const char* e = dev_env("WG_WINO_CS");
    cs_env = e ? atoi(e) : 0;
  }
  if (cs_env == 1 || C % 8 != 0 || K % small::kKN != 0) return 0;
  const int n_kb = C / 8;
  const int n_items = ((n_img * 49 + small::kMB - 1) / small::kMB) * (K / small::kKN);
  if (n_items > max_ctas) return 0;
  int best = 0;
  double best_t = 1e30;
  for (int cs : {2, 4, 8, 16}) {
    if (n_kb % cs != 0 || n_items * cs > max_ctas) continue;
    const int fit = cs == 2 ? small_max_clusters<2>() : cs == 4 ? small_max_clusters<4>()
                  : cs == 8 ? small_max_clusters<8>() : small_max_clusters<16>();
    if (n_items > fit) continue;
    const double t = (n_kb / cs) * 700.0 + (cs - 1.0) / cs * 32768.0 / 18.0 + 400.0;
    if (cs_env == cs) return cs;
    if (t < best_t) {
      best_t = t;
      best = cs;
    }
  }
  return cs_env > 1 ? 0 : best;
}

int wino_small_launch(const CUtensorMap& tmap_small, const float* u_plain, const float* scale, const float* shift,
                      float* y, int n_img, int C, int K, int relu, int out_padded, int cs, cudaStream_t stream) {
  static int debug = -1;  // WG_DEBUG_ABLATE & 64: phase timestamps (developer aid)
  if (debug < 0) {
    const char* e = dev_env("WG_DEBUG_ABLATE");
    debug = e ? atoi(e) : 0;
  }
  const int n_items = ((n_img * 49 + small::kMB - 1) / small::kMB) * (K / small::kKN);
#define WG_SMALL(CS_) \
  return launch_small<CS_>(tmap_small, u_plain, scale, shift, y, n_img, C, K, relu, out_padded, n_items, debug, stream)
  if (cs == 2) WG_SMALL(2);
  if (cs == 4) WG_SMALL(4);
  if (cs == 8) WG_SMALL(8);
  if (cs == 16) WG_SMALL(16);
#undef WG_SMALL
  return WG_ERR_ARG;
}

}  // namespace wg
