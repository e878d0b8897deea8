// Fused 1x1 conv (GEMM) + folded BatchNorm (+ ReLU) for sm_100a.
//
// Replaces kernel_512_one_128 / kernel_128_one_512 (/root/reference/Kernel128_one.cu:24-54, :244-273) and
// kernel_1024_one_256 / kernel_256_one_1024 (Kernel256_one.cu:26-56, :246-274): C[M x Cout] =
// act(scale * (A[M x Cin] * B[Cin x Cout]) + shift), A = NHWC activations (M = N*196 pixels), FP32 in / FP32 out.
//
// Persistent, warp-specialised: warp 0 = TMA producer (A tile through a 2-D tensor map with 128-byte swizzle,
// B tile as one bulk copy of a pre-swizzled image written once per layer by weight_pack_kernel), warp 1 = single
// thread issuing tcgen05.mma kind::tf32 (M=128, N=BN, K=8, four per 32-channel stage) into one of two TMEM
// accumulator buffers, warps 2..5 = epilogue overlapping the next tile's MMAs: tcgen05.ld -> scale/shift/ReLU ->
// 128-byte-swizzled shared staging (conflict-free 128-bit stores) -> per-warp TMA tensor stores of 32x32 sub-tiles
// (full 128-byte lines to L2; rows beyond M are clipped by the tensor map).
#include <stdlib.h>

#include <cuda_bf16.h>

#include "ptx.cuh"
#include "wg_internal.h"

namespace wg {

// Chain mode (out_padded): pixel m of the dense [N][H][W] map goes to (+1,+1) of the zero-bordered [N][Hf][Wf][Cout]
// frame a following 3x3 layer reads (Kernel128_winograd.cu:163,243 layout for 14x14 / 16x16; odd map sizes have a
// 2-wide bottom / right border, see ff::Geo); edge pixels also write their share of the border zeros.
// one_frame_index: frame pixel of dense pixel m and a border code (bit 0 top, bits 1-2 bottom rows, bit 3 left,
// bits 4-5 right columns) -- the divisions are done once per (item, row), not per stored vector.
__device__ __forceinline__ int2 one_frame_index(long long m, const OneGeo& g) {
  const int P = g.H * g.W;
  const int n = (int)(m / P), p = (int)(m - (long long)n * P), oy = p / g.W, ox = p - oy * g.W;
  const int code = (oy == 0 ? 1 : 0) | ((oy == g.H - 1 ? g.Hf - g.H - 1 : 0) << 1) | (ox == 0 ? 8 : 0) |
                   ((ox == g.W - 1 ? g.Wf - g.W - 1 : 0) << 4);
  return make_int2((n * g.Hf + oy + 1) * g.Wf + ox + 1, code);
}
__device__ __forceinline__ void one_frame_store(float* __restrict__ y, int2 pc, const OneGeo& g, int Cout, int col,
                                                float4 val) {
  float* q = y + (size_t)pc.x * Cout + col;
  *reinterpret_cast<float4*>(q) = val;
  if (pc.y == 0) return;
  const int dy0 = -(pc.y & 1), dy1 = (pc.y >> 1) & 3, dx0 = -((pc.y >> 3) & 1), dx1 = (pc.y >> 4) & 3;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const ptrdiff_t rs = (ptrdiff_t)g.Wf * Cout;
  for (int dy = dy0; dy <= dy1; ++dy)
    for (int dx = dx0; dx <= dx1; ++dx)
      if (dy | dx) *reinterpret_cast<float4*>(q + dy * rs + (ptrdiff_t)dx * Cout) = z4;
}

constexpr int kOneThreads = 32 * 6;
constexpr int kBK = 32;  // fp32 channels per stage = one 128-byte swizzle row

// WS = weight-stationary: the whole [BN x Cin] weight tile stays resident in shared memory (<= 128 KB) while the CTA
// walks the M-tiles of ONE N-tile; the stage ring then carries activations only.
// H16: bf16 operands -- the weight block is [4 k-chunks][BN couts][8 bf16] (K-major, no swizzle), half the bytes.
template <int BN, bool WS = false, bool H16 = false>
struct OneSmem {
  static constexpr int kStages = WS ? 4 : (BN == 256 ? 4 : 6);
  static constexpr uint32_t kABytes = 128 * 128;  // 128 rows x 128 B
  static constexpr uint32_t kBBytes = BN * (H16 ? 64 : 128);  // one 32-channel block of the weight tile
  static constexpr uint32_t kBResident = 128 * 1024;
  static constexpr uint32_t kStageOutBytes = 32 * 128;  // one warp's 32 rows x 32 fp32 columns
  static constexpr uint32_t kOffA = 0;
  static constexpr uint32_t kOffB = kOffA + kStages * kABytes;
  static constexpr uint32_t kOffOut = kOffB + (WS ? kBResident : kStages * kBBytes);  // [4 warps][2 buffers]
  static constexpr uint32_t kOffBar = kOffOut + 4 * 2 * kStageOutBytes;
  // + [4 epilogue warps][2 buffers] "residual sub-tile landed" + H16: [stages] "A stage converted into TMEM"
  static constexpr uint32_t kNumBars = 2 * kStages + 5 + 8 + kStages;
  static constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr uint32_t kOffFrame = kOffTmemPtr + 16;  // chain mode: int2 (frame pixel, border code) per tile row
  static constexpr uint32_t kTotal = kOffFrame + 128 * 8 + 1024;  // + slack for manual 1024-B alignment
  static_assert(kOffOut % 1024 == 0, "swizzled staging must be 1024-byte aligned");
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

// CL = thread-block cluster size: the CL CTAs of a cluster work on CL consecutive M-tiles of the same N-tile and share
// the weight tile -- each loads 1/CL of it and multicasts it to all (L2->SM traffic for B divided by CL; that feed,
// ~10 TB/s on B200, is what capped the non-clustered kernel on the Cin=1024 shape).
// WS (with CL == 1): weight-stationary schedule for Cin*BN*4 <= 128 KB (128->512): CTA b keeps N-tile b % n_ntiles
// for its whole life and loads that weight tile ONCE; per output tile the TMA unit then moves A + the output instead
// of A + B + the output (measured 31.9 -> 29.6 us at N=256; see one_launch for the variant that lost).
// PAIR (with CL == 2): the two CTAs of a cluster form a tcgen05 cta_group::2 pair -- two consecutive M-tiles of the same
// N-tile computed by ONE stream of M = 256 MMAs issued by the leader CTA. Each CTA loads its own activation tile and only
// HALF of the weight tile (rows [rank*BN/2, +BN/2), no multicast), and the tensor core reads each half from the shared
// memory it lives in: per SM the weight tile costs half the L2 -> SM traffic, half the shared-memory writes and half the
// B-operand reads. Hand-offs: the peer's (otherwise idle) MMA warp relays "stage landed" to the leader's full[] barrier,
// the leader's commits are multicast to both CTAs, the peer's epilogue warps arrive on the leader's acc_empty[].
// EXPERIMENT (WG_ONE_PAIR=1), bit-identical, measured at N=256: 256->1024 65.3 -> 64.7 us, 1024->256 56.7 -> 57.7,
// 512->128 27.2 -> 27.5: the weight traffic is not what these kernels wait for. (Neither is the epilogue's instruction
// count per se: eight epilogue warps ran 256->1024 in 78 us, epilogue warps storing the dense output themselves instead
// of TMA tensor stores in 85 us.) Default off.
// RES: residual add fused into the epilogue (the step that follows the reference's `_out` layers, which is why those
// stop before the ReLU -- Kernel128_one.cu:271-272, Kernel256_one.cu:273): y = [relu](scale * acc + shift + r). The
// residual sub-tile (32 rows x 32 couts, same geometry as the output sub-tile) is TMA-loaded straight INTO the warp's
// output staging buffer one chunk ahead, the epilogue adds in place and TMA-stores the same buffer: no extra shared
// memory, one extra HBM read of the output's size.
// H16: bf16 operands (the stated bf16 variant of the 1x1 path; fp32 I/O and accumulation, tolerance 1e-2). The activation
// still arrives as fp32 through TMA; four extra warps (6..9) convert each 128 x 32 stage to bf16 pairs and write it into
// TENSOR MEMORY (16 columns per stage, ring of kStages), from where the MMAs (kind::f16, K = 16, two per stage) read it
// as their A operand -- the same V-in-TMEM hand-off as the 3x3 kernels. B = bf16 weight image in shared memory.
template <int BN, int CL, bool WS = false, bool PAIR = false, bool RES = false, bool H16 = false>
__global__ void __launch_bounds__(H16 ? kOneThreads + 128 : kOneThreads, 1)
conv1x1_bn_act_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_y,
                      const __grid_constant__ CUtensorMap tmap_r, const float* __restrict__ w_img,
                      const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ y_padded,
                      long long m_rows, int Cin, int Cout, int relu, int bn_packed, int relu_after, int dev_flags,
                      const OneGeo geo, float* __restrict__ y_dbg) {
  using S = OneSmem<BN, WS, H16>;
  // dev_flags: always 0 in the product build. Developer build (WG_ONE_ABLATE): 1 no weight loads, 2 no activation loads,
  // 4 no output stores, 8 no MMAs -- results are garbage, only the time is of interest (profiles/one_ablation_r02.md).
  const int abl = kDev ? dev_flags : 0;
  if (kDev && (abl & 16) && threadIdx.x == 0)
    reinterpret_cast<long long*>(y_dbg)[1024 + 4 * blockIdx.x + 0] = (long long)globaltimer_ns();  // timeline: CTA entry
  static_assert(!RES || (CL == 1 && !PAIR), "the residual epilogue exists for the plain and weight-stationary schedules");
  static_assert(!WS || CL == 1, "weight-stationary schedule has no cluster variant");
  static_assert(!PAIR || (CL == 2 && !WS), "CTA pairs are clusters of 2");
  static_assert(!H16 || (BN == 128 && CL == 1 && !WS && !PAIR), "bf16 operands: plain schedule, 128-wide N-tiles");
  // two accumulator buffers (+ H16: the A ring, 16 columns per stage, after them; allocation is a power of two)
  constexpr uint32_t kTmemCols = H16 ? 512 : 2 * BN;
  constexpr uint32_t kACol0 = 2 * BN;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned stage buffers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* full = bars;
  uint64_t* empty = full + S::kStages;
  uint64_t* acc_full = empty + S::kStages;  // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint64_t* b_full = acc_empty + 2;         // WS: the resident weight tile has landed
  uint64_t* res_full = b_full + 1;          // RES: [epilogue warp][buffer]
  uint64_t* a_ready = res_full + 8;         // H16: [stage] converted A stage is in TMEM (4 converter warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::kOffTmemPtr);

  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kClusterMask = (uint16_t)((1u << CL) - 1u);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_y);
    for (int i = 0; i < S::kStages; ++i) {
      // PAIR, leader: own TMA bytes + the peer's relay; one multicast commit frees a stage in both CTAs
      mbar_init(&full[i], (PAIR && crank == 0) ? 2 : 1);
      // every CTA of the cluster must have consumed a stage before it is refilled; H16: the 4 converter warps (done
      // reading the fp32 stage) + the commit of the MMAs that read the weight block and the TMEM A slot
      mbar_init(&empty[i], H16 ? 5 : (PAIR ? 1 : CL));
      mbar_init(&a_ready[i], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], PAIR ? 8 : 4);
    }
    mbar_init(b_full, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&res_full[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_cg2<kTmemCols>(tmem_ptr);
    else tmem_alloc<kTmemCols>(tmem_ptr);
  }
  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_kb = Cin / kBK;
  const int n_ntiles = Cout / BN;
  const int n_mtiles = (int)((m_rows + 127) / 128);
  // an item = CL consecutive M-tiles x one N-tile, N-tile fastest; WS: an item = one M-tile of this CTA's fixed N-tile
  // (the host makes gridDim.x a multiple of n_ntiles)
  const int n_items = WS ? n_mtiles : ((n_mtiles + CL - 1) / CL) * n_ntiles;
  const int first_item = WS ? (int)blockIdx.x / n_ntiles : (int)blockIdx.x / CL;
  const int item_stride = WS ? (int)gridDim.x / n_ntiles : (int)gridDim.x / CL;
  const int ws_nt = (int)blockIdx.x % n_ntiles;
  // weight image: [Cout/bn_packed][Cin/32][bn_packed rows][128 B]; this kernel's N-tile may be a BN-row slice of a packed
  // tile (bn_packed is a multiple of BN; the swizzle only depends on row % 8)
  // (H16: [Cout/BN][Cin/32][4 k-chunks][BN couts][8 bf16], one contiguous block per (N-tile, 32-channel block))
  const size_t b_kb_stride = H16 ? (size_t)S::kBBytes : (size_t)bn_packed * 128;
  auto b_tile = [&](int nt) {
    if constexpr (H16) return reinterpret_cast<const uint8_t*>(w_img) + (size_t)nt * n_kb * S::kBBytes;
    const int col0 = nt * BN;
    return reinterpret_cast<const uint8_t*>(w_img) + ((size_t)(col0 / bn_packed) * n_kb * bn_packed + col0 % bn_packed) * 128;
  };
#define WG_ITEM_NT(item) (WS ? ws_nt : (item) % n_ntiles)
#define WG_ITEM_MT(item) (WS ? (item) : ((item) / n_ntiles) * CL + (int)crank)

  if (warp == 0) {
    if (elect_one()) {  // elect.sync, not lane == 0: lets ptxas keep descriptors in uniform registers (no per-MMA waterfall loop)
      uint32_t st = 0, ph = 0;
      if constexpr (WS) {  // the resident weight tile: requested before waiting for the previous kernel
        const uint8_t* b_src = b_tile(ws_nt);
        mbar_arrive_expect_tx(b_full, (uint32_t)n_kb * S::kBBytes);
        for (int kb = 0; kb < n_kb; ++kb)
          tma_bulk_g2s(smem + S::kOffB + kb * S::kBBytes, b_src + (size_t)kb * b_kb_stride, S::kBBytes, b_full);
      }
      pdl_wait();  // activations come from the previous kernel in the stream
      if (kDev && (abl & 16) && blockIdx.x == 0) reinterpret_cast<long long*>(y_dbg)[3] = clock64();  // timeline: past the wait
      if (kDev && (abl & 16)) reinterpret_cast<long long*>(y_dbg)[1024 + 4 * blockIdx.x + 1] = (long long)globaltimer_ns();
      for (int item = first_item; item < n_items; item += item_stride) {
        const int nt = WG_ITEM_NT(item);
        const int mt = WG_ITEM_MT(item);
        const uint8_t* b_src = b_tile(nt);
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&empty[st], ph ^ 1);
          const uint32_t a_bytes = (abl & 2) ? 0u : S::kABytes;
          const uint32_t b_bytes = (abl & 1) ? 0u : (WS ? 0u : (PAIR ? S::kBBytes / 2 : S::kBBytes));
          mbar_arrive_expect_tx(&full[st], a_bytes + b_bytes);
          if (!(abl & 2)) tma_tensor_2d_g2s(smem + S::kOffA + st * S::kABytes, &tmap_a, kb * kBK, mt * 128, &full[st]);
          if (abl & 1) {
          } else if constexpr (WS) {
          } else if constexpr (PAIR) {  // this CTA's half of the weight rows, at offset 0 of the stage in BOTH CTAs
            tma_bulk_g2s(smem + S::kOffB + st * S::kBBytes, b_src + (size_t)kb * b_kb_stride + crank * (S::kBBytes / 2),
                         S::kBBytes / 2, &full[st]);
          } else if constexpr (CL == 1) {
            tma_bulk_g2s(smem + S::kOffB + st * S::kBBytes, b_src + (size_t)kb * b_kb_stride, S::kBBytes, &full[st]);
          } else {
            constexpr uint32_t part = S::kBBytes / CL;  // rows [crank*BN/CL, (crank+1)*BN/CL) of the swizzled image
            tma_bulk_g2s_mcast(smem + S::kOffB + st * S::kBBytes + crank * part,
                               b_src + (size_t)kb * b_kb_stride + crank * part, part, &full[st], kClusterMask);
          }
          if (++st == S::kStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && PAIR && crank != 0) {
    // peer of a CTA pair: no MMAs to issue; relay "stage st has landed here" to the leader's full[st]
    if (elect_one()) {
      uint32_t st = 0, ph = 0;
      for (int item = first_item; item < n_items; item += item_stride)
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&full[st], ph);
          mbar_arrive_remote_plain(&full[st], 0);
          if (++st == S::kStages) { st = 0; ph ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // elect.sync, not lane == 0: lets ptxas keep descriptors in uniform registers (no per-MMA waterfall loop)
      constexpr uint32_t idesc = make_idesc(kFmtTF32, PAIR ? 256 : 128, BN);
      const uint32_t a_base = smem_u32(smem + S::kOffA);
      const uint32_t b_base = smem_u32(smem + S::kOffB);
      uint32_t st = 0, ph = 0;
      uint32_t it = 0;
      if constexpr (WS) mbar_wait(b_full, 0);
      // developer build, dev_flags & 16 (with 4 = no output stores): CTA 0 dumps clock64() stamps of its first 16 items
      // into y: long long y[8 * it + {0: MMA thread at item start, 1: accumulator buffer free, 2: last commit issued,
      // 4: epilogue warp 2 sees acc_full, 5: its chunk loop done}]
      long long* ts = reinterpret_cast<long long*>(y_dbg);
      const bool ts_on = kDev && (abl & 16) && blockIdx.x == 0;
      for (int item = first_item; item < n_items; item += item_stride, ++it) {
        const uint32_t buf = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        if (ts_on && it < 16) ts[8 * it + 0] = clock64();
        mbar_wait(&acc_empty[buf], aph ^ 1);
        tc_fence_after();
        if (ts_on && it < 16) ts[8 * it + 1] = clock64();
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&full[st], ph);
          if constexpr (H16) mbar_wait(&a_ready[st], ph);  // the converter warps have written this stage's A into TMEM
          tc_fence_after();
          if (abl & 8) {
          } else if constexpr (H16) {
            constexpr uint32_t idesc16 = make_idesc(kFmtBF16, 128, BN);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // B: [4 k-chunks][BN couts][16 B]; MMA k uses chunks 2k, 2k+1 (LBO = chunk pitch, SBO = 8 couts)
              const uint64_t b_desc = make_smem_desc(b_base + st * S::kBBytes + k * 2 * (BN * 16), BN * 16, 128, kLayoutNone);
              umma_f16_ts(tmem_base + buf * BN, tmem_base + kACol0 + st * 16 + k * 8, b_desc, idesc16,
                          (kb > 0 || k > 0) ? 1u : 0u);
            }
          } else {
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            const uint64_t a_desc = make_smem_desc(a_base + st * S::kABytes + k * 32, 0, 1024, kLayoutSW128);
            const uint64_t b_desc =
                make_smem_desc(b_base + (WS ? kb : (int)st) * S::kBBytes + k * 32, 0, 1024, kLayoutSW128);
            if constexpr (PAIR) umma_tf32_ss_cg2(tmem_base + buf * BN, a_desc, b_desc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else umma_tf32_ss(tmem_base + buf * BN, a_desc, b_desc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          }
          if constexpr (PAIR) umma_commit_mcast_cg2(&empty[st], kClusterMask);
          else if constexpr (CL == 1) umma_commit(&empty[st]);
          else umma_commit_mcast(&empty[st], kClusterMask);
          if (++st == S::kStages) { st = 0; ph ^= 1; }
        }
        if constexpr (PAIR) umma_commit_mcast_cg2(&acc_full[buf], kClusterMask);
        else umma_commit(&acc_full[buf]);
        if (ts_on && it < 16) ts[8 * it + 2] = clock64();
      }
    }
  } else if (H16 && warp >= 6) {
    // bf16 operands: converter warps. Thread = one row of the 128 x 32 fp32 stage (TMEM lane = row): 8 LDS.128 of its
    // swizzled 128-byte row, 16 bf16 pairs (column c = channels 2c | 2c+1 << 16), 4 tcgen05.st.x4.
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t a_base = smem_u32(smem + S::kOffA);
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + kACol0;
    uint32_t st = 0, ph = 0;
    for (int item = first_item; item < n_items; item += item_stride)
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&full[st], ph);
        const uint32_t src = a_base + st * S::kABytes + (uint32_t)row * 128;
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = ld_shared_v4(src + ((j ^ (row & 7)) << 4));
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // 8 channels -> 4 columns per x4 store
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[2 * q].x, v[2 * q].y);
          const __nv_bfloat162 p1 = __floats2bfloat162_rn(v[2 * q].z, v[2 * q].w);
          const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[2 * q + 1].x, v[2 * q + 1].y);
          const __nv_bfloat162 p3 = __floats2bfloat162_rn(v[2 * q + 1].z, v[2 * q + 1].w);
          tmem_st_x4(trow + st * 16 + q * 4, __uint_as_float(*reinterpret_cast<const uint32_t*>(&p0)),
                     __uint_as_float(*reinterpret_cast<const uint32_t*>(&p1)),
                     __uint_as_float(*reinterpret_cast<const uint32_t*>(&p2)),
                     __uint_as_float(*reinterpret_cast<const uint32_t*>(&p3)));
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a_ready[st]);
          mbar_arrive(&empty[st]);  // this warp is done with the fp32 stage
        }
        if (++st == S::kStages) { st = 0; ph ^= 1; }
      }
  } else {
    const int quad = warp & 3;  // TMEM lane quadrant this warp may touch
    uint8_t* stage_out = smem + S::kOffOut + quad * 2 * S::kStageOutBytes;
    const uint32_t stage_u32 = smem_u32(stage_out);
    uint32_t it = 0, chunk = 0;
    constexpr int kChunks = BN / 32;  // 32-cout sub-tiles per item
    uint64_t* rbar = res_full + quad * 2;
    int2* frame_tab = reinterpret_cast<int2*>(smem + S::kOffFrame);
    // RES: request the residual sub-tile of (item, chunk index ci) into staging buffer `b`
    auto res_request = [&](int item, int ci, uint32_t b) {
      const int nt = WG_ITEM_NT(item);
      const int mt = WG_ITEM_MT(item);
      mbar_arrive_expect_tx(&rbar[b], S::kStageOutBytes);
      tma_tensor_2d_g2s(stage_out + b * S::kStageOutBytes, &tmap_r, nt * BN + ci * 32, mt * 128 + quad * 32, &rbar[b]);
    };
    if constexpr (RES) {
      pdl_wait();  // the residual may come from the previous kernel in the stream
      if (lane == 0 && first_item < n_items && (long long)WG_ITEM_MT(first_item) * 128 + quad * 32 < m_rows)
        res_request(first_item, 0, 0);
    }
    for (int item = first_item; item < n_items; item += item_stride, ++it) {
      const int nt = WG_ITEM_NT(item);
      const int mt = WG_ITEM_MT(item);
      const uint32_t buf = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const float* sc = scale + nt * BN;
      const float* sh = shift + nt * BN;
      const bool rows_here = (long long)mt * 128 + quad * 32 < m_rows;  // warp-uniform: this warp's 32 rows exist
      if (y_padded != nullptr) {  // chain mode: where this warp's 32 rows go in the padded frame (read back after a warp sync)
        const long long m = (long long)mt * 128 + quad * 32 + lane;
        int2 pc = m < m_rows ? one_frame_index(m, geo) : make_int2(-1, 0);
        if (geo.interior_only) pc.y = 0;  // WG_OUT_INTERIOR_ONLY: the caller keeps the border zero
        frame_tab[quad * 32 + lane] = pc;
        __syncwarp();
      }
      mbar_wait(&acc_full[buf], aph);
      tc_fence_after();
      if (kDev && (abl & 16) && blockIdx.x == 0 && warp == 2 && lane == 0 && it < 16)
        reinterpret_cast<long long*>(y_dbg)[8 * it + 4] = clock64();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BN;
      // One 32-cout chunk whose accumulator columns are already in v[]: BN (+ReLU, + residual), staging, store.
      auto process_chunk = [&](float (&v)[32], int c0, bool cts, long long* cst) {
        // Folded BN of the chunk's 32 couts, all 16 loads issued before anything that orders memory (the shared-memory
        // stores below are asm volatile with a memory clobber: loads placed between them are NOT hoisted by the
        // compiler and each pair then costs a full L1 round trip -- 8 x ~50 clk per chunk, measured).
        float4 s4[8], h4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s4[j] = __ldg(reinterpret_cast<const float4*>(sc + c0 + 4 * j));
          h4[j] = __ldg(reinterpret_cast<const float4*>(sh + c0 + 4 * j));
        }
        if constexpr (RES) {
          // next sub-tile's residual goes into the OTHER buffer: the store that last read it (previous chunk) must be
          // done reading; then wait for this chunk's residual (requested one chunk ago)
          if (lane == 0) {
            tma_store_wait_read<0>();
            const int ci = c0 / 32 + 1;
            const int nitem = ci < kChunks ? item : item + item_stride;
            if (nitem < n_items && (long long)WG_ITEM_MT(nitem) * 128 + quad * 32 < m_rows)
              res_request(nitem, ci < kChunks ? ci : 0, (chunk + 1) & 1);
          }
          __syncwarp();
          if (rows_here) mbar_wait(&rbar[chunk & 1], (chunk >> 1) & 1);
        } else {
          // the staging buffer written two chunks ago must have been read by its TMA store
          if (lane == 0) tma_store_wait_read<1>();
          if (cts) cst[1] = clock64();
          __syncwarp();
        }
        const uint32_t dst = stage_u32 + (chunk & 1) * S::kStageOutBytes + lane * 128;
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
          const uint32_t sa = dst + ((j ^ (lane & 7)) << 4);  // 128-byte swizzle
          if constexpr (RES) {
            if (rows_here) {
              const float4 r4 = ld_shared_v4(sa);  // the TMA load used the same swizzle: this thread's own row
              o.x += r4.x, o.y += r4.y, o.z += r4.z, o.w += r4.w;
              if (relu_after) {
                o.x = fmaxf(o.x, 0.f);
                o.y = fmaxf(o.y, 0.f);
                o.z = fmaxf(o.z, 0.f);
                o.w = fmaxf(o.w, 0.f);
              }
            }
          }
          st_shared_v4(sa, o.x, o.y, o.z, o.w);
        }
        if (y_padded != nullptr) {
          // chain mode: write into the zero-bordered [N][16][16][Cout] frame a following 3x3 layer reads
          // (Kernel128_winograd.cu:163,243 layout). Rows are not affine in m there, so no TMA: 8 lanes write one
          // pixel's 128 bytes, 4 pixels per instruction; edge pixels also write their share of the border zeros.
          __syncwarp();
          const int j = lane & 7, rsub = lane >> 3;
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int r = i8 * 4 + rsub;
            const int2 pc = frame_tab[quad * 32 + r];
            if (pc.x >= 0) {
              const float4 val =
                  ld_shared_v4(stage_u32 + (chunk & 1) * S::kStageOutBytes + r * 128 + ((j ^ (r & 7)) << 4));
              one_frame_store(y_padded, pc, geo, Cout, nt * BN + c0 + j * 4, val);
            }
          }
          __syncwarp();  // staging buffer free again
          ++chunk;
          return;
        }
        if (cts) cst[3] = clock64();
        fence_proxy_async_smem();
        __syncwarp();
        if (cts) cst[4] = clock64();
        if (lane == 0 && rows_here && !(abl & 4)) {
          tma_tensor_2d_s2g(&tmap_y, stage_out + (chunk & 1) * S::kStageOutBytes, nt * BN + c0, mt * 128 + quad * 32);
          tma_store_commit();
        }
        if (cts) cst[5] = clock64();
        ++chunk;
      };
      // Chunk loop, software-pipelined over two register sets: the tcgen05.ld of chunk c + 1 (bound by the 64 B/clk of
      // TMEM read bandwidth: ~350 clk for this warp's 4 KB while all four epilogue warps load) is in flight while chunk
      // c is scaled, staged and stored. tcgen05.wait::ld covers every earlier load, so each load is issued right after
      // the wait for its predecessor.
      float va[32], vb[32];
      tmem_ld_x32(taddr, va);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 64) {
        // developer build timeline: phases of every chunk of CTA 0 / epilogue warp 2 / its third item
        const bool cts = kDev && (abl & 16) && blockIdx.x == 0 && warp == 2 && lane == 0 && it == 2;
        long long* cst = reinterpret_cast<long long*>(y_dbg) + 2048 + 8 * (c0 / 32);
        if (cts) cst[0] = clock64();
        tmem_ld_wait();
        tmem_ld_x32(taddr + c0 + 32, vb);
        if (cts) cst[2] = clock64();
        process_chunk(va, c0, cts, cst);
        if (cts) cst[8] = clock64();
        tmem_ld_wait();
        if (c0 + 64 < BN) {
          tmem_ld_x32(taddr + c0 + 64, va);
        }
        if (cts) cst[10] = clock64();
        process_chunk(vb, c0 + 32, cts, cst + 8);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && crank != 0) mbar_arrive_remote_plain(&acc_empty[buf], 0);  // the leader issues the MMAs of both CTAs
        else mbar_arrive(&acc_empty[buf]);
      }
      if (kDev && (abl & 16) && blockIdx.x == 0 && warp == 2 && lane == 0 && it < 16)
        reinterpret_cast<long long*>(y_dbg)[8 * it + 5] = clock64();
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all(); else __syncthreads();  // no CTA may leave while peers still multicast to it
  if (kDev && (abl & 16) && threadIdx.x == 0) {  // timeline: per-CTA globaltimer at exit (and the item count)
    reinterpret_cast<long long*>(y_dbg)[1024 + 4 * blockIdx.x + 2] = (long long)globaltimer_ns();
    reinterpret_cast<long long*>(y_dbg)[1024 + 4 * blockIdx.x + 3] = (n_items - first_item + item_stride - 1) / item_stride;
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc_cg2<kTmemCols>(tmem_base);
    else tmem_dealloc<kTmemCols>(tmem_base);
  }
#undef WG_ITEM_NT
#undef WG_ITEM_MT
}

// Small batches (latency): one cluster per (128-row M-tile, 64-cout sub-tile), split-K across its CS CTAs.
// With M = N*196 rows there are only a handful of 128-row tiles; what a layer costs at N=1 is (a) how fast the weights
// (the bulk of the bytes) stream out of L2 -- one SM's TMA unit delivers 37-100 B/clk, so they are spread over as many SMs as possible: 64-cout
// sub-tiles of the packed weight image and Cin/CS channels per CTA -- and (b) the reduction of the CS partial tiles,
// which goes over distributed shared memory (~20 B/clk per SM), so the partial tile is kept small (128 x 64 fp32).
// Each CTA pushes row r of its partial accumulator to the CTA owning that row (st.shared::cluster into a dedicated,
// bank-swizzled inbox [source][row][64]), then arrives on the owner's mbarrier (release.cluster); the owner waits for
// its 128 arrivals (acquire.cluster), sums the CS partials in fixed order, applies scale/shift/ReLU and writes 128-bit
// runs. No cluster barrier after the prologue, no atomics, deterministic.
constexpr int kNS = 64;

struct SmallSmem {
  static constexpr int kStages = 6;
  static constexpr uint32_t kABytes = 128 * 128;
  static constexpr uint32_t kBBytes = kNS * 128;
  static constexpr uint32_t kOffA = 0;
  static constexpr uint32_t kOffB = kOffA + kStages * kABytes;
  static constexpr uint32_t kOffInbox = kOffB + kStages * kBBytes;  // [CS sources][128/CS rows][64] fp32
  static constexpr uint32_t kOffBar = kOffInbox + 128 * kNS * 4;
  static constexpr uint32_t kNumBars = 2 * kStages + 2;
  static constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr uint32_t kTotal = kOffTmemPtr + 16 + 1024;
};

template <int CS>
__global__ void __launch_bounds__(kOneThreads, 1)
conv1x1_small_kernel(const __grid_constant__ CUtensorMap tmap_a, const float* __restrict__ w_img,
                     const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ y,
                     long long m_rows, int Cin, int Cout, int BN, int relu, int out_padded,
                     const float* __restrict__ residual, int relu_after, const OneGeo geo) {
  using S = SmallSmem;
  constexpr uint32_t kTmemCols = kNS;
  constexpr int RO = 128 / CS;  // rows of the tile each CTA finishes
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* full = bars;
  uint64_t* empty = full + S::kStages;
  uint64_t* acc_full = empty + S::kStages;
  uint64_t* inbox_full = acc_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::kOffTmemPtr);
  const uint32_t crank = CS > 1 ? cluster_ctarank() : 0u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    for (int i = 0; i < S::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(inbox_full, 128);  // RO rows x CS sources, one arrival per pushed row
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before();
  if constexpr (CS > 1) cluster_sync_all(); else __syncthreads();  // peers exist, their barriers are initialised
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_kb = Cin / kBK;
  const int n_sub = Cout / kNS;
  const int item = blockIdx.x / CS;
  const int ns = item % n_sub;
  const int mt = item / n_sub;
  const int kb_per = n_kb / CS, kb0 = (int)crank * kb_per;

  if (warp == 0) {
    if (elect_one()) {  // elect.sync, not lane == 0: lets ptxas keep descriptors in uniform registers (no per-MMA waterfall loop)
      // 64-row slice of the packed [Cout/BN][Cin/32][BN][32] image (the swizzle only depends on row % 8)
      const int col0 = ns * kNS;
      const uint8_t* b_src = reinterpret_cast<const uint8_t*>(w_img) +
                             ((size_t)(col0 / BN) * n_kb * BN + (size_t)(col0 % BN)) * 128;
      const size_t b_kb_stride = (size_t)BN * 128;
      // weights do not depend on the previous kernel in the stream: request them before waiting for it
      const int pre = kb_per < S::kStages ? kb_per : S::kStages;
      for (int i = 0; i < pre; ++i) {
        mbar_arrive_expect_tx(&full[i], S::kABytes + S::kBBytes);
        tma_bulk_g2s(smem + S::kOffB + i * S::kBBytes, b_src + (size_t)(kb0 + i) * b_kb_stride, S::kBBytes, &full[i]);
      }
      pdl_wait();
      for (int i = 0; i < pre; ++i)
        tma_tensor_2d_g2s(smem + S::kOffA + i * S::kABytes, &tmap_a, (kb0 + i) * kBK, mt * 128, &full[i]);
      uint32_t st = 0, ph = 1;  // stage ring continues after the `pre` primed stages
      for (int i = pre; i < kb_per; ++i) {
        mbar_wait(&empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&full[st], S::kABytes + S::kBBytes);
        tma_tensor_2d_g2s(smem + S::kOffA + st * S::kABytes, &tmap_a, (kb0 + i) * kBK, mt * 128, &full[st]);
        tma_bulk_g2s(smem + S::kOffB + st * S::kBBytes, b_src + (size_t)(kb0 + i) * b_kb_stride, S::kBBytes, &full[st]);
        if (++st == S::kStages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // elect.sync, not lane == 0: lets ptxas keep descriptors in uniform registers (no per-MMA waterfall loop)
      constexpr uint32_t idesc = make_idesc(kFmtTF32, 128, kNS);
      const uint32_t a_base = smem_u32(smem + S::kOffA);
      const uint32_t b_base = smem_u32(smem + S::kOffB);
      uint32_t st = 0, ph = 0;
      for (int kb = 0; kb < kb_per; ++kb) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) {
          const uint64_t a_desc = make_smem_desc(a_base + st * S::kABytes + k * 32, 0, 1024, kLayoutSW128);
          const uint64_t b_desc = make_smem_desc(b_base + st * S::kBBytes + k * 32, 0, 1024, kLayoutSW128);
          umma_tf32_ss(tmem_base, a_desc, b_desc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[st]);
        if (++st == S::kStages) { st = 0; ph ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else {
    const uint32_t inbox_local = smem_u32(smem + S::kOffInbox);
    {  // push this CTA's partial rows to their owners
      const int quad = warp & 3;
      const int r = quad * 32 + lane;  // row of the tile = TMEM lane
      const int owner = r / RO, lr = r % RO;
      const bool valid = (long long)mt * 128 + r < m_rows;  // rows beyond M are never pushed (nor summed)
      const bool warp_has_rows = (long long)mt * 128 + quad * 32 < m_rows;  // warp-uniform
      uint32_t dst, dst_bar;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                   : "=r"(dst)
                   : "r"(inbox_local + (uint32_t)((crank * RO + lr) * kNS * 4)), "r"(owner));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst_bar) : "r"(smem_u32(inbox_full)), "r"(owner));
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
      if (warp_has_rows) {
#pragma unroll
        for (int c0 = 0; c0 < kNS; c0 += 32) {
          float v[32];
          tmem_ld_x16(taddr + c0, v);  // warp-collective: all 32 lanes, valid row or not
          tmem_ld_x16(taddr + c0 + 16, v + 16);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t pos = (uint32_t)((c0 / 4 + j) ^ (lr & 15));  // 16-byte chunk, swizzled by row
              asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + pos * 16), "f"(v[4 * j]),
                           "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                           : "memory");
            }
          }
        }
      }
      tc_fence_before();
      asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(dst_bar) : "memory");
    }
    mbar_wait_cluster(inbox_full, 0);  // all CS partials of this CTA's rows have landed
    if (residual != nullptr) pdl_wait();  // the residual may have been written by the previous kernel in the stream
    const int t = threadIdx.x - 64;  // 0..127
    constexpr int kChunks = kNS / 4;  // 16-byte chunks per row
    for (int u = t; u < RO * kChunks; u += 128) {
      const int lr = u / kChunks, ch = u % kChunks;
      const long long m = (long long)mt * 128 + crank * RO + lr;
      if (m >= m_rows) continue;
      const uint32_t pos = (uint32_t)(ch ^ (lr & 15));
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < CS; ++p) {
        const float4 v = ld_shared_v4(inbox_local + (uint32_t)(((p * RO + lr) * kChunks + pos) * 16));
        acc.x += v.x;
        acc.y += v.y;
        acc.z += v.z;
        acc.w += v.w;
      }
      const int col = ns * kNS + ch * 4;
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale + col));
      const float4 h4 = __ldg(reinterpret_cast<const float4*>(shift + col));
      acc.x = fmaf(s4.x, acc.x, h4.x);
      acc.y = fmaf(s4.y, acc.y, h4.y);
      acc.z = fmaf(s4.z, acc.z, h4.z);
      acc.w = fmaf(s4.w, acc.w, h4.w);
      if (relu) {
        acc.x = fmaxf(acc.x, 0.f);
        acc.y = fmaxf(acc.y, 0.f);
        acc.z = fmaxf(acc.z, 0.f);
        acc.w = fmaxf(acc.w, 0.f);
      }
      if (residual != nullptr) {  // fused residual add (dense output only), optional ReLU on the sum
        const float4 r4 = *reinterpret_cast<const float4*>(residual + (size_t)m * Cout + col);
        acc.x += r4.x, acc.y += r4.y, acc.z += r4.z, acc.w += r4.w;
        if (relu_after) {
          acc.x = fmaxf(acc.x, 0.f);
          acc.y = fmaxf(acc.y, 0.f);
          acc.z = fmaxf(acc.z, 0.f);
          acc.w = fmaxf(acc.w, 0.f);
        }
      }
      if (!out_padded) {
        *reinterpret_cast<float4*>(y + (size_t)m * Cout + col) = acc;
      } else {
        int2 pc = one_frame_index(m, geo);
        if (geo.interior_only) pc.y = 0;
        one_frame_store(y, pc, geo, Cout, col, acc);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// Once per layer: W[Cin][Cout] (reference layout, Kernel128_one.cu:41-48) -> per (n-tile, 32-channel block) the
// K-major 128-byte-swizzled shared-memory image [BN couts][32 cin], RN-rounded to TF32. The reference's cuDNN half
// does the same [Cin][Cout] -> [Cout][Cin] transpose on the host (util.c:15-26, Kernel128_one.cu:131).
__global__ void weight_pack_kernel(const float* __restrict__ w, float* __restrict__ w_img, int Cin, int Cout, int BN,
                                   int op16) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cin * Cout) return;
  const int co = idx % Cout;
  const int ci = idx / Cout;
  const int nt = co / BN, r = co % BN;
  const int kb = ci / kBK, kk = ci % kBK;
  if (op16) {
    // bf16 image: [Cout/BN][Cin/32][4 k-chunks of 8 channels][BN couts][8 bf16] (K-major, no swizzle; RN)
    const size_t off = (((size_t)nt * (Cin / kBK) + kb) * 4 + (kk >> 3)) * (size_t)(BN * 8) + (size_t)r * 8 + (kk & 7);
    reinterpret_cast<__nv_bfloat16*>(w_img)[off] = __float2bfloat16_rn(w[idx]);
    return;
  }
  const int chunk = (kk >> 2) ^ (r & 7);
  const size_t off = ((size_t)nt * (Cin / kBK) + kb) * (size_t)(BN * kBK) + (size_t)r * kBK + chunk * 4 + (kk & 3);
  w_img[off] = to_tf32_rn(w[idx]);
}

static int encode_2d(CUtensorMap* tmap, const float* base, int inner, long long rows, int box_inner, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return WG_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)inner * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, wg::l2_promotion(),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WG_OK : WG_ERR_TMAP;
}

int one_make_tmap(CUtensorMap* tmap, const float* x, long long m_rows, int Cin) {
  return encode_2d(tmap, x, Cin, m_rows, kBK, 128);
}

int one_make_tmap_out(CUtensorMap* tmap, const float* y, long long m_rows, int Cout) {
  return encode_2d(tmap, y, Cout, m_rows, 32, 32);
}

// developer build only: WG_ONE_ABLATE=<bits> (see the kernel); 0 in the product build
static int one_dev_flags() {
  static int v = -1;
  if (v < 0) {
    const char* a = dev_env("WG_ONE_ABLATE");
    v = a ? atoi(a) : 0;
  }
  return v;
}

// developer build: buffer for the timeline stamps (WG_ONE_ABLATE & 16), set by tools/one_timeline.py through
// wg_dev_set_dbg_ptr(); never dereferenced otherwise
static float* g_one_dbg = nullptr;
static float* one_dbg_ptr() { return g_one_dbg; }
#ifdef WG_DEV_BUILD
extern "C" void wg_dev_set_dbg_ptr(void* p) { g_one_dbg = static_cast<float*>(p); }
#endif

struct OneRes {  // residual operand of a launch (RES instantiations)
  const CUtensorMap* tmap_r;
  int relu_after;
};

template <int BN, int CL, bool WS = false, bool PAIR = false, bool RES = false, bool H16 = false>
static int launch_one(const CUtensorMap& tmap, const CUtensorMap& tmap_y, const float* w_img, const float* scale,
                      const float* shift, float* y_padded, long long m_rows, int Cin, int Cout, int relu,
                      int max_ctas, cudaStream_t stream, const OneGeo& geo, int bn_packed = BN,
                      OneRes res = OneRes{nullptr, 0}) {
  using S = OneSmem<BN, WS, H16>;
  static unsigned long long configured = 0;  // per device: the attribute is a property of the function on ONE device
  int dev_ = 0;
  cudaGetDevice(&dev_);
  const unsigned long long dev_bit_ = 1ull << (dev_ & 63);
  if (!(configured & dev_bit_)) {
    cudaError_t e = cudaFuncSetAttribute(conv1x1_bn_act_kernel<BN, CL, WS, PAIR, RES, H16>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kTotal);
    if (e != cudaSuccess) return WG_ERR_CUDA;
    configured |= dev_bit_;
  }
  const long long n_mtiles = (m_rows + 127) / 128;
  const long long n_items = ((n_mtiles + CL - 1) / CL) * (Cout / BN);
  long long n_clusters = max_ctas / CL;
  if (n_clusters > n_items) n_clusters = n_items;
  if (n_clusters < 1) n_clusters = 1;
  if (WS) {  // every CTA owns one N-tile: grid = a multiple of the N-tile count
    const long long n_nt = Cout / BN;
    n_clusters = (max_ctas / n_nt) * n_nt;
    if (n_clusters > n_mtiles * n_nt) n_clusters = n_mtiles * n_nt;
    if (n_clusters < n_nt) n_clusters = n_nt;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_clusters * CL));
  cfg.blockDim = dim3(H16 ? kOneThreads + 128 : kOneThreads);
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv1x1_bn_act_kernel<BN, CL, WS, PAIR, RES, H16>, tmap, tmap_y,
                                     res.tmap_r ? *res.tmap_r : tmap_y, w_img, scale, shift, y_padded, m_rows, Cin, Cout,
                                     relu, bn_packed, res.relu_after, one_dev_flags(), geo, one_dbg_ptr());
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

// how many CS-clusters of the small kernel can be resident at once on this device (cached per device and CS)
template <int CS>
static int small_max_clusters() {
  static int cached[64];
  int dev_ = 0;
  cudaGetDevice(&dev_);
  int& slot = cached[dev_ & 63];
  if (slot == 0) {
    cudaFuncSetAttribute(conv1x1_small_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)SmallSmem::kTotal);
    if (CS > 8) cudaFuncSetAttribute(conv1x1_small_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS);
    cfg.blockDim = dim3(kOneThreads);
    cfg.dynamicSmemBytes = SmallSmem::kTotal;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CS > 1 ? 1 : 0;
    int n = 0;
    if (CS == 1) {
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev_);
    } else if (cudaOccupancyMaxActiveClusters(&n, conv1x1_small_kernel<CS>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    slot = n > 0 ? n : -1;
  }
  return slot > 0 ? slot : 0;
}

template <int CS>
static int launch_small(const CUtensorMap& tmap, const float* w_img, const float* scale, const float* shift, float* y,
                        int out_padded, long long m_rows, int Cin, int Cout, int BN, int relu, const float* residual,
                        int relu_after, const OneGeo& geo, cudaStream_t stream) {
  const long long n_items = ((m_rows + 127) / 128) * (Cout / kNS);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_items * CS));
  cfg.blockDim = dim3(kOneThreads);
  cfg.dynamicSmemBytes = SmallSmem::kTotal;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv1x1_small_kernel<CS>, tmap, w_img, scale, shift, y, m_rows, Cin, Cout,
                                     BN, relu, out_padded, residual, relu_after, geo);
  return e == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

// bf16 operands: one kernel for every batch size (128-wide N-tiles, A converted into TMEM by four extra warps)
static int one_bf16_launch(const CUtensorMap& tmap, const CUtensorMap& tmap_y, const CUtensorMap& tmap_res,
                           const float* w_img, const float* scale, const float* shift, float* y_padded, long long m_rows,
                           int Cin, int Cout, int relu, const float* residual, int relu_after, int max_ctas,
                           const OneGeo& geo, cudaStream_t stream) {
  if (residual)
    return launch_one<128, 1, false, false, true, true>(tmap, tmap_y, w_img, scale, shift, y_padded, m_rows, Cin, Cout,
                                                        relu, max_ctas, stream, geo, 128, OneRes{&tmap_res, relu_after});
  return launch_one<128, 1, false, false, false, true>(tmap, tmap_y, w_img, scale, shift, y_padded, m_rows, Cin, Cout,
                                                       relu, max_ctas, stream, geo, 128);
}

int one_launch(const CUtensorMap& tmap, const CUtensorMap& tmap_y, const CUtensorMap& tmap_res, const float* w_img,
               const float* scale, const float* shift, float* y, int out_padded, long long m_rows, int Cin, int Cout,
               int BN, int bf16, int relu, const float* residual, int relu_after, int max_ctas, const OneGeo& geo_in,
               cudaStream_t stream) {
  OneGeo geo = geo_in;
  geo.interior_only = (out_padded & 2) ? 1 : 0;  // out_padded: bit 0 = padded frame, bit 1 = interior only
  float* y_padded = (out_padded & 1) ? y : nullptr;
  out_padded &= 1;
  if (bf16)  // bf16 operands: the H16 instantiation of the throughput kernel, every batch size
    return one_bf16_launch(tmap, tmap_y, tmap_res, w_img, scale, shift, y_padded, m_rows, Cin, Cout, relu, residual,
                           relu_after, max_ctas, geo, stream);
  {
    // latency mode: every (M-tile, 64-cout sub-tile) fits on the chip at once -> the small kernel, split-K factor CS
    // chosen by a two-term model in clocks (weight/activation ingest of Cin/CS channels at ~36 B/clk per SM, DSMEM reduction
    // of (CS-1)/CS of a 32 KB partial at ~18 B/clk + fixed cost). WG_ONE_SPLITK=1 disables, WG_ONE_CS=n forces CS.
    static int sk_env = -1, cs_env = 0;
    if (sk_env < 0) {
      const char* e = dev_env("WG_ONE_SPLITK");
      sk_env = e ? atoi(e) : 0;
      const char* c = dev_env("WG_ONE_CS");
      cs_env = c ? atoi(c) : 0;
    }
    const int n_kb = Cin / kBK;
    const long long n_items = ((m_rows + 127) / 128) * (Cout / kNS);
    if (sk_env != 1 && Cout % kNS == 0 && n_items <= max_ctas) {
      int best = 0;
      double best_t = 1e30;
      for (int cs : {1, 2, 4, 8, 16}) {
        if (n_kb % cs != 0 || n_items * cs > max_ctas) continue;
        const int fit = cs == 1 ? small_max_clusters<1>() : cs == 2 ? small_max_clusters<2>()
                      : cs == 4 ? small_max_clusters<4>() : cs == 8 ? small_max_clusters<8>()
                                                                    : small_max_clusters<16>();
        if (n_items > fit) continue;
        const double t = (n_kb / cs) * 24576.0 / 36.0 + (cs > 1 ? (cs - 1.0) / cs * 32768.0 / 18.0 + 400.0 : 0.0);
        if (cs_env == cs || (cs_env == 0 && t < best_t)) {
          best = cs;
          best_t = cs_env == cs ? -1.0 : t;
        }
      }
#define WG_SMALL(CS_) \
  return launch_small<CS_>(tmap, w_img, scale, shift, y, out_padded, m_rows, Cin, Cout, BN, relu, residual, relu_after, \
                           geo, stream)
      if (best == 1) WG_SMALL(1);
      if (best == 2) WG_SMALL(2);
      if (best == 4) WG_SMALL(4);
      if (best == 8) WG_SMALL(8);
      if (best == 16) WG_SMALL(16);
#undef WG_SMALL
    }
  }
  // WG_ONE_CLUSTER=1|2|4 selects the cluster size. Default 1: measured on B200 (profiles/README.md) the multicast
  // variants are no faster (2) or slower (4) -- the limiter is per-SM ingest / shared-memory bandwidth, which
  // multicast does not reduce, not L2 output bandwidth.
  static int cl = -1;
  if (cl < 0) {
    const char* e = dev_env("WG_ONE_CLUSTER");
    cl = e ? atoi(e) : 1;
    if (cl != 1 && cl != 2 && cl != 4) cl = 1;
  }
  const int use = (m_rows <= 128) ? 1 : cl;  // a single M-tile has nobody to share the weight tile with
  {
    // weight-stationary schedule when the packed [BN x Cin] tile fits 128 KB and every CTA gets several M-tiles of one
    // N-tile. Measured at N=256: 128->512 31.9 -> 29.6 us. Narrower stationary tiles (a 128-row slice of a 256-row
    // packed tile, WG_ONE_WS_BN=128) make Cin=256 eligible but lose (256->1024: 64 -> 79 us: A is then re-read per
    // N-tile through the tensor-map path, ~37 B/clk per SM), so they are not used.
    static int ws_env = -1;  // WG_ONE_WS=0 disables
    if (ws_env < 0) {
      const char* e = dev_env("WG_ONE_WS");
      ws_env = e ? atoi(e) : 1;
    }
    static int ws_bn_env = -1;  // WG_ONE_WS_BN=128: also allow 128-row stationary slices (experiments)
    if (ws_bn_env < 0) {
      const char* e = dev_env("WG_ONE_WS_BN");
      ws_bn_env = e ? atoi(e) : 0;
    }
    int bn_ws = (long long)Cin * BN * 4 <= 128 * 1024 ? BN : 0;
    if (ws_bn_env == 128 && (long long)Cin * 128 * 4 <= 128 * 1024) bn_ws = 128;
    const long long n_mt = (m_rows + 127) / 128, n_nt = bn_ws ? Cout / bn_ws : 0;
    if (ws_env && use == 1 && bn_ws && max_ctas >= n_nt && n_mt >= 4 * (max_ctas / n_nt)) {
      if (bn_ws == 128 && residual)
        return launch_one<128, 1, true, false, true>(tmap, tmap_y, w_img, scale, shift, y_padded, m_rows, Cin, Cout, relu,
                                                     max_ctas, stream, geo, BN, OneRes{&tmap_res, relu_after});
      if (bn_ws == 128)
        return launch_one<128, 1, true>(tmap, tmap_y, w_img, scale, shift, y_padded, m_rows, Cin, Cout, relu, max_ctas,
                                        stream, geo, BN);
      if (residual)
        return launch_one<256, 1, true, false, true>(tmap, tmap_y, w_img, scale, shift, y_padded, m_rows, Cin, Cout, relu,
                                                     max_ctas, stream, geo, BN, OneRes{&tmap_res, relu_after});
      return launch_one<256, 1, true>(tmap, tmap_y, w_img, scale, shift, y_padded, m_rows, Cin, Cout, relu, max_ctas,
                                      stream, geo, BN);
    }
  }
  if (residual) {
    if (BN == 128)
      return launch_one<128, 1, false, false, true>(tmap, tmap_y, w_img, scale, shift, y_padded, m_rows, Cin, Cout, relu,
                                                    max_ctas, stream, geo, 128, OneRes{&tmap_res, relu_after});
    return launch_one<256, 1, false, false, true>(tmap, tmap_y, w_img, scale, shift, y_padded, m_rows, Cin, Cout, relu,
                                                  max_ctas, stream, geo, 256, OneRes{&tmap_res, relu_after});
  }
#define WG_ONE(BN_, CL_) \
  return launch_one<BN_, CL_>(tmap, tmap_y, w_img, scale, shift, y_padded, m_rows, Cin, Cout, relu, max_ctas, stream, geo)
  // CTA pairs (tcgen05 cta_group::2, M = 256 over two consecutive M-tiles, each CTA holding half of the weight tile):
  // WG_ONE_PAIR=1; experiment, default off (see the kernel).
  static int pair = -1;
  if (pair < 0) {
    const char* e = dev_env("WG_ONE_PAIR");
    pair = e ? (atoi(e) != 0) : 0;
  }
  if constexpr (kDev) {  // experiments that lost (see the kernel's header): CTA pairs, weight-tile multicast clusters
    if (pair && m_rows > 128 && max_ctas >= 2) {
      if (BN == 128)
        return launch_one<128, 2, false, true>(tmap, tmap_y, w_img, scale, shift, y_padded, m_rows, Cin, Cout, relu,
                                               max_ctas, stream, geo);
      if (BN == 256)
        return launch_one<256, 2, false, true>(tmap, tmap_y, w_img, scale, shift, y_padded, m_rows, Cin, Cout, relu,
                                               max_ctas, stream, geo);
    }
    if (BN == 128 && use == 2) WG_ONE(128, 2);
    if (BN == 128 && use == 4) WG_ONE(128, 4);
    if (BN == 256 && use == 2) WG_ONE(256, 2);
    if (BN == 256 && use == 4) WG_ONE(256, 4);
  }
  if (BN == 128) WG_ONE(128, 1);
  if (BN == 256) WG_ONE(256, 1);
#undef WG_ONE
  return WG_ERR_ARG;
}

int weight_pack_launch(const float* w_cin_cout, float* w_img, int Cin, int Cout, int BN, int op16,
                       cudaStream_t stream) {
  const int n = Cin * Cout;
  weight_pack_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w_cin_cout, w_img, Cin, Cout, BN, op16);
  return cudaGetLastError() == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

}  // namespace wg
