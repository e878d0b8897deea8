// Shared by the full-fold 3x3 kernels (wino_ff_kernel.cu, wino_ffw_kernel.cu): shared-memory / TMEM layout constants,
// the cout-slice rule of the filter image, operand rounding and the MMA wrapper. Internal.
#pragma once
#include "ptx.cuh"
#include "wg_internal.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace wg {

namespace ff {
constexpr int kWorkerWarps = 8, kProducerWarp = 8, kMmaWarp = 9;
constexpr int kRawRows = 48;                           // input rows (n*16+y) one 128-tile M-block can touch
constexpr uint32_t kRawBytes = kRawRows * 2 * 8 * 32;  // [ny][x parity][x/2][8 ch] fp32 = 24576
// P9 raw layout (conflict-free patch loads): four planes (y parity, x parity), each [24 row pairs][9 slots][8 ch] fp32
// with slot = x/2 + 1 (slot 0 = the out-of-bounds column x/2 = -1, zero-filled, never read). The row pitch of 9 slots
// puts the 32-byte slot of tile (ty, tx) at 9*(ty + dy/2) + tx + dx/2 + 1: with the tiles of an M-block numbered
// right-to-left inside each tile row, consecutive tiles sit in consecutive slots modulo 8 -- also across the end of a
// tile row -- so the 8 lanes of a quarter warp (16 bytes each, half selected by the 32-byte swizzle) hit 8 different
// bank groups. (Only a quarter warp that straddles two IMAGES still pays a second wavefront.)
constexpr uint32_t kPlaneBytes = 24 * 9 * 32;          // 6912 = 27 * 256
constexpr uint32_t kRawBytesP9 = 4 * kPlaneBytes;      // 27648
constexpr uint32_t kRawStride = kRawBytesP9;           // stage pitch of both layouts
constexpr int kRawStages = 3, kUBufs = 4;
constexpr int kKNmax = 96;
constexpr uint32_t kAccStride = 96;                    // TMEM: accumulator (a,b) at column (2a+b)*96 ...
constexpr uint32_t kVCol0 = 4 * kAccStride;            // ... V half jh at 384 + 64*jh, point (i, jj) at +8*(2i+jj)
constexpr uint32_t kUChunkMax = 8 * 2 * kKNmax * 16;   // 8 points x [2 k-chunks][KN couts][16 B] = 24576
constexpr int kEW = 32;                                // couts per epilogue chunk
constexpr uint32_t kStgRow = 2 * kEW * 4 + 16;         // [2 px][32 couts] fp32 per tile, rows padded by 16 B
constexpr uint32_t kStgBytes = 128 * kStgRow;
constexpr uint32_t kOffRaw = 0;
constexpr uint32_t kOffU = kOffRaw + kRawStages * kRawStride;
constexpr uint32_t kOffStg = kOffU + kUBufs * kUChunkMax;
constexpr uint32_t kOffPix = kOffStg + kStgBytes;      // first output pixel of each tile row (int[128]) + in-map mask (int[128])
constexpr uint32_t kOffBar = kOffPix + 2 * 128 * 4;
constexpr uint32_t kNumBars = 2 * kRawStages + 8 + 2 + 4 + 2;
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kTotal = kOffTmemPtr + 16;
static_assert(kOffU % 1024 == 0 && kOffStg % 128 == 0 && kOffBar % 8 == 0, "alignment");
static_assert(kTotal <= 227 * 1024, "shared memory budget");

// cout slices: ceil(K/96) of them, widths in multiples of 32 as even as possible, wider ones first
// (256 = 96 + 96 + 64, 128 = 64 + 64, 512 = 4 x 96 + 2 x 64). narrow: K/64 slices of 64 (K % 64 == 0) -- the second
// filter image of a layer, for launches whose items do not fill the SMs: a 64-wide slice leaves room for two V stages in
// TMEM and needs only two thirds of a 96-wide slice's MMA time per stage.
__host__ __device__ inline int n_slices(int K, int narrow = 0) { return narrow ? K / 64 : (K + 95) / 96; }
struct Slice { int kn, c0; };  // width and first cout
__host__ __device__ inline Slice slice(int K, int s, int narrow = 0) {
  if (narrow) return Slice{64, 64 * s};
  const int ns = n_slices(K), units = K / 32, base = units / ns, rem = units % ns;
  return Slice{32 * (base + (s < rem ? 1 : 0)), 32 * (s * base + (s < rem ? s : rem))};
}
__host__ __device__ inline int slice_of(int K, int k, int narrow = 0) {
  if (narrow) return k / 64;
  const int ns = n_slices(K), units = K / 32, base = units / ns, rem = units % ns;
  const int wide = rem * (base + 1) * 32;
  return k < wide ? k / ((base + 1) * 32) : rem + (k - wide) / (base * 32);
}
// a layer gets a narrow image when that slicing differs from the default one
__host__ __device__ inline bool has_narrow(int K) { return K % 64 == 0 && K / 64 != n_slices(K); }
}  // namespace ff

__device__ __forceinline__ float ff_tf32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

// Two packed fp32 in one 64-bit register (low word = first value): the transform warps' column / row passes run on
// Blackwell's 2-wide fp32 pipe (add.f32x2 / fma.f32x2), i.e. half the issue slots of scalar FADDs.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
  f2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2_t f2_sub(f2_t a, f2_t b) {  // a - b = fma(b, -1, a), exact
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(b), "l"(0xBF800000BF800000ull), "l"(a));
  return d;
}
// round both halves to TF32 (nearest, ties away): +0x1000 on each word; the MMA ignores the low 13 mantissa bits. A
// carry from the low into the high word would need a low word >= 0xFFFFF000 (a negative NaN payload): not a number the
// transform produces from finite inputs.
__device__ __forceinline__ f2_t f2_tf32(f2_t v) { return v + 0x0000100000001000ull; }
__device__ __forceinline__ float f2_lo(f2_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float f2_hi(f2_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ void ld_shared_f2x2(uint32_t addr, f2_t& a, f2_t& b) {  // 16 bytes = 4 fp32 = two pairs
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}

template <bool H16, bool CG2>
__device__ __forceinline__ void ff_umma(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  if constexpr (CG2) {
    if constexpr (H16) umma_f16_ts_cg2(d_tmem, a_tmem, b_desc, idesc, accumulate);
    else umma_tf32_ts_cg2(d_tmem, a_tmem, b_desc, idesc, accumulate);
  } else {
    if constexpr (H16) umma_f16_ts(d_tmem, a_tmem, b_desc, idesc, accumulate);
    else umma_tf32_ts(d_tmem, a_tmem, b_desc, idesc, accumulate);
  }
}

// two fp32 -> one 32-bit TMEM column of 16-bit operands (first value in the low half), round to nearest
__device__ __forceinline__ float ff_pack16(float lo, float hi, int fp16) {
  if (fp16) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
  }
  const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return __uint_as_float(*reinterpret_cast<const uint32_t*>(&b));
}

}  // namespace wg
