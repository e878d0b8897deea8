// Internal glue between the C-ABI (wg_api.cu) and the kernels. Not installed; the public surface is include/*.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/winograd_b200.h"

namespace wg {

// Product build: no environment variable changes what the library computes or which kernel it launches. The developer
// build (make dev -> tools/libwinograd_b200_dev.so, -DWG_DEV_BUILD) keeps the A/B knobs (WG_FF_*, WG_ONE_*, WG_WINO_*),
// the ablation instantiations and the superseded kernel generations that the profiles under profiles/ were taken with.
#ifdef WG_DEV_BUILD
constexpr bool kDev = true;
inline const char* dev_env(const char* name) { return getenv(name); }
#else
constexpr bool kDev = false;
inline const char* dev_env(const char*) { return nullptr; }
#endif

namespace ff {
// Spatial geometry of a 3x3 layer (runtime; the reference hard-codes 14x14 maps in 16x16 frames,
// Kernel128_winograd.cu:26-31,263-265). Output H x W, tiles of 2x2 outputs: TX x TY per image; input frame Hf x Wf =
// 2*TY+2 x 2*TX+2 (= H+2 x W+2 for even sizes; odd sizes carry one extra, ignored row / column so that frame rows pair
// up -- the raw tile is fetched as four (y parity, x parity) planes). Plane = [rp_box row pairs][SP slots][8 ch], slot =
// x/2 + 1 (slot 0 = the zero-filled column x/2 = -1), SP = TX + 2. rp_box * SP <= 216 slots (kPlaneBytes).
struct Geo {
  int H, W, Hf, Wf, TX, TY, TT, SP, RPI;  // TT = TX*TY tiles per image, RPI = Hf/2 frame row pairs per image
  int rp_box;                              // row pairs per TMA box (covers every M-block of up to mv_max tiles)
  int mv_max;                              // largest M-block (tiles) whose raw rows fit one plane
  uint32_t raw_bytes;                      // 4 * rp_box * SP * 32: bytes one raw stage delivers
};
__host__ __device__ inline bool geo_is_ref(const Geo& g) { return g.H == 14 && g.W == 14; }

}  // namespace ff

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled fetched through the runtime (no link-time libcuda dependency); nullptr on failure.
PFN_encodeTiled get_encode_tiled();

// L2 promotion of the activation tensor maps (WG_L2_PROMO=0|64|128|256 overrides; experiments)
CUtensorMapL2promotion l2_promotion();

// programmatic dependent launch on every product kernel (WG_PDL=0 disables; A/B measurements)
bool pdl_enabled();

// ---- 3x3 Winograd path (winograd_kernels.cu)
int wino_make_tmap(CUtensorMap* tmap, const float* x, int n_img, int C);
int wino_launch(const CUtensorMap& tmap, const void* u_img, const float* scale, const float* shift, float* y,
                int n_img, int C, int K, int KN, int bf16, int relu, int out_padded, int max_ctas,
                cudaStream_t stream);
int filter_transform_launch(const float* w_kcrs, void* u_img, int C, int K, int KN, int bf16, cudaStream_t stream);

// previous generation, developer build only (wino_tm_kernel.cu): V in tensor memory, half fold, cout slices of 48 / 32
#ifdef WG_DEV_BUILD
int wino_tm_cls(int K, int db);  // cluster size (CTAs sharing the raw tiles of one M-block)
int wino_tm_make_tmap(CUtensorMap* tmap, const float* x, int n_img, int C, int cls);
int wino_tm_choose_db(int C, int K);  // 1 = double-buffered V / 32-wide slices, 0 = one V stage / 48-wide slices
int filter_transform_tm_launch(const float* w_kcrs, float* u_img, int C, int K, int db, int op16,
                               cudaStream_t stream);  // op16: 0 tf32, 1 bf16, 2 fp16
int wino_tm_launch(const CUtensorMap& tmap, const float* u_img, const float* scale, const float* shift, float* y,
                   int n_img, int C, int K, int db, int op16, int relu, int out_padded, int max_ctas,
                   cudaStream_t stream);
#endif

// throughput kernel with V in tensor memory and the whole inverse transform folded into the MMAs (wino_ff_kernel.cu):
// 4 accumulators x 96-wide cout slices, its own filter image; same tensor map as the TM kernel (cls = 1)
int wino_ff_p9();  // raw-tile layout of the full-fold kernel (1 = parity planes with a 9-slot pitch)
int wino_ff_geo(int H, int W, ff::Geo* geo);  // WG_ERR_ARG when the map cannot be tiled (H, W < 3, or no M-block fits)
int wino_ff_make_tmap(CUtensorMap* tmap, const float* x, int n_img, int C, const ff::Geo& geo);
int wino_ff_cg2();  // 1 = CTA pairs (cta_group::2): filter image split in cout halves, clusters of 2
int wino_ff_has_narrow(int K);  // 1 = the layer also gets a filter image with all slices 64 wide (one-wave launches)
int filter_transform_ff_launch(const float* w_kcrs, float* u_img, int C, int K, int op16, int cg2, int narrow,
                               cudaStream_t stream);
int wino_ff_launch(const CUtensorMap& tmap, const float* x, const float* u_img, const float* u_img_narrow,
                   const float* scale, const float* shift, float* y, int n_img, int C, int K, int op16, int cg2, int relu,
                   int out_padded, int max_ctas, const ff::Geo& geo, cudaStream_t stream);

// the same kernel with sixteen transform warps, one group of 8 per V half (wino_ffw_kernel.cu); same filter image / map
int wino_ffw_launch(const CUtensorMap& tmap, const float* u_img, const float* scale, const float* shift, float* y,
                    int n_img, int C, int K, int op16, int cg2, int split, int narrow, int relu, int out_padded,
                    int mv, int grid, cudaStream_t stream);

// small-batch latency variant (wino_small_kernel.cu): TF32 only, filter in the plain KN=32 image
int wino_small_make_tmap(CUtensorMap* tmap, const float* x, int n_img, int C);
int wino_small_cs(int n_img, int C, int K, int max_ctas);  // cluster split factor, 0 = use the persistent kernel
int wino_small_launch(const CUtensorMap& tmap_small, const float* u_plain, const float* scale, const float* shift,
                      float* y, int n_img, int C, int K, int relu, int out_padded, int cs, cudaStream_t stream);

// ---- 3x3 as a direct convolution on the tensor core (conv3x3_direct_kernel.cu): TF32, the reference's 14x14 map in its
// 16x16 frame, Cin % 32 == 0, Cout % 128 == 0. Weight image: [Cout/128][Cin/32][9 taps][128 couts][32 channels].
// op16: 0 = TF32, 1 = bf16, 2 = fp16 ([K/128][C/64][9][128][64] 16-bit; needs Cin % 64 == 0)
int direct_pack_launch(const float* w_kcrs, float* w_img, int Cin, int Cout, int op16, cudaStream_t stream);
int direct16_launch(const CUtensorMap& tmap_x, const CUtensorMap& tmap_y, const float* w_img, const float* scale,
                    const float* shift, int n_img, int Cin, int Cout, int op16, int relu, int out_padded, int max_ctas,
                    int mixed, cudaStream_t stream);
int direct_make_tmap_in(CUtensorMap* tmap, const float* x, int n_img, int Cin);
int direct_make_tmap_out(CUtensorMap* tmap, const float* y, int n_img, int Cout, int out_padded);
// cl = cluster size (weight multicast), out_padded: 0 dense, 1 frame with its zero border; mixed = half-image tail items
int direct_launch(const CUtensorMap& tmap_x, const CUtensorMap& tmap_y, const float* w_img, const float* scale,
                  const float* shift, int n_img, int Cin, int Cout, int cl, int relu, int out_padded, int max_ctas,
                  int mixed, cudaStream_t stream);

// other map sizes (TF32): geometry as launch parameters, plain row stores (same file)
struct DirGeo {
  int H, W, Hf, Wf;
  int R, bands, G;
  int n_pad, halo, box_rows, n_boxes;
};
constexpr int kDirectGenMaxRows = 384, kDirect16GenMaxRows = 304;  // rows an activation stage holds (TF32 / 16-bit)
bool direct_gen_geo(int H, int W, int Hf, int Wf, int max_rows, DirGeo* out);
int direct16_gen_launch(const CUtensorMap& tmap_x, const float* w_img, const float* scale, const float* shift, float* y,
                        int n_img, int Cin, int Cout, int op16, int relu, int out_padded, int max_ctas, const DirGeo& g,
                        cudaStream_t stream);
int direct_gen_make_tmap_in(CUtensorMap* tmap, const float* x, int n_img, int Cin, const DirGeo& g);
int direct_gen_launch(const CUtensorMap& tmap_x, const float* w_img, const float* scale, const float* shift, float* y,
                      int n_img, int Cin, int Cout, int relu, int out_padded, int max_ctas, const DirGeo& g,
                      cudaStream_t stream);

// ---- 1x1, wide-Cout shapes at throughput sizes (conv1x1_t_kernel.cu): couts on M, pixels on N, resident weight slab
bool onet_eligible(long long m_rows, int Cin, int Cout, int max_ctas);
bool onet_pair(int Cout);
int onet_make_tmap_in(CUtensorMap* tmap, const float* x, long long m_rows, int Cin, int Cout);
int onet_make_tmap_out(CUtensorMap* tmap, const float* y, long long m_rows, int Cout);
int onet_launch(const CUtensorMap& tmap_x, const CUtensorMap& tmap_y, const CUtensorMap* tmap_r, const float* w_img,
                const float* scale, const float* shift, long long m_rows, int Cin, int Cout, int relu, int bn_packed,
                int relu_after, int max_ctas, cudaStream_t stream);

// frame output (WG_OUT_PADDED, 14x14 maps) of the transposed kernel: items of 16 image rows, one TMA store per frame row
bool onetf_eligible(int n_img, int Cin, int Cout, int max_ctas);
int onetf_make_tmap_in(CUtensorMap* tmap, const float* x, long long m_rows, int Cin, int Cout);
int onetf_make_tmap_out(CUtensorMap* tmap, const float* y_frame, int n_img, int Cout);
int onetf_launch(const CUtensorMap& tmap_x, const CUtensorMap& tmap_y, const float* w_img, const float* scale,
                 const float* shift, int n_img, int Cin, int Cout, int relu, int bn_packed, int interior_only,
                 int max_ctas, cudaStream_t stream);

// ---- 1x1 GEMM path (one_kernels.cu)
int one_make_tmap(CUtensorMap* tmap, const float* x, long long m_rows, int Cin);
int one_make_tmap_out(CUtensorMap* tmap, const float* y, long long m_rows, int Cout);
// spatial geometry of a 1x1 layer: H x W pixels per image; Hf x Wf = the padded frame written in chain mode
struct OneGeo {
  int H, W, Hf, Wf;
  int interior_only;  // chain mode: skip the border zeros (WG_OUT_INTERIOR_ONLY)
};
// tmap_res / residual / relu_after: fused residual add in the epilogue (null = none); bf16: bf16-operand kernel
int one_launch(const CUtensorMap& tmap, const CUtensorMap& tmap_y, const CUtensorMap& tmap_res, const float* w_img,
               const float* scale, const float* shift, float* y, int out_padded, long long m_rows, int Cin, int Cout,
               int BN, int bf16, int relu, const float* residual, int relu_after, int max_ctas, const OneGeo& geo,
               cudaStream_t stream);
int weight_pack_launch(const float* w_cin_cout, float* w_img, int Cin, int Cout, int BN, int op16, cudaStream_t stream);

}  // namespace wg
