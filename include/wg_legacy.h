/* wg_legacy.h -- the reference's six zero-argument entry points (its whole public API), re-implemented on top of
 * winograd_b200.h. Each call: reads the reference's data/<name>.bin files from the current directory, runs the fused
 * layer once on device 0 (N = 1), prints the reference's stdout lines and returns the packed int
 *     (my_elapsed_us << 16) | (baseline_elapsed_us & 0xFFFF)         (Kernel128_winograd.cu:433 there).
 * The product links no cuDNN, so baseline_elapsed_us is 0 unless a baseline hook is installed (wg_set_baseline_hook);
 * the "[max_error][error_cnt]" line compares against data/golden_<case>.bin (written by oracle/datagen.py) when
 * that file exists, or against the hook's output.
 *
 *   symbol             reference declaration          ./Test mode   layer
 *   kernel_128         Kernel128_winograd.h:20        0             3x3 128->128 + BN + ReLU, 16x16 padded frame out
 *   kernel_256         Kernel256_winograd.h:20        1             3x3 256->256 + BN + ReLU
 *   kernel_128_1_in    Kernel128_one.h:18             2             1x1 512->128 + BN + ReLU
 *   kernel_128_1_out   Kernel128_one.h:19             3             1x1 128->512 + BN
 *   kernel_256_1_in    Kernel256_one.h:18             4             1x1 1024->256 + BN + ReLU
 *   kernel_256_1_out   Kernel256_one.h:19             5             1x1 256->1024 + BN
 */
#ifndef WG_LEGACY_H_
#define WG_LEGACY_H_

#ifdef __cplusplus
extern "C" {
#endif

int kernel_128(void);
int kernel_256(void);
int kernel_128_1_in(void);
int kernel_128_1_out(void);
int kernel_256_1_in(void);
int kernel_256_1_out(void);

/* Optional same-process baseline (e.g. a cuDNN harness living OUTSIDE this library). Called after the product run
 * with host tensors in the reference's layouts; must fill y_baseline (dense [196][cout]) and return its elapsed
 * microseconds (or a negative value for "not available"). */
typedef int (*wg_baseline_fn)(int mode, int cin, int cout, int relu, const float* x_host, const float* w_host,
                              const float* scale, const float* shift, float* y_baseline);
void wg_set_baseline_hook(wg_baseline_fn fn);

/* Copy of the last product output (dense [196][cout], border stripped) for tests; returns element count. */
int wg_legacy_last_output(float* dst, int max_elems);

#ifdef __cplusplus
}
#endif
#endif
