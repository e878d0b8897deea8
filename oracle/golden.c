/* Plain-C CPU restatement of the fused conv + inference-BatchNorm (+ReLU) layer -- TEST INFRASTRUCTURE, NOT PRODUCT.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load liboracle.so.
 *
 * Follows the cuDNN call sequence the reference treats as truth (it has no CPU implementation of the path):
 *   3x3: valid cross-correlation of x[N][16][16][C] (NHWC, Kernel128_winograd.cu:335) with w[K][C][3][3] (KCRS, :343)
 *        -> [N][14][14][K] (:339), then the folded BN the reference kernels apply, o = scale*y + shift, and ReLU
 *        (Kernel128_winograd.cu:162-163; folding: data_generator.py:41-47);
 *   1x1: y[M][Cout] = x[M][Cin] . W[Cin][Cout] (Kernel128_one.cu:46-48), scale/shift, ReLU only when relu != 0
 *        (Kernel128_one.cu:53 vs :272).
 * Accumulation is double, results are rounded once to float. Pinned against oracle/golden.py (itself pinned against an
 * index-for-index emulation of the reference kernels) in tests/test_oracle.py. OpenMP over output pixels when built
 * with -fopenmp; oracle_threads() reports how many threads a call uses.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* y: dense [N][14][14][K] when out_padded == 0, else the reference's zero-bordered [N][16][16][K] frame
 * (Kernel128_winograd.cu:163,243). */
void oracle_conv3x3_bn_relu(const float* x, const float* w, const float* scale, const float* shift, float* y, int N,
                            int C, int K, int relu, int out_padded) {
  const int W = out_padded ? 16 : 14, off = out_padded ? 1 : 0;
  /* repack the filter to [r][s][c][k] so the inner loop runs over contiguous k */
  double* wt = (double*)malloc((size_t)9 * C * K * sizeof(double));
  for (int k = 0; k < K; ++k)
    for (int c = 0; c < C; ++c)
      for (int rs = 0; rs < 9; ++rs) wt[((size_t)rs * C + c) * K + k] = (double)w[((size_t)k * C + c) * 9 + rs];
  if (out_padded)
    for (size_t i = 0; i < (size_t)N * 256 * K; ++i) y[i] = 0.0f;
#pragma omp parallel
  {
    double* acc = (double*)malloc((size_t)K * sizeof(double));
#pragma omp for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
      for (int p = 0; p < 196; ++p) {
        const int oy = p / 14, ox = p % 14;
        for (int k = 0; k < K; ++k) acc[k] = 0.0;
        for (int r = 0; r < 3; ++r)
          for (int s = 0; s < 3; ++s) {
            const float* xp = x + (((size_t)n * 16 + oy + r) * 16 + ox + s) * C;
            const double* wp = wt + (size_t)(r * 3 + s) * C * K;
            for (int c = 0; c < C; ++c) {
              const double xv = (double)xp[c];
              const double* wk = wp + (size_t)c * K;
              for (int k = 0; k < K; ++k) acc[k] += xv * wk[k];
            }
          }
        float* yp = y + (((size_t)n * W + oy + off) * W + ox + off) * K;
        for (int k = 0; k < K; ++k) {
          double o = (double)scale[k] * acc[k] + (double)shift[k];
          if (relu && o < 0.0) o = 0.0;
          yp[k] = (float)o;
        }
      }
    free(acc);
  }
  free(wt);
}

void oracle_conv1x1_bn(const float* x, const float* w, const float* scale, const float* shift, float* y, long long M,
                       int Cin, int Cout, int relu) {
#pragma omp parallel
  {
    double* acc = (double*)malloc((size_t)Cout * sizeof(double));
#pragma omp for schedule(static)
    for (long long m = 0; m < M; ++m) {
      for (int k = 0; k < Cout; ++k) acc[k] = 0.0;
      const float* xp = x + (size_t)m * Cin;
      for (int c = 0; c < Cin; ++c) {
        const double xv = (double)xp[c];
        const float* wk = w + (size_t)c * Cout;
        for (int k = 0; k < Cout; ++k) acc[k] += xv * (double)wk[k];
      }
      float* yp = y + (size_t)m * Cout;
      for (int k = 0; k < Cout; ++k) {
        double o = (double)scale[k] * acc[k] + (double)shift[k];
        if (relu && o < 0.0) o = 0.0;
        yp[k] = (float)o;
      }
    }
    free(acc);
  }
}

/* data_generator.py:41-47 in float arithmetic, same operation order as the numpy expressions. */
void oracle_fold_bn(int K, const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                    float* scale_out, float* shift_out) {
  for (int k = 0; k < K; ++k) {
    const float sd = sqrtf(var[k] + eps);
    scale_out[k] = gamma[k] / sd;
    shift_out[k] = beta[k] - (gamma[k] * mean[k]) / sd;
  }
}

/* util.c:46-63 semantics; returns error_cnt and stores max_error. */
int oracle_output_checker(const float* A, const float* B, int len, int channel, int shift, float* max_error_out) {
  int cnt = 0;
  float mx = 0.f;
  const int pitch = len + 2 * shift;
  for (int i = 0; i < len; ++i)
    for (int j = 0; j < len; ++j)
      for (int k = 0; k < channel; ++k) {
        const float d = fabsf(A[((size_t)(i + shift) * pitch + j + shift) * channel + k] -
                              B[((size_t)i * len + j) * channel + k]);
        if (d > 1e-5f) ++cnt;
        if (d > mx) mx = d;
      }
  if (max_error_out) *max_error_out = mx;
  return cnt;
}
