// TEST INFRASTRUCTURE, NOT PRODUCT. One translation unit per reference source file (REF_UNIT = 0..3): the reference's
// .cu is #included from /root/reference WHERE IT LIES (never copied into this repo; -I/root/reference on the command
// line) and a small dump function launches its __global__ kernels exactly as its host function does, returning the
// raw output so the oracle can be checked against the reference's own arithmetic on a B200 (oracle/_ref/ref_dump).
// The reference's entry points themselves only print a max-error line (util.c:62) and keep the result private.
#include <cuda_runtime.h>

#if REF_UNIT == 0
#include "Kernel128_winograd.cu"
// launch sequence and buffer sizes: Kernel128_winograd.cu:235-265 (input buffer allocated twice as large and zeroed)
extern "C" int ref_dump_128w(const float* x, const float* u36, const float* bn_bias, const float* bn_scale,
                             float* out_padded) {
  const int C = 128, nIn = 16 * 16 * C, nOut = 16 * 16 * C, nW = 36 * C * C, nT = 16 * 36 * C;
  float *input, *output, *lw, *t_input, *ip, *lb, *ls;
  cudaMalloc(&input, nIn << 3);
  cudaMalloc(&output, nOut << 2);
  cudaMalloc(&lw, nW << 2);
  cudaMalloc(&t_input, nT << 2);
  cudaMalloc(&ip, nT << 2);
  cudaMalloc(&lb, C << 2);
  cudaMalloc(&ls, C << 2);
  cudaMemset(input, 0, nIn << 3);
  cudaMemset(output, 0, nOut << 2);
  cudaMemset(t_input, 0, nT << 2);
  cudaMemset(ip, 0, nT << 2);
  cudaMemcpy(input, x, nIn << 2, cudaMemcpyHostToDevice);
  cudaMemcpy(lw, u36, nW << 2, cudaMemcpyHostToDevice);
  cudaMemcpy(lb, bn_bias, C << 2, cudaMemcpyHostToDevice);
  cudaMemcpy(ls, bn_scale, C << 2, cudaMemcpyHostToDevice);
  kernel_128_winograd_BtdB<<<dim3(4, 4), dim3(128, 6), (6 * 6 * 128) << 2>>>(input, t_input);
  kernel_128_OuterProduct_128<<<dim3(36, 2), dim3(128, 8), (8 * 128 + 64 * 128 + 8 * 128) << 2>>>(t_input, lw, ip);
  kernel_128_winograd_AtIA<<<dim3(4, 4, 128), dim3(6, 6), (6 * 6) << 2>>>(ip, lb, ls, output);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(out_padded, output, nOut << 2, cudaMemcpyDeviceToHost);
  cudaFree(input); cudaFree(output); cudaFree(lw); cudaFree(t_input); cudaFree(ip); cudaFree(lb); cudaFree(ls);
  return (int)e;
}

#elif REF_UNIT == 1
#include "Kernel256_winograd.cu"
// Kernel256_winograd.cu:236-268
extern "C" int ref_dump_256w(const float* x, const float* u36, const float* bn_bias, const float* bn_scale,
                             float* out_padded) {
  const int C = 256, nIn = 16 * 16 * C, nOut = 16 * 16 * C, nW = 36 * C * C, nT = 16 * 36 * C;
  float *input, *output, *lw, *t_input, *ip, *lb, *ls;
  cudaMalloc(&input, nIn << 3);
  cudaMalloc(&output, nOut << 2);
  cudaMalloc(&lw, nW << 2);
  cudaMalloc(&t_input, nT << 2);
  cudaMalloc(&ip, nT << 2);
  cudaMalloc(&lb, C << 2);
  cudaMalloc(&ls, C << 2);
  cudaMemset(input, 0, nIn << 3);
  cudaMemset(output, 0, nOut << 2);
  cudaMemset(t_input, 0, nT << 2);
  cudaMemset(ip, 0, nT << 2);
  cudaMemcpy(input, x, nIn << 2, cudaMemcpyHostToDevice);
  cudaMemcpy(lw, u36, nW << 2, cudaMemcpyHostToDevice);
  cudaMemcpy(lb, bn_bias, C << 2, cudaMemcpyHostToDevice);
  cudaMemcpy(ls, bn_scale, C << 2, cudaMemcpyHostToDevice);
  kernel_256_winograd_BtdB<<<dim3(4, 4, 2), dim3(128, 6), (6 * 6 * 128) << 2>>>(input, t_input);
  kernel_256_OuterProduct_256<<<dim3(36, 2), dim3(256, 4), (8 * 256 + 32 * 256 + 8 * 256) << 2>>>(t_input, lw, ip);
  kernel_256_winograd_AtIA<<<dim3(4, 4, 256), dim3(6, 6), (6 * 6) << 2>>>(ip, lb, ls, output);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(out_padded, output, nOut << 2, cudaMemcpyDeviceToHost);
  cudaFree(input); cudaFree(output); cudaFree(lw); cudaFree(t_input); cudaFree(ip); cudaFree(lb); cudaFree(ls);
  return (int)e;
}

#elif REF_UNIT == 2
#include "Kernel128_one.cu"
static int run_one(int which, const float* x, const float* w, const float* bn_bias, const float* bn_scale, float* out,
                   int cin, int cout) {
  const int nIn = 196 * cin, nOut = 196 * cout, nW = cin * cout;
  float *in_, *out_, *w_, *b_, *s_;
  cudaMalloc(&in_, nIn << 3);  // Kernel128_one.cu:85,303 allocate twice the input
  cudaMalloc(&out_, nOut << 2);
  cudaMalloc(&w_, nW << 2);
  cudaMalloc(&b_, cout << 2);
  cudaMalloc(&s_, cout << 2);
  cudaMemcpy(in_, x, nIn << 2, cudaMemcpyHostToDevice);
  cudaMemcpy(w_, w, nW << 2, cudaMemcpyHostToDevice);
  cudaMemcpy(b_, bn_bias, cout << 2, cudaMemcpyHostToDevice);
  cudaMemcpy(s_, bn_scale, cout << 2, cudaMemcpyHostToDevice);
  if (which == 0)  // Kernel128_one.cu:98
    kernel_512_one_128<<<dim3(49), dim3(128, 4), (4 * 512 + 64 * 128 + 4 * 128 + 2 * 128) << 2>>>(in_, w_, b_, s_, out_);
  else  // Kernel128_one.cu:316
    kernel_128_one_512<<<dim3(49, 4), dim3(128, 4), (4 * 128 + 64 * 128 + 4 * 128 + 2 * 128) << 2>>>(in_, w_, b_, s_,
                                                                                                       out_);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(out, out_, nOut << 2, cudaMemcpyDeviceToHost);
  cudaFree(in_); cudaFree(out_); cudaFree(w_); cudaFree(b_); cudaFree(s_);
  return (int)e;
}
extern "C" int ref_dump_128_1_in(const float* x, const float* w, const float* b, const float* s, float* out) {
  return run_one(0, x, w, b, s, out, 512, 128);
}
extern "C" int ref_dump_128_1_out(const float* x, const float* w, const float* b, const float* s, float* out) {
  return run_one(1, x, w, b, s, out, 128, 512);
}

#elif REF_UNIT == 3
#include "Kernel256_one.cu"
static int run_one(int which, const float* x, const float* w, const float* bn_bias, const float* bn_scale, float* out,
                   int cin, int cout) {
  const int nIn = 196 * cin, nOut = 196 * cout, nW = cin * cout;
  float *in_, *out_, *w_, *b_, *s_;
  cudaMalloc(&in_, nIn << 3);
  cudaMalloc(&out_, nOut << 2);
  cudaMalloc(&w_, nW << 2);
  cudaMalloc(&b_, cout << 2);
  cudaMalloc(&s_, cout << 2);
  cudaMemcpy(in_, x, nIn << 2, cudaMemcpyHostToDevice);
  cudaMemcpy(w_, w, nW << 2, cudaMemcpyHostToDevice);
  cudaMemcpy(b_, bn_bias, cout << 2, cudaMemcpyHostToDevice);
  cudaMemcpy(s_, bn_scale, cout << 2, cudaMemcpyHostToDevice);
  if (which == 0)  // Kernel256_one.cu:100
    kernel_1024_one_256<<<dim3(49), dim3(256, 4), (4 * 1024 + 16 * 256 + 4 * 256 + 2 * 256) << 2>>>(in_, w_, b_, s_,
                                                                                                     out_);
  else  // Kernel256_one.cu:318
    kernel_256_one_1024<<<dim3(49, 4), dim3(256, 4), (4 * 256 + 32 * 256 + 4 * 256 + 2 * 256) << 2>>>(in_, w_, b_, s_,
                                                                                                       out_);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(out, out_, nOut << 2, cudaMemcpyDeviceToHost);
  cudaFree(in_); cudaFree(out_); cudaFree(w_); cudaFree(b_); cudaFree(s_);
  return (int)e;
}
extern "C" int ref_dump_256_1_in(const float* x, const float* w, const float* b, const float* s, float* out) {
  return run_one(0, x, w, b, s, out, 1024, 256);
}
extern "C" int ref_dump_256_1_out(const float* x, const float* w, const float* b, const float* s, float* out) {
  return run_one(1, x, w, b, s, out, 256, 1024);
}
#else
#error "REF_UNIT must be 0..3"
#endif
