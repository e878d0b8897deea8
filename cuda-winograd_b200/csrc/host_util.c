/* Host helpers exported with the reference's names and signatures (util.h:19-25 there) so that its Test.c and any
 * caller of these four symbols links against libwinograd_b200.so unchanged. Behaviour mirrors util.c:5-63 of the
 * reference: raw little-endian float32 files without header, exit(0) on a missing file, a checker that only prints. */
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/util.h"

uint64_t getTimeMicroseconds64(void) {
  struct timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  return (uint64_t)ts.tv_sec * 1000000ull + (uint64_t)(ts.tv_nsec / 1000);
}

/* weight is an [w][h] row-major matrix; returns the [h][w] transpose in fresh storage and frees the argument. */
float* transpose(float* weight, int h, int w) {
  float* t = (float*)malloc((size_t)w * h * sizeof(float));
  if (!t) {
    printf("Bad Malloc\n");
    exit(0);
  }
  for (int r = 0; r < w; ++r)
    for (int c = 0; c < h; ++c) t[(size_t)c * w + r] = weight[(size_t)r * h + c];
  free(weight);
  return t;
}

float* get_parameter(const char* filename, int size) {
  float* buf = (float*)malloc((size_t)size * sizeof(float));
  if (!buf) {
    printf("Bad Malloc\n");
    exit(0);
  }
  FILE* f = fopen(filename, "rb");
  if (!f) {
    printf("Bad file path: %p, %s\n", (void*)f, strerror(errno));
    exit(0);
  }
  size_t got = fread(buf, sizeof(float), (size_t)size, f);
  if (got != (size_t)size) memset(buf + got, 0, ((size_t)size - got) * sizeof(float));
  fclose(f);
  return buf;
}

/* A: [len+2*shift][len+2*shift][channel] frame with the result at (+shift,+shift); B: dense [len][len][channel].
 * Counts |a-b| > 1e-5 and prints "[max_error: %f][error_cnt: %d]" like the reference; also returns max_error. */
float output_checker(float* A, float* B, int len, int channel, int shift) {
  int error_cnt = 0;
  float max_error = 0.f;
  const int pitch = len + 2 * shift;
  for (int i = 0; i < len; ++i)
    for (int j = 0; j < len; ++j) {
      const float* a = A + ((size_t)(i + shift) * pitch + (j + shift)) * channel;
      const float* b = B + ((size_t)i * len + j) * channel;
      for (int k = 0; k < channel; ++k) {
        const float diff = fabsf(a[k] - b[k]);
        if (diff > 1e-5f) ++error_cnt;
        if (diff > max_error) max_error = diff;
      }
    }
  printf("[max_error: %f][error_cnt: %d]\n", max_error, error_cnt);
  return max_error;
}
