/* util.h -- the four host helpers of the reference (util.h:19-25 there), exported by libwinograd_b200.so with the
 * same names and signatures so that the reference's Test.c / callers build against this library unchanged. */
#ifndef WG_COMPAT_UTIL_H_
#define WG_COMPAT_UTIL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* malloc + fread of `size` float32 values; prints "Bad file path" and exit(0) if the file is missing (util.c:28-44).
 * A file shorter than `size` values is not an error here either (the reference ignores fread's return value and leaves
 * the tail uninitialised): the missing tail reads as zeros. No status code crosses this interface. */
float* get_parameter(const char* filename, int size);
/* [w][h] -> [h][w]; frees its argument (util.c:15-26). */
float* transpose(float* weight, int h, int w);
/* wall clock in microseconds, CLOCK_REALTIME (util.c:5-13). */
uint64_t getTimeMicroseconds64(void);
/* prints "[max_error: %f][error_cnt: %d]" for |A-B| > 1e-5; A is a (len+2*shift)^2 frame, B dense (util.c:46-63). */
float output_checker(float* A, float* B, int len, int channel, int shift);

#ifdef __cplusplus
}
#endif
#endif
