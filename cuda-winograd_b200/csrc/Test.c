/* ./Test <mode 0..5> -- the reference's command-line harness (Test.c:13-56 there) over libwinograd_b200.so:
 * 100 calls of the selected entry point, the first two discarded, packed return split into mine / baseline
 * microseconds, same output lines. WG_TEST_ITERS overrides the iteration count (>= 3). */
#include <stdio.h>
#include <stdlib.h>

#include "Kernel128_one.h"
#include "Kernel128_winograd.h"
#include "Kernel256_one.h"
#include "Kernel256_winograd.h"
#include "util.h"

typedef int (*entry_fn)(void);

int main(int argc, char** argv) {
  static const entry_fn entries[6] = {kernel_128,       kernel_256,      kernel_128_1_in,
                                      kernel_128_1_out, kernel_256_1_in, kernel_256_1_out};
  int iters = 100;
  const char* env = getenv("WG_TEST_ITERS");
  if (env && atoi(env) >= 3) iters = atoi(env);

  const int mode = argc == 2 ? atoi(argv[1]) : 0;
  long mine = 0, base = 0;
  for (int i = 0; i < iters; ++i) {
    printf("---- Iter: %d ----\n", i);
    const int res = (mode >= 0 && mode < 6) ? entries[mode]() : -1;
    if (i > 1) {
      mine += res >> 16;
      base += res & 0xFFFF;
    }
  }
  printf("Average Total Time: [Mine: %d us], [cuDNN: %d us]\n", (int)(mine / (iters - 2)), (int)(base / (iters - 2)));
  return 0;
}
