/* Kernel256_winograd.h -- compatibility shim: same file name as the reference's header so that `#include "Kernel256_winograd.h"` in its
 * Test.c keeps compiling. The declarations live in wg_legacy.h; the data/<name>.bin path table the reference keeps in
 * this header (Kernel256_winograd.h:8-18 there) lives in cuda-winograd_b200/csrc/legacy_entry.cu. */
#ifndef WG_COMPAT_KERNEL256_WINOGRAD_H_
#define WG_COMPAT_KERNEL256_WINOGRAD_H_
#include "wg_legacy.h"
#endif
