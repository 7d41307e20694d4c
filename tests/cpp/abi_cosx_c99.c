/* tests/cpp/abi_cosx_c99.c — include/occ_cosx_b200.h must be usable from plain C (C99): compile-and-link check plus
 * the argument checks that need no device. */
#include "../../include/occ_cosx_b200.h"
#include <stdio.h>
#include <string.h>

int main(void) {
  occ_cosx_handle *h = NULL;
  occ_cosx_options_t opt;
  occ_cosx_info_t info;
  memset(&opt, 0, sizeof(opt));
  opt.margin = 1.0;
  opt.device = -1;
  if (occ_cosx_create(NULL, 0, NULL, NULL, NULL, NULL, &opt, &h) == 0) return 1;
  if (strstr(occ_xc_last_error(), "null") == NULL) return 2;
  if (occ_cosx_compute_K(NULL, 0, NULL, 1e-16, 0.0, NULL) == 0) return 3;
  if (occ_cosx_compute_K_mo(NULL, 0, NULL, NULL, 0, 0, 1e-16, 0.0, NULL) == 0) return 3;
  if (occ_cosx_get_info(NULL, &info) == 0) return 4;
  occ_cosx_destroy(NULL);
  printf("ok\n");
  return 0;
}
