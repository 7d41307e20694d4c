/* tests/cpp/abi_c99.c — the C ABI header must be usable from plain C (C99): compile-and-link check plus the entry
 * points that need no device (error channel, functional tables, leaf partition). */
#include "../../include/occ_xc_b200.h"
#include <stdio.h>
#include <string.h>

int main(void) {
  if (occ_xc_abi_version() != OCC_XC_B200_ABI_VERSION) return 1;
  if (occ_xc_family(OCC_XC_GGA_X_PBE) != 1 || occ_xc_family(OCC_XC_MGGA_C_TPSS) != 2 || occ_xc_family(424242) != -1) return 2;
  if (occ_xc_exact_exchange(OCC_XC_HYB_GGA_XC_B3LYP) != 0.20) return 3;
  if (occ_xc_decay_cutoff(NULL, 1e-10, NULL) == 0) return 4;
  if (strstr(occ_xc_last_error(), "null") == NULL) return 5;
  /* one s shell at the origin, two leaves: partition over 2 ranks is a pure host computation */
  int l = 0, nprim = 1, off = 0, atom = 0;
  double origin[3] = {0, 0, 0}, e = 1.0, c = 1.0;
  occ_xc_basis_t b;
  b.nsh = 1; b.pure = 0; b.l = &l; b.origin = origin; b.nprim = &nprim; b.prim_off = &off; b.exps = &e; b.coefs = &c;
  b.shell_atom = &atom;
  occ_xc_leaf_t leaves[2] = {{0, 128, {0, 0, 0}, 1.0}, {128, 128, {0.5, 0, 0}, 1.0}};
  int rank[2], nb[2];
  if (occ_xc_partition_leaves(&b, 2, leaves, NULL, 1e-10, 2, rank, nb)) return 6;
  if (nb[0] != 1 || nb[1] != 1 || rank[0] != 0 || rank[1] != 1) return 7;
  printf("ok\n");
  return 0;
}
