/* Plain-C host of libdiaglib_b200.so: the reference's own test problem (main.f90:311-317,
 * a(i,i) = i + 1, a(i,j) = 1/(i+j), n = 1000, 10 roots, n_max = 15, tol 1e-8) solved with
 * lobpcg_driver and davidson_driver through the C ABI exactly as a Fortran host would call it
 * (every scalar by reference, same argument order as diaglib.f90:171-172 / 1483-1484).
 *
 *   gcc -O2 -Iinclude examples/c_host.c -o c_host -Ldiaglib_b200 -ldiaglib_b200 -Wl,-rpath,$PWD/diaglib_b200
 *
 * Exit status: 0 = both drivers converged, 3 = no CUDA device (there is no CPU fallback). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "diaglib_b200.h"

int main(void) {
  const int32_t n = 1000, n_targ = 10, n_max = 15, max_iter = 100, max_dav = 20, verbose = 0, gen_eig = 0;
  const double tol = 1e-8, shift = 0.0;
  if (diaglib_b200_init(0) != DIAGLIB_B200_OK) {
    printf("c_host: %s\n", diaglib_b200_last_message());
    return 3;
  }
  /* dense toy matrix stored as CSR (n entries per row) + its diagonal for the preconditioner */
  int64_t* rowptr = malloc((n + 1) * sizeof *rowptr);
  int32_t* col = malloc((size_t)n * n * sizeof *col);
  double* val = malloc((size_t)n * n * sizeof *val);
  double* diag = malloc(n * sizeof *diag);
  for (int i = 0; i < n; ++i) {
    rowptr[i] = (int64_t)i * n;
    for (int j = 0; j < n; ++j) {
      col[(size_t)i * n + j] = j;
      val[(size_t)i * n + j] = i == j ? i + 2.0 : 1.0 / (i + j + 2.0);   /* 1-based: a(i,i) = i + 1 */
    }
    diag[i] = i + 2.0;
  }
  rowptr[n] = (int64_t)n * n;
  if (diaglib_b200_set_csr(n, 0, rowptr, col, val, diag) != DIAGLIB_B200_OK) return 4;

  double* evec = malloc((size_t)n * n_max * sizeof *evec);
  double eig[15];
  int rc = 0;
  for (int drv = 0; drv < 2; ++drv) {
    /* start vectors: any deterministic non-orthogonal block; check_guess orthonormalises it */
    uint64_t s = 88172645463325252ull;
    for (size_t k = 0; k < (size_t)n * n_max; ++k) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      evec[k] = (double)(s >> 11) / 9007199254740992.0 - 0.5;
    }
    int32_t ok = 0;
    if (drv == 0)
      diaglib_b200_lobpcg_driver(&verbose, &gen_eig, &n, &n_targ, &n_max, &max_iter, &tol, &shift, diaglib_b200_csr_matvec,
                                 diaglib_b200_diag_precnd, NULL, eig, evec, &ok);
    else
      diaglib_b200_davidson_driver(&verbose, &n, &n_targ, &n_max, &max_iter, &tol, &max_dav, &shift, diaglib_b200_csr_matvec,
                                   diaglib_b200_diag_precnd, eig, evec, &ok);
    printf("%s ok=%d status=%d eig:", drv == 0 ? "lobpcg" : "davidson", ok, diaglib_b200_last_status());
    for (int i = 0; i < n_targ; ++i) printf(" %.12f", eig[i]);
    printf("\n");
    if (!ok || diaglib_b200_last_status() != 0) rc = 1;
  }
  diaglib_b200_finalize();
  free(rowptr); free(col); free(val); free(diag); free(evec);
  return rc;
}
